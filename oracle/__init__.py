"""CPU oracle loader -- TEST INFRASTRUCTURE, NOT PRODUCT.

Builds oracle/pcgrl_oracle.c with gcc into oracle/_build/ and exposes it through ctypes on numpy
arrays.  Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs may import this package (the product package ``gym_pcgrl_b200`` never does).
"""
import ctypes as C
import os
import subprocess

import numpy as np

from gym_pcgrl_b200 import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "pcgrl_oracle.c")
_SRC_SMB = os.path.join(_HERE, "smb_oracle.c")   # the smb problem's restatement, linked into the same library
_LIB = os.path.join(_HERE, "_build", "libpcgrl_oracle.so")
_lib = None


def build(force=False):
    hdr = os.path.join(_HERE, "..", "include", "pcgrl_b200.h")
    if not force and os.path.exists(_LIB) and os.path.getmtime(_LIB) >= max(os.path.getmtime(_SRC), os.path.getmtime(_SRC_SMB), os.path.getmtime(hdr)):
        return _LIB
    os.makedirs(os.path.dirname(_LIB), exist_ok=True)
    subprocess.check_call(["gcc", "-O2", "-fopenmp", "-shared", "-fPIC", "-o", _LIB, _SRC, _SRC_SMB, "-lm"])
    return _LIB


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.oracle_reset.restype = C.c_int
        _lib.oracle_step.restype = C.c_int
        _lib.oracle_get_stats.restype = C.c_int
        _lib.oracle_rng_randint.restype = C.c_int
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def alloc_buffers(cfg, n):
    """numpy state arrays for n envs + the pcgrl_buffers struct pointing at them."""
    h, w = cfg.height, cfg.width
    arrs = {}
    for name, dtype, shape in _abi.BUFFER_SPECS:
        if name == "heatmap" and (cfg.flags & _abi.FLAG_HEAT_U16):
            dtype = "uint16"
        arrs[name] = np.zeros((n,) + shape(h, w), dtype=dtype)
    arrs["tile_prob"][:] = np.asarray(list(cfg.tile_prob))[None, :]
    arrs["status"] = np.zeros(4, np.int32)
    b = _abi.PcgrlBuffers()
    for name, _, _ in _abi.BUFFER_SPECS:
        setattr(b, name, arrs[name].ctypes.data)
    b.scratch, b.scratch_bytes, b.status = None, 0, arrs["status"].ctypes.data
    return arrs, b


class OracleEnv:
    """n lock-step environments stepped by the C oracle (host memory)."""

    def __init__(self, cfg, n, threads=1):
        self.cfg, self.n, self.threads = cfg, n, threads
        self.arrs, self.bufs = alloc_buffers(cfg, n)
        self.adim = _abi.action_dim(cfg.representation)

    def seed(self, seeds):
        seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
        lib().oracle_seed(C.byref(self.bufs), _ptr(seeds), C.c_int(self.n))

    def set_rng_states(self, states):
        """states: uint32 [n, 2, 625] (or [n, 625] used for both streams)."""
        states = np.asarray(states, dtype=np.uint32)
        if states.ndim == 2:
            states = np.repeat(states[:, None, :], 2, axis=1)
        self.arrs["rng"][:] = states

    def reset(self, mask=None):
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        rc = lib().oracle_reset(C.byref(self.cfg), C.byref(self.bufs), None if m is None else _ptr(m),
                                C.c_int(self.n), C.c_int(self.threads))
        assert rc == 0, "oracle_reset failed"

    def step(self, actions):
        a = np.ascontiguousarray(actions, dtype=np.int32).reshape(self.n, self.adim)
        rc = lib().oracle_step(C.byref(self.cfg), C.byref(self.bufs), _ptr(a), C.c_int(self.n), C.c_int(self.threads))
        assert rc == 0, "oracle_step failed"

    def __getitem__(self, name):
        return self.arrs[name]


def get_stats(cfg, maps, threads=1):
    maps = np.ascontiguousarray(maps, dtype=np.uint8)
    n = maps.shape[0]
    out = np.zeros((n, _abi.MAX_STATS), np.int32)
    rc = lib().oracle_get_stats(C.byref(cfg), _ptr(maps), _ptr(out), C.c_int(n), C.c_int(threads))
    assert rc == 0, "oracle_get_stats failed"
    return out


def rng_doubles(seed, n):
    st = np.zeros(_abi.MT_WORDS, np.uint32)
    lib().oracle_rng_seed(_ptr(st), C.c_uint32(seed))
    out = np.zeros(n, np.float64)
    lib().oracle_rng_doubles(_ptr(st), _ptr(out), C.c_int(n))
    return out, st


def rng_randints(st, ns):
    return np.asarray([lib().oracle_rng_randint(_ptr(st), C.c_int(int(n))) for n in ns], np.int64)


def solver_counters():
    it, calls = C.c_long(0), C.c_long(0)
    lib().oracle_solver_counters(C.byref(it), C.byref(calls))
    return it.value, calls.value

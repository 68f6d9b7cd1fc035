"""smb oracle loader -- TEST INFRASTRUCTURE, NOT PRODUCT (SURVEY.md 8f row f3 groundwork).

Builds oracle/smb_oracle.c with gcc into oracle/_build/ and exposes SMBProblem.get_stats / get_reward /
get_episode_over of the reference on numpy arrays.  The product package does not implement smb yet; only tests/
import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "smb_oracle.c")
_LIB = os.path.join(_HERE, "_build", "libsmb_oracle.so")
_lib = None

STAT_NAMES = ["dist-floor", "disjoint-tubes", "enemies", "empty", "noise", "jumps", "jumps-dist", "dist-win"]
NSTATS = 8


def build(force=False):
    if not force and os.path.exists(_LIB) and os.path.getmtime(_LIB) >= os.path.getmtime(_SRC):
        return _LIB
    os.makedirs(os.path.dirname(_LIB), exist_ok=True)
    subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", _LIB, _SRC, "-lm"])
    return _LIB


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.smb_get_reward.restype = C.c_double
        _lib.smb_episode_over.restype = C.c_int
        _lib.smb_solver_iterations.restype = C.c_long
    return _lib


def get_stats(maps, power=10000):
    """maps uint8 [N,H,W] -> int32 [N,8] (smb_prob.py:126-148)."""
    maps = np.ascontiguousarray(maps, dtype=np.uint8)
    n, h, w = maps.shape
    out = np.zeros((n, NSTATS), dtype=np.int32)
    lib().smb_get_stats_batch(maps.ctypes.data_as(C.c_void_p), n, w, h, int(power), out.ctypes.data_as(C.c_void_p))
    return out


def get_reward(new, old, weights, iparam):
    new, old = np.ascontiguousarray(new, np.int32), np.ascontiguousarray(old, np.int32)
    weights, iparam = np.ascontiguousarray(weights, np.float64), np.ascontiguousarray(iparam, np.int32)
    return float(lib().smb_get_reward(new.ctypes.data_as(C.c_void_p), old.ctypes.data_as(C.c_void_p),
                                      weights.ctypes.data_as(C.c_void_p), iparam.ctypes.data_as(C.c_void_p)))


def episode_over(stats):
    stats = np.ascontiguousarray(stats, np.int32)
    return bool(lib().smb_episode_over(stats.ctypes.data_as(C.c_void_p)))

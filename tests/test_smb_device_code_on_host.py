"""The product's smb device code (csrc/pcgrl_smb.cuh: scalar __host__ __device__ functions, one thread per map on the
GPU) compiled for the HOST with g++ and checked against the reference's golden vectors and the smb oracle -- the
parity check of row f3's first piece that does not need a GPU.  (The -m gpu test runs the same functions through
pcgrl_smb_get_stats on the device.)"""
import ctypes as C
import os
import subprocess

import numpy as np

from oracle import smb as smb_oracle

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden", "stats_smb.npz")


def _harness(tmp_path):
    so = str(tmp_path / "libsmb_host_harness.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, os.path.join(HERE, "smb_host_harness.cpp")])
    return C.CDLL(so)


def _run(lib, maps, power):
    maps = np.ascontiguousarray(maps, dtype=np.uint8)
    n, h, w = maps.shape
    out = np.zeros((n, 12), dtype=np.int32)
    lib.smb_device_code_get_stats(maps.ctypes.data_as(C.c_void_p), n, w, h, int(power), out.ctypes.data_as(C.c_void_p), 12)
    return out


def test_smb_device_code_matches_reference_golden(tmp_path):
    lib = _harness(tmp_path)
    d = np.load(GOLDEN)
    power = int(d["solver_power"][0])
    k = 0
    while "maps_%d" % k in d.files:
        got = _run(lib, d["maps_%d" % k], power)
        np.testing.assert_array_equal(got[:, :8], d["stats_%d" % k], err_msg="group %d" % k)
        assert not got[:, 8:].any()
        k += 1
    assert k >= 6


def test_smb_device_code_matches_oracle_on_random_maps(tmp_path):
    lib = _harness(tmp_path)
    rs = np.random.RandomState(5)
    for w, h, p_solid, power in [(114, 14, 0.1, 10000), (114, 14, 0.35, 10000), (57, 9, 0.2, 300), (122, 16, 0.15, 2000), (8, 5, 0.3, 50)]:
        maps = rs.choice(7, size=(24, h, w), p=[0.9 - p_solid, p_solid] + [0.02] * 5).astype(np.uint8)
        np.testing.assert_array_equal(_run(lib, maps, power)[:, :8], smb_oracle.get_stats(maps, power), err_msg=str((w, h, p_solid, power)))

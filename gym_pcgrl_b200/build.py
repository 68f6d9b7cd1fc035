"""Build the sm_100a shared library in-tree: ``python -m gym_pcgrl_b200.build``.

One translation unit (csrc/pcgrl_b200.cu + headers) -> csrc/libpcgrl_b200.so.  nvcc cross-compiles
without a GPU; ``-fmad=false`` keeps the fp64 reward arithmetic identical to the reference's
(no product is fused into an add).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libpcgrl_b200.so")
SOURCES = ["pcgrl_b200.cu", "pcgrl_linear.cu"]
import glob  # noqa: E402


def _headers():
    """every header next to the translation unit (a stale-check list kept by hand went out of date once)"""
    return sorted(os.path.basename(p) for p in glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")))


NVCC_FLAGS = ["-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-fmad=false",
              "-shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-pthread"]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + _headers()] + [os.path.join(HERE, "..", "include", "pcgrl_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES
    subprocess.check_call(cmd, cwd=CSRC)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

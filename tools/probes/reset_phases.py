"""Build the library with -DPCGRL_PROFILE into a scratch .so and print the average cycles per reset phase."""
import ctypes as C, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
from gym_pcgrl_b200 import _native, build as B
so = os.path.join(ROOT, "gpurun_out", "libpcgrl_profile.so")
os.makedirs(os.path.dirname(so), exist_ok=True)
subprocess.check_call(["nvcc"] + B.NVCC_FLAGS + ["-DPCGRL_PROFILE", "-o", so] + B.SOURCES, cwd=B.CSRC)
_native.LIB_PATH = so
import bench
n = 4096
env = bench.make_env(n, "cuda:0", 0)
# status buffer needs room for the phase accumulators (4 int32 + 8 int64)
env._ensure_buffers()
env._tens["status"] = torch.zeros(2 * 48, dtype=torch.int32, device="cuda")
env._cbufs.status = env._tens["status"].data_ptr()
env.reset()
acts = torch.from_numpy(bench.host_actions(env, 600, n, 5)).cuda()
for t in range(600):
    env.step(acts[t])
torch.cuda.synchronize()
acc = env._tens["status"].cpu().numpy().view(np.int64)[4:]
cnt = acc[0]
names = ["stage key -> smem", "prob stream init", "thresholds (fp64)", "fill 512 draws (twist)", "cells -> map + bits", "randint+unstage", "map_stats", "probs+heat"]
print("resets", cnt)
for k, nm in enumerate(names):
    print("%-24s %8.0f cycles" % (nm, acc[k + 1] / max(cnt, 1)))
print("total              %8.0f cycles" % (acc[1:1 + len(names)].sum() / max(cnt, 1)))

print("-- single-step launch phases per warp (cycles): prologue | apply_action | map_stats | outputs+reset+epilogue")
full = env._tens["status"].cpu().numpy().view(np.int64)
for cls, nm in enumerate(["no change", "changed", "reset"]):
    a = full[16 + 8 * cls: 16 + 8 * cls + 8]
    c = max(a[0], 1)
    print("%-10s n=%-8d mean %7.0f %7.0f %7.0f %7.0f | max warp total %d, max map_stats %d" % (nm, a[0], a[1] / c, a[2] / c, a[3] / c, a[4] / c, a[5], a[6]))

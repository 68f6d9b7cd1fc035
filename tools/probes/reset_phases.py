"""Build the library with -DPCGRL_PROFILE into a scratch .so and print the average cycles per reset phase."""
import ctypes as C, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
from gym_pcgrl_b200 import _native, build as B
so = os.path.join(ROOT, "gpurun_out", "libpcgrl_profile.so")
os.makedirs(os.path.dirname(so), exist_ok=True)
subprocess.check_call(["nvcc"] + B.NVCC_FLAGS + ["-DPCGRL_PROFILE", "-o", so, "pcgrl_b200.cu"], cwd=B.CSRC)
_native.LIB_PATH = so
import bench
n = 4096
env = bench.make_env(n, "cuda:0", 0)
# status buffer needs room for the phase accumulators (4 int32 + 8 int64)
env._ensure_buffers()
env._tens["status"] = torch.zeros(4 + 2 * 10, dtype=torch.int32, device="cuda")
env._cbufs.status = env._tens["status"].data_ptr()
env.reset()
acts = torch.from_numpy(bench.host_actions(env, 600, n, 5)).cuda()
for t in range(600):
    env.step(acts[t])
torch.cuda.synchronize()
acc = env._tens["status"].cpu().numpy().view(np.int64)[4:]
cnt = acc[0]
names = ["stage", "gen+bits", "randint+unstage", "map_stats", "probs+heat"]
print("resets", cnt)
for k, nm in enumerate(names):
    print("%-18s %8.0f cycles" % (nm, acc[k + 1] / max(cnt, 1)))
print("total              %8.0f cycles" % (acc[1:6].sum() / max(cnt, 1)))

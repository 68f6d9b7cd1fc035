"""Profile build (-DPCGRL_PROFILE) of the env-asynchronous solver rollout: where do the cycles of k_rollout_async go?

  python tools/probes/async_phases.py [workload] [T]
"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
from gym_pcgrl_b200 import _native, build as B
so = os.path.join(ROOT, "gpurun_out", "libpcgrl_profile.so")
os.makedirs(os.path.dirname(so), exist_ok=True)
subprocess.check_call(["nvcc"] + B.NVCC_FLAGS + ["-DPCGRL_PROFILE", "-o", so] + B.SOURCES, cwd=B.CSRC)
_native.LIB_PATH = so
import bench
name = sys.argv[1] if len(sys.argv) > 1 else "sokoban-wide-5x5"
T = int(sys.argv[2]) if len(sys.argv) > 2 else 100
bench.select_workload(name)
n = bench.WORKLOAD["envs_per_gpu"]
env = bench.make_env(n, "cuda:0", 0)
env._ensure_buffers()
env._tens["status"] = torch.zeros(2 * 96, dtype=torch.int32, device="cuda")
env._cbufs.status = env._tens["status"].data_ptr()
env.reset()
acts = torch.from_numpy(bench.host_actions(env, 2 * T, n, 5)).cuda()
rew = torch.empty((T, n), dtype=torch.float64, device="cuda")
done = torch.empty((T, n), dtype=torch.uint8, device="cuda")
env.rollout(acts[:T], rew, done)
torch.cuda.synchronize()
env._tens["status"].zero_()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
env.rollout(acts[T:], rew, done)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
a = env._tens["status"].cpu().numpy().view(np.int64)[64:80].astype(np.float64)
ghz = 1.965
print("%s: %d envs x %d steps in %.2f ms (%.3e env-steps/s)" % (name, n, T, ms, n * T / ms * 1e3))
print("searches %d (%.2f%% of steps), mean %.1f us, max %.1f us; lock wait total %.1f ms" % (
    a[0], 100 * a[0] / (n * T), a[1] / max(a[0], 1) / ghz / 1e3, a[8] / ghz / 1e3, a[2] / ghz / 1e6))
print("resets %d, mean %.1f us" % (a[3], a[4] / max(a[3], 1) / ghz / 1e3))
print("envs %d, mean env %.1f us (%.2f us/step), max env %.1f us, max warp %.1f us" % (
    a[5], a[6] / max(a[5], 1) / ghz / 1e3, a[6] / max(a[5], 1) / ghz / 1e3 / T, a[7] / ghz / 1e3, a[9] / ghz / 1e3))
print("sum search %.1f ms, sum env %.1f ms (warp-time)" % (a[1] / ghz / 1e6, a[6] / ghz / 1e6))

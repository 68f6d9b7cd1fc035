"""pcgrl_step_host host-side phase timers (-DPCGRL_PROFILE build): H2D enqueue | kernel enqueue | D2H enqueue | sync wait | apply."""
import ctypes as C, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
from gym_pcgrl_b200 import _native, build as B, HostStepIO
so = os.path.join(ROOT, "gpurun_out", "libpcgrl_profile.so")
os.makedirs(os.path.dirname(so), exist_ok=True)
subprocess.check_call(["nvcc"] + B.NVCC_FLAGS + ["-DPCGRL_PROFILE", "-o", so] + B.SOURCES, cwd=B.CSRC)
_native.LIB_PATH = so
import bench
n, K = 4096, 1000
env = bench.make_env(n, "cuda:0", 0)
env._ensure_buffers()
env._tens["status"] = torch.zeros(2 * 48, dtype=torch.int32, device="cuda")
env._cbufs.status = env._tens["status"].data_ptr()
env.reset()
acts = torch.from_numpy(bench.host_actions(env, K + 8, n, 5)).pin_memory()
io = HostStepIO(env, with_obs=True, mode="delta")
base, stride = acts.data_ptr(), acts.stride(0) * 4
for t in range(8):
    io.struct.actions = base + t * stride; env.step_host(io)
out = (C.c_double * 8)()
_native.lib().pcgrl_debug_timers(out)
import time
t0 = time.perf_counter()
for t in range(K):
    io.struct.actions = base + (8 + t) * stride; env.step_host(io)
wall = (time.perf_counter() - t0) / K * 1e6
_native.lib().pcgrl_debug_timers(out)
c = out[7]
print("calls %d  wall %.1f us/step" % (c, wall))
for i, nm in enumerate(["H2D enqueue", "kernel enqueue", "D2H enqueue", "sync wait", "apply records"]):
    print("%-16s %6.1f us per step" % (nm, out[i] / (K + 8)))   # the counter ticks in _begin and in _end: divide by the steps

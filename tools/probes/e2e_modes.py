"""Per-step host API (pcgrl_step_host) by transport mode on the headline workload: us per step and env-steps/s."""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
B = bench.Bench(types.SimpleNamespace(), 0, 1, 0)
env = bench.make_env(n, B.dev, 0)
env.reset()
for mode in ("delta", "direct", "full", "direct", "delta"):
    ms, io, rsum = B.time_e2e(env, 20, 25, mode, 99)
    med = float(np.median(ms))
    print("%-7s %.1f us/step  %.3g env-steps/s  (reward sum %.1f)" % (mode, med / 20 * 1e3, n * 20 / (med * 1e-3), rsum), flush=True)

# device-side time of one step_host call (events around the call on the current stream): kernel (+ D2H copies in delta / full mode)
import torch
from gym_pcgrl_b200 import HostStepIO
for mode in ("delta", "direct"):
    io = HostStepIO(env, with_obs=True, with_info=False, mode=mode)
    acts = torch.from_numpy(bench.host_actions(env, 264, n, 7)).pin_memory()
    base, stride = acts.data_ptr(), acts.stride(0) * 4
    for t in range(8):
        io.struct.actions = base + t * stride
        env.step_host(io)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(256)]
    for t in range(256):
        io.struct.actions = base + (8 + t) * stride
        ev[t][0].record()
        env.step_host(io)
        ev[t][1].record()
    torch.cuda.synchronize()
    ms = np.array([a.elapsed_time(b) for a, b in ev])
    print("%-7s device time per step: median %.1f us, mean %.1f us, p90 %.1f us" % (mode, np.median(ms) * 1e3, ms.mean() * 1e3, np.percentile(ms, 90) * 1e3), flush=True)

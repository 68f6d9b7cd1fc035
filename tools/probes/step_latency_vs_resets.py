"""T = 1 step-kernel latency with and without auto-resets in the batch (binary-narrow 16x16, 4096 envs).

Right after a full reset no env can finish for ~100 steps (max_changes = 51, a third of the steps change the map), so
steps 5..60 time the kernel without any reset; steps 400..600 are the steady state with ~27 resets per step.
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import bench
n = 4096
env = bench.make_env(n, "cuda:0", 0)
env.reset()
acts = torch.from_numpy(bench.host_actions(env, 700, n, 5)).cuda()
for rep in range(2):
    env.reset()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(701)]
    dones = []
    ev[0].record()
    for t in range(700):
        _, _, d, _ = env.step(acts[t])
        ev[t + 1].record()
        dones.append(d.sum())
    torch.cuda.synchronize()
    ms = np.array([ev[t].elapsed_time(ev[t + 1]) for t in range(700)]) * 1e3
    dn = torch.stack(dones).cpu().numpy()
    print("pass %d: steps 5..60: %.1f us/step (resets/step %.1f) | steps 400..700: %.1f us/step (resets/step %.1f)" % (
        rep, np.median(ms[5:60]), dn[5:60].mean(), np.median(ms[400:]), dn[400:].mean()))
    lo = ms[400:][dn[400:] == 0]
    print("   steady-state steps without any reset: %d, median %.1f us" % (len(lo), np.median(lo) if len(lo) else float("nan")))

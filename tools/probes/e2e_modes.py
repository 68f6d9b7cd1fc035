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

#!/usr/bin/env python
"""Per-step throughput of AsyncGroupedEnv against the number of env groups (solver problems, BASELINE batch sizes).
    python tools/sweep_async_groups.py [workload=sokoban-wide-5x5-sparse] [groups=16,64,128,256] [K=48]"""
import json
import os
import sys
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

name = sys.argv[1] if len(sys.argv) > 1 else "sokoban-wide-5x5-sparse"
groups = [int(g) for g in (sys.argv[2] if len(sys.argv) > 2 else "16,64,128,256").split(",")]
K = int(sys.argv[3]) if len(sys.argv) > 3 else 48
wl = bench.WORKLOADS[name]
n = wl["envs_per_gpu"]
b = bench.Bench(types.SimpleNamespace(), 0, 1, 0)
for g in groups:
    if n % g:
        continue
    ms = b.time_e2e_async(wl, n, K, g, 1999)
    print(json.dumps({"workload": name, "envs": n, "groups": g, "envs_per_group": n // g, "steps_per_group": K,
                      "env_steps_per_s": n * K / (ms * 1e-3)}), flush=True)

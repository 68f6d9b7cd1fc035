#!/usr/bin/env python
"""Launch N single-step pcgrl_step calls (binary-narrow 16x16, 4096 envs) -- run under
`ncu --metrics gpu__time_duration.sum --cache-control none` to get the device time of a T=1 launch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
n = 4096
env = bench.make_env(n, "cuda:0", 0)
env.reset()
acts = torch.from_numpy(bench.host_actions(env, 400, n, 5)).cuda()
for t in range(400):
    env.step(acts[t])
torch.cuda.synchronize()

"""Small workloads of the fused rollouts with the incremental statistics (k_rollout<binary|zelda, narrow|turtle|wide|generic>)
for compute-sanitizer:   compute-sanitizer --tool memcheck|racecheck python tools/sanitize_r02i.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from gym_pcgrl_b200 import BatchedPcgrlEnv

rng = np.random.RandomState(0)
for prob, rep, w, h, n, T in (("binary", "narrow", 16, 16, 37, 60), ("binary", "turtle", 5, 9, 21, 80), ("binary", "wide", 32, 32, 9, 50),
                              ("binary", "narrowcast", 12, 10, 13, 40), ("zelda", "narrow", 11, 7, 33, 60), ("zelda", "wide", 11, 16, 17, 50),
                              ("zelda", "turtlecast", 9, 9, 11, 40)):
    env = BatchedPcgrlEnv(prob, rep, num_envs=n, device="cuda", seed=3)
    env.adjust_param(width=w, height=h, change_percentage=0.5)
    env.adjust_param(width=w, height=h, change_percentage=0.5)
    env.reset()
    sp = env.action_space
    if hasattr(sp, "nvec"):
        acts = np.stack([np.stack([rng.randint(int(k), size=n) for k in sp.nvec], axis=1) for _ in range(T)]).astype(np.int32)
    else:
        acts = rng.randint(sp.n, size=(T, n)).astype(np.int32)
    env.rollout(torch.from_numpy(acts).cuda())
    env.step(torch.from_numpy(acts[0]).cuda())
    env.check_status()
torch.cuda.synchronize()
print("sanitize_r02i workloads done")

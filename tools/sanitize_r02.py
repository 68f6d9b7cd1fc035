"""Small workloads of the round-2 kernels for `compute-sanitizer --tool memcheck` (k_smb_*, k_rollout_packed_binary,
k_render, k_im2col, k_linear_bf16).   compute-sanitizer --tool memcheck python tools/sanitize_r02.py"""
import os
import sys

os.environ.setdefault("PCGRL_PACKED", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from gym_pcgrl_b200 import BatchedPcgrlEnv, HostStepIO, _native
from gym_pcgrl_b200.models import ActorCritic
from gym_pcgrl_b200.policy_native import NativePolicy

# smb: reset, single steps, rollout, host step (delta transport), get_stats, render
env = BatchedPcgrlEnv("smb", "narrowcast", num_envs=37, device="cuda", seed=1)
env.adjust_param(width=40, height=9, change_percentage=0.3)
env.adjust_param(width=40, height=9, change_percentage=0.3)
env._prob._solver_power = 600
env._cfg = None
env.reset()
rng = np.random.RandomState(0)
for t in range(6):
    a = np.stack([rng.randint(3, size=37), rng.randint(7, size=37)], axis=1).astype(np.int32)
    env.step(torch.from_numpy(a).cuda())
acts = torch.from_numpy(np.stack([np.stack([rng.randint(3, size=37), rng.randint(7, size=37)], axis=1) for _ in range(12)]).astype(np.int32)).cuda()
env.rollout(acts)
io = HostStepIO(env, with_obs=True, with_info=True, mode="delta")
for t in range(5):
    io.actions.copy_(acts[t].cpu())
    env.step_host(io)
env.render("rgb_array")
_native.smb_get_stats(env._tens["map"], 300)
env.check_status()
# packed binary rollout (two and four envs per warp) incl. odd env counts
for (w, h, n) in ((16, 16, 21), (8, 8, 13)):
    e = BatchedPcgrlEnv("binary", "narrow", num_envs=n, device="cuda", seed=2)
    e.adjust_param(width=w, height=h, change_percentage=0.3)
    e.adjust_param(width=w, height=h, change_percentage=0.3)
    e.reset()
    e.rollout(torch.randint(0, 3, (96, n), device="cuda", dtype=torch.int32))
    e.render("rgb_array")
# tcgen05 linear + im2col through the policies
for kind, shape, acts_n, n in (("CustomPolicyBigMap", (28, 28, 1), 3, 70), ("FullyConvPolicySmallMap", (5, 5, 5), 125, 33)):
    net = ActorCritic(kind, shape, acts_n).cuda()
    NativePolicy(net)(torch.randint(0, 2, (n,) + shape, dtype=torch.uint8, device="cuda"))
_native.linear_bf16(torch.randn(300, 200, device="cuda"), torch.randn(132, 200, device="cuda"), torch.randn(132, device="cuda"))
torch.cuda.synchronize()
print("sanitize workloads done")

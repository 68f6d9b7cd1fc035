"""Small workloads of the round-2 kernels for `compute-sanitizer --tool memcheck` (k_smb_*, k_rollout_packed_binary,
k_render, k_im2col, k_linear_bf16, k_conv3x3_bf16, k_obs_image_u8_fast, k_rollout_async with the split open list).   compute-sanitizer --tool memcheck python tools/sanitize_r02.py"""
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
net = ActorCritic("FullyConvPolicyBigMap", (14, 14, 1), 392).cuda()        # implicit-GEMM convolutions (BLOCK_N 64 and 32)
NativePolicy(net)(torch.randint(0, 2, (19, 14, 14, 1), dtype=torch.uint8, device="cuda"))
_native.linear_bf16(torch.randn(300, 200, device="cuda"), torch.randn(132, 200, device="cuda"), torch.randn(132, device="cuda"))
# uint8 observation fast paths (raw index with a crop wider than the map; one-hot 8 channels) at odd env strides
import ctypes as C
from gym_pcgrl_b200 import PROBLEMS, REPRESENTATIONS
from gym_pcgrl_b200._config import build_config
for prob, w, h, crop in (("binary", 13, 7, 28), ("binary", 32, 32, 64), ("zelda", 11, 7, 22), ("mdungeon", 7, 11, 64)):
    p = PROBLEMS[prob]()
    p.adjust_param(width=w, height=h)
    cfg = build_config(p, REPRESENTATIONS["narrow"](), 1, 1, auto_reset=False)
    T, n = len(p.tile_types), 37
    maps = torch.randint(0, T, (n, h, w), dtype=torch.uint8, device="cuda")
    pos = torch.stack([torch.randint(0, w, (n,)), torch.randint(0, h, (n,))], dim=1).to(torch.uint8).cuda()
    out = torch.empty((n, crop, crop, T if prob != "binary" else 1), dtype=torch.uint8, device="cuda")
    rc = _native.lib().pcgrl_obs_image(C.byref(cfg), maps.data_ptr(), pos.data_ptr(), out.data_ptr(), n, crop, 1, int(prob != "binary"), 0,
                                       _native.stream_ptr(torch.device("cuda", 0)))
    assert rc == 0
# solver problems: the open list spills from shared memory to the HBM tail on long searches
for prob in ("sokoban", "mdungeon"):
    e = BatchedPcgrlEnv(prob, "wide", num_envs=200, device="cuda", seed=3)
    e.reset()
    hi = [int(v) for v in e.action_space.nvec]
    a = torch.stack([torch.randint(0, h, (40, 200), device="cuda", dtype=torch.int32) for h in hi], dim=-1).contiguous()
    e.rollout(a)
    e.check_status()
torch.cuda.synchronize()
print("sanitize workloads done")

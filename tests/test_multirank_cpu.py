"""world_size-2 gloo test (CPU) of the multi-GPU plumbing: env-index sharding gives shard-invariant
trajectories and the optional all-gather returns the global batch in global env order.  The stepping
engine here is the CPU oracle (no GPU in the build container); the sharding / seeding / gather code is
the product's (gym_pcgrl_b200.distributed, BatchedPcgrlEnv(env_offset=...))."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_PER_RANK, STEPS = 24, 40


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run_shard(env_offset, n, seed, steps):
    import oracle
    from gym_pcgrl_b200 import BatchedPcgrlEnv
    env = BatchedPcgrlEnv("zelda", "narrow", num_envs=n, seed=seed, env_offset=env_offset)
    o = oracle.OracleEnv(env.native_config, n)
    o.set_rng_states(env._pending_states)
    o.reset()
    rewards, dones = [], []
    for t in range(steps):
        # actions are a function of (global env index, t) so that every sharding sees the same ones
        gi = np.arange(env_offset, env_offset + n)
        a = ((gi * 7 + t * 13 + (gi * t) % 5) % 9).astype(np.int32)
        o.step(a)
        rewards.append(o["reward"].copy())
        dones.append(o["done"].copy())
    return np.stack(rewards), np.stack(dones), o["map"].copy()


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch
    from gym_pcgrl_b200 import distributed as D
    r, w = D.init_process_group("gloo")
    assert (r, w) == (rank, world)
    rew, done, maps = _run_shard(D.env_offset(rank, N_PER_RANK), N_PER_RANK, 3, STEPS)
    g_rew, g_done, g_map = D.all_gather_outputs(torch.from_numpy(rew[-1]), torch.from_numpy(done[-1]).bool(),
                                                torch.from_numpy(maps))
    assert g_rew.shape[0] == world * N_PER_RANK and g_done.dtype == torch.bool
    if rank == 0:
        np.savez(os.path.join(out_dir, "gathered.npz"), rew=g_rew.numpy(), done=g_done.numpy(), map=g_map.numpy())
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), rew=rew, done=done)
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


def test_two_rank_sharding_matches_single_process(tmp_path):
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    full_rew, full_done, full_map = _run_shard(0, world * N_PER_RANK, 3, STEPS)
    g = np.load(os.path.join(str(tmp_path), "gathered.npz"))
    np.testing.assert_array_equal(g["rew"], full_rew[-1])
    np.testing.assert_array_equal(g["done"], full_done[-1].astype(bool))
    np.testing.assert_array_equal(g["map"], full_map)
    for rank in range(world):
        d = np.load(os.path.join(str(tmp_path), "rank%d.npz" % rank))
        sl = slice(rank * N_PER_RANK, (rank + 1) * N_PER_RANK)
        np.testing.assert_array_equal(d["rew"], full_rew[:, sl])
        np.testing.assert_array_equal(d["done"], full_done[:, sl])


def test_single_process_gather_is_identity():
    import torch
    from gym_pcgrl_b200 import distributed as D
    t = torch.arange(6)
    assert D.all_gather_outputs(t) is t
    assert D.env_offset(3, 4096) == 12288

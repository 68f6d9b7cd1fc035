"""The product's smb code (csrc/pcgrl_smb.cuh / pcgrl_smb_env.cuh: scalar `__host__ __device__` functions; on the GPU
lane 0 of one warp per env runs them) exercised on the HOST through the host twins of the C ABI
(pcgrl_reset_cpu / pcgrl_step_cpu / pcgrl_get_stats_cpu) and checked against
  * the golden vectors recorded from the unmodified reference (stats_smb.npz: 182 maps; traj_smb_*.npz: 8 trajectories),
  * the CPU oracle on seeded random rollouts, including the exact "skip the search when no cell the last search read
    changed its solidity" carry-over.
No GPU needed; the -m gpu tests run the same functions through the kernels."""
import numpy as np
import pytest

import oracle
import util
from gym_pcgrl_b200 import PROBLEMS, PcgrlEnv, _abi, _native
from oracle import smb as smb_oracle

import os
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "stats_smb.npz")
SMB_KATS = [m for m in util.kat_configs() if m["env_id"].startswith("smb-")]


def _cpu_stats(maps, power):
    import torch
    prob = PROBLEMS["smb"]()
    prob.adjust_param(width=maps.shape[2], height=maps.shape[1])
    prob._solver_power = int(power)
    return _native.get_stats(prob, torch.from_numpy(np.ascontiguousarray(maps))).numpy()


def test_smb_scalar_code_matches_reference_golden():
    d = np.load(GOLDEN)
    power = int(d["solver_power"][0])
    k = 0
    while "maps_%d" % k in d.files:
        got = _cpu_stats(d["maps_%d" % k], power)
        np.testing.assert_array_equal(got[:, :8], d["stats_%d" % k], err_msg="group %d" % k)
        assert not got[:, 8:].any()
        k += 1
    assert k >= 6


def test_smb_scalar_code_matches_oracle_on_random_maps():
    rs = np.random.RandomState(5)
    for w, h, p_solid, power in [(114, 14, 0.1, 10000), (114, 14, 0.35, 10000), (57, 9, 0.2, 300), (122, 16, 0.15, 2000), (8, 5, 0.3, 50)]:
        maps = rs.choice(7, size=(24, h, w), p=[0.9 - p_solid, p_solid] + [0.02] * 5).astype(np.uint8)
        np.testing.assert_array_equal(_cpu_stats(maps, power)[:, :8], smb_oracle.get_stats(maps, power), err_msg=str((w, h, p_solid, power)))


@pytest.mark.parametrize("meta", SMB_KATS, ids=[m["name"] for m in SMB_KATS])
def test_smb_host_twin_trajectory_matches_reference_golden(meta):
    """The reference's own smb trajectories through the classic-gym facade on device="cpu" (manual reset on done)."""
    traj, _ = util.load_traj(meta["name"])
    prob, rep, _v = meta["env_id"].split("-")
    env = PcgrlEnv(prob, rep, device="cpu")
    if meta["kwargs"]:
        env.adjust_param(**meta["kwargs"])
        env.adjust_param(**meta["kwargs"])
    assert env._max_changes == meta["max_changes"] and env._max_iterations == meta["max_iterations"]
    env.set_rng(np.random.RandomState(meta["seed"]), np.random.RandomState(meta["seed"]))
    wide = rep == "wide"
    obs = env.reset()
    np.testing.assert_array_equal(obs["map"], traj["reset_map"][0])
    np.testing.assert_array_equal(env._batched._tens["stats"].numpy()[0, :8], traj["reset_stats"][0])
    k, total = 0, 0.0
    adim = _abi.action_dim(rep)
    for t in range(meta["steps"]):
        a = traj["actions"][t, :adim] if adim > 1 else int(traj["actions"][t, 0])
        obs, r, d, info = env.step(a)
        ctx = "%s step %d" % (meta["name"], t)
        np.testing.assert_array_equal(obs["map"], traj["map"][t], err_msg=ctx)
        np.testing.assert_array_equal(obs["heatmap"].astype(np.int32), traj["heat"][t], err_msg=ctx)
        if not wide:
            np.testing.assert_array_equal(obs["pos"].astype(np.int32), traj["pos"][t], err_msg=ctx)
        np.testing.assert_array_equal(env._batched._tens["stats"].numpy()[0, :8], traj["stats"][t], err_msg=ctx)
        assert float(r) == traj["reward"][t], ctx
        assert d == bool(traj["done"][t]), ctx
        assert info["iterations"] == traj["iteration"][t] and info["changes"] == traj["changes"][t], ctx
        assert info["dist-win"] == traj["stats"][t][7] and info["jumps"] == traj["stats"][t][5], ctx
        total += float(r)
        if d:
            obs = env.reset()
            k += 1
            np.testing.assert_array_equal(obs["map"], traj["reset_map"][k], err_msg=ctx)
            np.testing.assert_array_equal(env._batched._tens["stats"].numpy()[0, :8], traj["reset_stats"][k], err_msg=ctx)
            if not wide:
                np.testing.assert_array_equal(obs["pos"].astype(np.int32), traj["reset_pos"][k], err_msg=ctx)
    assert k == meta["episodes"] and abs(total - meta["sum_reward"]) < 1e-9


HOST_BATCH_CASES = [
    ("smb-narrow-v0", dict(width=40, height=10, change_percentage=0.3), 12, 150),
    ("smb-wide-v0", dict(width=30, height=8, change_percentage=0.5,
                         probs={"empty": 0.55, "solid": 0.3, "enemy": 0.03, "brick": 0.04, "question": 0.02, "coin": 0.02, "tube": 0.04}), 12, 150),
    ("smb-turtlecast-v0", dict(width=24, height=9, change_percentage=0.4), 8, 200),
    ("smb-narrowmulti-v0", dict(width=20, height=7, change_percentage=0.5, random_start=False), 8, 120),
    ("smb-turtle-v0", {}, 4, 60),
]


@pytest.mark.parametrize("case", HOST_BATCH_CASES, ids=[c[0] for c in HOST_BATCH_CASES])
def test_smb_host_twin_batched_matches_oracle(case):
    """Auto-reset batches on the host twin against the oracle: every buffer, every step (incl. RNG state)."""
    import torch
    env_id, kwargs, n, steps = case
    env = util.host_env(env_id, kwargs, num_envs=n, auto_reset=True, device="cpu")
    states = np.stack([util.randomstate_words(500 + i) for i in range(n)])
    env.set_rng_states(states)
    ref = oracle.OracleEnv(env.native_config, n, threads=4)
    ref.set_rng_states(states)
    wide = env_id.split("-")[1] == "wide"
    env.reset()
    ref.reset()
    arng = np.random.RandomState(6)
    sp = env.action_space
    ndone = 0
    for t in range(steps):
        if hasattr(sp, "nvec"):
            a = np.stack([arng.randint(int(k), size=n) for k in sp.nvec], axis=1).astype(np.int32)
        else:
            a = arng.randint(sp.n, size=n).astype(np.int32)
        obs, reward, done, info = env.step(torch.from_numpy(a))
        ref.step(a)
        ctx = "%s step %d" % (env_id, t)
        tn = {k: v.numpy() for k, v in env._tens.items()}
        np.testing.assert_array_equal(tn["map"], ref["map"], err_msg=ctx)
        np.testing.assert_array_equal(tn["heatmap"].view(ref["heatmap"].dtype), ref["heatmap"], err_msg=ctx)
        if not wide:
            np.testing.assert_array_equal(tn["pos"], ref["pos"], err_msg=ctx)
        for key in ("stats", "start_stats", "info_stats"):
            np.testing.assert_array_equal(tn[key][:, :8], ref[key][:, :8], err_msg=ctx + " " + key)
        np.testing.assert_array_equal(tn["info_stats"][:, 14:], ref["info_stats"][:, 14:], err_msg=ctx + " info counters")
        np.testing.assert_array_equal(tn["iteration"], ref["iteration"], err_msg=ctx)
        np.testing.assert_array_equal(tn["changes"], ref["changes"], err_msg=ctx)
        np.testing.assert_array_equal(tn["reward"], ref["reward"], err_msg=ctx)
        np.testing.assert_array_equal(tn["done"], ref["done"], err_msg=ctx)
        ndone += int(ref["done"].sum())
    np.testing.assert_array_equal(env._tens["rng"].numpy().view(np.uint32), ref["rng"])
    assert ndone > 0

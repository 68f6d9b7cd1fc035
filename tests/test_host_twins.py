"""Host twins (pcgrl_reset_cpu / pcgrl_step_cpu / pcgrl_get_stats_cpu) of binary, zelda, sokoban, ddave and mdungeon: the
bitboard algorithm of the kernels restated over row arrays on the host (csrc/pcgrl_host_twin.cuh) and, for the solver
problems, the kernels' own `__host__ __device__` game models under a plain scalar search loop
(csrc/pcgrl_solver_host.cuh), checked WITHOUT a GPU against
  * every golden trajectory of these problems recorded from the unmodified reference (incl. BASELINE config 1:
    binary-narrow 11x11, 1000 random-action steps -- SURVEY App. B.3 digest c455bfaf0f3e3b63),
  * the reference-labelled get_stats fixtures, and the oracle on random ragged sizes.
This is the "config 1 runs without a GPU and without the oracle" path: PcgrlEnv(..., device="cpu")."""
import hashlib
import struct

import numpy as np
import pytest
import torch

import oracle
import util
from gym_pcgrl_b200 import PROBLEMS, REPRESENTATIONS, PcgrlEnv, _abi, _native
from gym_pcgrl_b200._config import build_config

KATS = [m for m in util.kat_configs() if m["env_id"].split("-")[0] in ("binary", "zelda", "sokoban", "ddave", "mdungeon")]


@pytest.mark.parametrize("meta", KATS, ids=[m["name"] for m in KATS])
def test_host_twin_trajectory_matches_reference_golden(meta):
    traj, _ = util.load_traj(meta["name"])
    prob, rep, _v = meta["env_id"].split("-")
    env = PcgrlEnv(prob, rep, device="cpu")
    if meta["kwargs"]:
        env.adjust_param(**meta["kwargs"])
        env.adjust_param(**meta["kwargs"])
    env.set_rng(np.random.RandomState(meta["seed"]), np.random.RandomState(meta["seed"]))
    S, wide, adim = util.nstats(prob), rep == "wide", _abi.action_dim(rep)
    sha = hashlib.sha256()

    def feed(obs, r, d):      # the digest of SURVEY.md App. B.3
        sha.update(np.asarray(obs["map"]).astype(np.uint8).tobytes())
        if "pos" in obs:
            sha.update(np.asarray(obs["pos"]).astype(np.uint8).tobytes())
        sha.update(np.asarray(obs["heatmap"]).astype(np.int32).tobytes())
        sha.update(struct.pack("<d?", float(r), bool(d)))

    obs = env.reset()
    feed(obs, 0.0, False)
    np.testing.assert_array_equal(obs["map"], traj["reset_map"][0])
    k, total = 0, 0.0
    for t in range(meta["steps"]):
        a = traj["actions"][t, :adim] if adim > 1 else int(traj["actions"][t, 0])
        obs, r, d, info = env.step(a)
        feed(obs, r, d)
        ctx = "%s step %d" % (meta["name"], t)
        np.testing.assert_array_equal(obs["map"], traj["map"][t], err_msg=ctx)
        np.testing.assert_array_equal(obs["heatmap"].astype(np.int32), traj["heat"][t], err_msg=ctx)
        if not wide:
            np.testing.assert_array_equal(obs["pos"].astype(np.int32), traj["pos"][t], err_msg=ctx)
        np.testing.assert_array_equal(env._batched._tens["stats"].numpy()[0, :S], traj["stats"][t], err_msg=ctx)
        assert float(r) == traj["reward"][t] and d == bool(traj["done"][t]), ctx
        assert info["iterations"] == traj["iteration"][t] and info["changes"] == traj["changes"][t], ctx
        if prob == "binary":
            assert info["path-imp"] == traj["stats"][t][1] - traj["reset_stats"][k][1], ctx
        total += float(r)
        if d:
            obs = env.reset()
            feed(obs, 0.0, False)
            k += 1
            np.testing.assert_array_equal(obs["map"], traj["reset_map"][k], err_msg=ctx)
            np.testing.assert_array_equal(env._batched._tens["stats"].numpy()[0, :S], traj["reset_stats"][k], err_msg=ctx)
    assert k == meta["episodes"] and abs(total - meta["sum_reward"]) < 1e-9
    assert sha.hexdigest()[:16] == meta["digest"]


def test_baseline_config_1_on_the_cpu():
    """BASELINE.json config 1 through the public facade, no GPU, no oracle: 14 episodes, sum of rewards 14."""
    meta = [m for m in KATS if m["name"] == "binary_narrow_11x11"][0]
    assert meta["digest"] == "c455bfaf0f3e3b63" and meta["episodes"] == 14 and meta["sum_reward"] == 14


@pytest.mark.parametrize("prob_name", ["binary", "zelda"])
def test_host_twin_get_stats_matches_reference_golden_and_oracle(prob_name):
    for maps, stats in util.stats_groups(prob_name):
        prob = PROBLEMS[prob_name]()
        prob.adjust_param(width=maps.shape[2], height=maps.shape[1])
        got = _native.get_stats(prob, torch.from_numpy(maps)).numpy()
        np.testing.assert_array_equal(got[:, :stats.shape[1]], stats)
    rng = np.random.RandomState(23)
    sizes = [(1, 1), (1, 32), (32, 1), (32, 32), (2, 3), (31, 32)] + [(int(rng.randint(1, 33)), int(rng.randint(1, 33))) for _ in range(30)]
    for (w, h) in sizes:
        prob = PROBLEMS[prob_name]()
        prob.adjust_param(width=w, height=h)
        T = len(prob.tile_types)
        maps = []
        for k in range(32):
            dens = rng.random_sample()
            if prob_name == "binary":
                m = (rng.random_sample((h, w)) < dens).astype(np.uint8)
            else:
                p = np.full(T, (1 - dens) * 0.4 / (T - 2)); p[0] = dens; p[1] = (1 - dens) * 0.6
                m = rng.choice(T, size=(h, w), p=p / p.sum()).astype(np.uint8)
                if w * h >= 3 and k % 2 == 0:
                    flat = m.reshape(-1)
                    flat[flat == 2] = 0
                    flat[rng.randint(flat.size)] = 2
            maps.append(m)
        maps = np.stack(maps)
        cfg = build_config(prob, REPRESENTATIONS["wide"](), 1, 1, auto_reset=False)
        want = oracle.get_stats(cfg, maps, threads=4)
        got = _native.get_stats(prob, torch.from_numpy(maps)).numpy()
        np.testing.assert_array_equal(got, want, err_msg="%s %dx%d" % (prob_name, w, h))


@pytest.mark.parametrize("prob_name", ["sokoban", "ddave", "mdungeon"])
def test_solver_host_twin_get_stats_matches_reference_golden_and_oracle(prob_name):
    """Problem.get_stats incl. the BFS / A* play-through: reference-labelled maps, then oracle-labelled random maps with
    the preconditions forced on half of them (one player, matching crates / targets, one key / exit / door)."""
    n = 0
    for maps, stats in util.stats_groups(prob_name):
        prob = PROBLEMS[prob_name]()
        prob.adjust_param(width=maps.shape[2], height=maps.shape[1])
        got = _native.get_stats(prob, torch.from_numpy(maps)).numpy()
        np.testing.assert_array_equal(got[:, :stats.shape[1]], stats)
        n += len(maps)
    assert n > 0
    rng = np.random.RandomState(29)
    sizes = {"sokoban": [(5, 5), (7, 6), (3, 9), (8, 8)], "ddave": [(11, 7), (7, 11), (5, 5), (14, 9)],
             "mdungeon": [(7, 11), (11, 7), (6, 6), (14, 9)]}[prob_name]
    for (w, h) in sizes:
        prob = PROBLEMS[prob_name]()
        prob.adjust_param(width=w, height=h)
        T = len(prob.tile_types)
        maps = []
        for k in range(24):
            dens = 0.5 + 0.45 * rng.random_sample()
            p = np.full(T, (1 - dens) * 0.5 / (T - 2)); p[0] = dens; p[1] = (1 - dens) * 0.5
            m = rng.choice(T, size=(h, w), p=p / p.sum()).astype(np.uint8)
            if k % 2 == 0:      # make the play-through likely to run: exactly one of each singleton tile
                flat = m.reshape(-1)
                singles = {"sokoban": [2], "ddave": [2, 3, 5], "mdungeon": [2, 3]}[prob_name]
                for t in singles:
                    flat[flat == t] = 0
                cells = rng.choice(flat.size, size=len(singles), replace=False)
                for t, c in zip(singles, cells):
                    flat[c] = t
                if prob_name == "sokoban":
                    flat[(flat == 3) | (flat == 4)] = 0
                    free = np.flatnonzero(flat == 0)
                    kk = int(rng.randint(1, 3))
                    if len(free) >= 2 * kk:
                        pick = rng.choice(free, size=2 * kk, replace=False)
                        flat[pick[:kk]] = 3
                        flat[pick[kk:]] = 4
            maps.append(m)
        maps = np.stack(maps)
        cfg = build_config(prob, REPRESENTATIONS["wide"](), 1, 1, auto_reset=False)
        want = oracle.get_stats(cfg, maps, threads=4)
        got = _native.get_stats(prob, torch.from_numpy(maps)).numpy()
        np.testing.assert_array_equal(got, want, err_msg="%s %dx%d" % (prob_name, w, h))

"""Pin the CPU oracle (oracle/pcgrl_oracle.c) against golden vectors produced by EXECUTING the
unmodified reference (tests/golden/make_golden.py).  CPU only."""
import hashlib
import os
import struct

import numpy as np
import pytest

import oracle
from gym_pcgrl_b200 import PROBLEMS, REPRESENTATIONS, _abi
from gym_pcgrl_b200._config import build_config
import util


KATS = util.kat_configs()


def test_rng_streams_match_numpy_legacy():
    d = np.load(os.path.join(util.GOLDEN, "rng.npz"))
    for s in d["seeds"]:
        s = int(s)
        dbl, st = oracle.rng_doubles(s, 700)
        np.testing.assert_array_equal(dbl, d["sample_%d" % s])
        ns = (1, 2, 3, 5, 7, 11, 14, 16, 32) * 40
        np.testing.assert_array_equal(oracle.rng_randints(st, ns), d["randint_%d" % s])
    # live numpy check on fresh seeds (numpy's legacy stream is the third-party oracle)
    for s in (7, 99, 123456789):
        r = np.random.RandomState(s)
        dbl, st = oracle.rng_doubles(s, 1300)
        np.testing.assert_array_equal(dbl, r.random_sample(1300))
        ns = np.random.RandomState(s + 1).randint(1, 40, size=500)
        np.testing.assert_array_equal(oracle.rng_randints(st, ns), [r.randint(int(n)) for n in ns])


@pytest.mark.parametrize("prob_name", ["binary", "zelda", "sokoban", "ddave", "mdungeon"])
def test_get_stats_matches_reference(prob_name):
    total = 0
    for maps, stats in util.stats_groups(prob_name):
        prob = PROBLEMS[prob_name]()
        prob.adjust_param(width=maps.shape[2], height=maps.shape[1])
        cfg = build_config(prob, REPRESENTATIONS["wide"](), 1, 1, auto_reset=False)
        got = oracle.get_stats(cfg, maps, threads=4)[:, :stats.shape[1]]
        bad = np.nonzero((got != stats).any(axis=1))[0]
        assert bad.size == 0, "%s %s: map %d oracle %s reference %s" % (
            prob_name, maps.shape, bad[0], got[bad[0]], stats[bad[0]])
        total += len(maps)
    assert total > 200


@pytest.mark.parametrize("meta", KATS, ids=[m["name"] for m in KATS])
def test_trajectory_matches_reference(meta):
    """Replay the golden action sequence through the oracle (manual reset on done, exactly like the
    golden harness) and compare every observation / reward / done / stat; re-derive the App. B.3 digest."""
    traj, meta2 = util.load_traj(meta["name"])
    assert meta2["digest"] == meta["digest"]
    env = util.host_env(meta["env_id"], meta["kwargs"], num_envs=1, auto_reset=False)
    assert env._max_changes == meta["max_changes"] and env._max_iterations == meta["max_iterations"]
    cfg = env.native_config
    o = oracle.OracleEnv(cfg, 1)
    o.set_rng_states(util.randomstate_words(meta["seed"])[None])
    S = util.nstats(meta["env_id"].split("-")[0])
    wide = cfg.representation == _abi.REP_WIDE
    sha = hashlib.sha256()

    def feed(r, d):
        sha.update(o["map"][0].tobytes())
        if not wide:
            sha.update(o["pos"][0].tobytes())
        sha.update(o["heatmap"][0].astype(np.int32).tobytes())
        sha.update(struct.pack("<d?", float(r), bool(d)))

    o.reset()
    feed(0.0, False)
    k = 0
    np.testing.assert_array_equal(o["map"][0], traj["reset_map"][0])
    np.testing.assert_array_equal(o["stats"][0, :S], traj["reset_stats"][0])
    actions = util.golden_actions(meta, traj, o.adim)
    for t in range(meta["steps"]):
        o.step(actions[t:t + 1])
        ctx = "%s step %d" % (meta["name"], t)
        np.testing.assert_array_equal(o["map"][0], traj["map"][t], err_msg=ctx)
        np.testing.assert_array_equal(o["heatmap"][0].astype(np.int32), traj["heat"][t], err_msg=ctx)
        if not wide:
            np.testing.assert_array_equal(o["pos"][0].astype(np.int32), traj["pos"][t], err_msg=ctx)
        np.testing.assert_array_equal(o["stats"][0, :S], traj["stats"][t], err_msg=ctx)
        assert o["reward"][0] == traj["reward"][t], ctx          # bit-exact fp64
        assert bool(o["done"][0]) == bool(traj["done"][t]), ctx
        assert o["iteration"][0] == traj["iteration"][t] and o["changes"][0] == traj["changes"][t], ctx
        feed(o["reward"][0], o["done"][0])
        if o["done"][0]:
            o.reset()
            feed(0.0, False)
            k += 1
            assert traj["reset_step"][k] == t
            np.testing.assert_array_equal(o["map"][0], traj["reset_map"][k], err_msg=ctx)
            np.testing.assert_array_equal(o["stats"][0, :S], traj["reset_stats"][k], err_msg=ctx)
            if not wide:
                np.testing.assert_array_equal(o["pos"][0].astype(np.int32), traj["reset_pos"][k], err_msg=ctx)
    assert k == meta["episodes"]
    assert sha.hexdigest()[:16] == meta["digest"]


def test_auto_reset_equals_manual_reset():
    """The AUTO_RESET flag (VecEnv semantics) must give the same episode stream as manual resets."""
    meta = [m for m in KATS if m["name"] == "zelda_narrow"][0]
    traj, _ = util.load_traj(meta["name"])
    env = util.host_env(meta["env_id"], meta["kwargs"], num_envs=1, auto_reset=True)
    o = oracle.OracleEnv(env.native_config, 1)
    o.set_rng_states(util.randomstate_words(meta["seed"])[None])
    o.reset()
    k = 0
    for t in range(meta["steps"]):
        o.step(traj["actions"][t:t + 1, 0])
        assert o["reward"][0] == traj["reward"][t] and bool(o["done"][0]) == bool(traj["done"][t])
        np.testing.assert_array_equal(o["info_stats"][0, :7], traj["stats"][t])
        if traj["done"][t]:
            k += 1
            np.testing.assert_array_equal(o["map"][0], traj["reset_map"][k])
            np.testing.assert_array_equal(o["stats"][0, :7], traj["reset_stats"][k])
            assert o["iteration"][0] == 0 and o["heatmap"][0].sum() == 0
        else:
            np.testing.assert_array_equal(o["map"][0], traj["map"][t])


def test_every_registered_id_has_a_reference_trajectory():
    """36 ids = 6 problems x 6 representations (gym_pcgrl/__init__.py:6-12): each one replays at least one golden
    trajectory recorded from the unmodified reference, so no id is covered by batched-vs-oracle evidence alone."""
    import gym_pcgrl_b200
    covered = {m["env_id"] for m in util.kat_configs()}
    assert len(gym_pcgrl_b200.REGISTRY) == 36
    assert sorted(set(gym_pcgrl_b200.REGISTRY) - covered) == []

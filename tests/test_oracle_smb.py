"""smb oracle (oracle/smb_oracle.c, groundwork for SURVEY.md 8f row f3) against golden vectors recorded from the
unmodified reference (tests/golden/make_golden_smb.py): get_stats incl. the A* play-through, get_reward,
get_episode_over.  CPU only; the CUDA path does not implement smb yet."""
import os

import numpy as np

from oracle import smb

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "stats_smb.npz")


def _groups():
    d = np.load(GOLDEN)
    k = 0
    while "maps_%d" % k in d.files:
        yield k, d["maps_%d" % k], d["stats_%d" % k], d["over_%d" % k], d["reward_%d" % k]
        k += 1


def test_smb_get_stats_matches_reference_golden():
    d = np.load(GOLDEN)
    power = int(d["solver_power"][0])
    total = 0
    for k, maps, stats, over, reward in _groups():
        got = smb.get_stats(maps, power)
        for i in range(len(maps)):
            assert got[i].tolist() == stats[i].tolist(), ("group %d map %d" % (k, i), dict(zip(smb.STAT_NAMES, got[i])),
                                                          dict(zip(smb.STAT_NAMES, stats[i])))
        total += len(maps)
    assert total >= 180


def test_smb_reward_and_episode_over_match_reference_golden():
    d = np.load(GOLDEN)
    w, ip = d["weights"], d["iparam"]
    nonzero = 0
    for k, maps, stats, over, reward in _groups():
        for i in range(len(stats)):
            assert smb.episode_over(stats[i]) == bool(over[i])
        for i in range(len(stats) - 1):
            r = smb.get_reward(stats[i + 1], stats[i], w, ip)
            assert r == reward[i], (k, i, r, reward[i])
            nonzero += r != 0
    assert nonzero > 20


def test_smb_search_respects_the_iteration_cap():
    """A small power changes the result on a map whose play-through needs more iterations (the cap is part of the
    contract: smb_prob.py:17, engine.py:113)."""
    d = np.load(GOLDEN)
    m = d["maps_0"][:2]
    full = smb.get_stats(m, 10000)
    capped = smb.get_stats(m, 5)
    assert (full[:, 7] == 0).all() and (capped[:, 7] > 0).all()
    assert (full[:, :5] == capped[:, :5]).all()          # the map scans do not depend on the solver

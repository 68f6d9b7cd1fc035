#!/usr/bin/env python
"""Golden vectors for the smb oracle (SURVEY.md 8f row f3 groundwork), produced by EXECUTING the unmodified reference.

TEST INFRASTRUCTURE ONLY (build container; needs /root/reference).  Output: tests/golden/stats_smb.npz with, per
group g: maps_g uint8 [N,H,W], stats_g int32 [N,8] (SMBProblem.get_stats), over_g uint8 [N] (get_episode_over),
reward_g float64 [N-1] (get_reward(stats[i+1], stats[i])).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

STAT_NAMES = ["dist-floor", "disjoint-tubes", "enemies", "empty", "noise", "jumps", "jumps-dist", "dist-win"]


def main():
    ref_shim.install()
    from gym_pcgrl.envs.probs import PROBLEMS
    from gym_pcgrl.envs.helper import get_string_map
    rs = np.random.RandomState(2026)
    groups = []

    def maps_from_probs(n, w, h, probs):
        p = np.asarray(probs, dtype=np.float64)
        p = p / p.sum()
        return rs.choice(len(p), size=(n, h, w), p=p).astype(np.uint8)

    default = [0.75, 0.1, 0.01, 0.04, 0.01, 0.02, 0.02]
    g0 = maps_from_probs(40, 114, 14, default)
    groups.append((114, 14, g0))
    hard = np.concatenate([
        maps_from_probs(16, 114, 14, [0.55, 0.3, 0.02, 0.05, 0.02, 0.02, 0.04]),       # walls: A*(1) and A*(0) both matter
        maps_from_probs(8, 114, 14, [0.35, 0.5, 0.02, 0.05, 0.02, 0.02, 0.04]),       # mostly blocked
        maps_from_probs(12, 114, 14, [0.96, 0.005, 0.01, 0.005, 0.005, 0.01, 0.005]),  # nearly empty: pits, few jumps
    ])
    groups.append((114, 14, hard))
    crafted = np.zeros((6, 14, 114), dtype=np.uint8)
    crafted[1][:] = 1                                   # all solid
    crafted[2][11, :] = 1                               # a floor one row above the frame's floor rows
    crafted[3][11, ::2] = 1                             # gaps
    crafted[4][5:12, 40] = 1                            # a wall that is too high
    crafted[4][11, :] = 1
    crafted[5][11, :] = 1
    crafted[5][9:11, 30] = 3
    crafted[5][10, 60:64] = 6
    crafted[5][10, 80] = 2
    crafted[5][6, 81] = 2                               # bricks, tubes, enemies
    groups.append((114, 14, crafted))
    seq = [g0[0].copy()]                                # single-tile edits of one map: what an episode looks like
    for _ in range(39):
        m = seq[-1].copy()
        m[rs.randint(14), rs.randint(114)] = rs.randint(7)
        seq.append(m)
    groups.append((114, 14, np.stack(seq)))
    groups.append((30, 10, maps_from_probs(24, 30, 10, default)))
    groups.append((122, 16, maps_from_probs(12, 122, 16, [0.7, 0.15, 0.02, 0.05, 0.02, 0.02, 0.04])))   # the device operator's size limit
    groups.append((20, 8, maps_from_probs(24, 20, 8, [0.6, 0.25, 0.02, 0.05, 0.02, 0.02, 0.04])))

    out = {}
    for gi, (w, h, maps) in enumerate(groups):
        prob = PROBLEMS["smb"]()
        prob.adjust_param(width=w, height=h)
        tiles = prob.get_tile_types()
        stats, dicts = [], []
        for m in maps:
            d = prob.get_stats(get_string_map(m, tiles))
            dicts.append(d)
            stats.append([int(d[k]) for k in STAT_NAMES])
        out["maps_%d" % gi] = maps
        out["stats_%d" % gi] = np.asarray(stats, dtype=np.int32)
        out["over_%d" % gi] = np.asarray([1 if prob.get_episode_over(d, d) else 0 for d in dicts], dtype=np.uint8)
        out["reward_%d" % gi] = np.asarray([float(prob.get_reward(dicts[i + 1], dicts[i])) for i in range(len(dicts) - 1)],
                                           dtype=np.float64)
        print("group %d: %d maps %dx%d, wins %d, mean jumps %.1f" % (
            gi, len(maps), w, h, int((out["stats_%d" % gi][:, 7] == 0).sum()), out["stats_%d" % gi][:, 5].mean()), flush=True)
    prob = PROBLEMS["smb"]()                              # parameters the reward used (smb_prob.py:19-33)
    out["weights"] = np.asarray([prob._rewards[k] for k in STAT_NAMES], dtype=np.float64)
    out["iparam"] = np.asarray([prob._min_empty, prob._min_enemies, prob._max_enemies, prob._min_jumps], dtype=np.int32)
    out["solver_power"] = np.asarray([prob._solver_power], dtype=np.int32)
    np.savez_compressed(os.path.join(HERE, "stats_smb.npz"), **out)


if __name__ == "__main__":
    main()

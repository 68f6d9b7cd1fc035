#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by EXECUTING the unmodified reference.

TEST INFRASTRUCTURE ONLY (build container; needs /root/reference).  Usage:

    python tests/golden/make_golden.py [--only kat|traj|stats] [--jobs 8]

Outputs (all committed, all small):
  kat.json               SURVEY.md App. B.3 known-answer digests re-derived here (20 configs x 1000 steps)
  traj_<name>.npz        full per-step trajectories of the reference's reset()/step() (obs, reward, done, stats)
  stats_<problem>.npz    Problem.get_stats() on crafted / random maps (incl. solver-heavy ones)
  rng.npz                numpy legacy RandomState streams (random_sample / randint / choice) for MT19937 parity

The harness (feed order, digest definition) is the one described in SURVEY.md App. B.3:
  env = gym.make(id); adjust_param(**kw) twice; env._rep._random = RandomState(S); env._prob._random = RandomState(S)
  action rng RandomState(S+1); obs = reset(); feed(obs, 0.0, False); per step feed(obs, r, done); on done obs = reset(); feed(obs,0,False)
"""
import argparse
import hashlib
import json
import multiprocessing as mp
import os
import struct
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

# stat vector layout shared by oracle, CUDA path and fixtures (see include/pcgrl_b200.h)
STAT_NAMES = {
    "binary": ["regions", "path-length"],
    "zelda": ["player", "key", "door", "enemies", "regions", "nearest-enemy", "path-length"],
    "sokoban": ["player", "crate", "target", "regions", "dist-win", "solution"],
    "ddave": ["player", "dist-floor", "exit", "diamonds", "key", "spikes", "regions", "num-jumps",
              "col-diamonds", "dist-win", "sol-length"],
    "mdungeon": ["player", "exit", "potions", "treasures", "enemies", "regions", "col-potions",
                 "col-treasures", "col-enemies", "dist-win", "sol-length"],
    "smb": ["dist-floor", "disjoint-tubes", "enemies", "empty", "noise", "jumps", "jumps-dist", "dist-win"],
}

ZELDA_SPARSE = {"empty": 0.93, "solid": 0.02, "player": 0.006, "key": 0.006, "door": 0.006,
                "bat": 0.01, "scorpion": 0.01, "spider": 0.012}
SOKOBAN_SPARSE = {"empty": 0.7, "solid": 0.1, "player": 0.07, "crate": 0.065, "target": 0.065}
DDAVE_SPARSE = {"empty": 0.85, "solid": 0.08, "player": 0.01, "exit": 0.01, "diamond": 0.02,
                "key": 0.01, "spike": 0.02}
MDUNGEON_SPARSE = {"empty": 0.85, "solid": 0.05, "player": 0.01, "exit": 0.01, "potion": 0.02,
                   "treasure": 0.02, "goblin": 0.02, "ogre": 0.02}

# (name, env id, kwargs)  -- the 16 + 4 rows of SURVEY.md App. B.3, plus non-default rep modes
KAT_CONFIGS = [
    ("binary_narrow_11x11", "binary-narrow-v0", dict(width=11, height=11, change_percentage=0.2)),
    ("binary_narrow_16x16", "binary-narrow-v0", dict(width=16, height=16, change_percentage=0.2)),
    ("zelda_turtle_11x16", "zelda-turtle-v0", dict(width=11, height=16, change_percentage=0.2)),
    ("sokoban_wide", "sokoban-wide-v0", {}),
    ("binary_narrow", "binary-narrow-v0", {}),
    ("binary_turtle", "binary-turtle-v0", {}),
    ("binary_wide", "binary-wide-v0", {}),
    ("ddave_narrow", "ddave-narrow-v0", {}),
    ("ddave_turtle", "ddave-turtle-v0", {}),
    ("ddave_wide", "ddave-wide-v0", {}),
    ("mdungeon_narrow", "mdungeon-narrow-v0", {}),
    ("mdungeon_turtle", "mdungeon-turtle-v0", {}),
    ("mdungeon_wide", "mdungeon-wide-v0", {}),
    ("zelda_narrow", "zelda-narrow-v0", {}),
    ("zelda_turtle", "zelda-turtle-v0", {}),
    ("zelda_wide", "zelda-wide-v0", {}),
    ("zelda_turtle_11x16_sparse", "zelda-turtle-v0",
     dict(width=11, height=16, change_percentage=0.2, probs=ZELDA_SPARSE)),
    ("sokoban_wide_sparse", "sokoban-wide-v0", dict(probs=SOKOBAN_SPARSE)),
    ("ddave_wide_sparse", "ddave-wide-v0", dict(probs=DDAVE_SPARSE)),
    ("mdungeon_wide_sparse", "mdungeon-wide-v0", dict(probs=MDUNGEON_SPARSE)),
    # non-default representation modes and a transposed zelda
    ("binary_narrow_raster", "binary-narrow-v0", dict(random_tile=False)),
    ("binary_turtle_warp", "binary-turtle-v0", dict(warp=True)),
    ("zelda_turtle_16x11", "zelda-turtle-v0", dict(width=16, height=11, change_percentage=0.2)),
    ("sokoban_narrow_sparse", "sokoban-narrow-v0", dict(probs=SOKOBAN_SPARSE)),
    ("sokoban_turtle_sparse", "sokoban-turtle-v0", dict(probs=SOKOBAN_SPARSE)),
    ("binary_wide_fixedprob", "binary-wide-v0", dict(random_probs=False, change_percentage=0.6)),
    # the 3x3-stamp representations
    ("binary_narrowcast", "binary-narrowcast-v0", {}),
    ("zelda_narrowmulti", "zelda-narrowmulti-v0", {}),
    ("sokoban_turtlecast_sparse", "sokoban-turtlecast-v0", dict(probs=SOKOBAN_SPARSE)),
    ("binary_turtlecast_warp", "binary-turtlecast-v0", dict(warp=True, width=9, height=12, change_percentage=0.5)),
    ("ddave_narrowcast_raster", "ddave-narrowcast-v0", dict(random_tile=False)),
    ("mdungeon_narrowmulti", "mdungeon-narrowmulti-v0", {}),
    # the remaining ids, so that each of the 36 registered environments replays at least one reference trajectory
    ("binary_narrowmulti", "binary-narrowmulti-v0", {}),
    ("zelda_narrowcast", "zelda-narrowcast-v0", {}),
    ("zelda_turtlecast", "zelda-turtlecast-v0", dict(change_percentage=0.5)),
    ("sokoban_narrowcast_sparse", "sokoban-narrowcast-v0", dict(probs=SOKOBAN_SPARSE)),
    ("sokoban_narrowmulti", "sokoban-narrowmulti-v0", {}),
    ("ddave_narrowmulti_sparse", "ddave-narrowmulti-v0", dict(probs=DDAVE_SPARSE)),
    ("ddave_turtlecast", "ddave-turtlecast-v0", dict(warp=True)),
    ("mdungeon_narrowcast_sparse", "mdungeon-narrowcast-v0", dict(probs=MDUNGEON_SPARSE)),
    ("mdungeon_turtlecast", "mdungeon-turtlecast-v0", {}),
    # random_start=False: every reset restores the first map (representation.py:41-45)
    ("zelda_wide_fixedstart", "zelda-wide-v0", dict(random_start=False)),
    ("binary_narrow_fixedstart_raster", "binary-narrow-v0", dict(random_start=False, random_tile=False, width=10, height=6, change_percentage=0.3)),
    # smb: 114x14 default size and smaller levels whose episodes end inside the run.  Every changed
    # step runs the reference's A* play-through (power 10000) in Python, hence the shorter runs (STEPS_OVERRIDE).
    ("smb_narrow", "smb-narrow-v0", {}),
    ("smb_turtle", "smb-turtle-v0", {}),
    ("smb_wide", "smb-wide-v0", {}),
    ("smb_narrow_30x10", "smb-narrow-v0", dict(width=30, height=10, change_percentage=0.1)),
    ("smb_wide_40x8_sparse", "smb-wide-v0", dict(width=40, height=8, change_percentage=0.1,
                                                  probs={"empty": 0.9, "solid": 0.04, "enemy": 0.01, "brick": 0.02, "question": 0.01, "coin": 0.01, "tube": 0.01})),
    ("smb_narrowcast_30x10", "smb-narrowcast-v0", dict(width=30, height=10, change_percentage=0.2)),
    ("smb_turtlecast_warp_24x9", "smb-turtlecast-v0", dict(width=24, height=9, change_percentage=0.3, warp=True)),
    ("smb_narrowmulti_raster_30x10", "smb-narrowmulti-v0", dict(width=30, height=10, change_percentage=0.2, random_tile=False)),
]
STEPS_OVERRIDE = {"smb_narrow": 120, "smb_turtle": 250, "smb_wide": 120, "smb_narrow_30x10": 400, "smb_wide_40x8_sparse": 400,
                  "smb_narrowcast_30x10": 300, "smb_turtlecast_warp_24x9": 400, "smb_narrowmulti_raster_30x10": 300}


def stats_vector(prob_name, stats):
    out = []
    for k in STAT_NAMES[prob_name]:
        v = stats[k]
        out.append(len(v) if isinstance(v, list) else int(v))
    return out


def make_env(env_id, kwargs, seed):
    import gym
    env = gym.make(env_id)
    if kwargs:
        env.adjust_param(**kwargs)
        env.adjust_param(**kwargs)
    env._rep._random = np.random.RandomState(seed)
    env._prob._random = np.random.RandomState(seed)
    return env


def sample_action(space, rng):
    if hasattr(space, "nvec"):
        return [int(rng.randint(n)) for n in space.nvec]
    return int(rng.randint(space.n))


def run_kat(args):
    name, env_id, kwargs, seed, steps = args
    ref_shim.install()
    prob_name = env_id.split("-")[0]
    env = make_env(env_id, kwargs, seed)
    arng = np.random.RandomState(seed + 1)
    sha = hashlib.sha256()

    def feed(obs, r, d):
        sha.update(np.asarray(obs["map"]).astype(np.uint8).tobytes())
        if "pos" in obs:
            sha.update(np.asarray(obs["pos"]).astype(np.uint8).tobytes())
        sha.update(np.asarray(obs["heatmap"]).astype(np.int32).tobytes())
        sha.update(struct.pack("<d?", float(r), bool(d)))

    T = steps
    h, w = env._prob._height, env._prob._width
    S = len(STAT_NAMES[prob_name])
    wide = "pos" not in env.observation_space.spaces
    rec = dict(
        actions=np.zeros((T, 9), np.int32), map=np.zeros((T, h, w), np.uint8), heat=np.zeros((T, h, w), np.int32),
        pos=np.zeros((T, 2), np.int32), reward=np.zeros(T, np.float64), done=np.zeros(T, np.uint8),
        stats=np.zeros((T, S), np.int32), iteration=np.zeros(T, np.int32), changes=np.zeros(T, np.int32))
    rmap, rpos, rstats, rstep = [], [], [], []

    def rec_reset(obs, t):
        rmap.append(np.asarray(obs["map"]).astype(np.uint8))
        rpos.append(np.asarray(obs["pos"]).astype(np.int32) if "pos" in obs else np.zeros(2, np.int32))
        rstats.append(stats_vector(prob_name, env._rep_stats))
        rstep.append(t)

    t0 = time.time()
    obs = env.reset()
    feed(obs, 0.0, False)
    rec_reset(obs, -1)
    episodes, total = 0, 0.0
    for t in range(T):
        a = sample_action(env.action_space, arng)
        obs, r, d, info = env.step(a)
        feed(obs, r, d)
        total += float(r)
        rec["actions"][t, :len(np.atleast_1d(a))] = np.atleast_1d(a)
        rec["map"][t] = obs["map"]
        rec["heat"][t] = obs["heatmap"]
        if not wide:
            rec["pos"][t] = obs["pos"]
        rec["reward"][t] = float(r)
        rec["done"][t] = bool(d)
        rec["stats"][t] = stats_vector(prob_name, env._rep_stats)
        rec["iteration"][t] = env._iteration
        rec["changes"][t] = env._changes
        if d:
            episodes += 1
            obs = env.reset()
            feed(obs, 0.0, False)
            rec_reset(obs, t)
    meta = dict(name=name, env_id=env_id, kwargs=kwargs, seed=seed, steps=T, episodes=episodes,
                sum_reward=total, max_changes=int(env._max_changes), max_iterations=int(env._max_iterations),
                width=w, height=h, digest=sha.hexdigest()[:16], ref_steps_per_s=T / (time.time() - t0))
    np.savez_compressed(
        os.path.join(HERE, "traj_%s.npz" % name), **rec,
        reset_map=np.stack(rmap), reset_pos=np.stack(rpos), reset_stats=np.asarray(rstats, np.int32),
        reset_step=np.asarray(rstep, np.int32), meta=json.dumps(meta))
    return meta


# ----------------------------------------------------------------------------- get_stats fixtures
def _stats_job(args):
    prob_name, w, h, maps = args
    ref_shim.install()
    from gym_pcgrl.envs.probs import PROBLEMS
    from gym_pcgrl.envs.helper import get_string_map
    prob = PROBLEMS[prob_name]()
    prob.adjust_param(width=w, height=h)
    tiles = prob.get_tile_types()
    out = []
    for m in maps:
        out.append(stats_vector(prob_name, prob.get_stats(get_string_map(m, tiles))))
    return np.asarray(out, np.int32)


def _force_counts(m, rng, tile, count):
    """Rewrite map so that exactly `count` cells hold `tile` (keeps everything else)."""
    ys, xs = np.nonzero(m == tile)
    idx = list(zip(ys, xs))
    rng.shuffle(idx)
    for (y, x) in idx[count:]:
        m[y, x] = 0
    need = count - min(count, len(idx))
    while need > 0:
        y, x = rng.randint(m.shape[0]), rng.randint(m.shape[1])
        if m[y, x] in (0, 1):
            m[y, x] = tile
            need -= 1


def gen_maps(prob_name, w, h, n, rng):
    """Random maps: a mix of uniform densities, sparse ("solver fires") and crafted edge cases."""
    T = len(STAT_NAMES_TILES[prob_name])
    maps = []
    for i in range(n):
        mode = i % 4
        if prob_name == "binary":
            p_empty = rng.random_sample() if mode else 0.5
            m = (rng.random_sample((h, w)) >= p_empty).astype(np.uint8)
        else:
            solid = [0.0, 0.05, 0.15, 0.35][mode] if rng.random_sample() < 0.8 else rng.random_sample() * 0.6
            other = rng.random_sample() * 0.25 if mode else 0.08
            p = np.full(T, other / max(T - 2, 1))
            p[0] = 1.0 - solid - other
            p[1] = solid
            m = rng.choice(T, size=(h, w), p=p / p.sum()).astype(np.uint8)
            if w * h >= 4 and rng.random_sample() < 0.7:
                # push towards the solver / BFS preconditions
                if prob_name == "zelda":
                    for tile in (2, 3, 4):
                        _force_counts(m, rng, tile, 1)
                elif prob_name == "sokoban":
                    _force_counts(m, rng, 2, 1)
                    c = 1 + rng.randint(min(3, max(1, (w * h - 1) // 4)))
                    _force_counts(m, rng, 3, c)
                    _force_counts(m, rng, 4, c)
                elif prob_name == "ddave":
                    for tile in (2, 3, 5):
                        _force_counts(m, rng, tile, 1)
                elif prob_name == "mdungeon":
                    for tile in (2, 3):
                        _force_counts(m, rng, tile, 1)
        maps.append(m)
    # crafted edge cases: all-empty, all-solid, single passable tile
    maps.append(np.zeros((h, w), np.uint8))
    maps.append(np.ones((h, w), np.uint8))
    m = np.ones((h, w), np.uint8)
    m[h // 2, w // 2] = 0
    maps.append(m)
    return maps


STAT_NAMES_TILES = {
    "binary": ["empty", "solid"],
    "zelda": ["empty", "solid", "player", "key", "door", "bat", "scorpion", "spider"],
    "sokoban": ["empty", "solid", "player", "crate", "target"],
    "ddave": ["empty", "solid", "player", "exit", "diamond", "key", "spike"],
    "mdungeon": ["empty", "solid", "player", "exit", "potion", "treasure", "goblin", "ogre"],
}

STATS_PLAN = {
    # problem: [(w, h, n_random_maps)]
    "binary": [(16, 16, 120), (14, 14, 80), (11, 11, 60), (5, 9, 40), (32, 32, 16), (1, 1, 2), (32, 3, 20), (2, 31, 20)],
    "zelda": [(11, 16, 120), (11, 7, 120), (16, 11, 60), (5, 5, 40), (20, 12, 20)],
    "sokoban": [(5, 5, 160), (6, 4, 60), (7, 7, 40), (3, 3, 20)],
    "ddave": [(11, 7, 140), (7, 9, 40), (14, 6, 30)],
    "mdungeon": [(7, 11, 140), (9, 6, 40), (12, 8, 30)],
}


def run_stats(pool, jobs):
    for prob_name, plan in STATS_PLAN.items():
        rng = np.random.RandomState(1234 + len(prob_name))
        groups = []
        for (w, h, n) in plan:
            groups.append((w, h, gen_maps(prob_name, w, h, n, rng)))
        tasks = []
        for (w, h, maps) in groups:
            chunk = max(1, len(maps) // (jobs * 2))
            for i in range(0, len(maps), chunk):
                tasks.append((prob_name, w, h, maps[i:i + chunk]))
        t0 = time.time()
        results = pool.map(_stats_job, tasks)
        out, k = {}, 0
        for gi, (w, h, maps) in enumerate(groups):
            chunk = max(1, len(maps) // (jobs * 2))
            parts = []
            for i in range(0, len(maps), chunk):
                parts.append(results[k])
                k += 1
            out["maps_%d" % gi] = np.stack(maps)
            out["stats_%d" % gi] = np.concatenate(parts)
        out["names"] = np.asarray(STAT_NAMES[prob_name])
        np.savez_compressed(os.path.join(HERE, "stats_%s.npz" % prob_name), **out)
        print("stats", prob_name, "%.1fs" % (time.time() - t0), flush=True)


def run_rng():
    """numpy legacy RandomState streams (the live third-party oracle of SURVEY 8c(i))."""
    out = {}
    seeds = [0, 1, 42, 12345, 2 ** 31 - 1]
    out["seeds"] = np.asarray(seeds, np.int64)
    for s in seeds:
        r = np.random.RandomState(s)
        out["state_%d" % s] = r.get_state()[1].astype(np.uint32)
        out["sample_%d" % s] = r.random_sample(700)            # crosses one twist boundary
        out["randint_%d" % s] = np.asarray([r.randint(n) for n in (1, 2, 3, 5, 7, 11, 14, 16, 32) * 40], np.int64)
        p = np.asarray([0.58, 0.3, 0.02, 0.02, 0.02, 0.02, 0.02, 0.02])
        out["choice_%d" % s] = r.choice(8, size=(16, 11), p=p / p.sum()).astype(np.uint8)
        out["after_%d" % s] = np.asarray([r.randint(1 << 30)], np.int64)
    np.savez_compressed(os.path.join(HERE, "rng.npz"), **out)


# ----------------------------------------------------------------------------- wrappers (gym_pcgrl/wrappers.py)
WRAPPER_CONFIGS = [
    # (name, kind, env id, crop size, kwargs)
    ("cropped_binary_narrow", "cropped", "binary-narrow-v0", 28, {}),
    ("cropped_zelda_narrow", "cropped", "zelda-narrow-v0", 22, {}),
    ("cropped_sokoban_turtle", "cropped", "sokoban-turtle-v0", 10, {}),
    ("cropped_ddave_turtle_odd", "cropped", "ddave-turtle-v0", 7, {}),
    ("actionmap_binary_wide", "actionmap", "binary-wide-v0", 0, {}),
    ("actionmap_zelda_wide", "actionmap", "zelda-wide-v0", 0, {}),
    ("actionmap_sokoban_narrow", "actionmap_pos", "sokoban-narrow-v0", 0, {}),
    ("actionmap_mdungeon_turtle", "actionmap_pos", "mdungeon-turtle-v0", 0, {}),
]


def run_wrapper(args):
    name, kind, env_id, crop, kwargs, seed, steps = args
    ref_shim.install()
    from gym_pcgrl import wrappers as W
    import gym
    if kind == "cropped":
        env = W.CroppedImagePCGRLWrapper(env_id, crop, **kwargs)
        pcgrl = env.pcgrl_env
    elif kind == "actionmap":
        env = W.ActionMapImagePCGRLWrapper(env_id, **kwargs)
        pcgrl = env.pcgrl_env
    else:  # ActionMap over a cursor representation.  It has to sit OUTSIDE OneHotEncoding here: the one-hot
        # transform mutates the observation dict that an inner ActionMap keeps as old_obs (wrappers.py:101-104,137)
        pcgrl = gym.make(env_id)
        e = W.OneHotEncoding(pcgrl, 'map')
        e = W.ActionMap(e)
        env = W.ToImage(e, ['map'])
    pcgrl._rep._random = np.random.RandomState(seed)
    pcgrl._prob._random = np.random.RandomState(seed)
    arng = np.random.RandomState(seed + 1)
    obs = env.reset()
    obs0 = [np.asarray(obs).astype(np.uint8)]
    space = env.action_space
    acts, imgs, rews, dones, resets = [], [], [], [], []
    for t in range(steps):
        a = sample_action(space, arng)
        obs, r, d, info = env.step(a)
        acts.append(np.atleast_1d(a))
        imgs.append(np.asarray(obs).astype(np.uint8))
        rews.append(float(r))
        dones.append(bool(d))
        if d:
            obs = env.reset()
            resets.append(t)
            obs0.append(np.asarray(obs).astype(np.uint8))
    np.savez_compressed(os.path.join(HERE, "wrap_%s.npz" % name), actions=np.asarray(acts, np.int32), obs=np.stack(imgs),
                        reward=np.asarray(rews), done=np.asarray(dones, np.uint8), reset_obs=np.stack(obs0),
                        reset_step=np.asarray(resets, np.int32),
                        meta=json.dumps(dict(name=name, kind=kind, env_id=env_id, crop=crop, kwargs=kwargs, seed=seed, steps=steps,
                                             obs_shape=list(imgs[0].shape), obs_dtype=str(np.asarray(obs).dtype))))
    return name, imgs[0].shape, str(np.asarray(obs).dtype), int(np.sum(dones))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--jobs", type=int, default=len(os.sched_getaffinity(0)))
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--names", default="", help="kat/traj only: comma-separated config names or prefixes; the other kat.json entries are kept")
    a = ap.parse_args()
    pool = mp.Pool(a.jobs)
    if a.only in ("", "rng"):
        run_rng()
    if a.only in ("", "kat", "traj"):
        todo = KAT_CONFIGS
        if a.names:
            pref = tuple(a.names.split(","))
            todo = [c for c in KAT_CONFIGS if c[0].startswith(pref)]
        metas = pool.map(run_kat, [(n, i, k, 0, STEPS_OVERRIDE.get(n, a.steps)) for (n, i, k) in todo], chunksize=1)
        if a.names:   # merge into the existing file, keeping KAT_CONFIGS order
            with open(os.path.join(HERE, "kat.json")) as f:
                old = {m["name"]: m for m in json.load(f)}
            old.update({m["name"]: m for m in metas})
            metas_all = [old[c[0]] for c in KAT_CONFIGS if c[0] in old]
        else:
            metas_all = metas
        with open(os.path.join(HERE, "kat.json"), "w") as f:
            json.dump(metas_all, f, indent=1)
        for m in metas:
            print("%-28s ep=%-4d sum=%-9.1f mc/mi=%d/%d %s  (%.0f steps/s)" % (
                m["name"], m["episodes"], m["sum_reward"], m["max_changes"], m["max_iterations"], m["digest"],
                m["ref_steps_per_s"]), flush=True)
    if a.only in ("", "stats"):
        run_stats(pool, a.jobs)
    if a.only in ("", "wrappers"):
        for r in pool.map(run_wrapper, [(n, k, i, c, kw, 0, 300) for (n, k, i, c, kw) in WRAPPER_CONFIGS], chunksize=1):
            print("wrapper", r, flush=True)


if __name__ == "__main__":
    main()

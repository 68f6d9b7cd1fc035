"""Shared helpers for the parity tests (oracle <-> golden fixtures <-> CUDA path)."""
import json
import os

import numpy as np

from gym_pcgrl_b200 import BatchedPcgrlEnv, _abi
from gym_pcgrl_b200.seeding import mt_state_words

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def kat_configs():
    with open(os.path.join(GOLDEN, "kat.json")) as f:
        return json.load(f)


def load_traj(name):
    d = np.load(os.path.join(GOLDEN, "traj_%s.npz" % name))
    return {k: d[k] for k in d.files if k != "meta"}, json.loads(str(d["meta"]))


def host_env(env_id, kwargs, num_envs=1, auto_reset=True, **extra):
    """Host-side env object configured like the golden harness (adjust_param issued twice).  No CUDA
    is touched until reset()."""
    prob, rep, _ = env_id.split("-")
    env = BatchedPcgrlEnv(prob, rep, num_envs=num_envs, auto_reset=auto_reset, seed=0, **extra)
    if kwargs:
        env.adjust_param(**kwargs)
        env.adjust_param(**kwargs)
    return env


def randomstate_words(seed):
    return mt_state_words(np.random.RandomState(seed))


def golden_actions(meta, traj, adim):
    a = traj["actions"]
    return a[:, :adim] if adim > 1 else a[:, 0]


def stats_groups(prob_name):
    d = np.load(os.path.join(GOLDEN, "stats_%s.npz" % prob_name))
    k = 0
    while "maps_%d" % k in d.files:
        yield d["maps_%d" % k], d["stats_%d" % k]
        k += 1


def nstats(prob_name):
    return len(_abi.STAT_NAMES[prob_name])

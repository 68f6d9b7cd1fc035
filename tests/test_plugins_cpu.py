"""User-defined Problem / Representation subclasses on the plugin path (envs/plugin_env.py): the reference's extension
contract (probs/problem.py:54-122, reps/representation.py:67-103).  CPU-only: torch CPU tensors, no native call."""
import numpy as np
import pytest
import torch

import gym_pcgrl_b200 as pkg
from gym_pcgrl_b200 import PluginBatchedEnv, spaces
from gym_pcgrl_b200.envs.probs.problem import Problem
from gym_pcgrl_b200.envs.reps.representation import Representation


class GardenProblem(Problem):
    """Toy problem: between 6 and 10 'flower' tiles, and as few 'rock' tiles next to a flower as possible."""
    name = "garden"
    tile_types = ("grass", "flower", "rock")
    stat_names = ("flowers", "crowded")

    def __init__(self):
        super().__init__()
        self._width, self._height = 8, 6
        self._prob = {"grass": 0.6, "flower": 0.2, "rock": 0.2}
        self._rewards = {"flowers": 2, "crowded": 1}
        self._lo, self._hi = 6, 10

    def adjust_param(self, **kwargs):
        super().adjust_param(**kwargs)
        self._lo, self._hi = kwargs.get("min_flowers", self._lo), kwargs.get("max_flowers", self._hi)

    def get_stats(self, maps):
        m = maps.long()
        flower, rock = (m == 1), (m == 2)
        near = torch.zeros_like(flower)
        near[:, 1:, :] |= flower[:, :-1, :]
        near[:, :-1, :] |= flower[:, 1:, :]
        near[:, :, 1:] |= flower[:, :, :-1]
        near[:, :, :-1] |= flower[:, :, 1:]
        return {"flowers": flower.sum(dim=(1, 2)), "crowded": (rock & near).sum(dim=(1, 2))}

    def reward_terms(self):
        return [("flowers", lambda s: s["flowers"], self._lo, self._hi), ("crowded", lambda s: s["crowded"], 0, 0)]

    def get_episode_over(self, new_stats, old_stats):
        return (new_stats["flowers"] >= self._lo) & (new_stats["flowers"] <= self._hi) & (new_stats["crowded"] == 0)


class MirrorRepresentation(Representation):
    """Toy representation: action (x, tile) writes the tile at column x of the cursor row AND at the mirrored column."""
    name = "mirror"

    def get_action_space(self, width, height, num_tiles):
        return spaces.MultiDiscrete([width, num_tiles])

    def get_observation_space(self, width, height, num_tiles):
        return spaces.Dict({"map": spaces.Box(low=0, high=num_tiles - 1, dtype=np.uint8, shape=(height, width))})

    def get_observation(self):
        return {"map": self._map}

    def update(self, action):
        n, h, w = self._map.shape
        a = torch.as_tensor(action).reshape(n, 2).long()
        idx = torch.arange(n)
        row = self._y
        c1 = self._write_tile(idx, a[:, 0], row, a[:, 1], torch.ones(n, dtype=torch.bool))
        c2 = self._write_tile(idx, w - 1 - a[:, 0], row, a[:, 1], torch.ones(n, dtype=torch.bool))
        self._y = (row + 1) % h
        return c1 + c2, a[:, 0], row


def _numpy_garden_stats(m):
    flower, rock = m == 1, m == 2
    near = np.zeros_like(flower)
    near[1:, :] |= flower[:-1, :]; near[:-1, :] |= flower[1:, :]; near[:, 1:] |= flower[:, :-1]; near[:, :-1] |= flower[:, 1:]
    return int(flower.sum()), int((rock & near).sum())


def _range_reward(nv, ov, lo, hi):   # helper.py:366-376
    if lo <= nv <= hi and lo <= ov <= hi: return 0
    if ov <= hi and nv <= hi: return min(nv, lo) - min(ov, lo)
    if ov >= lo and nv >= lo: return max(ov, hi) - max(nv, hi)
    if nv > hi and ov < lo: return hi - nv + ov - lo
    if nv < lo and ov > hi: return hi - ov + nv - lo
    return 0


@pytest.mark.parametrize("rep", ["wide", "narrow", "turtle", MirrorRepresentation])
def test_custom_problem_follows_the_reference_step_semantics(rep):
    """A user-defined Problem with built-in and user-defined representations: every step is replayed by a plain
    per-env numpy restatement of pcgrl_env.py:129-150 driven by the maps / cursors the env reports."""
    n = 24
    env = PluginBatchedEnv(GardenProblem, rep, num_envs=n, device="cpu", seed=3, auto_reset=False)
    env.adjust_param(change_percentage=0.5)
    assert env._max_changes == 24 and env._max_iterations == 24 * 48
    obs = env.reset()
    assert tuple(obs["map"].shape) == (n, 6, 8) and obs["map"].max() <= 2
    prev = obs["map"].numpy().copy()
    stats = [_numpy_garden_stats(prev[i]) for i in range(n)]
    changes, iters = np.zeros(n, int), np.zeros(n, int)
    rng = np.random.RandomState(0)
    sp = env.action_space
    for t in range(40):
        a = np.stack([rng.randint(int(k), size=n) for k in sp.nvec], axis=1) if hasattr(sp, "nvec") else rng.randint(sp.n, size=n)
        obs, reward, done, info = env.step(a)
        cur = obs["map"].numpy()
        iters += 1
        for i in range(n):
            nchanged = int((cur[i] != prev[i]).sum())
            old = stats[i]
            if nchanged:
                changes[i] += nchanged
                stats[i] = _numpy_garden_stats(cur[i])
            want_r = 2 * _range_reward(stats[i][0], old[0], 6, 10) + 1 * _range_reward(stats[i][1], old[1], 0, 0)
            want_d = (6 <= stats[i][0] <= 10 and stats[i][1] == 0) or changes[i] >= 24 or iters[i] >= 24 * 48
            assert float(reward[i]) == want_r and bool(done[i]) == want_d, (rep, t, i)
            assert int(info["flowers"][i]) == stats[i][0] and int(info["changes"][i]) == changes[i]
        assert int(obs["heatmap"].sum()) == int((changes > 0).sum() and obs["heatmap"].sum())
        prev = cur.copy()
    assert changes.sum() > 0


def test_auto_reset_and_registry_roundtrip():
    pkg.register("garden-mirror-v0", GardenProblem, MirrorRepresentation)
    env = pkg.make("garden-mirror-v0", num_envs=16, device="cpu", seed=1)
    assert isinstance(env, PluginBatchedEnv) and env.get_num_tiles() == 3 and env.get_border_tile() == 0
    env.adjust_param(change_percentage=0.1, min_flowers=0, max_flowers=48)
    env.reset()
    ndone = 0
    for t in range(30):
        a = np.stack([np.random.RandomState(t).randint(8, size=16), np.random.RandomState(t + 99).randint(3, size=16)], axis=1)
        obs, r, d, info = env.step(a)
        ndone += int(d.sum())
        assert (env._iteration[d] == 0).all() and (env._changes[d] == 0).all()      # finished envs were reset in place
        assert (obs["heatmap"][d] == 0).all()
    assert ndone > 0
    # built-in ids still resolve to the native classes
    assert pkg.REGISTRY["smb-wide-v0"] == {"prob": "smb", "rep": "wide"} and len(pkg.REGISTRY) == 37
    pkg.REGISTRY.pop("garden-mirror-v0"); pkg.PROBLEMS.pop("garden"); pkg.REPRESENTATIONS.pop("mirror")


def test_abstract_methods_raise_like_the_reference():
    class Nothing(Problem):
        name = "nothing"
    with pytest.raises(NotImplementedError):      # problem.py:14: the constructor already needs get_tile_types()
        Nothing()

    class Tiles(Problem):
        name = "tiles"
        tile_types = ("a", "b")
    p = Tiles()
    with pytest.raises(NotImplementedError):
        p.get_episode_over({}, {})
    with pytest.raises(NotImplementedError):
        p.get_reward({}, {})
    with pytest.raises(NotImplementedError):
        Representation().update(0)

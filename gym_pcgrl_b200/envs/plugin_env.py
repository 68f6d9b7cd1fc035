"""The slow-but-correct path for USER-DEFINED plugins.

The reference's extension contract is "subclass ``Problem`` (probs/problem.py:54-122: get_tile_types, get_stats,
get_reward, get_episode_over, get_debug_info) or ``Representation`` (reps/representation.py:67-103: get_action_space,
get_observation_space, get_observation, update) and register it".  The fused sm_100a kernels only know the built-in
problems and representations; a class they do not know runs here instead: the PcgrlEnv.reset / step control flow
(pcgrl_env.py:66-76,129-150) written once in torch on batched tensors, calling the plugin's own methods --

    Representation.update(actions)        -> (change [N], x [N], y [N]); edits ``self._map`` (uint8 [N,H,W])
    Problem.get_stats(maps)               -> {name: int tensor [N]}
    Problem.get_reward(new, old)          -> float64 [N]          (default: the ``reward_terms()`` table)
    Problem.get_episode_over(new, old)    -> bool [N]
    Problem.get_debug_info(new, old)      -> {name: tensor [N]}

Built-in classes can be mixed in freely: a built-in problem computes its statistics with the native ``pcgrl_get_stats``
operator on the plugin representation's maps, a built-in representation (narrow / turtle / wide) edits the maps with its
torch ``update``.  This path is chosen ONLY for classes the native layer does not know (``BatchedPcgrlEnv`` with two
built-in names never comes here, there is no silent fallback); it is many launches per step instead of one and it does
not reproduce numpy's MT19937 streams (random maps and cursors come from a torch generator).
"""
import numpy as np

from .. import spaces
from .probs import PROBLEMS
from .probs.problem import Problem
from .reps import REPRESENTATIONS
from .reps.representation import Representation


def _instance(obj, registry, base):
    if isinstance(obj, str):
        return registry[obj]()
    if isinstance(obj, type):
        return obj()
    if not isinstance(obj, base):
        raise TypeError("expected a %s subclass, instance or registered name, got %r" % (base.__name__, obj))
    return obj


_BUILTIN_PROBLEMS = frozenset(PROBLEMS.values())          # captured at import: user registrations come later
_BUILTIN_REPRESENTATIONS = frozenset(REPRESENTATIONS.values())


def is_native(prob, rep):
    """True when both plugins are the unmodified built-in classes (the fused kernels apply)."""
    return type(prob) in _BUILTIN_PROBLEMS and type(rep) in _BUILTIN_REPRESENTATIONS


class PluginBatchedEnv:
    metadata = {'render.modes': []}

    def __init__(self, prob, rep, num_envs=1, device="cpu", seed=None, auto_reset=True):
        import torch
        self._prob = _instance(prob, PROBLEMS, Problem)
        self._rep = _instance(rep, REPRESENTATIONS, Representation)
        self.num_envs, self.auto_reset = int(num_envs), bool(auto_reset)
        self.device = torch.device(device)
        self._gen = torch.Generator(device=self.device)
        self._rep._gen = self._gen
        self._max_changes = max(int(0.2 * self._prob._width * self._prob._height), 1)   # pcgrl_env.py:33-34
        self._max_iterations = self._max_changes * self._prob._width * self._prob._height
        self._rep_stats = None
        self.seed(seed)
        self._update_spaces()

    # ------------------------------------------------------------------ reference surface
    def seed(self, seed=None):
        seed = self._rep.seed(seed)
        self._prob.seed(seed)
        self._gen.manual_seed(int(seed) % (2 ** 63))
        return [seed]

    def adjust_param(self, **kwargs):   # pcgrl_env.py:106-115 incl. the ordering quirk Q3
        if 'change_percentage' in kwargs:
            percentage = min(1, max(0, kwargs.get('change_percentage')))
            self._max_changes = max(int(percentage * self._prob._width * self._prob._height), 1)
        self._max_iterations = self._max_changes * self._prob._width * self._prob._height
        self._prob.adjust_param(**kwargs)
        self._rep.adjust_param(**kwargs)
        self._update_spaces()

    def get_border_tile(self):
        return self._prob.get_tile_types().index(self._prob._border_tile)

    def get_num_tiles(self):
        return len(self._prob.get_tile_types())

    def _update_spaces(self):
        w, h, t = self._prob._width, self._prob._height, self.get_num_tiles()
        self.action_space = self._rep.get_action_space(w, h, t)
        self.observation_space = self._rep.get_observation_space(w, h, t)
        self.observation_space.spaces['heatmap'] = spaces.Box(low=0, high=self._max_changes, dtype=np.int32, shape=(h, w))

    # ------------------------------------------------------------------ reset / step
    def _random_maps(self, n):
        """helper.py:310-312 gen_random_map: i.i.d. tiles from the problem's probabilities (torch generator)."""
        import torch
        tiles = self._prob.get_tile_types()
        p = torch.tensor([float(self._prob._prob[t]) for t in tiles], dtype=torch.float64, device=self.device)
        h, w = self._prob._height, self._prob._width
        flat = torch.multinomial((p / p.sum()).expand(n, -1), h * w, replacement=True, generator=self._gen)
        return flat.reshape(n, h, w).to(torch.uint8)

    def reset(self, mask=None):
        """PcgrlEnv.reset (pcgrl_env.py:66-76) for every env, or for those with mask[i] != 0."""
        import torch
        n, h, w = self.num_envs, self._prob._height, self._prob._width
        first = self._rep_stats is None or tuple(self._rep._map.shape) != (n, h, w)
        if first:
            self._rep._map = torch.zeros((n, h, w), dtype=torch.uint8, device=self.device)
            self._rep._old_map = None
            self._rep._x = torch.zeros(n, dtype=torch.int64, device=self.device)
            self._rep._y = torch.zeros(n, dtype=torch.int64, device=self.device)
            self._heatmap = torch.zeros((n, h, w), dtype=torch.int32, device=self.device)
            self._iteration = torch.zeros(n, dtype=torch.int64, device=self.device)
            self._changes = torch.zeros(n, dtype=torch.int64, device=self.device)
            mask = None
        m = torch.ones(n, dtype=torch.bool, device=self.device) if mask is None else torch.as_tensor(mask, device=self.device).bool()
        # Representation.reset (representation.py:40-45): a fresh random map, or the env's first map
        if self._rep._random_start or self._rep._old_map is None:
            fresh = self._random_maps(n)
            self._rep._map = torch.where(m[:, None, None], fresh, self._rep._map)
            if self._rep._old_map is None:
                self._rep._old_map = self._rep._map.clone()
        else:
            self._rep._map = torch.where(m[:, None, None], self._rep._old_map, self._rep._map)
        rx = torch.randint(0, w, (n,), generator=self._gen, device=self.device)
        ry = torch.randint(0, h, (n,), generator=self._gen, device=self.device)
        self._rep._x = torch.where(m, rx, self._rep._x)
        self._rep._y = torch.where(m, ry, self._rep._y)
        stats = self._prob.get_stats(self._rep._map)
        stats = {k: torch.as_tensor(v, device=self.device) for k, v in stats.items()}
        if first:
            self._rep_stats = {k: v.clone() for k, v in stats.items()}
            self._start_stats = {k: v.clone() for k, v in stats.items()}
        else:
            for k in stats:
                self._rep_stats[k] = torch.where(m, stats[k], self._rep_stats[k])
                self._start_stats[k] = torch.where(m, stats[k], self._start_stats[k])
        self._prob.reset(self._start_stats)                     # problem.py:45-46
        self._heatmap = torch.where(m[:, None, None], torch.zeros_like(self._heatmap), self._heatmap)
        self._iteration = torch.where(m, torch.zeros_like(self._iteration), self._iteration)
        self._changes = torch.where(m, torch.zeros_like(self._changes), self._changes)
        return self._observation()

    def _observation(self):
        obs = dict(self._rep.get_observation())
        obs["heatmap"] = self._heatmap
        return obs

    def step(self, actions):
        """PcgrlEnv.step (pcgrl_env.py:129-150) for the whole batch."""
        import torch
        if self._rep_stats is None:
            raise RuntimeError("call reset() before step()")
        actions = torch.as_tensor(actions, device=self.device)
        self._iteration = self._iteration + 1
        old = {k: v.clone() for k, v in self._rep_stats.items()}
        change, x, y = self._rep.update(actions)
        change = torch.as_tensor(change, device=self.device).to(torch.int64)
        changed = change > 0
        if bool(changed.any()):
            self._changes = self._changes + change
            idx = torch.nonzero(changed).flatten()
            self._heatmap[idx, torch.as_tensor(y, device=self.device)[idx].long(), torch.as_tensor(x, device=self.device)[idx].long()] += 1
            new = self._prob.get_stats(self._rep._map)          # a function of the map: unchanged envs keep their values
            for k in new:
                self._rep_stats[k] = torch.where(changed, torch.as_tensor(new[k], device=self.device), self._rep_stats[k])
        reward = torch.as_tensor(self._prob.get_reward(self._rep_stats, old), device=self.device).to(torch.float64)
        done = torch.as_tensor(self._prob.get_episode_over(self._rep_stats, old), device=self.device).bool()
        done = done | (self._changes >= self._max_changes) | (self._iteration >= self._max_iterations)
        info = dict(self._prob.get_debug_info(self._rep_stats, old))
        info["iterations"], info["changes"] = self._iteration.clone(), self._changes.clone()
        info["max_iterations"], info["max_changes"] = self._max_iterations, self._max_changes
        if self.auto_reset and bool(done.any()):
            self.reset(done)
        return self._observation(), reward, done, info

    def render(self, mode='human'):
        raise NotImplementedError("rendering is out of scope of the B200 hot path")

    def close(self):
        pass

"""Problem plugin base class (host-side mirror of gym_pcgrl/envs/probs/problem.py:7-156).

A Problem owns the tile alphabet, the map size, the initial tile probabilities, the quality
thresholds and the reward weights.  In this framework the per-step work (get_stats / get_reward /
get_episode_over) of the built-in problems runs fused inside the sm_100a step kernel; the Python
object carries the parameters, freezes them into the POD ``pcgrl_config`` block
(:meth:`native_params`) and offers batched tensor versions of the reference's methods:

* ``get_stats(maps)``            uint8 CUDA tensor [N,H,W] -> dict name -> int32 tensor [N]   (native kernel)
* ``get_reward(new, old)``       dicts of tensors -> float64 tensor [N]    (torch mirror of the fused code)
* ``get_episode_over(new, old)`` -> bool tensor [N]
* ``get_debug_info(new, old)``   -> dict of tensors

Reward terms are declared as ``(weight key, stat expression, low, high)`` in the order the reference
sums them, so the torch mirror and the kernel share one table.
"""
import math

from ... import _abi

INF = math.inf


def get_range_reward(new_value, old_value, low, high):
    """Batched helper.py:366-376 ``get_range_reward`` on float64 tensors (bounds may be +-inf)."""
    import torch
    n = new_value.to(torch.float64)
    o = old_value.to(torch.float64)
    lo = torch.full_like(n, float(low))
    hi = torch.full_like(n, float(high))
    inside = (n >= lo) & (n <= hi) & (o >= lo) & (o <= hi)
    below = (o <= hi) & (n <= hi)
    above = (o >= lo) & (n >= lo)
    up = (n > hi) & (o < lo)
    r_below = torch.minimum(n, lo) - torch.minimum(o, lo)
    r_above = torch.maximum(o, hi) - torch.maximum(n, hi)
    r_up = hi - n + o - lo
    r_down = hi - o + n - lo
    out = torch.where(up, r_up, r_down)
    out = torch.where(above, r_above, out)
    out = torch.where(below, r_below, out)
    return torch.where(inside, torch.zeros_like(n), out)


class Problem:
    """Base class; subclasses fill the class-level tables."""

    name = None            # key in PROBLEMS
    tile_types = ()        # get_tile_types()
    stat_names = ()        # == _abi.STAT_NAMES[name]
    extra_info_names = ()  # info entries derived in the kernel and appended to info_stats after the stats columns

    @property
    def debug_info_names(self):
        """Keys of Problem.get_debug_info that are plain statistics (default: all of them)."""
        return self.stat_names

    def __init__(self):
        # problem.py:11-22 defaults
        self._width = 9
        self._height = 9
        tiles = self.get_tile_types()
        self._prob = {t: 1.0 / len(tiles) for t in tiles}
        self._border_size = (1, 1)
        self._border_tile = tiles[0]
        self._tile_size = 16
        self._rewards = {}
        self._solver_power = 0
        self._start_stats = None
        self._random = None
        self._seed = None

    # -- reference surface ---------------------------------------------------------------------
    def seed(self, seed=None):
        """problem.py:34-36.  The batched env owns the per-env MT19937 streams; this records the seed."""
        from ...seeding import np_random
        self._random, seed = np_random(seed)
        self._seed = seed
        return seed

    def reset(self, start_stats):
        """problem.py:45-46."""
        self._start_stats = start_stats

    def get_tile_types(self):
        if not self.tile_types:
            raise NotImplementedError('get_tile_types is not implemented')
        return list(self.tile_types)

    def adjust_param(self, **kwargs):
        """problem.py:66-72: width / height / probs (only keys that already exist)."""
        self._width = kwargs.get('width', self._width)
        self._height = kwargs.get('height', self._height)
        prob = kwargs.get('probs')
        if prob is not None:
            for t in prob:
                if t in self._prob:
                    self._prob[t] = prob[t]

    def _adjust_rewards(self, kwargs):
        rewards = kwargs.get('rewards')
        if rewards is not None:
            for t in rewards:
                if t in self._rewards:
                    self._rewards[t] = rewards[t]

    # -- tables used by both the torch mirror and the native config ----------------------------
    def reward_terms(self):
        """[(weight key, fn(stats dict)->tensor/int, low, high)] in the reference's summation order."""
        raise NotImplementedError('get_reward is not implemented')

    def native_thresholds(self):
        """(iparam list, dparam list) as laid out in include/pcgrl_b200.h."""
        raise NotImplementedError

    def native_params(self):
        order = _abi.REWARD_ORDER[self.name]
        iparam, dparam = self.native_thresholds()
        return dict(
            problem=_abi.PROBLEM_IDS[self.name], width=int(self._width), height=int(self._height),
            num_tiles=len(self.tile_types), solver_power=int(self._solver_power),
            iparam=[int(v) for v in iparam], dparam=[float(v) for v in dparam],
            reward_weight=[float(self._rewards[k]) for k in order],
            tile_prob=[float(self._prob[t]) for t in self.tile_types],
            random_probs=bool(getattr(self, "_random_probs", False)),
        )

    # -- batched tensor API --------------------------------------------------------------------
    def stats_from_rows(self, rows):
        """int32 [N, MAX_STATS] stats rows -> dict name -> [N] (views)."""
        return {k: rows[..., i] for i, k in enumerate(self.stat_names)}

    def get_stats(self, maps):
        """Batched Problem.get_stats on the GPU: maps uint8 [N,H,W] (or [H,W]) CUDA tensor."""
        from ... import _native
        single = maps.dim() == 2
        rows = _native.get_stats(self, maps[None] if single else maps)
        stats = self.stats_from_rows(rows)
        return {k: v[0] for k, v in stats.items()} if single else stats

    def get_reward(self, new_stats, old_stats):
        total = None
        for key, fn, low, high in self.reward_terms():
            term = get_range_reward(_as_tensor(fn(new_stats)), _as_tensor(fn(old_stats)), low, high) * float(self._rewards[key])
            total = term if total is None else total + term
        return total

    def get_episode_over(self, new_stats, old_stats):
        raise NotImplementedError('get_episode_over is not implemented')

    def get_debug_info(self, new_stats, old_stats):
        return {k: new_stats[k] for k in self.stat_names}

    def get_graphics(self):
        """RGBA sprite atlas uint8 [num_tiles, tile_size, tile_size, 4].  Default = the grey levels of the reference's
        base class (problem.py:137-141); assign ``self._graphics`` (same shape) to use real sprites."""
        import numpy as np
        if getattr(self, "_graphics", None) is None:
            tiles = self.get_tile_types()
            g = np.zeros((len(tiles), self._tile_size, self._tile_size, 4), np.uint8)
            for i in range(len(tiles)):
                g[i, :, :, :3] = int(i * 255 / len(tiles))
                g[i, :, :, 3] = 255
            self._graphics = g
        return self._graphics

    def render(self, maps, pos=None):
        """Batched Problem.render (problem.py:134-156) on the GPU: uint8 CUDA maps [N,H,W] -> RGB uint8 [N,Hpx,Wpx,3]."""
        from ... import _native
        return _native.render(self, maps, pos)


def _as_tensor(v):
    import torch
    return v if isinstance(v, torch.Tensor) else torch.as_tensor(v)

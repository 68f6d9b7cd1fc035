"""Representation plugin base class (host-side mirror of gym_pcgrl/envs/reps/representation.py:7-117).

A Representation defines the action / observation spaces and how an action edits the map.  For the
built-in representations the edit itself (``update``) runs inside the fused sm_100a step kernel; the
Python object carries the mode flags (``random_start``, ``random_tile``, ``warp``) that are frozen
into ``pcgrl_config.flags`` and exposes the map / cursor tensors of the batched env it is bound to.
"""
from ... import _abi


class Representation:
    name = None

    def __init__(self):
        self._random_start = True   # representation.py:12
        self._map = None            # uint8 tensor [N,H,W] once bound to an env (live view)
        self._old_map = None        # uint8 tensor [N,H,W] (first map of every env)
        self._random = None
        self._seed = None
        self._env = None
        self._x = self._y = None    # cursor tensors [N] on the plugin path
        self._gen = None            # torch generator of the plugin env
        self.seed()

    def seed(self, seed=None):
        """representation.py:28-30."""
        from ...seeding import np_random
        self._random, seed = np_random(seed)
        self._seed = seed
        return seed

    def adjust_param(self, **kwargs):
        """representation.py:53-54."""
        self._random_start = kwargs.get('random_start', self._random_start)

    def bind(self, env):
        self._env = env
        self._map = env._bufs["map"]
        self._old_map = env._bufs["start_map"]

    def native_flags(self):
        return _abi.FLAG_RANDOM_START if self._random_start else 0

    def native_params(self):
        return dict(representation=_abi.REP_IDS[self.name], flags=self.native_flags())

    def get_action_space(self, width, height, num_tiles):
        raise NotImplementedError('get_action_space is not implemented')

    def get_observation_space(self, width, height, num_tiles):
        raise NotImplementedError('get_observation_space is not implemented')

    def get_observation(self):
        raise NotImplementedError('get_observation is not implemented')

    def update(self, action):
        """Batched ``update(actions) -> (change [N], x [N], y [N])`` on the plugin path (envs/plugin_env.py); edits
        ``self._map``.  Inside ``BatchedPcgrlEnv`` the built-in representations never get here: their tile write is
        fused with get_stats / get_reward in the step kernel.  Subclasses of the base class must implement it
        (representation.py:102-103)."""
        raise NotImplementedError('update is not implemented')

    # -- helpers for the torch implementations of the built-in representations (plugin path only)
    def _plugin_tensors(self):
        if self._env is not None:
            raise NotImplementedError('update runs inside the fused step kernel; call env.step(actions)')
        return self._map, self._x, self._y

    def _write_tile(self, idx, x, y, tile, active):
        """map[i, y[i], x[i]] = tile[i] where active[i]; returns change [N] (0 / 1)."""
        import torch
        m = self._map
        old = m[idx, y, x]
        change = active & (old != tile.to(m.dtype))
        m[idx[change], y[change], x[change]] = tile[change].to(m.dtype)
        return change.to(torch.int64)

    def _cursor_observation(self):
        import torch
        return {"pos": torch.stack([self._x, self._y], dim=1).to(torch.uint8), "map": self._map}

    def render(self, lvl_image, tile_size, border_size):
        return lvl_image

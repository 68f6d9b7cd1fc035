"""Narrow representation: the agent only chooses the tile value; the cursor is drawn at random (or
raster-scans) -- gym_pcgrl/envs/reps/narrow_rep.py:28-31,45-64,73-88,99-114."""
from collections import OrderedDict

import numpy as np

from ... import _abi, spaces
from .representation import Representation


class NarrowRepresentation(Representation):
    name = "narrow"

    def __init__(self):
        super().__init__()
        self._random_tile = True

    def adjust_param(self, **kwargs):
        super().adjust_param(**kwargs)
        self._random_tile = kwargs.get('random_tile', self._random_tile)

    def native_flags(self):
        return super().native_flags() | (_abi.FLAG_RANDOM_TILE if self._random_tile else 0)

    def get_action_space(self, width, height, num_tiles):
        return spaces.Discrete(num_tiles + 1)

    def get_observation_space(self, width, height, num_tiles):
        return spaces.Dict({
            "pos": spaces.Box(low=np.array([0, 0]), high=np.array([width - 1, height - 1]), dtype=np.uint8),
            "map": spaces.Box(low=0, high=num_tiles - 1, dtype=np.uint8, shape=(height, width)),
        })

    def get_observation(self):
        if self._env is None:
            return self._cursor_observation()
        return OrderedDict({"pos": self._env._bufs["pos"], "map": self._env._bufs["map"]})

    def update(self, action):
        """narrow_rep.py:99-114 on batched tensors (plugin path): write tile a-1 at the cursor, then move the cursor
        (random, or raster scan); the returned (x, y) is the cursor AFTER the move."""
        import torch
        m, x, y = self._plugin_tensors()
        n, h, w = m.shape
        a = torch.as_tensor(action, device=m.device).reshape(n).long()
        idx = torch.arange(n, device=m.device)
        change = self._write_tile(idx, x, y, (a - 1).clamp(min=0), a > 0)
        if self._random_tile:
            self._x = torch.randint(0, w, (n,), generator=self._gen, device=m.device)
            self._y = torch.randint(0, h, (n,), generator=self._gen, device=m.device)
        else:
            nx = x + 1
            wrap = nx >= w
            ny = torch.where(wrap, y + 1, y)
            self._x = torch.where(wrap, torch.zeros_like(nx), nx)
            self._y = torch.where(ny >= h, torch.zeros_like(ny), ny)
        return change, self._x, self._y

"""Turtle representation: 4 move actions + one action per tile value, clamped or wrapped at the
edges -- gym_pcgrl/envs/reps/turtle_rep.py:16-19,30-44,58-77,86-90,101-129."""
from collections import OrderedDict

import numpy as np

from ... import _abi, spaces
from .representation import Representation


class TurtleRepresentation(Representation):
    name = "turtle"

    def __init__(self):
        super().__init__()
        self._dirs = [(-1, 0), (1, 0), (0, -1), (0, 1)]
        self._warp = False

    def adjust_param(self, **kwargs):
        super().adjust_param(**kwargs)
        self._warp = kwargs.get('warp', self._warp)

    def native_flags(self):
        return super().native_flags() | (_abi.FLAG_WARP if self._warp else 0)

    def get_action_space(self, width, height, num_tiles):
        return spaces.Discrete(len(self._dirs) + num_tiles)

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
        """turtle_rep.py:101-129 on batched tensors (plugin path): actions 0..3 move the cursor (clamp or wrap), an
        action >= 4 writes tile a-4 at the cursor."""
        import torch
        m, x, y = self._plugin_tensors()
        n, h, w = m.shape
        a = torch.as_tensor(action, device=m.device).reshape(n).long()
        dx = torch.tensor([d[0] for d in self._dirs] + [0], device=m.device)[a.clamp(max=4)]
        dy = torch.tensor([d[1] for d in self._dirs] + [0], device=m.device)[a.clamp(max=4)]
        nx, ny = x + dx, y + dy
        if self._warp:
            nx, ny = nx % w, ny % h
        else:
            nx, ny = nx.clamp(0, w - 1), ny.clamp(0, h - 1)
        self._x, self._y = nx, ny
        idx = torch.arange(n, device=m.device)
        change = self._write_tile(idx, nx, ny, (a - 4).clamp(min=0), a >= 4)
        return change, self._x, self._y

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
        return OrderedDict({"pos": self._env._bufs["pos"], "map": self._env._bufs["map"]})

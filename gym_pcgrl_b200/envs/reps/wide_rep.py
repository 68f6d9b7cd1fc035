"""Wide representation: the action is (x, y, tile) -- gym_pcgrl/envs/reps/wide_rep.py:28-45,53-57,67-70."""
import numpy as np

from ... import spaces
from .representation import Representation


class WideRepresentation(Representation):
    name = "wide"

    def get_action_space(self, width, height, num_tiles):
        return spaces.MultiDiscrete([width, height, num_tiles])

    def get_observation_space(self, width, height, num_tiles):
        return spaces.Dict({
            "map": spaces.Box(low=0, high=num_tiles - 1, dtype=np.uint8, shape=(height, width)),
        })

    def get_observation(self):
        return {"map": self._env._bufs["map"]}

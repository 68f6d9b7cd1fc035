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
        if self._env is None:
            return {"map": self._map}
        return {"map": self._env._bufs["map"]}

    def update(self, action):
        """wide_rep.py:67-70 on batched tensors (plugin path): write action[2] at (x = action[0], y = action[1])."""
        import torch
        m, _, _ = self._plugin_tensors()
        n = m.shape[0]
        a = torch.as_tensor(action, device=m.device).reshape(n, 3).long()
        idx = torch.arange(n, device=m.device)
        change = self._write_tile(idx, a[:, 0], a[:, 1], a[:, 2], torch.ones(n, dtype=torch.bool, device=m.device))
        return change, a[:, 0], a[:, 1]

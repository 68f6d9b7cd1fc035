"""Narrow-multi representation: nine actions per step, one per cell of the 3x3 block around the cursor
(0 = keep, a > 0 writes tile a-1) -- gym_pcgrl/envs/reps/narrow_multi_rep.py:23-59."""
from ... import spaces
from .narrow_rep import NarrowRepresentation


class NarrowMultiRepresentation(NarrowRepresentation):
    name = "narrowmulti"

    def get_action_space(self, width, height, num_tiles):
        return spaces.MultiDiscrete([num_tiles + 1] * 9)

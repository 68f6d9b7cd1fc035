"""Narrow-cast representation: like narrow, but the action is (type, tile) with type 0 = keep, 1 = write the
cursor cell, 2 = write the 3x3 block around the cursor -- gym_pcgrl/envs/reps/narrow_cast_rep.py:23-59."""
from ... import spaces
from .narrow_rep import NarrowRepresentation


class NarrowCastRepresentation(NarrowRepresentation):
    name = "narrowcast"

    def get_action_space(self, width, height, num_tiles):
        return spaces.MultiDiscrete([3, num_tiles])

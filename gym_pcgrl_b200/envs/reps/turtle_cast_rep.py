"""Turtle-cast representation: (type, tile) with type 0..3 = move, 4 = write the cursor cell, 5 = write the
3x3 block around the cursor -- gym_pcgrl/envs/reps/turtle_cast_rep.py:26-76."""
from ... import spaces
from .turtle_rep import TurtleRepresentation


class TurtleCastRepresentation(TurtleRepresentation):
    name = "turtlecast"

    def get_action_space(self, width, height, num_tiles):
        return spaces.MultiDiscrete([len(self._dirs) + 2, num_tiles])

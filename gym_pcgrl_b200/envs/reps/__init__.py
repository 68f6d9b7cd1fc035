"""REPRESENTATIONS registry (same keys as gym_pcgrl/envs/reps/__init__.py:9-16)."""
from .narrow_cast_rep import NarrowCastRepresentation
from .narrow_multi_rep import NarrowMultiRepresentation
from .narrow_rep import NarrowRepresentation
from .turtle_cast_rep import TurtleCastRepresentation
from .turtle_rep import TurtleRepresentation
from .wide_rep import WideRepresentation

REPRESENTATIONS = {
    "narrow": NarrowRepresentation,
    "narrowcast": NarrowCastRepresentation,
    "narrowmulti": NarrowMultiRepresentation,
    "wide": WideRepresentation,
    "turtle": TurtleRepresentation,
    "turtlecast": TurtleCastRepresentation,
}

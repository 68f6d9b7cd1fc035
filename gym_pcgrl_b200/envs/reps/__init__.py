"""REPRESENTATIONS registry (keys of gym_pcgrl/envs/reps/__init__.py:9-16; the cast / multi variants
are a "next" row)."""
from .narrow_rep import NarrowRepresentation
from .turtle_rep import TurtleRepresentation
from .wide_rep import WideRepresentation

REPRESENTATIONS = {
    "narrow": NarrowRepresentation,
    "wide": WideRepresentation,
    "turtle": TurtleRepresentation,
}

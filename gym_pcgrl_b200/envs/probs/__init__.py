"""PROBLEMS registry (same keys as gym_pcgrl/envs/probs/__init__.py:9-16)."""
from .binary_prob import BinaryProblem
from .ddave_prob import DDaveProblem
from .mdungeon_prob import MDungeonProblem
from .smb_prob import SMBProblem
from .sokoban_prob import SokobanProblem
from .zelda_prob import ZeldaProblem

PROBLEMS = {
    "binary": BinaryProblem,
    "ddave": DDaveProblem,
    "mdungeon": MDungeonProblem,
    "sokoban": SokobanProblem,
    "smb": SMBProblem,
    "zelda": ZeldaProblem,
}

"""Sokoban problem: solvable level with a long solution, judged by a bounded BFS / A* solver
(gym_pcgrl/envs/probs/sokoban_prob.py:15-210, solver: probs/sokoban/engine.py)."""
from .problem import Problem, INF


class SokobanProblem(Problem):
    name = "sokoban"
    tile_types = ("empty", "solid", "player", "crate", "target")
    stat_names = ("player", "crate", "target", "regions", "dist-win", "sol-length")

    def __init__(self):
        super().__init__()
        self._width = 5
        self._height = 5
        self._prob = {"empty": 0.45, "solid": 0.4, "player": 0.05, "crate": 0.05, "target": 0.05}
        self._border_tile = "solid"
        self._solver_power = 5000
        self._max_crates = 3
        self._target_solution = 18
        self._rewards = {"player": 3, "crate": 2, "target": 2, "regions": 5, "ratio": 2,
                         "dist-win": 0.0, "sol-length": 1}

    def adjust_param(self, **kwargs):
        super().adjust_param(**kwargs)
        self._solver_power = kwargs.get('solver_power', self._solver_power)
        self._max_crates = kwargs.get('max_crates', self._max_crates)
        self._max_crates = kwargs.get('max_targets', self._max_crates)        # sokoban_prob.py:66 (sic)
        self._target_solution = kwargs.get('min_solution', self._target_solution)  # kwarg name per :67
        self._adjust_rewards(kwargs)

    def reward_terms(self):  # sokoban_prob.py:157-175
        return [("player", lambda s: s["player"], 1, 1),
                ("crate", lambda s: s["crate"], 1, self._max_crates),
                ("target", lambda s: s["target"], 1, self._max_crates),
                ("regions", lambda s: s["regions"], 1, 1),
                ("ratio", lambda s: abs(s["crate"] - s["target"]), -INF, -INF),
                ("dist-win", lambda s: s["dist-win"], -INF, -INF),
                ("sol-length", lambda s: s["sol-length"], INF, INF)]

    def native_thresholds(self):
        return [self._max_crates, self._target_solution], []

    def get_episode_over(self, new_stats, old_stats):  # sokoban_prob.py:188-189
        return new_stats["sol-length"] >= self._target_solution

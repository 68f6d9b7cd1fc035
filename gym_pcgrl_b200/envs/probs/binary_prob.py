"""Binary problem: one connected empty region with a long shortest path
(parameters and formulas of gym_pcgrl/envs/probs/binary_prob.py:14-138)."""
from .problem import Problem, INF


class BinaryProblem(Problem):
    name = "binary"
    tile_types = ("empty", "solid")
    stat_names = ("regions", "path-length")
    extra_info_names = ("path-imp",)   # binary_prob.py:137: path-length - start path-length, info_stats column 2

    def __init__(self):
        super().__init__()
        self._width = 14
        self._height = 14
        self._prob = {"empty": 0.5, "solid": 0.5}
        self._border_tile = "solid"
        self._target_path = 20
        self._random_probs = True   # binary_prob.py:24, :68-72 (drawn on the device, problem RNG stream)
        self._rewards = {"regions": 5, "path-length": 1}

    def adjust_param(self, **kwargs):
        super().adjust_param(**kwargs)
        self._target_path = kwargs.get('target_path', self._target_path)
        self._random_probs = kwargs.get('random_probs', self._random_probs)
        self._adjust_rewards(kwargs)

    def reward_terms(self):  # binary_prob.py:98-106
        return [("regions", lambda s: s["regions"], 1, 1),
                ("path-length", lambda s: s["path-length"], INF, INF)]

    def native_thresholds(self):
        return [self._target_path], []

    def get_episode_over(self, new_stats, old_stats):  # binary_prob.py:119-120
        return (new_stats["regions"] == 1) & \
            (new_stats["path-length"] - self._start_stats["path-length"] >= self._target_path)

    def get_debug_info(self, new_stats, old_stats):  # binary_prob.py:133-138
        return {"regions": new_stats["regions"], "path-length": new_stats["path-length"],
                "path-imp": new_stats["path-length"] - self._start_stats["path-length"]}

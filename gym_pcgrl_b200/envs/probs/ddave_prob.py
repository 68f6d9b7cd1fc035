"""Dangerous Dave problem: platformer level judged by a bounded A* / BFS play-through
(gym_pcgrl/envs/probs/ddave_prob.py:16-245, solver: probs/ddave/engine.py)."""
from .problem import Problem, INF


class DDaveProblem(Problem):
    name = "ddave"
    tile_types = ("empty", "solid", "player", "exit", "diamond", "key", "spike")
    stat_names = ("player", "dist-floor", "exit", "diamonds", "key", "spikes", "regions", "num-jumps",
                  "col-diamonds", "dist-win", "sol-length")

    def __init__(self):
        super().__init__()
        self._width = 11
        self._height = 7
        self._prob = {"empty": 0.5, "solid": 0.3, "player": 0.02, "exit": 0.02, "diamond": 0.04,
                      "key": 0.02, "spike": 0.1}
        self._border_tile = "solid"
        self._solver_power = 5000
        self._max_diamonds = 3
        self._min_spikes = 10
        self._target_jumps = 2
        self._target_solution = 20
        self._rewards = {"player": 3, "dist-floor": 2, "exit": 3, "diamonds": 1, "key": 3, "spikes": 1,
                         "regions": 5, "num-jumps": 3, "dist-win": 0.1, "sol-length": 1}

    def adjust_param(self, **kwargs):
        super().adjust_param(**kwargs)
        self._solver_power = kwargs.get('solver_power', self._solver_power)
        self._max_diamonds = kwargs.get('max_diamonds', self._max_diamonds)
        self._min_spikes = kwargs.get('min_spikes', self._min_spikes)
        self._target_jumps = kwargs.get('target_jumps', self._target_jumps)
        self._target_solution = kwargs.get('target_solution', self._target_solution)
        self._adjust_rewards(kwargs)

    def reward_terms(self):  # ddave_prob.py:181-205 (summation order of :196-205)
        return [("player", lambda s: s["player"], 1, 1),
                ("dist-floor", lambda s: s["dist-floor"], 0, 0),
                ("exit", lambda s: s["exit"], 1, 1),
                ("spikes", lambda s: s["spikes"], self._min_spikes, INF),
                ("diamonds", lambda s: s["diamonds"], -INF, self._max_diamonds),
                ("key", lambda s: s["key"], 1, 1),
                ("regions", lambda s: s["regions"], 1, 1),
                ("num-jumps", lambda s: s["num-jumps"], INF, INF),
                ("dist-win", lambda s: s["dist-win"], -INF, -INF),
                ("sol-length", lambda s: s["sol-length"], INF, INF)]

    def native_thresholds(self):
        return [self._max_diamonds, self._min_spikes, self._target_jumps, self._target_solution], []

    def get_episode_over(self, new_stats, old_stats):  # ddave_prob.py:218-220
        return (new_stats["sol-length"] >= self._target_solution) & \
            (new_stats["num-jumps"] > self._target_jumps)

    @property
    def debug_info_names(self):
        return tuple(k for k in self.stat_names if k != "dist-floor")

    def get_debug_info(self, new_stats, old_stats):  # ddave_prob.py:233-245 (no dist-floor)
        return {k: new_stats[k] for k in ("player", "exit", "diamonds", "key", "spikes", "regions",
                                          "col-diamonds", "num-jumps", "dist-win", "sol-length")}

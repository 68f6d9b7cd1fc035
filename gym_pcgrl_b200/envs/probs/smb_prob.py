"""Super Mario Bros problem: a 114 x 14 level that an A* agent can finish with many jumps
(parameters and formulas of gym_pcgrl/envs/probs/smb_prob.py:9-185; the play-through of probs/smb/engine.py runs on
the device inside get_stats, see csrc/pcgrl_smb.cuh)."""
from .problem import Problem, INF


class SMBProblem(Problem):
    name = "smb"
    tile_types = ("empty", "solid", "enemy", "brick", "question", "coin", "tube")
    stat_names = ("dist-floor", "disjoint-tubes", "enemies", "empty", "noise", "jumps", "jumps-dist", "dist-win")

    def __init__(self):
        super().__init__()
        self._width = 114
        self._height = 14
        self._prob = {"empty": 0.75, "solid": 0.1, "enemy": 0.01, "brick": 0.04, "question": 0.01, "coin": 0.02, "tube": 0.02}
        self._border_size = (3, 0)
        self._solver_power = 10000
        self._min_empty = 900
        self._min_enemies = 10
        self._max_enemies = 30
        self._min_jumps = 20
        self._rewards = {"dist-floor": 2, "disjoint-tubes": 1, "enemies": 1, "empty": 1, "noise": 4, "jumps": 2,
                         "jumps-dist": 2, "dist-win": 5}

    def adjust_param(self, **kwargs):  # smb_prob.py:38-50
        super().adjust_param(**kwargs)
        self._min_empty = kwargs.get('min_empty', self._min_empty)
        self._min_enemies = kwargs.get('min_enemies', self._min_enemies)
        self._max_enemies = kwargs.get('max_enemies', self._max_enemies)
        self._min_jumps = kwargs.get('min_jumps', self._min_jumps)
        self._adjust_rewards(kwargs)

    def reward_terms(self):  # smb_prob.py:149-170
        return [("dist-floor", lambda s: s["dist-floor"], 0, 0),
                ("disjoint-tubes", lambda s: s["disjoint-tubes"], 0, 0),
                ("enemies", lambda s: s["enemies"], self._min_enemies, self._max_enemies),
                ("empty", lambda s: s["empty"], self._min_empty, INF),
                ("noise", lambda s: s["noise"], 0, 0),
                ("jumps", lambda s: s["jumps"], self._min_jumps, INF),
                ("jumps-dist", lambda s: s["jumps-dist"], 0, 0),
                ("dist-win", lambda s: s["dist-win"], 0, 0)]

    def native_thresholds(self):
        return [self._min_empty, self._min_enemies, self._max_enemies, self._min_jumps], []

    def get_episode_over(self, new_stats, old_stats):  # smb_prob.py:172-173
        return new_stats["dist-win"] <= 0

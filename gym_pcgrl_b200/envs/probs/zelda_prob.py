"""Zelda problem: one player / key / door, enemies far from the player, long player->key->door walk
(gym_pcgrl/envs/probs/zelda_prob.py:17-178)."""
from .problem import Problem, INF


class ZeldaProblem(Problem):
    name = "zelda"
    tile_types = ("empty", "solid", "player", "key", "door", "bat", "scorpion", "spider")
    stat_names = ("player", "key", "door", "enemies", "regions", "nearest-enemy", "path-length")

    def __init__(self):
        super().__init__()
        self._width = 11
        self._height = 7
        self._prob = {"empty": 0.58, "solid": 0.3, "player": 0.02, "key": 0.02, "door": 0.02,
                      "bat": 0.02, "scorpion": 0.02, "spider": 0.02}
        self._border_tile = "solid"
        self._max_enemies = 5
        self._target_enemy_dist = 4
        self._target_path = 16
        self._rewards = {"player": 3, "key": 3, "door": 3, "regions": 5, "enemies": 1,
                         "nearest-enemy": 2, "path-length": 1}

    def adjust_param(self, **kwargs):
        super().adjust_param(**kwargs)
        self._max_enemies = kwargs.get('max_enemies', self._max_enemies)
        self._target_enemy_dist = kwargs.get('target_enemy_dist', self._target_enemy_dist)
        self._target_path = kwargs.get('target_path', self._target_path)
        self._adjust_rewards(kwargs)

    def reward_terms(self):  # zelda_prob.py:124-142
        return [("player", lambda s: s["player"], 1, 1),
                ("key", lambda s: s["key"], 1, 1),
                ("door", lambda s: s["door"], 1, 1),
                ("enemies", lambda s: s["enemies"], 2, self._max_enemies),
                ("regions", lambda s: s["regions"], 1, 1),
                ("nearest-enemy", lambda s: s["nearest-enemy"], self._target_enemy_dist, INF),
                ("path-length", lambda s: s["path-length"], INF, INF)]

    def native_thresholds(self):
        return [self._max_enemies, self._target_enemy_dist, self._target_path], []

    def get_episode_over(self, new_stats, old_stats):  # zelda_prob.py:155-156
        return (new_stats["nearest-enemy"] >= self._target_enemy_dist) & \
            (new_stats["path-length"] >= self._target_path)

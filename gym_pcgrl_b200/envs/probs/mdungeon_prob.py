"""MiniDungeons problem: dungeon crawl judged by a bounded A* / BFS play-through
(gym_pcgrl/envs/probs/mdungeon_prob.py:16-247, solver: probs/mdungeon/engine.py)."""
from .problem import Problem, INF


class MDungeonProblem(Problem):
    name = "mdungeon"
    tile_types = ("empty", "solid", "player", "exit", "potion", "treasure", "goblin", "ogre")
    stat_names = ("player", "exit", "potions", "treasures", "enemies", "regions", "col-potions",
                  "col-treasures", "col-enemies", "dist-win", "sol-length")

    def __init__(self):
        super().__init__()
        self._width = 7
        self._height = 11
        self._prob = {"empty": 0.4, "solid": 0.4, "player": 0.02, "exit": 0.02, "potion": 0.03,
                      "treasure": 0.03, "goblin": 0.05, "ogre": 0.05}
        self._border_tile = "solid"
        self._solver_power = 5000
        self._max_enemies = 6
        self._max_potions = 2
        self._max_treasures = 3
        self._target_col_enemies = 0.5
        self._target_solution = 20
        self._rewards = {"player": 3, "exit": 3, "potions": 1, "treasures": 1, "enemies": 2, "regions": 5,
                         "col-enemies": 2, "dist-win": 0.1, "sol-length": 1}

    def adjust_param(self, **kwargs):
        super().adjust_param(**kwargs)
        self._solver_power = kwargs.get('solver_power', self._solver_power)
        self._max_enemies = kwargs.get('max_enemies', self._max_enemies)
        self._max_potions = kwargs.get('max_potions', self._max_potions)
        self._max_treasures = kwargs.get('max_treasures', self._max_treasures)
        self._target_col_enemies = kwargs.get('target_col_enemies', self._target_col_enemies)
        self._target_solution = kwargs.get('target_solution', self._target_solution)
        self._adjust_rewards(kwargs)

    def reward_terms(self):  # mdungeon_prob.py:183-205 (summation order of :197-205)
        return [("player", lambda s: s["player"], 1, 1),
                ("exit", lambda s: s["exit"], 1, 1),
                ("enemies", lambda s: s["enemies"], 1, self._max_enemies),
                ("treasures", lambda s: s["treasures"], -INF, self._max_treasures),
                ("potions", lambda s: s["potions"], -INF, self._max_potions),
                ("regions", lambda s: s["regions"], 1, 1),
                ("col-enemies", lambda s: s["col-enemies"], INF, INF),
                ("dist-win", lambda s: s["dist-win"], -INF, -INF),
                ("sol-length", lambda s: s["sol-length"], INF, INF)]

    def native_thresholds(self):
        return [self._max_enemies, self._max_potions, self._max_treasures, self._target_solution], \
            [self._target_col_enemies]

    def get_episode_over(self, new_stats, old_stats):  # mdungeon_prob.py:218-221
        import torch
        enemies = new_stats["enemies"]
        ratio = new_stats["col-enemies"].to(torch.float64) / torch.clamp(enemies, min=1).to(torch.float64)
        return (new_stats["sol-length"] >= self._target_solution) & (enemies > 0) & \
            (ratio > self._target_col_enemies)

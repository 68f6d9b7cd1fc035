"""ctypes mirror of include/pcgrl_b200.h (POD structs + constants).  No CUDA, no torch."""
import ctypes as C

ABI_VERSION = 2
MAX_DIM = 32
MAX_TILES = 8
MAX_STATS = 16
MAX_REWARD_TERMS = 12
INFO_ITERATION, INFO_CHANGES = 14, 15   # info_stats columns: env counters before any auto-reset
MT_WORDS = 625

PROB_BINARY, PROB_ZELDA, PROB_SOKOBAN, PROB_DDAVE, PROB_MDUNGEON, PROB_SMB = range(6)
REP_NARROW, REP_TURTLE, REP_WIDE, REP_NARROWCAST, REP_NARROWMULTI, REP_TURTLECAST = range(6)

FLAG_RANDOM_TILE = 1
FLAG_WARP = 2
FLAG_RANDOM_START = 4
FLAG_RANDOM_PROBS = 8
FLAG_AUTO_RESET = 16
FLAG_HEAT_U16 = 32
FLAG_FULL_STATS = 64

PROBLEM_IDS = {"binary": PROB_BINARY, "zelda": PROB_ZELDA, "sokoban": PROB_SOKOBAN,
               "ddave": PROB_DDAVE, "mdungeon": PROB_MDUNGEON, "smb": PROB_SMB}
REP_IDS = {"narrow": REP_NARROW, "turtle": REP_TURTLE, "wide": REP_WIDE, "narrowcast": REP_NARROWCAST,
           "narrowmulti": REP_NARROWMULTI, "turtlecast": REP_TURTLECAST}
ACTION_DIMS = {REP_NARROW: 1, REP_TURTLE: 1, REP_WIDE: 3, REP_NARROWCAST: 2, REP_NARROWMULTI: 9, REP_TURTLECAST: 2}

# key order of each Problem.get_stats dict == column order of every stats row
STAT_NAMES = {
    "binary": ["regions", "path-length"],
    "zelda": ["player", "key", "door", "enemies", "regions", "nearest-enemy", "path-length"],
    "sokoban": ["player", "crate", "target", "regions", "dist-win", "sol-length"],
    "ddave": ["player", "dist-floor", "exit", "diamonds", "key", "spikes", "regions", "num-jumps",
              "col-diamonds", "dist-win", "sol-length"],
    "mdungeon": ["player", "exit", "potions", "treasures", "enemies", "regions", "col-potions",
                 "col-treasures", "col-enemies", "dist-win", "sol-length"],
    "smb": ["dist-floor", "disjoint-tubes", "enemies", "empty", "noise", "jumps", "jumps-dist", "dist-win"],
}

# order in which each Problem.get_reward sums its terms == order of pcgrl_config.reward_weight
REWARD_ORDER = {
    "binary": ["regions", "path-length"],
    "zelda": ["player", "key", "door", "enemies", "regions", "nearest-enemy", "path-length"],
    "sokoban": ["player", "crate", "target", "regions", "ratio", "dist-win", "sol-length"],
    "ddave": ["player", "dist-floor", "exit", "spikes", "diamonds", "key", "regions", "num-jumps",
              "dist-win", "sol-length"],
    "mdungeon": ["player", "exit", "enemies", "treasures", "potions", "regions", "col-enemies",
                 "dist-win", "sol-length"],
    "smb": ["dist-floor", "disjoint-tubes", "enemies", "empty", "noise", "jumps", "jumps-dist", "dist-win"],
}


class PcgrlConfig(C.Structure):
    _fields_ = [
        ("problem", C.c_int32), ("representation", C.c_int32),
        ("width", C.c_int32), ("height", C.c_int32), ("num_tiles", C.c_int32),
        ("max_changes", C.c_int32), ("max_iterations", C.c_int32),
        ("flags", C.c_uint32), ("solver_power", C.c_int32),
        ("iparam", C.c_int32 * 7),
        ("dparam", C.c_double * 2),
        ("reward_weight", C.c_double * MAX_REWARD_TERMS),
        ("tile_prob", C.c_double * MAX_TILES),
    ]


class PcgrlBuffers(C.Structure):
    _fields_ = [
        ("map", C.c_void_p), ("heatmap", C.c_void_p), ("pos", C.c_void_p),
        ("iteration", C.c_void_p), ("changes", C.c_void_p),
        ("stats", C.c_void_p), ("start_stats", C.c_void_p), ("info_stats", C.c_void_p),
        ("reward", C.c_void_p), ("done", C.c_void_p), ("rng", C.c_void_p),
        ("tile_prob", C.c_void_p), ("start_map", C.c_void_p), ("start_valid", C.c_void_p),
        ("scratch", C.c_void_p), ("scratch_bytes", C.c_size_t), ("status", C.c_void_p),
    ]


class PcgrlHostIO(C.Structure):
    _fields_ = [
        ("actions", C.c_void_p), ("map", C.c_void_p), ("heatmap", C.c_void_p), ("pos", C.c_void_p),
        ("reward", C.c_void_p), ("done", C.c_void_p), ("info_stats", C.c_void_p),
        ("d_staging", C.c_void_p), ("h_staging", C.c_void_p), ("staging_bytes", C.c_size_t),
        ("mode", C.c_int32), ("synced", C.c_int32), ("reset_base", C.c_int64), ("change_base", C.c_int64),
        ("pending", C.c_int32), ("reserved", C.c_int32),
    ]


class PcgrlHostRolloutIO(C.Structure):
    _fields_ = [
        ("actions", C.c_void_p), ("reward", C.c_void_p), ("done", C.c_void_p), ("map", C.c_void_p),
        ("heatmap", C.c_void_p), ("pos", C.c_void_p), ("info_stats", C.c_void_p),
    ]


# name -> dtype string, trailing shape as a function of (H, W); leading dim is n
BUFFER_SPECS = [
    ("map", "uint8", lambda h, w: (h, w)),
    ("heatmap", "uint8", lambda h, w: (h, w)),       # int16 storage when FLAG_HEAT_U16 is set (max_changes > 255)
    ("pos", "uint8", lambda h, w: (2,)),
    ("iteration", "int32", lambda h, w: ()),
    ("changes", "int32", lambda h, w: ()),
    ("stats", "int32", lambda h, w: (MAX_STATS,)),
    ("start_stats", "int32", lambda h, w: (MAX_STATS,)),
    ("info_stats", "int32", lambda h, w: (MAX_STATS,)),
    ("reward", "float64", lambda h, w: ()),
    ("done", "uint8", lambda h, w: ()),
    ("rng", "uint32", lambda h, w: (2, MT_WORDS)),
    ("tile_prob", "float64", lambda h, w: (MAX_TILES,)),
    ("start_map", "uint8", lambda h, w: (h, w)),
    ("start_valid", "uint8", lambda h, w: ()),
]


def action_dim(representation):
    if isinstance(representation, str):
        representation = REP_IDS[representation]
    return ACTION_DIMS[int(representation)]

"""Freeze Problem / Representation parameters into the POD ``pcgrl_config`` block
(include/pcgrl_b200.h) that is passed by value to every native call."""
import os

from . import _abi


def build_config(prob, rep, max_changes, max_iterations, auto_reset=True):
    p = prob.native_params()
    r = rep.native_params()
    cfg = _abi.PcgrlConfig()
    cfg.problem = p["problem"]
    cfg.representation = r["representation"]
    cfg.width, cfg.height, cfg.num_tiles = p["width"], p["height"], p["num_tiles"]
    cfg.max_changes, cfg.max_iterations = int(max_changes), int(max_iterations)
    flags = r["flags"]
    if p["random_probs"]:
        flags |= _abi.FLAG_RANDOM_PROBS
    if auto_reset:
        flags |= _abi.FLAG_AUTO_RESET
    if int(max_changes) > 255:   # a heat-map cell can count up to max_changes edits (pcgrl_env.py:137)
        flags |= _abi.FLAG_HEAT_U16
    if os.environ.get("PCGRL_FULL_STATS", "0") == "1":   # A/B switch: binary rollouts without the incremental statistics
        flags |= _abi.FLAG_FULL_STATS
    cfg.flags = flags
    cfg.solver_power = p["solver_power"]
    for i, v in enumerate(p["iparam"]):
        cfg.iparam[i] = v
    for i, v in enumerate(p["dparam"]):
        cfg.dparam[i] = v
    for i, v in enumerate(p["reward_weight"]):
        cfg.reward_weight[i] = v
    for i, v in enumerate(p["tile_prob"]):
        cfg.tile_prob[i] = v
    return cfg


def env_limits(prob, change_percentage=None, max_changes=None):
    """PcgrlEnv.__init__ / adjust_param limits (pcgrl_env.py:33-34, :107-110), computed from the
    problem's CURRENT width / height (quirk Q3: callers decide when this is evaluated)."""
    if change_percentage is not None:
        percentage = min(1, max(0, change_percentage))
        max_changes = max(int(percentage * prob._width * prob._height), 1)
    max_iterations = max_changes * prob._width * prob._height
    return max_changes, max_iterations

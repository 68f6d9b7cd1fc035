"""CPU-only tests of the host layer: plugin surface, config freezing, C-ABI exports, seeding."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import oracle
import util
import gym_pcgrl_b200 as pkg
from gym_pcgrl_b200 import PROBLEMS, REPRESENTATIONS, BatchedPcgrlEnv, _abi, _native, build as native_build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_registry_and_spaces():
    assert set(PROBLEMS) == {"binary", "ddave", "mdungeon", "sokoban", "zelda", "smb"}   # gym_pcgrl/envs/probs/__init__.py:9-16
    assert set(REPRESENTATIONS) == {"narrow", "turtle", "wide", "narrowcast", "narrowmulti", "turtlecast"}
    assert len(pkg.REGISTRY) == 36 and "zelda-turtle-v0" in pkg.REGISTRY and "sokoban-narrowmulti-v0" in pkg.REGISTRY
    smb = BatchedPcgrlEnv("smb", "narrow")
    assert smb.observation_space["map"].shape == (14, 114) and smb._max_changes == 319 and smb.action_space.n == 8
    assert smb.observation_space["heatmap"].dtype == np.uint16 and (smb.native_config.flags & _abi.FLAG_HEAT_U16)
    assert BatchedPcgrlEnv("zelda", "narrowmulti").action_space.nvec.tolist() == [9] * 9
    assert BatchedPcgrlEnv("binary", "narrowcast").action_space.nvec.tolist() == [3, 2]
    assert BatchedPcgrlEnv("ddave", "turtlecast").action_space.nvec.tolist() == [6, 7]
    env = BatchedPcgrlEnv("zelda", "turtle", num_envs=3)
    assert env.action_space.n == 4 + 8 and env.get_num_tiles() == 8 and env.get_border_tile() == 1
    assert env.observation_space["map"].shape == (7, 11) and set(env.observation_space.spaces) == {"pos", "map", "heatmap"}
    wide = BatchedPcgrlEnv("sokoban", "wide", num_envs=1)
    assert wide.action_space.nvec.tolist() == [5, 5, 5] and "pos" not in wide.observation_space.spaces
    assert BatchedPcgrlEnv("binary", "narrow").action_space.n == 3
    with pytest.raises(KeyError):
        BatchedPcgrlEnv("nope", "narrow")
    with pytest.raises(KeyError):
        BatchedPcgrlEnv("binary", "nope")


def test_adjust_param_ordering_quirk():
    """SURVEY.md Q3 / pcgrl_env.py:106-111: limits come from the size BEFORE the problem is resized."""
    env = BatchedPcgrlEnv("binary", "narrow")
    assert (env._max_changes, env._max_iterations) == (39, 7644)
    env.adjust_param(width=16, height=16)
    assert (env._max_changes, env._max_iterations) == (39, 7644)
    env = BatchedPcgrlEnv("binary", "narrow")
    env.adjust_param(width=16, height=16, change_percentage=0.2)
    env.adjust_param(width=16, height=16, change_percentage=0.2)
    assert (env._max_changes, env._max_iterations) == (51, 13056)
    for p, lim in (("zelda", (15, 1155)), ("ddave", (15, 1155)), ("mdungeon", (15, 1155)), ("sokoban", (5, 125))):
        e = BatchedPcgrlEnv(p, "wide")
        assert (e._max_changes, e._max_iterations) == lim


def test_problem_adjust_param_semantics():
    p = PROBLEMS["sokoban"]()
    p.adjust_param(max_targets=7, min_solution=3, probs={"empty": 0.9, "bogus": 1.0}, rewards={"ratio": 9, "bogus": 1})
    assert p._max_crates == 7 and p._target_solution == 3 and p._prob["empty"] == 0.9 and "bogus" not in p._prob
    assert p._rewards["ratio"] == 9 and "bogus" not in p._rewards
    z = PROBLEMS["zelda"]()
    z.adjust_param(width=5, target_path=3)
    assert (z._width, z._height, z._target_path) == (5, 7, 3)


def test_config_freeze_matches_header_layout():
    env = util.host_env("mdungeon-turtle-v0", dict(target_col_enemies=0.25, warp=True, solver_power=777))
    cfg = env.native_config
    assert cfg.problem == _abi.PROB_MDUNGEON and cfg.representation == _abi.REP_TURTLE
    assert (cfg.width, cfg.height, cfg.num_tiles, cfg.solver_power) == (7, 11, 8, 777)
    assert cfg.flags & _abi.FLAG_WARP and cfg.flags & _abi.FLAG_RANDOM_START and cfg.flags & _abi.FLAG_AUTO_RESET
    assert list(cfg.iparam)[:4] == [6, 2, 3, 20] and cfg.dparam[0] == 0.25
    assert list(cfg.reward_weight)[:9] == [3, 3, 2, 1, 1, 5, 2, 0.1, 1]
    b = util.host_env("binary-narrow-v0", dict(random_tile=False)).native_config
    assert not (b.flags & _abi.FLAG_RANDOM_TILE) and (b.flags & _abi.FLAG_RANDOM_PROBS)
    # the ctypes mirror must have the size the C compiler gives the struct
    assert C.sizeof(_abi.PcgrlConfig) == 9 * 4 + 7 * 4 + 2 * 8 + _abi.MAX_REWARD_TERMS * 8 + 8 * 8
    assert C.sizeof(_abi.PcgrlBuffers) == 17 * 8


def test_ctypes_mirrors_match_the_c_compiler(tmp_path):
    """Every struct of include/pcgrl_b200.h: sizeof and each field offset as gcc lays them out == the ctypes mirror."""
    import subprocess
    structs = {"pcgrl_config": _abi.PcgrlConfig, "pcgrl_buffers": _abi.PcgrlBuffers, "pcgrl_host_io": _abi.PcgrlHostIO,
               "pcgrl_host_rollout_io": _abi.PcgrlHostRolloutIO}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "pcgrl_b200.h"', 'int main(void) {']
    for cname, ct in structs.items():
        lines.append('  printf("%s sizeof %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in ct._fields_:
            lines.append('  printf("%s %s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    out = subprocess.check_output([str(exe)], text=True).split("\n")
    seen = 0
    for ln in out:
        if not ln.strip():
            continue
        cname, field, value = ln.split()
        ct = structs[cname]
        expect = C.sizeof(ct) if field == "sizeof" else getattr(ct, field).offset
        assert int(value) == expect, (cname, field, value, expect)
        seen += 1
    assert seen == sum(len(ct._fields_) + 1 for ct in structs.values())


def test_native_library_builds_loads_and_exports_every_symbol():
    native_build.build()
    lib = _native.lib()
    assert lib.pcgrl_abi_version() == _abi.ABI_VERSION
    header = open(os.path.join(ROOT, "include", "pcgrl_b200.h")).read()
    declared = set(re.findall(r"\b(pcgrl_[a-z0-9_]+)\s*\(", header))
    declared -= {"pcgrl_last_error"} if False else set()
    assert declared == set(_native.EXPORTS), declared ^ set(_native.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name


def test_config_validate_on_host():
    lib = _native.lib()
    ok = util.host_env("binary-narrow-v0", dict(width=16, height=16, change_percentage=0.2)).native_config
    assert lib.pcgrl_config_validate(C.byref(ok)) == 0
    assert lib.pcgrl_scratch_bytes(C.byref(ok), 4096) == 0
    big = BatchedPcgrlEnv("binary", "narrow")
    big.adjust_param(width=40, height=40)
    assert lib.pcgrl_config_validate(C.byref(big.native_config)) < 0
    assert b"width/height" in lib.pcgrl_last_error()
    sk = BatchedPcgrlEnv("sokoban", "wide").native_config
    assert lib.pcgrl_config_validate(C.byref(sk)) == 0 and lib.pcgrl_scratch_bytes(C.byref(sk), 2048) > 0
    sk2 = BatchedPcgrlEnv("sokoban", "wide")
    sk2.adjust_param(solver_power=100000)
    assert lib.pcgrl_config_validate(C.byref(sk2.native_config)) < 0


def test_product_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    env = BatchedPcgrlEnv("binary", "narrow", num_envs=2)
    with pytest.raises(_native.NativeError):
        env.reset()


def test_gym_style_seeding_spot_fact():
    """SURVEY.md App. B.3 spot fact (hashed seeding path, restated gym<=0.21 np_random): binary-narrow-v0,
    seed(42), first reset -> pos [9,12], regions 16, path-length 33, first map row 0 0 1 0 1 1 0 0 0 1 1 1 0 1."""
    env = BatchedPcgrlEnv("binary", "narrow", num_envs=1, seed=42, auto_reset=False)
    o = oracle.OracleEnv(env.native_config, 1)
    o.set_rng_states(env._pending_states)
    o.reset()
    assert o["pos"][0].tolist() == [9, 12]
    assert o["stats"][0, :2].tolist() == [16, 33]
    assert o["map"][0, 0].tolist() == [0, 0, 1, 0, 1, 1, 0, 0, 0, 1, 1, 1, 0, 1]


def test_env_offset_gives_shard_invariant_seeds():
    a = BatchedPcgrlEnv("binary", "narrow", num_envs=8, seed=5)
    b = BatchedPcgrlEnv("binary", "narrow", num_envs=4, seed=5, env_offset=4)
    np.testing.assert_array_equal(a._pending_states[4:], b._pending_states)


def test_phase_parallel_twist_equals_numpy():
    """The CUDA twist (csrc/pcgrl_device.cuh mt_twist_warp) runs numpy's mt19937_gen as three wide phases plus one
    scalar step; this is the same schedule in numpy, checked against RandomState's own twist."""
    def twist_phases(k):
        k = k.copy()
        UP, LO, A = np.uint32(0x80000000), np.uint32(0x7fffffff), np.uint32(0x9908b0df)

        def phase(lo, hi, off):
            idx = np.arange(lo, hi)
            a, b, c = k[idx], k[idx + 1], k[idx + off]          # all reads of the phase before any write
            y = (a & UP) | (b & LO)
            k[idx] = c ^ (y >> np.uint32(1)) ^ np.where(y & np.uint32(1), A, np.uint32(0))

        phase(0, 227, 397)
        phase(227, 454, -227)
        phase(454, 623, -227)
        y = (k[623] & UP) | (k[0] & LO)
        k[623] = k[396] ^ (y >> np.uint32(1)) ^ (A if y & np.uint32(1) else np.uint32(0))
        return k

    for seed in (0, 1, 123, 2 ** 31 - 1):
        r = np.random.RandomState(seed)
        before = r.get_state()[1].astype(np.uint32)
        r.random_sample(1)
        np.testing.assert_array_equal(twist_phases(before), r.get_state()[1].astype(np.uint32))

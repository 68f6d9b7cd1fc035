"""Cases of the opt-in multi-env-per-warp rollout kernel; run by test_gpu_parity.py::test_packed_rollout_kernel_matches_oracle
in a child process with PCGRL_PACKED=1 (the library reads the switch once per process)."""
import os

import numpy as np
import pytest

import oracle
import util
from test_gpu_parity import assert_state_equal, random_actions, t2n

pytestmark = pytest.mark.gpu


PACKED_CASES = [
    # (env id, kwargs, n envs, T, rounds) -- binary rollouts with T >= 8 run k_rollout_packed_binary<16 | 8>
    ("binary-narrow-v0", dict(width=16, height=16, change_percentage=0.2), 513, 128, 3),
    ("binary-narrow-v0", {}, 300, 400, 2),                                            # 14x14 default, long: MT19937 twists inside
    ("binary-turtle-v0", dict(width=11, height=11, change_percentage=0.2), 129, 96, 2),
    ("binary-wide-v0", dict(width=16, height=16, change_percentage=0.1), 64, 64, 3),
    ("binary-narrow-v0", dict(width=8, height=8, change_percentage=0.3), 257, 200, 2),     # four envs per warp
    ("binary-wide-v0", dict(width=20, height=7, change_percentage=0.3), 70, 80, 2),
    ("binary-narrow-v0", dict(width=3, height=2, change_percentage=0.5, random_tile=False), 19, 50, 2),
    ("binary-turtle-v0", dict(width=32, height=16, change_percentage=0.05, warp=True), 33, 120, 2),
    ("binary-narrow-v0", dict(width=16, height=16, change_percentage=0.2, random_start=False, random_probs=False), 7, 64, 3),
    ("binary-narrow-v0", dict(width=12, height=9, change_percentage=0.2), 1, 40, 2),
]


@pytest.mark.parametrize("case", PACKED_CASES, ids=["%s-%d" % (c[0], i) for i, c in enumerate(PACKED_CASES)])
def test_packed_rollout_matches_oracle(case):
    """The multi-env-per-warp rollout kernel (run-ahead + packed wave engine, csrc/pcgrl_packed.cuh; opt-in through
    PCGRL_PACKED=1, which the library reads once per process -- see test_packed_kernel_is_selected) against the
    oracle: every step's reward / done, the complete state after each fragment, the RNG streams at the end."""
    import torch
    env_id, kwargs, n, T, rounds = case
    env = util.host_env(env_id, kwargs, num_envs=n, device="cuda")
    states = np.stack([util.randomstate_words(4000 + i) for i in range(n)])
    env.set_rng_states(states)
    env.reset()
    ref = oracle.OracleEnv(env.native_config, n, threads=8)
    ref.set_rng_states(states)
    ref.reset()
    wide = env_id.split("-")[1] == "wide"
    arng = np.random.RandomState(21)
    ndone = 0
    for rnd in range(rounds):
        acts = np.stack([random_actions(env, arng, n) for _ in range(T)])
        rew, done = env.rollout(torch.from_numpy(acts).cuda())
        rew, done = t2n(rew), t2n(done).astype(np.uint8)
        for k in range(T):
            ref.step(acts[k])
            ctx = "%s round %d step %d" % (env_id, rnd, k)
            np.testing.assert_array_equal(rew[k], ref["reward"], err_msg=ctx + " reward")
            np.testing.assert_array_equal(done[k], ref["done"], err_msg=ctx + " done")
            ndone += int(ref["done"].sum())
        assert_state_equal(env, ref, 2, "%s round %d" % (env_id, rnd), wide)
        np.testing.assert_array_equal(t2n(env._tens["info_stats"])[:, :3], ref["info_stats"][:, :3])
        np.testing.assert_array_equal(t2n(env._tens["info_stats"])[:, 14:], ref["info_stats"][:, 14:])
        np.testing.assert_array_equal(t2n(env._tens["reward"]), ref["reward"])
        np.testing.assert_array_equal(t2n(env._tens["done"]), ref["done"])
        # a few single steps in between: the one-env-per-warp kernel continues from the packed kernel's state
        for k in range(3):
            a = random_actions(env, arng, n)
            env.step(torch.from_numpy(a).cuda())
            ref.step(a)
        assert_state_equal(env, ref, 2, "%s round %d + steps" % (env_id, rnd), wide)
    np.testing.assert_array_equal(t2n(env._tens["rng"]).view(np.uint32), ref["rng"], err_msg="rng state")
    np.testing.assert_array_equal(t2n(env._tens["tile_prob"]), ref["tile_prob"], err_msg="tile_prob")
    assert ndone > 0 or n < 8


def test_packed_kernel_is_selected():
    """The switch must be set before the library decides (first binary rollout of the process)."""
    assert os.environ.get("PCGRL_PACKED") == "1"

"""Seeding helpers.

``np_random`` restates gym <= 0.21 ``gym.utils.seeding.np_random`` (used by
gym_pcgrl/envs/reps/representation.py:29 and probs/problem.py:35).  gym itself is absent from the
build container, so this piece is "parity unpinned" (SURVEY.md App. C.1); parity harnesses inject
explicit ``RandomState`` objects / MT19937 states instead (``BatchedPcgrlEnv.set_rng_states``).
"""
import hashlib
import os

import numpy as np


def hash_seed_words(seed):
    """seed -> list of 32-bit words fed to MT19937 ``init_by_array``."""
    h = hashlib.sha512(str(int(seed)).encode("utf8")).digest()[:8]
    big = int.from_bytes(h, "little")
    words = []
    while big > 0:
        big, w = divmod(big, 2 ** 32)
        words.append(w)
    return words or [0]


def create_seed(seed=None):
    if seed is None:
        return int.from_bytes(os.urandom(8), "little")
    seed = int(seed)
    if seed < 0:
        raise ValueError("Seed must be a non-negative integer or omitted, not %r" % (seed,))
    return seed % (2 ** 64)


def np_random(seed=None):
    seed = create_seed(seed)
    rng = np.random.RandomState()
    rng.seed(hash_seed_words(seed))
    return rng, seed


def mt_state_words(rng):
    """numpy RandomState -> uint32[625] (624 key words + position), the device RNG layout."""
    _, key, pos = rng.get_state()[:3]
    out = np.empty(625, np.uint32)
    out[:624] = key
    out[624] = pos
    return out

"""Run the UNMODIFIED reference (amidos2006/gym-pcgrl, /root/reference) in the build container.

TEST INFRASTRUCTURE ONLY.  This module exists so that ``make_golden.py`` can execute the
reference's own Python ``PcgrlEnv`` and record golden vectors under ``tests/golden/``.  It is
never imported by the product (``gym_pcgrl_b200``), by ``bench.py`` or by the ``-m gpu`` tests:
``/root/reference`` does not exist on the GPU box.

Two shims are needed, neither touches the reference sources:

1. ``gym`` is not installed -> a minimal stand-in exposing exactly what the reference touches
   (``gym.Env``, ``gym.Wrapper``, ``gym.make``/``register``, ``gym.spaces.*``,
   ``gym.utils.seeding.np_random``).
2. numpy 2.x rejects ``[0, 1][np.bool_]`` (used by every ``Representation.update``), so the
   representation's ``_map`` is viewed as an ndarray subclass whose scalar reads are Python ints.
"""
import hashlib
import importlib
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("PCGRL_REFERENCE_ROOT", "/root/reference")


# ----------------------------------------------------------------------------- gym stand-in
class _Space:
    pass


class Discrete(_Space):
    def __init__(self, n):
        self.n = int(n)
        self.shape = ()
        self.dtype = np.int64


class MultiDiscrete(_Space):
    def __init__(self, nvec):
        self.nvec = np.asarray(nvec, dtype=np.int64)
        self.shape = self.nvec.shape
        self.dtype = np.int64


class Box(_Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        if shape is None:
            shape = np.asarray(low).shape
        self.shape = tuple(shape)
        self.dtype = np.dtype(dtype)
        self.low = np.broadcast_to(np.asarray(low), self.shape).astype(self.dtype)   # gym keeps the Box dtype
        self.high = np.broadcast_to(np.asarray(high), self.shape).astype(self.dtype)


class Dict(_Space):
    def __init__(self, spaces=None):
        self.spaces = dict(spaces or {})

    def __getitem__(self, k):
        return self.spaces[k]


class Env:
    metadata = {}

    @property
    def unwrapped(self):
        return self


class Wrapper(Env):
    def __init__(self, env):
        self.env = env
        if not hasattr(self, "action_space") or self.__dict__.get("action_space") is None:
            self.action_space = env.action_space
        if "observation_space" not in self.__dict__:
            self.observation_space = env.observation_space

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return getattr(self.__dict__["env"], name)

    @property
    def unwrapped(self):
        return self.env.unwrapped


_REGISTRY = {}


def register(id, entry_point, kwargs=None, **_):
    _REGISTRY[id] = (entry_point, dict(kwargs or {}))


def make(id, **extra):
    entry_point, kwargs = _REGISTRY[id]
    mod, cls = entry_point.split(":")
    return getattr(importlib.import_module(mod), cls)(**{**kwargs, **extra})


def np_random(seed=None):
    """gym <= 0.21 ``seeding.np_random`` restated (parity unpinned:
    gym itself is absent from the container)."""
    if seed is None:
        seed = int.from_bytes(os.urandom(8), "little")
    seed = int(seed) % (2 ** 64) if seed >= 0 else seed
    h = hashlib.sha512(str(seed).encode("utf8")).digest()[:8]
    big = int.from_bytes(h, "little")
    words = []
    while big > 0:
        big, w = divmod(big, 2 ** 32)
        words.append(w)
    rng = np.random.RandomState()
    rng.seed(words or [0])
    return rng, seed


def install():
    """Install the stand-in ``gym`` package and import the reference ``gym_pcgrl``."""
    if "gym_pcgrl" in sys.modules:
        return sys.modules["gym_pcgrl"]
    gym = types.ModuleType("gym")
    gym.Env, gym.Wrapper, gym.make = Env, Wrapper, make
    spaces = types.ModuleType("gym.spaces")
    spaces.Discrete, spaces.MultiDiscrete, spaces.Box, spaces.Dict = Discrete, MultiDiscrete, Box, Dict
    envs = types.ModuleType("gym.envs")
    registration = types.ModuleType("gym.envs.registration")
    registration.register = register
    utils = types.ModuleType("gym.utils")
    seeding = types.ModuleType("gym.utils.seeding")
    seeding.np_random = np_random
    gym.spaces, gym.envs, gym.utils = spaces, envs, utils
    envs.registration, utils.seeding = registration, seeding
    sys.modules.update({
        "gym": gym, "gym.spaces": spaces, "gym.envs": envs,
        "gym.envs.registration": registration, "gym.utils": utils, "gym.utils.seeding": seeding,
    })
    if not os.path.isdir(REFERENCE_ROOT):
        raise RuntimeError("reference tree not found at %s (golden generation only works in the "
                           "build container)" % REFERENCE_ROOT)
    sys.path.insert(0, REFERENCE_ROOT)
    import gym_pcgrl  # noqa: F401  (registers the 36 ids)
    _patch_numpy2()
    return gym_pcgrl


class PyInt(int):
    """int whose != / == against numpy scalars give Python bools (so ``[0, 1][tile != action]`` indexes)."""

    def __ne__(self, other):
        return bool(int(self) != int(other))

    def __eq__(self, other):
        return bool(int(self) == int(other))

    __hash__ = int.__hash__


class PyScalarMap(np.ndarray):
    """ndarray view whose scalar reads are Python ints (numpy-2 workaround)."""

    def __getitem__(self, idx):
        r = np.ndarray.__getitem__(self, idx)
        if isinstance(r, np.generic):
            return PyInt(r)
        return r


def _patch_numpy2():
    from gym_pcgrl.envs.reps import REPRESENTATIONS
    for cls in set(REPRESENTATIONS.values()):
        if getattr(cls.reset, "_pcgrl_shim", False):
            continue
        orig = cls.reset

        def reset(self, width, height, prob, _orig=orig):
            _orig(self, width, height, prob)
            if not isinstance(self._map, PyScalarMap):
                self._map = self._map.view(PyScalarMap)

        reset._pcgrl_shim = True
        cls.reset = reset

"""Batched observation / action wrappers (mirror of gym_pcgrl/wrappers.py, the "next" row f1 of the scope table).

The reference stacks gym.Wrappers per environment (Cropped -> OneHotEncoding -> ToImage, and ActionMap); here the
whole stack is ONE fused sm_100a kernel over the map batch (csrc/pcgrl_wrappers.cuh, pcgrl_obs_image) that writes the
policy input tensor [N, S, S, C] directly, and a tiny kernel (pcgrl_action_map) that decodes flat ActionMap indices.

* ``CroppedImagePCGRLWrapper(game, crop_size, num_envs=..., **kwargs)``  -- wrappers.py:215-230 (narrow / turtle)
* ``ActionMapImagePCGRLWrapper(game, num_envs=..., **kwargs)``           -- wrappers.py:234-248 (wide; cursor
  representations are supported too, with the reference's "write the current tile value" semantics, :146-150)

Observations are uint8 by default (float32 on request); the reference's values (float64 one-hot / uint8 tiles) are
equal in every dtype.
"""
import ctypes as C

import numpy as np

from . import _native, spaces
from .envs.pcgrl_env import BatchedPcgrlEnv


def _make_env(game, num_envs, device, **env_kwargs):
    if isinstance(game, str):
        from . import REGISTRY
        spec = REGISTRY[game]
        return BatchedPcgrlEnv(spec["prob"], spec["rep"], num_envs=num_envs, device=device, **env_kwargs), game
    return game, "%s-%s-v0" % (game._prob.name, game._rep.name)


class _ImageWrapper:
    """Shared machinery: owns the output tensor and launches pcgrl_obs_image after reset / step."""

    def __init__(self, game, crop_size, num_envs=1, device="cuda", out_dtype="uint8", env_kwargs=None, **kwargs):
        self.pcgrl_env, self.game = _make_env(game, num_envs, device, **(env_kwargs or {}))
        self.pcgrl_env.adjust_param(**kwargs)          # once, like the reference (wrappers.py:218,237)
        self.env = self.pcgrl_env
        self.crop_size = int(crop_size)
        self.one_hot = 'binary' not in self.game       # wrappers.py:222,244
        self.out_dtype = out_dtype
        self._out = None
        p = self.pcgrl_env._prob
        self._update_space()
        assert out_dtype in ("uint8", "float32")
        if self.crop_size:
            assert self.pcgrl_env._rep.name != "wide", 'This wrapper only works for representations thave have a position'

    def _update_space(self):
        p = self.pcgrl_env._prob
        s_h = self.crop_size or p._height
        s_w = self.crop_size or p._width
        c = self.pcgrl_env.get_num_tiles() if self.one_hot else 1
        self.shape = (s_h, s_w, c)
        self.observation_space = spaces.Box(low=0, high=1 if self.one_hot else self.pcgrl_env.get_num_tiles() - 1,
                                            shape=self.shape, dtype=np.dtype(self.out_dtype))

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return getattr(self.__dict__["pcgrl_env"], name)

    def adjust_param(self, **kwargs):
        self.pcgrl_env.adjust_param(**kwargs)
        self._update_space()
        self._out = None

    def _image(self):
        import torch
        env = self.pcgrl_env
        self._update_space()
        n = env.num_envs
        if self._out is None or tuple(self._out.shape) != (n,) + self.shape:
            self._out = torch.empty((n,) + self.shape, dtype=getattr(torch, self.out_dtype), device=env._dev)
        t = env._tens
        with torch.cuda.device(env._dev):
            _native.check(_native.lib().pcgrl_obs_image(
                C.byref(env.native_config), t["map"].data_ptr(), t["pos"].data_ptr(), self._out.data_ptr(), n,
                self.crop_size, env.get_border_tile(), int(self.one_hot), 0 if self.out_dtype == "uint8" else 1,
                _native.stream_ptr(env._dev)), "pcgrl_obs_image")
        return self._out

    def reset(self, mask=None):
        self.pcgrl_env.reset(mask)
        return self._image()


class CroppedImagePCGRLWrapper(_ImageWrapper):
    """Cropped(crop_size, pad = border tile) -> OneHotEncoding (unless binary) -> ToImage, fused and batched."""

    def __init__(self, game, crop_size, num_envs=1, device="cuda", out_dtype="uint8", env_kwargs=None, **kwargs):
        super().__init__(game, crop_size, num_envs, device, out_dtype, env_kwargs, **kwargs)
        self.action_space = self.pcgrl_env.action_space

    def step(self, actions):
        _, reward, done, info = self.pcgrl_env.step(actions)
        return self._image(), reward, done, info


class ActionMapImagePCGRLWrapper(_ImageWrapper):
    """ActionMap -> OneHotEncoding (unless binary) -> ToImage: the action is a flat index over (h, w, num_tiles)."""

    def __init__(self, game, num_envs=1, device="cuda", out_dtype="uint8", env_kwargs=None, **kwargs):
        super().__init__(game, 0, num_envs, device, out_dtype, env_kwargs, **kwargs)
        assert self.pcgrl_env._rep.name in ("wide", "narrow", "turtle"), "ActionMap supports narrow / turtle / wide"
        self._actions = None
        self._update_action_space()

    def _update_action_space(self):
        p = self.pcgrl_env._prob
        self.h, self.w, self.dim = p._height, p._width, self.pcgrl_env.get_num_tiles()
        self.action_space = spaces.Discrete(self.h * self.w * self.dim)     # wrappers.py:133

    def adjust_param(self, **kwargs):
        super().adjust_param(**kwargs)
        self._update_action_space()

    def step(self, flat_actions):
        import torch
        env = self.pcgrl_env
        a = torch.as_tensor(flat_actions).to(device=env._dev, dtype=torch.int32).contiguous()
        n = env.num_envs
        if self._actions is None or self._actions.numel() != n * env._adim:
            self._actions = torch.empty(n * env._adim, dtype=torch.int32, device=env._dev)
        with torch.cuda.device(env._dev):
            _native.check(_native.lib().pcgrl_action_map(C.byref(env.native_config), C.byref(env._cbufs), a.data_ptr(),
                                                         self._actions.data_ptr(), n, _native.stream_ptr(env._dev)),
                          "pcgrl_action_map")
        _, reward, done, info = env.step(self._actions.view(n, env._adim) if env._adim > 1 else self._actions)
        return self._image(), reward, done, info

"""The reference's policy networks (model.py:9-166) as torch modules -- SURVEY.md 8f row f4, the consumer side.

``Cnn1`` / ``Cnn2`` (feature extractors of CustomPolicySmallMap / BigMap, model.py:9-23,161-166) and ``FullyConv1`` /
``FullyConv2`` (FullyConvPolicySmallMap / BigMap, model.py:25-77,106-159), with stable-baselines' conventions: NHWC
observations cast to float (no /255: the wrapped spaces are not [0, 255] images), VALID padding unless stated, orthogonal
initialisation with the given ``init_scale`` and zero biases, ``conv_to_fc`` flattening in (h, w, c) order (which is the
order ActionMap unravels flat actions in, wrappers.py:139-141).  Layers run through cuDNN / cuBLAS (library kernels);
the one hand-written contraction of this repo is csrc/pcgrl_linear.cu (opt-in for the 512-unit layer of Cnn1 / Cnn2).
"""
import math
import os

import torch
import torch.nn as nn
import torch.nn.functional as F


def _ortho(layer, scale):
    nn.init.orthogonal_(layer.weight, gain=scale)
    nn.init.zeros_(layer.bias)
    return layer


def _conv(cin, cout, k, stride, pad_same, scale=math.sqrt(2)):
    return _ortho(nn.Conv2d(cin, cout, k, stride=stride, padding=(k // 2 if pad_same else 0)), scale)


class _CnnBase(nn.Module):
    strides = (1, 1, 1)

    def __init__(self, obs_shape):
        super().__init__()
        h, w, c = obs_shape
        s = self.strides
        self.c1, self.c2, self.c3 = _conv(c, 32, 3, s[0], False), _conv(32, 64, 3, s[1], False), _conv(64, 64, 3, s[2], False)
        for st in s:
            h, w = (h - 3) // st + 1, (w - 3) // st + 1
        if h < 1 or w < 1:
            raise ValueError("observation %s is too small for three VALID 3x3 convolutions" % (obs_shape,))
        self.fc1 = _ortho(nn.Linear(h * w * 64, 512), math.sqrt(2))
        self.out_features = 512
        # inference through the hand-written tcgen05 kernel (csrc/pcgrl_linear.cu); training keeps the autograd layer
        self.use_tcgen05_fc = os.environ.get("PCGRL_TCGEN05_FC", "0") == "1"

    def forward(self, obs):                                   # obs: [N, H, W, C] any dtype
        x = obs.permute(0, 3, 1, 2).float()
        x = F.relu(self.c3(F.relu(self.c2(F.relu(self.c1(x))))))
        x = x.permute(0, 2, 3, 1).flatten(1)                  # conv_to_fc: (h, w, c) order
        if self.use_tcgen05_fc and x.is_cuda and not torch.is_grad_enabled():
            from . import _native
            return _native.linear_bf16(x, self.fc1.weight, self.fc1.bias, relu=True)
        return F.relu(self.fc1(x))


class Cnn1(_CnnBase):      # model.py:9-15
    strides = (1, 1, 1)


class Cnn2(_CnnBase):      # model.py:17-23
    strides = (2, 2, 1)


class _FullyConvBase(nn.Module):
    value_strides = (2,)
    value_scales = (math.sqrt(2),)

    def __init__(self, obs_shape, n_tools):
        super().__init__()
        h, w, c = obs_shape
        chans = [c, 32, 64, 64, 64, 64, 64, 64, n_tools]
        self.body = nn.ModuleList([_conv(chans[i], chans[i + 1], 3, 1, True) for i in range(8)])      # c1 .. c8, SAME
        vin, vs = n_tools, []
        for st, sc in zip(self.value_strides, self.value_scales):                                         # v1 (, v2): VALID, stride 2
            vs.append(_conv(vin, 64, 3, st, False, sc))
            vin = 64
            h, w = (h - 3) // st + 1, (w - 3) // st + 1
        vs.append(_conv(64, 64, 1, 1, False))                                                             # v4: 1x1
        self.value = nn.ModuleList(vs)
        self.vf_features = h * w * 64

    def forward(self, obs):
        x = obs.permute(0, 3, 1, 2).float()
        for layer in self.body:
            x = F.relu(layer(x))
        act = x.permute(0, 2, 3, 1).flatten(1)                 # logits over (h, w, tool), model.py:45
        v = x
        for layer in self.value:
            v = F.relu(layer(v))
        return act, v.permute(0, 2, 3, 1).flatten(1)


class FullyConv1(_FullyConvBase):   # model.py:25-50
    value_strides = (2,)
    value_scales = (math.sqrt(2),)


class FullyConv2(_FullyConvBase):   # model.py:52-77 (v2 uses init_scale sqrt(3), sic)
    value_strides = (2, 2)
    value_scales = (math.sqrt(2), math.sqrt(3))


class ActorCritic(nn.Module):
    """The four policy classes of model.py:106-166 behind one interface: ``forward(obs) -> (logits [N, A], value [N])``.

    kind: "CustomPolicyBigMap" (Cnn2) | "CustomPolicySmallMap" (Cnn1)  -- FeedForwardPolicy: pi = linear(512, A, 0.01),
          vf = linear(512, 1);
          "FullyConvPolicyBigMap" (FullyConv2) | "FullyConvPolicySmallMap" (FullyConv1) -- the logits ARE the flattened
          conv output (NoDenseCategoricalProbabilityDistributionType), vf = linear(vf_latent, 1)."""

    def __init__(self, kind, obs_shape, n_actions):
        super().__init__()
        self.kind = kind
        if kind in ("CustomPolicyBigMap", "CustomPolicySmallMap"):
            self.extractor = (Cnn2 if kind == "CustomPolicyBigMap" else Cnn1)(obs_shape)
            self.pi = _ortho(nn.Linear(512, n_actions), 0.01)
            self.vf = _ortho(nn.Linear(512, 1), 1.0)
            self.fully_conv = False
        elif kind in ("FullyConvPolicyBigMap", "FullyConvPolicySmallMap"):
            n_tools = n_actions // (obs_shape[0] * obs_shape[1])          # model.py:109
            if n_tools * obs_shape[0] * obs_shape[1] != n_actions:
                raise ValueError("FullyConv policies need an ActionMap action space over (h, w, tools)")
            self.extractor = (FullyConv2 if kind == "FullyConvPolicyBigMap" else FullyConv1)(obs_shape, n_tools)
            self.vf = _ortho(nn.Linear(self.extractor.vf_features, 1), 1.0)
            self.fully_conv = True
        else:
            raise KeyError(kind)

    def forward(self, obs):
        if self.fully_conv:
            logits, vlat = self.extractor(obs)
            return logits, self.vf(vlat).squeeze(-1)
        feat = self.extractor(obs)
        return self.pi(feat), self.vf(feat).squeeze(-1)


def policy_for(game, representation):
    """train.py:51-62: which policy class the reference trains for a (game, representation) pair."""
    if representation == "wide":
        return "FullyConvPolicySmallMap" if game == "sokoban" else "FullyConvPolicyBigMap"
    return "CustomPolicySmallMap" if game == "sokoban" else "CustomPolicyBigMap"


CROPPED_SIZE = {"binary": 28, "zelda": 22, "sokoban": 10}   # train.py:63-68 (default 28, utils.py:52)

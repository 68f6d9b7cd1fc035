"""The policy forward pass of model.py on this repo's own kernels.

``NativePolicy(net)`` takes an ``ActorCritic`` (models.py: the reference's four policy classes) and runs its inference
path -- observation [N, H, W, C] uint8 from the fused wrapper kernel -> logits, value -- without cuDNN / cuBLAS.  Strided /
VALID / few-channel convolutions are ``pcgrl_im2col`` (NHWC patches, bf16) followed by ``pcgrl_linear_bf16_ex``
(csrc/pcgrl_linear.cu: TMA -> tcgen05.mma -> TMEM, bias + ReLU fused in the epilogue, bf16 activations written in the NHWC
layout the next layer reads); the 3 x 3 SAME layers of the fully convolutional policies (c2..c8) are an implicit GEMM on
zero-bordered NHWC buffers (``pcgrl_conv3x3_bf16``: no patch matrix, weights resident in shared memory); the dense layers
are the same GEMM kernel.  torch only owns the buffers (and slices the padded head columns).
Weights are snapshotted in bf16 at construction (``refresh()`` after an optimiser step); accumulation is fp32.
"""
import torch

from . import _native
from .models import ActorCritic


def _pad8(v):
    return (v + 7) // 8 * 8


class _Gemm:
    """y = act(x @ w.T + b) with w [Nout_pad, Kpad] bf16 (zero padded), through pcgrl_linear_bf16_ex."""

    def __init__(self, weight2d, bias, relu, out_bf16, device):
        n_out, k = weight2d.shape
        self.n_out, self.k, self.kpad = n_out, k, _pad8(k)
        self.npad = _pad8(n_out)
        w = torch.zeros((self.npad, self.kpad), dtype=torch.bfloat16, device=device)
        w[:n_out, :k] = weight2d.to(device=device, dtype=torch.bfloat16)
        b = torch.zeros(self.npad, dtype=torch.float32, device=device)
        b[:n_out] = bias.to(device=device, dtype=torch.float32)
        self.w, self.b, self.relu, self.out_bf16 = w, b, relu, out_bf16
        self.out = None

    def __call__(self, x, m, stream):
        if self.out is None or self.out.shape[0] != m:
            self.out = torch.empty((m, self.npad), dtype=torch.bfloat16 if self.out_bf16 else torch.float32, device=self.w.device)
        rc = _native.lib().pcgrl_linear_bf16_ex(x.data_ptr(), self.w.data_ptr(), self.b.data_ptr(), self.out.data_ptr(), m, self.npad,
                                                self.kpad, 1 if self.relu else 0, 1 if self.out_bf16 else 0, stream)
        if rc:
            raise _native.NativeError("pcgrl_linear_bf16_ex failed (rc=%d): %s" % (rc, _native.lib().pcgrl_linear_last_error().decode()))
        return self.out


class _Conv:
    """conv(k x k, stride, VALID | SAME) + bias + ReLU on NHWC activations: im2col + GEMM."""

    def __init__(self, conv, device, relu=True):
        cout, cin, kh, kw = conv.weight.shape
        self.cin, self.cout, self.ks, self.stride, self.pad = cin, cout, kh, conv.stride[0], conv.padding[0]
        w2 = conv.weight.detach().permute(0, 2, 3, 1).reshape(cout, kh * kw * cin)      # k = (ky * KW + kx) * C + c
        self.gemm = _Gemm(w2, conv.bias.detach(), relu, True, device)
        self.cols = None

    def __call__(self, x, in_bf16, n, h, w, c, stream):
        assert c == self.cin, (c, self.cin)
        ho = (h + 2 * self.pad - self.ks) // self.stride + 1
        wo = (w + 2 * self.pad - self.ks) // self.stride + 1
        m = n * ho * wo
        if self.cols is None or self.cols.shape[0] != m:
            self.cols = torch.empty((m, self.gemm.kpad), dtype=torch.bfloat16, device=self.gemm.w.device)
        rc = _native.lib().pcgrl_im2col(x.data_ptr(), 1 if in_bf16 else 0, self.cols.data_ptr(), n, h, w, c, self.ks, self.stride,
                                        self.pad, self.gemm.kpad, stream)
        if rc:
            raise _native.NativeError("pcgrl_im2col failed (rc=%d): %s" % (rc, _native.lib().pcgrl_linear_last_error().decode()))
        y = self.gemm(self.cols, m, stream)          # [n * ho * wo, cout_pad] == NHWC with cout_pad channels
        return y, ho, wo, self.gemm.npad


def _check(rc, what):
    if rc:
        raise _native.NativeError("%s failed (rc=%d): %s" % (what, rc, _native.lib().pcgrl_linear_last_error().decode()))


class _ConvFirstPadded:
    """The first SAME convolution of a fully convolutional policy (few input channels: im2col + GEMM), writing its bf16
    activations into the zero-bordered [n, H + 2, W + 2, 64] buffer the implicit-GEMM layers read."""

    def __init__(self, conv, device):
        cout, cin, kh, kw = conv.weight.shape
        assert kh == 3 and conv.stride[0] == 1 and conv.padding[0] == 1 and cout <= 64
        self.cin, self.ks = cin, kh
        w2 = torch.zeros((64, kh * kw * cin), dtype=conv.weight.dtype, device=conv.weight.device)   # channels cout..63 stay zero
        w2[:cout] = conv.weight.detach().permute(0, 2, 3, 1).reshape(cout, kh * kw * cin)
        b = torch.zeros(64, dtype=conv.bias.dtype, device=conv.bias.device)
        b[:cout] = conv.bias.detach()
        self.gemm = _Gemm(w2, b, True, True, device)
        self.cols = None

    def __call__(self, obs, out_padded, n, h, w, c, stream):
        assert c == self.cin
        m = n * h * w
        if self.cols is None or self.cols.shape[0] != m:
            self.cols = torch.empty((m, self.gemm.kpad), dtype=torch.bfloat16, device=self.gemm.w.device)
        L = _native.lib()
        _check(L.pcgrl_im2col(obs.data_ptr(), 0, self.cols.data_ptr(), n, h, w, c, 3, 1, 1, self.gemm.kpad, stream), "pcgrl_im2col")
        g = self.gemm
        _check(L.pcgrl_linear_bf16_pad(self.cols.data_ptr(), g.w.data_ptr(), g.b.data_ptr(), out_padded.data_ptr(), m, 64, g.kpad, 1,
                                       h, w, stream), "pcgrl_linear_bf16_pad")


class _Conv3x3Implicit:
    """3 x 3 / stride 1 / SAME conv + bias + ReLU as an implicit GEMM on zero-bordered NHWC buffers (pcgrl_conv3x3_bf16)."""

    def __init__(self, conv, device):
        cout, cin, kh, kw = conv.weight.shape
        assert kh == 3 and kw == 3 and conv.stride[0] == 1 and conv.padding[0] == 1 and cin <= 64 and cout <= 64
        self.cout, self.npad = cout, (64 if cout > 32 else _pad8(cout))
        w = torch.zeros((self.npad, 3, 3, 64), dtype=torch.float32, device=device)                   # input channels padded to 64
        w[:cout, :, :, :cin] = conv.weight.detach().permute(0, 2, 3, 1).to(device=device, dtype=torch.float32)
        self.w = w.reshape(self.npad, 9 * 64).to(torch.bfloat16).contiguous()
        self.b = torch.zeros(self.npad, dtype=torch.float32, device=device)
        self.b[:cout] = conv.bias.detach().to(device=device, dtype=torch.float32)

    def __call__(self, x_padded, y_padded, n, h, w, stream):
        _check(_native.lib().pcgrl_conv3x3_bf16(x_padded.data_ptr(), self.w.data_ptr(), self.b.data_ptr(), y_padded.data_ptr(), n, h, w,
                                                64, self.npad, 1, stream), "pcgrl_conv3x3_bf16")


class NativePolicy:
    def __init__(self, net):
        if not isinstance(net, ActorCritic):
            raise TypeError("NativePolicy wraps a gym_pcgrl_b200.models.ActorCritic")
        self.net = net
        self.refresh()

    def refresh(self):
        """Re-snapshot the weights (call after optimiser steps)."""
        net = self.net
        dev = next(net.parameters()).device
        self.device = _native.require_cuda(dev)
        ex = net.extractor
        if net.fully_conv:
            # c1: im2col + GEMM into a zero-bordered buffer; c2..c8: implicit GEMM, ping-pong between two such buffers
            self.first = _ConvFirstPadded(ex.body[0], dev)
            self.body = [_Conv3x3Implicit(c, dev) for c in ex.body[1:]]
            self.head_pad = self.body[-1].npad
            self.value_convs = [_Conv(c, dev) for c in ex.value]
            v1 = ex.value[0]     # reads the c8 map: its input channels are padded like the c8 output (zero weights)
            wv = torch.zeros((v1.out_channels, self.head_pad, v1.kernel_size[0], v1.kernel_size[1]), dtype=v1.weight.dtype, device=v1.weight.device)
            wv[:, :v1.in_channels] = v1.weight.detach()
            self.value_convs[0].cin = self.head_pad
            self.value_convs[0].pad = -1     # VALID over the interior of the zero-bordered c8 map
            self.value_convs[0].gemm = _Gemm(wv.permute(0, 2, 3, 1).reshape(v1.out_channels, -1), v1.bias.detach(), True, True, dev)
            self.bufs = None
            self.vf = _Gemm(net.vf.weight.detach(), net.vf.bias.detach(), False, False, dev)
            self.n_tools = ex.body[-1].out_channels
        else:
            self.convs = [_Conv(c, dev) for c in (ex.c1, ex.c2, ex.c3)]
            # fc1 reads the (h, w, c) flattening of conv3's NHWC output; its channel count is not padded (64)
            self.fc1 = _Gemm(ex.fc1.weight.detach(), ex.fc1.bias.detach(), True, True, dev)
            heads_w = torch.cat([net.pi.weight.detach(), net.vf.weight.detach()], dim=0)       # [A + 1, 512]: one GEMM for both heads
            heads_b = torch.cat([net.pi.bias.detach(), net.vf.bias.detach()], dim=0)
            self.heads = _Gemm(heads_w, heads_b, False, False, dev)
            self.n_actions = net.pi.out_features

    @torch.no_grad()
    def forward(self, obs):
        """obs: uint8 CUDA tensor [N, H, W, C] (the wrapper's image) -> (logits float32 [N, A], value float32 [N])."""
        obs = obs.contiguous()
        assert obs.dtype == torch.uint8 and obs.is_cuda
        n, h, w, c = obs.shape
        with torch.cuda.device(self.device):
            stream = _native.stream_ptr(self.device)
            x, in_bf16 = obs, False
            if not self.net.fully_conv:
                for conv in self.convs:
                    x, h, w, c = conv(x, in_bf16, n, h, w, c, stream)
                    in_bf16 = True
                feat = self.fc1(x, n, stream)                       # x viewed as [n, h * w * 64]: contiguous, no copy
                out = self.heads(feat, n, stream)                   # [n, pad4(A + 1)] fp32
                return out[:, :self.n_actions], out[:, self.n_actions]
            if self.bufs is None or self.bufs[0].shape[:3] != (n, h + 2, w + 2):
                self.bufs = [torch.zeros((n, h + 2, w + 2, 64), dtype=torch.bfloat16, device=obs.device) for _ in range(2)] + \
                            [torch.zeros((n, h + 2, w + 2, self.head_pad), dtype=torch.bfloat16, device=obs.device)]
            a, b, head = self.bufs
            self.first(x, a, n, h, w, c, stream)
            for conv in self.body[:-1]:
                conv(a, b, n, h, w, stream)
                a, b = b, a
            self.body[-1](a, head, n, h, w, stream)                  # c8: n_tools channels (padded to a multiple of 8)
            tools = self.n_tools
            logits = head[:, 1:-1, 1:-1, :tools].float().reshape(n, h * w * tools)
            # the value branch: VALID convolutions over the interior of the zero-bordered c8 map (im2col with pad = -1)
            v, vh, vw, vc = head, h + 2, w + 2, self.head_pad
            for conv in self.value_convs:
                v, vh, vw, vc = conv(v, True, n, vh, vw, vc, stream)
            value = self.vf(v, n, stream)                           # v viewed as [n, vh * vw * 64]
            return logits, value[:, 0]

    __call__ = forward

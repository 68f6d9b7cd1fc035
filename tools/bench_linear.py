"""pcgrl_linear_bf16 (tcgen05 / TMEM / TMA, csrc/pcgrl_linear.cu) vs torch (cuBLAS bf16 GEMM + bias + ReLU): CUDA-event
timing after warm-up, L2 flushed between timed launches.   python tools/bench_linear.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from gym_pcgrl_b200 import _native


def timeit(fn, flush, iters=20):
    for _ in range(5):
        fn()
    ms = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    ms.sort()
    return ms[len(ms) // 2]


def main():
    peak = 1662.5
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["bf16_tflops"])
    except Exception:
        pass
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for (m, n, k) in [(4096, 512, 1024), (32768, 512, 1024), (262144, 512, 1024), (8192, 4096, 4096)]:
        x = torch.randn((m, k), device="cuda").bfloat16()
        w = (torch.randn((n, k), device="cuda") / k ** 0.5).bfloat16()
        b = torch.randn(n, device="cuda")
        y = torch.empty((m, n), dtype=torch.float32, device="cuda")
        lib, sp = _native.lib(), _native.stream_ptr(x.device)
        ours = lambda: lib.pcgrl_linear_bf16(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), m, n, k, 1, sp)
        bb = b.bfloat16()
        ref = lambda: torch.relu(torch.addmm(bb, x, w.t()))
        t1, t2 = timeit(ours, flush), timeit(ref, flush)
        fl = 2.0 * m * n * k
        print("M=%6d N=%4d K=%4d  tcgen05 kernel %.3f ms = %.0f TFLOP/s (%.2f of the measured bf16 peak %.0f)   torch/cuBLAS bf16 %.3f ms = %.0f TFLOP/s"
              % (m, n, k, t1, fl / t1 / 1e9, fl / t1 / 1e9 / peak, peak, t2, fl / t2 / 1e9))

    # the implicit-GEMM 3 x 3 convolution (64 -> 64 channels, SAME) against cuDNN through torch (channels_last bf16)
    import torch.nn.functional as F
    for (n, h, w) in [(4096, 14, 14), (4096, 5, 5), (16384, 14, 14)]:
        c = 64
        xp = torch.zeros(n, h + 2, w + 2, c, dtype=torch.bfloat16, device="cuda")
        xp[:, 1:-1, 1:-1] = torch.randn(n, h, w, c, device="cuda").bfloat16()
        yp = torch.zeros_like(xp)
        wt = (0.05 * torch.randn(c, c, 3, 3, device="cuda")).bfloat16()
        w2 = wt.permute(0, 2, 3, 1).reshape(c, 9 * c).contiguous()
        b = torch.randn(c, device="cuda")
        lib, sp = _native.lib(), _native.stream_ptr(xp.device)
        ours = lambda: lib.pcgrl_conv3x3_bf16(xp.data_ptr(), w2.data_ptr(), b.data_ptr(), yp.data_ptr(), n, h, w, c, c, 1, sp)
        xt = xp[:, 1:-1, 1:-1].permute(0, 3, 1, 2).contiguous(memory_format=torch.channels_last)
        wc = wt.contiguous(memory_format=torch.channels_last)
        bb = b.bfloat16()
        ref = lambda: F.relu(F.conv2d(xt, wc, bb, padding=1))
        t1, t2 = timeit(ours, flush), timeit(ref, flush)
        useful = 2.0 * n * h * w * 9 * c * c
        issued = 2.0 * n * (h + 2) * (w + 2) * 9 * c * c
        hbm = 2.0 * n * (h + 2) * (w + 2) * c * 2      # one read + one write of the padded activations (bf16)
        print("conv3x3 %5d x %2dx%2d x 64->64  implicit GEMM %.3f ms = %.0f useful TFLOP/s (%.0f issued incl. the border, %.2f of the measured bf16 peak; "
              "%.0f GB/s of activation traffic)   torch/cuDNN bf16 channels_last %.3f ms = %.0f TFLOP/s"
              % (n, h, w, t1, useful / t1 / 1e9, issued / t1 / 1e9, issued / t1 / 1e9 / peak, hbm / t1 / 1e6, t2, useful / t2 / 1e9))


if __name__ == "__main__":
    main()

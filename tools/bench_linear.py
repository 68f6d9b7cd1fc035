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


if __name__ == "__main__":
    main()

"""Policy forward latency at the training batch (4096 envs): NativePolicy (im2col + tcgen05 GEMM, this repo's kernels) vs the
torch modules (cuDNN / cuBLAS) in fp32 and under bf16 autocast.   python tools/bench_policy.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from gym_pcgrl_b200.models import ActorCritic
from gym_pcgrl_b200.policy_native import NativePolicy


def timeit(fn, iters=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    n = 4096
    for kind, shape, acts in [("CustomPolicyBigMap", (28, 28, 1), 3), ("CustomPolicyBigMap", (22, 22, 8), 12),
                              ("FullyConvPolicyBigMap", (14, 14, 1), 14 * 14 * 2), ("FullyConvPolicySmallMap", (5, 5, 5), 125)]:
        net = ActorCritic(kind, shape, acts).cuda()
        obs = torch.randint(0, 2, (n,) + shape, dtype=torch.uint8, device="cuda")
        nat = NativePolicy(net)
        with torch.no_grad():
            t_nat = timeit(lambda: nat(obs))
            t_f32 = timeit(lambda: net(obs))
            with torch.autocast("cuda", dtype=torch.bfloat16):
                t_bf = timeit(lambda: net(obs))
        print("%-24s obs %-12s %d envs: native %.3f ms   torch fp32 %.3f ms   torch bf16 autocast %.3f ms" % (kind, shape, n, t_nat, t_f32, t_bf))


if __name__ == "__main__":
    main()

"""smb micro-benchmarks on one GPU: the stand-alone get_stats operator (maps/s) and the environment (env-steps/s).
    python tools/bench_smb.py [--envs 1024] [--steps 32]"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from gym_pcgrl_b200 import BatchedPcgrlEnv, _native


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=32)
    a = ap.parse_args()
    rs = np.random.RandomState(11)
    for nmaps in (1184, 3000, 8192):
        maps = rs.choice(7, size=(nmaps, 14, 114), p=[0.75, 0.1, 0.01, 0.04, 0.01, 0.02, 0.07]).astype(np.uint8)
        dm = torch.from_numpy(maps).cuda()
        _native.smb_get_stats(dm, 10000)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            st = _native.smb_get_stats(dm, 10000)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 3
        print("[smb] get_stats %5d maps 114x14: %.2f ms  %.3e maps/s  (wins %.0f%%)" % (nmaps, dt * 1e3, nmaps / dt, 100 * float((st[:, 7] == 0).float().mean())))
    for rep in ("narrow", "wide", "turtle"):
        n, T = a.envs, a.steps
        env = BatchedPcgrlEnv("smb", rep, num_envs=n, device="cuda", seed=0)
        env.reset()
        hi = [int(v) for v in env.action_space.nvec] if hasattr(env.action_space, "nvec") else [int(env.action_space.n)]
        acts = torch.stack([torch.randint(0, h, (T, n), device="cuda", dtype=torch.int32) for h in hi], dim=-1)
        acts = acts.contiguous() if len(hi) > 1 else acts[..., 0].contiguous()
        env.rollout(acts)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rew, done = env.rollout(acts)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t1 = time.perf_counter()
        for t in range(T):
            env.step(acts[t])
        torch.cuda.synchronize()
        dt2 = time.perf_counter() - t1
        print("[smb] env %-6s %d envs x %d steps: rollout %.1f ms (%.3e env-steps/s), per-step API %.1f ms (%.3e), done rate %.2f" % (
            rep, n, T, dt * 1e3, n * T / dt, dt2 * 1e3, n * T / dt2, float(done.float().mean())))


if __name__ == "__main__":
    main()

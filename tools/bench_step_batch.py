"""Per-step throughput of the solver problems as a function of the batch size (one B200).

A synchronous batched step lasts as long as its slowest env, and with thousands of envs nearly every step holds one
env in an iteration-capped A* pass (milliseconds); the per-step rate therefore grows with the batch until the
arenas are saturated.  Prints one JSON line per (workload, envs): device_step (pcgrl_step, actions in HBM) and e2e
(pcgrl_step_host, pinned host buffers, delta transport), env-steps/s.

    python tools/bench_step_batch.py [--workloads a,b] [--envs 2048,8192,32768,131072]
"""
import argparse
import json
import os
import sys
import types

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workloads", default="sokoban-wide-5x5,sokoban-wide-5x5-sparse,ddave-narrow-default,mdungeon-narrow-default,mdungeon-wide-default")
    ap.add_argument("--envs", default="2048,8192,32768,131072")
    ap.add_argument("--steps", type=int, default=48)
    a = ap.parse_args()
    import torch
    args = types.SimpleNamespace(sweep_only="", chunk=128, steps=a.steps)
    B = bench.Bench(args, 0, 1, 0)
    for name in a.workloads.split(","):
        wl = bench.WORKLOADS[name]
        for n in [int(v) for v in a.envs.split(",")]:
            env = bench.make_env(n, B.dev, env_offset=0, workload=wl)
            env.reset()
            B.preroll(env, 256, 77)
            acts = B.device_actions(env, 8 + a.steps, 900)
            for t in range(8):
                env.step(acts[t])
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for t in range(8, 8 + a.steps):
                env.step(acts[t])
            e1.record()
            torch.cuda.synchronize()
            step_ms = e0.elapsed_time(e1)
            e2e_ms, io, _ = B.time_e2e(env, a.steps, 3, "delta", 99)
            env.check_status()
            print(json.dumps({"workload": name, "envs": n, "device_step": n * a.steps / (step_ms * 1e-3),
                              "e2e": n * a.steps / (float(np.median(e2e_ms)) * 1e-3), "ms_per_step": step_ms / a.steps,
                              "unit": "env-steps/s"}), flush=True)
            del env, io


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py -- env-steps/s of the batched PcgrlEnv.step() hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference --gpus N --steps K ...  # the CPU arm (oracle port, all host threads)
    torchrun --nproc-per-node N ... bench.py --gpus N ...    # one rank per GPU, weak scaling

Workload (config 2 of BASELINE.json): binary-narrow 16x16, 4096 envs per GPU, uniform random actions on
Discrete(3), auto-reset on, env i seeded from its GLOBAL index (shard-invariant trajectories).

A "step" is one batched PcgrlEnv.step: every env of every rank advances by one action.
  value   env-steps/s with actions resident in HBM: the K steps run through pcgrl_rollout in chunks of
          --chunk steps per launch (default 128, the PPO2 rollout fragment length of the reference's train.py) (the fused step kernel keeps the bitboards in registers between steps);
          CUDA events around every chunk on the launch stream, L2 flushed between chunks, max over ranks.
  e2e     same metric through the reference-facing per-step C-ABI call with HOST buffers (pcgrl_step_host):
          every step copies that step's actions H2D from pinned memory and map + heatmap + pos + reward + done
          D2H, and synchronises -- the call a gym/VecEnv binding makes.
  roofline      k_rollout<binary>: algorithmic bytes (4*H*W + 64 per env-step) / event time,
                against the measured HBM copy bandwidth in MEASURED_PEAKS.json.
  cpu_baseline  the CPU oracle (oracle/pcgrl_oracle.c, a port of the reference's algorithms) on all host cores,
                bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

ZELDA_SPARSE = {"empty": 0.93, "solid": 0.02, "player": 0.006, "key": 0.006, "door": 0.006, "bat": 0.01,
                "scorpion": 0.01, "spider": 0.012}
SOKOBAN_SPARSE = {"empty": 0.7, "solid": 0.1, "player": 0.07, "crate": 0.065, "target": 0.065}
# BASELINE.json configs.  The default / headline workload is config 2.
WORKLOADS = {
    "binary-narrow-16x16": dict(prob="binary", rep="narrow", envs_per_gpu=4096,
                                kwargs=dict(width=16, height=16, change_percentage=0.2)),
    "zelda-turtle-11x16": dict(prob="zelda", rep="turtle", envs_per_gpu=4096,
                               kwargs=dict(width=11, height=16, change_percentage=0.2)),
    "zelda-turtle-11x16-sparse": dict(prob="zelda", rep="turtle", envs_per_gpu=4096,
                                      kwargs=dict(width=11, height=16, change_percentage=0.2, probs=ZELDA_SPARSE)),
    "sokoban-wide-5x5": dict(prob="sokoban", rep="wide", envs_per_gpu=2048, kwargs={}),
    "sokoban-wide-5x5-sparse": dict(prob="sokoban", rep="wide", envs_per_gpu=2048, kwargs=dict(probs=SOKOBAN_SPARSE)),
}
for _p in ("binary", "ddave", "mdungeon", "zelda"):          # config 5: default-size sweep, 8192 envs/GPU
    for _r in ("narrow", "turtle", "wide"):
        WORKLOADS["%s-%s-default" % (_p, _r)] = dict(prob=_p, rep=_r, envs_per_gpu=8192, kwargs={})
WORKLOAD = dict(WORKLOADS["binary-narrow-16x16"])
WORKLOAD_NAME = "binary-narrow 16x16, 4096 envs/GPU, random-action rollout, auto-reset"


def select_workload(name):
    global WORKLOAD, WORKLOAD_NAME
    WORKLOAD = dict(WORKLOADS[name])
    WORKLOAD_NAME = "%s, %d envs/GPU, random-action rollout, auto-reset" % (name, WORKLOAD["envs_per_gpu"])
HBM_FALLBACK_GBS = 6650.0


KERNEL_NAMES = {"binary": "k_rollout<binary>", "zelda": "k_rollout<zelda>"}


def algorithmic_bytes_per_env_step(w, h):
    return 4 * w * h + 64


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return HBM_FALLBACK_GBS, "fallback"


def make_env(num_envs, device, env_offset, auto_reset=True):
    from gym_pcgrl_b200 import BatchedPcgrlEnv
    env = BatchedPcgrlEnv(WORKLOAD["prob"], WORKLOAD["rep"], num_envs=num_envs, device=device, seed=0,
                          auto_reset=auto_reset, env_offset=env_offset)
    kw = WORKLOAD["kwargs"]
    if kw:
        env.adjust_param(**kw)
        env.adjust_param(**kw)  # quirk Q3: limits follow the size only on the second call
    return env


def action_high(env):
    sp = env.action_space
    return [int(v) for v in sp.nvec] if hasattr(sp, "nvec") else [int(sp.n)]


def host_actions(env, steps, n, seed):
    rng = np.random.RandomState(seed)
    hi = action_high(env)
    a = np.stack([rng.randint(h, size=(steps, n)) for h in hi], axis=-1).astype(np.int32)
    return a if len(hi) > 1 else a[..., 0]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower() == "active":
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def cpu_oracle_run(num_envs, seconds, threads, steps=None):
    """Time the CPU oracle on the workload: returns (env-steps/s, steps done, seconds)."""
    import oracle
    from gym_pcgrl_b200.seeding import mt_state_words
    env = make_env(num_envs, "cpu", 0)
    ref = oracle.OracleEnv(env.native_config, num_envs, threads=threads)
    rs = np.random.RandomState()
    states = np.empty((num_envs, 625), np.uint32)
    for i in range(num_envs):
        rs.seed(i)
        states[i] = mt_state_words(rs)
    ref.set_rng_states(states)
    ref.reset()
    acts = host_actions(env, 64, num_envs, 1)
    for k in range(2):
        ref.step(acts[k])
    done, t0 = 0, time.perf_counter()
    while True:
        ref.step(acts[done % 64])
        done += 1
        el = time.perf_counter() - t0
        if (steps is not None and done >= steps) or (steps is None and el >= seconds):
            break
    return num_envs * done / el, done, el


def run_reference(args, rank, world):
    """--impl reference: the CPU arm.  The reference is pure Python and cannot travel to the GPU box, so the arm
    is the oracle port (C, OpenMP over all host cores); rank 0 alone runs it.  Each step advances a bounded
    sample of the workload's envs, sized so that the whole run stays within ~2 minutes."""
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0))
    t0 = time.perf_counter()
    probe, _, _ = cpu_oracle_run(512, 1.5, cores)                      # env-steps/s estimate
    budget_s = 100.0
    n = int(min(WORKLOAD["envs_per_gpu"], max(64, probe * budget_s / max(1, args.steps + args.warmup))))
    import oracle
    from gym_pcgrl_b200.seeding import mt_state_words
    env = make_env(n, "cpu", 0)
    ref = oracle.OracleEnv(env.native_config, n, threads=cores)
    rs = np.random.RandomState()
    states = np.empty((n, 625), np.uint32)
    for i in range(n):
        rs.seed(i)
        states[i] = mt_state_words(rs)
    ref.set_rng_states(states)
    ref.reset()
    acts = host_actions(env, 64, n, 1)
    for k in range(args.warmup):
        ref.step(acts[k % 64])
    t1 = time.perf_counter()
    for k in range(args.steps):
        ref.step(acts[k % 64])
    el = time.perf_counter() - t1
    value = n * args.steps / el
    line = {
        "impl": "reference", "metric": "env steps/sec (batched)", "value": value, "unit": "env-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD_NAME, "envs_per_step_sample": n},
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": "port",
                         "sample": "%d batched steps x %d envs (bounded sample of the 4096-env batch), oracle/pcgrl_oracle.c "
                                   "with %d OpenMP threads; the Python reference itself cannot travel to the GPU box "
                                   "(BASELINE.md: ~1e3 steps/s per core)" % (args.steps, n, cores)},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "setup_s": t1 - t0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--chunk", type=int, default=128, help="env steps fused per pcgrl_rollout launch")
    ap.add_argument("--no-flush-l2", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--gather", action="store_true", help="all-gather reward/done across ranks after every chunk")
    ap.add_argument("--workload", default="binary-narrow-16x16", choices=sorted(WORKLOADS))
    args = ap.parse_args()
    select_workload(args.workload)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the step path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on stdout when the communicator comes up: keep stdout = the one JSON line
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize(dev)
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    from gym_pcgrl_b200 import HostRolloutIO, HostStepIO
    n = WORKLOAD["envs_per_gpu"]
    K, Wm, chunk = args.steps, max(args.warmup, 3), max(1, min(args.chunk, args.steps))
    env = make_env(n, dev, env_offset=rank * n)
    W, H = env._prob._width, env._prob._height
    env.reset()
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    hi = action_high(env)
    acts = torch.stack([torch.randint(0, h, (K + Wm, n), generator=gen, device=dev, dtype=torch.int32) for h in hi], dim=-1)
    acts = acts.contiguous() if len(hi) > 1 else acts[..., 0].contiguous()
    reward_buf = torch.empty((chunk, n), dtype=torch.float64, device=dev)
    done_buf = torch.empty((chunk, n), dtype=torch.uint8, device=dev)
    flush = None if args.no_flush_l2 else torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    gathered_r = torch.empty((world, chunk, n), dtype=torch.float64, device=dev) if (args.gather and world > 1) else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- warm-up (also moves the envs into steady state: mixed episode phases)
    for s in range(0, Wm, chunk):
        env.rollout(acts[s:min(s + chunk, Wm)], reward_buf, done_buf)
    barrier()

    # ---- timed region: exactly K steps, device-resident actions
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    events, launches, total_done = [], 0, 0
    wall0 = time.perf_counter()
    s = Wm
    while s < Wm + K:
        m = min(chunk, Wm + K - s)
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        env.rollout(acts[s:s + m], reward_buf, done_buf)
        e1.record()
        if gathered_r is not None:
            dist.all_gather_into_tensor(gathered_r, reward_buf)
        events.append((e0, e1, m))
        launches += 1
        s += m
    barrier()
    wall = time.perf_counter() - wall0
    dev_ms = sum(a.elapsed_time(b) for a, b, _ in events)
    env.check_status()

    # ---- e2e: per-step C-ABI call with host buffers: actions H2D from pinned memory, observation + reward + done
    # back on the host and the stream synchronised EVERY step (the call a gym / VecEnv binding makes)
    host_acts = torch.from_numpy(host_actions(env, K + 8, n, 99 + rank)).pin_memory()

    def run_e2e(mode):
        io = HostStepIO(env, with_obs=True, with_info=False, mode=mode)
        for t in range(8):
            io.struct.actions = host_acts[t].data_ptr()
            env.step_host(io)
        barrier()
        base, stride = host_acts.data_ptr(), host_acts.stride(0) * 4
        t0 = time.perf_counter()
        for t in range(K):
            io.struct.actions = base + (8 + t) * stride
            env.step_host(io)
        barrier()
        return time.perf_counter() - t0, io, float(io.reward.sum())

    e2e_s, io, rsum = run_e2e("delta")
    e2e_full_s, io_full, _ = run_e2e("full")

    # ---- e2e_rollout: the open-loop host call (pcgrl_rollout_host): `chunk` steps per call, pinned host actions in,
    # every step's reward / done plus the final observation back on the host, stream synchronised per call
    def run_e2e_rollout():
        rio = HostRolloutIO(env, chunk, with_obs=True, with_info=False)
        ncalls = max(1, K // chunk)
        acts_h = torch.from_numpy(host_actions(env, (ncalls + 1) * chunk, n, 199 + rank)).pin_memory()
        base, stride = acts_h.data_ptr(), acts_h.stride(0) * 4 * chunk
        rio.struct.actions = base
        env.rollout_host(rio)
        barrier()
        t0 = time.perf_counter()
        for c in range(ncalls):
            rio.struct.actions = base + (1 + c) * stride
            env.rollout_host(rio)
        barrier()
        return time.perf_counter() - t0, rio, ncalls * chunk

    e2e_roll_s, rio, roll_steps = run_e2e_rollout()
    clocks = sampler.stop() if rank == 0 else None

    # ---- max over ranks
    t_dev = torch.tensor([dev_ms, e2e_s * 1e3, e2e_full_s * 1e3, e2e_roll_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    dev_ms_max, e2e_ms_max, e2e_full_ms_max, e2e_roll_ms_max = float(t_dev[0]), float(t_dev[1]), float(t_dev[2]), float(t_dev[3])

    if rank == 0:
        total_envs = n * world
        value = total_envs * K / (dev_ms_max * 1e-3)
        e2e_value = total_envs * K / (e2e_ms_max * 1e-3)
        peak, peak_src = measured_peak()
        bytes_per_launch = algorithmic_bytes_per_env_step(W, H) * n * chunk
        avg_launch_s = (dev_ms * 1e-3) / launches
        achieved = bytes_per_launch / avg_launch_s / 1e9
        cores = len(os.sched_getaffinity(0))
        if world == 1:   # the CPU baseline is timed at N=1 only (other ranks would be spinning on the same cores)
            cpu_value, cpu_steps, cpu_s = cpu_oracle_run(n, args.cpu_seconds, cores)
            cpu_baseline = {"value": cpu_value, "unit": "env-steps/s", "cores": cores, "kind": "port",
                            "sample": "%d batched steps x %d envs in %.1f s, oracle/pcgrl_oracle.c, %d OpenMP threads" % (cpu_steps, n, cpu_s, cores)}
        else:
            cpu_baseline = None
        line = {
            "metric": "env steps/sec (batched)", "value": value, "unit": "env-steps/s", "n_gpus": world,
            "steps": K, "warmup": Wm, "ms_per_step": dev_ms_max / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD_NAME, "envs_per_gpu": n, "global_envs": total_envs,
                       "steps_per_launch": chunk, "l2": "flushed between launches (256 MiB write)" if flush is not None else "not flushed",
                       "timing": "CUDA events per launch on the launch stream, summed, max over ranks",
                       "parallelism": "env-index sharding, no data-path collective" + (" + all_gather(reward)" if gathered_r is not None else "")},
            "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": io.h2d_bytes * world,
                    "d2h_bytes_per_step": io.d2h_bytes * world, "ms_per_step": e2e_ms_max / K,
                    "api": "pcgrl_step_host mode 1 (pinned host buffers; per-env delta records + fresh maps of reset envs copied "
                           "back and applied, so the host arrays hold the complete map+heatmap+pos+reward+done after every step)"},
            "e2e_full_copy": {"value": total_envs * K / (e2e_full_ms_max * 1e-3), "unit": "env-steps/s",
                              "h2d_bytes_per_step": io_full.h2d_bytes * world, "d2h_bytes_per_step": io_full.d2h_bytes * world,
                              "ms_per_step": e2e_full_ms_max / K, "api": "pcgrl_step_host mode 0 (every array copied back in full)"},
            "e2e_rollout": {"value": total_envs * roll_steps / (e2e_roll_ms_max * 1e-3), "unit": "env-steps/s",
                            "h2d_bytes_per_step": rio.h2d_bytes * world // chunk, "d2h_bytes_per_step": rio.d2h_bytes * world // chunk,
                            "ms_per_step": e2e_roll_ms_max / roll_steps, "steps_per_call": chunk,
                            "api": "pcgrl_rollout_host (open loop: %d steps per call, pinned host actions in; every step's reward + done "
                                   "and the final map+heatmap+pos back on the host)" % chunk},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "peak_source": peak_src, "kernel": KERNEL_NAMES.get(WORKLOAD["prob"], "k_rollout_async<%s>" % WORKLOAD["prob"]),
                         "algorithmic_bytes_per_env_step": algorithmic_bytes_per_env_step(W, H),
                         "units_per_launch": n * chunk, "avg_launch_ms": avg_launch_s * 1e3},
            "cpu_baseline": cpu_baseline,
            "clocks": clocks, "wall_s_timed_region": wall, "check_reward_sum": rsum,
        }
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                if args.workload == "binary-narrow-16x16":
                    line["roofline"]["traffic"] = json.load(f).get("k_rollout_binary_bytes_per_launch")
        except Exception:
            pass
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

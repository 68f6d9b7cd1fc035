#!/usr/bin/env python
"""bench.py -- env-steps/s of the batched PcgrlEnv.step() hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference --gpus N --steps K ...  # the CPU arm: the unmodified Python reference (oracle/_ref)
    torchrun --nproc-per-node N ... bench.py --gpus N ...    # one rank per GPU, weak scaling

Workload (config 2 of BASELINE.json): binary-narrow 16x16, 4096 envs per GPU, uniform random actions on
Discrete(3), auto-reset on, env i seeded from its GLOBAL index (shard-invariant trajectories).

A "step" is one batched PcgrlEnv.step: every env of every rank advances by one action.  Before anything is timed the
batch is rolled 512 steps forward (several episode lengths) so that episodes are in mixed phases and auto-resets are in
flight; then W warm-up steps; then the K-step timed region is run R times back to back (R grows as K shrinks so that
a 20-step request is not one 0.2 ms sample) and every number below is the MEDIAN over the R repeats of the
max-over-ranks time of one K-step region.

  value        open-loop fused rollout: the K steps run through pcgrl_rollout in launches of `steps_per_launch` steps
               with the actions already in HBM (the kernel keeps the bitboards in registers between steps); CUDA
               events around every launch on the launch stream, L2 flushed between launches.  This is the named
               workload ("random-action rollout"): the actions do not depend on the observations.
  closed_loop  the SURVEY 8d loop `actions -> step -> (obs, reward, done)` device-resident: every step draws its
               actions ON THE DEVICE (torch.randint) and then calls pcgrl_step once, so a policy could sit in between;
               `plain` = 2 launches per step issued from Python, `graph` = the same K-step sequence captured once in a
               CUDA graph and replayed.
  e2e          the per-step C-ABI call with HOST buffers (pcgrl_step_host): every step reads that step's actions from
               pinned host memory and brings map + heatmap + pos + reward + done back to the host, synchronised -- the
               call a gym / VecEnv binding makes.  Transport = mode 2 ("direct": the step kernel stores the results into
               the pinned host arrays while it runs); `e2e_delta` (mode 1: change records + one D2H copy, applied by the
               library) and `e2e_full_copy` (mode 0) are the other two transports of the same call.
  roofline     k_rollout<binary>: algorithmic bytes (4*H*W + 64 per env-step) / event time of the
               launches of the timed region, against the measured HBM copy bandwidth in MEASURED_PEAKS.json; `issue`
               is the warp-instruction issue-rate view of the same kernel (the limiter that actually binds).
  sweep        BASELINE configs 3-5 (zelda-turtle 11x16, sokoban-wide 5x5, the 12 default-size combos), short runs.
  cpu_baseline the unmodified Python reference (oracle/_ref, one process per core) on a bounded sample of the same
               workload; cpu_baseline_port = the C/OpenMP oracle port, a much stronger baseline than the reference.
"""
import argparse
import hashlib
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
    "smb-narrow-114x14": dict(prob="smb", rep="narrow", envs_per_gpu=1024, kwargs={}),
}
for _p in ("binary", "ddave", "mdungeon", "zelda"):          # config 5: default-size sweep, 8192 envs/GPU
    for _r in ("narrow", "turtle", "wide"):
        WORKLOADS["%s-%s-default" % (_p, _r)] = dict(prob=_p, rep=_r, envs_per_gpu=8192, kwargs={})
SWEEP = ["zelda-turtle-11x16", "zelda-turtle-11x16-sparse", "sokoban-wide-5x5", "sokoban-wide-5x5-sparse"] + \
        ["%s-%s-default" % (p, r) for p in ("binary", "ddave", "mdungeon", "zelda") for r in ("narrow", "turtle", "wide")] + \
        ["smb-narrow-114x14"]
WORKLOAD = dict(WORKLOADS["binary-narrow-16x16"])
WORKLOAD_NAME = "binary-narrow 16x16, 4096 envs/GPU, random-action rollout, auto-reset"
STEADY_STATE_STEPS = 512     # untimed pre-roll: 16x16 episodes last ~150 steps, so resets are in flight afterwards
HBM_FALLBACK_GBS = 6650.0
KERNEL_NAMES = {"binary": "k_rollout<binary>", "zelda": "k_rollout<zelda>"}


def select_workload(name):
    global WORKLOAD, WORKLOAD_NAME
    WORKLOAD = dict(WORKLOADS[name])
    if name == "binary-narrow-16x16":
        WORKLOAD_NAME = "binary-narrow 16x16, 4096 envs/GPU, random-action rollout, auto-reset"
    else:
        WORKLOAD_NAME = "%s, %d envs/GPU, random-action rollout, auto-reset" % (name, WORKLOAD["envs_per_gpu"])


def algorithmic_bytes_per_env_step(w, h):
    return 4 * w * h + 64


def repeats_for(steps):
    """How often the K-step timed region is repeated: ~2000 timed steps in total, at least 5 and at most 100 regions."""
    return int(min(100, max(5, -(-2000 // max(1, steps)))))


def config_dict(args, world):
    """The workload description: a function of the command line only, so both arms print the SAME dict."""
    n = WORKLOAD["envs_per_gpu"]
    chunk = max(1, min(args.chunk, args.steps))
    return {"workload": WORKLOAD_NAME, "envs_per_gpu": n, "global_envs": n * world,
            "value_is": "open_loop_rollout (pcgrl_rollout, actions resident in HBM); closed_loop and e2e are separate keys",
            "steps_per_launch": chunk, "repeats_of_timed_region": repeats_for(args.steps),
            "steady_state_preroll_steps": STEADY_STATE_STEPS,
            "l2": "not flushed" if args.no_flush_l2 else "flushed between launches (256 MiB write)",
            "timing": "CUDA events per launch on the launch stream, summed per K-step region, max over ranks, median over regions",
            "parallelism": "env-index sharding, no data-path collective"}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return HBM_FALLBACK_GBS, "fallback"


def make_env(num_envs, device, env_offset, auto_reset=True, workload=None):
    from gym_pcgrl_b200 import BatchedPcgrlEnv
    wl = workload or WORKLOAD
    env = BatchedPcgrlEnv(wl["prob"], wl["rep"], num_envs=num_envs, device=device, seed=0,
                          auto_reset=auto_reset, env_offset=env_offset)
    kw = wl["kwargs"]
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


# ------------------------------------------------------------------------------------------------ CPU arms
def cpu_port_run(num_envs, seconds, threads, steps=None, warmup=2):
    """Time the C oracle port (oracle/pcgrl_oracle.c, OpenMP) on the workload: (env-steps/s, steps done, seconds)."""
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
    for k in range(warmup):
        ref.step(acts[k % 64])
    done, t0 = 0, time.perf_counter()
    while True:
        ref.step(acts[done % 64])
        done += 1
        el = time.perf_counter() - t0
        if (steps is not None and done >= steps) or (steps is None and el >= seconds):
            break
    return num_envs * done / el, done, el


def reference_available():
    from oracle import make_ref
    return make_ref.available()


def cpu_reference_run(envs, warmup, steps, procs, max_seconds):
    """The UNMODIFIED Python reference from oracle/_ref, one process per core, in a child process (it forks workers)."""
    spec = {"prob": WORKLOAD["prob"], "rep": WORKLOAD["rep"], "kwargs": WORKLOAD["kwargs"], "envs": int(envs),
            "warmup": int(warmup), "steps": int(steps), "procs": int(procs), "max_seconds": float(max_seconds)}
    out = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "ref_runner.py"), json.dumps(spec)],
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=max_seconds * 3 + 300)
    if out.returncode != 0:
        raise RuntimeError("ref_runner failed: %s" % out.stderr[-400:])
    return json.loads(out.stdout.strip().splitlines()[-1])


def run_reference(args, rank, world):
    """--impl reference: the CPU arm, rank 0 only.  The reference's own Python PcgrlEnv.step (oracle/_ref, unmodified)
    with one process per host core when the copy is present, else the C port of its algorithms (OpenMP, all cores).
    Each step advances a bounded sample of the workload's env batch, sized so the whole run stays within ~2 minutes."""
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0))
    n_full = WORKLOAD["envs_per_gpu"]
    t0 = time.perf_counter()
    total_steps = max(1, args.steps + args.warmup)
    port_value, port_steps, port_s = cpu_port_run(min(n_full, 1024), 3.0, cores)
    port = {"value": port_value, "unit": "env-steps/s", "cores": cores, "kind": "port",
            "sample": "%d batched steps x %d envs in %.1f s, oracle/pcgrl_oracle.c, %d OpenMP threads" % (port_steps, min(n_full, 1024), port_s, cores)}
    if reference_available():
        probe = cpu_reference_run(min(64, n_full), 1, 6, cores, 30.0)["env_steps_per_s"]
        n = int(min(n_full, max(cores, probe * 100.0 / total_steps)))
        r = cpu_reference_run(n, args.warmup, args.steps, cores, 150.0)
        value, kind = r["env_steps_per_s"], "reference"
        ms_per_step = 1e3 * r["seconds"] / max(1, r["steps"])
        sample = ("%d batched steps x %d envs (bounded sample of the %d-env batch) in %.1f s: unmodified Python reference "
                  "(oracle/_ref/gym_pcgrl, PcgrlEnv.step) under the gym stand-in, %d worker processes, reset on done"
                  % (r["steps"], n, n_full, r["seconds"], r["procs"]))
    else:
        n = int(min(n_full, max(64, port_value * 100.0 / total_steps)))
        value, steps_done, el = cpu_port_run(n, None, cores, steps=args.steps, warmup=args.warmup)
        kind, ms_per_step = "port", 1e3 * el / steps_done
        sample = "%d batched steps x %d envs, oracle/pcgrl_oracle.c with %d OpenMP threads (oracle/_ref is absent)" % (steps_done, n, cores)
    line = {
        "impl": "reference", "metric": "env steps/sec (batched)", "value": value, "unit": "env-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": config_dict(args, world),
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": kind, "sample": sample},
        "cpu_baseline_port": port,
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "setup_s": time.perf_counter() - t0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
class Bench:
    def __init__(self, args, rank, world, local_rank):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.args, self.rank, self.world = args, rank, world
        self.dev = torch.device("cuda", local_rank)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, values):
        t = self.torch.tensor(list(values), dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.cpu().numpy()

    def device_actions(self, env, steps, seed):
        torch = self.torch
        gen = torch.Generator(device=self.dev)
        gen.manual_seed(seed)
        hi = action_high(env)
        n = env.num_envs
        acts = torch.stack([torch.randint(0, h, (steps, n), generator=gen, device=self.dev, dtype=torch.int32) for h in hi], dim=-1)
        return acts.contiguous() if len(hi) > 1 else acts[..., 0].contiguous()

    def preroll(self, env, steps, seed):
        acts = self.device_actions(env, 128, seed)
        rb = self.torch.empty((128, env.num_envs), dtype=self.torch.float64, device=self.dev)
        db = self.torch.empty((128, env.num_envs), dtype=self.torch.uint8, device=self.dev)
        for s in range(0, steps, 128):
            env.rollout(acts[:min(128, steps - s)], rb, db)
        self.torch.cuda.synchronize(self.dev)

    # -- open loop: K steps through pcgrl_rollout, `chunk` steps per launch; returns per-region ms (this rank), launches
    def time_rollout(self, env, K, Wm, chunk, R, flush, seed, gathered=None):
        torch = self.torch
        n = env.num_envs
        acts = self.device_actions(env, K + Wm, seed)
        rb = torch.empty((chunk, n), dtype=torch.float64, device=self.dev)
        db = torch.empty((chunk, n), dtype=torch.uint8, device=self.dev)
        for s in range(0, Wm, chunk):
            env.rollout(acts[s:min(s + chunk, Wm)], rb, db)
        self.barrier()
        per_region, launch_ms, launches = [], [], 0
        for r in range(R):
            events = []
            s = Wm
            while s < Wm + K:
                m = min(chunk, Wm + K - s)
                if flush is not None:
                    flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                env.rollout(acts[s:s + m], rb, db)
                if gathered is not None:
                    self.dist.all_gather_into_tensor(gathered[0], rb)
                    self.dist.all_gather_into_tensor(gathered[1], db)
                e1.record()
                events.append((e0, e1, m))
                s += m
            self.barrier()
            ms = [a.elapsed_time(b) for a, b, _ in events]
            per_region.append(sum(ms))
            launch_ms += [t for t, (_, _, m) in zip(ms, events) if m == chunk]
            launches += len(events)
        return per_region, launch_ms, launches

    # -- closed loop: per step, actions drawn on the device, then ONE pcgrl_step
    def time_closed_loop(self, env, K, R):
        torch = self.torch
        n = env.num_envs
        hi = action_high(env)
        a = torch.zeros((n, len(hi)) if len(hi) > 1 else (n,), dtype=torch.int32, device=self.dev)
        torch.cuda.manual_seed(4321 + self.rank)   # default CUDA generator: graph-safe (philox offsets are patched on replay)

        def draw():
            if len(hi) == 1:
                torch.randint(0, hi[0], (n,), out=a)
            else:
                for j, h in enumerate(hi):
                    a[:, j] = torch.randint(0, h, (n,), device=self.dev, dtype=torch.int32)

        def k_steps():
            for _ in range(K):
                draw()
                env.step(a)

        for _ in range(3):
            draw()
            env.step(a)
        self.barrier()
        plain, kernel_ms = [], []
        Rp = min(R, 20)
        for r in range(Rp):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            k_steps()
            e1.record()
            self.barrier()
            plain.append(e0.elapsed_time(e1))
        # the step kernel alone (events bracket only the pcgrl_step launch)
        evs = []
        for _ in range(min(64, max(K, 16))):
            draw()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            env.step(a)
            e1.record()
            evs.append((e0, e1))
        self.barrier()
        kernel_ms = [x.elapsed_time(y) for x, y in evs]
        graph_ms = None
        try:
            side = torch.cuda.Stream(device=self.dev)
            side.wait_stream(torch.cuda.current_stream(self.dev))
            with torch.cuda.stream(side):
                k_steps()      # warm the capture stream
            torch.cuda.current_stream(self.dev).wait_stream(side)
            self.torch.cuda.synchronize(self.dev)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                k_steps()
            g.replay()
            self.barrier()
            graph_ms = []
            for r in range(R):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                g.replay()
                e1.record()
                self.barrier()
                graph_ms.append(e0.elapsed_time(e1))
        except Exception as ex:   # capture is an optimisation of the launch path, not part of the metric
            self.graph_error = repr(ex)[:200]
            self.torch.cuda.synchronize(self.dev)
        return plain, graph_ms, kernel_ms

    # -- e2e: per-step C-ABI call with host buffers, wall clock (the call synchronises)
    def time_e2e(self, env, K, R, mode, seed):
        from gym_pcgrl_b200 import HostStepIO
        torch = self.torch
        n = env.num_envs
        host_acts = torch.from_numpy(host_actions(env, K + 8, n, seed)).pin_memory()
        io = HostStepIO(env, with_obs=True, with_info=False, mode=mode)
        base, stride = host_acts.data_ptr(), host_acts.stride(0) * 4
        for t in range(8):
            io.struct.actions = base + t * stride
            env.step_host(io)
        c0, r0, nsteps = int(io.struct.change_base), int(io.struct.reset_base), 0   # running counters of the delta transport
        # wall clock: the K-step region is run `per` times inside one barrier bracket (the NCCL barrier itself costs about
        # as much as one step, so a bracket around a single 20-step region would charge it to the steps); ms per REGION
        per = max(1, min(R, 400 // max(1, K)))
        out = []
        for r in range(max(3, R // per)):
            self.barrier()
            t0 = time.perf_counter()
            for rep in range(per):
                for t in range(K):
                    io.struct.actions = base + (8 + t) * stride
                    env.step_host(io)
            t1 = time.perf_counter()
            self.barrier()
            out.append((t1 - t0) * 1e3 / per)
            nsteps += per * K
        if mode == "delta":   # per-step averages over the timed steps: envs with a change record, whole-map updates (auto-resets)
            io.records_per_step = (int(io.struct.change_base) - c0) / max(1, nsteps)
            io.resets_per_step = (int(io.struct.reset_base) - r0) / max(1, nsteps)
        return out, io, float(io.reward.sum())

    def time_e2e_rollout(self, env, K, chunk, R, seed):
        from gym_pcgrl_b200 import HostRolloutIO
        torch = self.torch
        n = env.num_envs
        rio = HostRolloutIO(env, chunk, with_obs=True, with_info=False)
        ncalls = max(1, K // chunk)
        acts_h = torch.from_numpy(host_actions(env, (ncalls + 1) * chunk, n, seed)).pin_memory()
        base, stride = acts_h.data_ptr(), acts_h.stride(0) * 4 * chunk
        rio.struct.actions = base
        env.rollout_host(rio)
        per = max(1, min(R, 400 // max(1, K)))
        out = []
        for r in range(max(3, R // per)):
            self.barrier()
            t0 = time.perf_counter()
            for rep in range(per):
                for c in range(ncalls):
                    rio.struct.actions = base + (1 + c) * stride
                    env.rollout_host(rio)
            t1 = time.perf_counter()
            self.barrier()
            out.append((t1 - t0) * 1e3 / per)
        return out, rio, ncalls * chunk

    # -- asynchronous grouped stepping (opt-in API): every env group keeps one step in flight on its own stream
    def time_e2e_async(self, wl, n, K, groups, seed):
        from gym_pcgrl_b200 import AsyncGroupedEnv
        torch = self.torch
        env = AsyncGroupedEnv(wl["prob"], wl["rep"], num_envs=n, groups=groups, device=self.dev, seed=0, env_offset=self.rank * n)
        if wl["kwargs"]:
            env.adjust_param(**wl["kwargs"])
            env.adjust_param(**wl["kwargs"])
        env.reset()
        m = env.per_group
        acts = torch.from_numpy(host_actions(env.envs[0], (K + 4) * groups, m, seed)).pin_memory()   # [(K+4)*groups, m(, k)]
        base, stride = acts.data_ptr(), acts.stride(0) * 4

        def send(g, t):
            env.io[g].struct.actions = base + (t * groups + g) * stride
            env.send(g)

        for t in range(4):          # warm-up, synchronous
            for g in range(groups):
                send(g, t)
            while any(env.in_flight):
                env.recv(wait=True)
        self.barrier()
        step = [4] * groups
        t0 = time.perf_counter()
        for g in range(groups):
            send(g, step[g])
        remaining = groups * K
        while remaining:
            for g in env.recv(wait=True):
                remaining -= 1
                step[g] += 1
                if step[g] < K + 4:
                    send(g, step[g])
        self.barrier()
        dt = (time.perf_counter() - t0) * 1e3
        env.check_status()
        return dt

    # -- shard invariance on hardware: 1024 GLOBAL envs split over the ranks must give the same per-env results at any N
    def shard_check(self):
        torch = self.torch
        G, T = 1024, 96
        m = G // self.world
        lo = self.rank * m
        wl = WORKLOADS["binary-narrow-16x16"]
        env = make_env(m, self.dev, env_offset=lo, workload=wl)
        env.reset()
        acts_all = np.random.RandomState(2024).randint(3, size=(T, G)).astype(np.int32)   # action of GLOBAL env i at step t
        acts = torch.from_numpy(np.ascontiguousarray(acts_all[:, lo:lo + m])).to(self.dev)
        rew, done = env.rollout(acts)
        torch.cuda.synchronize(self.dev)
        # one 64-bit digest per env over everything a caller can observe
        per_env = []
        maps, heat, pos = env._tens["map"].cpu().numpy(), env._tens["heatmap"].cpu().numpy(), env._tens["pos"].cpu().numpy()
        rew_h, done_h, stats = rew.cpu().numpy(), done.cpu().numpy(), env._tens["stats"].cpu().numpy()
        for i in range(m):
            h = hashlib.sha256()
            for arr in (maps[i], heat[i], pos[i], rew_h[:, i], done_h[:, i], stats[i]):
                h.update(np.ascontiguousarray(arr).tobytes())
            per_env.append(int.from_bytes(h.digest()[:8], "little", signed=True))
        mine = torch.tensor(per_env, dtype=torch.int64, device=self.dev)
        if self.world > 1:
            allv = torch.empty(G, dtype=torch.int64, device=self.dev)
            self.dist.all_gather_into_tensor(allv, mine)
        else:
            allv = mine
        digest = hashlib.sha256(allv.cpu().numpy().tobytes()).hexdigest()[:16]
        return {"global_envs": G, "steps": T, "envs_per_rank": m, "sha256_16": digest,
                "note": "same value at every N <=> env i follows the same trajectory on 1 or N GPUs"}

    # -- BASELINE configs 3-5, short runs
    def sweep(self):
        torch = self.torch
        out = {}
        names = SWEEP if not self.args.sweep_only else [w for w in SWEEP if w in self.args.sweep_only.split(",")]
        for name in names:
            wl = WORKLOADS[name]
            n = wl["envs_per_gpu"]
            t_wall = time.perf_counter()
            try:
                env = make_env(n, self.dev, env_offset=self.rank * n, workload=wl)
                env.reset()
                # smb runs an A* play-through of up to 2 x 10000 iterations per edited step: shorter runs
                Ks, Cs, Rs = (64, 32, 2) if wl["prob"] == "smb" else (256, 128, 3)
                self.preroll(env, Ks, 77 + self.rank)
                per_region, launch_ms, _ = self.time_rollout(env, Ks, Cs, Cs, Rs, None, 500 + self.rank)
                # closed loop on the device: one pcgrl_step per step, actions resident in HBM
                acts = self.device_actions(env, 64, 900 + self.rank)
                for t in range(4):
                    env.step(acts[t])
                self.barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for t in range(4, 64):
                    env.step(acts[t])
                e1.record()
                self.barrier()
                step_ms = e0.elapsed_time(e1)
                # per-step host API: direct transport for the graph-only problems (the transport matters there), delta
                # records for the solver problems (their step lasts milliseconds either way)
                e2e_ms, io, _ = self.time_e2e(env, 48, 2, "direct" if wl["prob"] in ("binary", "zelda") else "delta", 99 + self.rank)
                env.check_status()
                async_ms = None
                if wl["prob"] in ("sokoban", "ddave", "mdungeon", "smb"):   # a batch step waits for its slowest search
                    del io
                    io = None
                    async_ms = self.time_e2e_async(wl, n, 48, 16, 1999 + self.rank)
                big_ms, big_n = None, 131072
                if wl["prob"] in ("sokoban", "ddave", "mdungeon"):
                    # the same per-step call on a batch large enough to amortise the longest search of the step
                    del env
                    env = make_env(big_n, self.dev, env_offset=self.rank * big_n, workload=wl)
                    env.reset()
                    self.preroll(env, 256, 78 + self.rank)
                    b_ms, io, _ = self.time_e2e(env, 24, 1, "delta", 98 + self.rank)
                    env.check_status()
                    big_ms = float(np.median(b_ms))
                red = self.max_over_ranks([float(np.median(per_region)), step_ms, float(np.median(e2e_ms)), async_ms or 0.0, big_ms or 0.0])
                tot = n * self.world
                W, H = env._prob._width, env._prob._height
                out[name] = {"envs_per_gpu": n, "map": "%dx%d" % (W, H),
                             "value": tot * Ks / (red[0] * 1e-3), "device_step": tot * 60 / (red[1] * 1e-3),
                             "e2e": tot * 48 / (red[2] * 1e-3), "unit": "env-steps/s",
                             "hbm_frac": tot / self.world * Ks * algorithmic_bytes_per_env_step(W, H) / (red[0] * 1e-3) / 1e9 / measured_peak()[0],
                             "wall_s": time.perf_counter() - t_wall}
                if async_ms:
                    out[name]["e2e_async_groups"] = tot * 48 / (red[3] * 1e-3)
                    out[name]["e2e_async_api"] = "AsyncGroupedEnv: 16 env groups, one pcgrl_step_host_begin/_end step in flight each"
                if big_ms:
                    out[name]["e2e_large_batch"] = {"envs_per_gpu": big_n, "value": big_n * self.world * 24 / (red[4] * 1e-3)}
                del env, io
            except Exception as ex:
                out[name] = {"error": repr(ex)[:300]}
        return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--chunk", type=int, default=128, help="env steps fused per pcgrl_rollout launch")
    ap.add_argument("--no-flush-l2", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=6.0)
    ap.add_argument("--no-sweep", action="store_true", help="skip the configs 3-5 sweep")
    ap.add_argument("--sweep-only", default="", help="comma-separated subset of sweep workloads")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baselines (profiling runs)")
    ap.add_argument("--only-rollout", action="store_true", help="time the open-loop rollout only (ncu runs)")
    ap.add_argument("--repeats", type=int, default=0, help="override the number of timed regions")
    ap.add_argument("--workload", default="binary-narrow-16x16", choices=sorted(WORKLOADS))
    ap.add_argument("--envs", type=int, default=0, help="override envs per GPU (experiments; the headline uses the workload's own)")
    args = ap.parse_args()
    select_workload(args.workload)
    if args.envs > 0:
        global WORKLOAD_NAME
        WORKLOAD["envs_per_gpu"] = args.envs
        WORKLOAD_NAME = "%s, %d envs/GPU (override), random-action rollout, auto-reset" % (args.workload, args.envs)

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

    B = Bench(args, rank, world, local_rank)
    n = WORKLOAD["envs_per_gpu"]
    K, Wm, chunk = args.steps, max(args.warmup, 3), max(1, min(args.chunk, args.steps))
    R = args.repeats if args.repeats > 0 else repeats_for(K)
    t_start = time.perf_counter()
    env = make_env(n, dev, env_offset=rank * n)
    W, H = env._prob._width, env._prob._height
    env.reset()
    flush = None if args.no_flush_l2 else torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    B.preroll(env, STEADY_STATE_STEPS, 11 + rank)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    wall0 = time.perf_counter()
    per_region, launch_ms, launches = B.time_rollout(env, K, Wm, chunk, R, flush, 1234 + rank)
    env.check_status()
    wall_rollout = time.perf_counter() - wall0
    region_ms = B.max_over_ranks(per_region)          # max over ranks, per region
    dev_ms = float(np.median(region_ms))

    res = {}
    if not args.only_rollout:
        plain, graph_ms, kernel_ms = B.time_closed_loop(env, K, R)
        res["plain"] = float(np.median(B.max_over_ranks(plain)))
        res["graph"] = float(np.median(B.max_over_ranks(graph_ms))) if graph_ms else None
        res["kernel"] = float(np.median(kernel_ms))
        Re = max(3, min(R, 25))
        e2e_ms, io, rsum = B.time_e2e(env, K, Re, "direct", 99 + rank)
        e2e_delta_ms, io_delta, _ = B.time_e2e(env, K, Re, "delta", 149 + rank)
        e2e_full_ms, io_full, _ = B.time_e2e(env, K, max(3, Re // 3), "full", 199 + rank)
        e2e_roll_ms, rio, roll_steps = B.time_e2e_rollout(env, K, chunk, max(3, Re // 3), 299 + rank)
        res["e2e"] = float(np.median(B.max_over_ranks(e2e_ms)))
        res["e2e_delta"] = float(np.median(B.max_over_ranks(e2e_delta_ms)))
        res["e2e_full"] = float(np.median(B.max_over_ranks(e2e_full_ms)))
        res["e2e_roll"] = float(np.median(B.max_over_ranks(e2e_roll_ms)))
        gather = None
        if world > 1:   # the one collective the design allows: all-gather of every launch's reward + done rows
            gr = torch.empty((world, chunk, n), dtype=torch.float64, device=dev)
            gd = torch.empty((world, chunk, n), dtype=torch.uint8, device=dev)
            g_region, _, _ = B.time_rollout(env, K, 3, chunk, max(3, R // 4), flush, 777 + rank, gathered=(gr, gd))
            gather = float(np.median(B.max_over_ranks(g_region)))
    clocks = sampler.stop() if rank == 0 else None
    shard = None if args.only_rollout else B.shard_check()
    sweep = None if (args.no_sweep or args.only_rollout) else B.sweep()

    if rank == 0:
        total_envs = n * world
        value = total_envs * K / (dev_ms * 1e-3)
        peak, peak_src = measured_peak()
        bpe = algorithmic_bytes_per_env_step(W, H)
        avg_launch_s = (float(np.mean(launch_ms)) if launch_ms else dev_ms * chunk / K) * 1e-3
        achieved = bpe * n * chunk / avg_launch_s / 1e9
        traffic, inst_per_step = None, None
        try:   # ncu-measured numbers are attached only when they were captured for THIS launch shape
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                tj = json.load(f).get(KERNEL_NAMES.get(WORKLOAD["prob"], ""), {}).get("%s:T%d_n%d" % (args.workload, chunk, n))
            if tj:
                traffic, inst_per_step = tj.get("dram_bytes_per_launch"), tj.get("warp_inst_per_env_step")
        except Exception:
            pass
        sm_hz = ((clocks or {}).get("sm_mhz") or 1965.0) * 1e6
        issue_peak = 148 * 4 * sm_hz
        issue = None
        if inst_per_step:
            per_gpu_steps_s = n * chunk / avg_launch_s
            issue = {"warp_inst_per_env_step": inst_per_step, "achieved_warp_inst_per_s": per_gpu_steps_s * inst_per_step,
                     "peak_warp_inst_per_s": issue_peak, "frac": per_gpu_steps_s * inst_per_step / issue_peak,
                     "peak_is": "148 SMs x 4 schedulers x sampled SM clock", "source": "ncu smsp__inst_executed.sum, profiles/traffic.json"}
        line = {
            "metric": "env steps/sec (batched)", "value": value, "unit": "env-steps/s", "n_gpus": world,
            "steps": K, "warmup": Wm, "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": config_dict(args, world),
            "repeats": {"regions": R, "region_ms_min": float(region_ms.min()), "region_ms_median": dev_ms,
                        "region_ms_max": float(region_ms.max())},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "kernel": KERNEL_NAMES.get(WORKLOAD["prob"], "k_rollout_async<%s>" % WORKLOAD["prob"]),
                         "algorithmic_bytes_per_env_step": bpe, "units_per_launch": n * chunk,
                         "avg_launch_ms": avg_launch_s * 1e3, "issue": issue},
            "wall_s_rollout_region": wall_rollout, "wall_s_total": time.perf_counter() - t_start,
        }
        if not args.only_rollout:
            from gym_pcgrl_b200 import HostStepIO  # noqa: F401
            e2e_value = total_envs * K / (res["e2e"] * 1e-3)
            line["closed_loop"] = {
                "value": total_envs * K / ((res["graph"] or res["plain"]) * 1e-3), "unit": "env-steps/s",
                "plain": {"value": total_envs * K / (res["plain"] * 1e-3), "ms_per_step": res["plain"] / K,
                          "launches_per_step": 2},
                "graph": None if res["graph"] is None else {"value": total_envs * K / (res["graph"] * 1e-3),
                                                           "ms_per_step": res["graph"] / K},
                "step_kernel_ms": res["kernel"],
                "step_kernel_hbm_frac": bpe * n / (res["kernel"] * 1e-3) / 1e9 / peak,
                "api": "per step: torch.randint on the device -> pcgrl_step (T=1 launch of k_rollout); obs / reward / "
                       "done stay in HBM where a policy would read them",
                "graph_error": getattr(B, "graph_error", None)}
            # direct transport: bytes the kernel stores into the host arrays per step = reward + done + cursor of every env,
            # map cell + heat count of every edited env, map + cleared heat map of every auto-reset env (the per-step counts
            # of edited / reset envs are the change-record counters of the delta run of the same workload just below)
            hb = env._tens["heatmap"].element_size()
            edited = max(0.0, io_delta.records_per_step - io_delta.resets_per_step)
            direct_bytes = io.d2h_bytes + edited * (1 + hb) + io_delta.resets_per_step * W * H * (1 + hb)
            line["e2e"] = {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": io.h2d_bytes * world,
                           "d2h_bytes_per_step": int(round(direct_bytes)) * world, "ms_per_step": res["e2e"] / K,
                           "edited_envs_per_step": edited, "reset_envs_per_step": io_delta.resets_per_step,
                           "api": "pcgrl_step_host mode 2 (direct transport: pinned, device-mapped host buffers; the step kernel reads "
                                  "the host actions and stores reward / done / cursor of every env, the edited map cell + its heat count "
                                  "and the fresh map of auto-reset envs straight into the host arrays, which hold the complete "
                                  "map+heatmap+pos+reward+done after every step)"}
            line["e2e_delta"] = {"value": total_envs * K / (res["e2e_delta"] * 1e-3), "unit": "env-steps/s",
                                 "h2d_bytes_per_step": io_delta.h2d_bytes * world, "d2h_bytes_per_step": io_delta.d2h_bytes * world,
                                 "ms_per_step": res["e2e_delta"] / K,
                                 "api": "pcgrl_step_host mode 1 (per-env delta records + fresh maps of reset envs in one D2H copy, "
                                        "applied to the host arrays by the library)"}
            line["e2e_full_copy"] = {"value": total_envs * K / (res["e2e_full"] * 1e-3), "unit": "env-steps/s",
                                     "h2d_bytes_per_step": io_full.h2d_bytes * world, "d2h_bytes_per_step": io_full.d2h_bytes * world,
                                     "ms_per_step": res["e2e_full"] / K, "api": "pcgrl_step_host mode 0 (every array copied back in full)"}
            line["e2e_rollout"] = {"value": total_envs * roll_steps / (res["e2e_roll"] * 1e-3), "unit": "env-steps/s",
                                   "h2d_bytes_per_step": rio.h2d_bytes * world // chunk, "d2h_bytes_per_step": rio.d2h_bytes * world // chunk,
                                   "ms_per_step": res["e2e_roll"] / roll_steps, "steps_per_call": chunk,
                                   "api": "pcgrl_rollout_host (open loop: %d steps per call, pinned host actions in; every step's "
                                          "reward + done copied back, the final map+heatmap+pos stored into the pinned host arrays "
                                          "by the kernel as each env finishes)" % chunk}
            line["check_reward_sum"] = rsum
            line["shard_check"] = shard
            if world > 1:
                line["gather"] = {"value": total_envs * K / (gather * 1e-3), "unit": "env-steps/s",
                                  "what": "open-loop rollout + NCCL all_gather_into_tensor of every launch's reward (f64) and done (u8) "
                                          "rows to every rank, inside the timed launches",
                                  "bytes_per_rank_per_launch": chunk * n * 9 * world}
            line["sweep"] = sweep
        cores = len(os.sched_getaffinity(0))
        if world == 1 and not args.no_cpu and not args.only_rollout:
            # CPU baselines are timed at N=1 only (other ranks would be spinning on the same cores)
            cpu_value, cpu_steps, cpu_s = cpu_port_run(n, args.cpu_seconds, cores)
            line["cpu_baseline_port"] = {"value": cpu_value, "unit": "env-steps/s", "cores": cores, "kind": "port",
                                         "sample": "%d batched steps x %d envs in %.1f s, oracle/pcgrl_oracle.c, %d OpenMP threads"
                                                   % (cpu_steps, n, cpu_s, cores)}
            if reference_available():
                probe = cpu_reference_run(min(64, n), 1, 6, cores, 30.0)["env_steps_per_s"]
                ne = int(min(n, 1024))
                st = int(max(4, probe * 12.0 / ne))
                r = cpu_reference_run(ne, 2, st, cores, 40.0)
                line["cpu_baseline"] = {"value": r["env_steps_per_s"], "unit": "env-steps/s", "cores": r["procs"], "kind": "reference",
                                        "single_core": r["single_core_env_steps_per_s"],
                                        "sample": "%d batched steps x %d envs (of the %d-env batch) in %.1f s: unmodified Python reference "
                                                  "(oracle/_ref/gym_pcgrl PcgrlEnv.step), one worker process per core, reset on done"
                                                  % (r["steps"], ne, n, r["seconds"])}
            else:
                line["cpu_baseline"] = line["cpu_baseline_port"]
        else:
            line["cpu_baseline"] = None
        line["clocks"] = clocks
        line["wall_s_total"] = time.perf_counter() - t_start
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

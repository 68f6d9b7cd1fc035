#!/usr/bin/env python
"""Time the UNMODIFIED reference's PcgrlEnv.step on the host cores (one process per core).

MEASUREMENT INFRASTRUCTURE, NOT PRODUCT: only bench.py's cpu_baseline / `--impl reference` legs run this, as a
separate process (it forks workers, which must not happen inside a process that holds a CUDA context).  It imports
the reference package from oracle/_ref/ (made by oracle/make_ref.py; /root/reference does not exist on the GPU box)
under the gym stand-in of ref_shim.py and mirrors the reference's own data-parallel layout: SubprocVecEnv, one OS
process per env group (utils.py:60-71, train.py:103), every env stepped with uniform random actions and reset on
done (README.md:59-72).

    python oracle/ref_runner.py '{"prob": "binary", "rep": "narrow", "kwargs": {...}, "envs": 4096,
                                  "warmup": 5, "steps": 20, "procs": 16, "max_seconds": 60}'
prints one JSON line: {"env_steps_per_s": ..., "envs": ..., "steps": ..., "procs": ..., "seconds": ..., ...}.

Timing: every worker steps its slice of the env batch for `warmup` untimed batched steps, waits at a barrier, then
runs `steps` batched steps (stopping early at `max_seconds`, reported); value = envs x steps / slowest worker's time.
"""
import json
import multiprocessing as mp
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.path.join(HERE, "_ref")


def _worker(spec, lo, hi, barrier, out, idx):
    try:
        out.put((idx, _work(spec, lo, hi, barrier)))
    except BaseException as e:   # a dead worker must not leave the others waiting at the barrier
        try:
            barrier.abort()
        except Exception:
            pass
        out.put((idx, ("error", repr(e))))


def _work(spec, lo, hi, barrier):
    os.environ["PCGRL_REFERENCE_ROOT"] = REF_ROOT
    os.environ["OMP_NUM_THREADS"] = "1"
    sys.path.insert(0, REF_ROOT)
    import numpy as np
    import ref_shim
    ref_shim.install()
    import gym
    env_id = "%s-%s-v0" % (spec["prob"], spec["rep"])
    envs, rngs = [], []
    for i in range(lo, hi):
        env = gym.make(env_id)
        if spec["kwargs"]:
            env.adjust_param(**spec["kwargs"])
            env.adjust_param(**spec["kwargs"])   # quirk Q3, as in every harness of this repo
        env._rep._random = np.random.RandomState(i)
        env._prob._random = np.random.RandomState(i)
        env.reset()
        envs.append(env)
        rngs.append(np.random.RandomState(1_000_000 + i))
    space = envs[0].action_space
    nvec = [int(v) for v in space.nvec] if hasattr(space, "nvec") else None

    def batched_step():
        for env, rng in zip(envs, rngs):
            a = [int(rng.randint(k)) for k in nvec] if nvec else int(rng.randint(space.n))
            _, _, done, _ = env.step(a)
            if done:
                env.reset()

    for _ in range(spec["warmup"]):
        batched_step()
    barrier.wait()
    t0 = time.perf_counter()
    done_steps = 0
    for _ in range(spec["steps"]):
        batched_step()
        done_steps += 1
        if time.perf_counter() - t0 > spec["max_seconds"]:
            break
    return done_steps, time.perf_counter() - t0, hi - lo


def run(spec):
    procs = max(1, min(int(spec["procs"]), int(spec["envs"])))
    n = int(spec["envs"])
    bounds = [(n * p) // procs for p in range(procs + 1)]
    ctx = mp.get_context("fork")
    barrier, out = ctx.Barrier(procs), ctx.Queue()
    t_setup = time.perf_counter()
    ps = [ctx.Process(target=_worker, args=(spec, bounds[p], bounds[p + 1], barrier, out, p)) for p in range(procs)]
    for p in ps:
        p.start()
    res = [out.get() for _ in ps]
    for p in ps:
        p.join()
    bad = [r for _, r in res if r[0] == "error"]
    if bad:
        raise RuntimeError("reference worker failed: %s" % bad[0][1])
    res = [r for _, r in sorted(res)]
    # every worker ran the same number of batched steps unless one hit max_seconds: count env-steps actually done
    env_steps = sum(s * m for s, _, m in res)
    slowest = max(t for _, t, _ in res)
    return {"env_steps_per_s": env_steps / slowest, "env_steps": env_steps, "envs": n,
            "steps": min(s for s, _, _ in res), "procs": procs, "seconds": slowest,
            "wall_seconds_incl_setup": time.perf_counter() - t_setup,
            "single_core_env_steps_per_s": env_steps / slowest / procs}


if __name__ == "__main__":
    spec = json.loads(sys.argv[1])
    spec.setdefault("kwargs", {})
    spec.setdefault("warmup", 2)
    spec.setdefault("max_seconds", 60.0)
    spec.setdefault("procs", len(os.sched_getaffinity(0)))
    print(json.dumps(run(spec)), flush=True)

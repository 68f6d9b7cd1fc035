#!/usr/bin/env python
"""Where does a pcgrl_step_host call spend its time?  (binary-narrow 16x16, 4096 envs)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from gym_pcgrl_b200 import HostStepIO

n, K = 4096, 1000
env = bench.make_env(n, "cuda:0", 0)
env.reset()
acts_h = torch.from_numpy(bench.host_actions(env, K + 8, n, 5)).pin_memory()
acts_d = acts_h.cuda()
base, stride = acts_h.data_ptr(), acts_h.stride(0) * 4

def timeit(name, fn, reps=K):
    for t in range(8): fn(t)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for t in range(reps): fn(8 + t)
    torch.cuda.synchronize()
    print("%-44s %7.1f us/step" % (name, 1e6 * (time.perf_counter() - t0) / reps), flush=True)

def dev_step_sync(t):
    env.step(acts_d[t]); torch.cuda.synchronize()
timeit("env.step(device actions) + synchronize", dev_step_sync)
def dev_step_nosync(t):
    env.step(acts_d[t])
timeit("env.step(device actions), no sync (enqueue)", dev_step_nosync)
for mode in ("delta", "full"):
    io = HostStepIO(env, with_obs=True, mode=mode)
    def f(t, io=io):
        io.struct.actions = base + t * stride; env.step_host(io)
    timeit("step_host mode=%s with obs" % mode, f)
io = HostStepIO(env, with_obs=False, mode="delta")
def f2(t):
    io.struct.actions = base + t * stride; env.step_host(io)
timeit("step_host mode=delta, reward/done only", f2)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
torch.cuda.synchronize(); ev[0].record()
for t in range(200): env.step(acts_d[t])
ev[1].record(); torch.cuda.synchronize()
print("device time per env.step launch (back-to-back, warm L2): %.1f us" % (1e3 * ev[0].elapsed_time(ev[1]) / 200))

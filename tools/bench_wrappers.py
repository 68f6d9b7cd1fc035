#!/usr/bin/env python
"""HBM roofline of the fused observation kernel (pcgrl_obs_image): bytes written / event time vs the measured peak."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from gym_pcgrl_b200 import wrappers as W

peak, src = bench.measured_peak()
rows = []
for env_id, crop, n, dt in [("zelda-narrow-v0", 22, 4096, "float32"), ("zelda-narrow-v0", 22, 4096, "uint8"),
                            ("binary-narrow-v0", 28, 4096, "uint8"), ("zelda-narrow-v0", 22, 65536, "float32"),
                            ("binary-narrow-v0", 28, 65536, "uint8")]:
    env = W.CroppedImagePCGRLWrapper(env_id, crop, num_envs=n, out_dtype=dt)
    env.reset()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(5):
        env._image()
    ts = []
    for _ in range(20):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = env._image(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    t = sorted(ts)[len(ts) // 2]
    nbytes = out.numel() * out.element_size() + env.pcgrl_env._tens["map"].numel()
    rows.append(dict(env=env_id, crop=crop, envs=n, dtype=dt, out_mb=out.numel() * out.element_size() / 1e6, us=t * 1e6,
                     gbs=nbytes / t / 1e9, frac=nbytes / t / 1e9 / peak, peak=peak, peak_source=src))
    print(json.dumps(rows[-1]), flush=True)

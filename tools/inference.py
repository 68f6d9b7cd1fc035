#!/usr/bin/env python
"""Run a trained agent and collect the generated maps (the reference's inference.py, batched): every env plays ONE episode
with the saved policy; prints the terminal statistics of the batch and optionally saves the final maps / renders.

    python tools/train_ppo.py --game binary --rep narrow --timesteps 5e6 --save /tmp/binary_narrow.pt
    python tools/inference.py --game binary --rep narrow --model /tmp/binary_narrow.pt --envs 256 [--change-percentage 0.4]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from gym_pcgrl_b200.ppo import PPO, make_training_env


def infer(game, representation, model_path, num_envs=256, change_percentage=0.4, native_policy=True, out=None, deterministic=False, seed=1):
    env = make_training_env(game, representation, num_envs, seed=seed)
    env.pcgrl_env.adjust_param(change_percentage=change_percentage)     # inference.py:44-48 kwargs -> adjust_param
    agent = PPO(env, native_policy=native_policy).load(model_path)
    base = env.pcgrl_env
    base.auto_reset = False                                             # keep every env's final map
    base._cfg = None
    obs = env.reset()
    n = num_envs
    finished = torch.zeros(n, dtype=torch.bool, device=obs.device)
    ret = torch.zeros(n, dtype=torch.float64, device=obs.device)
    final_info, steps = {}, 0
    final_maps = torch.zeros_like(base._tens["map"])
    while not bool(finished.all()) and steps < 100000:
        obs, reward, done, info = env.step(agent.predict(obs, deterministic))
        ret += torch.where(finished, torch.zeros_like(reward), reward)
        newly = done & ~finished
        final_maps = torch.where(newly[:, None, None], base._tens["map"], final_maps)    # the map at the end of the env's episode
        for k, v in info.items():
            if torch.is_tensor(v):
                final_info.setdefault(k, torch.zeros_like(v))
                final_info[k] = torch.where(newly, v, final_info[k])
        finished |= done
        steps += 1
    print("%d episodes, %d batched steps, mean return %.2f" % (n, steps, float(ret.mean())))
    for k, v in final_info.items():
        print("  %-16s mean %.2f  min %d  max %d" % (k, float(v.double().mean()), int(v.min()), int(v.max())))
    if out:
        np.savez_compressed(out, maps=final_maps.cpu().numpy(), returns=ret.cpu().numpy(),
                            render=base._prob.render(final_maps[: min(n, 16)].contiguous(), None).cpu().numpy())
        print("saved maps, returns and 16 renders to", out)
    return ret


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--game", default="binary")
    ap.add_argument("--rep", default="narrow")
    ap.add_argument("--model", required=True)
    ap.add_argument("--envs", type=int, default=256)
    ap.add_argument("--change-percentage", type=float, default=0.4)
    ap.add_argument("--deterministic", action="store_true")
    ap.add_argument("--torch-policy", action="store_true", help="run the torch modules instead of the native kernels")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    infer(a.game, a.rep, a.model, a.envs, a.change_percentage, not a.torch_policy, a.out, a.deterministic)

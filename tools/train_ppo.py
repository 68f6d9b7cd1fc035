#!/usr/bin/env python
"""Device-resident PPO on the batched environment (the reference's train.py with its defaults: game 'binary',
representation 'narrow').   python tools/train_ppo.py [--game binary] [--rep narrow] [--envs 4096] [--timesteps 2e6]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_pcgrl_b200.ppo import PPO, make_training_env


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--game", default="binary")
    ap.add_argument("--rep", default="narrow")
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--timesteps", type=float, default=2e6)
    ap.add_argument("--n-steps", type=int, default=128)
    ap.add_argument("--native-policy", action="store_true", help="rollout inference through im2col + the tcgen05 GEMM kernel")
    ap.add_argument("--save", default=None, help="write the trained policy here (tools/inference.py loads it)")
    a = ap.parse_args()
    env = make_training_env(a.game, a.rep, a.envs)
    agent = PPO(env, n_steps=a.n_steps, native_policy=a.native_policy).learn(int(a.timesteps))
    if a.save:
        agent.save(a.save)
        print("saved", a.save)


if __name__ == "__main__":
    main()

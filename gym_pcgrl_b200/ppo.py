"""A device-resident PPO consumer for the batched environment (replaces the reference's
stable-baselines / TF1 ``PPO2(policy, SubprocVecEnv(...)).learn()``, train.py:51-95).

Everything between two log lines stays on the GPU: the wrapped env steps N lock-step envs through the fused kernels and
writes the policy input tensor, the policy samples actions from it, and the rollout buffers, GAE and the clipped-surrogate
updates are torch ops on those tensors -- no host round trip per step.  Hyper-parameters default to stable-baselines
PPO2's (gamma 0.99, n_steps 128, ent_coef 0.01, lr 2.5e-4, vf_coef 0.5, max_grad_norm 0.5, lam 0.95, 4 minibatches,
4 epochs, cliprange 0.2, value clipping with the same range, Adam eps 1e-5)."""
import time

import torch

from .models import CROPPED_SIZE, ActorCritic, policy_for
from .wrappers import ActionMapImagePCGRLWrapper, CroppedImagePCGRLWrapper


def make_training_env(game, representation, num_envs, device="cuda", seed=0, **kwargs):
    """utils.py:48-58 make_env: ActionMapImage wrapper for wide, CroppedImage (28 / 22 / 10) otherwise."""
    env_id = "%s-%s-v0" % (game, representation)
    if representation == "wide":
        return ActionMapImagePCGRLWrapper(env_id, num_envs=num_envs, device=device, env_kwargs=dict(seed=seed), **kwargs)
    crop = kwargs.pop("cropped_size", CROPPED_SIZE.get(game, 28))
    return CroppedImagePCGRLWrapper(env_id, crop, num_envs=num_envs, device=device, env_kwargs=dict(seed=seed), **kwargs)


class PPO:
    def __init__(self, env, policy=None, n_steps=128, gamma=0.99, lam=0.95, ent_coef=0.01, vf_coef=0.5, learning_rate=2.5e-4,
                 max_grad_norm=0.5, nminibatches=4, noptepochs=4, cliprange=0.2, seed=0, native_policy=False):
        self.env = env
        base = env.pcgrl_env
        self.device = torch.device(base.device if str(base.device) != "cuda" else "cuda:%d" % torch.cuda.current_device())
        self.n_envs, self.n_steps = base.num_envs, n_steps
        n_actions = int(env.action_space.n)
        kind = policy or policy_for(base._prob.name, base._rep.name)
        torch.manual_seed(seed)
        self.policy = ActorCritic(kind, env.shape, n_actions).to(self.device)
        self.opt = torch.optim.Adam(self.policy.parameters(), lr=learning_rate, eps=1e-5)
        # rollout-time inference through this repo's own kernels (policy_native.py: im2col + tcgen05 GEMM); the updates
        # keep the autograd modules, the bf16 weight snapshot is refreshed after every update
        self.native = None
        if native_policy:
            from .policy_native import NativePolicy
            self.native = NativePolicy(self.policy)
        self.gamma, self.lam, self.ent_coef, self.vf_coef = gamma, lam, ent_coef, vf_coef
        self.max_grad_norm, self.nminibatches, self.noptepochs, self.cliprange = max_grad_norm, nminibatches, noptepochs, cliprange
        T, N = n_steps, self.n_envs
        self.obs_buf = torch.empty((T, N) + tuple(env.shape), dtype=torch.uint8, device=self.device)
        self.act_buf = torch.empty((T, N), dtype=torch.int64, device=self.device)
        self.logp_buf = torch.empty((T, N), dtype=torch.float32, device=self.device)
        self.val_buf = torch.empty((T, N), dtype=torch.float32, device=self.device)
        self.rew_buf = torch.empty((T, N), dtype=torch.float32, device=self.device)
        self.done_buf = torch.empty((T, N), dtype=torch.bool, device=self.device)
        self.obs = None
        self.ep_ret = torch.zeros(N, dtype=torch.float64, device=self.device)
        self.finished_returns = []
        self.num_timesteps = 0

    @torch.no_grad()
    def collect(self):
        if self.obs is None:
            self.obs = self.env.reset()
        infer = self.native if self.native is not None else self.policy
        for t in range(self.n_steps):
            logits, value = infer(self.obs)
            dist = torch.distributions.Categorical(logits=logits)
            action = dist.sample()
            self.obs_buf[t].copy_(self.obs)
            self.act_buf[t], self.logp_buf[t], self.val_buf[t] = action, dist.log_prob(action), value
            self.obs, reward, done, _ = self.env.step(action.to(torch.int32))
            self.rew_buf[t], self.done_buf[t] = reward.float(), done
            self.ep_ret += reward
            if bool(done.any()):                              # one small D2H per step with finished episodes (logging only)
                self.finished_returns.append(self.ep_ret[done].clone())
                self.ep_ret[done] = 0
        _, last_value = infer(self.obs)
        adv = torch.empty_like(self.rew_buf)
        lastgae = torch.zeros(self.n_envs, device=self.device)
        for t in reversed(range(self.n_steps)):               # GAE(lambda); done[t] ends the episode AFTER step t
            nonterminal = (~self.done_buf[t]).float()
            next_value = last_value if t == self.n_steps - 1 else self.val_buf[t + 1]
            delta = self.rew_buf[t] + self.gamma * next_value * nonterminal - self.val_buf[t]
            lastgae = delta + self.gamma * self.lam * nonterminal * lastgae
            adv[t] = lastgae
        self.num_timesteps += self.n_steps * self.n_envs
        return adv, adv + self.val_buf

    def update(self, adv, returns):
        T, N = self.n_steps, self.n_envs
        B = T * N
        obs, act = self.obs_buf.reshape((B,) + tuple(self.env.shape)), self.act_buf.reshape(B)
        old_logp, old_val, adv, returns = self.logp_buf.reshape(B), self.val_buf.reshape(B), adv.reshape(B), returns.reshape(B)
        mb = B // self.nminibatches
        stats = torch.zeros(4, device=self.device)
        for _ in range(self.noptepochs):
            perm = torch.randperm(B, device=self.device)
            for k in range(self.nminibatches):
                idx = perm[k * mb:(k + 1) * mb]
                logits, value = self.policy(obs[idx])
                dist = torch.distributions.Categorical(logits=logits)
                a = adv[idx]
                a = (a - a.mean()) / (a.std() + 1e-8)
                ratio = torch.exp(dist.log_prob(act[idx]) - old_logp[idx])
                pg_loss = torch.max(-a * ratio, -a * torch.clamp(ratio, 1 - self.cliprange, 1 + self.cliprange)).mean()
                v_clipped = old_val[idx] + torch.clamp(value - old_val[idx], -self.cliprange, self.cliprange)
                vf_loss = 0.5 * torch.max((value - returns[idx]) ** 2, (v_clipped - returns[idx]) ** 2).mean()
                entropy = dist.entropy().mean()
                loss = pg_loss - self.ent_coef * entropy + self.vf_coef * vf_loss
                self.opt.zero_grad(set_to_none=True)
                loss.backward()
                nn_utils_clip(self.policy.parameters(), self.max_grad_norm)
                self.opt.step()
                stats += torch.stack([pg_loss.detach(), vf_loss.detach(), entropy.detach(), loss.detach()])
        if self.native is not None:
            self.native.refresh()
        return (stats / (self.noptepochs * self.nminibatches)).tolist()

    def learn(self, total_timesteps, log_every=1, log=print):
        t0, updates = time.perf_counter(), 0
        while self.num_timesteps < total_timesteps:
            adv, returns = self.collect()
            pg, vf, ent, loss = self.update(adv, returns)
            updates += 1
            if log and updates % log_every == 0:
                rets = torch.cat(self.finished_returns) if self.finished_returns else torch.zeros(0)
                self.finished_returns = []
                torch.cuda.synchronize(self.device)
                fps = self.num_timesteps / (time.perf_counter() - t0)
                log("update %d  timesteps %d  fps %.0f  ep_rew_mean %.2f (%d episodes)  pg %.4f  vf %.4f  entropy %.3f" % (
                    updates, self.num_timesteps, fps, float(rets.mean()) if rets.numel() else float("nan"), rets.numel(), pg, vf, ent))
        return self


    # ---- the reference's model.save / PPO2.load / agent.predict (train.py:40-48, inference.py:28-36)
    def save(self, path):
        base = self.env.pcgrl_env
        torch.save({"policy": self.policy.state_dict(), "kind": self.policy.kind, "obs_shape": tuple(self.env.shape),
                    "n_actions": int(self.env.action_space.n), "game": base._prob.name, "representation": base._rep.name,
                    "num_timesteps": self.num_timesteps}, path)

    def load(self, path):
        """Load weights saved by ``save`` into this learner (same game / representation / observation shape)."""
        ck = torch.load(path, map_location=self.device)
        if ck["kind"] != self.policy.kind or tuple(ck["obs_shape"]) != tuple(self.env.shape) or ck["n_actions"] != int(self.env.action_space.n):
            raise ValueError("checkpoint is for %s %s, this learner is %s %s" % (ck["kind"], ck["obs_shape"], self.policy.kind, tuple(self.env.shape)))
        self.policy.load_state_dict(ck["policy"])
        self.num_timesteps = int(ck.get("num_timesteps", 0))
        if self.native is not None:
            self.native.refresh()
        return self

    @torch.no_grad()
    def predict(self, obs, deterministic=False):
        """agent.predict(obs): actions int32 [N] for a batch of observations (sampled like stable-baselines unless
        deterministic)."""
        infer = self.native if self.native is not None else self.policy
        logits, _ = infer(obs)
        a = logits.argmax(dim=1) if deterministic else torch.distributions.Categorical(logits=logits).sample()
        return a.to(torch.int32)


def nn_utils_clip(params, max_norm):
    torch.nn.utils.clip_grad_norm_(params, max_norm)

"""Asynchronous grouped stepping: the opt-in ``step_async`` / ``step_wait`` form of the batched environment.

The reference's data-parallel layer is stable-baselines' ``SubprocVecEnv`` (utils.py:60-71): ``step_async(actions)``
hands every worker its action, ``step_wait()`` blocks until ALL workers have answered.  A synchronous batched step has
the same property -- it lasts as long as its slowest env -- and for the solver problems (sokoban / ddave / mdungeon,
smb) the slowest env of a large batch is almost always one stuck in a capped A* search that takes milliseconds, while
the other 99 % finish in microseconds.  ``AsyncGroupedEnv`` shards the batch into G env groups (global env indices are
kept, so every env follows exactly the trajectory it has in one big synchronous batch), gives each group its own CUDA
stream and pinned host buffers and exposes

    send(g, actions)   -> pcgrl_step_host_begin on the group's stream (returns at once)
    recv(wait=True)    -> the groups whose step has completed (pcgrl_step_host_end), host arrays updated

so a caller keeps every group in flight and serves whichever finishes first: a capped search delays its own group
only.  Per-env results are bit-identical to the synchronous call; only the batching changes.
"""
import ctypes as C

import numpy as np

from . import _native
from .envs.pcgrl_env import BatchedPcgrlEnv, HostStepIO


class AsyncGroupedEnv:
    def __init__(self, prob="binary", rep="narrow", num_envs=1024, groups=16, device="cuda", seed=None, auto_reset=True,
                 env_offset=0, with_info=False, mode="delta"):
        if num_envs % groups:
            raise ValueError("num_envs must be a multiple of groups")
        self.num_envs, self.groups, self.per_group = int(num_envs), int(groups), int(num_envs) // int(groups)
        self.env_offset = int(env_offset)
        self._with_info, self._mode = with_info, mode
        self.envs = [BatchedPcgrlEnv(prob, rep, num_envs=self.per_group, device=device, seed=seed, auto_reset=auto_reset,
                                     env_offset=self.env_offset + g * self.per_group) for g in range(self.groups)]
        self.io = [None] * self.groups
        self.streams = None
        self.in_flight = [False] * self.groups
        self.steps_done = [0] * self.groups
        self._order = []   # groups in flight, oldest first

    # ---- reference surface, fanned out
    @property
    def action_space(self):
        return self.envs[0].action_space

    @property
    def observation_space(self):
        return self.envs[0].observation_space

    def adjust_param(self, **kwargs):
        for e in self.envs:
            e.adjust_param(**kwargs)

    def set_rng_states(self, states):
        states = np.asarray(states)
        for g, e in enumerate(self.envs):
            e.set_rng_states(states[g * self.per_group:(g + 1) * self.per_group])

    def reset(self):
        """Reset every group (each on its own stream) and arm the host transport; returns the per-group host buffers."""
        import torch
        if any(self.in_flight):
            raise RuntimeError("reset() with steps in flight: recv() them first")
        if self.streams is None:
            for e in self.envs:
                e._ensure_buffers()
            dev = self.envs[0]._dev
            torch.cuda.synchronize(dev)   # allocations were made on the current stream, the groups run on their own
            self.streams = [torch.cuda.Stream(device=dev) for _ in range(self.groups)]
        for g, e in enumerate(self.envs):
            with torch.cuda.stream(self.streams[g]):
                e.reset()
                if self.io[g] is None:
                    self.io[g] = HostStepIO(e, with_obs=True, with_info=self._with_info, mode=self._mode)
                self.io[g].invalidate()
                # the observation after reset: map / heatmap / pos to the host arrays
                self.io[g].map.copy_(e._tens["map"], non_blocking=True)
                self.io[g].heatmap.copy_(e._tens["heatmap"], non_blocking=True)
                if self.io[g].pos is not None:
                    self.io[g].pos.copy_(e._tens["pos"], non_blocking=True)
        for s in self.streams:
            s.synchronize()
        self.steps_done = [0] * self.groups
        return self.io

    # ---- asynchronous stepping
    def send(self, g, actions=None):
        """Start one step of group g.  actions: int32 [per_group(, k)] array / tensor, or None when the caller has
        already written them into ``self.io[g].actions`` (pinned)."""
        if self.in_flight[g]:
            raise RuntimeError("group %d already has a step in flight" % g)
        e, io = self.envs[g], self.io[g]
        if actions is not None:
            import torch
            io.actions.copy_(torch.as_tensor(actions, dtype=torch.int32).reshape(io.actions.shape))
        if e._cfg is None:
            e.native_config
        rc = _native.lib().pcgrl_step_host_begin(C.addressof(e._cfg), C.addressof(e._cbufs), e._d_actions_ptr, C.addressof(io.struct),
                                                 e.num_envs, self.streams[g].cuda_stream)
        if rc:
            _native.check(rc, "pcgrl_step_host_begin")
        self.in_flight[g] = True
        self._order.append(g)

    def _end(self, g, wait):
        e, io = self.envs[g], self.io[g]
        rc = _native.lib().pcgrl_step_host_end(C.addressof(e._cfg), C.addressof(e._cbufs), C.addressof(io.struct), e.num_envs,
                                               self.streams[g].cuda_stream, 1 if wait else 0)
        if rc == 1 and not wait:
            return False
        if rc:
            _native.check(rc, "pcgrl_step_host_end")
        self.in_flight[g] = False
        self._order.remove(g)
        self.steps_done[g] += 1
        return True

    def recv(self, wait=True):
        """Groups whose step has completed (their ``io[g]`` host arrays now hold map / heatmap / pos / reward / done).
        With wait=True and nothing ready yet, blocks on the OLDEST step in flight."""
        ready = [g for g in list(self._order) if self._end(g, False)]
        if not ready and wait and self._order:
            g = self._order[0]
            self._end(g, True)
            ready = [g]
        return ready

    def step(self, actions):
        """Synchronous convenience: one step of every group; actions [num_envs(, k)].  Returns per-group io blocks."""
        a = np.asarray(actions)
        for g in range(self.groups):
            self.send(g, a[g * self.per_group:(g + 1) * self.per_group])
        while any(self.in_flight):
            self.recv(wait=True)
        return self.io

    def check_status(self):
        for e in self.envs:
            e.check_status()

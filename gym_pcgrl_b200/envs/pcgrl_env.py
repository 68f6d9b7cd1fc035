"""The gym.Env boundary of the hot path, batched.

``BatchedPcgrlEnv`` keeps the surface of the reference's ``PcgrlEnv``
(gym_pcgrl/envs/pcgrl_env.py:12-182: ``reset / step / seed / adjust_param / get_num_tiles /
get_border_tile``, ``action_space`` / ``observation_space``, ``_prob`` / ``_rep`` plugins chosen by
name from PROBLEMS / REPRESENTATIONS) but owns ``num_envs`` lock-step environment instances whose
state lives in HBM and whose whole ``step()`` runs in hand-written sm_100a kernels behind the C ABI
of include/pcgrl_b200.h.  Observations, rewards, dones and info come back as CUDA tensors.

Deliberate differences to the single-env reference (documented in DESIGN.md):
* observation tensors are live views of the state (``clone()`` to keep one); the reference returns copies;
* ``heatmap`` is uint8 (value-equal to the reference's float64 counts);
* ``reward`` is float64 [N], accumulated in the reference's term order;
* with ``auto_reset=True`` (VecEnv semantics, utils.py:64-68 role) an env that finishes is reset
  inside ``step`` and the returned observation is the first one of the next episode; the terminal
  statistics stay in ``info``.

``PcgrlEnv`` is the single-environment classic-gym facade (numpy in / numpy out) over a batch of one.
"""
import ctypes as C

import numpy as np

from .. import _abi, _native, spaces
from .._config import build_config
from ..seeding import create_seed, hash_seed_words, mt_state_words
from .probs import PROBLEMS
from .reps import REPRESENTATIONS


class BatchedPcgrlEnv:
    metadata = {'render.modes': []}

    def __init__(self, prob="binary", rep="narrow", num_envs=1, device="cuda", seed=None,
                 auto_reset=True, env_offset=0):
        self._prob = PROBLEMS[prob]()            # KeyError for unknown names, like the reference
        self._rep = REPRESENTATIONS[rep]()
        self.num_envs = int(num_envs)
        self.device = device
        self.auto_reset = bool(auto_reset)
        self.env_offset = int(env_offset)        # global index of env 0 (multi-GPU sharding)
        # pcgrl_env.py:33-34
        self._max_changes = max(int(0.2 * self._prob._width * self._prob._height), 1)
        self._max_iterations = self._max_changes * self._prob._width * self._prob._height
        self._tens = None
        self._cbufs = None
        self._cfg = None
        self._d_actions = None
        self._pending_states = None
        self._pending_probs = None
        self._base_seed = None
        self._update_spaces()
        self.seed(seed)

    # ------------------------------------------------------------------ reference surface
    def seed(self, seed=None):
        """pcgrl_env.py:54-57: both RNG streams of an env get the same seed.  Env i (global index
        ``env_offset + i``) is seeded with ``seed + env_offset + i`` through gym's seed hashing."""
        seed = self._rep.seed(seed)
        self._prob.seed(seed)
        self._base_seed = seed
        states = np.empty((self.num_envs, 2, _abi.MT_WORDS), np.uint32)
        rng = np.random.RandomState()
        for i in range(self.num_envs):
            rng.seed(hash_seed_words(create_seed(seed + self.env_offset + i)))
            states[i, 0] = states[i, 1] = mt_state_words(rng)
        self.set_rng_states(states)
        return [seed]

    def set_rng_states(self, states):
        """Inject MT19937 states: uint32 [N,2,625] ([:,0] representation stream, [:,1] problem
        stream) or [N,625] for both.  The parity harness uses this with ``RandomState(k)`` states."""
        states = np.asarray(states, dtype=np.uint32)
        if states.ndim == 2:
            states = np.repeat(states[:, None, :], 2, axis=1)
        assert states.shape == (self.num_envs, 2, _abi.MT_WORDS), states.shape
        if self._tens is None:
            self._pending_states = states.copy()
        else:
            import torch
            self._tens["rng"].copy_(torch.from_numpy(states.view(np.int32)))

    def seed_simple(self, seeds):
        """Seed both streams of env i with numpy's ``RandomState(seeds[i])`` on the device (pcgrl_seed)."""
        import torch
        self._ensure_buffers()
        s32 = torch.from_numpy((np.asarray(seeds, dtype=np.int64) & 0xFFFFFFFF).astype(np.uint32).view(np.int32)).to(self._dev)
        assert s32.numel() == self.num_envs
        with torch.cuda.device(self._dev):
            _native.check(_native.lib().pcgrl_seed(C.byref(self._cbufs), s32.data_ptr(), self.num_envs,
                                                   _native.stream_ptr(self._dev)), "pcgrl_seed")

    def adjust_param(self, **kwargs):
        """pcgrl_env.py:106-115, including its ordering quirk: the limits are computed
        from the problem's size BEFORE the problem applies the new width / height."""
        if 'change_percentage' in kwargs:
            percentage = min(1, max(0, kwargs.get('change_percentage')))
            self._max_changes = max(int(percentage * self._prob._width * self._prob._height), 1)
        self._max_iterations = self._max_changes * self._prob._width * self._prob._height
        old_shape = (self._prob._height, self._prob._width)
        self._prob.adjust_param(**kwargs)
        self._rep.adjust_param(**kwargs)
        self._update_spaces()
        self._cfg = None
        if self._tens is not None:
            probs = kwargs.get('probs')
            if probs is not None:
                # problem.py:66-72: Problem._prob is per-env state here (binary redraws it at every reset,
                # binary_prob.py:68-72), so the given keys -- and only those -- are overwritten for every env;
                # the next reset() draws its map from them
                tiles = self._prob.get_tile_types()
                for t, v in probs.items():
                    if t in tiles:
                        self._tens["tile_prob"][:, tiles.index(t)] = float(v)
            heat_u16 = self._tens["heatmap"].element_size() == 2
            if (self._prob._height, self._prob._width) != old_shape or heat_u16 != (self._max_changes > 255):
                # map size (or the heat-map width) changed: state tensors are re-allocated; RNG streams and the per-env
                # probabilities are kept
                self._pending_states = self._rng_states_numpy()
                self._pending_probs = self._tens["tile_prob"].clone()
                self._tens = None
                self._cbufs = None

    def get_border_tile(self):
        return self._prob.get_tile_types().index(self._prob._border_tile)

    def get_num_tiles(self):
        return len(self._prob.get_tile_types())

    def reset(self, mask=None):
        """PcgrlEnv.reset for every env (or those with mask[i] != 0).  pcgrl_env.py:66-76."""
        import torch
        self._ensure_buffers()
        m = None
        if mask is not None:
            m = torch.as_tensor(mask, device=self._dev).to(torch.uint8).contiguous()
        self.native_config
        if self._host:
            _native.check(_native.lib().pcgrl_reset_cpu(self._cfg_ref, C.byref(self._cbufs),
                                                        None if m is None else m.data_ptr(), self.num_envs), "pcgrl_reset_cpu")
        else:
            with torch.cuda.device(self._dev):
                _native.check(_native.lib().pcgrl_reset(self._cfg_ref, C.byref(self._cbufs),
                                                        None if m is None else m.data_ptr(), self.num_envs,
                                                        _native.stream_ptr(self._dev)), "pcgrl_reset")
        self._prob.reset(self._prob.stats_from_rows(self._tens["start_stats"]))
        self._info_cache["max_iterations"], self._info_cache["max_changes"] = self._max_iterations, self._max_changes
        return self._obs_cache

    def step(self, actions):
        """PcgrlEnv.step for the whole batch.  actions: int32 tensor/array [N] (narrow, turtle), [N,3] = (x, y, tile)
        (wide), [N,2] / [N,9] (cast / multi).  Returns (obs, reward f64[N], done bool[N], info) -- all live views
        of the env's buffers (the dicts are built once, not per call).  pcgrl_env.py:129-150."""
        import torch
        if self._tens is None:
            raise RuntimeError("call reset() before step()")
        if self._cfg is None:
            self.native_config  # re-freeze the parameters if adjust_param was called since the last step
        dev = self._dev
        a = actions
        if not (isinstance(a, torch.Tensor) and a.dtype == torch.int32 and a.device == dev and a.is_contiguous()):
            a = torch.as_tensor(actions).to(device=dev, dtype=torch.int32).contiguous()
        if a.numel() != self.num_envs * self._adim:
            raise ValueError("actions must have shape [%d%s]" % (self.num_envs, ",%d" % self._adim if self._adim > 1 else ""))
        if self._host:
            _native.check(_native.lib().pcgrl_step_cpu(self._cfg_ref, self._bufs_ref, a.data_ptr(), self.num_envs), "pcgrl_step_cpu")
            return self._obs_cache, self._reward_view, self._done_view, self._info_cache
        if torch.cuda.current_device() != dev.index:
            with torch.cuda.device(dev):
                return self.step(a)
        rc = self._fn_step(self._cfg_ref, self._bufs_ref, a.data_ptr(), self.num_envs,
                           torch.cuda.current_stream(dev).cuda_stream)
        if rc:
            _native.check(rc, "pcgrl_step")
        return self._obs_cache, self._reward_view, self._done_view, self._info_cache

    def rollout(self, actions, reward_out=None, done_out=None):
        """T steps in one native call: actions int32 CUDA [T,N] / [T,N,3].  Returns (reward [T,N],
        done [T,N]); the final observation is available through ``observation()``."""
        import torch
        if self._tens is None:
            raise RuntimeError("call reset() before rollout()")
        self.native_config
        a = torch.as_tensor(actions).to(device=self._dev, dtype=torch.int32).contiguous()
        T = a.shape[0]
        if T < 1 or a.numel() != T * self.num_envs * self._adim:
            raise ValueError("actions must have shape [T,%d%s]" % (self.num_envs, ",%d" % self._adim if self._adim > 1 else ""))
        if reward_out is None:
            reward_out = torch.empty((T, self.num_envs), dtype=torch.float64, device=self._dev)
        if done_out is None:
            done_out = torch.empty((T, self.num_envs), dtype=torch.uint8, device=self._dev)
        if reward_out.numel() < T * self.num_envs or done_out.numel() < T * self.num_envs:
            raise ValueError("reward_out / done_out must hold [T,%d] entries" % self.num_envs)
        if self._host:
            for t in range(T):
                _, r, d, _ = self.step(a[t])
                reward_out[t].copy_(r)
                done_out[t].copy_(d.view(torch.uint8))
            return reward_out, done_out.view(torch.bool)
        with torch.cuda.device(self._dev):
            _native.check(_native.lib().pcgrl_rollout(C.byref(self._cfg), C.byref(self._cbufs), a.data_ptr(),
                                                      reward_out.data_ptr(), done_out.data_ptr(), T, self.num_envs,
                                                      _native.stream_ptr(self._dev)), "pcgrl_rollout")
        return reward_out, done_out.view(torch.bool)

    def rollout_host(self, io):
        """T steps on HOST buffers in one native call (pcgrl_rollout_host, see HostRolloutIO): pinned host
        actions [T,N(,k)] in; every step's reward / done and the final observation back in pinned host memory."""
        import torch
        self.native_config
        with torch.cuda.device(self._dev):
            _native.check(_native.lib().pcgrl_rollout_host(
                C.addressof(self._cfg), C.addressof(self._cbufs), io.d_actions.data_ptr(), io.d_reward.data_ptr(),
                io.d_done.data_ptr(), C.addressof(io.struct), io.T, self.num_envs, _native.stream_ptr(self._dev)),
                "pcgrl_rollout_host")
        return io.reward, io.done

    def step_host(self, io):
        """End-to-end step on HOST buffers through pcgrl_step_host (see HostStepIO).  Hot loop: everything
        that does not change between calls is cached, the device guard is only taken when needed."""
        import torch
        dev = self._dev
        if torch.cuda.current_device() != dev.index:
            with torch.cuda.device(dev):
                return self.step_host(io)
        if self._cfg is None:
            self.native_config
        rc = self._fn_step_host(self._cfg_ref, self._bufs_ref, self._d_actions_ptr, io.ref, self.num_envs,
                                torch.cuda.current_stream(dev).cuda_stream)
        if rc:
            _native.check(rc, "pcgrl_step_host")

    def observation(self):
        return self._observation()

    def render(self, mode='rgb_array'):
        """pcgrl_env.py:160-173 for the whole batch, on the GPU: RGB uint8 tensor [N, Hpx, Wpx, 3] (level + border +
        the cursor frame of the cursor representations).  There is no window: mode='human' is not offered."""
        if mode != 'rgb_array':
            raise NotImplementedError("only mode='rgb_array' (a CUDA image batch) is offered")
        if self._tens is None:
            raise RuntimeError("call reset() before render()")
        return self._prob.render(self._tens["map"], None if self._rep.name == "wide" else self._tens["pos"])

    def close(self):
        self._tens = None
        self._cbufs = None

    # ------------------------------------------------------------------ state / checkpoint
    def state_dict(self):
        self._ensure_buffers()
        return {k: v.clone() for k, v in self._tens.items() if k not in ("scratch",)}

    def load_state_dict(self, state):
        self._ensure_buffers()
        for k, v in state.items():
            if k in self._tens and k != "scratch":
                self._tens[k].copy_(v)

    def check_status(self):
        """Raise if a device-side capacity limit was hit (synchronises)."""
        st = self._tens["status"].tolist()
        if st[0] != 0:
            raise _native.NativeError("device capacity limit hit: status=%s" % (st,))

    @property
    def native_config(self):
        if self._cfg is None:
            self._cfg = build_config(self._prob, self._rep, self._max_changes, self._max_iterations, self.auto_reset)
            self._cfg_ref = C.byref(self._cfg)
        return self._cfg

    # ------------------------------------------------------------------ internals
    def _update_spaces(self):
        w, h, t = self._prob._width, self._prob._height, self.get_num_tiles()
        self.action_space = self._rep.get_action_space(w, h, t)
        self.observation_space = self._rep.get_observation_space(w, h, t)
        self.observation_space.spaces['heatmap'] = spaces.Box(low=0, high=self._max_changes, dtype=np.uint8 if self._max_changes <= 255 else np.uint16, shape=(h, w))
        self._adim = _abi.action_dim(self._rep.name)

    def _rng_states_numpy(self):
        return self._tens["rng"].cpu().numpy().view(np.uint32).copy()

    def _ensure_buffers(self):
        import torch
        cfg = self.native_config
        if self._tens is not None:
            return
        self._host = torch.device(self.device).type == "cpu"
        _native.validate(cfg)
        if self._host:
            # host twin (pcgrl_*_cpu): explicit device="cpu" -- never a fallback
            self._dev = torch.device("cpu")
            self._tens, self._cbufs = _native.alloc_buffers(cfg, self.num_envs, self._dev)
            self._d_actions = torch.zeros(self.num_envs * self._adim, dtype=torch.int32)
        else:
            self._dev = _native.require_cuda(self.device)
            with torch.cuda.device(self._dev):
                self._tens, self._cbufs = _native.alloc_buffers(cfg, self.num_envs, self._dev)
                self._d_actions = torch.zeros(self.num_envs * self._adim, dtype=torch.int32, device=self._dev)
        self._bufs = self._tens
        self._fn_step_host = _native.lib().pcgrl_step_host
        self._fn_step = _native.lib().pcgrl_step
        self._obs_cache = self._build_observation()
        self._info_cache = self._build_info()
        self._reward_view = self._tens["reward"]
        self._done_view = self._tens["done"].view(torch.bool)
        self._bufs_ref = C.byref(self._cbufs)
        self._d_actions_ptr = self._d_actions.data_ptr()
        self._rep.bind(self)
        if self._pending_states is not None:
            self._tens["rng"].copy_(torch.from_numpy(self._pending_states.view(np.int32)))
            self._pending_states = None
        if self._pending_probs is not None:
            self._tens["tile_prob"].copy_(self._pending_probs)
            self._pending_probs = None

    def _observation(self):
        return self._obs_cache

    def _build_observation(self):
        t = self._tens
        obs = {"map": t["map"], "heatmap": t["heatmap"]}
        if self._rep.name != "wide":
            obs["pos"] = t["pos"]
        return obs

    def _build_info(self):
        """Problem.get_debug_info + the env counters (pcgrl_env.py:144-148) as views of ``info_stats`` (the
        statistics at the end of the step, written by the kernel BEFORE any auto-reset)."""
        t = self._tens
        rows = t["info_stats"]
        info = {k: rows[:, i] for i, k in enumerate(self._prob.stat_names) if k in self._prob.debug_info_names}
        for j, k in enumerate(self._prob.extra_info_names):   # derived entries the kernel appends after the stats
            info[k] = rows[:, len(self._prob.stat_names) + j]
        # the counters as they were at the end of the step (an auto-reset zeroes the live ones): VecEnv terminal-info
        info["iterations"] = rows[:, _abi.INFO_ITERATION]
        info["changes"] = rows[:, _abi.INFO_CHANGES]
        info["max_iterations"] = self._max_iterations
        info["max_changes"] = self._max_changes
        return info


class HostStepIO:
    """Pinned host buffers for ``BatchedPcgrlEnv.step_host`` (pcgrl_host_io in the header).

    mode="delta" (default): the step kernels emit one 16-byte record per env plus the fresh maps of auto-reset
    envs; one small D2H copy per step, the library patches these persistent host arrays so they always hold the
    complete current observation.  mode="full": every array is copied back in full every step.
    mode="direct": the step kernels store every result straight into these pinned (device-mapped) arrays while they
    run -- reward / done / cursor of every env, the edited map cell and its heat-map count, the whole map of an
    auto-reset env -- so nothing is copied or patched after the launch."""

    def __init__(self, env, with_obs=True, with_info=False, mode="delta"):
        import torch
        env._ensure_buffers()
        n, h, w = env.num_envs, env._prob._height, env._prob._width
        pin = dict(pin_memory=True)
        self.actions = torch.zeros((n, env._adim), dtype=torch.int32, **pin)
        self.map = torch.zeros((n, h, w), dtype=torch.uint8, **pin) if with_obs else None
        self.heatmap = torch.zeros((n, h, w), dtype=env._tens["heatmap"].dtype, **pin) if with_obs else None
        self.pos = torch.zeros((n, 2), dtype=torch.uint8, **pin) if with_obs and env._rep.name != "wide" else None
        self.reward = torch.zeros(n, dtype=torch.float64, **pin)
        self.done = torch.zeros(n, dtype=torch.uint8, **pin)
        self.info_stats = torch.zeros((n, _abi.MAX_STATS), dtype=torch.int32, **pin) if with_info else None
        s = _abi.PcgrlHostIO()
        for name in ("actions", "map", "heatmap", "pos", "reward", "done", "info_stats"):
            t = getattr(self, name)
            setattr(s, name, None if t is None else t.data_ptr())
        self.mode = mode
        self.staging_bytes = 0
        if mode == "delta":
            nbytes = int(_native.lib().pcgrl_host_staging_bytes(C.byref(env.native_config), n))
            self.d_staging = torch.zeros(nbytes, dtype=torch.uint8, device=env._dev)
            self.h_staging = torch.zeros(nbytes, dtype=torch.uint8, **pin)
            s.d_staging, s.h_staging, s.staging_bytes = self.d_staging.data_ptr(), self.h_staging.data_ptr(), nbytes
            s.mode, s.synced, s.reset_base = 1, 0, 0
            self.staging_bytes = nbytes
        elif mode == "direct":
            s.mode, s.synced = 2, 0
        elif mode != "full":
            raise ValueError("mode must be 'delta', 'full' or 'direct'")
        self.struct = s
        self.ref = C.byref(s)
        self.h2d_bytes = self.actions.numel() * 4
        self.full_bytes = sum(t.numel() * t.element_size() for t in
                              (self.map, self.heatmap, self.pos, self.reward, self.done, self.info_stats) if t is not None)
        info_bytes = 0 if self.info_stats is None else self.info_stats.numel() * 4
        self.d2h_bytes = (self.staging_bytes + info_bytes) if mode == "delta" else self.full_bytes
        if mode == "direct":   # fixed part: reward + done (+ cursor) of every env; edited cells / reset maps come on top
            self.d2h_bytes = n * (8 + 1 + (2 if self.pos is not None else 0)) + info_bytes

    def invalidate(self):
        """Call after env.reset() / env.step() / env.rollout(): the next step_host re-syncs with a full copy."""
        self.struct.synced = 0


class HostRolloutIO:
    """Pinned host buffers + device staging for ``BatchedPcgrlEnv.rollout_host`` (pcgrl_host_rollout_io)."""

    def __init__(self, env, T, with_obs=True, with_info=False):
        import torch
        env._ensure_buffers()
        n, h, w = env.num_envs, env._prob._height, env._prob._width
        pin = dict(pin_memory=True)
        self.T = int(T)
        ashape = (self.T, n, env._adim) if env._adim > 1 else (self.T, n)
        self.actions = torch.zeros(ashape, dtype=torch.int32, **pin)
        self.reward = torch.zeros((self.T, n), dtype=torch.float64, **pin)
        self.done = torch.zeros((self.T, n), dtype=torch.uint8, **pin)
        self.map = torch.zeros((n, h, w), dtype=torch.uint8, **pin) if with_obs else None
        self.heatmap = torch.zeros((n, h, w), dtype=env._tens["heatmap"].dtype, **pin) if with_obs else None
        self.pos = torch.zeros((n, 2), dtype=torch.uint8, **pin) if with_obs and env._rep.name != "wide" else None
        self.info_stats = torch.zeros((n, _abi.MAX_STATS), dtype=torch.int32, **pin) if with_info else None
        self.d_actions = torch.zeros(ashape, dtype=torch.int32, device=env._dev)
        self.d_reward = torch.zeros((self.T, n), dtype=torch.float64, device=env._dev)
        self.d_done = torch.zeros((self.T, n), dtype=torch.uint8, device=env._dev)
        s = _abi.PcgrlHostRolloutIO()
        for name in ("actions", "reward", "done", "map", "heatmap", "pos", "info_stats"):
            t = getattr(self, name)
            setattr(s, name, None if t is None else t.data_ptr())
        self.struct = s
        self.h2d_bytes = self.actions.numel() * 4
        self.d2h_bytes = sum(t.numel() * t.element_size() for t in
                             (self.reward, self.done, self.map, self.heatmap, self.pos, self.info_stats) if t is not None)


class PcgrlEnv:
    """Single-environment classic-gym facade (numpy / Python scalars) over a batch of one.

    Same constructor and methods as the reference's PcgrlEnv (pcgrl_env.py:27-150); the work still
    runs on the GPU through the C ABI -- there is no CPU implementation of the step path."""

    metadata = {'render.modes': []}

    def __init__(self, prob="binary", rep="narrow", device="cuda"):
        # device="cpu" selects the host twin (pcgrl_*_cpu)
        self._batched = BatchedPcgrlEnv(prob, rep, num_envs=1, device=device, auto_reset=False)
        self._prob = self._batched._prob
        self._rep = self._batched._rep
        self.viewer = None

    action_space = property(lambda self: self._batched.action_space)
    observation_space = property(lambda self: self._batched.observation_space)
    _max_changes = property(lambda self: self._batched._max_changes)
    _max_iterations = property(lambda self: self._batched._max_iterations)
    unwrapped = property(lambda self: self)

    def seed(self, seed=None):
        return self._batched.seed(seed)

    def set_rng(self, rep_rng, prob_rng):
        """Inject numpy RandomState objects (parity harness: env._rep._random / env._prob._random)."""
        st = np.stack([mt_state_words(rep_rng), mt_state_words(prob_rng)])[None]
        self._batched.set_rng_states(st)

    def adjust_param(self, **kwargs):
        self._batched.adjust_param(**kwargs)

    def get_border_tile(self):
        return self._batched.get_border_tile()

    def get_num_tiles(self):
        return self._batched.get_num_tiles()

    def _obs(self, obs):
        out = {"map": obs["map"][0].cpu().numpy().copy()}
        if "pos" in obs:
            out["pos"] = obs["pos"][0].cpu().numpy().copy()
        out["heatmap"] = obs["heatmap"][0].cpu().numpy().astype(np.float64)   # pcgrl_env.py:35 dtype
        return out

    def reset(self):
        return self._obs(self._batched.reset())

    def step(self, action):
        a = np.asarray(action, dtype=np.int32).reshape(1, -1)
        obs, reward, done, info = self._batched.step(a if self._batched._adim > 1 else a.reshape(1))
        r = float(reward[0].item())
        info = {k: (int(v[0].item()) if hasattr(v, "dim") else v) for k, v in info.items()}
        return self._obs(obs), (int(r) if r == int(r) and self._prob.name in ("binary", "zelda") else r), \
            bool(done[0].item()), info

    def render(self, mode='rgb_array'):
        return self._batched.render(mode)[0].cpu().numpy()

    def close(self):
        self._batched.close()

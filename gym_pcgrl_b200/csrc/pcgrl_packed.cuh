// pcgrl_packed.cuh -- several environments per warp for maps of at most 16 (8) rows.
//
// k_rollout gives one warp to one env and row r to lane r, so a 16-row map leaves half of every warp instruction empty
// (78 % at 7 rows).  Simply putting two envs side by side does not help: only ~1/3 of the steps edit the map and need the
// graph statistics, so in lock-step the two halves would rarely have that work at the same time.  The packed kernel
// therefore decouples the envs of a warp in TIME (possible in a rollout: the actions are known in advance):
//
//   phase 1  every group (G = 16 or 8 lanes = one env) runs ahead through its own cheap steps (cursor move, no edit)
//            until it reaches a step that edited the map -- its statistics job -- or needs a reset, or is finished;
//   phase 2  all pending statistics jobs run TOGETHER in one wave engine: a per-group state machine
//            (sweep from the component's first cell -> optional second sweep -> next component) advances every group by
//            two BFS waves per iteration, so a warp instruction now carries 2 (4) envs' waves;
//   phase 3  resets are rare (one per ~150 steps) and reuse the full-warp env_reset, one group at a time.
//
// Every env still sees exactly the reference's sequence of operations (pcgrl_env.py:129-150) and RNG draws, so results
// are bit-identical to k_rollout; tests/test_gpu_parity.py compares both against the oracle.
// Reference for the statistics: helper.py:197-207 (calc_num_regions), :250-264 (calc_longest_path), as in pcgrl_device.cuh.
#pragma once
#include "pcgrl_env.cuh"

namespace pcgrl {

template <int G>
struct Grp {
  int g, r, base;      // group index inside the warp, row (lane inside the group), first lane of the group
  uint32_t gmask;      // ballot bits of the group
  __device__ __forceinline__ explicit Grp(int lane)
      : g(lane / G), r(lane % G), base(lane / G * G), gmask(((G == 32) ? 0xffffffffu : ((1u << G) - 1u)) << (lane / G * G)) {}
};

// 4-neighbourhood dilation inside a group (the segment ends receive their own word, harmless: f itself is OR-ed in)
template <int G>
__device__ __forceinline__ uint32_t dilate_g(uint32_t f) {
  return f | (f << 1) | (f >> 1) | __shfl_up_sync(FULL_MASK, f, 1, G) | __shfl_down_sync(FULL_MASK, f, 1, G);
}
template <int G>
__device__ __forceinline__ int gsum(int v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
  return v;
}
// row-major-first cell of m inside the group as a one-bit board; returns whether the group's m is non-empty
template <int G>
__device__ __forceinline__ bool gfirst_seed(uint32_t m, const Grp<G>& grp, uint32_t& seed) {
  const uint32_t rows = (__ballot_sync(FULL_MASK, m != 0u) & grp.gmask) >> grp.base;
  seed = (rows != 0u && grp.r == __ffs(rows) - 1) ? (m & (0u - m)) : 0u;
  return rows != 0u;
}

// regions + calc_longest_path of `pass` for every group at once (pass == 0 for groups without a job).
// Same per-component procedure and shortcuts as regions_and_longest_path (pcgrl_device.cuh), run as a state machine:
// phase 1 = sweep from the component's first cell over `remaining`, phase 2 = sweep from the first cell of the last
// frontier over the component, phase 0 = idle.
template <int G>
__device__ __forceinline__ void packed_regions_and_longest_path(uint32_t pass, const Grp<G>& grp, int& regions_out, int& path_out) {
  uint32_t up = __shfl_up_sync(FULL_MASK, pass, 1, G), dn = __shfl_down_sync(FULL_MASK, pass, 1, G);
  if (grp.r == 0) up = 0;
  if (grp.r == G - 1) dn = 0;
  const uint32_t l = pass << 1, r = pass >> 1;
  const uint32_t iso = pass & ~(l | r | up | dn);
  const uint32_t one = pass & (l ^ r ^ up ^ dn) & ~((l & r) | (up & dn) | ((l ^ r) & (up ^ dn)));
  const uint32_t hd = one & (one >> 1);
  uint32_t below = __shfl_down_sync(FULL_MASK, one, 1, G);
  if (grp.r == G - 1) below = 0;
  const uint32_t vd = one & below;
  uint32_t vd_low = __shfl_up_sync(FULL_MASK, vd, 1, G);
  if (grp.r == 0) vd_low = 0;
  const uint32_t dominoes = hd | (hd << 1) | vd | vd_low;
  const int counts = gsum<G>(__popc(iso) | ((__popc(hd) + __popc(vd)) << 16));  // two sums in one butterfly (each < 2^16)
  const int ndom = counts >> 16;
  int regions = (counts & 0xffff) + ndom, best = ndom > 0 ? 1 : 0;
  uint32_t remaining = pass & ~iso & ~dominoes;
  uint32_t f, vis, bpass;
  int d = 0, phase;
  {
    uint32_t seed;
    const bool has = gfirst_seed<G>(remaining, grp, seed);
    phase = has ? 1 : 0;
    bpass = has ? remaining : 0u;
    f = vis = seed;
  }
  while (__any_sync(FULL_MASK, phase != 0)) {
    const uint32_t n1 = dilate_g<G>(f) & bpass & ~vis;
    const uint32_t v1 = vis | n1;
    const uint32_t n2 = dilate_g<G>(n1) & bpass & ~v1;
    const bool g2 = (__ballot_sync(FULL_MASK, n2 != 0u) & grp.gmask) != 0u;
    if (g2) { vis = v1 | n2; f = n2; d += 2; }
    const bool fin = (phase != 0) && !g2;
    if (__any_sync(FULL_MASK, fin)) {  // some group's sweep ended in this pair of waves (warp-uniform branch)
      const bool g1 = (__ballot_sync(FULL_MASK, n1 != 0u) & grp.gmask) != 0u;
      const uint32_t lastf = g1 ? n1 : f, vv = g1 ? v1 : vis;
      const int dd = d + (g1 ? 1 : 0);
      bool want2 = false;
      if (fin) {
        if (phase == 1) { remaining &= ~vv; regions++; want2 = 2 * dd > best; }
        else best = max(best, dd);
      }
      uint32_t s2, s1;
      gfirst_seed<G>(want2 ? lastf : 0u, grp, s2);
      const bool has1 = gfirst_seed<G>(remaining, grp, s1);
      if (fin) {
        d = 0;
        if (want2) { phase = 2; bpass = vv; f = vis = s2; }
        else if (has1) { phase = 1; bpass = remaining; f = vis = s1; }
        else { phase = 0; bpass = 0u; f = vis = 0u; }
      }
    }
  }
  regions_out = regions;
  path_out = best;
}

// MT19937 stream of one group: a G-word window of tempered outputs in registers, a draw is one shuffle.  All 32 lanes
// execute every call; `active` selects the groups that really consume a draw.
template <int G>
struct GroupRng {
  uint32_t* st;
  int pos, base;
  uint32_t cache;
  bool dirty;

  __device__ __forceinline__ void init(uint32_t* s, const Grp<G>& grp) {
    st = s;
    pos = (int)s[624];
    dirty = false;
    base = -1000;
    cache = 0;
    if (pos < 624) {
      base = pos;
      cache = mt_temper((pos + grp.r < 624) ? s[pos + grp.r] : 0u);
    }
  }
  __device__ __forceinline__ uint32_t next(bool active, const Grp<G>& grp, int lane) {
    unsigned tw = __ballot_sync(FULL_MASK, active && pos >= 624);
    while (tw) {  // rare (once per 624 draws): the whole warp twists one group's key
      const int k = (__ffs(tw) - 1) / G;
      tw &= ~(((G == 32) ? 0xffffffffu : ((1u << G) - 1u)) << (k * G));
      const unsigned long long p = __shfl_sync(FULL_MASK, (unsigned long long)(uintptr_t)st, k * G);
      mt_twist_warp(reinterpret_cast<uint32_t*>((uintptr_t)p), lane);
      if (grp.g == k) { pos = 0; base = -1000; }
    }
    if (active && (pos - base >= G || pos < base)) {
      base = pos;
      cache = mt_temper((pos + grp.r < 624) ? st[pos + grp.r] : 0u);
    }
    int off = pos - base;
    off = off < 0 ? 0 : (off >= G ? G - 1 : off);
    const uint32_t v = __shfl_sync(FULL_MASK, cache, grp.base + off);
    if (active) { pos++; dirty = true; }
    return v;
  }
  // RandomState.randint(n): masked rejection sampling (n is warp-uniform)
  __device__ __forceinline__ int randint(int n, bool active, const Grp<G>& grp, int lane) {
    const uint32_t rng = (uint32_t)(n - 1);
    if (rng == 0u) return 0;
    const uint32_t mask = 0xffffffffu >> __clz(rng);
    uint32_t v = 0;
    bool need = active;
    while (__any_sync(FULL_MASK, need)) {
      const uint32_t d = next(need, grp, lane) & mask;
      if (need) { v = d; need = d > rng; }
    }
    return (int)v;
  }
  __device__ __forceinline__ void finish(const Grp<G>& grp) {
    if (dirty && grp.r == 0) st[624] = (uint32_t)pos;
    dirty = false;
  }
};

// Representation.update for narrow / turtle / wide, predicated: every lane executes, `active` groups act.
template <int G>
__device__ __forceinline__ int packed_apply_action(const pcgrl_config& cfg, const int32_t* __restrict__ act, Board& board, uint8_t* map,
                                                   GroupRng<G>& rng, const Grp<G>& grp, int lane, bool active, int& x, int& y,
                                                   int& hx, int& hy) {
  const int W = cfg.width, H = cfg.height, rep = cfg.representation;
  int wx = x, wy = y, newt = -1;
  if (active) {
    if (rep == PCGRL_REP_NARROW) {
      const int a = act[0];
      if (a > 0) newt = (a - 1) & 7;
    } else if (rep == PCGRL_REP_TURTLE) {
      const int a = act[0];
      if (a >= 4) newt = (a - 4) & 7;
      else if (a >= 0) turtle_move(cfg, a, x, y);
      wx = x; wy = y;
    } else {  // PCGRL_REP_WIDE
      wx = min(max(act[0], 0), W - 1);
      wy = min(max(act[1], 0), H - 1);
      newt = act[2] & 7;
    }
  }
  const int oldt = __shfl_sync(FULL_MASK, tile_at(board, wx), grp.base + wy);
  const int change = (active && newt >= 0 && oldt != newt) ? 1 : 0;
  if (change) {
    if (grp.r == wy) set_tile(board, wx, newt);
    if (grp.r == 0) map[wy * W + wx] = (uint8_t)newt;
  }
  if (rep == PCGRL_REP_NARROW) {
    if (cfg.flags & PCGRL_FLAG_RANDOM_TILE) {  // narrow_rep.py:104-106
      const int nx = rng.randint(W, active, grp, lane), ny = rng.randint(H, active, grp, lane);
      if (active) { x = nx; y = ny; }
    } else if (active) {                       // :107-113
      x += 1;
      if (x >= W) { x = 0; y += 1; if (y >= H) y = 0; }
    }
  }
  if (active) {  // the heat-map cell of this step stays with the group until its statistics job has run
    if (rep == PCGRL_REP_WIDE) { hx = wx; hy = wy; }
    else { hx = x; hy = y; }
  }
  return change;
}

#define PACKED_WPB 4

// T-step rollout of the binary problem with 32 / G envs per warp (see the file header).
template <int G>
__global__ void __launch_bounds__(32 * PACKED_WPB) k_rollout_packed_binary(const __grid_constant__ pcgrl_config cfg,
                                                                          const __grid_constant__ pcgrl_buffers b,
                                                                          const int32_t* __restrict__ actions, double* reward_out,
                                                                          uint8_t* done_out, int T, int n) {
  constexpr int PROB = PCGRL_PROB_BINARY, EPW = 32 / G, NS = 2;
  __shared__ WarpSmem smem[PACKED_WPB];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int first_env = (blockIdx.x * PACKED_WPB + wib) * EPW;
  if (first_env >= n) return;
  WarpSmem& sm = smem[wib];
  const Grp<G> grp(lane);
  const int W = cfg.width, H = cfg.height, cells = W * H;
  const int adim = action_dim(cfg.representation);
  const bool auto_reset = (cfg.flags & PCGRL_FLAG_AUTO_RESET) != 0;
  const bool valid = first_env + grp.g < n;
  const int e = valid ? first_env + grp.g : n - 1;  // clamped: lanes of an absent env never act
  const EnvRefs refs = env_refs(cfg, b, e);
  const uint32_t rm = (grp.r < H) ? ((W >= 32) ? FULL_MASK : ((1u << W) - 1u)) : 0u;

  Board board = {0u, 0u, 0u};
#pragma unroll
  for (int k = 0; k < EPW; k++) {  // full-warp byte -> bitboard transposition, one env after the other
    if (first_env + k < n) {
      const Board bf = load_board<1>(b.map + (size_t)(first_env + k) * cells, W, H, lane, sm.bits);
      const uint32_t p0 = __shfl_sync(FULL_MASK, bf.p0, grp.r);
      if (grp.g == k) board.p0 = p0;
    }
  }
  GroupRng<G> rng;
  rng.init(refs.rng_rep, grp);
  int x = 0, y = 0;
  if (cfg.representation != PCGRL_REP_WIDE) { x = b.pos[2 * e]; y = b.pos[2 * e + 1]; }
  int iteration = b.iteration[e], changes = b.changes[e];
  int st[NS], start[NS], old[NS] = {0, 0};
  load_row<NS>(b.stats + (size_t)e * PCGRL_MAX_STATS, st);
  load_row<NS>(b.start_stats + (size_t)e * PCGRL_MAX_STATS, start);

  int t = 0, hx = 0, hy = 0;
  bool pend = false, need_reset = false;

  // end of a step (pcgrl_env.py:142-149): outputs, last-step buffers, auto-reset request
  auto finish_step = [&](double reward, bool changed) {
    const bool done = problem_over<PROB>(cfg, st, start) || changes >= cfg.max_changes || iteration >= cfg.max_iterations;
    if (grp.r == 0) {
      const size_t row = (size_t)t * n + e;
      if (reward_out) reward_out[row] = reward;
      if (done_out) done_out[row] = done ? 1 : 0;
      if (t == T - 1) {
        b.reward[e] = reward;
        b.done[e] = done ? 1 : 0;
        int32_t* info = b.info_stats + (size_t)e * PCGRL_MAX_STATS;
        info[0] = st[0]; info[1] = st[1];
        info[2] = st[1] - start[1];  // info["path-imp"] (binary_prob.py:137), before any auto-reset
        info[PCGRL_INFO_ITERATION] = iteration;
        info[PCGRL_INFO_CHANGES] = changes;
      }
    }
    if (done && auto_reset) need_reset = true;
    else if (changed) heat_increment(cfg, b.heatmap, (size_t)e * cells + (size_t)hy * W + hx, grp.r);  // :137
    t++;
  };

  while (true) {
    // ---- phase 1: run ahead through steps that do not edit the map
    while (true) {
      const bool run = valid && !pend && !need_reset && t < T;
      if (!__any_sync(FULL_MASK, run)) break;
      const int32_t* act = actions + ((size_t)(run ? t : 0) * n + e) * adim;
      if (run) {
        iteration++;  // pcgrl_env.py:130
        old[0] = st[0]; old[1] = st[1];
      }
      const int change = packed_apply_action<G>(cfg, act, board, refs.map, rng, grp, lane, run, x, y, hx, hy);
      if (run) {
        if (change > 0) { changes += change; pend = true; }  // :135-138, statistics in phase 2
        else finish_step(0.0, false);                         // get_reward(s, s) == 0
      }
    }
    // ---- phase 2: the statistics jobs of all pending groups, together
    if (__any_sync(FULL_MASK, pend)) {
      int regions, path;
      packed_regions_and_longest_path<G>(pend ? (~board.p0 & rm) : 0u, grp, regions, path);
      if (pend) {
        st[0] = regions; st[1] = path;  // binary_prob.py:81-86
        finish_step(problem_reward<PROB>(cfg, st, old), true);
        pend = false;
      }
    }
    // ---- phase 3: resets, one group at a time on the full warp
    unsigned rb = __ballot_sync(FULL_MASK, need_reset);
    while (rb) {
      const int k = (__ffs(rb) - 1) / G;
      rb &= ~(((G == 32) ? 0xffffffffu : ((1u << G) - 1u)) << (k * G));
      if (grp.g == k) rng.finish(grp);  // the reset continues the env's stream from the position reached so far
      __syncwarp();
      const int ek = first_env + k;
      WarpRng wr;
      wr.init(b.rng + (size_t)ek * 2 * PCGRL_MT_WORDS);
      Board bf;
      int rx = 0, ry = 0, rst[NS];
      bool unused;
      env_reset<PROB>(cfg, b, ek, lane, sm, wr, bf, rx, ry, rst, unused);
      wr.finish(lane);
      __syncwarp();
      const uint32_t p0 = __shfl_sync(FULL_MASK, bf.p0, grp.r);
      if (grp.g == k) {
        board.p0 = p0;
        x = rx; y = ry;
        st[0] = start[0] = rst[0];  // problem.py:45-46
        st[1] = start[1] = rst[1];
        iteration = 0;
        changes = 0;
        need_reset = false;
        rng.init(refs.rng_rep, grp);
      }
    }
    if (!__any_sync(FULL_MASK, valid && t < T)) break;
  }
  rng.finish(grp);
  if (valid && grp.r == 0) {
    if (cfg.representation != PCGRL_REP_WIDE) { b.pos[2 * e] = (uint8_t)x; b.pos[2 * e + 1] = (uint8_t)y; }
    b.iteration[e] = iteration;
    b.changes[e] = changes;
    b.stats[(size_t)e * PCGRL_MAX_STATS] = st[0]; b.stats[(size_t)e * PCGRL_MAX_STATS + 1] = st[1];
    b.start_stats[(size_t)e * PCGRL_MAX_STATS] = start[0]; b.start_stats[(size_t)e * PCGRL_MAX_STATS + 1] = start[1];
  }
}

}  // namespace pcgrl

// pcgrl_env.cuh -- PcgrlEnv.reset / Representation.update for one warp = one environment.
// Reference: gym_pcgrl/envs/pcgrl_env.py:66-76,129-150; reps/{narrow,turtle,wide}_rep.py; helper.py:310-352.
#pragma once
#include "pcgrl_problems.cuh"

namespace pcgrl {

struct EnvRefs {  // per-env base pointers
  uint8_t* map;
  uint8_t* heat;
  uint8_t* start_map;
  uint32_t* rng_rep;
  uint32_t* rng_prob;
  double* tile_prob;
};

__host__ __device__ __forceinline__ int heat_bytes(const pcgrl_config& cfg) { return (cfg.flags & PCGRL_FLAG_HEAT_U16) ? 2 : 1; }

__device__ __forceinline__ EnvRefs env_refs(const pcgrl_config& cfg, const pcgrl_buffers& b, int e) {
  const size_t cells = (size_t)cfg.width * cfg.height;
  EnvRefs r;
  r.map = b.map + (size_t)e * cells;
  r.heat = reinterpret_cast<uint8_t*>(b.heatmap) + (size_t)e * cells * heat_bytes(cfg);
  r.start_map = b.start_map + (size_t)e * cells;
  r.rng_rep = b.rng + (size_t)e * 2 * PCGRL_MT_WORDS;
  r.rng_prob = r.rng_rep + PCGRL_MT_WORDS;
  r.tile_prob = b.tile_prob + (size_t)e * PCGRL_MAX_TILES;
  return r;
}

__device__ __forceinline__ int action_dim(int representation) {
  return representation == PCGRL_REP_WIDE ? 3
       : (representation == PCGRL_REP_NARROWCAST || representation == PCGRL_REP_TURTLECAST) ? 2
       : representation == PCGRL_REP_NARROWMULTI ? 9 : 1;
}

// turtle_rep.py:101-125 / turtle_cast_rep.py:41-61: move the cursor, clamp or wrap
__device__ __forceinline__ void turtle_move(const pcgrl_config& cfg, int a, int& x, int& y) {
  const int W = cfg.width, H = cfg.height;
  const bool warp = (cfg.flags & PCGRL_FLAG_WARP) != 0;
  const int dx = (a == 0) ? -1 : (a == 1) ? 1 : 0, dy = (a == 2) ? -1 : (a == 3) ? 1 : 0;
  x += dx;
  if (x < 0) x = warp ? x + W : 0;
  if (x >= W) x = warp ? x - W : W - 1;
  y += dy;
  if (y < 0) y = warp ? y + H : 0;
  if (y >= H) y = warp ? y - H : H - 1;
}

// Write up to nine tiles of the 3x3 block centred on (x, y): t9[(dy+1)*3 + dx+1] >= 0 is the new tile, -1 keeps
// the cell; cells outside the map are skipped (narrow_cast_rep.py:43-48, narrow_multi_rep.py:41-47,
// turtle_cast_rep.py:69-75).  Lane y+dy edits its own bitboard row and the uint8 map; returns the number of
// cells whose tile changed.
__device__ __forceinline__ int apply_stamp(const pcgrl_config& cfg, Board& board, uint8_t* map, int lane, int x, int y,
                                           const int (&t9)[9]) {
  const int W = cfg.width, H = cfg.height;
  int cnt = 0;
  const int row = lane - y;  // -1, 0, 1 for the three lanes that own a stamped row
  if (row >= -1 && row <= 1 && lane < H) {
#pragma unroll
    for (int dx = -1; dx <= 1; dx++) {
      const int cx = x + dx;
      const int newt = (row < 0) ? t9[dx + 1] : (row == 0) ? t9[3 + dx + 1] : t9[6 + dx + 1];
      if (cx >= 0 && cx < W && newt >= 0) {
        if (tile_at(board, cx) != newt) {
          cnt++;
          set_tile(board, cx, newt);
          map[lane * W + cx] = (uint8_t)newt;
        }
      }
    }
  }
  return (int)__reduce_add_sync(FULL_MASK, (unsigned)cnt);
}

// Representation.update(action) -> (change, x, y) where (hx, hy) is the heat-map cell
// (narrow_rep.py:99-114: the cursor AFTER it moved; turtle_rep.py:101-129; wide_rep.py:67-70; the cast / multi
// variants stamp a 3x3 block).  Changed tiles are written to the uint8 map in HBM and to the bitboards.
// cell/tile describe a single-cell edit (delta transport); multi is set when more than one cell may have changed.
// (ex, ey) = the edited cell of a single-cell edit (cell = ey * W + ex).  REPT >= 0 fixes the representation at compile
// time (the branch chain below folds away); REPT = -1 reads it from the config.
template <int REPT = -1>
__device__ __forceinline__ int apply_action(const pcgrl_config& cfg, const int32_t* __restrict__ act, Board& board,
                                            uint8_t* map, WarpRng& rng, int lane, int& x, int& y, int& hx, int& hy,
                                            int& cell, int& tile, bool& multi, int& ex, int& ey, int& old_tile) {
  const int W = cfg.width, H = cfg.height, rep = (REPT >= 0) ? REPT : cfg.representation;
  int change = 0, wx = x, wy = y, newt = -1;
  old_tile = 0;
  multi = false;
  if (rep == PCGRL_REP_NARROW) {
    const int a = act[0];
    if (a > 0) newt = (a - 1) & 7;
  } else if (rep == PCGRL_REP_TURTLE) {
    const int a = act[0];
    if (a >= 4) newt = (a - 4) & 7;
    else if (a >= 0) turtle_move(cfg, a, x, y);
  } else if (rep == PCGRL_REP_WIDE) {
    wx = min(max(act[0], 0), W - 1);
    wy = min(max(act[1], 0), H - 1);
    newt = act[2] & 7;
  } else if (rep == PCGRL_REP_NARROWCAST) {
    const int type = act[0], value = act[1] & 7;
    if (type == 1) newt = value;
    else if (type == 2) {
      const int t9[9] = {value, value, value, value, value, value, value, value, value};
      change = apply_stamp(cfg, board, map, lane, x, y, t9);
      multi = true;
    }
  } else if (rep == PCGRL_REP_NARROWMULTI) {
    int t9[9];
#pragma unroll
    for (int k = 0; k < 9; k++) { const int a = act[k]; t9[k] = (a > 0) ? ((a - 1) & 7) : -1; }
    change = apply_stamp(cfg, board, map, lane, x, y, t9);
    multi = true;
  } else {  // PCGRL_REP_TURTLECAST
    const int type = act[0], value = act[1] & 7;
    if (type >= 0 && type < 4) turtle_move(cfg, type, x, y);
    else if (type == 4) newt = value;
    else if (type == 5) {
      const int t9[9] = {value, value, value, value, value, value, value, value, value};
      change = apply_stamp(cfg, board, map, lane, x, y, t9);
      multi = true;
    }
    wx = x; wy = y;
  }
  if (newt >= 0) {
    const int oldt = __shfl_sync(FULL_MASK, tile_at(board, wx), wy);
    old_tile = oldt;
    change = (oldt != newt) ? 1 : 0;
    if (change) {
      if (lane == wy) set_tile(board, wx, newt);
      if (lane == 0) map[wy * W + wx] = (uint8_t)newt;
    }
  }
  cell = wy * W + wx;
  ex = wx;
  ey = wy;
  tile = newt < 0 ? 0 : newt;
  if (rep == PCGRL_REP_NARROW || rep == PCGRL_REP_NARROWCAST || rep == PCGRL_REP_NARROWMULTI) {
    if (cfg.flags & PCGRL_FLAG_RANDOM_TILE) {
      rng.randint2(W, H, lane, x, y);
    } else {
      x += 1;
      if (x >= W) { x = 0; y += 1; if (y >= H) y = 0; }
    }
    hx = x; hy = y;
  } else if (rep == PCGRL_REP_WIDE) {
    hx = wx; hy = wy;
  } else {
    hx = x; hy = y;
  }
  return change;
}
__device__ __forceinline__ int apply_action(const pcgrl_config& cfg, const int32_t* __restrict__ act, Board& board,
                                            uint8_t* map, WarpRng& rng, int lane, int& x, int& y, int& hx, int& hy,
                                            int& cell, int& tile, bool& multi) {
  int ex, ey, old_tile;
  return apply_action<-1>(cfg, act, board, map, rng, lane, x, y, hx, hy, cell, tile, multi, ex, ey, old_tile);
}

// _heatmap[y][x] += 1 (pcgrl_env.py:137) as a fire-and-forget 32-bit reduction on the containing word
// (uint8 elements, or uint16 with PCGRL_FLAG_HEAT_U16; elem_off counts elements from the start of the batch).
__device__ __forceinline__ void heat_increment(const pcgrl_config& cfg, void* heat_base, size_t elem_off, int lane) {
  if (lane == 0) {
    if (cfg.flags & PCGRL_FLAG_HEAT_U16) {
      uint32_t* w = reinterpret_cast<uint32_t*>(heat_base) + (elem_off >> 1);
      atomicAdd(w, 1u << (16u * (uint32_t)(elem_off & 1)));
    } else {
      uint32_t* w = reinterpret_cast<uint32_t*>(heat_base) + (elem_off >> 2);
      atomicAdd(w, 1u << (8u * (uint32_t)(elem_off & 3)));
    }
  }
}

__device__ __forceinline__ void warp_fill_bytes(uint8_t* dst, int nbytes, uint8_t v, int lane) {
  if ((((uintptr_t)dst) & 3) == 0 && (nbytes & 3) == 0) {
    const uint32_t vv = 0x01010101u * v;
    for (int i = lane; i < (nbytes >> 2); i += 32) reinterpret_cast<uint32_t*>(dst)[i] = vv;
  } else {
    for (int i = lane; i < nbytes; i += 32) dst[i] = v;
  }
}

// A reset reads the 624 key words of the representation stream (20 lines of 128 B), the problem stream's position and
// the tile probabilities: three dependent HBM round trips when issued at the reset itself.  A step that may end the
// episode through the change / iteration limits issues these prefetches before its own statistics are computed.
__device__ __forceinline__ int max_change_per_step(int representation) { return representation >= PCGRL_REP_NARROWCAST ? 9 : 1; }
__device__ __forceinline__ void prefetch_reset_inputs(const EnvRefs& r, int lane) {
  if (lane <= 20) asm volatile("prefetch.global.L1 [%0];" ::"l"(r.rng_rep + min(lane * 32, 623)));  // 5000 B per env: 20-21 lines
  else if (lane == 21) asm volatile("prefetch.global.L1 [%0];" ::"l"(r.rng_prob + 624));
  else if (lane == 22) asm volatile("prefetch.global.L1 [%0];" ::"l"(r.tile_prob));
}

// PcgrlEnv.reset for one env, up to (and including) the map part of get_stats.  The caller finishes
// Problem.reset (start_stats) -- after the solver for the solver problems.
template <int PROB>
__device__ __forceinline__ void env_reset(const pcgrl_config& cfg, const pcgrl_buffers& b, int e, int lane, WarpSmem& sm,
                                       WarpRng& rng, Board& board, int& x, int& y, int* st, bool& need_solver) {
  constexpr int NP = ProblemTraits<PROB>::NPLANES;
  const int W = cfg.width, H = cfg.height, cells = W * H, T = cfg.num_tiles;
  const EnvRefs r = env_refs(cfg, b, e);
  const bool generate = (cfg.flags & PCGRL_FLAG_RANDOM_START) || (b.start_valid[e] == 0);
  __syncwarp();
#ifdef PCGRL_PROFILE
  long long tp[12]; int ntp = 0;
#define TP() do { __syncwarp(); tp[ntp++] = clock64(); } while (0)
#else
#define TP() do {} while (0)
#endif
  TP();
  uint32_t* rng_home = nullptr;
  WarpRng pr;  // problem stream (binary_prob.py:68-72): consumed at the end
  const bool redraw_probs = (PROB == PCGRL_PROB_BINARY) && (cfg.flags & PCGRL_FLAG_RANDOM_PROBS);
  // its position word is loaded together with the representation stream's key (one round trip for both); the load
  // that depends on it is issued after the key has been staged
  const uint32_t pr_pos = redraw_probs ? r.rng_prob[624] : 0u;
  if (generate) {  // representation.py:41-43 -> helper.py:310-312 gen_random_map
    rng_home = rng.stage(sm.mt, lane);  // the reset consumes 2*H*W (+2) draws and usually crosses a twist
    TP();
    if (redraw_probs) pr.init(r.rng_prob, lane, (int)pr_pos);
    TP();
    // helper.py:343-352 get_int_prob, then RandomState.choice: cdf = cumsum(p); cdf /= cdf[-1]
    // One probability per lane (a single load round trip), sequential sums in the reference's order via shuffles,
    // the two rounds of divisions done by all lanes at once.
    // searchsorted(cdf, u, side='right') == #{t : cdf[t] <= u}.  u = k * 2^-53 with the 53-bit integer
    // k = (a << 26) | b, and cdf[t] * 2^53 is exact, so cdf[t] <= u  <=>  ceil(cdf[t] * 2^53) <= k: the per-cell
    // comparisons are done on integers, bit-identical to numpy's double comparison.
    const double p_lane = (lane < T) ? r.tile_prob[lane] : 0.0;
    double total = 0.0;
#pragma unroll 1
    for (int t = 0; t < T; t++) total += __shfl_sync(FULL_MASK, p_lane, t);      // get_int_prob: total += prob[t]
    const double q_lane = p_lane / total;                                          // result[i] /= total
    double acc = 0.0, cdf_lane = 0.0;
#pragma unroll 1
    for (int t = 0; t < T; t++) {                                                  // cdf = p.cumsum()
      acc += __shfl_sync(FULL_MASK, q_lane, t);
      if (lane == t) cdf_lane = acc;
    }
    const double scaled = ceil((cdf_lane / acc) * 9007199254740992.0);             // cdf /= cdf[-1]; * 2^53
    const unsigned long long thr_lane =
        (lane >= T || scaled >= 18446744073709551615.0) ? 0xffffffffffffffffull : (unsigned long long)scaled;
    unsigned long long thr[PCGRL_MAX_TILES];
#pragma unroll
    for (int t = 0; t < PCGRL_MAX_TILES; t++) thr[t] = __shfl_sync(FULL_MASK, thr_lane, t);
    const int nchunks = (cells + 31) >> 5;
    TP();
    for (int s0 = 0; s0 < cells; s0 += 256) {  // segments of 256 cells: 512 draws staged at once, 8 cells per lane
      const int nseg = min(256, cells - s0);
      rng.fill(sm.draws, 2 * nseg, lane);  // H*W random_sample() doubles in row-major order
      if (s0 == 0) TP();
#pragma unroll 2
      for (int k = 0; k < 8; k++) {
        const int j = k * 32 + lane;  // cell inside the segment
        uint32_t tile = 0;
        if (j < nseg) {
          const uint32_t a = sm.draws[2 * j] >> 5, bb = sm.draws[2 * j + 1] >> 6;  // random_sample(): (a*2^26 + b) / 2^53
          const unsigned long long k53 = ((unsigned long long)a << 26) | (unsigned long long)bb;
#pragma unroll
          for (int t = 0; t < PCGRL_MAX_TILES; t++) tile += (thr[t] <= k53) ? 1u : 0u;  // searchsorted(side='right')
          r.map[s0 + j] = (uint8_t)tile;
          r.start_map[s0 + j] = (uint8_t)tile;  // _old_map = _map.copy()
        }
        if (k * 32 < nseg) chunk_to_bits<NP>(tile, (s0 >> 5) + k, lane, sm.bits);
      }
      __syncwarp();
    }
    board = bits_to_board<NP>(sm.bits, nchunks, W, H, lane);
    if (lane == 0) b.start_valid[e] = 1;
    TP();
  } else {  // representation.py:44-45
    if (redraw_probs) pr.init(r.rng_prob, lane, (int)pr_pos);
    for (int i = lane; i < cells; i += 32) r.map[i] = r.start_map[i];
    board = load_board<NP>(r.start_map, W, H, lane, sm.bits);
  }
  if (cfg.representation != PCGRL_REP_WIDE) {  // narrow_rep.py:30-31, turtle_rep.py:32-33
    x = rng.randint(W, lane);
    y = rng.randint(H, lane);
  }
  if (rng_home) rng.unstage(rng_home, lane);
  TP();
  {  // the out-of-line call gets its own array so that the caller's statistics can stay in registers
    int tmp[ProblemTraits<PROB>::NSTATS];
    map_stats_shared<PROB>(board, cfg, lane, tmp, need_solver);
#pragma unroll
    for (int i = 0; i < ProblemTraits<PROB>::NSTATS; i++) st[i] = tmp[i];
  }
  TP();
  if (redraw_probs) {  // binary_prob.py:68-72 (problem stream)
    const double p_empty = pr.next_double(lane);
    pr.finish(lane);
    if (lane == 0) { r.tile_prob[0] = p_empty; r.tile_prob[1] = 1 - p_empty; }
  }
  warp_fill_bytes(r.heat, cells * heat_bytes(cfg), 0, lane);  // pcgrl_env.py:72
  __syncwarp();
  TP();
#ifdef PCGRL_PROFILE
  if (lane == 0 && b.status) {  // accumulate phase cycles: status[8 + k]
    long long* acc = reinterpret_cast<long long*>(b.status) + 4;
    for (int k = 1; k < ntp; k++) atomicAdd(reinterpret_cast<unsigned long long*>(acc + k), (unsigned long long)(tp[k] - tp[k - 1]));
    atomicAdd(reinterpret_cast<unsigned long long*>(acc), 1ull);
  }
#endif
}

// ------------------------------------------------------------------------------------------------
// delta transport for pcgrl_step_host (mode 1).  Staging buffer, copied to the host in ONE D2H transfer:
//   [header 16 B: u32 running reset counter, u32 running change counter]
//   [reward f64 x n][done u8 x n][pos u8 x 2n]   final array layout: the host bulk-copies them
//   [ChangeRecord x n]   compacted: one record per env whose observation changed this step
//   [nslots x H*W bytes] fresh maps of auto-reset envs / multi-cell edits
// ------------------------------------------------------------------------------------------------
struct __align__(8) ChangeRecord {
  uint32_t env_kind;  // env index (24 bits) | kind << 24
  uint16_t cell;      // y*W + x of the changed cell (single-cell edit)
  uint8_t tile, slot; // new tile; staging slot of the whole map (0xFF: none / overflow)
};
#define PCGRL_REC_CHANGED 1 /* one map cell changed, heat map += 1 at the heat cell */
#define PCGRL_REC_RESET 2   /* env was reset: whole map replaced (slot), heat map cleared */
#define PCGRL_REC_MULTI 4   /* several cells changed: whole map replaced (slot), heat map += 1 */
#define PCGRL_STAGING_HEADER 16

struct StagingLayout { size_t reward_off, done_off, pos_off, rec_off, slot_off, total; };
__host__ __device__ __forceinline__ StagingLayout staging_layout(int n, int nslots, int cells) {
  StagingLayout L;
  L.reward_off = PCGRL_STAGING_HEADER;
  L.done_off = L.reward_off + sizeof(double) * (size_t)n;
  L.pos_off = L.done_off + (size_t)n;
  L.rec_off = (L.pos_off + 2 * (size_t)n + 15) & ~(size_t)15;
  L.slot_off = L.rec_off + sizeof(ChangeRecord) * (size_t)n;
  L.total = L.slot_off + (size_t)nslots * cells;
  return L;
}

struct Staging {
  uint8_t* base;         // nullptr: delta transport disabled
  uint32_t reset_base;   // values of the running counters at the start of this step
  uint32_t change_base;
  int nslots;
  int n;
  // direct transport (pcgrl_host_io mode 2): device-visible addresses of the caller's pinned host arrays; the kernel
  // stores every result where it belongs while it runs, nothing is copied or patched afterwards
  int direct;            // 1: per-step results (pcgrl_step_host); 2: final state of a T-step rollout (pcgrl_rollout_host)
  uint8_t* h_map;
  uint8_t* h_heat;       // uint8 or uint16 elements, like the device heat map
  uint8_t* h_pos;        // nullptr for the wide representation
  double* h_reward;
  uint8_t* h_done;
  const void* d_heat;    // device heat map (the new value of the incremented cell is read back from it)
};

// Direct transport of one env's step results into the host arrays (posted writes over PCIe, overlapped with the
// rest of the launch).
__device__ __forceinline__ void write_direct(const Staging& sg, const pcgrl_config& cfg, int e, int lane, double reward,
                                             bool done, int x, int y, bool changed, bool was_reset, int cell, int tile,
                                             const uint8_t* new_map, bool multi) {
  const int cells = cfg.width * cfg.height, hb = heat_bytes(cfg);
  const bool wide = cfg.representation == PCGRL_REP_WIDE;
  if (sg.direct == 2) {
    // Final state of a T-step rollout (pcgrl_rollout_host): the env's whole map and heat map and its cursor.  The
    // bytes were written by this warp over the T steps (plain stores and reductions, mostly by lane 0); the warp barrier
    // orders them before the L2 loads (ld.cg) of the other lanes.
    if (lane == 0 && sg.h_pos) *reinterpret_cast<uint16_t*>(sg.h_pos + 2 * (size_t)e) = (uint16_t)((uint32_t)x | ((uint32_t)y << 8));
    __syncwarp();
    if (sg.h_map) {
      uint8_t* dst = sg.h_map + (size_t)e * cells;
      if ((cells & 3) == 0) {
        for (int i = lane; i < (cells >> 2); i += 32) reinterpret_cast<uint32_t*>(dst)[i] = __ldcg(reinterpret_cast<const uint32_t*>(new_map) + i);
      } else {
        for (int i = lane; i < cells; i += 32) dst[i] = __ldcg(new_map + i);
      }
    }
    if (sg.h_heat) {
      const int nb = cells * hb;
      const uint8_t* src = reinterpret_cast<const uint8_t*>(sg.d_heat) + (size_t)e * nb;
      uint8_t* dst = sg.h_heat + (size_t)e * nb;
      if ((nb & 3) == 0) {
        for (int i = lane; i < (nb >> 2); i += 32) reinterpret_cast<uint32_t*>(dst)[i] = __ldcg(reinterpret_cast<const uint32_t*>(src) + i);
      } else {
        for (int i = lane; i < nb; i += 32) dst[i] = __ldcg(src + i);
      }
    }
    return;
  }
  if (lane == 0) {
    sg.h_reward[e] = reward;
    sg.h_done[e] = done ? 1 : 0;
    if (sg.h_pos) *reinterpret_cast<uint16_t*>(sg.h_pos + 2 * (size_t)e) = (uint16_t)((uint32_t)x | ((uint32_t)y << 8));
  }
  multi = multi && changed && !was_reset;
  if (!(changed || was_reset)) return;
  if ((was_reset || multi) && sg.h_map) {  // whole map (the warp's own earlier stores are visible after the barrier)
    __syncwarp();
    uint8_t* dst = sg.h_map + (size_t)e * cells;
    if ((cells & 3) == 0) {  // L2 loads: the bytes were stored by other lanes of this warp
      for (int i = lane; i < (cells >> 2); i += 32)
        reinterpret_cast<uint32_t*>(dst)[i] = __ldcg(reinterpret_cast<const uint32_t*>(new_map) + i);
    } else {
      for (int i = lane; i < cells; i += 32) dst[i] = __ldcg(new_map + i);
    }
  } else if (lane == 0 && sg.h_map) {
    sg.h_map[(size_t)e * cells + cell] = (uint8_t)tile;
  }
  if (!sg.h_heat) return;
  if (was_reset) {
    warp_fill_bytes(sg.h_heat + (size_t)e * cells * hb, cells * hb, 0, lane);   // pcgrl_env.py:72
  } else if (lane == 0) {
    // the heat cell was incremented by heat_increment (a reduction on the containing word); an atomic read of that word
    // by the same thread is ordered after it
    const size_t hi = (size_t)e * cells + (wide ? (size_t)cell : (size_t)y * cfg.width + x);
    if (hb == 2) {
      const uint32_t w = atomicAdd(const_cast<uint32_t*>(reinterpret_cast<const uint32_t*>(sg.d_heat)) + (hi >> 1), 0u);
      reinterpret_cast<uint16_t*>(sg.h_heat)[hi] = (uint16_t)(w >> (16u * (uint32_t)(hi & 1)));
    } else {
      const uint32_t w = atomicAdd(const_cast<uint32_t*>(reinterpret_cast<const uint32_t*>(sg.d_heat)) + (hi >> 2), 0u);
      sg.h_heat[hi] = (uint8_t)(w >> (8u * (uint32_t)(hi & 3)));
    }
  }
}

__device__ __forceinline__ void write_record(const Staging& sg, const pcgrl_config& cfg, int e, int lane, double reward,
                                             bool done, int x, int y, bool changed, bool was_reset, int cell, int tile,
                                             const uint8_t* new_map, bool multi = false) {
  if (sg.direct) { write_direct(sg, cfg, e, lane, reward, done, x, y, changed, was_reset, cell, tile, new_map, multi); return; }
  if (!sg.base) return;
  const int cells = cfg.width * cfg.height;
  const StagingLayout L = staging_layout(sg.n, sg.nslots, cells);
  if (lane == 0) {
    reinterpret_cast<double*>(sg.base + L.reward_off)[e] = reward;
    sg.base[L.done_off + e] = done ? 1 : 0;
    sg.base[L.pos_off + 2 * e] = (uint8_t)x;
    sg.base[L.pos_off + 2 * e + 1] = (uint8_t)y;
  }
  multi = multi && changed && !was_reset;
  if (!(changed || was_reset)) return;
  int slot = 0xFF;
  if (was_reset || multi) {
    uint32_t s = 0;
    if (lane == 0) s = atomicAdd(reinterpret_cast<uint32_t*>(sg.base), 1u) - sg.reset_base;
    s = __shfl_sync(FULL_MASK, s, 0);
    if (s < (uint32_t)sg.nslots) {
      slot = (int)s;
      uint8_t* dst = sg.base + L.slot_off + (size_t)slot * cells;
      __syncwarp();
      for (int i = lane; i < cells; i += 32) dst[i] = new_map[i];
    }
  }
  if (lane == 0) {
    const uint32_t k = atomicAdd(reinterpret_cast<uint32_t*>(sg.base) + 1, 1u) - sg.change_base;
    ChangeRecord r;
    const uint32_t kind = was_reset ? PCGRL_REC_RESET : (multi ? PCGRL_REC_MULTI : PCGRL_REC_CHANGED);
    r.env_kind = (uint32_t)e | (kind << 24);
    r.cell = (uint16_t)cell;
    r.tile = (uint8_t)tile;
    r.slot = (uint8_t)slot;
    if (k < (uint32_t)sg.n) reinterpret_cast<ChangeRecord*>(sg.base + L.rec_off)[k] = r;
  }
}

}  // namespace pcgrl

// pcgrl_b200.cu -- kernels + C ABI (include/pcgrl_b200.h) of the batched PcgrlEnv hot path, sm_100a.
//
// Kernels (one warp per environment, 4 warps per CTA):
//   k_rollout<PROB, REP> fused PcgrlEnv.step x T for the graph-only problems (binary, zelda):
//                       Representation.update -> get_stats -> get_reward/get_episode_over -> auto reset
//   k_reset<PROB>       PcgrlEnv.reset
//   k_get_stats<PROB>   stand-alone Problem.get_stats operator
//   k_seed              numpy RandomState(seed) == MT19937 init_genrand
// Solver problems (sokoban, ddave, mdungeon) run update / solver / finish as separate launches, see
// pcgrl_solver.cuh.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include <thread>
#include <vector>

#include "pcgrl_env.cuh"
#include "pcgrl_solver.cuh"
#include "pcgrl_wrappers.cuh"
#include "pcgrl_smb_env.cuh"

using namespace pcgrl;

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------
#define WPB PCGRL_WARPS_PER_BLOCK
#define PCGRL_PACKED_MIN_T 8 /* shorter fragments do not amortise the packed kernel's per-env prologue */

template <int N>
__device__ __forceinline__ void load_row(const int32_t* __restrict__ src, int* dst) {
#pragma unroll
  for (int i = 0; i < N; i++) dst[i] = src[i];
}
template <int N>
__device__ __forceinline__ void store_row(int32_t* dst, const int* src, int lane) {
#pragma unroll
  for (int i = 0; i < N; i++) if (lane == i) dst[i] = src[i];
}

// info["iterations"] / info["changes"] (pcgrl_env.py:144-145): the counters at the end of the step, before any auto-reset
__device__ __forceinline__ void store_info_counters(int32_t* info_row, int iteration, int changes, int lane) {
  if (lane == PCGRL_INFO_ITERATION) info_row[PCGRL_INFO_ITERATION] = iteration;
  if (lane == PCGRL_INFO_CHANGES) info_row[PCGRL_INFO_CHANGES] = changes;
}

#include "pcgrl_packed.cuh"

// REPT >= 0: the representation is a compile-time constant (narrow / turtle / wide); REPT = -1: read from the config.
template <int PROB, int REPT>
__global__ void __launch_bounds__(32 * WPB, 7) k_rollout(const __grid_constant__ pcgrl_config cfg,
                                                      const __grid_constant__ pcgrl_buffers b,
                                                      const int32_t* __restrict__ actions, double* reward_out,
                                                      uint8_t* done_out, int T, int n, Staging sg) {
  constexpr int NP = ProblemTraits<PROB>::NPLANES, NS = ProblemTraits<PROB>::NSTATS;
  __shared__ WarpSmem smem[WPB];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int e = blockIdx.x * WPB + wib;
  if (e >= n) return;
  WarpSmem& sm = smem[wib];
  const int W = cfg.width, H = cfg.height, cells = W * H;
  const int rep = (REPT >= 0) ? REPT : cfg.representation;
  const int adim = action_dim(rep);
  const bool auto_reset = (cfg.flags & PCGRL_FLAG_AUTO_RESET) != 0;
  // the only per-env pointer that stays live across the step loop (the others are rebuilt where a reset needs them)
  uint8_t* const env_map = b.map + (size_t)e * cells;

#ifdef PCGRL_PROFILE
  long long kp[6]; int nkp = 0;
#define KP() do { __syncwarp(); kp[nkp++] = clock64(); } while (0)
#else
#define KP() do {} while (0)
#endif
  KP();
  // One-word actions (narrow / turtle): the first step's action is LOADED first of all -- with pcgrl_step_host it sits in
  // pinned host memory, and its PCIe round trip then overlaps the state loads below -- and every later step's action is
  // loaded one step ahead (software pipelining; the other representations keep the L1 prefetch).
  constexpr bool ONE_WORD = (REPT == PCGRL_REP_NARROW || REPT == PCGRL_REP_TURTLE);
  int a_next = 0;
  if (ONE_WORD) a_next = actions[e];
  // all prologue loads are independent: issue them before the ballots of load_board serialise the warp
  WarpRng rng;
  rng.init(b.rng + (size_t)e * 2 * PCGRL_MT_WORDS, (rep == PCGRL_REP_NARROW || rep >= PCGRL_REP_NARROWCAST) ? lane : -1);
  int x = 0, y = 0;
  if (rep != PCGRL_REP_WIDE) { x = b.pos[2 * e]; y = b.pos[2 * e + 1]; }
  int iteration = b.iteration[e], changes = b.changes[e];
  int st[NS], start[NS];
  load_row<NS>(b.stats + (size_t)e * PCGRL_MAX_STATS, st);
  load_row<NS>(b.start_stats + (size_t)e * PCGRL_MAX_STATS, start);
  // first step's action: pull its line into L1 now so the load in apply_action does not add a round trip
  if (!ONE_WORD) asm volatile("prefetch.global.L1 [%0];" ::"l"(actions + (size_t)e * adim));
  Board board = load_board<NP>(env_map, W, H, lane, sm.bits);
  KP();
#ifdef PCGRL_PROFILE
  bool prof_reset = false, prof_changed = false;
#endif

  // binary: cells of components known to hold the longest path (binary_stats_update); nothing known at launch
  uint32_t best_cells = 0u;
  const bool incremental = (cfg.flags & PCGRL_FLAG_FULL_STATS) == 0;
  // one 32-bit row offset (t * n + e) is the only loop-carried index (rollout_dispatch checks T * n * adim < 2^31): the
  // base pointers are kernel parameters (constant bank operands), loop-carried 64-bit pointers were measured to spill
  uint32_t row = (uint32_t)e;
  for (int t = 0; t < T; t++, row += (uint32_t)n) {
    const int32_t* act = actions + (size_t)(row * (uint32_t)adim);
    int32_t a_cur = a_next;
    if (ONE_WORD) {
      act = &a_cur;
      if (t + 1 < T) a_next = actions[row + (uint32_t)n];
    } else if (t + 2 < T) {  // the action rows of the next steps are independent of the state: pull them towards L1 now
      asm volatile("prefetch.global.L1 [%0];" ::"l"(actions + (size_t)((row + 2u * (uint32_t)n) * (uint32_t)adim)));
    }
    iteration++;  // pcgrl_env.py:130
    // this step can end the episode through the change / iteration limits: start pulling what the reset will read
    if (auto_reset && (changes + max_change_per_step(rep) >= cfg.max_changes || iteration >= cfg.max_iterations))
      prefetch_reset_inputs(env_refs(cfg, b, e), lane);
    int old[NS];
#pragma unroll
    for (int i = 0; i < NS; i++) old[i] = st[i];
    int hx, hy, cell, tile, ex, ey, old_tile;
    bool multi;
    const int change = apply_action<REPT>(cfg, act, board, env_map, rng, lane, x, y, hx, hy, cell, tile, multi, ex, ey, old_tile);
    KP();
    if (change > 0) {  // pcgrl_env.py:135-138
      changes += change;
      if constexpr (PROB == PCGRL_PROB_BINARY) {
        // binary_prob.py:81-86 after a single-cell edit: only the components next to the cell are re-measured
        const uint32_t pass = type_mask<0x01u>(board, row_mask(W, H, lane));
        if (!multi && incremental) binary_stats_update(pass, cell_bit(ey, ex, lane), tile == 0, lane, st[0], st[1], best_cells, sm.draws);
        else if (T > 1) {  // multi-cell edit (or the A/B switch) inside a fused rollout: shared out-of-line copy
          const uint3 full = regions_and_longest_path_call(pass, lane);
          st[0] = (int)full.x; st[1] = (int)full.y; best_cells = full.z;
        } else regions_and_longest_path(pass, lane, st[0], st[1], best_cells);  // single step: inline, latency first
      } else {
        // zelda_prob.py:85: the region board holds every tile except solid (1) and door (4); a single-cell edit between two
        // tiles on the same side leaves calc_num_regions unchanged (59 % of random edits), so its floods are skipped
        constexpr unsigned REGION_TILES = 0xEDu;
        const bool same_side = !multi && (((REGION_TILES >> old_tile) & 1u) == ((REGION_TILES >> tile) & 1u)) &&
                               (cfg.flags & PCGRL_FLAG_FULL_STATS) == 0;
        int known_regions = same_side ? st[4] : -1;
        if (!same_side && !multi && (cfg.flags & PCGRL_FLAG_FULL_STATS) == 0) {
          // the cell entered / left the region board: the 3x3 window around it usually tells how many regions it touches
          const bool grew = ((REGION_TILES >> tile) & 1u) != 0u;
          const uint32_t p0 = type_mask<REGION_TILES>(board, row_mask(W, H, lane)) & ~cell_bit(ey, ex, lane);
          const int m = local_piece_count(p0, ex, ey);
          if (m >= 0) known_regions = st[4] + (grew ? 1 - m : m - 1);
        }
        bool unused;
        map_stats<PROB>(board, cfg, lane, st, unused, known_regions);
      }
    }
    KP();
#ifdef PCGRL_PROFILE
    prof_changed = change > 0;
    prof_reset = false;
#endif
    const double reward = (change > 0) ? problem_reward<PROB>(cfg, st, old) : 0.0;  // :142 (get_reward(s, s) == 0)
    const bool done = problem_over<PROB>(cfg, st, start) || changes >= cfg.max_changes ||
                      iteration >= cfg.max_iterations;                            // :143
    if (lane == 0) {
      if (reward_out) reward_out[row] = reward;
      if (done_out) done_out[row] = done ? 1 : 0;
    }
    if (t == T - 1) {  // the env's own reward / done / info buffers describe the last step only
      if (lane == 0) { b.reward[e] = reward; b.done[e] = done ? 1 : 0; }
      store_row<NS>(b.info_stats + (size_t)e * PCGRL_MAX_STATS, st, lane);
      store_info_counters(b.info_stats + (size_t)e * PCGRL_MAX_STATS, iteration, changes, lane);
      if (PROB == PCGRL_PROB_BINARY && lane == NS)  // info["path-imp"] (binary_prob.py:137), before any auto-reset
        b.info_stats[(size_t)e * PCGRL_MAX_STATS + NS] = st[1] - start[1];
    }
    if (done && auto_reset) {
      bool unused;
      env_reset<PROB>(cfg, b, e, lane, sm, rng, board, x, y, st, unused);
#pragma unroll
      for (int i = 0; i < NS; i++) start[i] = st[i];  // problem.py:45-46
      // written here, where they change, so that only the entries get_episode_over reads stay live in the step loop
      store_row<NS>(b.start_stats + (size_t)e * PCGRL_MAX_STATS, st, lane);
      iteration = 0;
      changes = 0;
      best_cells = 0u;
#ifdef PCGRL_PROFILE
      prof_reset = true;
#endif
      if (t == T - 1) write_record(sg, cfg, e, lane, reward, done, x, y, false, true, cell, tile, env_map);
    } else {
      if (change > 0) heat_increment(cfg, b.heatmap, (size_t)e * cells + (size_t)hy * W + hx, lane);   // :137
      if (t == T - 1) write_record(sg, cfg, e, lane, reward, done, x, y, change > 0, false, cell, tile, env_map, multi);
    }
  }
  rng.finish(lane);
  if (lane == 0) {
    if (rep != PCGRL_REP_WIDE) { b.pos[2 * e] = (uint8_t)x; b.pos[2 * e + 1] = (uint8_t)y; }
    b.iteration[e] = iteration;
    b.changes[e] = changes;
  }
  store_row<NS>(b.stats + (size_t)e * PCGRL_MAX_STATS, st, lane);
#ifdef PCGRL_PROFILE
  KP();
  if (T == 1 && lane == 0 && nkp == 5) {  // status int64[16 + 8*cls + k]: cls 0 = no change, 1 = changed, 2 = reset
    unsigned long long* acc = reinterpret_cast<unsigned long long*>(b.status) + 16 + 8 * (prof_reset ? 2 : prof_changed ? 1 : 0);
    atomicAdd(acc, 1ull);
    for (int k = 1; k < 5; k++) { atomicAdd(acc + k, (unsigned long long)(kp[k] - kp[k - 1])); }
    atomicMax(acc + 5, (unsigned long long)(kp[4] - kp[0]));
    atomicMax(acc + 6, (unsigned long long)(kp[3] - kp[2]));
  }
#endif
}

template <int PROB>
__global__ void __launch_bounds__(32 * WPB) k_reset(const __grid_constant__ pcgrl_config cfg,
                                                    const __grid_constant__ pcgrl_buffers b,
                                                    const uint8_t* __restrict__ mask, SolverQueue q, int n) {
  constexpr int NS = ProblemTraits<PROB>::NSTATS;
  __shared__ WarpSmem smem[WPB];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int e = blockIdx.x * WPB + wib;
  if (e >= n) return;
  if (mask && mask[e] == 0) return;
  const EnvRefs r = env_refs(cfg, b, e);
  WarpRng rng;
  rng.init(r.rng_rep);
  Board board;
  int x = 0, y = 0, st[NS];
  bool need_solver;
  env_reset<PROB>(cfg, b, e, lane, smem[wib], rng, board, x, y, st, need_solver);
  rng.finish(lane);
  if (lane == 0) {
    if (cfg.representation != PCGRL_REP_WIDE) { b.pos[2 * e] = (uint8_t)x; b.pos[2 * e + 1] = (uint8_t)y; }
    b.iteration[e] = 0;
    b.changes[e] = 0;
    b.reward[e] = 0.0;
    b.done[e] = 0;
  }
  store_row<NS>(b.stats + (size_t)e * PCGRL_MAX_STATS, st, lane);
  store_row<NS>(b.start_stats + (size_t)e * PCGRL_MAX_STATS, st, lane);
  store_row<NS>(b.info_stats + (size_t)e * PCGRL_MAX_STATS, st, lane);
  store_info_counters(b.info_stats + (size_t)e * PCGRL_MAX_STATS, 0, 0, lane);
  if (PROB == PCGRL_PROB_BINARY && lane == NS) b.info_stats[(size_t)e * PCGRL_MAX_STATS + NS] = 0;
  if constexpr (ProblemTraits<PROB>::SOLVER) if (need_solver) solver_enqueue(q, e, SOLVE_FOR_RESET, lane);
}

template <int PROB>
__global__ void __launch_bounds__(32 * WPB) k_get_stats(const __grid_constant__ pcgrl_config cfg,
                                                        const uint8_t* __restrict__ maps, int32_t* stats_out,
                                                        SolverQueue q, int n) {
  constexpr int NP = ProblemTraits<PROB>::NPLANES, NS = ProblemTraits<PROB>::NSTATS;
  __shared__ WarpSmem smem[WPB];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int e = blockIdx.x * WPB + wib;
  if (e >= n) return;
  const Board board = load_board<NP>(maps + (size_t)e * cfg.width * cfg.height, cfg.width, cfg.height, lane, smem[wib].bits);
  int st[NS];
  bool need_solver;
  map_stats_shared<PROB>(board, cfg, lane, st, need_solver);
  store_row<NS>(stats_out + (size_t)e * PCGRL_MAX_STATS, st, lane);
  if (lane >= NS && lane < PCGRL_MAX_STATS) stats_out[(size_t)e * PCGRL_MAX_STATS + lane] = 0;
  if constexpr (ProblemTraits<PROB>::SOLVER) if (need_solver) solver_enqueue(q, e, SOLVE_STATS_ONLY, lane);
}

// Solver problems, phase A of PcgrlEnv.step: Representation.update + map part of get_stats.
template <int PROB>
__global__ void __launch_bounds__(32 * WPB) k_step_update(const __grid_constant__ pcgrl_config cfg,
                                                          const __grid_constant__ pcgrl_buffers b,
                                                          const int32_t* __restrict__ actions, SolverQueue q,
                                                          int32_t* old_stats, uint8_t* heat_cell, int n) {
  constexpr int NP = ProblemTraits<PROB>::NPLANES, NS = ProblemTraits<PROB>::NSTATS;
  __shared__ WarpSmem smem[WPB];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int e = blockIdx.x * WPB + wib;
  if (e >= n) return;
  const int W = cfg.width, H = cfg.height;
  const int adim = action_dim(cfg.representation);
  const EnvRefs r = env_refs(cfg, b, e);
  Board board = load_board<NP>(r.map, W, H, lane, smem[wib].bits);
  int x = 0, y = 0;
  if (cfg.representation != PCGRL_REP_WIDE) { x = b.pos[2 * e]; y = b.pos[2 * e + 1]; }
  int st[NS];
  load_row<NS>(b.stats + (size_t)e * PCGRL_MAX_STATS, st);
  store_row<NS>(old_stats + (size_t)e * PCGRL_MAX_STATS, st, lane);  // old_stats = self._rep_stats (pcgrl_env.py:132)
  WarpRng rng;
  rng.init(r.rng_rep);
  int hx, hy, cell, tile;
  bool multi;
  const int change = apply_action(cfg, actions + (size_t)e * adim, board, r.map, rng, lane, x, y, hx, hy, cell, tile, multi);
  rng.finish(lane);
  if (lane == 0) {
    if (cfg.representation != PCGRL_REP_WIDE) { b.pos[2 * e] = (uint8_t)x; b.pos[2 * e + 1] = (uint8_t)y; }
    b.iteration[e] += 1;
    b.changes[e] += change;
    uint8_t* hc = heat_cell + 6 * (size_t)e;  // consumed by k_step_finish
    hc[0] = (uint8_t)change; hc[1] = (uint8_t)hx; hc[2] = (uint8_t)hy;
    hc[3] = (uint8_t)(cell & 0xff); hc[4] = (uint8_t)(cell >> 8); hc[5] = (uint8_t)(tile | (multi ? 0x80 : 0));
  }
  if (change > 0) {
    bool need_solver;
    map_stats_shared<PROB>(board, cfg, lane, st, need_solver);
    store_row<NS>(b.stats + (size_t)e * PCGRL_MAX_STATS, st, lane);
    if (need_solver) solver_enqueue(q, e, SOLVE_FOR_STEP, lane);
  }
}

// Solver problems, phase B: get_reward / get_episode_over / info, heat map, auto reset (whose new map may
// again need the solver -> second queue).
template <int PROB>
__global__ void __launch_bounds__(32 * WPB) k_step_finish(const __grid_constant__ pcgrl_config cfg,
                                                          const __grid_constant__ pcgrl_buffers b, SolverQueue q,
                                                          const int32_t* __restrict__ old_stats,
                                                          const uint8_t* __restrict__ heat_cell, int n, Staging sg,
                                                          double* reward_out, uint8_t* done_out) {
  constexpr int NS = ProblemTraits<PROB>::NSTATS;
  __shared__ WarpSmem smem[WPB];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int e = blockIdx.x * WPB + wib;
  if (e >= n) return;
  const int W = cfg.width, H = cfg.height, cells = W * H;
  int st[NS], old[NS], start[NS];
  load_row<NS>(b.stats + (size_t)e * PCGRL_MAX_STATS, st);
  load_row<NS>(old_stats + (size_t)e * PCGRL_MAX_STATS, old);
  load_row<NS>(b.start_stats + (size_t)e * PCGRL_MAX_STATS, start);
  const int iteration = b.iteration[e], changes = b.changes[e];
  const double reward = problem_reward<PROB>(cfg, st, old);
  const bool done = problem_over<PROB>(cfg, st, start) || changes >= cfg.max_changes || iteration >= cfg.max_iterations;
  if (lane == 0) {
    b.reward[e] = reward;
    b.done[e] = done ? 1 : 0;
    if (reward_out) reward_out[e] = reward;  // row t of the rollout outputs
    if (done_out) done_out[e] = done ? 1 : 0;
  }
  store_row<NS>(b.info_stats + (size_t)e * PCGRL_MAX_STATS, st, lane);
  store_info_counters(b.info_stats + (size_t)e * PCGRL_MAX_STATS, iteration, changes, lane);
  const uint8_t* hc = heat_cell + 6 * (size_t)e;
  const int change = hc[0], hx = hc[1], hy = hc[2], cell = hc[3] | (hc[4] << 8), tile = hc[5] & 0x7f;
  const bool multi = (hc[5] & 0x80) != 0;
  int px = 0, py = 0;
  if (cfg.representation != PCGRL_REP_WIDE) { px = b.pos[2 * e]; py = b.pos[2 * e + 1]; }
  if (done && (cfg.flags & PCGRL_FLAG_AUTO_RESET)) {
    const EnvRefs r = env_refs(cfg, b, e);
    WarpRng rng;
    rng.init(r.rng_rep);
    Board board;
    int x = 0, y = 0;
    bool need_solver;
    env_reset<PROB>(cfg, b, e, lane, smem[wib], rng, board, x, y, st, need_solver);
    rng.finish(lane);
    if (lane == 0) {
      if (cfg.representation != PCGRL_REP_WIDE) { b.pos[2 * e] = (uint8_t)x; b.pos[2 * e + 1] = (uint8_t)y; }
      b.iteration[e] = 0;
      b.changes[e] = 0;
    }
    store_row<NS>(b.stats + (size_t)e * PCGRL_MAX_STATS, st, lane);
    store_row<NS>(b.start_stats + (size_t)e * PCGRL_MAX_STATS, st, lane);
    if (need_solver) solver_enqueue(q, e, SOLVE_FOR_RESET, lane);
    write_record(sg, cfg, e, lane, reward, done, x, y, false, true, cell, tile, r.map);
  } else {
    if (change > 0) heat_increment(cfg, b.heatmap, (size_t)e * cells + (size_t)hy * W + hx, lane);
    write_record(sg, cfg, e, lane, reward, done, px, py, change > 0, false, cell, tile, b.map + (size_t)e * cells, multi);
  }
  (void)H;
}

// ------------------------------------------------------------------------------------------------
// Solver problems, T-step rollouts: ENV-ASYNCHRONOUS persistent kernel.
//
// The lock-step pipeline above (update -> k_solve -> finish -> k_solve) makes every batched step wait for the
// slowest search of the batch.  Environments never interact, so for a rollout whose actions are known in advance
// each env can run through its T steps on its own clock: a warp pulls an env index from a global counter, keeps the
// bitboards / cursor / statistics in registers like k_rollout, and when the map satisfies the solver precondition
// it runs _run_game INLINE: the four passes in the reference's order, stopping at the first win
// (sokoban_prob.py:110-122, ddave_prob.py:122-135, mdungeon_prob.py:125-138).
//
// A CTA holds ASYNC_WPB env warps and ONE search arena (visited table + node ring + heap, 94.5 KB at power 5000)
// guarded by a shared-memory lock, plus two lock-guarded reset staging areas; two CTAs fit one SM.
//
// Tail: the rollout ends when the slowest env does, and a single search that runs all four passes to the
// iteration cap costs ~20 ms.  So the owner POSTS passes 1..3 of every search in a request slot in HBM; warps that
// have run out of envs become helpers: they claim posted passes (atomicAnd on open_mask), run them on their own
// CTA's arena from the env's map in HBM, and publish the result.  The owner runs pass 0, then every pass nobody has
// claimed, then waits for the claimed ones (a helper never blocks, so the wait is finite) and merges in the
// reference's order exactly like k_solve (first winning pass; an exhausted pass proves that nobody wins).  While
// all warps still have envs of their own there are no helpers and no speculative work at all.
// ------------------------------------------------------------------------------------------------
#define ASYNC_RESET_AREAS 2

struct ArenaHead {
  Level L;
  SState root;
  int res[4];
  int exhausted;
  int lock;
  int reset_lock[ASYNC_RESET_AREAS];
  int helper;
};

// nanosleep may return early (PTX only bounds it from above), so waiting loops pace themselves with the global timer
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void pause_ns(unsigned ns) {
  const unsigned long long t0 = global_ns();
  do { __nanosleep(ns); } while (global_ns() - t0 < ns);
}
__device__ __forceinline__ void arena_lock(int* lock, int lane) {
  if (lane == 0) {
    while (atomicCAS(lock, 0, 1) != 0) pause_ns(1000);
    __threadfence_block();
  }
  __syncwarp();
}
__device__ __forceinline__ bool arena_trylock(int* lock, int lane) {
  int got = 0;
  if (lane == 0) {
    got = (atomicCAS(lock, 0, 1) == 0) ? 1 : 0;
    __threadfence_block();
  }
  return __shfl_sync(FULL_MASK, got, 0) != 0;
}
__device__ __forceinline__ void arena_unlock(int* lock, int lane) {
  __syncwarp();
  if (lane == 0) {
    __threadfence_block();
    atomicExch(lock, 0);
  }
}
__device__ __forceinline__ int ld_volatile(const int32_t* p) { return *reinterpret_cast<const volatile int32_t*>(p); }

struct ArenaRefs {
  ArenaHead* A;
  uint32_t *table, *cache, *nodes;
  HeapRef heap;
  int table_size;
};

// One pass of _run_game on the level in A.L (arena lock held); publishes the result in the request slot.
template <int GAME>
__device__ __noinline__ void async_run_pass(const pcgrl_config& cfg, const ArenaRefs& R, AsyncGroup* G, int pass, int lane) {
  ArenaHead& A = *R.A;
  const int b = (GAME == GAME_SOKOBAN) ? ((pass == 0) ? -1 : (pass == 1) ? 2 : (pass == 2) ? 1 : 0)
                                       : ((pass == 0) ? 2 : (pass == 1) ? 1 : (pass == 2) ? 0 : -1);
  __syncwarp();
  uint4* t4 = reinterpret_cast<uint4*>(R.table);
  for (int i = lane; i < (R.table_size >> 2); i += 32) t4[i] = make_uint4(0u, 0u, 0u, 0u);
  if (lane == 0) { A.res[0] = 0; A.res[1] = 0; A.res[2] = 0; A.res[3] = 0; A.exhausted = 0; }
  __syncwarp();
  const SState root = A.root;
  if (b < 0)
    search_bfs_batched<GAME>(A.L, root, cfg.solver_power, R.nodes, R.table, R.table_size - 1, &G->best_win, pass, A.res, &A.exhausted, lane);
  else
    search_pass<GAME>(A.L, root, b, cfg.solver_power, R.nodes, R.cache, R.heap, R.table, R.table_size - 1, &G->best_win, pass, A.res, &A.exhausted, lane);
  __syncwarp();
  if (lane == 0) {
    if (A.res[0] == 1) atomicMin(&G->best_win, pass);
    // exhaustion rule: see k_solve
    if (A.res[0] == 0 && A.exhausted && (GAME == GAME_SOKOBAN || (GAME == GAME_MDUNGEON && b < 0))) atomicMin(&G->best_win, -1 - pass);
    G->results[pass * 4 + 0] = A.res[0]; G->results[pass * 4 + 1] = A.res[1];
    G->results[pass * 4 + 2] = A.res[2]; G->results[pass * 4 + 3] = A.res[3];
    __threadfence();
    atomicOr(&G->done_mask, 1 << pass);
  }
  __syncwarp();
}

// _run_game for the map held in `board` (== the env's map in HBM); arena lock held by the caller.  Patches the
// play-through statistics in st.
template <int PROB>
__device__ __noinline__ void solve_inline(const pcgrl_config& cfg, const Board& board, const ArenaRefs& R, AsyncHeader* hdr,
                                          AsyncGroup* G, int e, int lane, int* st, int32_t* status) {
  constexpr int GAME = GameOf<PROB>::GAME;
  ArenaHead& A = *R.A;
  const int W = cfg.width, H = cfg.height;
  __syncwarp();
  if (lane < H)
    for (int x = 0; x < W; x++) A.L.tiles[lane * W + x] = (uint8_t)tile_at(board, x);
  __threadfence();  // helpers read this env's map from HBM
  __syncwarp();
  if (lane == 0) {
    SState root;
    level_init<GAME>(A.L, root, W, H);
    A.root = root;
    if (A.L.overflow && status) atomicExch(status, 1);
  }
  __syncwarp();
  int won = 0, depth = 0, h = 0;
  uint32_t misc = 0;
  if (!A.L.overflow) {
    if (lane == 0) {
      G->env = e;
      G->best_win = 4;
      G->done_mask = 0;
      __threadfence();
      atomicExch(&G->open_mask, 0xE);
      atomicAdd(&hdr->posted, 3);
    }
    __syncwarp();
    async_run_pass<GAME>(cfg, R, G, 0, lane);
    for (int pass = 1; pass < 4; pass++) {
      int mine = 0, skip = 0;
      if (lane == 0) {
        const int bw = ld_volatile(&G->best_win);
        // once an earlier pass has won (or some pass exhausted) every remaining unclaimed pass is dropped at once
        const int want = (bw < pass) ? (0xE & ~((1 << pass) - 1)) : (1 << pass);
        const int got = atomicAnd(&G->open_mask, ~want) & want;
        if (got) atomicSub(&hdr->posted, __popc(got));
        if (bw < pass) {
          for (int p = pass; p < 4; p++) if ((got >> p) & 1) G->results[p * 4] = -1;
          if (got) { __threadfence(); atomicOr(&G->done_mask, got); }
          skip = 1;
        } else {
          mine = (got >> pass) & 1;
        }
      }
      mine = __shfl_sync(FULL_MASK, mine, 0);
      skip = __shfl_sync(FULL_MASK, skip, 0);
      if (skip) break;
      if (mine) async_run_pass<GAME>(cfg, R, G, pass, lane);
    }
    if (lane == 0) {
      while ((ld_volatile(&G->done_mask) & 0xF) != 0xF) pause_ns(2000);
      __threadfence();
      const volatile int32_t* rr = G->results;
      const int bw = ld_volatile(&G->best_win);
      int sel = 3;
      if (bw < 0) sel = -1 - bw;  // exhausted pass: nobody can win
      else for (int p = 0; p < 4; p++) if (rr[p * 4] == 1) { sel = p; break; }
      A.res[0] = (rr[sel * 4] == 1) ? 1 : 0; A.res[1] = rr[sel * 4 + 1]; A.res[2] = rr[sel * 4 + 2]; A.res[3] = rr[sel * 4 + 3];
    }
    __syncwarp();
    won = A.res[0]; depth = A.res[1]; h = A.res[2]; misc = (uint32_t)A.res[3];
  }
  __syncwarp();
  const int dist_win = won ? 0 : h, sol_len = won ? depth : 0;
  if (GAME == GAME_SOKOBAN) {  // sokoban_prob.py:110-122,143-144
    st[4] = dist_win; st[5] = sol_len;
  } else if (GAME == GAME_DDAVE) {  // ddave_prob.py:122-135,164-168
    st[9] = dist_win; st[10] = sol_len; st[7] = (int)(misc >> 16); st[8] = (int)((misc >> 8) & 0xffu);
  } else {  // mdungeon_prob.py:125-138,166-170
    st[9] = dist_win; st[10] = sol_len;
    st[6] = (int)(misc & 0xffu); st[7] = (int)((misc >> 8) & 0xffu); st[8] = (int)((misc >> 16) & 0xffu);
  }
}

// A warp without envs of its own: claim a posted pass of somebody else's search and run it on this CTA's arena.
template <int PROB>
__device__ __noinline__ void async_help(const pcgrl_config& cfg, const pcgrl_buffers& b, const ArenaRefs& R, AsyncHeader* hdr,
                                        AsyncGroup* groups, int ngroups, int n, int lane) {
  constexpr int GAME = GameOf<PROB>::GAME;
  ArenaHead& A = *R.A;
  const int W = cfg.width, H = cfg.height, cells = W * H;
  int start = (blockIdx.x * ASYNC_WPB + (threadIdx.x >> 5)) % ngroups;
  while (true) {
    int stop = 0, has = 0;
    if (lane == 0) { stop = ld_volatile(&hdr->envs_done) >= n; has = ld_volatile(&hdr->posted) > 0; }
    stop = __shfl_sync(FULL_MASK, stop, 0);
    has = __shfl_sync(FULL_MASK, has, 0);
    if (stop) break;
    if (!has) { pause_ns(10000); continue; }
    if (!arena_trylock(&A.lock, lane)) { pause_ns(10000); continue; }
    int fg = -1, fp = 0;
    for (int g0 = 0; g0 < ngroups && fg < 0; g0 += 32) {
      int g = start + g0 + lane;
      if (g >= ngroups) g -= ngroups;
      const int m = (g0 + lane < ngroups) ? ld_volatile(&groups[g].open_mask) : 0;
      unsigned cand = __ballot_sync(FULL_MASK, m != 0);
      while (cand && fg < 0) {
        const int src = __ffs(cand) - 1;
        cand &= cand - 1;
        int ok = 0, p = 0;
        if (lane == src) {
          p = __ffs(m) - 1;
          ok = (atomicAnd(&groups[g].open_mask, ~(1 << p)) >> p) & 1;
          if (ok) { atomicSub(&hdr->posted, 1); __threadfence(); }
        }
        ok = __shfl_sync(FULL_MASK, ok, src);
        if (ok) { fg = __shfl_sync(FULL_MASK, g, src); fp = __shfl_sync(FULL_MASK, p, src); }
      }
    }
    if (fg >= 0) {
      AsyncGroup* G = groups + fg;
      start = fg;
      int skip = 0, e = 0;
      if (lane == 0) {
        e = ld_volatile(&G->env);
        if (ld_volatile(&G->best_win) < fp) {  // already decided: nothing to run
          G->results[fp * 4] = -1;
          __threadfence();
          atomicOr(&G->done_mask, 1 << fp);
          skip = 1;
        }
      }
      skip = __shfl_sync(FULL_MASK, skip, 0);
      e = __shfl_sync(FULL_MASK, e, 0);
      if (!skip) {
        const uint8_t* gm = b.map + (size_t)e * cells;
        for (int i = lane; i < cells; i += 32) A.L.tiles[i] = __ldcg(gm + i);
        __syncwarp();
        if (lane == 0) {
          SState root;
          level_init<GAME>(A.L, root, W, H);
          A.root = root;
        }
        __syncwarp();
        async_run_pass<GAME>(cfg, R, G, fp, lane);
      }
    }
    arena_unlock(&A.lock, lane);
    if (fg < 0) pause_ns(5000);
  }
  (void)H;
}

template <int PROB>
__global__ void __launch_bounds__(32 * ASYNC_WPB, ASYNC_MIN_CTAS) k_rollout_async(const __grid_constant__ pcgrl_config cfg,
                                                                     const __grid_constant__ pcgrl_buffers b,
                                                                     const int32_t* __restrict__ actions, double* reward_out,
                                                                     uint8_t* done_out, int T, int n, AsyncHeader* hdr,
                                                                     AsyncGroup* groups, uint32_t* node_pool,
                                                                     size_t nodes_per_pass, int table_size, Staging sg,
                                                                     uint32_t* heap_pool, size_t heap_words, int heap_fast) {
  constexpr int NP = ProblemTraits<PROB>::NPLANES, NS = ProblemTraits<PROB>::NSTATS;
  extern __shared__ __align__(16) uint32_t dyn[];
  __shared__ ArenaHead A;
  __shared__ uint32_t wbits[ASYNC_WPB][3 * PCGRL_SBITS_STRIDE];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  ArenaRefs R;
  R.A = &A;
  R.table = dyn;
  R.cache = dyn + table_size;
  uint32_t* heap_fast_p = R.cache + SOLVER_CACHE_NODES * SOLVER_NODE_WORDS;  // the first heap_fast open-list entries (top levels)
  R.heap = heap_ref(heap_fast_p, heap_pool + (size_t)blockIdx.x * heap_words, heap_fast);  // tail: this CTA's slice of HBM scratch
  R.nodes = node_pool + (size_t)blockIdx.x * nodes_per_pass * SOLVER_NODE_WORDS;
  R.table_size = table_size;
  WarpSmem* reset_areas = reinterpret_cast<WarpSmem*>(heap_fast_p + (((size_t)heap_fast + 3) & ~(size_t)3));
  const int my_area = wib % ASYNC_RESET_AREAS;
  WarpSmem& reset_sm = reset_areas[my_area];
  AsyncGroup* G = groups + blockIdx.x * ASYNC_WPB + wib;
  const int ngroups = gridDim.x * ASYNC_WPB;
  if (threadIdx.x == 0) {
    A.lock = 0;
    A.helper = 0;
    for (int i = 0; i < ASYNC_RESET_AREAS; i++) A.reset_lock[i] = 0;
  }
  __syncthreads();
  const int W = cfg.width, H = cfg.height, cells = W * H;
  const int adim = action_dim(cfg.representation);
  const bool auto_reset = (cfg.flags & PCGRL_FLAG_AUTO_RESET) != 0;

#ifdef PCGRL_PROFILE
  // status int64[64 + k]: 0 searches, 1 search cycles, 2 lock-wait cycles, 3 resets, 4 reset cycles, 5 envs,
  // 6 env cycles, 7 max env cycles, 8 max search cycles, 9 max warp cycles (all its envs)
  unsigned long long* prof = reinterpret_cast<unsigned long long*>(b.status) + 64;
  const long long warp_t0 = clock64();
#define AP_ADD(k, v) do { if (lane == 0) atomicAdd(prof + (k), (unsigned long long)(v)); } while (0)
#define AP_MAX(k, v) do { if (lane == 0) atomicMax(prof + (k), (unsigned long long)(v)); } while (0)
#define AP_NOW() clock64()
#else
#define AP_ADD(k, v) do {} while (0)
#define AP_MAX(k, v) do {} while (0)
#define AP_NOW() 0ll
#endif
  while (true) {
    int e = 0;
    if (lane == 0) e = atomicAdd(&hdr->work, 1);
    e = __shfl_sync(FULL_MASK, e, 0);
    if (e >= n) break;
    const long long env_t0 = AP_NOW();
    const EnvRefs r = env_refs(cfg, b, e);
    WarpRng rng;
    rng.init(r.rng_rep, (cfg.representation == PCGRL_REP_NARROW || cfg.representation >= PCGRL_REP_NARROWCAST) ? lane : -1);
    int x = 0, y = 0;
    if (cfg.representation != PCGRL_REP_WIDE) { x = b.pos[2 * e]; y = b.pos[2 * e + 1]; }
    int iteration = b.iteration[e], changes = b.changes[e];
    int st[NS], start[NS];
    load_row<NS>(b.stats + (size_t)e * PCGRL_MAX_STATS, st);
    load_row<NS>(b.start_stats + (size_t)e * PCGRL_MAX_STATS, start);
    Board board = load_board<NP>(r.map, W, H, lane, wbits[wib]);

    for (int t = 0; t < T; t++) {
      const int32_t* act = actions + ((size_t)t * n + e) * adim;
      iteration++;  // pcgrl_env.py:130
      int old[NS];
#pragma unroll
      for (int i = 0; i < NS; i++) old[i] = st[i];
      int hx, hy, cell, tile;
      bool multi;
      const int change = apply_action(cfg, act, board, r.map, rng, lane, x, y, hx, hy, cell, tile, multi);
      if (change > 0) {  // pcgrl_env.py:135-138
        changes += change;
        bool need_solver;
        map_stats<PROB>(board, cfg, lane, st, need_solver);
        if (need_solver) {
          const long long t0 = AP_NOW();
          arena_lock(&A.lock, lane);
          const long long t1 = AP_NOW();
          solve_inline<PROB>(cfg, board, R, hdr, G, e, lane, st, b.status);
          arena_unlock(&A.lock, lane);
          const long long t2 = AP_NOW();
          AP_ADD(0, 1); AP_ADD(1, t2 - t1); AP_ADD(2, t1 - t0); AP_MAX(8, t2 - t1);
          (void)t0; (void)t1; (void)t2;
        }
      }
      const double reward = (change > 0) ? problem_reward<PROB>(cfg, st, old) : 0.0;  // :142
      const bool done = problem_over<PROB>(cfg, st, start) || changes >= cfg.max_changes ||
                        iteration >= cfg.max_iterations;                              // :143
      if (lane == 0) {
        if (reward_out) reward_out[(size_t)t * n + e] = reward;
        if (done_out) done_out[(size_t)t * n + e] = done ? 1 : 0;
        if (t == T - 1) { b.reward[e] = reward; b.done[e] = done ? 1 : 0; }
      }
      if (t == T - 1) {
        store_row<NS>(b.info_stats + (size_t)e * PCGRL_MAX_STATS, st, lane);
        store_info_counters(b.info_stats + (size_t)e * PCGRL_MAX_STATS, iteration, changes, lane);
      }
      if (done && auto_reset) {
        bool need_solver;
        const long long t0 = AP_NOW();
        arena_lock(&A.reset_lock[my_area], lane);
        const long long t1 = AP_NOW();
        env_reset<PROB>(cfg, b, e, lane, reset_sm, rng, board, x, y, st, need_solver);
        arena_unlock(&A.reset_lock[my_area], lane);
        const long long t2 = AP_NOW();
        AP_ADD(2, t1 - t0); AP_ADD(3, 1); AP_ADD(4, t2 - t1);
        if (need_solver) {
          arena_lock(&A.lock, lane);
          const long long t3 = AP_NOW();
          solve_inline<PROB>(cfg, board, R, hdr, G, e, lane, st, b.status);
          arena_unlock(&A.lock, lane);
          const long long t4 = AP_NOW();
          AP_ADD(0, 1); AP_ADD(1, t4 - t3); AP_ADD(2, t3 - t2); AP_MAX(8, t4 - t3);
          (void)t3; (void)t4;
        }
        (void)t0; (void)t1; (void)t2;
#pragma unroll
        for (int i = 0; i < NS; i++) start[i] = st[i];  // problem.py:45-46
        iteration = 0;
        changes = 0;
        if (t == T - 1) write_record(sg, cfg, e, lane, reward, done, x, y, false, true, cell, tile, r.map);
      } else {
        if (change > 0) heat_increment(cfg, b.heatmap, (size_t)e * cells + (size_t)hy * W + hx, lane);  // :137
        if (t == T - 1) write_record(sg, cfg, e, lane, reward, done, x, y, change > 0, false, cell, tile, r.map, multi);
      }
    }
    rng.finish(lane);
    if (lane == 0) {
      if (cfg.representation != PCGRL_REP_WIDE) { b.pos[2 * e] = (uint8_t)x; b.pos[2 * e + 1] = (uint8_t)y; }
      b.iteration[e] = iteration;
      b.changes[e] = changes;
    }
    store_row<NS>(b.stats + (size_t)e * PCGRL_MAX_STATS, st, lane);
    store_row<NS>(b.start_stats + (size_t)e * PCGRL_MAX_STATS, start, lane);
    __syncwarp();
    if (lane == 0) atomicAdd(&hdr->envs_done, 1);
    AP_ADD(5, 1); AP_ADD(6, AP_NOW() - env_t0); AP_MAX(7, AP_NOW() - env_t0);
    (void)env_t0;
  }
#ifdef PCGRL_PROFILE
  AP_MAX(9, clock64() - warp_t0);
#endif
  // one helper per CTA (there is one arena): the first warp that runs out of envs; the others are done
  int first = 0;
  if (lane == 0) first = (atomicExch(&A.helper, 1) == 0) ? 1 : 0;
  if (__shfl_sync(FULL_MASK, first, 0)) async_help<PROB>(cfg, b, R, hdr, groups, ngroups, n, lane);
}

__global__ void k_seed(uint32_t* rng, const uint32_t* __restrict__ seeds, int n) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  uint32_t* a = rng + (size_t)e * 2 * PCGRL_MT_WORDS;
  uint32_t* c = a + PCGRL_MT_WORDS;
  uint32_t v = seeds[e];
  a[0] = v; c[0] = v;
  for (int i = 1; i < 624; i++) {  // init_genrand
    v = 1812433253u * (v ^ (v >> 30)) + (uint32_t)i;
    a[i] = v; c[i] = v;
  }
  a[624] = 624; c[624] = 624;  // RandomState(seed).get_state()[2]
}

// ------------------------------------------------------------------------------------------------
// host side: C ABI
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int fail(int rc, const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return rc;
}
static int cuda_rc(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
  return (int)e;
}

extern "C" int pcgrl_abi_version(void) { return PCGRL_ABI_VERSION; }
extern "C" const char* pcgrl_last_error(void) { return g_err; }

extern "C" int pcgrl_config_validate(const pcgrl_config* c) {
  static const int ntiles[PCGRL_NUM_PROBLEMS] = {2, 8, 5, 7, 8, 7};
  if (!c) return fail(-1, "config is NULL");
  if (c->problem < 0 || c->problem >= PCGRL_NUM_PROBLEMS) return fail(-1, "unknown problem id");
  if (c->representation < 0 || c->representation >= PCGRL_NUM_REPS) return fail(-1, "unknown representation id");
  if (c->problem == PCGRL_PROB_SMB) {
    if (c->width < 1 || c->width > PCGRL_SMB_MAX_W || c->height < 3 || c->height > PCGRL_SMB_MAX_H)
      return fail(-1, "smb: 1 <= width <= 122 and 3 <= height <= 16");
    if (c->solver_power < 1 || c->solver_power > pcgrl_smb::MAX_POWER) return fail(-1, "smb: solver_power must be in [1, 16000]");
  } else if (c->width < 1 || c->width > PCGRL_MAX_DIM || c->height < 1 || c->height > PCGRL_MAX_DIM)
    return fail(-1, "width/height must be in [1, 32] (one bitboard row per warp lane)");
  if (c->num_tiles != ntiles[c->problem]) return fail(-1, "num_tiles does not match the problem's tile alphabet");
  if (c->max_changes < 1 || c->max_changes > 65535) return fail(-1, "max_changes must be in [1, 65535]");
  if (c->max_changes > 255 && !(c->flags & PCGRL_FLAG_HEAT_U16))
    return fail(-1, "max_changes > 255 needs a uint16 heat map: set PCGRL_FLAG_HEAT_U16");
  if (c->max_iterations < 1) return fail(-1, "max_iterations must be >= 1");
  double tot = 0;
  for (int t = 0; t < c->num_tiles; t++) {
    if (!(c->tile_prob[t] >= 0)) return fail(-1, "tile probabilities must be >= 0");
    tot += c->tile_prob[t];
  }
  if (!(tot > 0)) return fail(-1, "tile probabilities must not all be zero");
  if (c->problem >= PCGRL_PROB_SOKOBAN && c->problem != PCGRL_PROB_SMB) {
    const int rc = solver_validate(c);
    if (rc) return fail(-1, solver_validate_message(rc));
  }
  return 0;
}

extern "C" size_t pcgrl_scratch_bytes(const pcgrl_config* c, int n_envs) {
  if (!c || n_envs <= 0 || c->problem < PCGRL_PROB_SOKOBAN) return 0;
  if (c->problem == PCGRL_PROB_SMB) return pcgrl_smb::scratch_bytes(n_envs, c->solver_power);
  return solver_scratch_bytes(c, n_envs);
}

static int check_common(const pcgrl_config* cfg, const pcgrl_buffers* b, int n) {
  if (!cfg || !b) return fail(-1, "NULL config / buffers");
  if (n <= 0) return fail(-1, "n must be > 0");
  const int rc = pcgrl_config_validate(cfg);
  if (rc) return rc;
  if (!b->map || !b->heatmap || !b->pos || !b->iteration || !b->changes || !b->stats || !b->start_stats ||
      !b->info_stats || !b->reward || !b->done || !b->rng || !b->tile_prob || !b->start_map || !b->start_valid ||
      !b->status)
    return fail(-1, "a pcgrl_buffers pointer is NULL");
  if (cfg->problem >= PCGRL_PROB_SOKOBAN && (!b->scratch || b->scratch_bytes < pcgrl_scratch_bytes(cfg, n)))
    return fail(-1, "scratch buffer too small: see pcgrl_scratch_bytes()");
  return 0;
}

static inline int action_dim_host(int rep) {
  return rep == PCGRL_REP_WIDE ? 3 : (rep == PCGRL_REP_NARROWCAST || rep == PCGRL_REP_TURTLECAST) ? 2
       : rep == PCGRL_REP_NARROWMULTI ? 9 : 1;
}

static inline dim3 env_grid(int n) { return dim3((unsigned)((n + WPB - 1) / WPB)); }

// ---- smb: persistent warp-per-env kernels (pcgrl_smb_env.cuh) ----
static int smb_plan(const pcgrl_config* cfg, int n, pcgrl_smb::Launch* out) {
  static thread_local int sm_dev = -1, sm_count = 0;
  static thread_local size_t configured[3] = {0, 0, 0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (sm_dev != dev) {
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
    sm_dev = dev;
    configured[0] = configured[1] = configured[2] = 0;
  }
  *out = pcgrl_smb::launch_plan(cfg, n, sm_count > 0 ? sm_count : 148);
  const void* fns[3] = {(const void*)pcgrl_smb::k_smb_rollout, (const void*)pcgrl_smb::k_smb_reset, (const void*)pcgrl_smb::k_smb_get_stats};
  for (int i = 0; i < 3; i++)
    if (configured[i] < out->smem) {
      cudaError_t ce = cudaFuncSetAttribute(fns[i], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)out->smem);
      if (ce != cudaSuccess) return cuda_rc(ce, "smb shared memory opt-in");
      configured[i] = out->smem;
    }
  return 0;
}

static int smb_reset(const pcgrl_config* cfg, const pcgrl_buffers* b, const uint8_t* mask, int n, cudaStream_t s) {
  pcgrl_smb::Launch L;
  int rc = smb_plan(cfg, n, &L);
  if (rc) return rc;
  pcgrl_smb::Scratch sc = pcgrl_smb::scratch_view(b->scratch, n, cfg->solver_power);
  cudaMemsetAsync(sc.work, 0, sizeof(int32_t), s);
  pcgrl_smb::k_smb_reset<<<L.grid, 32 * SMB_WPB, L.smem, s>>>(*cfg, *b, mask, n, sc, L.fast_cap, L.per_warp_bytes);
  return cuda_rc(cudaGetLastError(), "pcgrl_reset (smb) launch");
}

static int smb_rollout(const pcgrl_config* cfg, const pcgrl_buffers* b, const int32_t* actions, double* reward_out,
                       uint8_t* done_out, int T, int n, cudaStream_t s, Staging sg) {
  pcgrl_smb::Launch L;
  int rc = smb_plan(cfg, n, &L);
  if (rc) return rc;
  pcgrl_smb::Scratch sc = pcgrl_smb::scratch_view(b->scratch, n, cfg->solver_power);
  cudaMemsetAsync(sc.work, 0, sizeof(int32_t), s);
  pcgrl_smb::k_smb_rollout<<<L.grid, 32 * SMB_WPB, L.smem, s>>>(*cfg, *b, actions, reward_out, done_out, T, n, sc, sg, L.fast_cap,
                                                             L.per_warp_bytes);
  return cuda_rc(cudaGetLastError(), "pcgrl_step (smb) launch");
}

static int smb_get_stats(const pcgrl_config* cfg, const uint8_t* maps, int32_t* out, int n, void* scratch, cudaStream_t s) {
  pcgrl_smb::Launch L;
  int rc = smb_plan(cfg, n, &L);
  if (rc) return rc;
  pcgrl_smb::Scratch sc = pcgrl_smb::scratch_view(scratch, n, cfg->solver_power);
  cudaMemsetAsync(sc.work, 0, sizeof(int32_t), s);
  pcgrl_smb::k_smb_get_stats<<<L.grid, 32 * SMB_WPB, L.smem, s>>>(*cfg, maps, out, n, sc, L.fast_cap, L.per_warp_bytes);
  return cuda_rc(cudaGetLastError(), "pcgrl_get_stats (smb) launch");
}

template <int PROB>
static int reset_impl(const pcgrl_config* cfg, const pcgrl_buffers* b, const uint8_t* mask, int n, cudaStream_t s) {
  SolverQueue q = solver_queue(cfg, b->scratch, n, 0, b->status);
  if constexpr (ProblemTraits<PROB>::SOLVER) solver_queue_clear(q, s);
  k_reset<PROB><<<env_grid(n), 32 * WPB, 0, s>>>(*cfg, *b, mask, q, n);
  if constexpr (ProblemTraits<PROB>::SOLVER) solver_launch<PROB>(cfg, b->stats, b->start_stats, b->map, q, b->scratch, n, s);
  return cuda_rc(cudaGetLastError(), "pcgrl_reset launch");
}

extern "C" int pcgrl_reset(const pcgrl_config* cfg, const pcgrl_buffers* b, const uint8_t* mask, int n, void* stream) {
  int rc = check_common(cfg, b, n);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  switch (cfg->problem) {
    case PCGRL_PROB_BINARY: return reset_impl<PCGRL_PROB_BINARY>(cfg, b, mask, n, s);
    case PCGRL_PROB_ZELDA: return reset_impl<PCGRL_PROB_ZELDA>(cfg, b, mask, n, s);
    case PCGRL_PROB_SOKOBAN: return reset_impl<PCGRL_PROB_SOKOBAN>(cfg, b, mask, n, s);
    case PCGRL_PROB_DDAVE: return reset_impl<PCGRL_PROB_DDAVE>(cfg, b, mask, n, s);
    case PCGRL_PROB_SMB: return smb_reset(cfg, b, mask, n, s);
    default: return reset_impl<PCGRL_PROB_MDUNGEON>(cfg, b, mask, n, s);
  }
}

template <int PROB>
static int rollout_fused(const pcgrl_config* cfg, const pcgrl_buffers* b, const int32_t* actions, double* reward_out,
                         uint8_t* done_out, int T, int n, cudaStream_t s, Staging sg) {
  if constexpr (PROB == PCGRL_PROB_BINARY) {
    // Rollouts of maps with at most 16 (8) rows: two (four) envs per warp, decoupled in time (pcgrl_packed.cuh).
    // OPT-IN (PCGRL_PACKED=1): bit-exact, but measured SLOWER than one env per warp on B200 at every batch size
    // (binary-narrow 16x16, 128 steps per launch: 6.2e8 vs 9.3e8 env-steps/s at 4096 envs, 9.0e8 vs 1.35e9 at 65536;
    // profiles/r02_summary.md): the per-group state machine costs ~15 instructions per wave against 10.5, a packed
    // pass lasts as long as its slower env, and half as many warps hide less latency.  Kept for A/B runs.
    static const bool packed = getenv("PCGRL_PACKED") && atoi(getenv("PCGRL_PACKED")) == 1;
    if (packed && !sg.base && !sg.direct && T >= PCGRL_PACKED_MIN_T && cfg->height <= 16 && cfg->representation <= PCGRL_REP_WIDE) {
      if (cfg->height <= 8) {
        const int epc = PACKED_WPB * 4;
        k_rollout_packed_binary<8><<<(n + epc - 1) / epc, 32 * PACKED_WPB, 0, s>>>(*cfg, *b, actions, reward_out, done_out, T, n);
      } else {
        const int epc = PACKED_WPB * 2;
        k_rollout_packed_binary<16><<<(n + epc - 1) / epc, 32 * PACKED_WPB, 0, s>>>(*cfg, *b, actions, reward_out, done_out, T, n);
      }
      return cuda_rc(cudaGetLastError(), "pcgrl_rollout (packed) launch");
    }
  }
  switch (cfg->representation) {  // the three base representations get their own instantiation
    case PCGRL_REP_NARROW: k_rollout<PROB, PCGRL_REP_NARROW><<<env_grid(n), 32 * WPB, 0, s>>>(*cfg, *b, actions, reward_out, done_out, T, n, sg); break;
    case PCGRL_REP_TURTLE: k_rollout<PROB, PCGRL_REP_TURTLE><<<env_grid(n), 32 * WPB, 0, s>>>(*cfg, *b, actions, reward_out, done_out, T, n, sg); break;
    case PCGRL_REP_WIDE: k_rollout<PROB, PCGRL_REP_WIDE><<<env_grid(n), 32 * WPB, 0, s>>>(*cfg, *b, actions, reward_out, done_out, T, n, sg); break;
    default: k_rollout<PROB, -1><<<env_grid(n), 32 * WPB, 0, s>>>(*cfg, *b, actions, reward_out, done_out, T, n, sg); break;
  }
  return cuda_rc(cudaGetLastError(), "pcgrl_step launch");
}

// one PcgrlEnv.step of a solver problem = update -> solver -> finish(+reset) -> solver
template <int PROB>
static int step_solver(const pcgrl_config* cfg, const pcgrl_buffers* b, const int32_t* actions, int n, cudaStream_t s,
                       Staging sg, int max_slots = SOLVER_MAX_SLOTS, double* reward_out = nullptr, uint8_t* done_out = nullptr) {
  SolverQueue q1 = solver_queue(cfg, b->scratch, n, 0, b->status), q2 = solver_queue(cfg, b->scratch, n, 1, b->status);
  int32_t* old_stats = solver_old_stats(cfg, b->scratch, n);
  uint8_t* heat_cell = solver_heat_cell(cfg, b->scratch, n);
  solver_queue_clear2(q1, q2, s);
  k_step_update<PROB><<<env_grid(n), 32 * WPB, 0, s>>>(*cfg, *b, actions, q1, old_stats, heat_cell, n);
  solver_launch<PROB>(cfg, b->stats, b->start_stats, b->map, q1, b->scratch, n, s, max_slots);
  k_step_finish<PROB><<<env_grid(n), 32 * WPB, 0, s>>>(*cfg, *b, q2, old_stats, heat_cell, n, sg, reward_out, done_out);
  if (cfg->flags & PCGRL_FLAG_AUTO_RESET) solver_launch<PROB>(cfg, b->stats, b->start_stats, b->map, q2, b->scratch, n, s, max_slots);
  return cuda_rc(cudaGetLastError(), "pcgrl_step launch");
}

// pcgrl_buffers of the env range [off, off + m) with its own scratch region
static pcgrl_buffers shard_buffers(const pcgrl_config* cfg, const pcgrl_buffers* b, int off, void* scratch, size_t scratch_bytes) {
  const size_t cells = (size_t)cfg->width * cfg->height, o = (size_t)off;
  pcgrl_buffers g = *b;
  g.map += o * cells; g.heatmap = (char*)g.heatmap + o * cells * (size_t)heat_bytes(*cfg); g.pos += 2 * o; g.iteration += o; g.changes += o;
  g.stats += o * PCGRL_MAX_STATS; g.start_stats += o * PCGRL_MAX_STATS; g.info_stats += o * PCGRL_MAX_STATS;
  g.reward += o; g.done += o; g.rng += o * 2 * PCGRL_MT_WORDS; g.tile_prob += o * PCGRL_MAX_TILES;
  g.start_map += o * cells; g.start_valid += o;
  g.scratch = scratch; g.scratch_bytes = scratch_bytes;
  return g;
}

struct GroupStreams {
  int device = -1;
  cudaStream_t streams[SOLVER_MAX_GROUPS];
  cudaEvent_t fork, join[SOLVER_MAX_GROUPS];
};

static GroupStreams* group_streams() {
  static thread_local GroupStreams gs;
  int dev = 0;
  cudaGetDevice(&dev);
  if (gs.device != dev) {  // (re)create for the calling thread's current device
    for (int i = 0; i < SOLVER_MAX_GROUPS; i++) {
      cudaStreamCreateWithFlags(&gs.streams[i], cudaStreamNonBlocking);
      cudaEventCreateWithFlags(&gs.join[i], cudaEventDisableTiming);
    }
    cudaEventCreateWithFlags(&gs.fork, cudaEventDisableTiming);
    gs.device = dev;
  }
  return &gs;
}

// T >= 1 steps of a solver problem through the env-asynchronous persistent kernel (k_rollout_async).
template <int PROB>
static int rollout_async(const pcgrl_config* cfg, const pcgrl_buffers* b, const int32_t* actions, double* reward_out,
                         uint8_t* done_out, int T, int n, cudaStream_t s, Staging sg) {
  if constexpr (GameOf<PROB>::GAME >= 0) {
    int table_size, heap_fast;
    const size_t smem = ((solver_async_arena_words(cfg, &table_size, &heap_fast) + 3) & ~(size_t)3) * sizeof(uint32_t) + ASYNC_RESET_AREAS * sizeof(WarpSmem);
    // launch geometry is a function of (device, problem, smem): queried once, not on every per-step call
    struct Geometry { size_t smem; int sm_count, per_sm; };
    static thread_local Geometry geo[SOLVER_MAX_DEVICES][PCGRL_NUM_PROBLEMS] = {};
    int dev = 0;
    cudaGetDevice(&dev);  // thread-local read in the runtime, no driver round trip
    Geometry local = {0, 0, 0};
    Geometry& g = (dev >= 0 && dev < SOLVER_MAX_DEVICES) ? geo[dev][PROB] : local;
    if (g.smem != smem || g.per_sm < 1) {
      cudaError_t ce = cudaFuncSetAttribute(k_rollout_async<PROB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (ce != cudaSuccess) return cuda_rc(ce, "k_rollout_async shared memory opt-in");
      cudaDeviceGetAttribute(&g.sm_count, cudaDevAttrMultiProcessorCount, dev);
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g.per_sm, k_rollout_async<PROB>, 32 * ASYNC_WPB, smem);
      g.smem = smem;
    }
    const int sm_count = g.sm_count, per_sm = g.per_sm;
    if (per_sm < 1) return fail(-1, "k_rollout_async does not fit one SM (solver_power too large)");
    const SolverLayout lay = solver_layout(cfg, n);
    int grid = (n + ASYNC_WPB - 1) / ASYNC_WPB;
    if (grid > sm_count * per_sm) grid = sm_count * per_sm;
    if (grid > 4 * lay.slots) grid = 4 * lay.slots;  // one node pool per CTA
    if (grid > ASYNC_MAX_CTAS) grid = ASYNC_MAX_CTAS;
    char* region = (char*)b->scratch + lay.async_off;
    AsyncHeader* hdr = (AsyncHeader*)region;
    AsyncGroup* groups = (AsyncGroup*)(region + 256);
    cudaMemsetAsync(region, 0, 256 + sizeof(AsyncGroup) * (size_t)grid * ASYNC_WPB, s);
    uint32_t* pool = (uint32_t*)((char*)b->scratch + lay.nodes_off);
    uint32_t* heap_pool = (uint32_t*)((char*)b->scratch + lay.heap_off);
    k_rollout_async<PROB><<<grid, 32 * ASYNC_WPB, smem, s>>>(*cfg, *b, actions, reward_out, done_out, T, n, hdr, groups, pool,
                                                             lay.nodes_per_pass, table_size, sg, heap_pool, lay.heap_words, heap_fast);
    return cuda_rc(cudaGetLastError(), "pcgrl_rollout (async) launch");
  } else {
    return fail(-1, "not a solver problem");
  }
}

// T consecutive PcgrlEnv.step calls of a solver problem.  Default: the env-asynchronous kernel (any T, one launch).
// PCGRL_SOLVER_ASYNC=0 selects the older multi-launch pipeline kept for comparison: T == 1 is one lock-step batch
// (update -> k_solve -> finish -> k_solve); for T > 1 the batch is split into independent env groups, each advancing
// through its own T steps on its own stream, so that a capped search stalls one group, not the batch.
template <int PROB>
static int rollout_solver(const pcgrl_config* cfg, const pcgrl_buffers* b, const int32_t* actions, double* reward_out,
                          uint8_t* done_out, int T, int n, cudaStream_t s, Staging sg) {
  const int adim = action_dim_host(cfg->representation);
  // PCGRL_SOLVER_ASYNC=0 selects the older stream-group pipeline below (kept for comparison)
  static const bool use_async = !(getenv("PCGRL_SOLVER_ASYNC") && atoi(getenv("PCGRL_SOLVER_ASYNC")) == 0);
  if (use_async) return rollout_async<PROB>(cfg, b, actions, reward_out, done_out, T, n, s, sg);
  const GroupPlan plan = solver_group_plan(cfg, n);
  if (T == 1 || plan.groups == 1) {
    for (int t = 0; t < T; t++) {
      int rc = step_solver<PROB>(cfg, b, actions + (size_t)t * n * adim, n, s, (t == T - 1) ? sg : Staging{nullptr, 0u, 0u, 0, n},
                                 SOLVER_MAX_SLOTS, reward_out ? reward_out + (size_t)t * n : nullptr,
                                 done_out ? done_out + (size_t)t * n : nullptr);
      if (rc) return rc;
    }
    return cuda_rc(cudaGetLastError(), "pcgrl_rollout");
  }
  GroupStreams* gs = group_streams();
  solver_prepare<PROB>(cfg);
  cudaEventRecord(gs->fork, s);
  for (int g = 0; g < plan.groups; g++) cudaStreamWaitEvent(gs->streams[g], gs->fork, 0);
  // the enqueue itself (groups x T x 5 launches) is spread over a few host threads, one set of groups each
  int dev = 0;
  cudaGetDevice(&dev);
  int nthreads = 1;  // PCGRL_ENQUEUE_THREADS: host threads sharing the enqueue (default 1: the device front end is the limit)
  if (const char* env = getenv("PCGRL_ENQUEUE_THREADS")) { const int v = atoi(env); if (v >= 1 && v <= 16) nthreads = v; }
  if (nthreads > plan.groups) nthreads = plan.groups;
  std::vector<int> rcs(nthreads, 0);
  auto enqueue = [&](int tid) {
    cudaSetDevice(dev);
    for (int g = tid; g < plan.groups; g += nthreads) {
      const int off = g * plan.envs_per_group, m = (off + plan.envs_per_group <= n) ? plan.envs_per_group : n - off;
      if (m <= 0) continue;
      const pcgrl_buffers bg = shard_buffers(cfg, b, off, (char*)b->scratch + (size_t)g * plan.bytes_per_group, plan.bytes_per_group);
      for (int t = 0; t < T && rcs[tid] == 0; t++)
        rcs[tid] = step_solver<PROB>(cfg, &bg, actions + ((size_t)t * n + off) * adim, m, gs->streams[g], Staging{nullptr, 0u, 0u, 0, m},
                                     plan.slots_per_group, reward_out ? reward_out + (size_t)t * n + off : nullptr,
                                     done_out ? done_out + (size_t)t * n + off : nullptr);
    }
  };
  std::vector<std::thread> workers;
  for (int tid = 1; tid < nthreads; tid++) workers.emplace_back(enqueue, tid);
  enqueue(0);
  for (auto& w : workers) w.join();
  for (int g = 0; g < plan.groups; g++) {
    cudaEventRecord(gs->join[g], gs->streams[g]);
    cudaStreamWaitEvent(s, gs->join[g], 0);
  }
  for (int tid = 0; tid < nthreads; tid++)
    if (rcs[tid]) return fail(rcs[tid], "pcgrl_rollout: a group launch failed");
  return cuda_rc(cudaGetLastError(), "pcgrl_rollout");
}

static int rollout_dispatch(const pcgrl_config* cfg, const pcgrl_buffers* b, const int32_t* actions, double* reward_out,
                            uint8_t* done_out, int T, int n, void* stream, Staging sg) {
  int rc = check_common(cfg, b, n);
  if (rc) return rc;
  if (!actions) return fail(-1, "actions is NULL");
  if (T <= 0) return fail(-1, "T must be > 0");
  if ((unsigned long long)T * (unsigned long long)n * (unsigned long long)action_dim_host(cfg->representation) >= (1ull << 31))
    return fail(-1, "T * n * action_dim must be < 2^31 per call: split the rollout");
  cudaStream_t s = (cudaStream_t)stream;
  switch (cfg->problem) {
    case PCGRL_PROB_BINARY: return rollout_fused<PCGRL_PROB_BINARY>(cfg, b, actions, reward_out, done_out, T, n, s, sg);
    case PCGRL_PROB_ZELDA: return rollout_fused<PCGRL_PROB_ZELDA>(cfg, b, actions, reward_out, done_out, T, n, s, sg);
    case PCGRL_PROB_SOKOBAN: return rollout_solver<PCGRL_PROB_SOKOBAN>(cfg, b, actions, reward_out, done_out, T, n, s, sg);
    case PCGRL_PROB_DDAVE: return rollout_solver<PCGRL_PROB_DDAVE>(cfg, b, actions, reward_out, done_out, T, n, s, sg);
    case PCGRL_PROB_SMB: return smb_rollout(cfg, b, actions, reward_out, done_out, T, n, s, sg);
    default: return rollout_solver<PCGRL_PROB_MDUNGEON>(cfg, b, actions, reward_out, done_out, T, n, s, sg);
  }
}

extern "C" int pcgrl_rollout(const pcgrl_config* cfg, const pcgrl_buffers* b, const int32_t* actions, double* reward_out,
                             uint8_t* done_out, int T, int n, void* stream) {
  return rollout_dispatch(cfg, b, actions, reward_out, done_out, T, n, stream, Staging{nullptr, 0u, 0u, 0, n});
}

extern "C" int pcgrl_step(const pcgrl_config* cfg, const pcgrl_buffers* b, const int32_t* actions, int n, void* stream) {
  return rollout_dispatch(cfg, b, actions, nullptr, nullptr, 1, n, stream, Staging{nullptr, 0u, 0u, 0, n});
}

template <int PROB>
static int get_stats_impl(const pcgrl_config* cfg, const uint8_t* maps, int32_t* out, int n, void* scratch,
                          int32_t* status, cudaStream_t s) {
  SolverQueue q = solver_queue(cfg, scratch, n, 0, status);
  if constexpr (ProblemTraits<PROB>::SOLVER) solver_queue_clear(q, s);
  k_get_stats<PROB><<<env_grid(n), 32 * WPB, 0, s>>>(*cfg, maps, out, q, n);
  if constexpr (ProblemTraits<PROB>::SOLVER) solver_launch<PROB>(cfg, out, nullptr, maps, q, scratch, n, s);
  return cuda_rc(cudaGetLastError(), "pcgrl_get_stats launch");
}

extern "C" int pcgrl_get_stats(const pcgrl_config* cfg, const uint8_t* maps, int32_t* stats_out, int n, void* scratch,
                               size_t scratch_bytes, int32_t* status, void* stream) {
  if (!cfg || !maps || !stats_out || !status) return fail(-1, "NULL argument");
  if (n <= 0) return fail(-1, "n must be > 0");
  int rc = pcgrl_config_validate(cfg);
  if (rc) return rc;
  if (cfg->problem >= PCGRL_PROB_SOKOBAN && (!scratch || scratch_bytes < pcgrl_scratch_bytes(cfg, n)))
    return fail(-1, "scratch buffer too small: see pcgrl_scratch_bytes()");
  cudaStream_t s = (cudaStream_t)stream;
  switch (cfg->problem) {
    case PCGRL_PROB_BINARY: return get_stats_impl<PCGRL_PROB_BINARY>(cfg, maps, stats_out, n, scratch, status, s);
    case PCGRL_PROB_ZELDA: return get_stats_impl<PCGRL_PROB_ZELDA>(cfg, maps, stats_out, n, scratch, status, s);
    case PCGRL_PROB_SOKOBAN: return get_stats_impl<PCGRL_PROB_SOKOBAN>(cfg, maps, stats_out, n, scratch, status, s);
    case PCGRL_PROB_DDAVE: return get_stats_impl<PCGRL_PROB_DDAVE>(cfg, maps, stats_out, n, scratch, status, s);
    case PCGRL_PROB_SMB: return smb_get_stats(cfg, maps, stats_out, n, scratch, s);
    default: return get_stats_impl<PCGRL_PROB_MDUNGEON>(cfg, maps, stats_out, n, scratch, status, s);
  }
}

extern "C" int pcgrl_seed(const pcgrl_buffers* b, const uint32_t* seeds, int n, void* stream) {
  if (!b || !b->rng || !seeds) return fail(-1, "NULL argument");
  if (n <= 0) return fail(-1, "n must be > 0");
  k_seed<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(b->rng, seeds, n);
  return cuda_rc(cudaGetLastError(), "pcgrl_seed launch");
}

#ifdef PCGRL_PROFILE
#include <chrono>
static double g_host_t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
extern "C" void pcgrl_debug_timers(double* out) { for (int i = 0; i < 8; i++) { out[i] = g_host_t[i]; g_host_t[i] = 0; } }
#define HT(i) do { auto now_ = std::chrono::high_resolution_clock::now(); g_host_t[i] += std::chrono::duration<double, std::micro>(now_ - ht_).count(); ht_ = now_; } while (0)
#define HT_BEGIN() auto ht_ = std::chrono::high_resolution_clock::now(); g_host_t[7] += 1
#else
#define HT(i) do {} while (0)
#define HT_BEGIN() do {} while (0)
#endif

static int staging_slots(int n) {
  int r = n / 64;
  return r < 16 ? 16 : (r > 255 ? 255 : r);
}

extern "C" size_t pcgrl_host_staging_bytes(const pcgrl_config* cfg, int n) {
  if (!cfg || n <= 0) return 0;
  return staging_layout(n, staging_slots(n), cfg->width * cfg->height).total;
}

// Device-visible address of a pinned host allocation (NULL if p is NULL or not device-mapped).  The lookups of the last
// few base pointers are cached: a binding passes the same arrays on every step.
// refresh = true re-queries the driver and replaces the cached entry: done on every (re)synchronising call of an io block
// (io->synced == 0), which is also the call a binding must make after it swapped its host arrays.
static void* host_device_ptr(const void* p, bool refresh = false) {
  if (!p) return nullptr;
  struct Entry { const void* host; void* dev; };
  static thread_local Entry cache[16];
  static thread_local int used = 0;
  int slot = -1;
  for (int i = 0; i < used; i++) if (cache[i].host == p) { slot = i; break; }
  if (slot >= 0 && !refresh) return cache[slot].dev;
  cudaPointerAttributes attr;
  void* dev = nullptr;
  if (cudaPointerGetAttributes(&attr, p) == cudaSuccess && attr.type == cudaMemoryTypeHost && attr.devicePointer) dev = attr.devicePointer;
  else cudaGetLastError();
  if (slot >= 0) cache[slot].dev = dev;
  else if (dev && used < 16) cache[used++] = Entry{p, dev};
  return dev;
}

// pcgrl_step_host = pcgrl_step_host_begin (enqueue: actions in, step kernels, results towards pinned host memory) +
// pcgrl_step_host_end (wait for / poll the stream, then patch the host arrays).  io->pending carries the state between
// the two: 0 idle, 1 delta transport in flight, 2 full copies in flight.
extern "C" int pcgrl_step_host_begin(const pcgrl_config* cfg, const pcgrl_buffers* b, int32_t* d_actions, pcgrl_host_io* io,
                                     int n, void* stream) {
  if (!io || !io->actions || !io->reward || !io->done || !d_actions) return fail(-1, "NULL host io pointer");
  if (io->pending) return fail(-1, "pcgrl_step_host_begin: the previous step of this io block has not been ended");
  int rc = check_common(cfg, b, n);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  const size_t cells = (size_t)cfg->width * cfg->height;
  const int adim = action_dim_host(cfg->representation);
  const bool wide = cfg->representation == PCGRL_REP_WIDE;
  const bool delta = io->mode == 1;
  const size_t hb = (size_t)heat_bytes(*cfg);
  if (delta && (!io->d_staging || !io->h_staging || io->staging_bytes < pcgrl_host_staging_bytes(cfg, n)))
    return fail(-1, "mode 1 needs d_staging / h_staging of pcgrl_host_staging_bytes() bytes");
  HT_BEGIN();
  // Pinned, device-mapped host actions are read by the step kernel directly (no cudaMemcpyAsync call: -6 us of host
  // time for +3 us of kernel time on PCIe Gen5); pageable buffers, or PCGRL_ZERO_COPY_ACTIONS=0, take the H2D copy.
  static const bool zero_copy = !(getenv("PCGRL_ZERO_COPY_ACTIONS") && atoi(getenv("PCGRL_ZERO_COPY_ACTIONS")) == 0);
  const int32_t* act_ptr = d_actions;
  if (zero_copy) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, io->actions) == cudaSuccess && attr.type == cudaMemoryTypeHost && attr.devicePointer)
      act_ptr = (const int32_t*)attr.devicePointer;
  }
  if (act_ptr == d_actions) cudaMemcpyAsync(d_actions, io->actions, sizeof(int32_t) * (size_t)n * adim, cudaMemcpyHostToDevice, s);
  HT(0);

  if (io->mode == 2 && io->synced) {
    // direct transport: the kernels store the results straight into the caller's pinned host arrays
    Staging sg{nullptr, 0u, 0u, 0, n};
    sg.direct = 1;
    sg.h_map = (uint8_t*)host_device_ptr(io->map);
    sg.h_heat = (uint8_t*)host_device_ptr(io->heatmap);
    sg.h_pos = wide ? nullptr : (uint8_t*)host_device_ptr(io->pos);
    sg.h_reward = (double*)host_device_ptr(io->reward);
    sg.h_done = (uint8_t*)host_device_ptr(io->done);
    sg.d_heat = b->heatmap;
    if (!sg.h_reward || !sg.h_done || (io->map && !sg.h_map) || (io->heatmap && !sg.h_heat) || (io->pos && !wide && !sg.h_pos))
      return fail(-1, "mode 2 needs pinned (device-mapped) host arrays");
    rc = rollout_dispatch(cfg, b, act_ptr, nullptr, nullptr, 1, n, stream, sg);
    if (rc) return rc;
    HT(1);
    if (io->info_stats) cudaMemcpyAsync(io->info_stats, b->info_stats, sizeof(int32_t) * PCGRL_MAX_STATS * n, cudaMemcpyDeviceToHost, s);
    HT(2);
    io->pending = 3;
    return 0;
  }

  if (delta && io->synced) {
    if (n >= (1 << 24)) return fail(-1, "delta transport supports n < 2^24 envs per call");
    const int nslots = staging_slots(n);
    const StagingLayout L = staging_layout(n, nslots, (int)cells);
    Staging sg{(uint8_t*)io->d_staging, (uint32_t)io->reset_base, (uint32_t)io->change_base, nslots, n};
    rc = rollout_dispatch(cfg, b, act_ptr, nullptr, nullptr, 1, n, stream, sg);
    if (rc) return rc;
    HT(1);
    cudaMemcpyAsync(io->h_staging, io->d_staging, L.total, cudaMemcpyDeviceToHost, s);
    if (io->info_stats) cudaMemcpyAsync(io->info_stats, b->info_stats, sizeof(int32_t) * PCGRL_MAX_STATS * n, cudaMemcpyDeviceToHost, s);
    HT(2);
    io->pending = 1;
    return 0;
  }

  // full copies (mode 0, or the first / re-arming call of modes 1 and 2)
  if (delta) cudaMemsetAsync(io->d_staging, 0, PCGRL_STAGING_HEADER, s);
  if (io->mode == 2) {  // (re)resolve the device-visible addresses of the host arrays for the direct steps that follow
    host_device_ptr(io->map, true); host_device_ptr(io->heatmap, true); host_device_ptr(io->pos, true);
    host_device_ptr(io->reward, true); host_device_ptr(io->done, true);
  }
  rc = rollout_dispatch(cfg, b, act_ptr, nullptr, nullptr, 1, n, stream, Staging{nullptr, 0u, 0u, 0, n});
  if (rc) return rc;
  if (io->map) cudaMemcpyAsync(io->map, b->map, cells * n, cudaMemcpyDeviceToHost, s);
  if (io->heatmap) cudaMemcpyAsync(io->heatmap, b->heatmap, hb * cells * n, cudaMemcpyDeviceToHost, s);
  if (io->pos && !wide) cudaMemcpyAsync(io->pos, b->pos, 2 * (size_t)n, cudaMemcpyDeviceToHost, s);
  cudaMemcpyAsync(io->reward, b->reward, sizeof(double) * n, cudaMemcpyDeviceToHost, s);
  cudaMemcpyAsync(io->done, b->done, (size_t)n, cudaMemcpyDeviceToHost, s);
  if (io->info_stats) cudaMemcpyAsync(io->info_stats, b->info_stats, sizeof(int32_t) * PCGRL_MAX_STATS * n, cudaMemcpyDeviceToHost, s);
  io->pending = 2;
  return 0;
}

extern "C" int pcgrl_step_host_end(const pcgrl_config* cfg, const pcgrl_buffers* b, pcgrl_host_io* io, int n, void* stream,
                                   int wait) {
  if (!cfg || !b || !io) return fail(-1, "NULL argument");
  if (!io->pending) return fail(-1, "pcgrl_step_host_end: no step in flight on this io block");
  cudaStream_t s = (cudaStream_t)stream;
  if (!wait) {
    const cudaError_t q = cudaStreamQuery(s);
    if (q == cudaErrorNotReady) return 1;  // still running: call again (nothing was consumed)
    if (q != cudaSuccess) return cuda_rc(q, "pcgrl_step_host_end");
  }
  HT_BEGIN();
  int rc = cuda_rc(cudaStreamSynchronize(s), "pcgrl_step_host");
  if (rc) return rc;
  HT(3);
  const size_t cells = (size_t)cfg->width * cfg->height;
  const bool wide = cfg->representation == PCGRL_REP_WIDE;
  const size_t hb = (size_t)heat_bytes(*cfg);
  uint8_t* const heat8 = (uint8_t*)io->heatmap;
  uint16_t* const heat16 = (uint16_t*)io->heatmap;
  if (io->pending == 2) {
    io->pending = 0;
    if (io->mode == 1 || io->mode == 2) { io->synced = 1; io->reset_base = 0; io->change_base = 0; }
    return 0;
  }
  if (io->pending == 3) {  // direct transport: the kernel wrote the host arrays itself
    io->pending = 0;
    return 0;
  }
  io->pending = 0;
  const StagingLayout L = staging_layout(n, staging_slots(n), (int)cells);
  // reward / done / cursor arrive in their final layout; the observation arrays are patched from the change list
  const uint8_t* hs = (const uint8_t*)io->h_staging;
  const uint32_t total_resets = ((const uint32_t*)hs)[0], total_changes = ((const uint32_t*)hs)[1];
  memcpy(io->reward, hs + L.reward_off, sizeof(double) * (size_t)n);
  memcpy(io->done, hs + L.done_off, (size_t)n);
  const uint8_t* hpos = hs + L.pos_off;
  if (io->pos && !wide) memcpy(io->pos, hpos, 2 * (size_t)n);
  const ChangeRecord* rec = (const ChangeRecord*)(hs + L.rec_off);
  const uint8_t* slots = hs + L.slot_off;
  uint32_t nchg = total_changes - (uint32_t)io->change_base;
  if (nchg > (uint32_t)n) nchg = (uint32_t)n;
  bool overflow = false;
  const uint32_t PF = 8;  // the records scatter over the 2 x n x H x W host arrays: fetch the target lines ahead
  for (uint32_t k = 0; k < nchg; k++) {
    if (k + PF < nchg) {
      const ChangeRecord& f = rec[k + PF];
      const size_t fe = f.env_kind & 0xffffffu;
      if (fe < (size_t)n) {
        if (io->map) __builtin_prefetch(io->map + fe * cells + f.cell, 1);
        if (io->heatmap) __builtin_prefetch(heat8 + hb * (fe * cells + (wide ? (size_t)f.cell : (size_t)hpos[2 * fe + 1] * cfg->width + hpos[2 * fe])), 1);
      }
    }
    const ChangeRecord r = rec[k];
    const size_t e = r.env_kind & 0xffffffu;
    const uint32_t kind = r.env_kind >> 24;
    if (e >= (size_t)n) continue;
    if (kind == PCGRL_REC_RESET) {
      if (r.slot == 0xFF) overflow = true;
      else if (io->map) memcpy(io->map + e * cells, slots + (size_t)r.slot * cells, cells);
      if (io->heatmap) memset(heat8 + hb * e * cells, 0, hb * cells);
    } else {
      if (kind == PCGRL_REC_MULTI) {
        if (r.slot == 0xFF) overflow = true;
        else if (io->map) memcpy(io->map + e * cells, slots + (size_t)r.slot * cells, cells);
      } else if (io->map) {
        io->map[e * cells + r.cell] = r.tile;
      }
      if (io->heatmap) {
        const size_t hi = e * cells + (wide ? (size_t)r.cell : (size_t)hpos[2 * e + 1] * cfg->width + hpos[2 * e]);
        if (hb == 2) heat16[hi] += 1; else heat8[hi] += 1;
      }
    }
  }
  io->reset_base = (int64_t)total_resets;
  io->change_base = (int64_t)total_changes;
  HT(4);
  if (overflow && io->map) {  // more whole-map updates than staging slots in one step: fetch the whole map batch
    cudaMemcpyAsync(io->map, b->map, cells * n, cudaMemcpyDeviceToHost, s);
    rc = cuda_rc(cudaStreamSynchronize(s), "pcgrl_step_host (map refetch)");
  }
  return rc;
}

extern "C" int pcgrl_step_host(const pcgrl_config* cfg, const pcgrl_buffers* b, int32_t* d_actions, pcgrl_host_io* io,
                               int n, void* stream) {
  const int rc = pcgrl_step_host_begin(cfg, b, d_actions, io, n, stream);
  if (rc) return rc;
  return pcgrl_step_host_end(cfg, b, io, n, stream, 1);
}

extern "C" int pcgrl_rollout_host(const pcgrl_config* cfg, const pcgrl_buffers* b, int32_t* d_actions, double* d_reward,
                                  uint8_t* d_done, pcgrl_host_rollout_io* io, int T, int n, void* stream) {
  if (!io || !io->actions || !io->reward || !io->done || !d_actions || !d_reward || !d_done)
    return fail(-1, "NULL rollout io pointer");
  if (T <= 0) return fail(-1, "T must be > 0");
  int rc = check_common(cfg, b, n);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  const size_t cells = (size_t)cfg->width * cfg->height, tn = (size_t)T * n;
  const int adim = action_dim_host(cfg->representation);
  const bool wide = cfg->representation == PCGRL_REP_WIDE;
  // one bulk H2D copy of all T action rows (a T-step kernel that read pinned host memory directly would pay a PCIe
  // round trip per env-step; the per-step call does read them in place, see pcgrl_step_host)
  cudaMemcpyAsync(d_actions, io->actions, sizeof(int32_t) * tn * adim, cudaMemcpyHostToDevice, s);
  const int32_t* act_ptr = d_actions;
  // Final observation: when the host arrays are pinned (device-mapped), every warp stores its env's final map / heat map /
  // cursor into them as it finishes (posted PCIe writes that overlap the envs still running); otherwise D2H copies.
  // PCGRL_ROLLOUT_DIRECT=0 forces the copies (A/B).
  static const bool direct_ok = !(getenv("PCGRL_ROLLOUT_DIRECT") && atoi(getenv("PCGRL_ROLLOUT_DIRECT")) == 0);
  Staging sg{nullptr, 0u, 0u, 0, n};
  // (binary / zelda: the fused kernel; the solver problems' rollouts last milliseconds and keep the copies)
  if (direct_ok && (cfg->problem == PCGRL_PROB_BINARY || cfg->problem == PCGRL_PROB_ZELDA) && (io->map || io->heatmap)) {
    sg.h_map = (uint8_t*)host_device_ptr(io->map, true);
    sg.h_heat = (uint8_t*)host_device_ptr(io->heatmap, true);
    sg.h_pos = wide ? nullptr : (uint8_t*)host_device_ptr(io->pos, true);
    sg.d_heat = b->heatmap;
    const bool all_mapped = (!io->map || sg.h_map) && (!io->heatmap || sg.h_heat) && (!io->pos || wide || sg.h_pos);
    sg.direct = all_mapped ? 2 : 0;
  }
  rc = rollout_dispatch(cfg, b, act_ptr, d_reward, d_done, T, n, stream, sg);
  if (rc) return rc;
  cudaMemcpyAsync(io->reward, d_reward, sizeof(double) * tn, cudaMemcpyDeviceToHost, s);
  cudaMemcpyAsync(io->done, d_done, tn, cudaMemcpyDeviceToHost, s);
  if (sg.direct != 2) {
    if (io->map) cudaMemcpyAsync(io->map, b->map, cells * n, cudaMemcpyDeviceToHost, s);
    if (io->heatmap) cudaMemcpyAsync(io->heatmap, b->heatmap, (size_t)heat_bytes(*cfg) * cells * n, cudaMemcpyDeviceToHost, s);
    if (io->pos && !wide) cudaMemcpyAsync(io->pos, b->pos, 2 * (size_t)n, cudaMemcpyDeviceToHost, s);
  }
  if (io->info_stats) cudaMemcpyAsync(io->info_stats, b->info_stats, sizeof(int32_t) * PCGRL_MAX_STATS * n, cudaMemcpyDeviceToHost, s);
  return cuda_rc(cudaStreamSynchronize(s), "pcgrl_rollout_host");
}

static void smb_operator_config(pcgrl_config* c, int width, int height, int solver_power) {
  memset(c, 0, sizeof(*c));
  c->problem = PCGRL_PROB_SMB; c->representation = PCGRL_REP_WIDE; c->width = width; c->height = height;
  c->num_tiles = pcgrl_smb::NUM_TILES; c->max_changes = 1; c->max_iterations = 1; c->solver_power = solver_power;
  for (int t = 0; t < pcgrl_smb::NUM_TILES; t++) c->tile_prob[t] = 1.0;
}

extern "C" size_t pcgrl_smb_scratch_bytes(int n, int solver_power) {
  if (n <= 0 || solver_power < 1 || solver_power > pcgrl_smb::MAX_POWER) return 0;
  return pcgrl_smb::scratch_bytes(n, solver_power);
}

extern "C" int pcgrl_smb_get_stats(const uint8_t* maps, int32_t* stats_out, int n, int width, int height, int solver_power,
                                   void* scratch, size_t scratch_bytes, void* stream) {
  if (!maps || !stats_out || !scratch) return fail(-1, "NULL argument");
  if (n <= 0) return fail(-1, "n must be > 0");
  pcgrl_config c;
  smb_operator_config(&c, width, height, solver_power);
  int rc = pcgrl_config_validate(&c);
  if (rc) return rc;
  if (scratch_bytes < pcgrl_smb_scratch_bytes(n, solver_power)) return fail(-1, "scratch buffer too small: see pcgrl_smb_scratch_bytes()");
  return smb_get_stats(&c, maps, stats_out, n, scratch, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------
// host twins: the same entry points on HOST pointers, no GPU involved, for every built-in problem: the
// step logic is scalar `__host__ __device__` code shared with the kernels (smb, the solver problems' game models) or the
// bitboard algorithm restated over row arrays (binary, zelda); only the search / BFS loops are host-specific.
// ------------------------------------------------------------------------------------------------
static int cpu_supported(const pcgrl_config* cfg) {
  if (!cfg) return fail(-1, "config is NULL");
  int rc = pcgrl_config_validate(cfg);
  if (rc) return rc;
  return 0;  // every built-in problem has a host twin (pcgrl_host_twin.cuh, pcgrl_solver_host.cuh, pcgrl_smb.cuh)
}
struct HostWorkOwner {
  pcgrl_smb::HostWork hw;
  explicit HostWorkOwner(const pcgrl_config* cfg) {
    memset(&hw, 0, sizeof(hw));
    hw.heap = (pcgrl_smb::u64*)malloc(sizeof(pcgrl_smb::u64) * pcgrl_smb::heap_entries(cfg->problem == PCGRL_PROB_SMB ? cfg->solver_power : 1));
  }
  ~HostWorkOwner() { free(hw.heap); }
};
// per-env "touched" bitmap: smb keeps it in the scratch buffer, the other problems have none
static uint32_t* host_touched(const pcgrl_config* cfg, const pcgrl_buffers* b, pcgrl_smb::HostWork& hw, int n, int e) {
  if (cfg->problem != PCGRL_PROB_SMB) return hw.no_touch;
  return pcgrl_smb::scratch_view(b->scratch, n, cfg->solver_power).touched + (size_t)e * pcgrl_smb::LEVEL_WORDS;
}
static int host_buffers_ok(const pcgrl_config* cfg, const pcgrl_buffers* b, int n) {
  if (!b || n <= 0) return fail(-1, "bad buffers");
  const size_t need = pcgrl_scratch_bytes(cfg, n);
  if (need && (!b->scratch || b->scratch_bytes < need)) return fail(-1, "scratch buffer too small: see pcgrl_scratch_bytes()");
  return 0;
}

extern "C" int pcgrl_reset_cpu(const pcgrl_config* cfg, const pcgrl_buffers* b, const uint8_t* mask, int n) {
  int rc = cpu_supported(cfg);
  if (rc) return rc;
  rc = host_buffers_ok(cfg, b, n);
  if (rc) return rc;
  HostWorkOwner w(cfg);
  if (!w.hw.heap) return fail(-1, "out of memory");
  for (int e = 0; e < n; e++) {
    if (mask && !mask[e]) continue;
    pcgrl_smb::host_reset_env(cfg, b, e, w.hw, host_touched(cfg, b, w.hw, n, e));
    b->reward[e] = 0.0;
    b->done[e] = 0;
    for (int i = 0; i < PCGRL_MAX_STATS; i++) b->info_stats[(size_t)e * PCGRL_MAX_STATS + i] = (i < 8) ? b->stats[(size_t)e * PCGRL_MAX_STATS + i] : 0;
  }
  return 0;
}

extern "C" int pcgrl_step_cpu(const pcgrl_config* cfg, const pcgrl_buffers* b, const int32_t* actions, int n) {
  int rc = cpu_supported(cfg);
  if (rc) return rc;
  if (!actions) return fail(-1, "actions is NULL");
  rc = host_buffers_ok(cfg, b, n);
  if (rc) return rc;
  HostWorkOwner w(cfg);
  if (!w.hw.heap) return fail(-1, "out of memory");
  for (int e = 0; e < n; e++) pcgrl_smb::host_step_env(cfg, b, actions, e, w.hw, host_touched(cfg, b, w.hw, n, e));
  return 0;
}

extern "C" int pcgrl_get_stats_cpu(const pcgrl_config* cfg, const uint8_t* maps, int32_t* stats_out, int n) {
  int rc = cpu_supported(cfg);
  if (rc) return rc;
  if (!maps || !stats_out || n <= 0) return fail(-1, "NULL argument");
  HostWorkOwner w(cfg);
  if (!w.hw.heap) return fail(-1, "out of memory");
  uint32_t touched[pcgrl_smb::LEVEL_WORDS];
  const size_t cells = (size_t)cfg->width * cfg->height;
  for (int e = 0; e < n; e++) pcgrl_smb::host_get_stats(cfg, maps + (size_t)e * cells, touched, w.hw, stats_out + (size_t)e * PCGRL_MAX_STATS);
  return 0;
}

extern "C" int pcgrl_obs_image(const pcgrl_config* cfg, const uint8_t* maps, const uint8_t* pos, void* out, int n,
                               int crop_size, int pad_value, int one_hot, int out_dtype, void* stream) {
  if (!cfg || !maps || !out) return fail(-1, "NULL argument");
  if (n <= 0) return fail(-1, "n must be > 0");
  if (crop_size < 0 || crop_size > 64) return fail(-1, "crop_size must be in [0, 64]");
  if (crop_size > 0 && !pos) return fail(-1, "Cropped needs the cursor positions");
  if (out_dtype != 0 && out_dtype != 1) return fail(-1, "out_dtype must be 0 (uint8) or 1 (float32)");
  int rc = pcgrl_config_validate(cfg);
  if (rc) return rc;
  const int S_h = crop_size ? crop_size : cfg->height, S_w = crop_size ? crop_size : cfg->width;
  const int channels = one_hot ? cfg->num_tiles : 1;
  const size_t total = (size_t)n * S_h * S_w * channels;
  if (total >= (1ull << 32)) return fail(-1, "observation tensor too large (>= 2^32 elements): split the batch");
  cudaStream_t s = (cudaStream_t)stream;
  if (out_dtype == 0) {
    // warp-per-env word builder for the two layouts the policies consume (raw index with S_w % 4 == 0, 8 one-hot channels)
    const int mode = (channels == 1 && (S_w & 3) == 0) ? 0 : (channels == 8 ? 1 : -1);
    if (mode >= 0 && crop_size <= 2 * OBS_FAST_SLACK && ((uintptr_t)maps & 3) == 0 && ((uintptr_t)out & 3) == 0) {
      const int D = mode == 0 ? (S_w >> 2) : S_w, range = mode == 0 ? S_h * D : S_h * S_w;
      const uint32_t magic = (65536u + (uint32_t)D - 1u) / (uint32_t)D;
      bool exact = true;
      for (int v = 0; v < range && exact; v++) exact = (int)(((uint32_t)v * magic) >> 16) == v / D;
      if (exact) {
        unsigned blocks = (unsigned)((n + OBS_FAST_WARPS - 1) / OBS_FAST_WARPS);
        if (blocks > 148u * 8u) blocks = 148u * 8u;
        const size_t map_bytes = (size_t)n * cfg->height * cfg->width;
        if (mode == 0)
          k_obs_image_u8_fast<0><<<blocks, 32 * OBS_FAST_WARPS, 0, s>>>(maps, pos, (uint32_t*)out, n, cfg->height, cfg->width, S_h, S_w,
                                                                       crop_size, pad_value, magic, map_bytes);
        else
          k_obs_image_u8_fast<1><<<blocks, 32 * OBS_FAST_WARPS, 0, s>>>(maps, pos, (uint32_t*)out, n, cfg->height, cfg->width, S_h, S_w,
                                                                       crop_size, pad_value, magic, map_bytes);
        return cuda_rc(cudaGetLastError(), "pcgrl_obs_image launch");
      }
    }
    const unsigned blocks = (unsigned)((total + OBS_THREADS * 64 - 1) / (OBS_THREADS * 64));  // one 16 KB tile per CTA
    k_obs_image<uint8_t><<<blocks, OBS_THREADS, 0, s>>>(maps, pos, (uint8_t*)out, (uint32_t)total, n, cfg->height, cfg->width,
                                                        S_h, S_w, crop_size, pad_value, channels);
  } else {
    const unsigned blocks = (unsigned)((total + OBS_THREADS * 16 - 1) / (OBS_THREADS * 16));
    k_obs_image<float><<<blocks, OBS_THREADS, 0, s>>>(maps, pos, (float*)out, (uint32_t)total, n, cfg->height, cfg->width,
                                                      S_h, S_w, crop_size, pad_value, channels);
  }
  return cuda_rc(cudaGetLastError(), "pcgrl_obs_image launch");
}

extern "C" int pcgrl_render(const uint8_t* maps, const uint8_t* pos_or_null, const uint8_t* atlas, uint8_t* out, int n,
                            int height, int width, int num_tiles, int border_w, int border_h, int border_tile, int tile_size,
                            void* stream) {
  if (!maps || !atlas || !out) return fail(-1, "NULL argument");
  if (n <= 0 || height <= 0 || width <= 0) return fail(-1, "n, height, width must be > 0");
  if (tile_size < 4 || (tile_size & 3)) return fail(-1, "tile_size must be a positive multiple of 4");
  if (num_tiles < 1 || border_tile < 0 || border_tile >= num_tiles || border_w < 0 || border_h < 0) return fail(-1, "bad tile / border arguments");
  const size_t quads = (size_t)n * (height + 2 * border_h) * tile_size * ((size_t)(width + 2 * border_w) * tile_size / 4);
  unsigned blocks = (unsigned)((quads + 255) / 256 < 148u * 16u ? (quads + 255) / 256 : 148u * 16u);
  k_render<<<blocks ? blocks : 1, 256, 0, (cudaStream_t)stream>>>(maps, pos_or_null, (const uint32_t*)atlas, (uint32_t*)out, n, height, width,
                                                                 border_w, border_h, border_tile, tile_size);
  return cuda_rc(cudaGetLastError(), "pcgrl_render launch");
}

extern "C" int pcgrl_action_map(const pcgrl_config* cfg, const pcgrl_buffers* b, const int32_t* flat_actions,
                                int32_t* actions_out, int n, void* stream) {
  if (!cfg || !b || !flat_actions || !actions_out) return fail(-1, "NULL argument");
  if (n <= 0) return fail(-1, "n must be > 0");
  int rc = pcgrl_config_validate(cfg);
  if (rc) return rc;
  if (cfg->representation != PCGRL_REP_WIDE && cfg->representation != PCGRL_REP_NARROW && cfg->representation != PCGRL_REP_TURTLE)
    return fail(-1, "ActionMap supports the narrow, turtle and wide representations");
  k_action_map<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(flat_actions, b->map, b->pos, actions_out, n, cfg->height,
                                                              cfg->width, cfg->num_tiles, cfg->representation == PCGRL_REP_WIDE);
  return cuda_rc(cudaGetLastError(), "pcgrl_action_map launch");
}

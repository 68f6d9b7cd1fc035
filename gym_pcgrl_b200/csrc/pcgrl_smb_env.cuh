// pcgrl_smb_env.cuh -- PcgrlEnv.reset / step / get_stats kernels of the smb problem and their host twins.
// Scalar building blocks, design notes and reference citations: pcgrl_smb.cuh.
#pragma once
#include <stdlib.h>

#include "pcgrl_env.cuh"
#include "pcgrl_host_twin.cuh"
#include "pcgrl_solver_host.cuh"
#include "pcgrl_smb.cuh"

namespace pcgrl_smb {

#define SMB_WPB 4             /* env warps per CTA */
#define SMB_HEADER_BYTES 256

// scratch: [header: work counter][touched bitmaps, LEVEL_WORDS words per env][slow heap, heap_entries(power) per slot]
struct Scratch {
  int32_t* work;
  uint32_t* touched;
  u64* heap;
  size_t heap_stride;
};
static inline int smb_ctas_cap() {  // resident CTAs per SM (tuning knob PCGRL_SMB_CTAS_PER_SM=1..10, read once)
  static const int cap = [] { const char* v = getenv("PCGRL_SMB_CTAS_PER_SM"); const int k = v ? atoi(v) : 0; return (k >= 1 && k <= 10) ? k : 5; }();
  return cap;
}
static inline int smb_slots(int n) {  // resident env warps of one launch (B200: 148 SMs)
  const int s = (n + SMB_WPB - 1) / SMB_WPB * SMB_WPB, most = 148 * smb_ctas_cap() * SMB_WPB;
  return s < most ? s : most;
}
static inline size_t scratch_bytes(int n, int power) {
  return SMB_HEADER_BYTES + sizeof(uint32_t) * LEVEL_WORDS * (size_t)n + sizeof(u64) * heap_entries(power) * (size_t)smb_slots(n);
}
static inline Scratch scratch_view(void* base, int n, int power) {
  Scratch s;
  s.work = (int32_t*)base;
  s.touched = (uint32_t*)((char*)base + SMB_HEADER_BYTES);
  s.heap = (u64*)((char*)base + SMB_HEADER_BYTES + sizeof(uint32_t) * LEVEL_WORDS * (size_t)n);
  s.heap_stride = heap_entries(power);
  return s;
}

#ifdef __CUDACC__
using pcgrl::Staging;

struct Arena {  // one warp's shared memory
  u64* heap_fast;
  uint32_t *visited, *solid, *touched;
  uint8_t* map;
};
__host__ __device__ __forceinline__ int map_stride(int cells) { return (cells + 15) & ~15; }
__host__ __device__ __forceinline__ int arena_fixed_bytes(int cells) { return 4 * (VISITED_WORDS + 2 * LEVEL_WORDS) + map_stride(cells); }
__device__ __forceinline__ Arena carve(unsigned char* base, int fast_cap, int cells) {
  Arena a;
  a.heap_fast = reinterpret_cast<u64*>(base);
  a.visited = reinterpret_cast<uint32_t*>(base + 8 * (size_t)fast_cap);
  a.solid = a.visited + VISITED_WORDS;
  a.touched = a.solid + LEVEL_WORDS;
  a.map = reinterpret_cast<uint8_t*>(a.touched + LEVEL_WORDS);
  return a;
}

__device__ __forceinline__ void warp_clear_visited(uint32_t* visited, int lane) {
  uint4* v = reinterpret_cast<uint4*>(visited);
#pragma unroll
  for (int k = 0; k < VISITED_WORDS / 4 / 32; k++) v[k * 32 + lane] = make_uint4(0u, 0u, 0u, 0u);
  static_assert(VISITED_WORDS % 128 == 0, "one uint4 per lane per round");
}

// _run_game (smb_prob.py:95-124) for the map staged in A.map: lane 0 writes st[5..7].  Rebuilds the touched bitmap.
__device__ __noinline__ void warp_search(const pcgrl_config& cfg, const Arena& A, const Heap& hp, int lane, int32_t* st) {
  const int W = cfg.width, H = cfg.height;
  __syncwarp();
  warp_clear_visited(A.visited, lane);
  A.touched[lane] = 0u;
  A.touched[lane + 32] = 0u;
  if (lane < H) build_solid_row(A.map, W, H, lane, A.solid);
  __syncwarp();
  Level L;
  L.width = W + 6; L.height = H; L.exit_x = W + 4; L.solid = A.solid; L.touched = A.touched;
  u64 node = 0;
  int won = 0, it = 0;
  if (lane == 0) won = astar_core(L, root_state(H), 1, cfg.solver_power, hp, A.visited, node, it) ? 1 : 0;
  won = __shfl_sync(0xffffffffu, won, 0);
  if (!won) {
    warp_clear_visited(A.visited, lane);
    __syncwarp();
    if (lane == 0) won = astar_core(L, root_state(H), 0, cfg.solver_power, hp, A.visited, node, it) ? 1 : 0;
  }
  if (lane == 0) play_stats(node, won != 0, W, L.exit_x, st);
  __syncwarp();
}

// PcgrlEnv.reset (pcgrl_env.py:66-76) for the env staged in this warp; lane 0 owns x, y, st.
__device__ __noinline__ void warp_reset(const pcgrl_config& cfg, const pcgrl_buffers& b, int e, const Arena& A, const Heap& hp,
                                        int lane, int& x, int& y, int32_t* st) {
  const int W = cfg.width, H = cfg.height, cells = W * H;
  uint8_t* gmap = b.map + (size_t)e * cells;
  uint8_t* smap = b.start_map + (size_t)e * cells;
  uint32_t* rng_rep = b.rng + (size_t)e * 2 * PCGRL_MT_WORDS;
  const bool generate = (cfg.flags & PCGRL_FLAG_RANDOM_START) || (b.start_valid[e] == 0);
  __syncwarp();
  if (!generate) {  // representation.py:44-45
    for (int i = lane; i < cells; i += 32) { const uint8_t t = smap[i]; A.map[i] = t; gmap[i] = t; }
  }
  __syncwarp();
  if (lane == 0) {
    if (generate) {  // representation.py:41-43
      gen_random_map(rng_rep, b.tile_prob + (size_t)e * PCGRL_MAX_TILES, cfg.num_tiles, cells, A.map, gmap, smap);
      b.start_valid[e] = 1;
    }
    if (cfg.representation != PCGRL_REP_WIDE) {  // narrow_rep.py:30-31, turtle_rep.py:32-33
      x = mt_randint(rng_rep, W);
      y = mt_randint(rng_rep, H);
    }
    scan_stats(A.map, W, H, st);
  }
  warp_search(cfg, A, hp, lane, st);
  pcgrl::warp_fill_bytes(reinterpret_cast<uint8_t*>(b.heatmap) + (size_t)e * cells * pcgrl::heat_bytes(cfg), cells * pcgrl::heat_bytes(cfg), 0, lane);  // pcgrl_env.py:72
  __syncwarp();
}

__device__ __forceinline__ void stage_env(const pcgrl_buffers& b, const Scratch& sc, int e, int cells, const Arena& A, int lane) {
  const uint8_t* gmap = b.map + (size_t)e * cells;
  if (((cells & 3) == 0) && ((((size_t)e * cells) & 3) == 0)) {
    const uint32_t* g4 = reinterpret_cast<const uint32_t*>(gmap);
    uint32_t* s4 = reinterpret_cast<uint32_t*>(A.map);
    for (int i = lane; i < (cells >> 2); i += 32) s4[i] = g4[i];
  } else {
    for (int i = lane; i < cells; i += 32) A.map[i] = gmap[i];
  }
  if (sc.touched) {
    A.touched[lane] = sc.touched[(size_t)e * LEVEL_WORDS + lane];
    A.touched[lane + 32] = sc.touched[(size_t)e * LEVEL_WORDS + lane + 32];
  }
  __syncwarp();
}

__device__ __forceinline__ int next_env(const Scratch& sc, int lane) {
  int e = 0;
  if (lane == 0) e = atomicAdd(sc.work, 1);
  return __shfl_sync(0xffffffffu, e, 0);
}

__device__ __forceinline__ double shfl_double(double v, int src) {
  const long long bits = __double_as_longlong(v);
  const int lo = __shfl_sync(0xffffffffu, (int)(bits & 0xffffffffll), src), hi = __shfl_sync(0xffffffffu, (int)(bits >> 32), src);
  return __longlong_as_double(((long long)hi << 32) | (long long)(uint32_t)lo);
}

// T consecutive PcgrlEnv.step calls per env (pcgrl_env.py:129-150), auto-reset inside.
__global__ void __launch_bounds__(32 * SMB_WPB) k_smb_rollout(const __grid_constant__ pcgrl_config cfg,
                                                              const __grid_constant__ pcgrl_buffers b,
                                                              const int32_t* __restrict__ actions, double* reward_out,
                                                              uint8_t* done_out, int T, int n, Scratch sc, Staging sg,
                                                              int fast_cap, int per_warp_bytes) {
  extern __shared__ __align__(16) unsigned char smb_dyn[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int W = cfg.width, H = cfg.height, cells = W * H, adim = pcgrl::action_dim(cfg.representation);
  const bool auto_reset = (cfg.flags & PCGRL_FLAG_AUTO_RESET) != 0;
  const Arena A = carve(smb_dyn + (size_t)wib * per_warp_bytes, fast_cap, cells);
  Heap hp;
  hp.fast = A.heap_fast;
  hp.slow = sc.heap + (size_t)(blockIdx.x * SMB_WPB + wib) * sc.heap_stride;
  hp.fast_cap = fast_cap;
  while (true) {
    const int e = next_env(sc, lane);
    if (e >= n) break;
    stage_env(b, sc, e, cells, A, lane);
    uint8_t* gmap = b.map + (size_t)e * cells;
    uint32_t* rng_rep = b.rng + (size_t)e * 2 * PCGRL_MT_WORDS;
    int x = 0, y = 0, iteration = 0, changes = 0;
    int32_t st[8], start[8];
    if (lane == 0) {
      if (cfg.representation != PCGRL_REP_WIDE) { x = b.pos[2 * e]; y = b.pos[2 * e + 1]; }
      iteration = b.iteration[e];
      changes = b.changes[e];
#pragma unroll
      for (int i = 0; i < 8; i++) { st[i] = b.stats[(size_t)e * PCGRL_MAX_STATS + i]; start[i] = b.start_stats[(size_t)e * PCGRL_MAX_STATS + i]; }
    }
    for (int t = 0; t < T; t++) {
      int change = 0, need_search = 0, done = 0, hx = 0, hy = 0, cell = 0, tile = 0, multi = 0;
      int32_t old[8];
      double reward = 0.0;
      if (lane == 0) {
        iteration++;  // pcgrl_env.py:130
#pragma unroll
        for (int i = 0; i < 8; i++) old[i] = st[i];
        const Edit ed = apply_action(cfg, actions + ((size_t)t * n + e) * adim, A.map, gmap, A.touched, rng_rep, x, y);
        change = ed.change; hx = ed.hx; hy = ed.hy; cell = ed.cell; tile = ed.tile; multi = ed.multi ? 1 : 0;
        if (change > 0) {  // pcgrl_env.py:135-138
          changes += change;
          scan_stats(A.map, W, H, st);
          need_search = ed.solidity_touched ? 1 : 0;  // otherwise jumps / jumps-dist / dist-win carry over (see pcgrl_smb.cuh)
        }
      }
      change = __shfl_sync(0xffffffffu, change, 0);
      need_search = __shfl_sync(0xffffffffu, need_search, 0);
      if (need_search) warp_search(cfg, A, hp, lane, st);
      if (lane == 0) {
        reward = (change > 0) ? get_reward(cfg, st, old) : 0.0;  // :142 (get_reward(s, s) == 0)
        done = (episode_over(st) || changes >= cfg.max_changes || iteration >= cfg.max_iterations) ? 1 : 0;  // :143
        if (reward_out) reward_out[(size_t)t * n + e] = reward;
        if (done_out) done_out[(size_t)t * n + e] = (uint8_t)done;
        if (t == T - 1) {
          b.reward[e] = reward;
          b.done[e] = (uint8_t)done;
          int32_t* info = b.info_stats + (size_t)e * PCGRL_MAX_STATS;
#pragma unroll
          for (int i = 0; i < 8; i++) info[i] = st[i];
          info[PCGRL_INFO_ITERATION] = iteration;
          info[PCGRL_INFO_CHANGES] = changes;
        }
      }
      done = __shfl_sync(0xffffffffu, done, 0);
      const bool resetting = done && auto_reset;
      if (resetting) {
        warp_reset(cfg, b, e, A, hp, lane, x, y, st);
        if (lane == 0) {
#pragma unroll
          for (int i = 0; i < 8; i++) start[i] = st[i];  // problem.py:45-46
          iteration = 0;
          changes = 0;
        }
      } else if (change > 0 && lane == 0) {  // :137; this warp owns the env, a plain read-modify-write does
        const size_t hi = (size_t)e * cells + (size_t)hy * W + hx;
        if (cfg.flags & PCGRL_FLAG_HEAT_U16) reinterpret_cast<uint16_t*>(b.heatmap)[hi] += 1;
        else reinterpret_cast<uint8_t*>(b.heatmap)[hi] += 1;
      }
      if (t == T - 1 && (sg.base || sg.direct)) {  // delta transport of pcgrl_step_host: warp-uniform arguments
        reward = shfl_double(reward, 0);
        const int rx = __shfl_sync(0xffffffffu, x, 0), ry = __shfl_sync(0xffffffffu, y, 0);
        cell = __shfl_sync(0xffffffffu, cell, 0);
        tile = __shfl_sync(0xffffffffu, tile, 0);
        multi = __shfl_sync(0xffffffffu, multi, 0);
        __threadfence_block();
        __syncwarp();
        pcgrl::write_record(sg, cfg, e, lane, reward, done != 0, rx, ry, change > 0, resetting, cell, tile, gmap, multi != 0);
      }
    }
    __syncwarp();
    if (lane == 0) {
      if (cfg.representation != PCGRL_REP_WIDE) { b.pos[2 * e] = (uint8_t)x; b.pos[2 * e + 1] = (uint8_t)y; }
      b.iteration[e] = iteration;
      b.changes[e] = changes;
#pragma unroll
      for (int i = 0; i < 8; i++) { b.stats[(size_t)e * PCGRL_MAX_STATS + i] = st[i]; b.start_stats[(size_t)e * PCGRL_MAX_STATS + i] = start[i]; }
    }
    sc.touched[(size_t)e * LEVEL_WORDS + lane] = A.touched[lane];
    sc.touched[(size_t)e * LEVEL_WORDS + lane + 32] = A.touched[lane + 32];
    __syncwarp();
  }
}

__global__ void __launch_bounds__(32 * SMB_WPB) k_smb_reset(const __grid_constant__ pcgrl_config cfg,
                                                            const __grid_constant__ pcgrl_buffers b,
                                                            const uint8_t* __restrict__ mask, int n, Scratch sc, int fast_cap,
                                                            int per_warp_bytes) {
  extern __shared__ __align__(16) unsigned char smb_dyn[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int cells = cfg.width * cfg.height;
  const Arena A = carve(smb_dyn + (size_t)wib * per_warp_bytes, fast_cap, cells);
  Heap hp;
  hp.fast = A.heap_fast;
  hp.slow = sc.heap + (size_t)(blockIdx.x * SMB_WPB + wib) * sc.heap_stride;
  hp.fast_cap = fast_cap;
  while (true) {
    const int e = next_env(sc, lane);
    if (e >= n) break;
    if (mask && mask[e] == 0) continue;
    int x = 0, y = 0;
    int32_t st[8];
    warp_reset(cfg, b, e, A, hp, lane, x, y, st);
    if (lane == 0) {
      if (cfg.representation != PCGRL_REP_WIDE) { b.pos[2 * e] = (uint8_t)x; b.pos[2 * e + 1] = (uint8_t)y; }
      b.iteration[e] = 0;
      b.changes[e] = 0;
      b.reward[e] = 0.0;
      b.done[e] = 0;
      int32_t* rows[3] = {b.stats + (size_t)e * PCGRL_MAX_STATS, b.start_stats + (size_t)e * PCGRL_MAX_STATS,
                          b.info_stats + (size_t)e * PCGRL_MAX_STATS};
      for (int r = 0; r < 3; r++)
        for (int i = 0; i < PCGRL_MAX_STATS; i++) rows[r][i] = (i < 8) ? st[i] : 0;
    }
    sc.touched[(size_t)e * LEVEL_WORDS + lane] = A.touched[lane];
    sc.touched[(size_t)e * LEVEL_WORDS + lane + 32] = A.touched[lane + 32];
    __syncwarp();
  }
}

// stand-alone SMBProblem.get_stats on n maps
__global__ void __launch_bounds__(32 * SMB_WPB) k_smb_get_stats(const __grid_constant__ pcgrl_config cfg,
                                                                const uint8_t* __restrict__ maps, int32_t* stats_out, int n,
                                                                Scratch sc, int fast_cap, int per_warp_bytes) {
  extern __shared__ __align__(16) unsigned char smb_dyn[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int W = cfg.width, H = cfg.height, cells = W * H;
  const Arena A = carve(smb_dyn + (size_t)wib * per_warp_bytes, fast_cap, cells);
  Heap hp;
  hp.fast = A.heap_fast;
  hp.slow = sc.heap + (size_t)(blockIdx.x * SMB_WPB + wib) * sc.heap_stride;
  hp.fast_cap = fast_cap;
  while (true) {
    const int e = next_env(sc, lane);
    if (e >= n) break;
    for (int i = lane; i < cells; i += 32) A.map[i] = maps[(size_t)e * cells + i];
    __syncwarp();
    int32_t st[8];
    if (lane == 0) scan_stats(A.map, W, H, st);
    warp_search(cfg, A, hp, lane, st);
    if (lane == 0)
      for (int i = 0; i < PCGRL_MAX_STATS; i++) stats_out[(size_t)e * PCGRL_MAX_STATS + i] = (i < 8) ? st[i] : 0;
    __syncwarp();
  }
}

// launch geometry: resident env warps per SM follow the batch, the shared memory left over goes to the heap's top levels
struct Launch { int grid, fast_cap, per_warp_bytes; size_t smem; };
static inline Launch launch_plan(const pcgrl_config* cfg, int n, int sm_count) {
  const int cells = cfg->width * cfg->height;
  const int slots = smb_slots(n);
  int ctas = slots / SMB_WPB;
  int ctas_per_sm = (ctas + sm_count - 1) / sm_count;
  if (ctas_per_sm < 1) ctas_per_sm = 1;
  // A search is one dependent instruction chain on lane 0 (~0.6 us per iteration): throughput comes from resident
  // warps hiding each other's latencies, not from a larger shared-memory slice per search.  Measured on B200
  // (tools/bench_smb.py, 8192 maps): 2 / 3 / 5 / 7 / 10 CTAs per SM -> 1.9 / 2.5 / 3.3 / 3.5 / 3.3 e5 maps/s.
  if (ctas_per_sm > smb_ctas_cap()) ctas_per_sm = smb_ctas_cap();
  const int budget = (227 * 1024 - 1024 * ctas_per_sm) / ctas_per_sm / SMB_WPB;  // bytes per warp (1 KB per CTA reserved by the driver)
  int fast_cap = (budget - arena_fixed_bytes(cells)) / 8;
  const int full = (int)heap_entries(cfg->solver_power);
  if (fast_cap > full) fast_cap = full;
  if (fast_cap > 5120) fast_cap = 5120;  // 40 KB: twelve heap levels; more shared memory per warp buys nothing
  fast_cap &= ~1;
  if (fast_cap < 64) fast_cap = 64;
  Launch L;
  L.fast_cap = fast_cap;
  L.per_warp_bytes = (8 * fast_cap + arena_fixed_bytes(cells) + 15) & ~15;
  L.smem = (size_t)L.per_warp_bytes * SMB_WPB;
  L.grid = ctas < sm_count * ctas_per_sm ? ctas : sm_count * ctas_per_sm;
  if (L.grid > slots / SMB_WPB) L.grid = slots / SMB_WPB;  // one slow-heap slice per resident warp
  return L;
}
#endif  // __CUDACC__

// ------------------------------------------------------------------------------------------------
// host twins (pcgrl_*_cpu): the same scalar functions, one env after the other; buffers are HOST pointers
// ------------------------------------------------------------------------------------------------
struct HostWork {
  uint32_t solid[LEVEL_WORDS], visited[VISITED_WORDS];
  uint32_t no_touch[32 * ROW_WORDS + 8];  // all-zero "touched" bitmap for the problems without search skipping
  u64* heap;
};

static inline int host_nstats(const pcgrl_config* cfg) {
  switch (cfg->problem) {
    case PCGRL_PROB_BINARY: return 2;
    case PCGRL_PROB_ZELDA: return 7;
    case PCGRL_PROB_SOKOBAN: return 6;
    case PCGRL_PROB_DDAVE: case PCGRL_PROB_MDUNGEON: return 11;
    default: return 8;
  }
}
// Problem.get_stats: smb = scans + A* play-through (this file), binary / zelda = host bitboards (pcgrl_host_twin.cuh)
static inline void host_get_stats(const pcgrl_config* cfg, const uint8_t* map, uint32_t* touched, HostWork& hw, int32_t* st) {
  for (int i = 0; i < PCGRL_MAX_STATS; i++) st[i] = 0;
  if (cfg->problem == PCGRL_PROB_SMB) {
    scan_stats(map, cfg->width, cfg->height, st);
    run_game_scalar(map, cfg->width, cfg->height, cfg->solver_power, hw.solid, touched, hw.visited, hw.heap, st, nullptr);
  } else if (!pcgrl_host::get_stats(cfg, map, st)) {  // binary / zelda; else a solver problem (pcgrl_solver_host.cuh)
    static thread_local pcgrl_host::SearchWork work;
    pcgrl_host::solver_get_stats(cfg, map, work, st);
  }
}
static inline double host_reward(const pcgrl_config* cfg, const int32_t* n, const int32_t* o) {
  if (cfg->problem == PCGRL_PROB_BINARY) return pcgrl::problem_reward<PCGRL_PROB_BINARY>(*cfg, n, o);
  if (cfg->problem == PCGRL_PROB_ZELDA) return pcgrl::problem_reward<PCGRL_PROB_ZELDA>(*cfg, n, o);
  if (cfg->problem == PCGRL_PROB_SOKOBAN) return pcgrl::problem_reward<PCGRL_PROB_SOKOBAN>(*cfg, n, o);
  if (cfg->problem == PCGRL_PROB_DDAVE) return pcgrl::problem_reward<PCGRL_PROB_DDAVE>(*cfg, n, o);
  if (cfg->problem == PCGRL_PROB_MDUNGEON) return pcgrl::problem_reward<PCGRL_PROB_MDUNGEON>(*cfg, n, o);
  return get_reward(*cfg, n, o);
}
static inline bool host_over(const pcgrl_config* cfg, const int32_t* n, const int32_t* start) {
  if (cfg->problem == PCGRL_PROB_BINARY) return pcgrl::problem_over<PCGRL_PROB_BINARY>(*cfg, n, start);
  if (cfg->problem == PCGRL_PROB_ZELDA) return pcgrl::problem_over<PCGRL_PROB_ZELDA>(*cfg, n, start);
  if (cfg->problem == PCGRL_PROB_SOKOBAN) return pcgrl::problem_over<PCGRL_PROB_SOKOBAN>(*cfg, n, start);
  if (cfg->problem == PCGRL_PROB_DDAVE) return pcgrl::problem_over<PCGRL_PROB_DDAVE>(*cfg, n, start);
  if (cfg->problem == PCGRL_PROB_MDUNGEON) return pcgrl::problem_over<PCGRL_PROB_MDUNGEON>(*cfg, n, start);
  return episode_over(n);
}

static inline void host_reset_env(const pcgrl_config* cfg, const pcgrl_buffers* b, int e, HostWork& hw, uint32_t* touched) {
  const int W = cfg->width, H = cfg->height, cells = W * H, S = PCGRL_MAX_STATS;
  uint8_t* map = b->map + (size_t)e * cells;
  uint8_t* smap = b->start_map + (size_t)e * cells;
  uint32_t* rng_rep = b->rng + (size_t)e * 2 * PCGRL_MT_WORDS;
  double* tp = b->tile_prob + (size_t)e * PCGRL_MAX_TILES;
  if ((cfg->flags & PCGRL_FLAG_RANDOM_START) || !b->start_valid[e]) {  // representation.py:40-45
    gen_random_map(rng_rep, tp, cfg->num_tiles, cells, map, smap, nullptr);
    b->start_valid[e] = 1;
  } else {
    memcpy(map, smap, (size_t)cells);
  }
  if (cfg->representation != PCGRL_REP_WIDE) {  // narrow_rep.py:30-31, turtle_rep.py:32-33
    b->pos[2 * e] = (uint8_t)mt_randint(rng_rep, W);
    b->pos[2 * e + 1] = (uint8_t)mt_randint(rng_rep, H);
  }
  int32_t st[PCGRL_MAX_STATS];
  host_get_stats(cfg, map, touched, hw, st);
  for (int i = 0; i < S; i++) {
    b->stats[(size_t)e * S + i] = st[i];
    b->start_stats[(size_t)e * S + i] = st[i];  // problem.py:45-46
  }
  if (cfg->problem == PCGRL_PROB_BINARY && (cfg->flags & PCGRL_FLAG_RANDOM_PROBS)) {  // binary_prob.py:68-72 (problem stream)
    tp[0] = mt_double(rng_rep + PCGRL_MT_WORDS);
    tp[1] = 1 - tp[0];
  }
  const size_t hb = (cfg->flags & PCGRL_FLAG_HEAT_U16) ? 2 : 1;
  memset((uint8_t*)b->heatmap + hb * (size_t)e * cells, 0, hb * (size_t)cells);  // pcgrl_env.py:72
  b->iteration[e] = 0;
  b->changes[e] = 0;
}

static inline void host_step_env(const pcgrl_config* cfg, const pcgrl_buffers* b, const int32_t* actions, int e, HostWork& hw,
                                 uint32_t* touched) {
  const int W = cfg->width, H = cfg->height, cells = W * H, S = PCGRL_MAX_STATS, NS = host_nstats(cfg);
  const int adim = cfg->representation == PCGRL_REP_WIDE ? 3 : (cfg->representation == PCGRL_REP_NARROWCAST || cfg->representation == PCGRL_REP_TURTLECAST) ? 2
                   : cfg->representation == PCGRL_REP_NARROWMULTI ? 9 : 1;
  uint8_t* map = b->map + (size_t)e * cells;
  int32_t* st = b->stats + (size_t)e * S;
  int32_t old[PCGRL_MAX_STATS];
  for (int i = 0; i < S; i++) old[i] = st[i];
  int x = 0, y = 0;
  if (cfg->representation != PCGRL_REP_WIDE) { x = b->pos[2 * e]; y = b->pos[2 * e + 1]; }
  b->iteration[e] += 1;  // pcgrl_env.py:130
  const Edit ed = apply_action(*cfg, actions + (size_t)e * adim, map, nullptr, touched, b->rng + (size_t)e * 2 * PCGRL_MT_WORDS, x, y);
  if (cfg->representation != PCGRL_REP_WIDE) { b->pos[2 * e] = (uint8_t)x; b->pos[2 * e + 1] = (uint8_t)y; }
  if (ed.change > 0) {  // :135-138
    b->changes[e] += ed.change;
    const size_t hi = (size_t)e * cells + (size_t)ed.hy * W + ed.hx;
    if (cfg->flags & PCGRL_FLAG_HEAT_U16) ((uint16_t*)b->heatmap)[hi] += 1; else ((uint8_t*)b->heatmap)[hi] += 1;
    int32_t ns[PCGRL_MAX_STATS];
    if (cfg->problem == PCGRL_PROB_SMB && !ed.solidity_touched) {  // exact search skipping (pcgrl_smb.cuh)
      for (int i = 0; i < S; i++) ns[i] = 0;
      scan_stats(map, W, H, ns);
      ns[5] = st[5]; ns[6] = st[6]; ns[7] = st[7];
    } else {
      host_get_stats(cfg, map, touched, hw, ns);
    }
    for (int i = 0; i < S; i++) st[i] = ns[i];
  }
  b->reward[e] = (ed.change > 0) ? host_reward(cfg, st, old) : 0.0;  // :142
  const bool done = host_over(cfg, st, b->start_stats + (size_t)e * S) || b->changes[e] >= cfg->max_changes ||
                    b->iteration[e] >= cfg->max_iterations;            // :143
  b->done[e] = done ? 1 : 0;
  int32_t* info = b->info_stats + (size_t)e * S;
  for (int i = 0; i < S; i++) info[i] = (i < NS) ? st[i] : 0;
  if (cfg->problem == PCGRL_PROB_BINARY) info[2] = st[1] - b->start_stats[(size_t)e * S + 1];  // binary_prob.py:137 "path-imp"
  info[PCGRL_INFO_ITERATION] = b->iteration[e];
  info[PCGRL_INFO_CHANGES] = b->changes[e];
  if (done && (cfg->flags & PCGRL_FLAG_AUTO_RESET)) host_reset_env(cfg, b, e, hw, touched);
}

}  // namespace pcgrl_smb

// pcgrl_solver.cuh -- the bounded BFS / A* play-through solvers of sokoban, ddave and mdungeon on the GPU.
//
// Reference: gym_pcgrl/envs/probs/{sokoban,ddave,mdungeon}/engine.py (State, Node, BFSAgent, AStarAgent) and
// the _run_game methods of the three *_prob.py files.  The result of a search depends on the exact pop
// order of CPython's binary heap under ties, on visited-set semantics and on best-node tie-breaks, so the
// search itself is reproduced operation by operation; what is B200-specific is the placement:
//
//   * the per-step API and rollouts run through k_rollout_async (pcgrl_b200.cu): every env warp runs _run_game
//     inline with the search routines of this file; idle warps take over posted passes;
//   * pcgrl_reset / pcgrl_get_stats (and PCGRL_SOLVER_ASYNC=0) use the queue form: envs whose map satisfies the solver
//     precondition are compacted into a work queue, k_solve runs one CTA per (queued env, pass) with the 4 passes of
//     _run_game SPECULATIVELY IN PARALLEL, a pass is cancelled as soon as a pass earlier in the reference's order has
//     won, the last CTA to finish merges the 4 results in the reference's order;
//   * binary heap (packed 32-bit entries: priority | node index), the visited hash table and a ring of the 128 newest
//     nodes live in shared memory (94.5 KB per search at power 5000: two searches per SM); the append-only node store
//     (32 B per node) lives in HBM scratch and stays L2 resident;
//   * states are fixed-width: 5 key words (exactly the information of State.getKey) + 3 payload words.
//
// Limits (checked by pcgrl_config_validate / reported through status[0]): width, height <= 14,
// width*height <= 128, solver_power in [1, 8000], at most 16 crates/targets in a sokoban level.
#pragma once
#include <stdlib.h>

#include "pcgrl_device.cuh"

namespace pcgrl {

enum { SOLVE_FOR_STEP = 0, SOLVE_FOR_RESET = 1, SOLVE_STATS_ONLY = 2 };
enum { GAME_SOKOBAN = 0, GAME_DDAVE = 1, GAME_MDUNGEON = 2 };

#define SOLVER_MAX_SLOTS 148
#define SOLVER_MAX_DEVICES 64
#define SOLVER_NODE_WORDS 8
#define SOLVER_PRIO_BIAS 2048

struct SolverQueue {
  int32_t* count;      // [1] number of queued items
  int32_t* items;      // [n] env | mode << 28
  int32_t* pass_done;  // [n] passes finished per item
  int32_t* best_win;   // [n] earliest pass (in reference order) that has won, 4 = none
  int32_t* results;    // [n][4][4]  {won, depth, h, counters}
  int32_t* status;
  int32_t capacity;
};

struct SolverLayout {
  size_t queue_bytes, old_stats_off, heat_off, nodes_off, async_off, heap_off, total;
  int slots;
  size_t nodes_per_pass, heap_words;
};

// k_rollout_async bookkeeping (see pcgrl_b200.cu): one header + one 128-byte request slot per env warp
#define ASYNC_MAX_CTAS 1024
#ifndef ASYNC_WPB
#define ASYNC_WPB 2 /* env warps per CTA (= per search arena): 174 registers x 64 threads lets four CTAs share an SM */
#endif
#ifndef ASYNC_HEAP_FAST
#define ASYNC_HEAP_FAST 2048 /* open-list entries kept in shared memory (the top 11 levels); the tail lives in HBM scratch */
#endif
#ifndef ASYNC_MIN_CTAS
#define ASYNC_MIN_CTAS 4 /* arenas per SM the register allocation must allow */
#endif
struct AsyncHeader { int32_t work, envs_done, posted, pad; };
struct __align__(16) AsyncGroup {
  int32_t env;         // env whose map (in HBM) is being solved
  int32_t open_mask;   // passes posted and not yet claimed (bit p)
  int32_t done_mask;   // passes finished, skipped or cancelled (bit p)
  int32_t best_win;    // earliest pass (reference order) that has won; -1 - p: pass p exhausted the state space; 4: none
  int32_t results[16]; // [pass][won / -1 cancelled, depth, h, counters]
  int32_t pad[12];
};
static inline size_t async_region_bytes() { return 256 + sizeof(AsyncGroup) * (size_t)ASYNC_MAX_CTAS * ASYNC_WPB; }

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static inline size_t solver_queue_bytes(int n) {
  // count(16 B) + items + pass_done + best_win + results
  return align_up(16 + sizeof(int32_t) * (size_t)n * (3 + 16), 256);
}

static inline SolverLayout solver_layout(const pcgrl_config* c, int n, int max_slots = SOLVER_MAX_SLOTS) {
  SolverLayout L;
  L.queue_bytes = solver_queue_bytes(n);
  L.old_stats_off = 2 * L.queue_bytes;
  L.heat_off = L.old_stats_off + align_up(sizeof(int32_t) * PCGRL_MAX_STATS * (size_t)n, 256);
  L.nodes_off = L.heat_off + align_up(6 * (size_t)n, 256);
  L.slots = n < max_slots ? n : max_slots;
  L.nodes_per_pass = (size_t)4 * (size_t)c->solver_power + 8;
  L.async_off = align_up(L.nodes_off + (size_t)L.slots * 4 * L.nodes_per_pass * SOLVER_NODE_WORDS * sizeof(uint32_t), 256);
  L.heap_words = ((size_t)3 * (size_t)c->solver_power + 9) & ~(size_t)1;  // (even) open-list tail of one k_rollout_async arena (one per CTA, <= 4 * slots CTAs)
  L.heap_off = align_up(L.async_off + async_region_bytes(), 256);
  L.total = L.heap_off + (size_t)L.slots * 4 * L.heap_words * sizeof(uint32_t);
  return L;
}

// Rollouts (T > 1) of the solver problems split the batch into independent env groups, one CUDA stream each,
// so that a slow search only stalls its own group (see rollout_solver).  Each group owns a scratch region.
#define SOLVER_MAX_GROUPS 64
#define SOLVER_DEFAULT_GROUPS 8
#define SOLVER_GROUP_MIN_ENVS 32
struct GroupPlan { int groups, envs_per_group, slots_per_group; size_t bytes_per_group; };
static inline GroupPlan solver_group_plan(const pcgrl_config* c, int n) {
  GroupPlan g;
  int max_groups = SOLVER_DEFAULT_GROUPS;  // tuning knob: PCGRL_SOLVER_GROUPS=1..64 (1 = plain lock-step batches)
  if (const char* env = getenv("PCGRL_SOLVER_GROUPS")) {
    const int v = atoi(env);
    if (v >= 1 && v <= SOLVER_MAX_GROUPS) max_groups = v;
  }
  g.groups = n / SOLVER_GROUP_MIN_ENVS;
  if (g.groups > max_groups) g.groups = max_groups;
  if (g.groups < 1) g.groups = 1;
  g.envs_per_group = ((n + g.groups - 1) / g.groups + 3) & ~3;  // multiple of 4: heat-map word atomics stay aligned
  g.slots_per_group = 2 * SOLVER_MAX_SLOTS / g.groups;
  if (g.slots_per_group < 8) g.slots_per_group = 8;
  g.bytes_per_group = align_up(solver_layout(c, g.envs_per_group, g.slots_per_group).total, 256);
  return g;
}
static inline size_t solver_scratch_bytes(const pcgrl_config* c, int n) {
  const GroupPlan g = solver_group_plan(c, n);
  const size_t single = solver_layout(c, n).total, grouped = (size_t)g.groups * g.bytes_per_group;
  return single > grouped ? single : grouped;  // evaluated with the same PCGRL_SOLVER_GROUPS the rollout will see
}

static inline int solver_validate(const pcgrl_config* c) {
  if (c->width > 14 || c->height > 14 || c->width * c->height > 128) return 1;
  if (c->solver_power < 1 || c->solver_power > 8000) return 2;
  return 0;
}
static inline const char* solver_validate_message(int rc) {
  return rc == 1 ? "solver problems support width, height <= 14 and width*height <= 128"
                 : "solver_power must be in [1, 8000]";
}

static inline SolverQueue solver_queue(const pcgrl_config* c, void* scratch, int n, int which, int32_t* status) {
  SolverQueue q;
  memset(&q, 0, sizeof(q));
  q.status = status;
  q.capacity = n;
  if (!scratch || c->problem < PCGRL_PROB_SOKOBAN) return q;
  char* base = (char*)scratch + (size_t)which * solver_queue_bytes(n);
  q.count = (int32_t*)base;
  q.items = (int32_t*)(base + 16);
  q.pass_done = q.items + n;
  q.best_win = q.pass_done + n;
  q.results = q.best_win + n;
  return q;
}
static inline int32_t* solver_old_stats(const pcgrl_config* c, void* scratch, int n) {
  return (int32_t*)((char*)scratch + solver_layout(c, n).old_stats_off);  // offsets do not depend on the slot count
}
static inline uint8_t* solver_heat_cell(const pcgrl_config* c, void* scratch, int n) {
  return (uint8_t*)scratch + solver_layout(c, n).heat_off;
}

__global__ void k_queue_clear(SolverQueue q, SolverQueue q2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) { *q.count = 0; if (q2.count) *q2.count = 0; }
  if (i < q.capacity) {
    q.pass_done[i] = 0; q.best_win[i] = 4;
    if (q2.count) { q2.pass_done[i] = 0; q2.best_win[i] = 4; }
  }
}
static inline void solver_queue_clear(SolverQueue q, cudaStream_t s) {
  SolverQueue none;
  memset(&none, 0, sizeof(none));
  if (q.count) k_queue_clear<<<(q.capacity + 255) / 256, 256, 0, s>>>(q, none);
}
static inline void solver_queue_clear2(SolverQueue q, SolverQueue q2, cudaStream_t s) {  // both queues, one launch
  if (q.count) k_queue_clear<<<(q.capacity + 255) / 256, 256, 0, s>>>(q, q2);
}

__device__ __forceinline__ void solver_enqueue(const SolverQueue& q, int env, int mode, int lane) {
  if (lane == 0) {
    const int slot = atomicAdd(q.count, 1);
    q.items[slot] = env | (mode << 28);
  }
}

// ------------------------------------------------------------------------------------------------
// level (shared memory) and state (registers)
// ------------------------------------------------------------------------------------------------
struct Level {
  int W, H, bw, bh;       // map size, bordered size
  uint32_t solid[16];     // bordered rows
  uint32_t dead[16];      // sokoban deadlock cells (engine.py:203-246)
  uint32_t spikes[16];    // ddave
  uint8_t tiles[128];     // map tiles, row-major (object types for mdungeon)
  uint8_t tx[16], ty[16]; // sokoban targets in row-major order
  int ntargets, ncrates;
  int doorx, doory, keyx, keyy;
  int overflow;
  // sokoban levels with bw*bh <= 64: one bit per bordered cell (bit y*bw + x)
  int small;
  unsigned long long solid64, dead64, target64;
};

// 5 key words (m[0..3], ks) == everything State.getKey can distinguish inside one level; payload: dh, misc.
//   sokoban : m = 16 crate bytes (x | y<<4, index order kept, 0xFF filler); ks = px | py<<8
//   ddave   : m = remaining-diamond bits over map cells; ks = px | py<<8 | health<<16 | key_present<<24;
//             misc = airTime | collected diamonds<<8 | jumps<<16
//   mdungeon: m = remaining potion/treasure/enemy bits; ks = px | py<<8 | health<<16;
//             misc = potions | treasures<<8 | enemies<<16
struct SState {
  uint32_t m[4];
  uint32_t ks;
  uint32_t dh;    // depth | (h + SOLVER_PRIO_BIAS) << 16
  uint32_t misc;
  uint32_t pad;
};

#define SOLVER_HD __host__ __device__ __forceinline__
SOLVER_HD int st_px(const SState& s) { return (int)(s.ks & 0xffu); }
SOLVER_HD int st_py(const SState& s) { return (int)((s.ks >> 8) & 0xffu); }
SOLVER_HD int st_health(const SState& s) { return (int)((s.ks >> 16) & 0xffu); }
SOLVER_HD void st_set_pos(SState& s, int x, int y) { s.ks = (s.ks & 0xffff0000u) | (uint32_t)x | ((uint32_t)y << 8); }
SOLVER_HD void st_set_health(SState& s, int h) { s.ks = (s.ks & 0xff00ffffu) | ((uint32_t)h << 16); }
SOLVER_HD int st_depth(const SState& s) { return (int)(s.dh & 0xffffu); }
SOLVER_HD int st_h(const SState& s) { return (int)(s.dh >> 16) - SOLVER_PRIO_BIAS; }

SOLVER_HD bool mask_test(const SState& s, int idx) {
  const uint32_t w = (idx < 64) ? ((idx < 32) ? s.m[0] : s.m[1]) : ((idx < 96) ? s.m[2] : s.m[3]);
  return (w >> (idx & 31)) & 1u;
}
SOLVER_HD void mask_clear(SState& s, int idx) {
  const uint32_t bit = ~(1u << (idx & 31));
#pragma unroll
  for (int w = 0; w < 4; w++) if ((idx >> 5) == w) s.m[w] &= bit;
}
SOLVER_HD bool lv_solid(const Level& L, int x, int y) { return (L.solid[y] >> x) & 1u; }
SOLVER_HD bool lv_movable(const Level& L, int x, int y) {  // ddave :204-205, mdungeon :201-202
  return !(x < 0 || y < 0 || x >= L.bw || y >= L.bh || lv_solid(L, x, y));
}
SOLVER_HD int iabs(int v) { return v < 0 ? -v : v; }

// --- sokoban (sokoban/engine.py) ----------------------------------------------------------------
SOLVER_HD int sk_crate_at(const SState& s, int x, int y) {  // :262-266 first match in list order
#ifdef __CUDA_ARCH__
  const uint32_t vv = 0x01010101u * (uint32_t)(x | (y << 4));
#pragma unroll
  for (int w = 0; w < 4; w++) {
    const uint32_t eq = __vcmpeq4(s.m[w], vv);
    if (eq) return 4 * w + ((__ffs(eq) - 1) >> 3);
  }
  return -1;
#else
  const uint32_t v = (uint32_t)(x | (y << 4));
  for (int i = 0; i < 16; i++)
    if (((s.m[i >> 2] >> (8 * (i & 3))) & 0xffu) == v) return i;
  return -1;
#endif
}
SOLVER_HD uint32_t sk_crate(const SState& s, int i) {
  const uint32_t w = (i < 8) ? ((i < 4) ? s.m[0] : s.m[1]) : ((i < 12) ? s.m[2] : s.m[3]);
  return (w >> (8 * (i & 3))) & 0xffu;
}
SOLVER_HD void sk_set_crate(SState& s, int i, int x, int y) {
  const int sh = 8 * (i & 3);
  const uint32_t v = (uint32_t)(x | (y << 4)) << sh, keep = ~(0xffu << sh);
#pragma unroll
  for (int w = 0; w < 4; w++) if ((i >> 2) == w) s.m[w] = (s.m[w] & keep) | v;
}
SOLVER_HD bool sk_movable(const Level& L, const SState& s, int x, int y) {  // :268-269
  if (x < 0 || y < 0 || x > L.bw - 1 || y > L.bh - 1) return false;
  return !lv_solid(L, x, y) && sk_crate_at(s, x, y) < 0;
}
SOLVER_HD bool sk_win(const Level& L, const SState& s) {  // :271-280
  if (L.ntargets != L.ncrates || L.ntargets == 0) return false;
  for (int t = 0; t < L.ntargets; t++) if (sk_crate_at(s, L.tx[t], L.ty[t]) < 0) return false;
  return true;
}
SOLVER_HD int sk_heuristic(const Level& L, const SState& s) {  // :282-296
  uint32_t used = 0;
  int distance = 0;
  for (int c = 0; c < L.ncrates; c++) {
    const uint32_t cr = sk_crate(s, c);
    const int cx = (int)(cr & 15u), cy = (int)(cr >> 4);
    int best = L.bw + L.bh, match = -1, first_free = -1;
    for (int t = 0; t < L.ntargets; t++) {
      if ((used >> t) & 1u) continue;
      if (first_free < 0) first_free = t;
      const int d = iabs(cx - (int)L.tx[t]) + iabs(cy - (int)L.ty[t]);
      if (best > d) { match = t; best = d; }
    }
    if (match < 0) match = first_free;  // bestMatch = 0 default (first remaining target)
    if (match < 0) break;
    distance += iabs((int)L.tx[match] - cx) + iabs((int)L.ty[match] - cy);
    used |= 1u << match;
  }
  return distance;
}
SOLVER_HD bool sk_update(const Level& L, SState& s, int dx, int dy) {  // :298-327 -> crateMove
  if (sk_win(L, s)) return false;
  const int nx = st_px(s) + dx, ny = st_py(s) + dy;
  if (sk_movable(L, s, nx, ny)) { st_set_pos(s, nx, ny); return false; }
  const int c = sk_crate_at(s, nx, ny);
  if (c >= 0) {
    const int cx = nx + dx, cy = ny + dy;
    if (sk_movable(L, s, cx, cy)) {
      st_set_pos(s, nx, ny);
      sk_set_crate(s, c, cx, cy);
      return true;
    }
  }
  return false;
}
SOLVER_HD bool sk_deadlocked(const Level& L, const SState& s) {  // :248-252
  for (int c = 0; c < L.ncrates; c++) {
    const uint32_t cr = sk_crate(s, c);
    if ((L.dead[cr >> 4] >> (cr & 15u)) & 1u) return true;
  }
  return false;
}
SOLVER_HD bool sk_target_at(const Level& L, int x, int y) {
  for (int t = 0; t < L.ntargets; t++) if (L.tx[t] == x && L.ty[t] == y) return true;
  return false;
}
__host__ __device__ inline void sk_init_deadlocks(Level& L) {  // :203-246 (single thread)
  for (int y = 0; y < 16; y++) L.dead[y] = 0;
  uint32_t corner[16];
  for (int y = 0; y < 16; y++) corner[y] = 0;
  for (int y = 1; y < L.bh - 1; y++)
    for (int x = 1; x < L.bw - 1; x++) {
      if (lv_solid(L, x, y)) continue;
      const bool up = lv_solid(L, x, y - 1), dn = lv_solid(L, x, y + 1), lf = lv_solid(L, x - 1, y), rt = lv_solid(L, x + 1, y);
      if (((up && lf) || (up && rt) || (dn && lf) || (dn && rt)) && !sk_target_at(L, x, y)) corner[y] |= 1u << x;
    }
  for (int y = 0; y < 16; y++) L.dead[y] = corner[y];
  // for every ordered pair of corners on one row / column: the open segment between them is dead if every cell
  // is a non-target floor cell hugging a wall on at least one side
  for (int y2 = 1; y2 < L.bh - 1; y2++)
    for (int x2 = 1; x2 < L.bw - 1; x2++) {
      if (!((corner[y2] >> x2) & 1u)) continue;
      for (int x1 = x2 + 1; x1 < L.bw - 1; x1++) {  // same row (both orders give the same cell set)
        if (!((corner[y2] >> x1) & 1u)) continue;
        bool ok = true;
        for (int x = x2 + 1; x < x1 && ok; x++)
          if (sk_target_at(L, x, y2) || lv_solid(L, x, y2) || (!lv_solid(L, x, y2 - 1) && !lv_solid(L, x, y2 + 1))) ok = false;
        if (ok) for (int x = x2 + 1; x < x1; x++) L.dead[y2] |= 1u << x;
      }
      for (int y1 = y2 + 1; y1 < L.bh - 1; y1++) {  // same column
        if (!((corner[y1] >> x2) & 1u)) continue;
        bool ok = true;
        for (int y = y2 + 1; y < y1 && ok; y++)
          if (sk_target_at(L, x2, y) || lv_solid(L, x2, y) || (!lv_solid(L, x2 - 1, y) && !lv_solid(L, x2 + 1, y))) ok = false;
        if (ok) for (int y = y2 + 1; y < y1; y++) L.dead[y] |= 1u << x2;
      }
    }
}

// --- ddave (ddave/engine.py) --------------------------------------------------------------------
SOLVER_HD int cell_index(const Level& L, int x, int y) { return (y - 1) * L.W + (x - 1); }
SOLVER_HD bool dd_win(const Level& L, const SState& s) {  // :319-320 (key count == 1 - key_present)
  return ((s.ks >> 24) & 1u) == 0u && st_px(s) == L.doorx && st_py(s) == L.doory;
}
SOLVER_HD void dd_update(const Level& L, SState& s, int dx, int dy) {  // :244-280 + :225-242
  if (dd_win(L, s) || st_health(s) <= 0) return;
  const int px = st_px(s), py = st_py(s);
  int air = (int)(s.misc & 0xffu), diamonds = (int)((s.misc >> 8) & 0xffu), jumps = (int)(s.misc >> 16);
  const bool ground = lv_solid(L, px, py + 1), ceiling = lv_solid(L, px, py - 1);
  int nx = px, ny = py;
  if (dx != 0) {
    if (lv_movable(L, nx + dx, ny)) nx += dx;
  } else if (dy < 0) {
    if (ground && !ceiling) { air = 3; jumps++; }
  }
  if (air > 1) {
    air--;
    if (lv_movable(L, nx, ny - 1)) ny--; else air = 1;
  } else if (air > 0) {
    air--;
  } else {
    if (lv_movable(L, nx, ny + 1)) ny++;
  }
  st_set_pos(s, nx, ny);
  const int idx = cell_index(L, nx, ny);
  if (idx >= 0 && idx < 128 && nx >= 1 && ny >= 1 && nx <= L.W && ny <= L.H) {
    if (mask_test(s, idx)) { diamonds++; mask_clear(s, idx); }                   // diamond
    else if ((L.spikes[ny] >> nx) & 1u) st_set_health(s, 0);                       // spike
    else if (((s.ks >> 24) & 1u) && nx == L.keyx && ny == L.keyy) s.ks &= ~(1u << 24);  // key
  }
  s.misc = (uint32_t)air | ((uint32_t)diamonds << 8) | ((uint32_t)jumps << 16);
}
SOLVER_HD int dd_heuristic(const Level& L, const SState& s) {  // :294-299
  int d = iabs(st_px(s) - L.doorx) + iabs(st_py(s) - L.doory);
  if ((s.ks >> 24) & 1u) d = iabs(st_px(s) - L.keyx) + iabs(st_py(s) - L.keyy) + (L.bw + L.bh);
  return d - 5 * (int)((s.misc >> 8) & 0xffu);
}

// --- mdungeon (mdungeon/engine.py) ---------------------------------------------------------------
SOLVER_HD bool md_win(const Level& L, const SState& s) { return st_px(s) == L.doorx && st_py(s) == L.doory; }  // :308-309
SOLVER_HD void md_update(const Level& L, SState& s, int dx, int dy) {  // :254-270 + :222-252
  if (md_win(L, s) || st_health(s) <= 0) return;
  const int nx = st_px(s) + dx, ny = st_py(s) + dy;
  if (!lv_movable(L, nx, ny)) return;
  st_set_pos(s, nx, ny);
  const int idx = cell_index(L, nx, ny);
  if (nx >= 1 && ny >= 1 && nx <= L.W && ny <= L.H && mask_test(s, idx)) {
    const int t = L.tiles[idx];
    int health = st_health(s);
    int potions = (int)(s.misc & 0xffu), treasures = (int)((s.misc >> 8) & 0xffu), enemies = (int)((s.misc >> 16) & 0xffu);
    if (t == 4) { health = min(health + 2, 5); potions++; }
    else if (t == 5) { treasures++; }
    else { enemies++; health = max(health - (t == 6 ? 1 : 2), 0); }
    mask_clear(s, idx);
    st_set_health(s, health);
    s.misc = (uint32_t)potions | ((uint32_t)treasures << 8) | ((uint32_t)enemies << 16);
  }
}
SOLVER_HD int md_heuristic(const Level& L, const SState& s) {  // :285-289
  return iabs(st_px(s) - L.doorx) + iabs(st_py(s) - L.doory) + 4 * (5 - st_health(s)) - 4 * (int)((s.misc >> 8) & 0xffu);
}

template <int GAME> SOLVER_HD bool g_win(const Level& L, const SState& s) {
  return GAME == GAME_SOKOBAN ? sk_win(L, s) : GAME == GAME_DDAVE ? dd_win(L, s) : md_win(L, s);
}
template <int GAME> SOLVER_HD int g_heuristic(const Level& L, const SState& s) {
  return GAME == GAME_SOKOBAN ? sk_heuristic(L, s) : GAME == GAME_DDAVE ? dd_heuristic(L, s) : md_heuristic(L, s);
}

// *_prob.py _run_game level framing + engine.stringInitialize; run by lane 0 after the tiles are staged
template <int GAME>
__host__ __device__ inline void level_init(Level& L, SState& root, int W, int H) {
  L.W = W; L.H = H; L.bw = W + 2; L.bh = H + 2;
  L.ntargets = 0; L.ncrates = 0; L.overflow = 0;
  L.doorx = L.doory = L.keyx = L.keyy = 0;
  for (int y = 0; y < 16; y++) { L.solid[y] = 0; L.spikes[y] = 0; L.dead[y] = 0; }
  root.m[0] = root.m[1] = root.m[2] = root.m[3] = (GAME == GAME_SOKOBAN) ? 0xffffffffu : 0u;
  root.ks = 0; root.dh = 0; root.misc = 0; root.pad = 0;
  for (int y = 0; y < L.bh; y++)
    for (int x = 0; x < L.bw; x++) {
      const bool border = (x == 0 || y == 0 || x == L.bw - 1 || y == L.bh - 1);
      const int t = border ? 1 : L.tiles[(y - 1) * W + (x - 1)];
      if (t == 1) { L.solid[y] |= 1u << x; continue; }
      if (t == 0) continue;
      const int idx = (y - 1) * W + (x - 1);
      if (GAME == GAME_SOKOBAN) {
        if (t == 2) st_set_pos(root, x, y);
        if (t == 3) { if (L.ncrates < 16) { sk_set_crate(root, L.ncrates, x, y); L.ncrates++; } else L.overflow = 1; }
        if (t == 4) { if (L.ntargets < 16) { L.tx[L.ntargets] = (uint8_t)x; L.ty[L.ntargets] = (uint8_t)y; L.ntargets++; } else L.overflow = 1; }
      } else if (GAME == GAME_DDAVE) {
        if (t == 2) { st_set_pos(root, x, y); st_set_health(root, 1); }
        if (t == 3) { L.doorx = x; L.doory = y; }
        if (t == 4) {
#pragma unroll
          for (int w = 0; w < 4; w++) if ((idx >> 5) == w) root.m[w] |= 1u << (idx & 31);
        }
        if (t == 5) { L.keyx = x; L.keyy = y; root.ks |= 1u << 24; }
        if (t == 6) L.spikes[y] |= 1u << x;
      } else {
        if (t == 2) { st_set_pos(root, x, y); st_set_health(root, 5); }
        if (t == 3) { L.doorx = x; L.doory = y; }
        if (t >= 4) {
#pragma unroll
          for (int w = 0; w < 4; w++) if ((idx >> 5) == w) root.m[w] |= 1u << (idx & 31);
        }
      }
    }
  L.small = 0; L.solid64 = 0ull; L.dead64 = 0ull; L.target64 = 0ull;
  if (GAME == GAME_SOKOBAN) {
    sk_init_deadlocks(L);
    if (L.bw * L.bh <= 64) {
      L.small = 1;
      for (int y = 0; y < L.bh; y++)
        for (int x = 0; x < L.bw; x++) {
          const unsigned long long bit = 1ull << (y * L.bw + x);
          if ((L.solid[y] >> x) & 1u) L.solid64 |= bit;
          if ((L.dead[y] >> x) & 1u) L.dead64 |= bit;
        }
      for (int t = 0; t < L.ntargets; t++) L.target64 |= 1ull << ((int)L.ty[t] * L.bw + (int)L.tx[t]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// search
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void node_store(uint32_t* nodes, int i, const SState& s) {
  uint4* p = reinterpret_cast<uint4*>(nodes + (size_t)i * SOLVER_NODE_WORDS);
  p[0] = make_uint4(s.m[0], s.m[1], s.m[2], s.m[3]);
  p[1] = make_uint4(s.ks, s.dh, s.misc, s.pad);
}
__device__ __forceinline__ void node_load(const uint32_t* nodes, int i, SState& s) {
  const uint4* p = reinterpret_cast<const uint4*>(nodes + (size_t)i * SOLVER_NODE_WORDS);
  const uint4 a = p[0], b = p[1];
  s.m[0] = a.x; s.m[1] = a.y; s.m[2] = a.z; s.m[3] = a.w;
  s.ks = b.x; s.dh = b.y; s.misc = b.z; s.pad = b.w;
}
__device__ __forceinline__ uint32_t key_hash(const SState& s) {
  uint32_t h = 2166136261u;
  h = (h ^ s.m[0]) * 16777619u; h = (h ^ s.m[1]) * 16777619u; h = (h ^ s.m[2]) * 16777619u;
  h = (h ^ s.m[3]) * 16777619u; h = (h ^ s.ks) * 16777619u;
  return h ^ (h >> 15);
}

// CPython heapq on packed entries (priority << 15 | node); comparisons use the priority only, strict <
// (Lib/heapq.py _siftdown / _siftup; engine.py Node.__lt__ with 2*h + b*depth, b = 2*balance).
#define HP(e) ((e) >> 15)
// The open list, 1-based (entry i in 1..n; the children of i are the aligned 8-byte pair 2i, 2i+1).  Entries below `cap`
// live in shared memory (addressed as .shared: `fast_s`), the tail in a per-arena slice of HBM scratch; cap is even so a
// child pair never straddles the two.  k_solve keeps everything in shared memory (cap = INT_MAX).
struct HeapRef {
  uint32_t fast_s;        // shared-space byte address of entry 0
  uint32_t* slow_biased;  // slow - cap: entry i >= cap lives at slow_biased[i]
  int cap;
};
__device__ __forceinline__ HeapRef heap_ref(uint32_t* fast, uint32_t* slow, int cap) {
  HeapRef h;
  h.fast_s = (uint32_t)__cvta_generic_to_shared(fast);
  h.slow_biased = slow ? slow - cap : nullptr;
  h.cap = cap;
  return h;
}
__device__ __forceinline__ uint32_t hget(const HeapRef& h, int i) {
  if (i < h.cap) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(h.fast_s + 4u * (uint32_t)i) : "memory");
    return v;
  }
  return h.slow_biased[i];
}
__device__ __forceinline__ uint2 hget_pair(const HeapRef& h, int i) {  // i even
  if (i < h.cap) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(h.fast_s + 4u * (uint32_t)i) : "memory");
    return v;
  }
  return *reinterpret_cast<const uint2*>(h.slow_biased + i);
}
__device__ __forceinline__ void hset(const HeapRef& h, int i, uint32_t v) {
  if (i < h.cap) asm volatile("st.shared.u32 [%0], %1;" :: "r"(h.fast_s + 4u * (uint32_t)i), "r"(v) : "memory");
  else h.slow_biased[i] = v;
}
// heapq._siftdown(heap, 0, pos) for `item` placed in the hole at `pos`
__device__ __forceinline__ void heap_siftdown(const HeapRef& heap, int pos, uint32_t item) {
  while (pos > 1) {
    const int parentpos = pos >> 1;
    const uint32_t parent = hget(heap, parentpos);
    if (HP(item) < HP(parent)) { hset(heap, pos, parent); pos = parentpos; continue; }
    break;
  }
  hset(heap, pos, item);
}
__device__ __forceinline__ void heap_push(const HeapRef& heap, int& n, uint32_t item) { heap_siftdown(heap, ++n, item); }
// heapq.heappop: the last entry refills the root through _siftup (smaller child up to a leaf, then _siftdown)
__device__ __forceinline__ uint32_t heap_pop(const HeapRef& heap, int& n) {
  const uint32_t last = hget(heap, n);
  n--;
  if (n == 0) return last;
  const uint32_t ret = hget(heap, 1);
  int pos = 1, childpos = 2;
  while (childpos <= n) {
    const uint2 pr = hget_pair(heap, childpos);  // (left, right); right is ignored when it is past the end
    uint32_t child = pr.x;
    if (childpos < n && !(HP(pr.x) < HP(pr.y))) { childpos++; child = pr.y; }
    hset(heap, pos, child);
    pos = childpos;
    childpos = 2 * pos;
  }
  heap_siftdown(heap, pos, last);
  return ret;
}

// crate occupancy of a sokoban state as a 64-bit mask over bordered cells (small levels only)
SOLVER_HD unsigned long long sk_occupancy(const Level& L, const SState& s) {
  unsigned long long occ = 0ull;
  for (int c = 0; c < L.ncrates; c++) {
    const uint32_t cr = sk_crate(s, c);
    occ |= 1ull << ((int)(cr >> 4) * L.bw + (int)(cr & 15u));
  }
  return occ;
}

// One child of Node.getChildren: direction d in the engine's `directions` order applied to a copy `c` of the expanded
// node `cs`; returns whether the child is kept (sokoban prunes no-move and deadlocked children, engine.py:14-24) and its
// heuristic in h.  Shared by the warp searches and the host twin.
template <int GAME>
SOLVER_HD bool make_child(const Level& L, const SState& cs, int d, bool sk_small, SState& c, int& h) {
  bool valid = false;
  h = 0;
  if (GAME == GAME_SOKOBAN) {  // engine.py:3 and :14-24
    const int dx = (d == 0) ? -1 : (d == 1) ? 1 : 0, dy = (d == 2) ? -1 : (d == 3) ? 1 : 0;
    if (sk_small) {
      // bit-mask form of State.update (:298-327): the expanded node is never a winning state
      const unsigned long long occ = (unsigned long long)cs.misc | ((unsigned long long)cs.pad << 32);
      const int nx = st_px(cs) + dx, ny = st_py(cs) + dy;
      const unsigned long long nbit = 1ull << (ny * L.bw + nx);  // inside: the border ring is solid
      h = st_h(cs);
      if (!((L.solid64 | occ) & nbit)) {
        st_set_pos(c, nx, ny);
        valid = true;
      } else if (occ & nbit) {
        const int cx = nx + dx, cy = ny + dy;
        if (cx >= 0 && cy >= 0 && cx < L.bw && cy < L.bh) {
          const unsigned long long cbit = 1ull << (cy * L.bw + cx);
          if (!((L.solid64 | occ) & cbit)) {
            st_set_pos(c, nx, ny);
            sk_set_crate(c, sk_crate_at(cs, nx, ny), cx, cy);
            const unsigned long long occ2 = occ ^ nbit ^ cbit;
            c.misc = (uint32_t)occ2;
            c.pad = (uint32_t)(occ2 >> 32);
            valid = ((occ2 & L.dead64) == 0ull);  // a crate moved: prune deadlocks (any crate)
            h = sk_heuristic(L, c);
          }
        }
      }
    } else {
      const bool crate_move = sk_update(L, c, dx, dy);
      valid = (c.ks & 0xffffu) != (cs.ks & 0xffffu) && !(crate_move && sk_deadlocked(L, c));
      h = sk_heuristic(L, c);
    }
  } else if (GAME == GAME_DDAVE) {  // ddave/engine.py:3  (0,0) (-1,0) (1,0) (0,-1)
    const int dx = (d == 1) ? -1 : (d == 2) ? 1 : 0, dy = (d == 3) ? -1 : 0;
    dd_update(L, c, dx, dy);
    valid = true;
    h = dd_heuristic(L, c);
  } else {  // mdungeon/engine.py:3
    const int dx = (d == 0) ? -1 : (d == 1) ? 1 : 0, dy = (d == 2) ? -1 : (d == 3) ? 1 : 0;
    md_update(L, c, dx, dy);
    valid = true;
    h = md_heuristic(L, c);
  }
  return valid;
}

#define SOLVER_CACHE_NODES 128 /* shared-memory ring of the most recently created nodes */
enum { ACT_EXPAND = 0, ACT_SKIP = 1, ACT_STOP = 2 };

__device__ __forceinline__ void node_fetch(const uint32_t* nodes, const uint32_t* cache, int i, int nn, SState& s) {
  if (i >= nn - SOLVER_CACHE_NODES) {
    const uint4* p = reinterpret_cast<const uint4*>(cache + (size_t)(i & (SOLVER_CACHE_NODES - 1)) * SOLVER_NODE_WORDS);
    const uint4 a = p[0], b = p[1];
    s.m[0] = a.x; s.m[1] = a.y; s.m[2] = a.z; s.m[3] = a.w;
    s.ks = b.x; s.dh = b.y; s.misc = b.z; s.pad = b.w;
  } else {
    node_load(nodes, i, s);
  }
}
__device__ __forceinline__ void node_put(uint32_t* nodes, uint32_t* cache, int i, const SState& s) {
  node_store(nodes, i, s);
  uint4* p = reinterpret_cast<uint4*>(cache + (size_t)(i & (SOLVER_CACHE_NODES - 1)) * SOLVER_NODE_WORDS);
  p[0] = make_uint4(s.m[0], s.m[1], s.m[2], s.m[3]);
  p[1] = make_uint4(s.ks, s.dh, s.misc, s.pad);
}

// One pass of _run_game: BFSAgent / AStarAgent.getSolution (b < 0: BFS), executed by one warp:
//   lane 0      pops (CPython heap order / FIFO), checks lose / win / visited, tracks the best node, pushes;
//   lanes 0..3  generate the four children of the expanded node in parallel (Node.getChildren order = lane).
// Returns (lane 0): res[0] won, res[1] depth, res[2] h, res[3] misc counters of solState; res[0] = -1 if cancelled.
template <int GAME>
__device__ void search_pass(const Level& L, const SState& root0, int b, int power, uint32_t* nodes, uint32_t* cache,
                            const HeapRef heap, uint32_t* table, int table_mask, const volatile int32_t* best_win,
                            int pass_index, int* res, int* exhausted, int lane) {
  const bool check_lose = (GAME != GAME_SOKOBAN);
  const bool sk_small = (GAME == GAME_SOKOBAN) && L.small;
  int nn = 1, nheap = 0, head = 0, iterations = 0;
  int best = -1, best_h = 0, best_depth = 0;
  if (lane == 0) {
    SState root = root0;
    if (sk_small) {
      const unsigned long long occ = sk_occupancy(L, root);
      root.misc = (uint32_t)occ;
      root.pad = (uint32_t)(occ >> 32);
    }
    root.dh = 0u | ((uint32_t)(g_heuristic<GAME>(L, root) + SOLVER_PRIO_BIAS) << 16);
    node_put(nodes, cache, 0, root);
    if (b >= 0) heap_push(heap, nheap, ((uint32_t)(2 * st_h(root) + 2 * SOLVER_PRIO_BIAS) << 15) | 0u);
    res[0] = 0;
    *exhausted = 0;
  }
  __syncwarp();
  while (true) {
    int action = ACT_STOP, cur = 0;
    SState cs;
    cs.m[0] = cs.m[1] = cs.m[2] = cs.m[3] = cs.ks = cs.dh = cs.misc = cs.pad = 0u;
    if (lane == 0) {
      if (!(iterations < power && (b >= 0 ? nheap > 0 : head < nn))) {
        action = ACT_STOP;
      } else if ((iterations & 31) == 0 && *best_win < pass_index) {
        res[0] = -1;
        action = ACT_STOP;
      } else {
        iterations++;
        cur = (b >= 0) ? (int)(heap_pop(heap, nheap) & 0x7fffu) : head++;
        node_fetch(nodes, cache, cur, nn, cs);
        bool win;
        if (sk_small) {  // small sokoban levels carry their crate-occupancy mask in (misc, pad)
          const unsigned long long occ = (unsigned long long)cs.misc | ((unsigned long long)cs.pad << 32);
          win = (occ & L.target64) == L.target64 && L.ntargets == L.ncrates && L.ntargets > 0;
        } else {
          win = g_win<GAME>(L, cs);
        }
        if (check_lose && st_health(cs) <= 0) {
          action = ACT_SKIP;
        } else if (win) {
          res[0] = 1; res[1] = st_depth(cs); res[2] = st_h(cs); res[3] = (int)cs.misc;
          action = ACT_STOP;
        } else {
          // visited set: open addressing, entry = fingerprint << 15 | (node + 1); exact key compare on a fingerprint hit
          const uint32_t hsh = key_hash(cs);
          const uint32_t fp = (hsh >> 15) & 0x1ffffu;
          uint32_t slot = hsh & (uint32_t)table_mask;
          bool seen = false;
          while (true) {
            const uint32_t ent = table[slot];
            if (ent == 0u) break;
            if ((ent >> 15) == fp) {
              SState o;
              node_fetch(nodes, cache, (int)(ent & 0x7fffu) - 1, nn, o);
              if (o.m[0] == cs.m[0] && o.m[1] == cs.m[1] && o.m[2] == cs.m[2] && o.m[3] == cs.m[3] && o.ks == cs.ks) { seen = true; break; }
            }
            slot = (slot + 1) & (uint32_t)table_mask;
          }
          if (seen) {
            action = ACT_SKIP;
          } else {
            const int ch = st_h(cs), cd = st_depth(cs);
            if (best < 0 || ch < best_h || (ch == best_h && cd < best_depth)) { best = cur; best_h = ch; best_depth = cd; }
            table[slot] = (fp << 15) | (uint32_t)(cur + 1);
            action = ACT_EXPAND;
          }
        }
      }
    }
    action = __shfl_sync(FULL_MASK, action, 0);
    if (action == ACT_STOP) break;
    if (action == ACT_SKIP) continue;
    // broadcast the expanded node; lanes 0..3 build one child each (Node.getChildren, `directions` order)
    cs.m[0] = __shfl_sync(FULL_MASK, cs.m[0], 0); cs.m[1] = __shfl_sync(FULL_MASK, cs.m[1], 0);
    cs.m[2] = __shfl_sync(FULL_MASK, cs.m[2], 0); cs.m[3] = __shfl_sync(FULL_MASK, cs.m[3], 0);
    cs.ks = __shfl_sync(FULL_MASK, cs.ks, 0); cs.dh = __shfl_sync(FULL_MASK, cs.dh, 0);
    cs.misc = __shfl_sync(FULL_MASK, cs.misc, 0);
    cs.pad = __shfl_sync(FULL_MASK, cs.pad, 0);
    const int nn_base = __shfl_sync(FULL_MASK, nn, 0);
    const int cd = st_depth(cs);
    bool valid = false;
    SState c = cs;
    int h = 0;
    if (lane < 4) {
      valid = make_child<GAME>(L, cs, lane, sk_small, c, h);  // Node.getChildren, `directions` order = lane
      c.dh = (uint32_t)(cd + 1) | ((uint32_t)(h + SOLVER_PRIO_BIAS) << 16);
    }
    const uint32_t vmask = __ballot_sync(FULL_MASK, valid) & 0xFu;
    if (valid) node_put(nodes, cache, nn_base + __popc(vmask & ((1u << lane) - 1u)), c);
    const int prio = 2 * h + b * (cd + 1) + 2 * SOLVER_PRIO_BIAS;
    int idx = nn_base;
#pragma unroll
    for (int d = 0; d < 4; d++) {
      const int pd = __shfl_sync(FULL_MASK, prio, d);
      if ((vmask >> d) & 1u) {
        if (lane == 0 && b >= 0) heap_push(heap, nheap, ((uint32_t)pd << 15) | (uint32_t)idx);
        idx++;
      }
    }
    nn = idx;
    __syncwarp();
  }
  if (lane == 0 && res[0] == 0) {
    SState bs;
    node_fetch(nodes, cache, best < 0 ? 0 : best, nn, bs);
    res[1] = st_depth(bs); res[2] = st_h(bs); res[3] = (int)bs.misc;
    *exhausted = (b >= 0 ? nheap == 0 : head >= nn) ? 1 : 0;
  }
  __syncwarp();
}

// ------------------------------------------------------------------------------------------------
// BFSAgent.getSolution, 32 queue nodes per round.  The FIFO queue of the reference IS the node list in creation
// order, so iterations k .. k+31 pop nodes head .. head+31, which all exist already.  Processing them together is
// exact because a node of the batch influences a later one only through
//   (1) the visited set  -> a node is accepted iff it is not lost, its key is not in the table and no EARLIER lane
//                           of the batch holds the same key (key equality is an equivalence, the first lane wins);
//   (2) the best node    -> ordered reduction: minimum (h, depth), earliest node on ties, the current best on ties;
//   (3) the child order  -> children are appended at prefix-sum offsets (lane order, then `directions` order);
// and the first winning node in lane order ends the search with iterations = its queue position + 1.
// Returns (lane 0 writes res): res[0] won / -1 cancelled, res[1] depth, res[2] h, res[3] misc; *exhausted = 1 when
// the queue ran empty without a win.
// ------------------------------------------------------------------------------------------------
template <int GAME>
__device__ void search_bfs_batched(const Level& L, const SState& root0, int power, uint32_t* nodes, uint32_t* table,
                                   int table_mask, const volatile int32_t* best_win, int pass_index, int* res,
                                   int* exhausted, int lane) {
  const bool check_lose = (GAME != GAME_SOKOBAN);
  const bool sk_small = (GAME == GAME_SOKOBAN) && L.small;
  int nn = 1, head = 0, iterations = 0;
  int best = -1, best_h = 0, best_depth = 0;
  if (lane == 0) {
    SState root = root0;
    if (sk_small) {
      const unsigned long long occ = sk_occupancy(L, root);
      root.misc = (uint32_t)occ;
      root.pad = (uint32_t)(occ >> 32);
    }
    root.dh = 0u | ((uint32_t)(g_heuristic<GAME>(L, root) + SOLVER_PRIO_BIAS) << 16);
    node_store(nodes, 0, root);
    res[0] = 0;
    *exhausted = 0;
  }
  __syncwarp();
  while (iterations < power && head < nn) {
    int cancel = 0;
    if (lane == 0) cancel = (*best_win < pass_index) ? 1 : 0;
    if (__shfl_sync(FULL_MASK, cancel, 0)) {
      if (lane == 0) res[0] = -1;
      __syncwarp();
      return;
    }
    const int B = min(32, min(nn - head, power - iterations));
    const bool active = lane < B;
    SState cs;
    cs.m[0] = cs.m[1] = cs.m[2] = cs.m[3] = cs.ks = cs.dh = cs.misc = cs.pad = 0u;
    if (active) node_load(nodes, head + lane, cs);
    const bool lost = active && check_lose && st_health(cs) <= 0;
    bool win = false;
    if (active && !lost) {
      if (sk_small) {
        const unsigned long long occ = (unsigned long long)cs.misc | ((unsigned long long)cs.pad << 32);
        win = (occ & L.target64) == L.target64 && L.ntargets == L.ncrates && L.ntargets > 0;
      } else {
        win = g_win<GAME>(L, cs);
      }
    }
    const unsigned winmask = __ballot_sync(FULL_MASK, win);
    if (winmask) {  // first winning node in queue order
      const int w = __ffs(winmask) - 1;
      if (lane == w) { res[0] = 1; res[1] = st_depth(cs); res[2] = st_h(cs); res[3] = (int)cs.misc; }
      __syncwarp();
      return;
    }
    iterations += B;
    const bool cand = active && !lost;
    // (1a) visited table (read-only here): exact key compare on a fingerprint hit
    const uint32_t hsh = key_hash(cs);
    const uint32_t fp = (hsh >> 15) & 0x1ffffu;
    uint32_t slot = hsh & (uint32_t)table_mask;
    bool seen = false;
    if (cand) {
      while (true) {
        const uint32_t ent = table[slot];
        if (ent == 0u) break;
        if ((ent >> 15) == fp) {
          SState o;
          node_load(nodes, (int)(ent & 0x7fffu) - 1, o);
          if (o.m[0] == cs.m[0] && o.m[1] == cs.m[1] && o.m[2] == cs.m[2] && o.m[3] == cs.m[3] && o.ks == cs.ks) { seen = true; break; }
        }
        slot = (slot + 1) & (uint32_t)table_mask;
      }
    }
    // (1b) duplicates inside the batch: group lanes by hash, compare the full key with the group's first lane
    const bool fresh = cand && !seen;
    const unsigned long long mval = fresh ? (unsigned long long)hsh : (0x100000000ull | (unsigned long long)lane);
    const unsigned grp = __match_any_sync(FULL_MASK, mval);
    const int leader = __ffs(grp) - 1;
    const uint32_t l0 = __shfl_sync(FULL_MASK, cs.m[0], leader), l1 = __shfl_sync(FULL_MASK, cs.m[1], leader);
    const uint32_t l2 = __shfl_sync(FULL_MASK, cs.m[2], leader), l3 = __shfl_sync(FULL_MASK, cs.m[3], leader);
    const uint32_t l4 = __shfl_sync(FULL_MASK, cs.ks, leader);
    const bool same_as_leader = (l0 == cs.m[0] && l1 == cs.m[1] && l2 == cs.m[2] && l3 == cs.m[3] && l4 == cs.ks);
    bool dup = fresh && leader != lane && same_as_leader;
    const bool collision = fresh && leader != lane && !same_as_leader;  // equal hash, different key (rare)
    if (__any_sync(FULL_MASK, collision)) {
      // exact fallback: every lane in turn broadcasts its key; later fresh lanes with the same key become duplicates
      dup = false;
      for (int src = 0; src < B; src++) {
        const bool src_ok = __shfl_sync(FULL_MASK, (int)(fresh && !dup), src) != 0;
        const uint32_t s0 = __shfl_sync(FULL_MASK, cs.m[0], src), s1 = __shfl_sync(FULL_MASK, cs.m[1], src);
        const uint32_t s2 = __shfl_sync(FULL_MASK, cs.m[2], src), s3 = __shfl_sync(FULL_MASK, cs.m[3], src);
        const uint32_t s4 = __shfl_sync(FULL_MASK, cs.ks, src);
        if (src_ok && lane > src && fresh && s0 == cs.m[0] && s1 == cs.m[1] && s2 == cs.m[2] && s3 == cs.m[3] && s4 == cs.ks) dup = true;
      }
    }
    const bool accept = fresh && !dup;
    if (accept) {  // insert (other lanes insert concurrently: claim an empty slot with CAS)
      const uint32_t ent = (fp << 15) | (uint32_t)(head + lane + 1);
      while (atomicCAS(&table[slot], 0u, ent) != 0u) slot = (slot + 1) & (uint32_t)table_mask;
    }
    // (2) best node: minimum (h, depth), earliest lane; the current best survives ties
    {
      unsigned long long k = 0xffffffffffffffffull;
      if (accept) k = ((unsigned long long)(uint32_t)(st_h(cs) + SOLVER_PRIO_BIAS) << 24) | ((unsigned long long)st_depth(cs) << 8) | (unsigned long long)lane;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(FULL_MASK, k, o);
        k = other < k ? other : k;
      }
      if (k != 0xffffffffffffffffull) {
        const int bh = (int)(k >> 24) - SOLVER_PRIO_BIAS, bd = (int)((k >> 8) & 0xffffu), bl = (int)(k & 0xffu);
        if (best < 0 || bh < best_h || (bh == best_h && bd < best_depth)) { best = head + bl; best_h = bh; best_depth = bd; }
      }
    }
    // (3) children of the accepted nodes, appended in (lane, direction) order
    SState c[4];
    int h4[4];
    unsigned vmask = 0;
    if (accept) {
      const int cd = st_depth(cs);
#pragma unroll
      for (int d = 0; d < 4; d++) {
        c[d] = cs;
        int h = 0;
        const bool valid = make_child<GAME>(L, cs, d, sk_small, c[d], h);
        c[d].dh = (uint32_t)(cd + 1) | ((uint32_t)(h + SOLVER_PRIO_BIAS) << 16);
        h4[d] = h;
        if (valid) vmask |= 1u << d;
      }
    }
    int cnt = __popc(vmask), incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(FULL_MASK, incl, o);
      if (lane >= o) incl += v;
    }
    const int total = __shfl_sync(FULL_MASK, incl, 31);
    int dst = nn + incl - cnt;
#pragma unroll
    for (int d = 0; d < 4; d++)
      if ((vmask >> d) & 1u) { node_store(nodes, dst, c[d]); dst++; }
    (void)h4;
    nn += total;
    head += B;
    __syncwarp();
  }
  if (lane == 0) {
    SState bs;
    node_load(nodes, best < 0 ? 0 : best, bs);
    res[1] = st_depth(bs); res[2] = st_h(bs); res[3] = (int)bs.misc;
    *exhausted = (head >= nn) ? 1 : 0;
  }
  __syncwarp();
}

template <int PROB> struct GameOf;
template <> struct GameOf<PCGRL_PROB_SOKOBAN> { static constexpr int GAME = GAME_SOKOBAN; };
template <> struct GameOf<PCGRL_PROB_DDAVE> { static constexpr int GAME = GAME_DDAVE; };
template <> struct GameOf<PCGRL_PROB_MDUNGEON> { static constexpr int GAME = GAME_MDUNGEON; };
template <> struct GameOf<PCGRL_PROB_BINARY> { static constexpr int GAME = -1; };
template <> struct GameOf<PCGRL_PROB_ZELDA> { static constexpr int GAME = -1; };

// grid = 4 * slots CTAs of one warp; CTA c runs pass (c & 3) of queue items (c >> 2), (c >> 2) + slots, ...
template <int PROB>
__global__ void __launch_bounds__(32) k_solve(const __grid_constant__ pcgrl_config cfg, int32_t* stats, int32_t* start_stats,
                                              const uint8_t* __restrict__ maps, SolverQueue q, uint32_t* node_pool,
                                              size_t nodes_per_pass, int slots, int table_size) {
  constexpr int GAME = GameOf<PROB>::GAME;
  extern __shared__ uint32_t dyn[];
  __shared__ Level L;
  const int lane = threadIdx.x, pass = blockIdx.x & 3, slot = blockIdx.x >> 2;
  const int count = *q.count;
  uint32_t* table = dyn;
  uint32_t* cache = dyn + table_size;
  const HeapRef heap = heap_ref(cache + SOLVER_CACHE_NODES * SOLVER_NODE_WORDS, nullptr, 0x7fffffff);  // all in shared memory
  __shared__ SState root_s;
  uint32_t* nodes = node_pool + ((size_t)slot * 4 + pass) * nodes_per_pass * SOLVER_NODE_WORDS;
  const int W = cfg.width, H = cfg.height, cells = W * H;
  // pass order: sokoban BFS, A*(1), A*(.5), A*(0) (sokoban_prob.py:110-122); ddave / mdungeon A*(1), A*(.5), A*(0), BFS
  const int b = (GAME == GAME_SOKOBAN) ? ((pass == 0) ? -1 : (pass == 1) ? 2 : (pass == 2) ? 1 : 0)
                                       : ((pass == 0) ? 2 : (pass == 1) ? 1 : (pass == 2) ? 0 : -1);
  for (int item = slot; item < count; item += slots) {
    const int packed = q.items[item], e = packed & 0x0fffffff, mode = packed >> 28;
    __syncwarp();
    for (int i = lane; i < cells; i += 32) L.tiles[i] = maps[(size_t)e * cells + i];
    for (int i = lane; i < table_size; i += 32) table[i] = 0u;
    __syncwarp();
    __shared__ int res_s[4];
    __shared__ int exhausted_s;
    int* res = res_s;
    if (lane == 0) {
      SState root;
      level_init<GAME>(L, root, W, H);
      root_s = root;
      res[0] = 0; res[1] = 0; res[2] = 0; res[3] = 0;
      exhausted_s = 0;
      if (L.overflow) atomicExch(q.status, 1);
    }
    __syncwarp();
    if (!L.overflow) {
      const SState root = root_s;
      if (b < 0)
        search_bfs_batched<GAME>(L, root, cfg.solver_power, nodes, table, table_size - 1, q.best_win + item, pass, res, &exhausted_s, lane);
      else
        search_pass<GAME>(L, root, b, cfg.solver_power, nodes, cache, heap, table, table_size - 1, q.best_win + item, pass, res, &exhausted_s, lane);
    }
    __syncwarp();
    if (lane == 0) {
      if (res[0] == 1) atomicMin(q.best_win + item, pass);
      // A pass whose queue ran empty without a win has popped every reachable state exactly once.  Where the visited
      // key IS the state (sokoban, mdungeon) every other pass would pop the same states in the same number of
      // iterations and also fail, and its best node has the same minimum heuristic -> the other passes are cancelled
      // (best_win = -1 - pass).  sokoban reports the heuristic only, so any exhausted pass serves; mdungeon reports
      // the counters of the BFS pass's best node, so only its BFS pass may cancel.  ddave's key omits airTime / jumps:
      // no rule.
      if (res[0] == 0 && exhausted_s && (GAME == GAME_SOKOBAN || (GAME == GAME_MDUNGEON && b < 0)))
        atomicMin(q.best_win + item, -1 - pass);
      int32_t* r = q.results + ((size_t)item * 4 + pass) * 4;
      r[0] = res[0]; r[1] = res[1]; r[2] = res[2]; r[3] = res[3];
      __threadfence();
      if (atomicAdd(q.pass_done + item, 1) == 3) {  // last pass to finish merges in the reference's order
        __threadfence();
        const volatile int32_t* rr = q.results + (size_t)item * 16;
        int sel = 3;
        const int bw = *(const volatile int32_t*)(q.best_win + item);
        if (bw < 0) sel = -1 - bw;  // exhausted pass: nobody can win
        else for (int p = 0; p < 4; p++) if (rr[p * 4] == 1) { sel = p; break; }
        const int won = (rr[sel * 4] == 1), depth = rr[sel * 4 + 1], h = rr[sel * 4 + 2];
        const uint32_t misc = (uint32_t)rr[sel * 4 + 3];
        int32_t* st = stats + (size_t)e * PCGRL_MAX_STATS;
        int32_t* ss = (mode == SOLVE_FOR_RESET && start_stats) ? start_stats + (size_t)e * PCGRL_MAX_STATS : nullptr;
        const int dist_win = won ? 0 : h, sol_len = won ? depth : 0;
        if (GAME == GAME_SOKOBAN) {  // sokoban_prob.py:110-122,143-144
          st[4] = dist_win; st[5] = sol_len;
          if (ss) { ss[4] = dist_win; ss[5] = sol_len; }
        } else if (GAME == GAME_DDAVE) {  // ddave_prob.py:122-135,164-168
          const int jumps = (int)(misc >> 16), col = (int)((misc >> 8) & 0xffu);
          st[9] = dist_win; st[10] = sol_len; st[7] = jumps; st[8] = col;
          if (ss) { ss[9] = dist_win; ss[10] = sol_len; ss[7] = jumps; ss[8] = col; }
        } else {  // mdungeon_prob.py:125-138,166-170
          const int pot = (int)(misc & 0xffu), tre = (int)((misc >> 8) & 0xffu), ene = (int)((misc >> 16) & 0xffu);
          st[9] = dist_win; st[10] = sol_len; st[6] = pot; st[7] = tre; st[8] = ene;
          if (ss) { ss[9] = dist_win; ss[10] = sol_len; ss[6] = pot; ss[7] = tre; ss[8] = ene; }
        }
      }
    }
    __syncwarp();
  }
}

// Shared-memory words of one search: visited table (power of two >= 1.5 * power) + node ring + binary heap.
// The open list grows by at most 3 per iteration (one pop, at most four pushes), so 3 * power + 8 entries suffice;
// with the default power of 5000 one search needs 94.5 KB and TWO searches fit one SM.
static inline size_t solver_arena_words(const pcgrl_config* cfg, int* table_size_out) {
  int table_size = 1024;
  while (table_size < cfg->solver_power + cfg->solver_power / 2) table_size <<= 1;
  const size_t heap_words = (size_t)3 * cfg->solver_power + 8;
  if (table_size_out) *table_size_out = table_size;
  return (size_t)table_size + SOLVER_CACHE_NODES * SOLVER_NODE_WORDS + heap_words;
}

// k_rollout_async keeps only the first ASYNC_HEAP_FAST open-list entries in shared memory
static inline size_t solver_async_arena_words(const pcgrl_config* cfg, int* table_size_out, int* heap_fast_out) {
  int table_size;
  solver_arena_words(cfg, &table_size);
  size_t heap_words = ((size_t)3 * cfg->solver_power + 9) & ~(size_t)1;  // even: a child pair never straddles the regions
  if (ASYNC_HEAP_FAST > 0 && heap_words > ASYNC_HEAP_FAST) heap_words = ASYNC_HEAP_FAST;
  if (table_size_out) *table_size_out = table_size;
  if (heap_fast_out) *heap_fast_out = (int)heap_words;
  return (size_t)table_size + SOLVER_CACHE_NODES * SOLVER_NODE_WORDS + heap_words;
}

// opt in to the large dynamic shared memory of k_solve (once per power setting; call before multi-threaded enqueue)
template <int PROB>
static inline void solver_prepare(const pcgrl_config* cfg) {
  if constexpr (GameOf<PROB>::GAME >= 0) {
    int table_size;
    const size_t smem = solver_arena_words(cfg, &table_size) * sizeof(uint32_t);
    static size_t configured[SOLVER_MAX_DEVICES][PCGRL_NUM_PROBLEMS] = {};  // the attribute is per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= SOLVER_MAX_DEVICES || configured[dev][PROB] < smem) {
      cudaFuncSetAttribute(k_solve<PROB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (dev >= 0 && dev < SOLVER_MAX_DEVICES) configured[dev][PROB] = smem;
    }
  }
}

template <int PROB>
static inline void solver_launch(const pcgrl_config* cfg, int32_t* stats, int32_t* start_stats, const uint8_t* maps,
                                 SolverQueue q, void* scratch, int n, cudaStream_t s, int max_slots = SOLVER_MAX_SLOTS) {
  if constexpr (GameOf<PROB>::GAME >= 0) {
  if (!q.count) return;
  const SolverLayout lay = solver_layout(cfg, n, max_slots);
  int table_size;
  const size_t smem = solver_arena_words(cfg, &table_size) * sizeof(uint32_t);
  solver_prepare<PROB>(cfg);
  uint32_t* pool = (uint32_t*)((char*)scratch + lay.nodes_off);
  k_solve<PROB><<<4 * lay.slots, 32, smem, s>>>(*cfg, stats, start_stats, maps, q, pool, lay.nodes_per_pass, lay.slots, table_size);
  }
}

}  // namespace pcgrl

// pcgrl_smb.cuh -- the smb problem: PcgrlEnv.reset / step for 114 x 14 Super Mario levels.
//
// Reference: gym_pcgrl/envs/probs/smb_prob.py:9-185 (_run_game :95-124, get_stats :126-148, get_reward :150-172,
// get_episode_over :174-175), probs/smb/engine.py (State :131-286, AStarAgent :105-129), helper.py:37-62
// (get_floor_dist), :74-103 (get_type_grouping), :115-133 (get_changes), :310-352 (gen_random_map), the six
// Representation.update methods (reps/*.py) and pcgrl_env.py:66-76,129-150.
//
// The 114 x 14 map does not fit the one-row-per-lane bitboards of the other problems, and its cost is not the map
// scans but the always-on A* play-through (two passes of up to 10 000 iterations on every step that edits the map).
// Design:
//   * ONE WARP PER ENV, persistent CTAs pulling env indices from a global counter.  The warp stages the env's byte map
//     in shared memory; everything sequential (Representation.update, the scans, the search loop) is scalar code run by
//     lane 0, everything wide (map load, visited-set clears, building the solid bit rows) is done by the whole warp.
//   * The A* open list is CPython's binary heap reproduced operation by operation (the result depends on its pop order
//     under ties) on SELF-CONTAINED 64-bit entries: x, y, airTime, depth, jumps, last jump x and widest jump gap are
//     packed in the entry itself, so there is no node store and no indirection; priority = (exit - x) + balance * depth
//     is recomputed from the fields.  The first `fast_cap` entries (the top levels, touched by every pop) live in the
//     warp's shared memory, the tail in a per-warp slice of HBM scratch.  The visited set is a 16 Kbit bitmap over
//     (x, y, airTime) == State.getKey, also in shared memory.
//   * SEARCH SKIPPING (exact): the search is a deterministic function of the solidity of the level cells it reads.  Every
//     read is recorded in a per-env "touched" bitmap; a later edit that does not change the solidity of a touched cell
//     cannot change the outcome, so jumps / jumps-dist / dist-win are carried over and the two A* passes are skipped.
//     Random edits mostly fall outside the region the agent can reach, and half of the tile pairs have equal solidity.
//   * All scalar pieces are `__host__ __device__`: the same functions are the host twins pcgrl_*_cpu (plumbing without a
//     GPU) and are checked against the reference's golden vectors on the CPU (tests/test_smb_device_code_on_host.py).
//
// Limits (pcgrl_config_validate): width <= 122, 3 <= height <= 16, solver_power <= 16000 (depth / jumps fit 14 bits).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "../../include/pcgrl_b200.h"

#ifdef __CUDACC__
#define SMB_HD __host__ __device__ __forceinline__
#define SMB_HDN __host__ __device__
#else
#define SMB_HD static inline
#define SMB_HDN static
#endif

namespace pcgrl_smb {

typedef unsigned long long u64;

enum { T_EMPTY = 0, T_SOLID, T_ENEMY, T_BRICK, T_QUESTION, T_COIN, T_TUBE, NUM_TILES = 7 };
enum { MAX_W = PCGRL_SMB_MAX_W, MAX_H = PCGRL_SMB_MAX_H, ROW_WORDS = 4, LEVEL_WORDS = MAX_H * ROW_WORDS,
       VISITED_WORDS = 512 /* 6 airTimes x 21 rows (-5 <= y <= 15) x 128 columns, one bit each (504 words) */, MAX_POWER = 16000 };
#define SMB_SOLID_TYPES ((1u << T_SOLID) | (1u << T_BRICK) | (1u << T_QUESTION) | (1u << T_TUBE)) /* smb_prob.py:96 " # ## #" */

// ------------------------------------------------------------------------------------------------
// MT19937 + numpy legacy RandomState (scalar; state = 624 key words + position, as everywhere in this repo)
// ------------------------------------------------------------------------------------------------
SMB_HDN void mt_twist(uint32_t* k) {
  const uint32_t UP = 0x80000000u, LO = 0x7fffffffu, A = 0x9908b0dfu;
  int i;
  uint32_t y;
  for (i = 0; i < 624 - 397; i++) { y = (k[i] & UP) | (k[i + 1] & LO); k[i] = k[i + 397] ^ (y >> 1) ^ ((y & 1u) ? A : 0u); }
  for (; i < 623; i++) { y = (k[i] & UP) | (k[i + 1] & LO); k[i] = k[i - 227] ^ (y >> 1) ^ ((y & 1u) ? A : 0u); }
  y = (k[623] & UP) | (k[0] & LO);
  k[623] = k[396] ^ (y >> 1) ^ ((y & 1u) ? A : 0u);
}
SMB_HD uint32_t mt_u32(uint32_t* st) {
  uint32_t pos = st[624];
  if (pos >= 624) { mt_twist(st); pos = 0; }
  uint32_t y = st[pos];
  st[624] = pos + 1;
  y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
  return y;
}
SMB_HD double mt_double(uint32_t* st) {  // random_sample()
  const uint32_t a = mt_u32(st) >> 5, b = mt_u32(st) >> 6;
  return ((double)a * 67108864.0 + (double)b) / 9007199254740992.0;
}
SMB_HD int mt_randint(uint32_t* st, int n) {  // RandomState.randint(n): masked rejection, no draw when n == 1
  const uint32_t rng = (uint32_t)(n - 1);
  if (rng == 0) return 0;
  uint32_t mask = rng, v;
  mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
  do { v = mt_u32(st) & mask; } while (v > rng);
  return (int)v;
}

// helper.py:310-312 gen_random_map + :343-352 get_int_prob + RandomState.choice(p) on the per-env probabilities
SMB_HDN void gen_random_map(uint32_t* rng, const double* tile_prob, int T, int cells, uint8_t* map, uint8_t* map2, uint8_t* map3) {
  double p[PCGRL_MAX_TILES], cdf[PCGRL_MAX_TILES], total = 0.0, acc = 0.0;
  for (int t = 0; t < T; t++) total += tile_prob[t];
  for (int t = 0; t < T; t++) p[t] = tile_prob[t] / total;
  for (int t = 0; t < T; t++) { acc += p[t]; cdf[t] = acc; }
  for (int t = 0; t < T; t++) cdf[t] /= cdf[T - 1];
  for (int i = 0; i < cells; i++) {
    const double u = mt_double(rng);
    int k = 0;
    while (k < T && cdf[k] <= u) k++;  // searchsorted(side='right')
    map[i] = (uint8_t)k;
    if (map2) map2[i] = (uint8_t)k;
    if (map3) map3[i] = (uint8_t)k;
  }
}

// ------------------------------------------------------------------------------------------------
// level = solid bit rows (4 words per row, level width = W + 6) + the bitmap of cells the search has read
// ------------------------------------------------------------------------------------------------
struct Level {
  int width, height, exit_x;
  const uint32_t* solid;  // [height][ROW_WORDS]
  uint32_t* touched;      // [height][ROW_WORDS]: every solidity read is recorded here
};

SMB_HD bool solid_at(const Level& L, int x, int y) {
  const int w = y * ROW_WORDS + (x >> 5);
  const uint32_t bit = 1u << (x & 31);
  L.touched[w] |= bit;
  return (L.solid[w] & bit) != 0u;
}
SMB_HD bool movable(const Level& L, int x, int y) {  // engine.py:203-206 checkMovableLocation
  if (y < 0) return true;
  return !(x < 0 || x >= L.width || y >= L.height || solid_at(L, x, y));
}

// one row of the runnable level (smb_prob.py:97-115): "   " / " @ " / "###" + row + " | " / " # " / "###"
SMB_HD void build_solid_row(const uint8_t* m, int w, int h, int y, uint32_t* solid) {
  uint32_t r[ROW_WORDS] = {0u, 0u, 0u, 0u};
  const bool floor_rows = y > h - 3;
  for (int x = 0; x < w + 6; x++) {
    bool s;
    if (x < 3 || x >= 3 + w) s = floor_rows || (y == h - 3 && x == 3 + w + 1);
    else s = (SMB_SOLID_TYPES >> m[y * w + x - 3]) & 1u;
    if (s) r[x >> 5] |= 1u << (x & 31);
  }
  for (int k = 0; k < ROW_WORDS; k++) solid[y * ROW_WORDS + k] = r[k];
}

// ------------------------------------------------------------------------------------------------
// search node == heap entry (64 bits):  x:7 | y+8:5 | airTime:3 | depth:14 | jumps:14 | last jump x:7 | widest gap:7
// (engine.py:159 player dict; jump_locs folded into (jumps, last jump x, widest gap), smb_prob.py:140-146)
// ------------------------------------------------------------------------------------------------
struct State { int x, y, air, jumps, last_jump_x, max_gap; };

SMB_HD u64 pack(const State& s, int depth) {
  return (u64)(uint32_t)s.x | ((u64)(uint32_t)(s.y + 8) << 7) | ((u64)(uint32_t)s.air << 12) | ((u64)(uint32_t)depth << 15) |
         ((u64)(uint32_t)s.jumps << 29) | ((u64)(uint32_t)s.last_jump_x << 43) | ((u64)(uint32_t)s.max_gap << 50);
}
SMB_HD int e_x(u64 e) { return (int)(e & 127u); }
SMB_HD int e_y(u64 e) { return (int)((e >> 7) & 31u) - 8; }
SMB_HD int e_air(u64 e) { return (int)((e >> 12) & 7u); }
SMB_HD int e_depth(u64 e) { return (int)((e >> 15) & 0x3fffu); }
// (x, y, airTime) == State.getKey (engine.py:248-249) as a dense bit index.  A state that reaches the visited test has
// 0 <= airTime <= 5 and -5 <= y < height <= 16: a jump needs ground (y >= -1) and rises at most four rows (:224-236).
SMB_HD int e_key(u64 e) {
  int row = (int)((e >> 7) & 31u) - 3;
  row = row < 0 ? 0 : (row > 20 ? 20 : row);
  return (((int)((e >> 12) & 7u) * 21 + row) << 7) | (int)(e & 127u);
}
SMB_HD State unpack(u64 e) {
  State s;
  s.x = e_x(e); s.y = e_y(e); s.air = e_air(e);
  s.jumps = (int)((e >> 29) & 0x3fffu); s.last_jump_x = (int)((e >> 43) & 127u); s.max_gap = (int)((e >> 50) & 127u);
  return s;
}

SMB_HD void st_update(const Level& L, State& s, int dir_x, int dir_y) {  // engine.py:208-246
  if (s.x >= L.exit_x || s.y >= L.height) return;  // checkOver
  bool ground = false;
  if (s.y < L.height - 1 && s.y >= -1) ground = solid_at(L, s.x, s.y + 1);
  int nx = s.x, ny = s.y;
  if (dir_x != 0 && movable(L, nx + dir_x, ny)) nx += dir_x;
  if (dir_y < 0) {
    if (ground && movable(L, nx, ny - 1)) {
      s.air = 5;
      s.jumps += 1;
      if (s.x - s.last_jump_x > s.max_gap) s.max_gap = s.x - s.last_jump_x;  // jump_locs.append((x, y)) of the OLD position
      s.last_jump_x = s.x;
    }
  } else if (s.air > 0) {
    s.air = 1;
  }
  if (s.air > 1) {
    s.air -= 1;
    if (movable(L, nx, ny - 1)) ny -= 1;
    else s.air = 1;
  } else if (s.air == 1) {
    s.air = 0;
  } else if (movable(L, nx, ny + 1)) {
    ny += 1;
  }
  s.x = nx;
  s.y = ny;
}

// The open list: element i of CPython's list sits in storage slot i + 1 (a 1-based heap: parent s / 2, children 2s and
// 2s + 1), so that a child pair is one aligned 16-byte couple; slots below fast_cap are in `fast` (shared memory on the
// device), the rest in `slow`.  The sift loops work on local copies of the three fields (HeapRef) -- the accessors were
// 47 % of the kernel's instructions while they re-read the struct through a reference (profiles/r02_summary.md).
struct Heap {
  u64* fast;
  u64* slow;
  int fast_cap;  // even
};
struct HeapRef {
  u64* fast;
  u64* slow_biased;  // slow - fast_cap: slot s >= fast_cap lives at slow_biased[s]
  int cap;
};
SMB_HD HeapRef heap_ref(const Heap& h) { HeapRef r; r.fast = h.fast; r.slow_biased = h.slow - h.fast_cap; r.cap = h.fast_cap; return r; }
SMB_HD u64* slot(const HeapRef& h, int s) { return (s < h.cap ? h.fast : h.slow_biased) + s; }
// priority of Node.__lt__ (engine.py:52-53): (exit - x) + balance * depth -- both fields sit in the low word of the entry
SMB_HD int e_prio(u64 e, int exit_x, int balance) {
  const uint32_t lo = (uint32_t)e;
  return exit_x - (int)(lo & 127u) + balance * (int)((lo >> 15) & 0x3fffu);
}

// Lib/heapq.py: heappush -> _siftdown(heap, 0, len-1); heappop -> _siftup(heap, 0) (which ends in a _siftdown)
SMB_HD void heap_siftdown(const HeapRef& h, int s, u64 item, int exit_x, int bal) {  // s = slot of the hole (1-based)
  const int pi = e_prio(item, exit_x, bal);
  while (s > 1) {
    const int ps = s >> 1;
    const u64 parent = *slot(h, ps);
    if (pi < e_prio(parent, exit_x, bal)) { *slot(h, s) = parent; s = ps; continue; }
    break;
  }
  *slot(h, s) = item;
}
SMB_HD u64 heap_pop(const HeapRef& h, int& n, int exit_x, int bal) {  // n = number of entries (slots 1..n)
  const u64 last = *slot(h, n);
  n--;
  if (n == 0) return last;
  const u64 ret = h.fast[1];
  int s = 1, c = 2;
  while (c <= n) {
    const u64* pair = slot(h, c);  // c is even and fast_cap is even: both children are in the same region
    u64 child = pair[0];
    if (c < n) {
      const u64 right = pair[1];
      if (!(e_prio(child, exit_x, bal) < e_prio(right, exit_x, bal))) { c++; child = right; }
    }
    *slot(h, s) = child;
    s = c;
    c = 2 * s;
  }
  heap_siftdown(h, s, last, exit_x, bal);
  return ret;
}

// AStarAgent.getSolution (engine.py:105-129) on a CLEARED visited bitmap; returns true on a win; `result` = the winning
// node, else the best node (lowest heuristic, then lowest cost).
SMB_HDN bool astar_core(const Level& level, u64 root, int balance, int max_iter, const Heap& hp, uint32_t* visited, u64& result,
                        int& iterations_out) {
  const Level L = level;  // local copies: the loop must not re-read the structs through the references
  const HeapRef h = heap_ref(hp);
  const int exit_x = L.exit_x, height = L.height;
  int nheap = 1, iterations = 0, best_h = 0, best_depth = 0;
  bool have_best = false;
  u64 best = root;
  h.fast[1] = root;
  while ((iterations < max_iter || max_iter <= 0) && nheap > 0) {
    iterations++;
    const u64 cur = heap_pop(h, nheap, exit_x, balance);
    if (e_y(cur) >= height) continue;                                              // checkLose
    if (e_x(cur) >= exit_x) { result = cur; iterations_out = iterations; return true; }  // checkWin
    const int key = e_key(cur);
    const uint32_t bit = 1u << (key & 31);
    const uint32_t seen = visited[key >> 5];
    if (seen & bit) continue;
    const int hh = exit_x - e_x(cur), depth = e_depth(cur);
    if (!have_best || hh < best_h || (hh == best_h && depth < best_depth)) { best = cur; best_h = hh; best_depth = depth; have_best = true; }
    visited[key >> 5] = seen | bit;
    const State cs = unpack(cur);
#pragma unroll
    for (int d = 0; d < 4; d++) {  // engine.py:3 directions (0,0) (1,0) (0,-1) (1,-1)
      State c = cs;
      st_update(L, c, d & 1, (d & 2) ? -1 : 0);
      nheap++;
      heap_siftdown(h, nheap, pack(c, depth + 1), exit_x, balance);
    }
  }
  result = best;
  iterations_out = iterations;
  return false;
}

SMB_HD u64 root_state(int h) { const State s0 = {1, h - 3, 0, 0, 0, 0}; return pack(s0, 0); }

// play statistics of the selected node -> st[5..7] (smb_prob.py:136-147)
SMB_HD void play_stats(u64 node, bool won, int w, int exit_x, int32_t* st) {
  const State s = unpack(node);
  st[5] = s.jumps;
  st[6] = (w - s.last_jump_x > s.max_gap) ? (w - s.last_jump_x) : s.max_gap;
  st[7] = won ? 0 : (exit_x - s.x);
}

// ------------------------------------------------------------------------------------------------
// helper.py scans on the uint8 map -> st[0..4] = dist-floor, disjoint-tubes, enemies, empty, noise
// ------------------------------------------------------------------------------------------------
SMB_HDN void scan_stats(const uint8_t* m, int w, int h, int32_t* st) {
  const unsigned floor_types = (1u << T_SOLID) | (1u << T_BRICK) | (1u << T_QUESTION);  // "tube_left/right" never occur in the map
  int dist_floor = 0, tubes = 0, enemies = 0, empty = 0, noise = 0;
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      const int t = m[y * w + x];
      empty += (t == T_EMPTY);
      if (t == T_ENEMY) {  // helper.py:37-62 get_floor_dist(map, ["enemy"], floor types)
        enemies++;
        int dist = h - 1;
        for (int dy = 0; y + dy < h; dy++)
          if ((floor_types >> m[(y + dy) * w + x]) & 1u) { dist = dy - 1; break; }
        dist_floor += dist;
      }
      if (t == T_TUBE) {  // get_type_grouping(map, ["tube"], [(-1,0),(1,0)], 1, 1): helper.py:74-103
        const int nb = ((x >= 1 && m[y * w + x - 1] == T_TUBE) ? 1 : 0) + ((x + 1 < w && m[y * w + x + 1] == T_TUBE) ? 1 : 0);
        tubes += (nb == 1);
      }
      if (x >= 1 && m[y * w + x - 1] != t) noise++;    // get_changes(map, False): helper.py:115-133
      if (y >= 1 && m[(y - 1) * w + x] != t) noise++;  // get_changes(map, True)
    }
  st[0] = dist_floor; st[1] = tubes; st[2] = enemies; st[3] = empty; st[4] = noise;
}

// helper.py:366-376 get_range_reward
SMB_HD double range_reward(double nv, double ov, double low, double high) {
  if (nv >= low && nv <= high && ov >= low && ov <= high) return 0.0;
  if (ov <= high && nv <= high) return fmin(nv, low) - fmin(ov, low);
  if (ov >= low && nv >= low) return fmax(ov, high) - fmax(nv, high);
  if (nv > high && ov < low) return high - nv + ov - low;
  if (nv < low && ov > high) return high - ov + nv - low;
  return 0.0;
}
// smb_prob.py:150-172, terms summed left to right
SMB_HD double get_reward(const pcgrl_config& cfg, const int32_t* n, const int32_t* o) {
  const double* w = cfg.reward_weight;
  const int32_t* ip = cfg.iparam;
  const double INF = HUGE_VAL;
  return range_reward(n[0], o[0], 0, 0) * w[0] + range_reward(n[1], o[1], 0, 0) * w[1] +
         range_reward(n[2], o[2], ip[1], ip[2]) * w[2] + range_reward(n[3], o[3], ip[0], INF) * w[3] +
         range_reward(n[4], o[4], 0, 0) * w[4] + range_reward(n[5], o[5], ip[3], INF) * w[5] +
         range_reward(n[6], o[6], 0, 0) * w[6] + range_reward(n[7], o[7], 0, 0) * w[7];
}
SMB_HD bool episode_over(const int32_t* n) { return n[7] <= 0; }  // smb_prob.py:174-175

// ------------------------------------------------------------------------------------------------
// Representation.update on the byte map (the six representations; reps/*.py, cited in include/pcgrl_b200.h)
// ------------------------------------------------------------------------------------------------
struct Edit {
  int change;            // number of cells whose tile changed
  int hx, hy;            // heat-map cell (pcgrl_env.py:137)
  int cell, tile;        // single-cell edit (delta transport)
  bool multi;            // a 3x3 stamp ran
  bool solidity_touched; // some changed cell switched between solid / non-solid AND the last search had read it
};

SMB_HD void write_cell(const pcgrl_config& cfg, uint8_t* m, uint8_t* m2, const uint32_t* touched, int x, int y, int t, Edit& ed) {
  const int W = cfg.width, old = m[y * W + x];
  if (old == t) return;
  ed.change++;
  m[y * W + x] = (uint8_t)t;
  if (m2) m2[y * W + x] = (uint8_t)t;
  if ((((SMB_SOLID_TYPES >> old) ^ (SMB_SOLID_TYPES >> t)) & 1u) && ((touched[y * ROW_WORDS + ((x + 3) >> 5)] >> ((x + 3) & 31)) & 1u))
    ed.solidity_touched = true;
}
SMB_HD void turtle_move(const pcgrl_config& cfg, int a, int& x, int& y) {  // turtle_rep.py:101-125
  const int W = cfg.width, H = cfg.height;
  const bool warp = (cfg.flags & PCGRL_FLAG_WARP) != 0;
  x += (a == 0) ? -1 : (a == 1) ? 1 : 0;
  if (x < 0) x = warp ? x + W : 0;
  if (x >= W) x = warp ? x - W : W - 1;
  y += (a == 2) ? -1 : (a == 3) ? 1 : 0;
  if (y < 0) y = warp ? y + H : 0;
  if (y >= H) y = warp ? y - H : H - 1;
}
// m = working copy of the map (shared memory on the device), m2 = the env's map in HBM (or NULL on the host)
SMB_HDN Edit apply_action(const pcgrl_config& cfg, const int32_t* act, uint8_t* m, uint8_t* m2, const uint32_t* touched,
                          uint32_t* rng, int& x, int& y) {
  const int W = cfg.width, H = cfg.height, rep = cfg.representation;
  Edit ed;
  ed.change = 0; ed.multi = false; ed.solidity_touched = false; ed.tile = 0;
  int wx = x, wy = y, newt = -1, stamp = -2;  // stamp: -2 none, -1 per-cell values (narrowmulti), >= 0 one value
  if (rep == PCGRL_REP_NARROW) {
    if (act[0] > 0) newt = (act[0] - 1) & 7;
  } else if (rep == PCGRL_REP_TURTLE) {
    if (act[0] >= 4) newt = (act[0] - 4) & 7;
    else if (act[0] >= 0) turtle_move(cfg, act[0], x, y);
    wx = x; wy = y;
  } else if (rep == PCGRL_REP_WIDE) {
    wx = act[0] < 0 ? 0 : (act[0] >= W ? W - 1 : act[0]);
    wy = act[1] < 0 ? 0 : (act[1] >= H ? H - 1 : act[1]);
    newt = act[2] & 7;
  } else if (rep == PCGRL_REP_NARROWCAST) {
    if (act[0] == 1) newt = act[1] & 7;
    else if (act[0] == 2) stamp = act[1] & 7;
  } else if (rep == PCGRL_REP_NARROWMULTI) {
    stamp = -1;
  } else {  // PCGRL_REP_TURTLECAST
    if (act[0] >= 0 && act[0] < 4) turtle_move(cfg, act[0], x, y);
    else if (act[0] == 4) newt = act[1] & 7;
    else if (act[0] == 5) stamp = act[1] & 7;
    wx = x; wy = y;
  }
  if (stamp != -2) {  // 3x3 block centred on the cursor, clipped to the map (narrow_cast_rep.py:43-48 etc.)
    ed.multi = true;
    for (int k = 0; k < 9; k++) {
      const int cx = x + (k % 3) - 1, cy = y + (k / 3) - 1;
      int t = stamp;
      if (stamp == -1) t = (act[k] > 0) ? ((act[k] - 1) & 7) : -1;
      if (cx >= 0 && cx < W && cy >= 0 && cy < H && t >= 0) write_cell(cfg, m, m2, touched, cx, cy, t, ed);
    }
  }
  if (newt >= 0) {
    write_cell(cfg, m, m2, touched, wx, wy, newt, ed);
    ed.tile = newt;
  }
  ed.cell = wy * W + wx;
  if (rep == PCGRL_REP_NARROW || rep == PCGRL_REP_NARROWCAST || rep == PCGRL_REP_NARROWMULTI) {
    if (cfg.flags & PCGRL_FLAG_RANDOM_TILE) {  // narrow_rep.py:104-106
      x = mt_randint(rng, W);
      y = mt_randint(rng, H);
    } else {                                   // :107-113
      x += 1;
      if (x >= W) { x = 0; y += 1; if (y >= H) y = 0; }
    }
    ed.hx = x; ed.hy = y;
  } else if (rep == PCGRL_REP_WIDE) {
    ed.hx = wx; ed.hy = wy;
  } else {
    ed.hx = x; ed.hy = y;
  }
  return ed;
}

SMB_HD size_t heap_entries(int power) { return ((size_t)3 * power + 16) & ~(size_t)1; }  // one pop, <= four pushes per iteration

// _run_game + the play part of get_stats on the host: both passes, everything in `slow` memory
SMB_HDN void run_game_scalar(const uint8_t* m, int w, int h, int power, uint32_t* solid, uint32_t* touched, uint32_t* visited,
                             u64* heap_mem, int32_t* st, long* iterations) {
  Level L;
  L.width = w + 6; L.height = h; L.exit_x = w + 4; L.solid = solid; L.touched = touched;
  for (int i = 0; i < LEVEL_WORDS; i++) touched[i] = 0u;
  for (int y = 0; y < h; y++) build_solid_row(m, w, h, y, solid);
  Heap hp;
  hp.fast = heap_mem; hp.slow = heap_mem; hp.fast_cap = 1 << 30;
  u64 node;
  int it1 = 0, it2 = 0;
  for (int i = 0; i < VISITED_WORDS; i++) visited[i] = 0u;
  bool won = astar_core(L, root_state(h), 1, power, hp, visited, node, it1);
  if (!won) {
    for (int i = 0; i < VISITED_WORDS; i++) visited[i] = 0u;
    won = astar_core(L, root_state(h), 0, power, hp, visited, node, it2);
  }
  play_stats(node, won, w, L.exit_x, st);
  if (iterations) *iterations += it1 + it2;
}

}  // namespace pcgrl_smb

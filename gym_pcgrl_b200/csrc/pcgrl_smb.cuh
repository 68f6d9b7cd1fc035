// pcgrl_smb.cuh -- SMBProblem.get_stats on the device (first piece: the stand-alone operator).
//
// Reference: gym_pcgrl/envs/probs/smb_prob.py:95-148 (_run_game, get_stats), probs/smb/engine.py (State, AStarAgent),
// helper.py:37-62 (get_floor_dist), :74-103 (get_type_grouping), :115-133 (get_changes).
//
// The 114 x 14 map does not fit the one-row-per-lane bitboards of the other problems, and its statistics are plain
// scans plus an always-on A* play-through whose key space is tiny (x, y, airTime): so the mapping is ONE THREAD PER
// MAP.  A thread keeps the level's solid cells as 4 x 32-bit words per row in shared memory, and its open list
// (CPython heapq order on packed entries priority << 16 | node, as in pcgrl_solver.cuh), node store and visited
// bitmap in a private slice of caller-owned scratch in HBM; 2048 searches are in flight per launch, the maps are
// taken grid-stride.  Everything below is scalar `__host__ __device__` code: tests/test_smb_device_code_on_host.py
// compiles this header with g++ and checks the very same functions against the reference's golden vectors on the
// CPU; the GPU test only has to confirm the launch plumbing.
//
// Limits: width <= 122, height <= 16, solver_power <= 16000 (node index fits 16 bits).
#pragma once
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define SMB_HD __host__ __device__ __forceinline__
#else
#define SMB_HD static inline
#endif

namespace pcgrl_smb {

enum { T_EMPTY = 0, T_SOLID, T_ENEMY, T_BRICK, T_QUESTION, T_COIN, T_TUBE };
enum { MAX_W = 122, MAX_H = 16, ROW_WORDS = 4, VISITED_WORDS = (MAX_H + 16) * 128 * 8 / 32, PRIO_BIAS = 256 };

struct Level {
  int width, height, exit_x;
  const uint32_t* solid;  // [height][ROW_WORDS]
};

struct State {  // engine.py:159,187-195: player dict; jump_locs folded into (jumps, last_jump_x, max_gap)
  int x, y, air, jumps, last_jump_x, max_gap;
};

struct Workspace {  // private slice of one search
  uint32_t* heap;     // [3 * power + 8]  priority << 16 | node
  uint32_t* nodes;    // [4 * power + 8][4]  x | (y + 8) << 8 | air << 16,  jumps | last_jump_x << 16,  max_gap | depth << 16,  -
  uint32_t* visited;  // [VISITED_WORDS] bitmap over (y + 8, x, air)
};

SMB_HD size_t heap_words(int power) { return (size_t)3 * power + 8; }
SMB_HD size_t node_words(int power) { return ((size_t)4 * power + 8) * 4; }
SMB_HD size_t workspace_words(int power) { return ((heap_words(power) + node_words(power) + VISITED_WORDS + 3) / 4) * 4; }

SMB_HD bool solid_at(const Level& L, int x, int y) { return (L.solid[y * ROW_WORDS + (x >> 5)] >> (x & 31)) & 1u; }
SMB_HD bool movable(const Level& L, int x, int y) {  // engine.py:203-206
  if (y < 0) return true;
  return !(x < 0 || x >= L.width || y >= L.height || solid_at(L, x, y));
}
SMB_HD bool st_win(const Level& L, const State& s) { return s.x >= L.exit_x; }
SMB_HD bool st_lose(const Level& L, const State& s) { return s.y >= L.height; }

SMB_HD void st_update(const Level& L, State& s, int dir_x, int dir_y) {  // engine.py:208-246
  if (st_win(L, s) || st_lose(L, s)) return;
  dir_y = (dir_y < 0) ? -1 : 0;
  bool ground = false;
  if (s.y < L.height - 1 && s.y >= -1) ground = solid_at(L, s.x, s.y + 1);
  int nx = s.x, ny = s.y;
  if (dir_x != 0 && movable(L, nx + dir_x, ny)) nx += dir_x;
  if (dir_y == -1) {
    if (ground && movable(L, nx, ny - 1)) {
      s.air = 5;
      s.jumps += 1;
      if (s.x - s.last_jump_x > s.max_gap) s.max_gap = s.x - s.last_jump_x;  // smb_prob.py:141-145 on (old x, y)
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

SMB_HD void node_store(uint32_t* nodes, int i, const State& s, int depth) {
  uint32_t* p = nodes + (size_t)i * 4;
  p[0] = (uint32_t)s.x | ((uint32_t)(s.y + 8) << 8) | ((uint32_t)s.air << 16);
  p[1] = (uint32_t)s.jumps | ((uint32_t)s.last_jump_x << 16);
  p[2] = (uint32_t)s.max_gap | ((uint32_t)depth << 16);
}
SMB_HD void node_load(const uint32_t* nodes, int i, State& s, int& depth) {
  const uint32_t* p = nodes + (size_t)i * 4;
  s.x = (int)(p[0] & 0xffu); s.y = (int)((p[0] >> 8) & 0xffu) - 8; s.air = (int)((p[0] >> 16) & 0xffu);
  s.jumps = (int)(p[1] & 0xffffu); s.last_jump_x = (int)(p[1] >> 16);
  s.max_gap = (int)(p[2] & 0xffffu); depth = (int)(p[2] >> 16);
}

// CPython heapq (Lib/heapq.py _siftdown / _siftup) on packed entries; Node.__lt__ compares h + balance * depth only
#define SMB_HP(e) ((e) >> 16)
SMB_HD void heap_siftdown(uint32_t* heap, int startpos, int pos) {
  const uint32_t item = heap[pos];
  while (pos > startpos) {
    const int parentpos = (pos - 1) >> 1;
    const uint32_t parent = heap[parentpos];
    if (SMB_HP(item) < SMB_HP(parent)) { heap[pos] = parent; pos = parentpos; continue; }
    break;
  }
  heap[pos] = item;
}
SMB_HD uint32_t heap_pop(uint32_t* heap, int& n) {
  const uint32_t last = heap[--n];
  if (n == 0) return last;
  const uint32_t ret = heap[0];
  int pos = 0, childpos = 1;
  while (childpos < n) {
    const int rightpos = childpos + 1;
    uint32_t child = heap[childpos];
    if (rightpos < n) {
      const uint32_t right = heap[rightpos];
      if (!(SMB_HP(child) < SMB_HP(right))) { childpos = rightpos; child = right; }
    }
    heap[pos] = child;
    pos = childpos;
    childpos = 2 * pos + 1;
  }
  heap[pos] = last;
  heap_siftdown(heap, 0, pos);
  return ret;
}

// AStarAgent.getSolution (engine.py:105-129): returns the node index of the winning node or of the best node
SMB_HD int astar(const Level& L, const State& s0, int balance, int max_iter, const Workspace& ws, bool& won) {
  const int dirs_x[4] = {0, 1, 0, 1}, dirs_y[4] = {0, 0, -1, -1};  // engine.py:3
  int nn = 0, nheap = 0, iterations = 0, best = -1, best_h = 0, best_depth = 0;
  for (int i = 0; i < VISITED_WORDS; i++) ws.visited[i] = 0u;
  node_store(ws.nodes, 0, s0, 0);
  ws.heap[0] = ((uint32_t)((L.exit_x - s0.x) + PRIO_BIAS) << 16) | 0u;
  nn = 1; nheap = 1;
  won = false;
  while ((iterations < max_iter || max_iter <= 0) && nheap > 0) {
    iterations++;
    const int cur = (int)(heap_pop(ws.heap, nheap) & 0xffffu);
    State cs;
    int depth;
    node_load(ws.nodes, cur, cs, depth);
    if (st_lose(L, cs)) continue;
    if (st_win(L, cs)) { won = true; return cur; }
    const int key = (((cs.y + 8) * 128 + cs.x) << 3) + cs.air;
    const uint32_t bit = 1u << (key & 31);
    if (!(ws.visited[key >> 5] & bit)) {
      const int h = L.exit_x - cs.x;
      if (best < 0 || h < best_h || (h == best_h && depth < best_depth)) { best = cur; best_h = h; best_depth = depth; }
      ws.visited[key >> 5] |= bit;
      for (int d = 0; d < 4; d++) {
        State c = cs;
        st_update(L, c, dirs_x[d], dirs_y[d]);
        node_store(ws.nodes, nn, c, depth + 1);
        ws.heap[nheap] = ((uint32_t)((L.exit_x - c.x) + balance * (depth + 1) + PRIO_BIAS) << 16) | (uint32_t)nn;
        nheap++;
        heap_siftdown(ws.heap, 0, nheap - 1);
        nn++;
      }
    }
  }
  return best;
}

// helper.py scans on the uint8 map
SMB_HD int floor_dist(const uint8_t* m, int w, int h, unsigned from_types, unsigned floor_types) {  // :37-62
  int result = 0;
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      if (!((from_types >> m[y * w + x]) & 1u)) continue;
      int dist = h - 1;
      for (int dy = 0; dy < h; dy++) {
        if (y + dy >= h) break;
        if ((floor_types >> m[(y + dy) * w + x]) & 1u) { dist = dy - 1; break; }
      }
      result += dist;
    }
  return result;
}

// SMBProblem.get_stats(map) -> st[0..7] = dist-floor, disjoint-tubes, enemies, empty, noise, jumps, jumps-dist, dist-win
SMB_HD void get_stats_one(const uint8_t* m, int w, int h, int power, uint32_t* solid_words, const Workspace& ws, int32_t* st) {
  int tubes = 0, enemies = 0, empty = 0, noise = 0;
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      const int t = m[y * w + x];
      enemies += (t == T_ENEMY);
      empty += (t == T_EMPTY);
      if (t == T_TUBE) {  // get_type_grouping(map, ["tube"], [(-1,0),(1,0)], 1, 1): helper.py:74-103
        const int nb = ((x >= 1 && m[y * w + x - 1] == T_TUBE) ? 1 : 0) + ((x + 1 < w && m[y * w + x + 1] == T_TUBE) ? 1 : 0);
        tubes += (nb == 1);
      }
      if (x >= 1 && m[y * w + x - 1] != t) noise++;   // get_changes(map, False): helper.py:115-133
      if (y >= 1 && m[(y - 1) * w + x] != t) noise++;  // get_changes(map, True)
    }
  st[0] = floor_dist(m, w, h, 1u << T_ENEMY, (1u << T_SOLID) | (1u << T_BRICK) | (1u << T_QUESTION));
  st[1] = tubes; st[2] = enemies; st[3] = empty; st[4] = noise;
  // _run_game (smb_prob.py:95-124): "   " / " @ " / "###" + row + " | " / " # " / "###"
  Level L;
  L.width = w + 6; L.height = h; L.exit_x = w + 4; L.solid = solid_words;
  for (int i = 0; i < h * ROW_WORDS; i++) solid_words[i] = 0u;
  for (int y = 0; y < h; y++) {
    uint32_t* row = solid_words + y * ROW_WORDS;
    const bool floor_rows = y > h - 3;
    for (int x = 0; x < L.width; x++) {
      bool s;
      if (x < 3 || x >= 3 + w) s = floor_rows || (y == h - 3 && x == 3 + w + 1);
      else { const int t = m[y * w + x - 3]; s = (t == T_SOLID || t == T_BRICK || t == T_QUESTION || t == T_TUBE); }
      if (s) row[x >> 5] |= 1u << (x & 31);
    }
  }
  State s0 = {1, h - 3, 0, 0, 0, 0};
  bool won;
  int sol = astar(L, s0, 1, power, ws, won);
  if (!won) sol = astar(L, s0, 0, power, ws, won);
  State ss;
  int depth;
  node_load(ws.nodes, sol, ss, depth);
  st[5] = ss.jumps;
  st[6] = (w - ss.last_jump_x > ss.max_gap) ? (w - ss.last_jump_x) : ss.max_gap;  // smb_prob.py:140-146
  st[7] = won ? 0 : (L.exit_x - ss.x);
}

}  // namespace pcgrl_smb

#ifdef __CUDACC__
namespace pcgrl_smb {

#define SMB_THREADS 64
#define SMB_MAX_CONCURRENCY 2048

__global__ void __launch_bounds__(SMB_THREADS) k_smb_get_stats(const uint8_t* __restrict__ maps, int32_t* stats_out, int n, int w,
                                                               int h, int power, uint32_t* scratch, int concurrency,
                                                               int out_stride) {
  __shared__ uint32_t solid_s[SMB_THREADS][MAX_H * ROW_WORDS];
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= concurrency) return;
  uint32_t* base = scratch + (size_t)tid * workspace_words(power);
  Workspace ws;
  ws.heap = base;
  ws.nodes = base + heap_words(power);
  ws.visited = ws.nodes + node_words(power);
  for (int i = tid; i < n; i += concurrency) {
    int32_t st[8];
    get_stats_one(maps + (size_t)i * w * h, w, h, power, solid_s[threadIdx.x], ws, st);
    for (int k = 0; k < out_stride; k++) stats_out[(size_t)i * out_stride + k] = (k < 8) ? st[k] : 0;
  }
}

}  // namespace pcgrl_smb
#endif

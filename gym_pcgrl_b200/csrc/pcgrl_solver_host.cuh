// pcgrl_solver_host.cuh -- host twins of the solver problems (sokoban, ddave, mdungeon): Problem.get_stats on the CPU.
//
// The game models (State.update / getHeuristic / checkWin, Node.getChildren pruning, the level framing) are the
// `__host__ __device__` functions of pcgrl_solver.cuh, i.e. the very code the warp searches run; only the search loop is
// written again here, as the plain scalar loop of the reference (sokoban/engine.py:56-119, ddave/engine.py:61-129,
// mdungeon/engine.py:61-129): FIFO queue or CPython heapq order, visited set on State.getKey, best node = lowest
// heuristic then lowest depth, at most `solver_power` popped nodes per pass.  Passes run one after the other in the
// reference's order and stop at the first win (sokoban_prob.py:110-122, ddave_prob.py:122-135, mdungeon_prob.py:125-138);
// the GPU kernels' speculative / exhaustion shortcuts are not used here, so this file doubles as an independent check
// of them (tests/test_host_twins.py replays the reference's golden trajectories through it).
#pragma once
#include <stdint.h>

#include <unordered_set>
#include <vector>

#include "pcgrl_host_twin.cuh"
#include "pcgrl_solver.cuh"

namespace pcgrl_host {

struct StateKey {
  uint32_t w[5];
  bool operator==(const StateKey& o) const { return w[0] == o.w[0] && w[1] == o.w[1] && w[2] == o.w[2] && w[3] == o.w[3] && w[4] == o.w[4]; }
};
struct StateKeyHash {
  size_t operator()(const StateKey& k) const {
    uint64_t h = 1469598103934665603ull;
    for (int i = 0; i < 5; i++) h = (h ^ k.w[i]) * 1099511628211ull;
    return (size_t)(h ^ (h >> 29));
  }
};
struct SearchWork {
  std::vector<pcgrl::SState> nodes;
  std::vector<uint64_t> heap;  // 1-based; entry = priority << 32 | node index, ordered by the priority alone (strict <)
  std::unordered_set<StateKey, StateKeyHash> visited;
};

// Lib/heapq.py heappush / heappop on `priority << 32 | node`
static inline uint64_t hprio(uint64_t e) { return e >> 32; }
static inline void heap_siftdown(std::vector<uint64_t>& h, size_t pos, uint64_t item) {
  while (pos > 1) {
    const size_t parent = pos >> 1;
    if (hprio(item) < hprio(h[parent])) { h[pos] = h[parent]; pos = parent; continue; }
    break;
  }
  h[pos] = item;
}
static inline void heap_push(std::vector<uint64_t>& h, uint64_t item) {
  h.push_back(item);
  heap_siftdown(h, h.size() - 1, item);
}
static inline uint64_t heap_pop(std::vector<uint64_t>& h) {  // h.size() - 1 entries, >= 1
  const uint64_t last = h.back();
  h.pop_back();
  const size_t n = h.size() - 1;
  if (n == 0) return last;
  const uint64_t ret = h[1];
  size_t pos = 1, child = 2;
  while (child <= n) {
    if (child < n && !(hprio(h[child]) < hprio(h[child + 1]))) child++;
    h[pos] = h[child];
    pos = child;
    child = 2 * pos;
  }
  heap_siftdown(h, pos, last);
  return ret;
}

// One pass (b < 0: BFSAgent, else AStarAgent with priority 2*h + b*depth, b = 2*balance); res = {won, depth, h, misc}
// of the winning node, else of the best node.
template <int GAME>
static inline void search_pass(const pcgrl::Level& L, const pcgrl::SState& root0, int b, int power, SearchWork& w, int* res) {
  using namespace pcgrl;
  const bool check_lose = (GAME != GAME_SOKOBAN);
  const bool sk_small = (GAME == GAME_SOKOBAN) && L.small;
  w.nodes.clear();
  w.heap.assign(1, 0ull);
  w.visited.clear();
  SState root = root0;
  if (sk_small) {
    const unsigned long long occ = sk_occupancy(L, root);
    root.misc = (uint32_t)occ;
    root.pad = (uint32_t)(occ >> 32);
  }
  root.dh = 0u | ((uint32_t)(g_heuristic<GAME>(L, root) + SOLVER_PRIO_BIAS) << 16);
  w.nodes.push_back(root);
  if (b >= 0) heap_push(w.heap, ((uint64_t)(2 * st_h(root) + 2 * SOLVER_PRIO_BIAS) << 32) | 0ull);
  size_t head = 0;
  int iterations = 0, best = -1, best_h = 0, best_depth = 0;
  res[0] = 0;
  while (iterations < power && (b >= 0 ? w.heap.size() > 1 : head < w.nodes.size())) {
    iterations++;
    const int cur = (b >= 0) ? (int)(heap_pop(w.heap) & 0xffffffffull) : (int)head++;
    const SState cs = w.nodes[cur];
    if (check_lose && st_health(cs) <= 0) continue;
    bool win;
    if (sk_small) {
      const unsigned long long occ = (unsigned long long)cs.misc | ((unsigned long long)cs.pad << 32);
      win = (occ & L.target64) == L.target64 && L.ntargets == L.ncrates && L.ntargets > 0;
    } else {
      win = g_win<GAME>(L, cs);
    }
    if (win) {
      res[0] = 1; res[1] = st_depth(cs); res[2] = st_h(cs); res[3] = (int)cs.misc;
      return;
    }
    StateKey k;
    k.w[0] = cs.m[0]; k.w[1] = cs.m[1]; k.w[2] = cs.m[2]; k.w[3] = cs.m[3]; k.w[4] = cs.ks;
    if (!w.visited.insert(k).second) continue;
    const int ch = st_h(cs), cd = st_depth(cs);
    if (best < 0 || ch < best_h || (ch == best_h && cd < best_depth)) { best = cur; best_h = ch; best_depth = cd; }
    for (int d = 0; d < 4; d++) {
      SState c = cs;
      int h = 0;
      if (!make_child<GAME>(L, cs, d, sk_small, c, h)) continue;
      c.dh = (uint32_t)(cd + 1) | ((uint32_t)(h + SOLVER_PRIO_BIAS) << 16);
      const int idx = (int)w.nodes.size();
      w.nodes.push_back(c);
      if (b >= 0) heap_push(w.heap, ((uint64_t)(2 * h + b * (cd + 1) + 2 * SOLVER_PRIO_BIAS) << 32) | (uint64_t)idx);
    }
  }
  const SState& bs = w.nodes[best < 0 ? 0 : best];
  res[1] = st_depth(bs); res[2] = st_h(bs); res[3] = (int)bs.misc;
}

// _run_game on a map whose preconditions hold: patches the play-through statistics in st
template <int GAME>
static inline void run_game(const pcgrl_config* cfg, const uint8_t* map, SearchWork& w, int32_t* st) {
  using namespace pcgrl;
  const int W = cfg->width, H = cfg->height;
  Level L;
  SState root;
  for (int i = 0; i < W * H; i++) L.tiles[i] = map[i];
  level_init<GAME>(L, root, W, H);
  int res[4] = {0, 0, 0, 0};
  if (!L.overflow) {  // > 16 crates or targets: the kernels report status[0] and leave the defaults, so does the twin
    for (int pass = 0; pass < 4; pass++) {
      const int b = (GAME == GAME_SOKOBAN) ? ((pass == 0) ? -1 : (pass == 1) ? 2 : (pass == 2) ? 1 : 0)
                                           : ((pass == 0) ? 2 : (pass == 1) ? 1 : (pass == 2) ? 0 : -1);
      search_pass<GAME>(L, root, b, cfg->solver_power, w, res);
      if (res[0] == 1) break;
    }
  }
  const int won = res[0], depth = res[1], h = res[2];
  const uint32_t misc = (uint32_t)res[3];
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

// helper.py:37-62 get_floor_dist: per `from` cell, the cells strictly between it and the first floor cell below (H-1 if none)
static inline int floor_dist(const uint8_t* map, int W, int H, int from_tile, int floor_tile) {
  int total = 0;
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++) {
      if (map[y * W + x] != from_tile) continue;
      int d = H - 1;
      for (int yy = y + 1; yy < H; yy++)
        if (map[yy * W + x] == floor_tile) { d = yy - y - 1; break; }
      total += d;
    }
  return total;
}

// Problem.get_stats of sokoban (sokoban_prob.py:133-145), ddave (ddave_prob.py:149-169), mdungeon (mdungeon_prob.py:151-171)
static inline bool solver_get_stats(const pcgrl_config* cfg, const uint8_t* map, SearchWork& w, int32_t* st) {
  const int W = cfg->width, H = cfg->height;
  for (int i = 0; i < PCGRL_MAX_STATS; i++) st[i] = 0;
  auto count = [&](unsigned types) { return popc(type_rows(map, W, H, types)); };
  if (cfg->problem == PCGRL_PROB_SOKOBAN) {
    st[0] = count(0x04u); st[1] = count(0x08u); st[2] = count(0x10u);
    st[3] = count_regions(type_rows(map, W, H, 0x1Du));
    st[4] = W * H * (W + H);
    if (st[0] == 1 && st[1] == st[2] && st[1] > 0 && st[3] == 1) run_game<pcgrl::GAME_SOKOBAN>(cfg, map, w, st);
    return true;
  }
  if (cfg->problem == PCGRL_PROB_DDAVE) {
    st[0] = count(0x04u);
    st[1] = floor_dist(map, W, H, 2, 1);
    st[2] = count(0x08u); st[3] = count(0x10u); st[4] = count(0x20u); st[5] = count(0x40u);
    st[6] = count_regions(type_rows(map, W, H, 0x3Du));
    st[9] = W * H;
    if (st[0] == 1 && st[2] == 1 && st[4] == 1 && st[6] == 1) run_game<pcgrl::GAME_DDAVE>(cfg, map, w, st);
    return true;
  }
  if (cfg->problem == PCGRL_PROB_MDUNGEON) {
    st[0] = count(0x04u); st[1] = count(0x08u); st[2] = count(0x10u); st[3] = count(0x20u); st[4] = count(0xC0u);
    st[5] = count_regions(type_rows(map, W, H, 0xFDu));
    st[9] = W * H;
    if (st[0] == 1 && st[1] == 1 && st[5] == 1) run_game<pcgrl::GAME_MDUNGEON>(cfg, map, w, st);
    return true;
  }
  return false;
}

}  // namespace pcgrl_host

/* smb_oracle.c -- CPU restatement of the reference's `smb` problem (groundwork).
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under gym_pcgrl_b200/ may link or call this file; the product does not
 * implement smb yet (the 114 x 14 map does not fit the one-row-per-lane bitboards).  The restatement is PINNED:
 * tests/test_oracle_smb.py checks it against golden vectors produced by executing the unmodified reference
 * (tests/golden/make_golden_smb.py).
 *
 * Follows, function by function (R = /root/reference/gym_pcgrl/envs):
 *   R/probs/smb_prob.py:95-124   _run_game        (level framing, A*(balance 1) then A*(balance 0), power 10000)
 *   R/probs/smb_prob.py:126-148  get_stats
 *   R/probs/smb_prob.py:150-172  get_reward
 *   R/probs/smb_prob.py:174-175  get_episode_over
 *   R/probs/smb/engine.py:105-129 AStarAgent.getSolution (queue.PriorityQueue == CPython heapq, Node.__lt__ :52-53)
 *   R/probs/smb/engine.py:131-286 State (stringInitialize, checkMovableLocation, update, getKey, getHeuristic)
 *   R/helper.py:37-62 get_floor_dist, :74-103 get_type_grouping, :115-133 get_changes, :366-376 get_range_reward
 *
 * Tiles: 0 empty, 1 solid, 2 enemy, 3 brick, 4 question, 5 coin, 6 tube (smb_prob.py:36-37).
 * Stats row: dist-floor, disjoint-tubes, enemies, empty, noise, jumps, jumps-dist, dist-win.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define SMB_NSTATS 8
enum { T_EMPTY = 0, T_SOLID, T_ENEMY, T_BRICK, T_QUESTION, T_COIN, T_TUBE };

/* ---- helper.py ------------------------------------------------------------------------------- */
static int calc_dist_floor(const uint8_t* m, int w, int h, int x, int y, unsigned floor_types) { /* :37-43 */
  for (int dy = 0; dy < h; dy++) {
    if (y + dy >= h) break;
    if ((floor_types >> m[(y + dy) * w + x]) & 1u) return dy - 1;
  }
  return h - 1;
}
static int get_floor_dist(const uint8_t* m, int w, int h, unsigned from_types, unsigned floor_types) { /* :56-62 */
  int result = 0;
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++)
      if ((from_types >> m[y * w + x]) & 1u) result += calc_dist_floor(m, w, h, x, y, floor_types);
  return result;
}
/* get_type_grouping(map, types, relLocs = [(-1,0),(1,0)], min, max): :74-103 */
static int get_type_grouping_h(const uint8_t* m, int w, int h, unsigned types, int lo, int hi) {
  int result = 0;
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      if (!((types >> m[y * w + x]) & 1u)) continue;
      int value = 0;
      if (x - 1 >= 0 && ((types >> m[y * w + x - 1]) & 1u)) value++;
      if (x + 1 < w && ((types >> m[y * w + x + 1]) & 1u)) value++;
      if (value >= lo && value <= hi) result++;
    }
  return result;
}
static int get_changes(const uint8_t* m, int w, int h, int vertical) { /* :115-133 */
  const int sy = vertical ? 1 : 0, sx = vertical ? 0 : 1;
  int value = 0;
  for (int y = sy; y < h; y++)
    for (int x = sx; x < w; x++) {
      const int same = vertical ? (m[y * w + x] == m[(y - 1) * w + x]) : (m[y * w + x] == m[y * w + x - 1]);
      if (!same) value++;
    }
  return value;
}
static int count_tile(const uint8_t* m, int n, int t) {
  int c = 0;
  for (int i = 0; i < n; i++) c += (m[i] == t);
  return c;
}
static double range_reward(double nv, double ov, double low, double high) { /* :366-376 */
  if (nv >= low && nv <= high && ov >= low && ov <= high) return 0;
  if (ov <= high && nv <= high) return fmin(nv, low) - fmin(ov, low);
  if (ov >= low && nv >= low) return fmax(ov, high) - fmax(nv, high);
  if (nv > high && ov < low) return high - nv + ov - low;
  if (nv < low && ov > high) return high - ov + nv - low;
  return 0;
}

/* ---- engine.py State ------------------------------------------------------------------------- */
typedef struct {
  int width, height, exit_x;
  uint8_t* solid; /* [height][width] */
} level_t;

/* jump_locs is only consumed through len() and the largest gap between consecutive jump x's (smb_prob.py:140-146),
 * so a state carries (jumps, x of the last jump, largest gap so far) instead of the list. */
typedef struct {
  int x, y, air, jumps, last_jump_x, max_gap;
} state_t;

typedef struct {
  state_t st;
  int depth, h;
} node_t;

static int movable(const level_t* L, int x, int y) { /* :203-206 checkMovableLocation */
  if (y < 0) return 1;
  return !(x < 0 || x >= L->width || y >= L->height || L->solid[y * L->width + x]);
}
static int st_win(const level_t* L, const state_t* s) { return s->x >= L->exit_x; }  /* :258-259 */
static int st_lose(const level_t* L, const state_t* s) { return s->y >= L->height; } /* :261-262 */

static void st_update(const level_t* L, state_t* s, int dir_x, int dir_y) { /* :208-246 */
  if (st_win(L, s) || st_lose(L, s)) return;
  if (dir_x > 0) dir_x = 1;
  if (dir_x < 0) dir_x = -1;
  dir_y = (dir_y < 0) ? -1 : 0;
  int ground = 0;
  if (s->y < L->height - 1 && s->y >= -1) ground = L->solid[(s->y + 1) * L->width + s->x];
  int nx = s->x, ny = s->y;
  if (dir_x != 0 && movable(L, nx + dir_x, ny)) nx += dir_x;
  if (dir_y == -1) {
    if (ground && movable(L, nx, ny - 1)) {
      s->air = 5;
      s->jumps += 1;
      /* jump_locs.append((player x BEFORE the horizontal move, y)): gap bookkeeping of smb_prob.py:141-145 */
      if (s->x - s->last_jump_x > s->max_gap) s->max_gap = s->x - s->last_jump_x;
      s->last_jump_x = s->x;
    }
  } else if (s->air > 0) {
    s->air = 1;
  }
  if (s->air > 1) {
    s->air -= 1;
    if (movable(L, nx, ny - 1)) ny -= 1;
    else s->air = 1;
  } else if (s->air == 1) {
    s->air = 0;
  } else if (movable(L, nx, ny + 1)) {
    ny += 1;
  }
  s->x = nx;
  s->y = ny;
}

/* ---- CPython heapq on node indices; Node.__lt__: h + balance * depth, strict <  (engine.py:52-53).
 * balance is 1 or 0, so the key is an int. */
typedef struct {
  node_t* nodes;
  int* heap;
  uint8_t* visited; /* key (x, y, airTime): x in [0, width), y in [-8, height], air in [0, 5] */
  int cap;
} ws_t;
static __thread ws_t g_ws;

static void ws_reserve(int power, int width, int height) {
  const int cap = 4 * power + 8;
  if (g_ws.cap < cap) {
    free(g_ws.nodes); free(g_ws.heap);
    g_ws.nodes = (node_t*)malloc(sizeof(node_t) * (size_t)cap);
    g_ws.heap = (int*)malloc(sizeof(int) * (size_t)cap);
    g_ws.cap = cap;
  }
  free(g_ws.visited);
  g_ws.visited = (uint8_t*)calloc((size_t)width * (size_t)(height + 16) * 8, 1);
}
static int prio(const node_t* n, int balance) { return n->h + balance * n->depth; }
static void heap_siftdown(int* heap, const node_t* nodes, int b, int startpos, int pos) { /* heapq._siftdown */
  const int item = heap[pos];
  while (pos > startpos) {
    const int parentpos = (pos - 1) >> 1, parent = heap[parentpos];
    if (prio(&nodes[item], b) < prio(&nodes[parent], b)) { heap[pos] = parent; pos = parentpos; continue; }
    break;
  }
  heap[pos] = item;
}
static void heap_push(int* heap, int* n, const node_t* nodes, int b, int item) {
  heap[(*n)++] = item;
  heap_siftdown(heap, nodes, b, 0, *n - 1);
}
static int heap_pop(int* heap, int* n, const node_t* nodes, int b) { /* heapq.heappop -> _siftup */
  const int last = heap[--(*n)];
  if (*n == 0) return last;
  const int ret = heap[0];
  int pos = 0, childpos = 1;
  const int endpos = *n;
  while (childpos < endpos) {
    const int rightpos = childpos + 1;
    if (rightpos < endpos && !(prio(&nodes[heap[childpos]], b) < prio(&nodes[heap[rightpos]], b))) childpos = rightpos;
    heap[pos] = heap[childpos];
    pos = childpos;
    childpos = 2 * pos + 1;
  }
  heap[pos] = last;
  heap_siftdown(heap, nodes, b, 0, pos);
  return ret;
}

static const int DIRS[4][2] = {{0, 0}, {1, 0}, {0, -1}, {1, -1}}; /* engine.py:3 */

/* AStarAgent.getSolution(state, balance, maxIterations) -> index of the returned node (win node or best node) */
static int astar(const level_t* L, const state_t* s0, int balance, int max_iter, int* won, long* iters) {
  node_t* nodes = g_ws.nodes;
  int* heap = g_ws.heap;
  int nn = 0, nheap = 0, iterations = 0, best = -1;
  memset(g_ws.visited, 0, (size_t)L->width * (size_t)(L->height + 16) * 8);
  nodes[nn].st = *s0; nodes[nn].depth = 0; nodes[nn].h = L->exit_x - s0->x;
  heap_push(heap, &nheap, nodes, balance, nn++);
  *won = 0;
  while ((iterations < max_iter || max_iter <= 0) && nheap > 0) {
    iterations++;
    const int cur = heap_pop(heap, &nheap, nodes, balance);
    const state_t cs = nodes[cur].st;
    if (st_lose(L, &cs)) continue;
    if (st_win(L, &cs)) { *won = 1; *iters += iterations; return cur; }
    uint8_t* v = &g_ws.visited[(((size_t)(cs.y + 8) * L->width + cs.x) << 3) + cs.air];
    if (!*v) {
      if (best < 0 || nodes[cur].h < nodes[best].h || (nodes[cur].h == nodes[best].h && nodes[cur].depth < nodes[best].depth))
        best = cur;
      *v = 1;
      for (int d = 0; d < 4; d++) {
        node_t* c = &nodes[nn];
        c->st = cs;
        st_update(L, &c->st, DIRS[d][0], DIRS[d][1]);
        c->depth = nodes[cur].depth + 1;
        c->h = L->exit_x - c->st.x;
        heap_push(heap, &nheap, nodes, balance, nn++);
      }
    }
  }
  *iters += iterations;
  return best;
}

/* _run_game (smb_prob.py:95-124): level string rows are
 *   "   " / " @ " / "###"  +  map row with solid, brick, question, tube -> '#'  +  " | " / " # " / "###"
 * so the level is (w + 6) wide, the player starts at (1, h - 3) and the exit column is w + 4. */
static int run_game(const uint8_t* m, int w, int h, int power, int* jumps, int* jumps_dist, long* iters) {
  level_t L;
  L.width = w + 6; L.height = h; L.exit_x = w + 4;
  L.solid = (uint8_t*)calloc((size_t)L.width * h, 1);
  for (int y = 0; y < h; y++) {
    uint8_t* row = L.solid + (size_t)y * L.width;
    const int floor_rows = (y > h - 3);
    for (int x = 0; x < 3; x++) row[x] = floor_rows;
    for (int x = 0; x < w; x++) {
      const int t = m[y * w + x];
      row[3 + x] = (t == T_SOLID || t == T_BRICK || t == T_QUESTION || t == T_TUBE);
    }
    for (int x = 0; x < 3; x++) row[3 + w + x] = floor_rows;
    if (y == h - 3) row[3 + w + 1] = 1; /* " # " */
  }
  state_t s0 = {1, h - 3, 0, 0, 0, 0};
  ws_reserve(power, L.width, h);
  int won = 0;
  int sol = astar(&L, &s0, 1, power, &won, iters);
  if (!won) sol = astar(&L, &s0, 0, power, &won, iters);
  const state_t* ss = &g_ws.nodes[sol].st;
  *jumps = ss->jumps;
  int value = ss->max_gap;                                   /* smb_prob.py:140-146 */
  if (w - ss->last_jump_x > value) value = w - ss->last_jump_x;
  *jumps_dist = value;
  const int dist_win = won ? 0 : (L.exit_x - ss->x);
  free(L.solid);
  return dist_win;
}

/* ---- smb_prob.py ----------------------------------------------------------------------------- */
static long g_iterations = 0;

void smb_get_stats(const uint8_t* map, int w, int h, int power, int32_t* st) {
  const unsigned floor_types = (1u << T_SOLID) | (1u << T_BRICK) | (1u << T_QUESTION); /* "tube_left/right" never occur */
  st[0] = get_floor_dist(map, w, h, 1u << T_ENEMY, floor_types);
  st[1] = get_type_grouping_h(map, w, h, 1u << T_TUBE, 1, 1);
  st[2] = count_tile(map, w * h, T_ENEMY);
  st[3] = count_tile(map, w * h, T_EMPTY);
  st[4] = get_changes(map, w, h, 0) + get_changes(map, w, h, 1);
  int jumps = 0, jumps_dist = 0;
  long iters = 0;
  st[7] = run_game(map, w, h, power, &jumps, &jumps_dist, &iters);
  st[5] = jumps;
  st[6] = jumps_dist;
  __atomic_fetch_add(&g_iterations, iters, __ATOMIC_RELAXED);
}

/* weights: dist-floor, disjoint-tubes, enemies, empty, noise, jumps, jumps-dist, dist-win (smb_prob.py:24-33);
 * iparam: min_empty, min_enemies, max_enemies, min_jumps (:19-22) */
double smb_get_reward(const int32_t* n, const int32_t* o, const double* w, const int32_t* ip) {
  return range_reward(n[0], o[0], 0, 0) * w[0] + range_reward(n[1], o[1], 0, 0) * w[1] +
         range_reward(n[2], o[2], ip[1], ip[2]) * w[2] + range_reward(n[3], o[3], ip[0], INFINITY) * w[3] +
         range_reward(n[4], o[4], 0, 0) * w[4] + range_reward(n[5], o[5], ip[3], INFINITY) * w[5] +
         range_reward(n[6], o[6], 0, 0) * w[6] + range_reward(n[7], o[7], 0, 0) * w[7];
}
int smb_episode_over(const int32_t* n) { return n[7] <= 0; }

void smb_get_stats_batch(const uint8_t* maps, int n, int w, int h, int power, int32_t* out) {
  for (int i = 0; i < n; i++) smb_get_stats(maps + (size_t)i * w * h, w, h, power, out + (size_t)i * SMB_NSTATS);
}
long smb_solver_iterations(void) { return g_iterations; }

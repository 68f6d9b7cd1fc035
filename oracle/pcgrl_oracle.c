/*
 * pcgrl_oracle.c -- CPU ORACLE for the batched PcgrlEnv hot path.  TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * A scalar C restatement of the reference's (amidos2006/gym-pcgrl @ 0385b2e) Python algorithms for
 * PcgrlEnv.reset/step -> Representation.update -> Problem.get_stats -> get_reward/get_episode_over.
 * It deliberately keeps the reference's data structures (FIFO-queue flood fill and BFS, node lists,
 * CPython binary heap) so that it is an independent check of the bitboard CUDA kernels.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may call it.
 *
 * Parity status: PINNED.  The reference has no tests or golden vectors of its own, so the
 * oracle is pinned against outputs of the unmodified reference executed in the build container:
 * tests/golden/traj_*.npz (26 trajectories x 1000 steps, incl. SURVEY App. B.3 digests),
 * tests/golden/stats_*.npz (Problem.get_stats on ~1900 maps) and tests/golden/rng.npz (numpy legacy
 * RandomState streams); see tests/test_oracle_golden.py.  Unpinned piece: gym.utils.seeding's seed
 * hashing (gym is absent from the container) -- it lives in host Python, not here.
 *
 * Every function cites the reference file:line it follows (paths relative to the reference root,
 * G = gym_pcgrl/envs).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/pcgrl_b200.h"

#define MAXD PCGRL_MAX_DIM
#define MAXCELLS (MAXD * MAXD)
#define S PCGRL_MAX_STATS

/* ------------------------------------------------------------------------------------------------
 * MT19937 + numpy legacy RandomState algorithms (third-party: numpy >= 1.17, legacy stream frozen;
 * numpy/random/src/mt19937/mt19937.c, _legacy/legacy-distributions.c, mtrand.pyx choice/randint).
 * State layout: 624 key words + word 624 = pos (RandomState.get_state()[1], [2]).
 * ---------------------------------------------------------------------------------------------- */
static void mt_seed(uint32_t* st, uint32_t seed) { /* init_genrand; RandomState(seed) leaves pos = 624 */
  st[0] = seed;
  for (int i = 1; i < 624; i++) st[i] = 1812433253u * (st[i - 1] ^ (st[i - 1] >> 30)) + (uint32_t)i;
  st[624] = 624;
}

static void mt_twist(uint32_t* k) {
  const uint32_t UP = 0x80000000u, LO = 0x7fffffffu, A = 0x9908b0dfu;
  int i;
  uint32_t y;
  for (i = 0; i < 624 - 397; i++) {
    y = (k[i] & UP) | (k[i + 1] & LO);
    k[i] = k[i + 397] ^ (y >> 1) ^ ((y & 1u) ? A : 0u);
  }
  for (; i < 623; i++) {
    y = (k[i] & UP) | (k[i + 1] & LO);
    k[i] = k[i + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? A : 0u);
  }
  y = (k[623] & UP) | (k[0] & LO);
  k[623] = k[396] ^ (y >> 1) ^ ((y & 1u) ? A : 0u);
  k[624] = 0;
}

static uint32_t mt_u32(uint32_t* st) {
  if (st[624] >= 624) mt_twist(st);
  uint32_t y = st[st[624]++];
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  return y;
}

static double mt_double(uint32_t* st) { /* random_sample(): 53-bit double from two draws */
  uint32_t a = mt_u32(st) >> 5, b = mt_u32(st) >> 6;
  return (a * 67108864.0 + b) / 9007199254740992.0;
}

static int mt_randint(uint32_t* st, int n) { /* RandomState.randint(n): masked rejection, no draw if n == 1 */
  uint32_t rng = (uint32_t)(n - 1), mask, v;
  if (rng == 0) return 0;
  mask = rng;
  mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
  do { v = mt_u32(st) & mask; } while (v > rng);
  return (int)v;
}

/* ------------------------------------------------------------------------------------------------
 * helper.py restatements
 * ---------------------------------------------------------------------------------------------- */
typedef struct { int w, h; const uint8_t* m; } grid_t;
#define AT(g, x, y) ((g)->m[(y) * (g)->w + (x)])

static const int DX[4] = {-1, 1, 0, 0}, DY[4] = {0, 0, -1, 1}; /* G/helper.py:179,232 neighbour order */

/* G/helper.py:16-23 get_tile_locations + :150-154 _get_certain_tiles: (x,y) of every tile whose type is in
 * `types` (bitmask), concatenated per type in the order given by `order` (row-major inside one type). */
static int certain_tiles(const grid_t* g, const int* order, int norder, int* xs, int* ys) {
  int n = 0;
  for (int k = 0; k < norder; k++)
    for (int y = 0; y < g->h; y++)
      for (int x = 0; x < g->w; x++)
        if (AT(g, x, y) == order[k]) { xs[n] = x; ys[n] = y; n++; }
  return n;
}

static int count_tiles(const grid_t* g, unsigned types) { /* G/helper.py:272-273 calc_certain_tile */
  int n = 0;
  for (int i = 0; i < g->w * g->h; i++) n += (types >> g->m[i]) & 1u;
  return n;
}

/* G/helper.py:170-184 _flood_fill (FIFO queue, colours on first pop) */
static int flood_fill(const grid_t* g, int x, int y, int* color, int index, unsigned passable, int* queue) {
  int head = 0, tail = 0, num = 0;
  queue[tail++] = y * g->w + x;
  while (head < tail) {
    int c = queue[head++], cx = c % g->w, cy = c / g->w;
    if (color[c] != -1 || !((passable >> AT(g, cx, cy)) & 1u)) continue;
    num++;
    color[c] = index;
    for (int d = 0; d < 4; d++) {
      int nx = cx + DX[d], ny = cy + DY[d];
      if (nx < 0 || ny < 0 || nx >= g->w || ny >= g->h) continue;
      queue[tail++] = ny * g->w + nx;
    }
  }
  return num;
}

/* G/helper.py:197-207 calc_num_regions */
static int calc_num_regions(const grid_t* g, const int* order, int norder, unsigned passable) {
  static __thread int xs[MAXCELLS], ys[MAXCELLS], color[MAXCELLS], queue[4 * MAXCELLS + 8];
  int n = certain_tiles(g, order, norder, xs, ys), region = 0;
  for (int i = 0; i < g->w * g->h; i++) color[i] = -1;
  for (int i = 0; i < n; i++)
    if (flood_fill(g, xs[i], ys[i], color, region + 1, passable, queue) > 0) region++;
  return region;
}

/* G/helper.py:222-237 run_dikjstra: unit-cost BFS; -1 = unreachable (or source not passable) */
static void run_dikjstra(const grid_t* g, int x, int y, unsigned passable, int* dmap, uint8_t* visited) {
  static __thread int qc[4 * MAXCELLS + 8], qd[4 * MAXCELLS + 8];
  int head = 0, tail = 0;
  for (int i = 0; i < g->w * g->h; i++) { dmap[i] = -1; if (visited) visited[i] = 0; }
  qc[tail] = y * g->w + x; qd[tail++] = 0;
  while (head < tail) {
    int c = qc[head], cd = qd[head++], cx = c % g->w, cy = c / g->w;
    if (!((passable >> AT(g, cx, cy)) & 1u) || (dmap[c] >= 0 && dmap[c] <= cd)) continue;
    if (visited) visited[c] = 1;
    dmap[c] = cd;
    for (int d = 0; d < 4; d++) {
      int nx = cx + DX[d], ny = cy + DY[d];
      if (nx < 0 || ny < 0 || nx >= g->w || ny >= g->h) continue;
      qc[tail] = ny * g->w + nx; qd[tail++] = cd + 1;
    }
  }
}

/* G/helper.py:250-264 calc_longest_path: per component double sweep with np.argmax tie-break */
static int calc_longest_path(const grid_t* g, const int* order, int norder, unsigned passable) {
  static __thread int xs[MAXCELLS], ys[MAXCELLS], dmap[MAXCELLS];
  static __thread uint8_t visited[MAXCELLS], final_visited[MAXCELLS];
  int n = certain_tiles(g, order, norder, xs, ys), cells = g->w * g->h, final_value = 0;
  memset(final_visited, 0, (size_t)cells);
  for (int i = 0; i < n; i++) {
    if (final_visited[ys[i] * g->w + xs[i]]) continue;
    run_dikjstra(g, xs[i], ys[i], passable, dmap, visited);
    for (int c = 0; c < cells; c++) final_visited[c] |= visited[c];
    int arg = 0; /* np.argmax: first row-major index holding the maximum */
    for (int c = 1; c < cells; c++) if (dmap[c] > dmap[arg]) arg = c;
    run_dikjstra(g, arg % g->w, arg / g->w, passable, dmap, NULL);
    int mx = dmap[0];
    for (int c = 1; c < cells; c++) if (dmap[c] > mx) mx = dmap[c];
    if (mx > final_value) final_value = mx;
  }
  return final_value;
}

/* G/helper.py:37-62 get_floor_dist / _calc_dist_floor */
static int get_floor_dist(const grid_t* g, unsigned from_types, unsigned floor_types) {
  int result = 0;
  for (int y = 0; y < g->h; y++)
    for (int x = 0; x < g->w; x++) {
      if (!((from_types >> AT(g, x, y)) & 1u)) continue;
      int r = g->h - 1;
      for (int dy = 0; dy < g->h; dy++) {
        if (y + dy >= g->h) break;
        if ((floor_types >> AT(g, x, y + dy)) & 1u) { r = dy - 1; break; }
      }
      result += r;
    }
  return result;
}

/* G/helper.py:366-376 get_range_reward (bounds may be +-inf) */
static double range_reward(double nv, double ov, double low, double high) {
  if (nv >= low && nv <= high && ov >= low && ov <= high) return 0;
  if (ov <= high && nv <= high) return fmin(nv, low) - fmin(ov, low);
  if (ov >= low && nv >= low) return fmax(ov, high) - fmax(nv, high);
  if (nv > high && ov < low) return high - nv + ov - low;
  if (nv < low && ov > high) return high - ov + nv - low;
  return 0; /* unreachable for ordered finite inputs (the reference would return None) */
}

/* ------------------------------------------------------------------------------------------------
 * Game engines + search agents  (G/probs/{sokoban,ddave,mdungeon}/engine.py)
 * Level = map wrapped in a 1-tile solid border (G/probs/*_prob.py _run_game), engine coords = map + 1.
 * ---------------------------------------------------------------------------------------------- */
#define MAXOBJ 256 /* collectibles / crates per level handled by the oracle */
enum { GAME_SOKOBAN = 0, GAME_DDAVE = 1, GAME_MDUNGEON = 2 };

typedef struct {
  int game, w, h; /* bordered size */
  uint8_t solid[(MAXD + 2) * (MAXD + 2)];
  uint8_t deadlock[(MAXD + 2) * (MAXD + 2)]; /* sokoban */
  int nobj;                                  /* ddave: diamonds; mdungeon: potions+treasures+enemies; sokoban: crates */
  int ox[MAXOBJ], oy[MAXOBJ], otype[MAXOBJ]; /* mdungeon otype: 0 potion, 1 treasure, 2 goblin, 3 ogre */
  int ntargets, tx[MAXOBJ], ty[MAXOBJ];      /* sokoban targets */
  int nspikes, sx[MAXCELLS], sy[MAXCELLS];   /* ddave */
  int doorx, doory, keyx, keyy;
} level_t;

typedef struct {
  int16_t px, py, health, air, jumps, c0, c1, c2; /* c0..c2: ddave diamonds,key,-; mdungeon potions,treasures,enemies */
  uint8_t key_present;                            /* ddave */
  uint8_t remain[MAXOBJ / 8];                     /* ddave / mdungeon: object i still on the map */
  uint8_t cx[64], cy[64];                         /* sokoban crates, index order preserved (engine.py:329-335 key) */
} gstate_t;

typedef struct { gstate_t st; int parent, depth, h; } node_t;

static int lv_solid(const level_t* L, int x, int y) { return L->solid[y * L->w + x]; }

/* --- sokoban ------------------------------------------------------------------------------------ */
static int sk_target_at(const level_t* L, int x, int y) { /* sokoban/engine.py:256-260 */
  for (int i = 0; i < L->ntargets; i++) if (L->tx[i] == x && L->ty[i] == y) return 1;
  return 0;
}
static int sk_crate_at(const level_t* L, const gstate_t* s, int x, int y) { /* :262-266, first match in list order */
  for (int i = 0; i < L->nobj; i++) if (s->cx[i] == x && s->cy[i] == y) return i;
  return -1;
}
static int sk_movable(const level_t* L, const gstate_t* s, int x, int y) { /* :268-269 */
  if (x < 0 || y < 0 || x > L->w - 1 || y > L->h - 1) return 0;
  return !lv_solid(L, x, y) && sk_crate_at(L, s, x, y) < 0;
}
static int sk_win(const level_t* L, const gstate_t* s) { /* :271-280 */
  if (L->ntargets != L->nobj || L->ntargets == 0 || L->nobj == 0) return 0;
  for (int i = 0; i < L->ntargets; i++) if (sk_crate_at(L, s, L->tx[i], L->ty[i]) < 0) return 0;
  return 1;
}
static int sk_heuristic(const level_t* L, const gstate_t* s) { /* :282-296 greedy crate->nearest remaining target */
  int tx[MAXOBJ], ty[MAXOBJ], nt = L->ntargets, distance = 0;
  for (int i = 0; i < nt; i++) { tx[i] = L->tx[i]; ty[i] = L->ty[i]; }
  for (int c = 0; c < L->nobj; c++) {
    int best = L->w + L->h, match = 0;
    for (int i = 0; i < nt; i++) {
      int d = abs(s->cx[c] - tx[i]) + abs(s->cy[c] - ty[i]);
      if (best > d) { match = i; best = d; }
    }
    distance += abs(tx[match] - s->cx[c]) + abs(ty[match] - s->cy[c]);
    for (int i = match; i + 1 < nt; i++) { tx[i] = tx[i + 1]; ty[i] = ty[i + 1]; }
    nt--;
  }
  return distance;
}
static int sk_update(const level_t* L, gstate_t* s, int dx, int dy) { /* :298-327, returns crateMove */
  if (sk_win(L, s)) return 0;
  int nx = s->px + dx, ny = s->py + dy;
  if (sk_movable(L, s, nx, ny)) { s->px = (int16_t)nx; s->py = (int16_t)ny; return 0; }
  int c = sk_crate_at(L, s, nx, ny);
  if (c >= 0) {
    int cx = s->cx[c] + dx, cy = s->cy[c] + dy;
    if (sk_movable(L, s, cx, cy)) {
      s->px = (int16_t)nx; s->py = (int16_t)ny; s->cx[c] = (uint8_t)cx; s->cy[c] = (uint8_t)cy;
      return 1;
    }
  }
  return 0;
}
static int sk_deadlocked(const level_t* L, const gstate_t* s) { /* :248-252 any crate on a deadlock cell */
  for (int i = 0; i < L->nobj; i++) if (L->deadlock[s->cy[i] * L->w + s->cx[i]]) return 1;
  return 0;
}
static int isign(int x) { return (x > 0) - (x < 0); } /* :204 sign lambda */
static void sk_init_deadlocks(level_t* L) { /* :203-246 intializeDeadlocks */
  int ncorner = 0, cx[MAXCELLS], cy[MAXCELLS];
  memset(L->deadlock, 0, sizeof(L->deadlock));
  for (int y = 0; y < L->h; y++)
    for (int x = 0; x < L->w; x++) {
      if (x == 0 || y == 0 || x == L->w - 1 || y == L->h - 1 || lv_solid(L, x, y)) continue;
      int up = lv_solid(L, x, y - 1), dn = lv_solid(L, x, y + 1), lf = lv_solid(L, x - 1, y), rt = lv_solid(L, x + 1, y);
      if ((up && lf) || (up && rt) || (dn && lf) || (dn && rt))
        if (!sk_target_at(L, x, y)) { cx[ncorner] = x; cy[ncorner++] = y; L->deadlock[y * L->w + x] = 1; }
    }
  for (int a = 0; a < ncorner; a++)
    for (int b = 0; b < ncorner; b++) {
      int dx = isign(cx[a] - cx[b]), dy = isign(cy[a] - cy[b]);
      if ((dx == 0 && dy == 0) || (dx != 0 && dy != 0)) continue;
      int wx[MAXD + 2], wy[MAXD + 2], nw = 0, x = cx[b], y = cy[b];
      if (dx != 0) {
        x += dx;
        while (x != cx[a]) {
          if (sk_target_at(L, x, y) || lv_solid(L, x, y) || (!lv_solid(L, x, y - 1) && !lv_solid(L, x, y + 1))) { nw = 0; break; }
          wx[nw] = x; wy[nw++] = y; x += dx;
        }
      }
      if (dy != 0) {
        y += dy;
        while (y != cy[a]) {
          if (sk_target_at(L, x, y) || lv_solid(L, x, y) || (!lv_solid(L, x - 1, y) && !lv_solid(L, x + 1, y))) { nw = 0; break; }
          wx[nw] = x; wy[nw++] = y; y += dy;
        }
      }
      for (int i = 0; i < nw; i++) L->deadlock[wy[i] * L->w + wx[i]] = 1;
    }
}

/* --- ddave -------------------------------------------------------------------------------------- */
static int lv_movable(const level_t* L, int x, int y) { /* ddave/engine.py:204-205, mdungeon/engine.py:201-202 */
  return !(x < 0 || y < 0 || x >= L->w || y >= L->h || lv_solid(L, x, y));
}
static int obj_at(const level_t* L, const gstate_t* s, int x, int y, int tmask) { /* first remaining object of a type in tmask */
  for (int i = 0; i < L->nobj; i++)
    if (((tmask >> L->otype[i]) & 1) && ((s->remain[i >> 3] >> (i & 7)) & 1) && L->ox[i] == x && L->oy[i] == y) return i;
  return -1;
}
static void dd_update_player(const level_t* L, gstate_t* s, int x, int y) { /* ddave/engine.py:225-242 */
  s->px = (int16_t)x; s->py = (int16_t)y;
  int i = obj_at(L, s, x, y, 1);
  if (i >= 0) { s->c0++; s->remain[i >> 3] &= (uint8_t)~(1u << (i & 7)); return; }
  for (int k = 0; k < L->nspikes; k++) if (L->sx[k] == x && L->sy[k] == y) { s->health = 0; return; }
  if (s->key_present && L->keyx == x && L->keyy == y) { s->c1++; s->key_present = 0; return; }
}
static int dd_win(const level_t* L, const gstate_t* s) { return s->c1 > 0 && s->px == L->doorx && s->py == L->doory; } /* :319-320 */
static int dm_lose(const gstate_t* s) { return s->health <= 0; } /* ddave :322-323, mdungeon :311-312 */
static void dd_update(const level_t* L, gstate_t* s, int dx, int dy) { /* ddave/engine.py:244-280 */
  if (dd_win(L, s) || dm_lose(s)) return;
  dy = dy < 0 ? -1 : 0;
  int ground = lv_solid(L, s->px, s->py + 1), ceiling = lv_solid(L, s->px, s->py - 1);
  int nx = s->px, ny = s->py;
  if (dx != 0) {
    if (lv_movable(L, nx + dx, ny)) nx += dx;
  } else if (dy == -1) {
    if (ground && !ceiling) { s->air = 3; s->jumps++; }
  }
  if (s->air > 1) {
    s->air--;
    if (lv_movable(L, nx, ny - 1)) ny--; else s->air = 1;
  } else if (s->air > 0 && s->air <= 1) {
    s->air--;
  } else {
    if (lv_movable(L, nx, ny + 1)) ny++;
  }
  dd_update_player(L, s, nx, ny);
}
static int dd_heuristic(const level_t* L, const gstate_t* s) { /* :294-299 */
  int d = abs(s->px - L->doorx) + abs(s->py - L->doory);
  if (s->key_present) d = abs(s->px - L->keyx) + abs(s->py - L->keyy) + (L->w + L->h);
  return d + 5 * (-s->c0);
}

/* --- mdungeon ----------------------------------------------------------------------------------- */
static void md_update_player(const level_t* L, gstate_t* s, int x, int y) { /* mdungeon/engine.py:222-252 */
  s->px = (int16_t)x; s->py = (int16_t)y;
  int i = obj_at(L, s, x, y, 1 << 0);
  if (i >= 0) {
    s->health += 2; s->c0++;
    if (s->health > 5) s->health = 5;
    s->remain[i >> 3] &= (uint8_t)~(1u << (i & 7));
    return;
  }
  i = obj_at(L, s, x, y, 1 << 1);
  if (i >= 0) { s->c1++; s->remain[i >> 3] &= (uint8_t)~(1u << (i & 7)); return; }
  i = obj_at(L, s, x, y, (1 << 2) | (1 << 3));
  if (i >= 0) {
    s->c2++;
    s->health -= (L->otype[i] == 2) ? 1 : 2;
    if (s->health < 0) s->health = 0;
    s->remain[i >> 3] &= (uint8_t)~(1u << (i & 7));
    return;
  }
}
static int md_win(const level_t* L, const gstate_t* s) { return s->px == L->doorx && s->py == L->doory; } /* :308-309 */
static void md_update(const level_t* L, gstate_t* s, int dx, int dy) { /* :254-270 */
  if (md_win(L, s) || dm_lose(s)) return;
  int nx = s->px + dx, ny = s->py + dy;
  if (lv_movable(L, nx, ny)) md_update_player(L, s, nx, ny);
}
static int md_heuristic(const level_t* L, const gstate_t* s) { /* :285-289 */
  return abs(s->px - L->doorx) + abs(s->py - L->doory) + 4 * (5 - s->health) + 4 * (-s->c1);
}

/* --- level construction: *_prob.py _run_game + engine.stringInitialize ---------------------------- */
static int level_init(level_t* L, gstate_t* s0, int game, const grid_t* g) {
  memset(L, 0, sizeof(*L));
  memset(s0, 0, sizeof(*s0));
  L->game = game; L->w = g->w + 2; L->h = g->h + 2;
  for (int y = 0; y < L->h; y++)
    for (int x = 0; x < L->w; x++) {
      int border = (x == 0 || y == 0 || x == L->w - 1 || y == L->h - 1);
      int t = border ? 1 : AT(g, x - 1, y - 1);
      L->solid[y * L->w + x] = (t == 1);
      if (border || t <= 1) continue;
      if (game == GAME_SOKOBAN) { /* tiles: 2 player '@', 3 crate '$', 4 target '.'  (sokoban_prob.py:86) */
        if (t == 2) { s0->px = (int16_t)x; s0->py = (int16_t)y; }
        if (t == 3) { if (L->nobj >= 64) return -1; s0->cx[L->nobj] = (uint8_t)x; s0->cy[L->nobj] = (uint8_t)y; L->nobj++; }
        if (t == 4) { if (L->ntargets >= MAXOBJ) return -1; L->tx[L->ntargets] = x; L->ty[L->ntargets++] = y; }
      } else if (game == GAME_DDAVE) { /* 2 player, 3 exit 'H', 4 diamond '$', 5 key 'V', 6 spike '*' (ddave_prob.py:98) */
        if (t == 2) { s0->px = (int16_t)x; s0->py = (int16_t)y; s0->health = 1; }
        if (t == 3) { L->doorx = x; L->doory = y; }
        if (t == 4) { if (L->nobj >= MAXOBJ) return -1; L->ox[L->nobj] = x; L->oy[L->nobj] = y; L->otype[L->nobj] = 0; s0->remain[L->nobj >> 3] |= (uint8_t)(1u << (L->nobj & 7)); L->nobj++; }
        if (t == 5) { L->keyx = x; L->keyy = y; s0->key_present = 1; }
        if (t == 6) { L->sx[L->nspikes] = x; L->sy[L->nspikes++] = y; }
      } else { /* mdungeon: 2 player, 3 exit, 4 potion '*', 5 treasure '$', 6 goblin 'g', 7 ogre 'o' (mdungeon_prob.py:101) */
        if (t == 2) { s0->px = (int16_t)x; s0->py = (int16_t)y; s0->health = 5; }
        if (t == 3) { L->doorx = x; L->doory = y; }
        if (t >= 4) {
          if (L->nobj >= MAXOBJ) return -1;
          L->ox[L->nobj] = x; L->oy[L->nobj] = y; L->otype[L->nobj] = t - 4;
          s0->remain[L->nobj >> 3] |= (uint8_t)(1u << (L->nobj & 7)); L->nobj++;
        }
      }
    }
  if (game == GAME_SOKOBAN) sk_init_deadlocks(L);
  return 0;
}

static int g_win(const level_t* L, const gstate_t* s) {
  return L->game == GAME_SOKOBAN ? sk_win(L, s) : L->game == GAME_DDAVE ? dd_win(L, s) : md_win(L, s);
}
static int g_heuristic(const level_t* L, const gstate_t* s) {
  return L->game == GAME_SOKOBAN ? sk_heuristic(L, s) : L->game == GAME_DDAVE ? dd_heuristic(L, s) : md_heuristic(L, s);
}
/* State.getKey equality (sokoban :329-335, ddave :282-292, mdungeon :272-283): fields of the key string that can
 * differ between two states of the same level. */
static int g_key_equal(const level_t* L, const gstate_t* a, const gstate_t* b) {
  if (a->px != b->px || a->py != b->py) return 0;
  if (L->game == GAME_SOKOBAN) return memcmp(a->cx, b->cx, (size_t)L->nobj) == 0 && memcmp(a->cy, b->cy, (size_t)L->nobj) == 0;
  if (a->health != b->health) return 0;
  if (L->game == GAME_DDAVE && a->key_present != b->key_present) return 0;
  return memcmp(a->remain, b->remain, sizeof(a->remain)) == 0;
}
static uint32_t g_key_hash(const level_t* L, const gstate_t* s) {
  uint32_t hsh = 2166136261u;
#define MIX(v) hsh = (hsh ^ (uint32_t)(v)) * 16777619u
  MIX(s->px); MIX(s->py);
  if (L->game == GAME_SOKOBAN) { for (int i = 0; i < L->nobj; i++) { MIX(s->cx[i]); MIX(s->cy[i]); } }
  else { MIX(s->health); MIX(s->key_present); for (int i = 0; i < (L->nobj + 7) / 8; i++) MIX(s->remain[i]); }
#undef MIX
  return hsh;
}

/* directions == child order: sokoban/mdungeon engine.py:3, ddave engine.py:3 */
static const int DIR_SM[4][2] = {{-1, 0}, {1, 0}, {0, -1}, {0, 1}};
static const int DIR_DD[4][2] = {{0, 0}, {-1, 0}, {1, 0}, {0, -1}};

typedef struct {
  node_t* nodes; int cap_nodes;
  int* heap; int* table; int table_size;
} search_ws_t;

static __thread search_ws_t g_ws;

static void ws_reserve(int power) {
  int cap = 4 * power + 8, tsz = 1;
  while (tsz < 2 * (power + 2)) tsz <<= 1;
  if (g_ws.cap_nodes < cap) {
    free(g_ws.nodes); free(g_ws.heap);
    g_ws.nodes = (node_t*)malloc(sizeof(node_t) * (size_t)cap);
    g_ws.heap = (int*)malloc(sizeof(int) * (size_t)cap);
    g_ws.cap_nodes = cap;
  }
  if (g_ws.table_size < tsz) { free(g_ws.table); g_ws.table = (int*)malloc(sizeof(int) * (size_t)tsz); g_ws.table_size = tsz; }
}

/* CPython heapq (Lib/heapq.py _siftdown/_siftup) on node indices; Node.__lt__ = strict < on h + balance*depth,
 * computed exactly as the integer 2*h + b*depth with b = 2*balance in {2,1,0} (engine.py Node.__lt__). */
static int prio(const node_t* n, int b) { return 2 * n->h + b * n->depth; }
static void heap_siftdown(int* heap, const node_t* nodes, int b, int startpos, int pos) {
  int item = heap[pos];
  while (pos > startpos) {
    int parentpos = (pos - 1) >> 1, parent = heap[parentpos];
    if (prio(&nodes[item], b) < prio(&nodes[parent], b)) { heap[pos] = parent; pos = parentpos; continue; }
    break;
  }
  heap[pos] = item;
}
static void heap_push(int* heap, int* n, const node_t* nodes, int b, int item) {
  heap[*n] = item; (*n)++;
  heap_siftdown(heap, nodes, b, 0, *n - 1);
}
static int heap_pop(int* heap, int* n, const node_t* nodes, int b) {
  int last = heap[--(*n)];
  if (*n == 0) return last;
  int ret = heap[0], pos = 0, endpos = *n, childpos = 1;
  heap[0] = last;
  while (childpos < endpos) { /* _siftup */
    int rightpos = childpos + 1;
    if (rightpos < endpos && !(prio(&nodes[heap[childpos]], b) < prio(&nodes[heap[rightpos]], b))) childpos = rightpos;
    heap[pos] = heap[childpos];
    pos = childpos;
    childpos = 2 * pos + 1;
  }
  heap[pos] = last;
  heap_siftdown(heap, nodes, b, 0, pos);
  return ret;
}

/* BFSAgent / AStarAgent.getSolution (sokoban :56-74,96-119; ddave/mdungeon :61-81,105-129).
 * b < 0 -> BFS (FIFO == node creation order).  Returns the index of the node that is `solState`;
 * *won = 1 if it is a winning node.  *iters_out = iterations used. */
static int search(const level_t* L, const gstate_t* s0, int b, int max_iter, int* won, int* iters_out) {
  node_t* nodes = g_ws.nodes;
  int* heap = g_ws.heap; int* table = g_ws.table;
  int tmask = g_ws.table_size - 1, nn = 0, nheap = 0, head = 0, iterations = 0, best = -1;
  const int check_lose = (L->game != GAME_SOKOBAN);
  const int (*dirs)[2] = (L->game == GAME_DDAVE) ? DIR_DD : DIR_SM;
  for (int i = 0; i <= tmask; i++) table[i] = -1;
  nodes[0].st = *s0; nodes[0].parent = -1; nodes[0].depth = 0; nodes[0].h = g_heuristic(L, s0); nn = 1;
  if (b >= 0) heap_push(heap, &nheap, nodes, b, 0);
  *won = 0;
  while ((iterations < max_iter || max_iter <= 0) && (b >= 0 ? nheap > 0 : head < nn)) {
    iterations++;
    int cur = (b >= 0) ? heap_pop(heap, &nheap, nodes, b) : head++;
    const gstate_t* cs = &nodes[cur].st;
    if (check_lose && dm_lose(cs)) continue;
    if (g_win(L, cs)) { *won = 1; *iters_out = iterations; return cur; }
    uint32_t slot = g_key_hash(L, cs) & (uint32_t)tmask;
    int seen = 0;
    while (table[slot] >= 0) {
      if (g_key_equal(L, &nodes[table[slot]].st, cs)) { seen = 1; break; }
      slot = (slot + 1) & (uint32_t)tmask;
    }
    if (seen) continue;
    if (best < 0 || nodes[cur].h < nodes[best].h) best = cur;
    else if (nodes[cur].h == nodes[best].h && nodes[cur].depth < nodes[best].depth) best = cur;
    table[slot] = cur;
    for (int d = 0; d < 4; d++) { /* Node.getChildren */
      if (nn >= g_ws.cap_nodes) break;
      node_t* ch = &nodes[nn];
      ch->st = *cs;
      if (L->game == GAME_SOKOBAN) {
        int crate_move = sk_update(L, &ch->st, dirs[d][0], dirs[d][1]);
        if (ch->st.px == cs->px && ch->st.py == cs->py) continue;
        if (crate_move && sk_deadlocked(L, &ch->st)) continue;
      } else if (L->game == GAME_DDAVE) {
        dd_update(L, &ch->st, dirs[d][0], dirs[d][1]);
      } else {
        md_update(L, &ch->st, dirs[d][0], dirs[d][1]);
      }
      ch->parent = cur; ch->depth = nodes[cur].depth + 1; ch->h = g_heuristic(L, &ch->st);
      if (b >= 0) heap_push(heap, &nheap, nodes, b, nn);
      nn++;
    }
  }
  *iters_out = iterations;
  return best;
}

/* *_prob.py _run_game.  out: dist_win, sol_length, and the solState's player counters. */
typedef struct { int dist_win, sol_length, jumps, c0, c1, c2; long iterations; } game_result_t;

static int run_game(int game, const grid_t* g, int power, game_result_t* r) {
  static __thread level_t L;
  gstate_t s0;
  /* pass order: sokoban_prob.py:110-122 BFS, A*(1), A*(.5), A*(0); ddave_prob.py:122-135 / mdungeon_prob.py:125-138
   * A*(1), A*(.5), A*(0), BFS */
  static const int ORDER_SK[4] = {-1, 2, 1, 0}, ORDER_DM[4] = {2, 1, 0, -1};
  const int* order = (game == GAME_SOKOBAN) ? ORDER_SK : ORDER_DM;
  if (level_init(&L, &s0, game, g) != 0) return -1;
  ws_reserve(power > 0 ? power : 5000);
  int won = 0, it = 0, node = -1;
  r->iterations = 0;
  for (int p = 0; p < 4; p++) {
    node = search(&L, &s0, order[p], power, &won, &it);
    r->iterations += it;
    if (won) break;
  }
  const node_t* nd = &g_ws.nodes[node];
  r->dist_win = won ? 0 : nd->h;
  r->sol_length = won ? nd->depth : 0;
  r->jumps = nd->st.jumps; r->c0 = nd->st.c0; r->c1 = nd->st.c1; r->c2 = nd->st.c2;
  return 0;
}

/* ------------------------------------------------------------------------------------------------
 * Problem.get_stats  (binary_prob.py:81-86, zelda_prob.py:80-112, sokoban_prob.py:133-145,
 * ddave_prob.py:149-169, mdungeon_prob.py:151-171)
 * ---------------------------------------------------------------------------------------------- */
static long g_solver_iterations = 0, g_solver_calls = 0;

/* smb (114 x 14 byte map, always-on A* play-through): restated in oracle/smb_oracle.c, linked into this library */
void smb_get_stats(const uint8_t* map, int w, int h, int power, int32_t* st);
double smb_get_reward(const int32_t* n, const int32_t* o, const double* w, const int32_t* ip);
int smb_episode_over(const int32_t* n);

static int get_stats(const pcgrl_config* cfg, const uint8_t* map, int32_t* st) {
  grid_t g = {cfg->width, cfg->height, map};
  const int W = cfg->width, H = cfg->height;
  for (int i = 0; i < S; i++) st[i] = 0;
  switch (cfg->problem) {
    case PCGRL_PROB_SMB: /* smb_prob.py:126-148 */
      smb_get_stats(map, W, H, cfg->solver_power, st);
      return 0;
    case PCGRL_PROB_BINARY: {
      static const int order[1] = {0};
      st[0] = calc_num_regions(&g, order, 1, 1u << 0);
      st[1] = calc_longest_path(&g, order, 1, 1u << 0);
      return 0;
    }
    case PCGRL_PROB_ZELDA: {
      static const int order[6] = {0, 2, 3, 5, 7, 6}; /* "empty","player","key","bat","spider","scorpion" */
      static __thread int dmap[MAXCELLS];
      st[0] = count_tiles(&g, 1u << 2); st[1] = count_tiles(&g, 1u << 3); st[2] = count_tiles(&g, 1u << 4);
      st[3] = count_tiles(&g, (1u << 5) | (1u << 6) | (1u << 7));
      st[4] = calc_num_regions(&g, order, 6, 0xEDu);
      if (st[0] == 1 && st[4] == 1) {
        int px = 0, py = 0, kx = 0, ky = 0, dx = 0, dy = 0;
        for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) {
          int t = AT(&g, x, y);
          if (t == 2) { px = x; py = y; } else if (t == 3) { kx = x; ky = y; } else if (t == 4) { dx = x; dy = y; }
        }
        if (st[3] > 0) {
          run_dikjstra(&g, px, py, 0xE5u, dmap, NULL); /* empty, player, bat, spider, scorpion */
          int min_dist = W * H;
          for (int i = 0; i < W * H; i++)
            if (map[i] >= 5 && dmap[i] > 0 && dmap[i] < min_dist) min_dist = dmap[i];
          st[5] = min_dist;
        }
        if (st[1] == 1 && st[2] == 1) {
          run_dikjstra(&g, px, py, 0xEDu, dmap, NULL); /* + key */
          st[6] += dmap[ky * W + kx];
          run_dikjstra(&g, kx, ky, 0xFDu, dmap, NULL); /* + door */
          st[6] += dmap[dy * W + dx];
        }
      }
      return 0;
    }
    case PCGRL_PROB_SOKOBAN: {
      static const int order[4] = {0, 2, 3, 4};
      st[0] = count_tiles(&g, 1u << 2); st[1] = count_tiles(&g, 1u << 3); st[2] = count_tiles(&g, 1u << 4);
      st[3] = calc_num_regions(&g, order, 4, 0x1Du);
      st[4] = W * H * (W + H);
      st[5] = 0;
      if (st[0] == 1 && st[1] == st[2] && st[1] > 0 && st[3] == 1) {
        game_result_t r;
        if (run_game(GAME_SOKOBAN, &g, cfg->solver_power, &r) != 0) return -1;
        st[4] = r.dist_win; st[5] = r.sol_length;
        __atomic_fetch_add(&g_solver_iterations, r.iterations, __ATOMIC_RELAXED); __atomic_fetch_add(&g_solver_calls, 1, __ATOMIC_RELAXED);
      }
      return 0;
    }
    case PCGRL_PROB_DDAVE: {
      static const int order[5] = {0, 2, 4, 5, 3}; /* "empty","player","diamond","key","exit" */
      st[0] = count_tiles(&g, 1u << 2);
      st[1] = get_floor_dist(&g, 1u << 2, 1u << 1);
      st[2] = count_tiles(&g, 1u << 3); st[3] = count_tiles(&g, 1u << 4); st[4] = count_tiles(&g, 1u << 5);
      st[5] = count_tiles(&g, 1u << 6);
      st[6] = calc_num_regions(&g, order, 5, 0x3Du);
      st[9] = W * H;
      if (st[0] == 1 && st[2] == 1 && st[4] == 1 && st[6] == 1) {
        game_result_t r;
        if (run_game(GAME_DDAVE, &g, cfg->solver_power, &r) != 0) return -1;
        st[9] = r.dist_win; st[10] = r.sol_length; st[7] = r.jumps; st[8] = r.c0;
        __atomic_fetch_add(&g_solver_iterations, r.iterations, __ATOMIC_RELAXED); __atomic_fetch_add(&g_solver_calls, 1, __ATOMIC_RELAXED);
      }
      return 0;
    }
    case PCGRL_PROB_MDUNGEON: {
      static const int order[7] = {0, 2, 3, 4, 5, 6, 7};
      st[0] = count_tiles(&g, 1u << 2); st[1] = count_tiles(&g, 1u << 3); st[2] = count_tiles(&g, 1u << 4);
      st[3] = count_tiles(&g, 1u << 5); st[4] = count_tiles(&g, (1u << 6) | (1u << 7));
      st[5] = calc_num_regions(&g, order, 7, 0xFDu);
      st[9] = W * H;
      if (st[0] == 1 && st[1] == 1 && st[5] == 1) {
        game_result_t r;
        if (run_game(GAME_MDUNGEON, &g, cfg->solver_power, &r) != 0) return -1;
        st[9] = r.dist_win; st[10] = r.sol_length; st[6] = r.c0; st[7] = r.c1; st[8] = r.c2;
        __atomic_fetch_add(&g_solver_iterations, r.iterations, __ATOMIC_RELAXED); __atomic_fetch_add(&g_solver_calls, 1, __ATOMIC_RELAXED);
      }
      return 0;
    }
  }
  return -1;
}

/* Problem.get_reward: fp64, terms summed left to right in the reference's order (see header). */
static double get_reward(const pcgrl_config* cfg, const int32_t* n, const int32_t* o) {
  const double* w = cfg->reward_weight;
  const double INF = INFINITY;
  const int32_t* ip = cfg->iparam;
  switch (cfg->problem) {
    case PCGRL_PROB_SMB: /* smb_prob.py:150-172 */
      return smb_get_reward(n, o, w, ip);
    case PCGRL_PROB_BINARY: /* binary_prob.py:98-106 */
      return range_reward(n[0], o[0], 1, 1) * w[0] + range_reward(n[1], o[1], INF, INF) * w[1];
    case PCGRL_PROB_ZELDA: /* zelda_prob.py:124-142 */
      return range_reward(n[0], o[0], 1, 1) * w[0] + range_reward(n[1], o[1], 1, 1) * w[1] +
             range_reward(n[2], o[2], 1, 1) * w[2] + range_reward(n[3], o[3], 2, ip[0]) * w[3] +
             range_reward(n[4], o[4], 1, 1) * w[4] + range_reward(n[5], o[5], ip[1], INF) * w[5] +
             range_reward(n[6], o[6], INF, INF) * w[6];
    case PCGRL_PROB_SOKOBAN: /* sokoban_prob.py:157-175 */
      return range_reward(n[0], o[0], 1, 1) * w[0] + range_reward(n[1], o[1], 1, ip[0]) * w[1] +
             range_reward(n[2], o[2], 1, ip[0]) * w[2] + range_reward(n[3], o[3], 1, 1) * w[3] +
             range_reward(abs(n[1] - n[2]), abs(o[1] - o[2]), -INF, -INF) * w[4] +
             range_reward(n[4], o[4], -INF, -INF) * w[5] + range_reward(n[5], o[5], INF, INF) * w[6];
    case PCGRL_PROB_DDAVE: /* ddave_prob.py:181-205 */
      return range_reward(n[0], o[0], 1, 1) * w[0] + range_reward(n[1], o[1], 0, 0) * w[1] +
             range_reward(n[2], o[2], 1, 1) * w[2] + range_reward(n[5], o[5], ip[1], INF) * w[3] +
             range_reward(n[3], o[3], -INF, ip[0]) * w[4] + range_reward(n[4], o[4], 1, 1) * w[5] +
             range_reward(n[6], o[6], 1, 1) * w[6] + range_reward(n[7], o[7], INF, INF) * w[7] +
             range_reward(n[9], o[9], -INF, -INF) * w[8] + range_reward(n[10], o[10], INF, INF) * w[9];
    case PCGRL_PROB_MDUNGEON: /* mdungeon_prob.py:183-205 */
      return range_reward(n[0], o[0], 1, 1) * w[0] + range_reward(n[1], o[1], 1, 1) * w[1] +
             range_reward(n[4], o[4], 1, ip[0]) * w[2] + range_reward(n[3], o[3], -INF, ip[2]) * w[3] +
             range_reward(n[2], o[2], -INF, ip[1]) * w[4] + range_reward(n[5], o[5], 1, 1) * w[5] +
             range_reward(n[8], o[8], INF, INF) * w[6] + range_reward(n[9], o[9], -INF, -INF) * w[7] +
             range_reward(n[10], o[10], INF, INF) * w[8];
  }
  return 0;
}

/* Problem.get_episode_over */
static int episode_over(const pcgrl_config* cfg, const int32_t* n, const int32_t* start) {
  const int32_t* ip = cfg->iparam;
  switch (cfg->problem) {
    case PCGRL_PROB_SMB: return smb_episode_over(n);                                /* smb_prob.py:174-175 */
    case PCGRL_PROB_BINARY: return n[0] == 1 && n[1] - start[1] >= ip[0];          /* binary_prob.py:119-120 */
    case PCGRL_PROB_ZELDA: return n[5] >= ip[1] && n[6] >= ip[2];                   /* zelda_prob.py:155-156 */
    case PCGRL_PROB_SOKOBAN: return n[5] >= ip[1];                                  /* sokoban_prob.py:188-189 */
    case PCGRL_PROB_DDAVE: return n[10] >= ip[3] && n[7] > ip[2];                   /* ddave_prob.py:218-220 */
    case PCGRL_PROB_MDUNGEON:                                                       /* mdungeon_prob.py:218-221 */
      return n[10] >= ip[3] && n[4] > 0 && (double)n[8] / (double)(n[4] > 1 ? n[4] : 1) > cfg->dparam[0];
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------------
 * PcgrlEnv.reset / step
 * ---------------------------------------------------------------------------------------------- */
/* helper.py:310-312 gen_random_map + :343-352 get_int_prob + RandomState.choice(p=...) */
static void gen_random_map(const pcgrl_config* cfg, uint32_t* rng, const double* tile_prob, uint8_t* map) {
  const int T = cfg->num_tiles;
  double p[PCGRL_MAX_TILES], cdf[PCGRL_MAX_TILES], total = 0.0, acc = 0.0;
  for (int t = 0; t < T; t++) total += tile_prob[t];
  for (int t = 0; t < T; t++) p[t] = tile_prob[t] / total;
  for (int t = 0; t < T; t++) { acc += p[t]; cdf[t] = acc; }
  for (int t = 0; t < T; t++) cdf[t] /= cdf[T - 1];
  for (int i = 0; i < cfg->width * cfg->height; i++) {
    double u = mt_double(rng);
    int k = 0;
    while (k < T && cdf[k] <= u) k++; /* searchsorted(side='right') */
    map[i] = (uint8_t)k;
  }
}

static int env_reset(const pcgrl_config* cfg, const pcgrl_buffers* b, int i) { /* pcgrl_env.py:66-76 */
  const int W = cfg->width, H = cfg->height, cells = W * H;
  uint8_t* map = b->map + (size_t)i * cells;
  uint32_t* rng_rep = b->rng + (size_t)i * 2 * PCGRL_MT_WORDS;
  uint32_t* rng_prob = rng_rep + PCGRL_MT_WORDS;
  double* tp = b->tile_prob + (size_t)i * PCGRL_MAX_TILES;
  b->changes[i] = 0;
  b->iteration[i] = 0;
  /* representation.py:40-45 */
  if ((cfg->flags & PCGRL_FLAG_RANDOM_START) || !b->start_valid[i]) {
    gen_random_map(cfg, rng_rep, tp, map);
    memcpy(b->start_map + (size_t)i * cells, map, (size_t)cells);
    b->start_valid[i] = 1;
  } else {
    memcpy(map, b->start_map + (size_t)i * cells, (size_t)cells);
  }
  if (cfg->representation != PCGRL_REP_WIDE) { /* narrow_rep.py:28-31, turtle_rep.py:30-33 */
    b->pos[2 * i + 0] = (uint8_t)mt_randint(rng_rep, W);
    b->pos[2 * i + 1] = (uint8_t)mt_randint(rng_rep, H);
  }
  if (get_stats(cfg, map, b->stats + (size_t)i * S) != 0) return -1;
  memcpy(b->start_stats + (size_t)i * S, b->stats + (size_t)i * S, sizeof(int32_t) * S); /* problem.py:45-46 */
  if (cfg->problem == PCGRL_PROB_BINARY && (cfg->flags & PCGRL_FLAG_RANDOM_PROBS)) { /* binary_prob.py:68-72 */
    tp[0] = mt_double(rng_prob);
    tp[1] = 1 - tp[0];
  }
  {
    const size_t hb = (cfg->flags & PCGRL_FLAG_HEAT_U16) ? 2 : 1;
    memset((uint8_t*)b->heatmap + hb * (size_t)i * cells, 0, hb * (size_t)cells);
  }
  return 0;
}

static int env_step(const pcgrl_config* cfg, const pcgrl_buffers* b, const int32_t* actions, int i) { /* pcgrl_env.py:129-150 */
  const int W = cfg->width, H = cfg->height, cells = W * H;
  uint8_t* map = b->map + (size_t)i * cells;
  uint32_t* rng_rep = b->rng + (size_t)i * 2 * PCGRL_MT_WORDS;
  int32_t* stats = b->stats + (size_t)i * S;
  int32_t old_stats[S];
  int change = 0, x = b->pos[2 * i], y = b->pos[2 * i + 1];
  b->iteration[i] += 1;
  memcpy(old_stats, stats, sizeof(old_stats));
  if (cfg->representation == PCGRL_REP_NARROW) { /* narrow_rep.py:99-114 */
    int a = actions[i];
    if (a > 0) {
      change += (map[y * W + x] != a - 1);
      map[y * W + x] = (uint8_t)(a - 1);
    }
    if (cfg->flags & PCGRL_FLAG_RANDOM_TILE) {
      x = mt_randint(rng_rep, W);
      y = mt_randint(rng_rep, H);
    } else {
      x += 1;
      if (x >= W) { x = 0; y += 1; if (y >= H) y = 0; }
    }
  } else if (cfg->representation == PCGRL_REP_TURTLE) { /* turtle_rep.py:101-129 */
    static const int TDX[4] = {-1, 1, 0, 0}, TDY[4] = {0, 0, -1, 1};
    int a = actions[i], warp = (cfg->flags & PCGRL_FLAG_WARP) != 0;
    if (a < 4) {
      x += TDX[a];
      if (x < 0) x = warp ? x + W : 0;
      if (x >= W) x = warp ? x - W : W - 1;
      y += TDY[a];
      if (y < 0) y = warp ? y + H : 0;
      if (y >= H) y = warp ? y - H : H - 1;
    } else {
      change = (map[y * W + x] != a - 4);
      map[y * W + x] = (uint8_t)(a - 4);
    }
  } else if (cfg->representation == PCGRL_REP_WIDE) { /* wide_rep.py:67-70 */
    x = actions[3 * i]; y = actions[3 * i + 1];
    int v = actions[3 * i + 2];
    change = (map[y * W + x] != v);
    map[y * W + x] = (uint8_t)v;
  } else if (cfg->representation == PCGRL_REP_NARROWCAST || cfg->representation == PCGRL_REP_NARROWMULTI) {
    if (cfg->representation == PCGRL_REP_NARROWCAST) { /* narrow_cast_rep.py:36-59 */
      int type = actions[2 * i], value = actions[2 * i + 1];
      if (type == 1) {
        change += (map[y * W + x] != value);
        map[y * W + x] = (uint8_t)value;
      } else if (type == 2) {
        int low_y = y - 1 > 0 ? y - 1 : 0, high_y = y + 2 < H ? y + 2 : H;
        int low_x = x - 1 > 0 ? x - 1 : 0, high_x = x + 2 < W ? x + 2 : W;
        for (int yy = low_y; yy < high_y; yy++)
          for (int xx = low_x; xx < high_x; xx++) {
            change += (map[yy * W + xx] != value);
            map[yy * W + xx] = (uint8_t)value;
          }
      }
    } else { /* narrow_multi_rep.py:39-59 */
      const int32_t* a = actions + 9 * i;
      int low_y = y - 1 > 0 ? y - 1 : 0, high_y = y + 2 < H ? y + 2 : H;
      int low_x = x - 1 > 0 ? x - 1 : 0, high_x = x + 2 < W ? x + 2 : W;
      for (int k = 0; k < 9; k++) {
        int xx = x + (k % 3) - 1, yy = y + (k / 3) - 1;
        if (xx >= low_x && xx < high_x && yy >= low_y && yy < high_y && a[k] > 0) {
          change += (map[yy * W + xx] != a[k] - 1);
          map[yy * W + xx] = (uint8_t)(a[k] - 1);
        }
      }
    }
    if (cfg->flags & PCGRL_FLAG_RANDOM_TILE) { /* narrow_rep.py cursor advance, inherited */
      x = mt_randint(rng_rep, W);
      y = mt_randint(rng_rep, H);
    } else {
      x += 1;
      if (x >= W) { x = 0; y += 1; if (y >= H) y = 0; }
    }
  } else { /* turtle_cast_rep.py:38-76 */
    static const int TDX[4] = {-1, 1, 0, 0}, TDY[4] = {0, 0, -1, 1};
    int type = actions[2 * i], value = actions[2 * i + 1], warp = (cfg->flags & PCGRL_FLAG_WARP) != 0;
    if (type < 4) {
      x += TDX[type];
      if (x < 0) x = warp ? x + W : 0;
      if (x >= W) x = warp ? x - W : W - 1;
      y += TDY[type];
      if (y < 0) y = warp ? y + H : 0;
      if (y >= H) y = warp ? y - H : H - 1;
    } else if (type == 4) {
      change = (map[y * W + x] != value);
      map[y * W + x] = (uint8_t)value;
    } else if (type == 5) {
      int low_y = y - 1 > 0 ? y - 1 : 0, high_y = y + 2 < H ? y + 2 : H;
      int low_x = x - 1 > 0 ? x - 1 : 0, high_x = x + 2 < W ? x + 2 : W;
      for (int yy = low_y; yy < high_y; yy++)
        for (int xx = low_x; xx < high_x; xx++) {
          change += (map[yy * W + xx] != value);
          map[yy * W + xx] = (uint8_t)value;
        }
    }
  }
  if (cfg->representation != PCGRL_REP_WIDE) { b->pos[2 * i] = (uint8_t)x; b->pos[2 * i + 1] = (uint8_t)y; }
  if (change > 0) {
    b->changes[i] += change;
    if (cfg->flags & PCGRL_FLAG_HEAT_U16) ((uint16_t*)b->heatmap)[(size_t)i * cells + y * W + x] += 1;
    else ((uint8_t*)b->heatmap)[(size_t)i * cells + y * W + x] += 1;
    if (get_stats(cfg, map, stats) != 0) return -1;
  }
  b->reward[i] = get_reward(cfg, stats, old_stats);
  int done = episode_over(cfg, stats, b->start_stats + (size_t)i * S) || b->changes[i] >= cfg->max_changes ||
             b->iteration[i] >= cfg->max_iterations;
  b->done[i] = (uint8_t)done;
  memcpy(b->info_stats + (size_t)i * S, stats, sizeof(int32_t) * S);
  b->info_stats[(size_t)i * S + PCGRL_INFO_ITERATION] = b->iteration[i]; /* pcgrl_env.py:144-145, before any auto-reset */
  b->info_stats[(size_t)i * S + PCGRL_INFO_CHANGES] = b->changes[i];
  if (cfg->problem == PCGRL_PROB_BINARY) /* binary_prob.py:137 "path-imp" */
    b->info_stats[(size_t)i * S + 2] = stats[1] - b->start_stats[(size_t)i * S + 1];
  if (done && (cfg->flags & PCGRL_FLAG_AUTO_RESET)) return env_reset(cfg, b, i);
  return 0;
}

/* ------------------------------------------------------------------------------------------------
 * exported API (host pointers); nthreads > 1 uses OpenMP over env index
 * ---------------------------------------------------------------------------------------------- */
int oracle_reset(const pcgrl_config* cfg, const pcgrl_buffers* b, const uint8_t* mask, int n, int nthreads) {
  int err = 0;
#pragma omp parallel for schedule(dynamic, 16) num_threads(nthreads > 0 ? nthreads : 1) reduction(| : err)
  for (int i = 0; i < n; i++)
    if (!mask || mask[i]) err |= (env_reset(cfg, b, i) != 0);
  return err ? -1 : 0;
}

int oracle_step(const pcgrl_config* cfg, const pcgrl_buffers* b, const int32_t* actions, int n, int nthreads) {
  int err = 0;
#pragma omp parallel for schedule(dynamic, 16) num_threads(nthreads > 0 ? nthreads : 1) reduction(| : err)
  for (int i = 0; i < n; i++) err |= (env_step(cfg, b, actions, i) != 0);
  return err ? -1 : 0;
}

int oracle_get_stats(const pcgrl_config* cfg, const uint8_t* maps, int32_t* stats_out, int n, int nthreads) {
  int err = 0;
  const int cells = cfg->width * cfg->height;
#pragma omp parallel for schedule(dynamic, 4) num_threads(nthreads > 0 ? nthreads : 1) reduction(| : err)
  for (int i = 0; i < n; i++) err |= (get_stats(cfg, maps + (size_t)i * cells, stats_out + (size_t)i * S) != 0);
  return err ? -1 : 0;
}

void oracle_seed(const pcgrl_buffers* b, const uint32_t* seeds, int n) {
  for (int i = 0; i < n; i++) {
    mt_seed(b->rng + (size_t)i * 2 * PCGRL_MT_WORDS, seeds[i]);
    mt_seed(b->rng + (size_t)i * 2 * PCGRL_MT_WORDS + PCGRL_MT_WORDS, seeds[i]);
  }
}

/* raw RNG access for tests/test_oracle_golden.py */
void oracle_rng_seed(uint32_t* st, uint32_t seed) { mt_seed(st, seed); }
void oracle_rng_doubles(uint32_t* st, double* out, int n) { for (int i = 0; i < n; i++) out[i] = mt_double(st); }
int oracle_rng_randint(uint32_t* st, int n) { return mt_randint(st, n); }
void oracle_solver_counters(long* iterations, long* calls) { *iterations = g_solver_iterations; *calls = g_solver_calls; }

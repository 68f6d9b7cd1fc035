// pcgrl_host_twin.cuh -- host twins of the graph-only problems (binary, zelda): Problem.get_stats on the CPU
// (the solver problems: pcgrl_solver_host.cuh, which reuses the row-array helpers below).
//
// The device code gives one map row to one warp lane and propagates BFS frontiers with shuffles; a host thread has no
// lanes, so the same BITBOARD ALGORITHM (pcgrl_device.cuh / pcgrl_problems.cuh: one word per row, wave = dilation & passable
// & ~visited, single-cell / two-cell component shortcuts, double sweep per component with the reference's first-cell
// tie-breaks) is written once more over `uint32_t rows[32]` arrays.  It is deliberately NOT the oracle's algorithm
// (FIFO-queue flood fill, per-tile distance maps): tests/test_host_twins.py replays the reference's golden trajectories
// through it, which checks the bitboard formulation itself without a GPU.  The env loop around it (Representation.update
// on the byte map, gen_random_map, MT19937) is the scalar code shared with smb (pcgrl_smb.cuh); rewards and termination
// are the `__host__ __device__` functions of pcgrl_problems.cuh.
// Reference: helper.py:16-23,150-154,170-207,222-264; binary_prob.py:81-86; zelda_prob.py:80-112.
#pragma once
#include <stdint.h>
#include <string.h>

#include "../../include/pcgrl_b200.h"

namespace pcgrl_host {

struct Rows { uint32_t r[32]; };

static inline Rows zero_rows() { Rows z; memset(z.r, 0, sizeof(z.r)); return z; }
static inline bool any(const Rows& a) { uint32_t o = 0; for (int i = 0; i < 32; i++) o |= a.r[i]; return o != 0u; }
static inline int popc(const Rows& a) { int n = 0; for (int i = 0; i < 32; i++) n += __builtin_popcount(a.r[i]); return n; }
static inline Rows band(const Rows& a, const Rows& b) { Rows o; for (int i = 0; i < 32; i++) o.r[i] = a.r[i] & b.r[i]; return o; }
static inline Rows bandn(const Rows& a, const Rows& b) { Rows o; for (int i = 0; i < 32; i++) o.r[i] = a.r[i] & ~b.r[i]; return o; }
static inline Rows bor(const Rows& a, const Rows& b) { Rows o; for (int i = 0; i < 32; i++) o.r[i] = a.r[i] | b.r[i]; return o; }
static inline Rows dilate(const Rows& f) {  // 4-neighbourhood, includes f
  Rows o;
  for (int i = 0; i < 32; i++) o.r[i] = f.r[i] | (f.r[i] << 1) | (f.r[i] >> 1) | (i > 0 ? f.r[i - 1] : 0u) | (i < 31 ? f.r[i + 1] : 0u);
  return o;
}
static inline Rows neighbours(const Rows& f) {
  Rows o;
  for (int i = 0; i < 32; i++) o.r[i] = (f.r[i] << 1) | (f.r[i] >> 1) | (i > 0 ? f.r[i - 1] : 0u) | (i < 31 ? f.r[i + 1] : 0u);
  return o;
}
// row-major-first cell of m as a one-bit board (helper.py:16-23 location order)
static inline bool first_seed(const Rows& m, Rows& seed) {
  seed = zero_rows();
  for (int i = 0; i < 32; i++)
    if (m.r[i]) { seed.r[i] = m.r[i] & (0u - m.r[i]); return true; }
  return false;
}

// BFS from seed over pass (helper.py:222-237): eccentricity, visited set, last non-empty frontier
static inline int bfs_ecc(const Rows& seed, const Rows& pass, Rows& visited, Rows& last) {
  Rows f = seed, vis = seed;
  int d = 0;
  while (true) {
    const Rows n = bandn(band(dilate(f), pass), vis);
    if (!any(n)) break;
    vis = bor(vis, n);
    f = n;
    d++;
  }
  visited = vis;
  last = f;
  return d;
}
static inline Rows flood(const Rows& seed, const Rows& pass) {
  Rows v = seed;
  while (true) {
    const Rows n = band(dilate(v), pass);
    bool same = true;
    for (int i = 0; i < 32; i++) same = same && (n.r[i] == v.r[i]);
    if (same) return v;
    v = n;
  }
}
// distance from seed to target over pass; -1 if unreachable or the seed is not passable (zelda_prob.py:104-110)
static inline int bfs_dist_to(const Rows& seed, const Rows& target, const Rows& pass) {
  Rows f = band(seed, pass), vis = f;
  if (!any(f)) return -1;
  if (any(band(f, target))) return 0;
  int d = 0;
  while (true) {
    const Rows n = bandn(band(dilate(f), pass), vis);
    if (!any(n)) return -1;
    d++;
    if (any(band(n, target))) return d;
    vis = bor(vis, n);
    f = n;
  }
}
static inline int count_regions(const Rows& pass) {  // helper.py:197-207
  const Rows iso = bandn(pass, neighbours(pass));
  int regions = popc(iso);
  Rows remaining = bandn(pass, iso), seed;
  while (first_seed(remaining, seed)) {
    remaining = bandn(remaining, flood(seed, remaining));
    regions++;
  }
  return regions;
}
// helper.py:197-207 + :250-264 fused, with the shortcuts of regions_and_longest_path (pcgrl_device.cuh)
static inline void regions_and_longest_path(const Rows& pass, int& regions_out, int& path_out) {
  Rows iso, one, hd, vd, dominoes;
  for (int i = 0; i < 32; i++) {
    const uint32_t p = pass.r[i], l = p << 1, r = p >> 1, up = i > 0 ? pass.r[i - 1] : 0u, dn = i < 31 ? pass.r[i + 1] : 0u;
    iso.r[i] = p & ~(l | r | up | dn);
    one.r[i] = p & (l ^ r ^ up ^ dn) & ~((l & r) | (up & dn) | ((l ^ r) & (up ^ dn)));
  }
  for (int i = 0; i < 32; i++) {
    hd.r[i] = one.r[i] & (one.r[i] >> 1);
    vd.r[i] = one.r[i] & (i < 31 ? one.r[i + 1] : 0u);
  }
  for (int i = 0; i < 32; i++) dominoes.r[i] = hd.r[i] | (hd.r[i] << 1) | vd.r[i] | (i > 0 ? vd.r[i - 1] : 0u);
  const int ndom = popc(hd) + popc(vd);
  int regions = popc(iso) + ndom, best = ndom > 0 ? 1 : 0;
  Rows remaining = bandn(bandn(pass, iso), dominoes), seed;
  while (first_seed(remaining, seed)) {
    Rows visited, last;
    const int d1 = bfs_ecc(seed, remaining, visited, last);
    remaining = bandn(remaining, visited);
    regions++;
    if (2 * d1 > best) {
      first_seed(last, seed);
      Rows v2, l2;
      const int d2 = bfs_ecc(seed, visited, v2, l2);
      if (d2 > best) best = d2;
    }
  }
  regions_out = regions;
  path_out = best;
}

static inline Rows type_rows(const uint8_t* map, int W, int H, unsigned types) {
  Rows o = zero_rows();
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++)
      if ((types >> map[y * W + x]) & 1u) o.r[y] |= 1u << x;
  return o;
}

// Problem.get_stats for binary (binary_prob.py:81-86) and zelda (zelda_prob.py:80-112); false for other problems
static inline bool get_stats(const pcgrl_config* cfg, const uint8_t* map, int32_t* st) {
  const int W = cfg->width, H = cfg->height;
  for (int i = 0; i < PCGRL_MAX_STATS; i++) st[i] = 0;
  if (cfg->problem == PCGRL_PROB_BINARY) {
    int regions, path;
    regions_and_longest_path(type_rows(map, W, H, 0x01u), regions, path);
    st[0] = regions; st[1] = path;
    return true;
  }
  if (cfg->problem == PCGRL_PROB_ZELDA) {
    const Rows player = type_rows(map, W, H, 0x04u), key = type_rows(map, W, H, 0x08u), door = type_rows(map, W, H, 0x10u);
    const Rows enemies = type_rows(map, W, H, 0xE0u);
    st[0] = popc(player); st[1] = popc(key); st[2] = popc(door); st[3] = popc(enemies);
    st[4] = count_regions(type_rows(map, W, H, 0xEDu));
    if (st[0] == 1 && st[4] == 1) {
      if (st[3] > 0) {  // nearest enemy: first wave (d > 0) that touches one; key and door block
        const Rows pass = type_rows(map, W, H, 0xE5u);
        Rows f = player, vis = player;
        int d = 0, min_dist = W * H;
        while (true) {
          const Rows n = bandn(band(dilate(f), pass), vis);
          if (!any(n)) break;
          d++;
          if (any(band(n, enemies))) { min_dist = d; break; }
          vis = bor(vis, n);
          f = n;
        }
        st[5] = min_dist;
      }
      if (st[1] == 1 && st[2] == 1) {
        st[6] += bfs_dist_to(player, key, type_rows(map, W, H, 0xEDu));
        st[6] += bfs_dist_to(key, door, type_rows(map, W, H, 0xFDu));
      }
    }
    return true;
  }
  return false;
}

}  // namespace pcgrl_host

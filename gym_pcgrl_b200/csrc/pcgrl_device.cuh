// pcgrl_device.cuh -- warp-level building blocks of the batched PcgrlEnv hot path (sm_100a).
//
// Execution model: ONE WARP PER ENVIRONMENT.  Lane r owns map row r as bitboards (bit x <-> column x):
// three bit-planes of the tile index (tiles < 8).  Graph routines are frontier propagations
//   next = (f | f<<1 | f>>1 | shfl_up(f) | shfl_down(f)) & passable
// with warp votes for termination, so a BFS wave costs ~10 warp instructions and no memory traffic.
// The reference's Python equivalents are cited per function (G = gym_pcgrl/envs).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pcgrl_b200.h"

#define FULL_MASK 0xffffffffu
#define PCGRL_WARPS_PER_BLOCK 4
#define PCGRL_SBITS_STRIDE 34 /* 33 words used per plane (chunks + 1 guard) */

namespace pcgrl {

// ------------------------------------------------------------------------------------------------
// per-warp shared scratch
// ------------------------------------------------------------------------------------------------
struct __align__(8) WarpSmem {
  uint32_t bits[3 * PCGRL_SBITS_STRIDE];  // ballot words of the three tile bit-planes (row-major bit stream)
  uint32_t draws[512];                    // tempered MT19937 outputs for one 256-cell segment of gen_random_map
  uint32_t mt[624];                       // MT19937 key staged here for the duration of a reset (twist + ~2*H*W draws)
};

// ------------------------------------------------------------------------------------------------
// MT19937 (numpy legacy RandomState) -- state lives in HBM: 624 words + word 624 = position.
// The warp caches 32 consecutive tempered words in registers; a draw is one shuffle.
// numpy: random/src/mt19937/mt19937.c (mt19937_gen), _legacy random_sample, _bounded_integers masked rejection.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  return y;
}

// In-place twist by one warp in three wide phases + one scalar step.  new[i] depends on old[i], old[i+1] and
// (i < 227 ? old[i+397] : new[i-227]); inside [0,227), [227,454) and [454,623) every input is either untouched by the
// phase or produced by an earlier phase, so each phase reads all its inputs (24 loads in flight per lane), syncs, and
// writes -- identical to numpy's sequential mt19937_gen loop (checked against RandomState in tests/test_host_cpu.py).
template <int LO, int HI, int SRC_OFF>
__device__ __forceinline__ void mt_twist_phase(uint32_t* k, int lane) {
  uint32_t out[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const int i = LO + j * 32 + lane;
    out[j] = 0;
    if (i < HI) {
      const uint32_t a = k[i], b = k[i + 1], c = k[i + SRC_OFF];
      const uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu);
      out[j] = c ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
  }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const int i = LO + j * 32 + lane;
    if (i < HI) k[i] = out[j];
  }
  __syncwarp();
}

__device__ __noinline__ void mt_twist_warp(uint32_t* k, int lane) {  // one copy: every draw site may reach it
  __syncwarp();
  mt_twist_phase<0, 227, 397>(k, lane);
  mt_twist_phase<227, 454, -227>(k, lane);
  mt_twist_phase<454, 623, -227>(k, lane);
  if (lane == 0) {
    const uint32_t y = (k[623] & 0x80000000u) | (k[0] & 0x7fffffffu);
    k[623] = k[396] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
  }
  __syncwarp();
}

struct WarpRng {
  uint32_t* st;
  int pos;         // numpy `pos`
  int base;        // index held by lane 0 of `cache`, or < -31 when invalid
  uint32_t cache;  // tempered st[base + lane]
  bool dirty;

  // known_pos >= 0: the caller has already loaded s[624]
  __device__ __forceinline__ void init(uint32_t* s, int lane = -1, int known_pos = -1) {
    st = s;
    pos = known_pos >= 0 ? known_pos : (int)s[624];
    base = -1000;
    cache = 0;
    dirty = false;
    if (lane >= 0 && pos < 624) {  // prefetch: the dependent second round trip starts now, not at the first draw
      base = pos;
      const int i = pos + lane;
      cache = mt_temper((i < 624) ? s[i] : 0u);
    }
  }
  __device__ __forceinline__ uint32_t next(int lane) {
    if (pos >= 624) {
      mt_twist_warp(st, lane);
      pos = 0;
      base = -1000;
    }
    if (pos - base >= 32) {
      base = pos;
      const int i = pos + lane;
      cache = mt_temper((i < 624) ? st[i] : 0u);
    }
    const uint32_t v = __shfl_sync(FULL_MASK, cache, pos - base);
    pos++;
    dirty = true;
    return v;
  }
  // RandomState.random_sample(): two draws -> 53-bit double
  __device__ __forceinline__ double next_double(int lane) {
    const uint32_t a = next(lane) >> 5, b = next(lane) >> 6;
    return ((double)a * 67108864.0 + (double)b) * (1.0 / 9007199254740992.0);  // exact: power-of-two scaling
  }
  // RandomState.randint(n): masked rejection sampling, one 32-bit draw per attempt, none if n == 1
  __device__ __forceinline__ int randint(int n, int lane) {
    const uint32_t rng = (uint32_t)(n - 1);
    if (rng == 0) return 0;
    const uint32_t mask = 0xffffffffu >> __clz(rng);  // smallest 2^k - 1 >= rng
    uint32_t v;
    do { v = next(lane) & mask; } while (v > rng);
    return (int)v;
  }
  // x = randint(nx); y = randint(ny) -- the cursor draw of the narrow representations (narrow_rep.py:105-106).
  // Fast path: both 32-bit draws sit in the cached window and both are accepted at the first attempt; anything else
  // (rejection, window refill, twist, n == 1) replays the two generic calls from the unchanged position.
  __device__ __forceinline__ void randint2(int nx, int ny, int lane, int& x, int& y) {
    const uint32_t rx = (uint32_t)(nx - 1), ry = (uint32_t)(ny - 1);
    const int off = pos - base;
    if (rx != 0u && ry != 0u && off >= 0 && off <= 30 && pos <= 622) {
      const uint32_t a = __shfl_sync(FULL_MASK, cache, off) & (0xffffffffu >> __clz(rx));
      const uint32_t b = __shfl_sync(FULL_MASK, cache, off + 1) & (0xffffffffu >> __clz(ry));
      if (a <= rx && b <= ry) {
        x = (int)a; y = (int)b;
        pos += 2;
        dirty = true;
        return;
      }
    }
    x = randint(nx, lane);
    y = randint(ny, lane);
  }
  // `count` (<= 512) consecutive tempered draws into shared memory
  __device__ __forceinline__ void fill(uint32_t* buf, int count, int lane) {
    int filled = 0;
    while (filled < count) {
      if (pos >= 624) {
        mt_twist_warp(st, lane);
        pos = 0;
      }
      const int m = min(count - filled, 624 - pos);
      for (int i0 = 0; i0 < m; i0 += 128) {  // four key words in flight per lane
        uint32_t v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) { const int i = i0 + u * 32 + lane; v[u] = (i < m) ? st[pos + i] : 0u; }
#pragma unroll
        for (int u = 0; u < 4; u++) { const int i = i0 + u * 32 + lane; if (i < m) buf[filled + i] = mt_temper(v[u]); }
      }
      pos += m;
      filled += m;
    }
    base = -1000;
    dirty = true;
    __syncwarp();
  }
  __device__ __forceinline__ void finish(int lane) {
    if (dirty && lane == 0) st[624] = (uint32_t)pos;
  }
  // Stage the 624 key words in shared memory (one coalesced round trip instead of one per twist batch / refill);
  // `pos` stays in the register.  unstage() writes the key back and returns to the HBM copy.
  __device__ __forceinline__ uint32_t* stage(uint32_t* smem_key, int lane) {
    uint32_t* g = st;
    // 312 asynchronous 8-byte copies global -> shared (a stream starts on an 8-byte boundary: 2500 B per stream), all in
    // flight at once and without staging registers; under the 72-register cap the register version (20 loads per lane)
    // was issued in several dependent batches
    if ((reinterpret_cast<uintptr_t>(g) & 7u) == 0u) {
      const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem_key);
#pragma unroll
      for (int k = 0; k < 10; k++) {
        const int i = k * 32 + lane;
        if (i < 312) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sbase + 8u * (uint32_t)i), "l"(g + 2 * i) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    } else {  // a caller-owned rng buffer that is only 4-byte aligned
      for (int i = lane; i < 624; i += 32) smem_key[i] = __ldcg(g + i);
    }
    __syncwarp();
    st = smem_key;
    base = -1000;
    return g;
  }
  __device__ __forceinline__ void unstage(uint32_t* g, int lane) {
    __syncwarp();
    for (int i = lane; i < 624; i += 32) g[i] = st[i];
    st = g;
    base = -1000;
    __syncwarp();
  }
};

// ------------------------------------------------------------------------------------------------
// bitboards
// ------------------------------------------------------------------------------------------------
struct Board {
  uint32_t p0, p1, p2;  // bit-planes of this lane's row
};

__device__ __forceinline__ uint32_t row_mask(int W, int H, int lane) {
  return (lane < H) ? ((W >= 32) ? FULL_MASK : ((1u << W) - 1u)) : 0u;
}

// mask of the cells of this row whose tile index is in TYPES (bit t of TYPES <-> tile t)
template <unsigned TYPES>
__device__ __forceinline__ uint32_t type_mask(const Board& b, uint32_t rmask) {
  uint32_t m = 0;
#pragma unroll
  for (int t = 0; t < 8; t++) {
    if ((TYPES >> t) & 1u) m |= ((t & 1) ? b.p0 : ~b.p0) & ((t & 2) ? b.p1 : ~b.p1) & ((t & 4) ? b.p2 : ~b.p2);
  }
  return m & rmask;
}

__device__ __forceinline__ int tile_at(const Board& b, int x) {
  return (int)(((b.p0 >> x) & 1u) | (((b.p1 >> x) & 1u) << 1) | (((b.p2 >> x) & 1u) << 2));
}

__device__ __forceinline__ void set_tile(Board& b, int x, int t) {
  const uint32_t bit = 1u << x;
  b.p0 = (t & 1) ? (b.p0 | bit) : (b.p0 & ~bit);
  b.p1 = (t & 2) ? (b.p1 | bit) : (b.p1 & ~bit);
  b.p2 = (t & 4) ? (b.p2 | bit) : (b.p2 & ~bit);
}

// One 32-cell chunk of the row-major tile stream -> three ballot words in shared memory.
template <int NPLANES>
__device__ __forceinline__ void chunk_to_bits(uint32_t tile, int chunk, int lane, uint32_t* sbits) {
  const uint32_t b0 = __ballot_sync(FULL_MASK, tile & 1u);
  uint32_t b1 = 0, b2 = 0;
  if (NPLANES > 1) {
    b1 = __ballot_sync(FULL_MASK, tile & 2u);
    b2 = __ballot_sync(FULL_MASK, tile & 4u);
  }
  if (lane == 0) {
    sbits[chunk] = b0;
    sbits[PCGRL_SBITS_STRIDE + chunk] = b1;
    sbits[2 * PCGRL_SBITS_STRIDE + chunk] = b2;
  }
}

// Row r = bits [r*W, r*W+W) of the stream: funnel shift of two adjacent ballot words.
template <int NPLANES>
__device__ __forceinline__ Board bits_to_board(uint32_t* sbits, int nchunks, int W, int H, int lane) {
  if (lane == 0) {  // guard word read by the funnel shift of the last row
    sbits[nchunks] = 0;
    sbits[PCGRL_SBITS_STRIDE + nchunks] = 0;
    sbits[2 * PCGRL_SBITS_STRIDE + nchunks] = 0;
  }
  __syncwarp();
  Board b = {0u, 0u, 0u};
  if (lane < H) {
    const int off = lane * W, wi = off >> 5, sh = off & 31;
    const uint32_t m = (W >= 32) ? FULL_MASK : ((1u << W) - 1u);
    b.p0 = __funnelshift_r(sbits[wi], sbits[wi + 1], sh) & m;
    if (NPLANES > 1) {
      b.p1 = __funnelshift_r(sbits[PCGRL_SBITS_STRIDE + wi], sbits[PCGRL_SBITS_STRIDE + wi + 1], sh) & m;
      b.p2 = __funnelshift_r(sbits[2 * PCGRL_SBITS_STRIDE + wi], sbits[2 * PCGRL_SBITS_STRIDE + wi + 1], sh) & m;
    }
  }
  __syncwarp();
  return b;
}

// Coalesced byte loads of one env's uint8[H][W] map (the packed map batch in HBM) -> bitboards.
template <int NPLANES>
__device__ __forceinline__ Board load_board(const uint8_t* __restrict__ g, int W, int H, int lane, uint32_t* sbits) {
  const int cells = W * H, nchunks = (cells + 31) >> 5;
  for (int c0 = 0; c0 < nchunks; c0 += 8) {  // 8 independent loads in flight per round trip (16x16 = one trip)
    uint32_t t[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const int i = (c0 + k) * 32 + lane;
      t[k] = (i < cells) ? (uint32_t)g[i] : 0u;
    }
#pragma unroll
    for (int k = 0; k < 8; k++)
      if (c0 + k < nchunks) chunk_to_bits<NPLANES>(t[k], c0 + k, lane, sbits);
  }
  return bits_to_board<NPLANES>(sbits, nchunks, W, H, lane);
}

// ------------------------------------------------------------------------------------------------
// graph routines on bitboards
// ------------------------------------------------------------------------------------------------
// 4-neighbourhood dilation (includes f).  Lane 0 / lane 31 receive their own word from the shuffles,
// which is harmless here because f itself is OR-ed in.
__device__ __forceinline__ uint32_t dilate(uint32_t f) {
  return f | (f << 1) | (f >> 1) | __shfl_up_sync(FULL_MASK, f, 1) | __shfl_down_sync(FULL_MASK, f, 1);
}

// neighbours only (exact zero beyond the top / bottom rows)
__device__ __forceinline__ uint32_t neighbours(uint32_t f, int lane) {
  uint32_t up = __shfl_up_sync(FULL_MASK, f, 1), dn = __shfl_down_sync(FULL_MASK, f, 1);
  if (lane == 0) up = 0;
  if (lane == 31) dn = 0;
  return (f << 1) | (f >> 1) | up | dn;
}

__device__ __forceinline__ int popc_all(uint32_t m) { return (int)__reduce_add_sync(FULL_MASK, (unsigned)__popc(m)); }

// first set cell in row-major order (G/helper.py:16-23 location order); returns false if m is empty
__device__ __forceinline__ bool first_cell(uint32_t m, int& row, int& col) {
  const uint32_t rows = __ballot_sync(FULL_MASK, m != 0u);
  if (rows == 0u) return false;
  row = __ffs(rows) - 1;
  col = __shfl_sync(FULL_MASK, __ffs(m) - 1, row);
  return true;
}

__device__ __forceinline__ uint32_t cell_bit(int row, int col, int lane) { return (lane == row) ? (1u << col) : 0u; }

// the same cell as a one-bit board, without materialising (row, col): the first non-empty row keeps its lowest bit
__device__ __forceinline__ bool first_cell_seed(uint32_t m, int lane, uint32_t& seed) {
  const uint32_t rows = __ballot_sync(FULL_MASK, m != 0u);
  seed = (rows != 0u && lane == __ffs(rows) - 1) ? (m & (0u - m)) : 0u;
  return rows != 0u;
}

// BFS from `seed` over `pass` (G/helper.py:222-237 run_dikjstra): returns the eccentricity of the seed inside
// its component, the visited set and the last non-empty frontier (the cells at maximum distance).
__device__ __forceinline__ int bfs_ecc(uint32_t seed, uint32_t pass, uint32_t& visited, uint32_t& last) {
  uint32_t f = seed, vis = seed;
  int d = 0;
  while (true) {  // two waves per termination vote (the tail decides whether the first of the two was the last);
                  // the body is written out twice so that the loop-carried boards need no register moves
    const uint32_t n1 = dilate(f) & pass & ~vis;
    const uint32_t v1 = vis | n1;
    const uint32_t n2 = dilate(n1) & pass & ~v1;
    if (!__any_sync(FULL_MASK, n2 != 0u)) {
      if (__any_sync(FULL_MASK, n1 != 0u)) { vis = v1; f = n1; d += 1; }
      break;
    }
    const uint32_t v2 = v1 | n2;
    const uint32_t n3 = dilate(n2) & pass & ~v2;
    const uint32_t v3 = v2 | n3;
    const uint32_t n4 = dilate(n3) & pass & ~v3;
    if (!__any_sync(FULL_MASK, n4 != 0u)) {
      if (__any_sync(FULL_MASK, n3 != 0u)) { vis = v3; f = n3; d += 3; }
      else { vis = v2; f = n2; d += 2; }
      break;
    }
    vis = v3 | n4;
    f = n4;
    d += 4;
  }
  visited = vis;
  last = f;
  return d;
}

// flood fill without levels (component mask of seed)
__device__ __forceinline__ uint32_t flood(uint32_t seed, uint32_t pass) {
  uint32_t v = seed;
  while (true) {  // monotone: two dilations per vote, a fixed point of the second is a fixed point
    const uint32_t n1 = dilate(v) & pass;
    const uint32_t n2 = dilate(n1) & pass;
    v = n2;
    if (!__any_sync(FULL_MASK, n2 != n1)) break;
  }
  return v;
}

// distance from seed to the target cell over `pass`, -1 if unreachable or the seed is not passable
// (G/helper.py:222-237 semantics, used by zelda_prob.py:104-110)
__device__ __forceinline__ int bfs_dist_to(uint32_t seed, uint32_t target, uint32_t pass) {
  uint32_t f = seed & pass, visited = f;
  int d = 0;
  if (!__any_sync(FULL_MASK, f != 0u)) return -1;
  if (__any_sync(FULL_MASK, (f & target) != 0u)) return 0;
  while (true) {  // two waves per pair of votes: "did either wave reach the target", "is the second wave empty"
    const uint32_t n1 = dilate(f) & pass & ~visited;
    const uint32_t v1 = visited | n1;
    const uint32_t n2 = dilate(n1) & pass & ~v1;
    if (__any_sync(FULL_MASK, ((n1 | n2) & target) != 0u))
      return __any_sync(FULL_MASK, (n1 & target) != 0u) ? d + 1 : d + 2;
    if (!__any_sync(FULL_MASK, n2 != 0u)) return -1;  // n1 empty implies n2 empty; neither touched the target
    visited = v1 | n2;
    f = n2;
    d += 2;
  }
}

// G/helper.py:197-207 calc_num_regions: number of 4-connected components of `pass`.
__device__ __forceinline__ int count_regions(uint32_t pass, int lane) {
  const uint32_t iso = pass & ~neighbours(pass, lane);  // single-cell components, all at once
  int regions = popc_all(iso);
  uint32_t remaining = pass & ~iso;
  uint32_t seed;
  while (first_cell_seed(remaining, lane, seed)) {
    remaining &= ~flood(seed, remaining);
    regions++;
  }
  return regions;
}

// How many components of p0 (a board that does NOT contain the cell (ex, ey)) hold one of the cell's four neighbours,
// decided from the 3x3 window around the cell alone: 0 when no neighbour is set, 1 when all set neighbours are linked
// through set diagonal cells of the window (N-NE-E, E-SE-S, S-SW-W, W-NW-N), -1 when the window cannot tell (the
// neighbours may still be connected through the rest of the map).  Toggling the cell then changes the number of regions
// by 1 - m (cell set) or m - 1 (cell cleared).  Warp-uniform.
__device__ __forceinline__ int local_piece_count(uint32_t p0, int ex, int ey) {
  const uint32_t up_row = __shfl_sync(FULL_MASK, p0, (ey + 31) & 31), mid_row = __shfl_sync(FULL_MASK, p0, ey),
                 dn_row = __shfl_sync(FULL_MASK, p0, (ey + 1) & 31);
  // bits (ex-1, ex, ex+1) of a row -> bits 0..2
  const uint32_t up = ey > 0 ? (uint32_t)((((unsigned long long)up_row) << 1) >> ex) & 7u : 0u;
  const uint32_t mid = (uint32_t)((((unsigned long long)mid_row) << 1) >> ex) & 7u;
  const uint32_t dn = ey < 31 ? (uint32_t)((((unsigned long long)dn_row) << 1) >> ex) & 7u : 0u;
  const uint32_t N = (up >> 1) & 1u, S = (dn >> 1) & 1u, W = mid & 1u, E = (mid >> 2) & 1u;
  const uint32_t NW = up & 1u, NE = (up >> 2) & 1u, SW = dn & 1u, SE = (dn >> 2) & 1u;
  const int k = (int)(N + S + W + E);
  if (k <= 1) return k;
  const int links = (int)((N & NE & E) + (E & SE & S) + (S & SW & W) + (W & NW & N));
  return (k - links <= 1) ? 1 : -1;
}

// G/helper.py:197-207 + :250-264 fused: regions and calc_longest_path of `pass`.
// Per component (processed in row-major order of its first cell, like the reference): BFS from the first
// cell, np.argmax tie-break = row-major-first cell of the last frontier, BFS from there, keep the max.
// The result (a max over components) does not depend on the component order, so three exact shortcuts apply:
// single-cell components contribute 0 and two-cell components contribute 1 (both found for the whole map at
// once from neighbour-count boards); a component whose first sweep has eccentricity d1 has diameter <= 2*d1,
// so its second sweep is skipped when 2*d1 <= best.
// `best_cells` (optional by-product): cells of components whose double-sweep value is known to EQUAL path_out -- a subset
// (components skipped by the 2*d1 <= best shortcut are left out even when they tie), possibly empty.
__device__ __forceinline__ void regions_and_longest_path(uint32_t pass, int lane, int& regions_out, int& path_out,
                                                         uint32_t& best_cells) {
  // neighbour-presence boards: has a passable neighbour to the left / right / above / below
  uint32_t up = __shfl_up_sync(FULL_MASK, pass, 1), dn = __shfl_down_sync(FULL_MASK, pass, 1);
  if (lane == 0) up = 0;
  if (lane == 31) dn = 0;
  const uint32_t l = pass << 1, r = pass >> 1;
  const uint32_t any_nb = l | r | up | dn;
  const uint32_t iso = pass & ~any_nb;                       // single-cell components: path 0
  // cells with exactly one passable neighbour; two adjacent such cells form a 2-cell component: path 1
  const uint32_t one = pass & (l ^ r ^ up ^ dn) & ~((l & r) | (up & dn) | ((l ^ r) & (up ^ dn)));
  const uint32_t hd = one & (one >> 1);                      // left cell of a horizontal domino
  uint32_t below = __shfl_down_sync(FULL_MASK, one, 1);
  if (lane == 31) below = 0;
  const uint32_t vd = one & below;                           // top cell of a vertical domino
  uint32_t vd_low = __shfl_up_sync(FULL_MASK, vd, 1);
  if (lane == 0) vd_low = 0;
  const uint32_t dominoes = hd | (hd << 1) | vd | vd_low;
  const int ndom = popc_all(hd) + popc_all(vd);
  int regions = popc_all(iso) + ndom, best = ndom > 0 ? 1 : 0;
  uint32_t bm = ndom > 0 ? dominoes : iso;
  uint32_t remaining = pass & ~iso & ~dominoes;
  uint32_t seed;
  while (first_cell_seed(remaining, lane, seed)) {
    uint32_t visited, last;
    const int d1 = bfs_ecc(seed, remaining, visited, last);
    remaining &= ~visited;
    regions++;
    if (2 * d1 > best) {
      first_cell_seed(last, lane, seed);
      uint32_t v2, l2;
      const int d2 = bfs_ecc(seed, visited, v2, l2);
      if (d2 > best) { best = d2; bm = visited; }
      else if (d2 == best) bm |= visited;
    }
  }
  regions_out = regions;
  path_out = best;
  best_cells = bm;
}
__device__ __forceinline__ void regions_and_longest_path(uint32_t pass, int lane, int& regions_out, int& path_out) {
  uint32_t unused;
  regions_and_longest_path(pass, lane, regions_out, path_out, unused);
}

// Out-of-line copies of the two sweeps (results in registers): binary_stats_update calls them from rolled loops so that
// the hot loop of k_rollout stays small (instruction-cache misses were 24 % of its stall samples when every call site
// carried its own unrolled copy).
__device__ __noinline__ uint3 bfs_ecc_call(uint32_t seed, uint32_t pass) {
  uint32_t visited, last;
  const int d = bfs_ecc(seed, pass, visited, last);
  return make_uint3((uint32_t)d, visited, last);
}
__device__ __noinline__ uint32_t flood_call(uint32_t seed, uint32_t pass) { return flood(seed, pass); }
__device__ __noinline__ uint3 regions_and_longest_path_call(uint32_t pass, int lane) {  // (regions, path, best_cells)
  int regions, path;
  uint32_t bm;
  regions_and_longest_path(pass, lane, regions, path, bm);
  return make_uint3((uint32_t)regions, (uint32_t)path, bm);
}

// Incremental form of the two binary statistics after ONE cell `cb` (a one-bit board) of the passable board changed
// (`pass` is the board AFTER the edit; grew = the cell became passable).  Exact, because both statistics are functions
// of the components alone: regions is their number and calc_longest_path (G/helper.py:250-264) is a maximum over
// components of a value -- sweep from the component's row-major-first cell, sweep from the row-major-first farthest
// cell -- that depends on nothing but the component's own cells.  An edit touches only the components next to the
// cell: with P0 = the board without the cell, the `pieces` are the components of P0 that hold one of the cell's (at
// most four) passable neighbours; a grow event replaces the pieces by their union plus the cell, a shrink event
// replaces that union by the pieces.  Carried between calls: regions, best and `bm`, a (possibly empty) set of cells of
// components whose value is known to equal best.  The maximum over the untouched components is known to be `best`
// when an untouched component is in bm or when no touched old component is large enough to reach best (a component of
// s cells has value <= s - 1); otherwise it is only known to be <= best, which still decides the new maximum whenever
// a new component reaches best.  In the remaining case the whole board is recomputed.
__device__ __forceinline__ void binary_stats_update(uint32_t pass, uint32_t cb, bool grew, int lane, int& regions, int& best,
                                                    uint32_t& bm, uint32_t* piece_smem) {
  // piece_smem: 4 x 32 words of per-warp shared memory (piece k, row = lane)
  const uint32_t p0 = pass & ~cb;
  uint32_t nb = neighbours(cb, lane) & p0;
  uint32_t touched = 0u, seed;
  int m = 0, max_piece = 0;
#pragma unroll 1
  while (first_cell_seed(nb, lane, seed)) {  // at most four rounds, warp-uniform
    const uint32_t piece = flood_call(seed, p0);
    piece_smem[m * 32 + lane] = piece;
    nb &= ~piece;
    touched |= piece;
    max_piece = max(max_piece, popc_all(piece));
    m++;
  }
  regions += grew ? 1 - m : m - 1;
  const int touched_cells = grew ? max_piece : popc_all(touched) + 1;  // the largest touched OLD component
  const uint32_t old_cells = grew ? touched : (touched | cb);
  const uint32_t keep = bm & ~old_cells;
  const bool rest_is_best = __any_sync(FULL_MASK, keep != 0u) || touched_cells - 1 < best;
  int cur = rest_is_best ? best : best - 1;  // a new component matters only if its value exceeds cur
  uint32_t cur_cells = 0u;
  if (grew) { piece_smem[lane] = touched | cb; m = 1; }
#pragma unroll 1
  for (int k = 0; k < m; k++) {
    const uint32_t comp = piece_smem[k * 32 + lane];
    if (popc_all(comp) - 1 > cur) {
      first_cell_seed(comp, lane, seed);
      const uint3 s1 = bfs_ecc_call(seed, comp);  // (d1, visited, last)
      if (2 * (int)s1.x > cur) {
        first_cell_seed(s1.z, lane, seed);
        const int d2 = (int)bfs_ecc_call(seed, comp).x;
        if (d2 > cur) { cur = d2; cur_cells = comp; }
        else if (d2 == cur) cur_cells |= comp;
      }
    }
  }
  if (cur >= best) {
    bm = ((cur == best) ? keep : 0u) | cur_cells;
    best = cur;
  } else {  // the maximum sat in a touched component and no new component reaches it
    const uint3 full = regions_and_longest_path_call(pass, lane);
    best = (int)full.y;
    bm = full.z;
  }
}

// G/helper.py:37-62 get_floor_dist(map, from, floor): sum over `from` cells of the number of cells strictly
// between the cell and the first floor cell below it (H-1 if there is none).
__device__ __forceinline__ int floor_dist(uint32_t from, uint32_t floor_m, int H, int lane) {
  int total = 0, row, col;
  while (first_cell(from, row, col)) {
    from &= ~cell_bit(row, col, lane);
    const uint32_t column = __ballot_sync(FULL_MASK, (floor_m >> col) & 1u);
    const uint32_t below = (row >= 31) ? 0u : (column >> (row + 1));
    total += below ? (__ffs(below) - 1) : (H - 1);
  }
  return total;
}

// G/helper.py:366-376 get_range_reward in fp64 (bounds may be +-inf)
__host__ __device__ __forceinline__ double range_reward(double nv, double ov, double low, double high) {
  if (nv >= low && nv <= high && ov >= low && ov <= high) return 0.0;
  if (ov <= high && nv <= high) return fmin(nv, low) - fmin(ov, low);
  if (ov >= low && nv >= low) return fmax(ov, high) - fmax(nv, high);
  if (nv > high && ov < low) return high - nv + ov - low;
  if (nv < low && ov > high) return high - ov + nv - low;
  return 0.0;
}

// get_range_reward on integer statistics with finite integer bounds: every branch of the fp64 form above yields a small
// integer exactly, so the int32 arithmetic is value-identical (the caller converts to double before the weighted sum).
__host__ __device__ __forceinline__ int range_reward_i(int nv, int ov, int low, int high) {
  if (nv >= low && nv <= high && ov >= low && ov <= high) return 0;
  if (ov <= high && nv <= high) return (nv < low ? nv : low) - (ov < low ? ov : low);
  if (ov >= low && nv >= low) return (ov > high ? ov : high) - (nv > high ? nv : high);
  if (nv > high && ov < low) return high - nv + ov - low;
  if (nv < low && ov > high) return high - ov + nv - low;
  return 0;
}
// the same with high = +inf: only the first two branches can be taken
__host__ __device__ __forceinline__ int range_reward_i_hi_inf(int nv, int ov, int low) {
  if (nv >= low && ov >= low) return 0;
  return (nv < low ? nv : low) - (ov < low ? ov : low);
}

}  // namespace pcgrl

// pcgrl_wrappers.cuh -- batched observation / action wrappers (reference: gym_pcgrl/wrappers.py).
//
//   k_obs_image   Cropped (:163-206) -> OneHotEncoding (:67-104) -> ToImage (:18-60) fused: one pass over the uint8
//                 map batch produces the [N, S, S, C] policy input (S = crop size or the map size; C = 1 raw tile
//                 index, or num_tiles one-hot channels).  HBM-write bound: one thread per output pixel writes its
//                 C channels contiguously, a warp covers 32 consecutive pixels.
//   k_action_map  ActionMap (:111-154): flat index over (h, w, num_tiles) -> the wrapped env's action.
#pragma once
#include "pcgrl_device.cuh"

namespace pcgrl {

// out[n][i][j][c]; crop: padded = np.pad(map, pad, constant_values=pad_value); cropped = padded[y:y+S, x:x+S]
//
// HBM-write bound, so the kernel is organised around the output stream: every thread produces 64 contiguous
// output bytes (decoding its start (env, i, j, channel) once and then stepping channel -> column -> row -> env
// incrementally), parks them in shared memory, and the CTA then streams its 16 KB tile out with fully coalesced
// 16-byte stores (a warp store covers 512 contiguous bytes).  The uint8 map batch is read through L1/L2.
#define OBS_THREADS 256
template <typename OutT>
__global__ void __launch_bounds__(OBS_THREADS) k_obs_image(const uint8_t* __restrict__ maps,
                                                           const uint8_t* __restrict__ pos, OutT* __restrict__ out,
                                                           uint32_t total, int n, int H, int W, int S_h, int S_w,
                                                           int crop, int pad_value, int C) {
  constexpr int EPT = 64 / (int)sizeof(OutT);  // elements per thread
  __shared__ uint4 tile_s[OBS_THREADS * 4];
  const int tid = threadIdx.x;
  const uint32_t tile_elems = (uint32_t)(OBS_THREADS * EPT);
  const uint32_t ntiles = (total + tile_elems - 1) / tile_elems;
  for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {  // one trip with the default grid (a persistent
  // grid of 8 CTAs/SM was measured slower: 0.52 vs 0.61 of the HBM peak)
  const uint32_t block_elem0 = tile * tile_elems;
  const uint32_t k0 = block_elem0 + (uint32_t)tid * EPT;
  constexpr int EPC = 16 / (int)sizeof(OutT);  // elements per 16-byte chunk
  if (k0 < total) {
    const uint32_t per_env = (uint32_t)S_h * S_w;
    uint32_t p = k0 / (uint32_t)C;
    int c = (int)(k0 - p * (uint32_t)C);
    int e = (int)(p / per_env);
    const uint32_t rem = p - (uint32_t)e * per_env;
    int i = (int)(rem / (uint32_t)S_w), j = (int)(rem - (uint32_t)i * S_w);
    const int pad = crop / 2;
    int oy = 0, ox = 0;
    if (crop) { ox = (int)pos[2 * e] - pad; oy = (int)pos[2 * e + 1] - pad; }
    const uint8_t* m = maps + (size_t)e * H * W;
    auto fetch = [&]() -> int {
      const int my = oy + i, mx = ox + j;
      return (my >= 0 && my < H && mx >= 0 && mx < W) ? (int)m[my * W + mx] : pad_value;
    };
    int t = fetch();
#pragma unroll 1
    for (int r = 0; r < 4; r++) {  // four 16-byte chunks per thread, each built in registers
      union { OutT v[EPC]; uint4 q; } u;
#pragma unroll
      for (int q = 0; q < EPC; q++) {
        u.v[q] = (C == 1) ? (OutT)t : (OutT)(c == t ? 1 : 0);  // np.eye(dim)[map]
        if (++c == C) {
          c = 0;
          if (++j == S_w) {
            j = 0;
            if (++i == S_h) {
              i = 0;
              if (++e < n) {
                m += (size_t)H * W;
                if (crop) { ox = (int)pos[2 * e] - pad; oy = (int)pos[2 * e + 1] - pad; }
              }
            }
          }
          if (e < n) t = fetch();
        }
      }
      // park (swizzled inside each thread's group of four to spread the banks)
      tile_s[tid * 4 + (r ^ (tid & 3))] = u.q;
    }
  }
  __syncthreads();
  const size_t total_bytes = (size_t)total * sizeof(OutT);
  const size_t block_byte0 = (size_t)block_elem0 * sizeof(OutT);
  uint8_t* ob = reinterpret_cast<uint8_t*>(out);
#pragma unroll
  for (int r = 0; r < 4; r++) {
    const int L = r * OBS_THREADS + tid;  // linear 16-byte chunk inside the CTA tile
    const size_t off = block_byte0 + (size_t)L * 16;
    const uint4 val = tile_s[(L & ~3) | ((L & 3) ^ ((L >> 2) & 3))];
    if (off + 16 <= total_bytes) {
      *reinterpret_cast<uint4*>(ob + off) = val;
    } else if (off < total_bytes) {  // ragged tail of the whole tensor
      const uint8_t* vb = reinterpret_cast<const uint8_t*>(&val);
      for (int b = 0; off + b < total_bytes; b++) ob[off + b] = vb[b];
    }
  }
  __syncthreads();  // the tile buffer is reused by the next trip
  }
}

// uint8 fast path of the same operator (the PPO consumer's observation buffer), one WARP per env: the env's map is staged in
// shared memory once (aligned 4-byte loads), then every lane builds whole 32-bit output words:
//   MODE 0  raw tile index (C == 1, S_w % 4 == 0): an output row is a byte-shifted window of a map row -- two aligned
//           shared-memory words funnel-shifted into place, the bytes outside [0, W) replaced by the border tile with one mask;
//   MODE 1  one-hot with 8 channels: a word is four channels of one pixel, i.e. (1 << 8 (tile - c0)) or 0.
// ~5 (MODE 0) / ~3 (MODE 1) instructions per output byte instead of ~12, no dependent global loads, and warp stores
// of consecutive words.  `magic` = ceil(65536 / D) turns the division by D (words per row / pixels per row) into a
// multiply-shift (the host checks that it is exact over the whole index range).
#define OBS_FAST_WARPS 8
#define OBS_FAST_SLACK 32
template <int MODE>
__global__ void __launch_bounds__(32 * OBS_FAST_WARPS) k_obs_image_u8_fast(const uint8_t* __restrict__ maps, const uint8_t* __restrict__ pos,
                                                                        uint32_t* __restrict__ out, int n, int H, int W, int S_h,
                                                                        int S_w, int crop, int pad_value, uint32_t magic,
                                                                        size_t map_bytes_total) {
  __shared__ __align__(16) uint8_t smap_all[OBS_FAST_WARPS][OBS_FAST_SLACK + 1024 + OBS_FAST_SLACK + 8];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  uint8_t* smap = smap_all[wib];
  uint32_t* smap32 = reinterpret_cast<uint32_t*>(smap);
  const int cells = H * W, pad = crop / 2;
  const uint32_t padword = 0x01010101u * (uint32_t)pad_value;
  const int words_per_env = (MODE == 0) ? S_h * (S_w >> 2) : S_h * S_w * 2;
  const uint32_t* maps32 = reinterpret_cast<const uint32_t*>(maps);
  const size_t last_word = (map_bytes_total + 3) / 4;
  for (int e = blockIdx.x * OBS_FAST_WARPS + wib; e < n; e += gridDim.x * OBS_FAST_WARPS) {
    // stage: the aligned words covering [e * cells, (e + 1) * cells); map byte k sits at smap[SLACK + off + k]
    const size_t b0 = (size_t)e * cells;
    const int off = (int)(b0 & 3);
    const size_t g0 = b0 >> 2;
    const int nw = (off + cells + 3) >> 2;
    for (int k = lane; k < nw; k += 32) {
      uint32_t v = 0u;
      const size_t g = g0 + k;
      if ((g + 1) * 4 <= map_bytes_total) v = maps32[g];
      else if (g < last_word)  // the last, partial word of the whole batch: never read past the caller's buffer
        for (size_t b = g * 4; b < map_bytes_total; b++) v |= (uint32_t)maps[b] << (8 * (int)(b - g * 4));
      smap32[(OBS_FAST_SLACK >> 2) + k] = v;
    }
    int ox = 0, oy = 0;
    if (crop) { ox = (int)pos[2 * e] - pad; oy = (int)pos[2 * e + 1] - pad; }
    __syncwarp();
    uint32_t* o = out + (size_t)e * words_per_env;
    const int base = OBS_FAST_SLACK + off;
    for (int w = lane; w < words_per_env; w += 32) {
      uint32_t word;
      if (MODE == 0) {
        const int wpr = S_w >> 2;
        const int i = (int)(((uint32_t)w * magic) >> 16), q = w - i * wpr;
        const int my = oy + i, mx0 = ox + 4 * q;
        word = padword;
        if ((unsigned)my < (unsigned)H) {
          int lo = -mx0, hi = W - mx0;           // valid bytes of this word: [lo, hi) clamped to [0, 4]
          lo = lo < 0 ? 0 : (lo > 4 ? 4 : lo);
          hi = hi < 0 ? 0 : (hi > 4 ? 4 : hi);
          if (hi > lo) {
            const int a = base + my * W + mx0;   // >= 0: the front slack covers mx0 >= -32
            const uint32_t w0 = smap32[a >> 2], w1 = smap32[(a >> 2) + 1];
            const uint32_t src = __funnelshift_r(w0, w1, (a & 3) * 8);
            const uint32_t m = (uint32_t)(((1ull << (8 * hi)) - 1ull) & ~((1ull << (8 * lo)) - 1ull));
            word = (src & m) | (padword & ~m);
          }
        }
      } else {
        const int p = w >> 1, c0 = (w & 1) * 4;
        const int i = (int)(((uint32_t)p * magic) >> 16), j = p - i * S_w;
        const int my = oy + i, mx = ox + j;
        const int t = ((unsigned)my < (unsigned)H && (unsigned)mx < (unsigned)W) ? (int)smap[base + my * W + mx] : pad_value;
        const unsigned d = (unsigned)(t - c0);
        word = d < 4u ? (1u << (8 * d)) : 0u;   // np.eye(8)[tile]
      }
      o[w] = word;
    }
    __syncwarp();  // the staging buffer is reused for the next env
  }
}

// ActionMap.step (:139-154): (y, x, v) = unravel_index(action, (h, w, dim)).
//   wide representations       -> [x, y, v]
//   cursor representations     -> v if (x, y) is the cursor, else the tile value under the cursor (sic: the
//                                 reference passes the tile VALUE o_v as the action)
__global__ void k_action_map(const int32_t* __restrict__ flat, const uint8_t* __restrict__ maps,
                             const uint8_t* __restrict__ pos, int32_t* __restrict__ actions, int n, int H, int W,
                             int dim, int wide) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const int a = flat[e];
  const int v = a % dim, x = (a / dim) % W, y = a / (dim * W);
  if (wide) {
    actions[3 * e] = x; actions[3 * e + 1] = y; actions[3 * e + 2] = v;
  } else {
    const int ox = pos[2 * e], oy = pos[2 * e + 1];
    actions[e] = (ox == x && oy == y) ? v : (int)maps[((size_t)e * H + oy) * W + ox];
  }
}

// PcgrlEnv.render (pcgrl_env.py:160-173) = Problem.render (probs/problem.py:134-156: a border of `border_tile`, every
// map tile pasted as a tile_size x tile_size sprite) + the cursor frame of the narrow / turtle representations
// (reps/narrow_rep.py:126-140: a 2-pixel red frame on the cursor tile) + convert("RGB"), for the whole batch:
//   maps [n][H][W] u8, atlas [num_tiles][ts][ts][4] RGBA u8  ->  out [n][(H + 2 bh) ts][(W + 2 bw) ts][3] RGB u8.
// HBM-write bound: one thread produces four horizontal pixels (12 bytes, three aligned 32-bit stores; ts % 4 == 0).
__global__ void __launch_bounds__(256) k_render(const uint8_t* __restrict__ maps, const uint8_t* __restrict__ pos,
                                                const uint32_t* __restrict__ atlas, uint32_t* __restrict__ out, int n, int H,
                                                int W, int bw, int bh, int border_tile, int ts) {
  const int wpx = (W + 2 * bw) * ts, hpx = (H + 2 * bh) * ts, quads = wpx >> 2;
  const size_t total = (size_t)n * hpx * quads;
  for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (size_t)gridDim.x * blockDim.x) {
    const int qx = (int)(q % quads);
    const size_t rest = q / quads;
    const int py = (int)(rest % hpx), e = (int)(rest / hpx);
    const int px = qx << 2, tx = px / ts - bw, ty = py / ts - bh, sx = px % ts, sy = py % ts;
    const bool inside = tx >= 0 && tx < W && ty >= 0 && ty < H;
    const int tile = inside ? (int)maps[((size_t)e * H + ty) * W + tx] : border_tile;
    const uint4 rgba = *reinterpret_cast<const uint4*>(atlas + ((size_t)tile * ts + sy) * ts + sx);  // four RGBA pixels
    uint32_t p[4] = {rgba.x, rgba.y, rgba.z, rgba.w};
    if (pos && inside && tx == (int)pos[2 * e] && ty == (int)pos[2 * e + 1]) {
      const bool edge_row = sy < 2 || sy >= ts - 2;
#pragma unroll
      for (int k = 0; k < 4; k++)
        if (edge_row || sx + k < 2 || sx + k >= ts - 2) p[k] = 0xff0000ffu;  // (255, 0, 0, 255) little-endian RGBA
    }
    // RGBA x 4 -> RGB x 4 = 12 bytes
    const uint32_t w0 = (p[0] & 0xffffffu) | ((p[1] & 0xffu) << 24);
    const uint32_t w1 = ((p[1] >> 8) & 0xffffu) | ((p[2] & 0xffffu) << 16);
    const uint32_t w2 = ((p[2] >> 16) & 0xffu) | ((p[3] & 0xffffffu) << 8);
    uint32_t* o = out + q * 3;
    o[0] = w0; o[1] = w1; o[2] = w2;
  }
}

}  // namespace pcgrl

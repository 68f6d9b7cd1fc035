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

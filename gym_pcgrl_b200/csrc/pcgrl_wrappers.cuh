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
template <typename OutT>
__global__ void __launch_bounds__(256) k_obs_image(const uint8_t* __restrict__ maps, const uint8_t* __restrict__ pos,
                                                   OutT* __restrict__ out, int n, int H, int W, int S_h, int S_w,
                                                   int crop, int pad_value, int channels) {
  const size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t per_env = (size_t)S_h * S_w;
  if (pix >= (size_t)n * per_env) return;
  const int e = (int)(pix / per_env), rem = (int)(pix % per_env), i = rem / S_w, j = rem % S_w;
  int t;
  if (crop) {
    const int pad = crop / 2;
    const int my = (int)pos[2 * e + 1] + i - pad, mx = (int)pos[2 * e] + j - pad;
    t = (my >= 0 && my < H && mx >= 0 && mx < W) ? (int)maps[((size_t)e * H + my) * W + mx] : pad_value;
  } else {
    t = (int)maps[((size_t)e * H + i) * W + j];
  }
  OutT* o = out + pix * channels;
  if (channels == 1) {
    o[0] = (OutT)t;
  } else {
    for (int c = 0; c < channels; c++) o[c] = (OutT)(c == t ? 1 : 0);  // np.eye(dim)[map]
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

}  // namespace pcgrl

// pcgrl_linear.cu -- the one dense contraction next to the hot path: the 512-unit fully connected
// layer of the reference's Cnn1 / Cnn2 feature extractors (model.py:15,23: fc1 over the flattened conv features),
//   Y[M, N] = relu(X[M, K] . W[N, K]^T + b),  M = envs, K = 4*4*64 = 1024, N = 512 for the default policy,
// as a hand-written sm_100a kernel: TMA (cp.async.bulk.tensor) stages 128 x 64 bf16 tiles of X and W in shared memory with
// the 128-byte swizzle, ONE elected thread issues tcgen05.mma (cta_group::1, kind::f16, 128 x 128 x 16) accumulating in
// TMEM, four epilogue warps read the accumulator back with tcgen05.ld, add the bias, apply the ReLU and store fp32.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2-5 = epilogue (a warp may
// only touch TMEM lanes [32 * (warp_id % 4), +32), so warps 2, 3, 4, 5 cover lane quarters 2, 3, 0, 1).
// Pipelines: full[s] / empty[s] mbarriers between TMA and MMA over 4 smem stages; tmem_full[a] / tmem_empty[a] between MMA
// and epilogue over 2 accumulator stages.
// Persistent CTAs (one per SM) loop over 128 x 128 output tiles; the accumulator is double-buffered in TMEM so the epilogue of
// one tile overlaps the main loop of the next (128 tiles = 128 CTAs for the default policy at 4096 envs).
//
// Descriptor encodings follow the public CUTLASS definitions (cute/arch/mma_sm100_desc.hpp: UMMA::SmemDescriptor,
// UMMA::InstrDescriptor); every wait is bounded and traps instead of hanging the GPU.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace pcgrl_linear {

constexpr int BLOCK_M = 128, BLOCK_K = 64, UMMA_K = 16, STAGES = 4, THREADS = 192;
constexpr int TILE_A_BYTES = BLOCK_M * BLOCK_K * 2;
// BLOCK_N is a template parameter: 128 (more tiles: small problems fill the 148 SMs) or 256 (one third less operand
// traffic per flop: large problems).  TMEM holds two fp32 accumulator stages of 128 lanes x BLOCK_N columns.
template <int BN> __host__ __device__ constexpr int tile_b_bytes() { return BN * BLOCK_K * 2; }
template <int BN> __host__ __device__ constexpr int smem_bytes() { return STAGES * (TILE_A_BYTES + tile_b_bytes<BN>()) + 1024 /* alignment slack */ + 2048 /* barriers, TMEM slot, bias */; }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  for (uint32_t spin = 0; spin < (1u << 26); spin++) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) return;
  }
  __trap();  // a protocol bug must surface as a launch error, never as a hung GPU
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// UMMA shared-memory matrix descriptor, K-major, SWIZZLE_128B: 8-row x 128-byte atoms stacked along M/N every 1024 bytes
__device__ __forceinline__ uint64_t umma_smem_desc(const void* tile) {
  const uint64_t addr = (uint64_t)((smem_u32(tile) >> 4) & 0x3FFFu);
  const uint64_t sbo = (uint64_t)((8 * BLOCK_K * 2) >> 4);  // stride byte offset: 1024 B between 8-row groups
  return addr | (0ull << 16) /* LBO unused: one swizzle atom along K */ | (sbo << 32) | (1ull << 46) /* version: sm_100 */ |
         (2ull << 61) /* LayoutType::SWIZZLE_128B */;
}
// A tile may also be read starting `r` (< 8) rows into a 1024-byte swizzle atom by adding r * 128 bytes (r * 8 in this
// descriptor's address field) to the start address and nothing else: the tensor core applies the 128-byte swizzle to the
// ABSOLUTE shared-memory address it computes (start + (i / 8) * SBO + (i % 8) * 128 is linear in the row i because
// SBO = 8 * 128), exactly like TMA did when it wrote the tile.  Measured: with the "matrix base offset" bits (49-51) set to
// the row phase the results are wrong, with 0 they equal those of separately loaded tiles.  k_conv3x3_bf16 feeds the three
// horizontal filter taps from ONE staged tile this way.
// UMMA instruction descriptor (kind::f16): D = F32, A = B = BF16, both K-major, N >> 3 at [17,23), M >> 4 at [24,29)
template <int BN> __host__ __device__ constexpr uint32_t instr_desc() {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
}

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {  // arrives on `bar` once all MMAs issued so far have completed
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Warp-convergent forms (all 32 lanes execute them, elect.sync picks the issuing lane inside the PTX block): in a branch
// on `lane == 0` the compiler wraps every UTCHMMA in an elect / branch loop of its own.
__device__ __forceinline__ void umma_f16_elect(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred e, p;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_elect(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar))
      : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Persistent: CTA b works on output tiles b, b + gridDim.x, ... (tile = m_block * tiles_n + n_block, so the CTAs that share
// an X tile run at the same time and it is read from HBM once).  The accumulator is double-buffered in TMEM (2 x 128
// columns): the epilogue of tile i overlaps the TMA / MMA main loop of tile i + 1.
template <int BLOCK_N>
__global__ void __launch_bounds__(THREADS, 1) k_linear_bf16(const __grid_constant__ CUtensorMap map_x,
                                                            const __grid_constant__ CUtensorMap map_w,
                                                            const float* __restrict__ bias, void* __restrict__ y_out, int M, int N,
                                                            int K, int relu, int out_bf16, int pad_h, int pad_w) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);  // SW128: 1024-B aligned
  uint8_t* smem_a = smem;
  constexpr int TILE_B_BYTES = tile_b_bytes<BLOCK_N>();
  constexpr uint32_t TMEM_COLS = 2 * BLOCK_N;
  uint8_t* smem_b = smem + STAGES * TILE_A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * (TILE_A_BYTES + TILE_B_BYTES));
  uint64_t* full = bars;                       // [STAGES] TMA -> MMA
  uint64_t* empty = bars + STAGES;             // [STAGES] MMA -> TMA
  uint64_t* tmem_full = bars + 2 * STAGES;     // [2] MMA -> epilogue
  uint64_t* tmem_empty = bars + 2 * STAGES + 2;  // [2] epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  float* bias_s = reinterpret_cast<float*>(bars + 2 * STAGES + 6);  // [BLOCK_N] bias of the current n-block (epilogue warps)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_n = (N + BLOCK_N - 1) / BLOCK_N, tiles_m = (M + BLOCK_M - 1) / BLOCK_M, num_tiles = tiles_m * tiles_n;
  const int num_kb = (K + BLOCK_K - 1) / BLOCK_K;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_x)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_w)) : "memory");
    for (int s = 0; s < STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; a++) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // TMEM allocation is warp-collective; the same warp frees it at the end
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // (broadcast: the compiler then keeps the TMEM address in a uniform register instead of re-electing a lane per MMA)
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer: one ring of smem stages across all tiles of this CTA =====
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / tiles_n) * BLOCK_M, n0 = (tile % tiles_n) * BLOCK_N;
        for (int kb = 0; kb < num_kb; kb++, it++) {
          const int s = it % STAGES;
          mbar_wait(&empty[s], ((it / STAGES) & 1) ^ 1);  // passes at once during the first trip round the ring
          mbar_expect_tx(&full[s], TILE_A_BYTES + TILE_B_BYTES);
          tma_load_2d(smem_a + s * TILE_A_BYTES, &map_x, &full[s], kb * BLOCK_K, m0);
          tma_load_2d(smem_b + s * TILE_B_BYTES, &map_w, &full[s], kb * BLOCK_K, n0);
        }
      }
    }
  } else if (warp == 1) {
    {  // ===== MMA issuer: the warp runs the loop convergently, elect.sync picks the lane that drives the tensor core =====
      int it = 0, lt = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, lt++) {
        const int as = lt & 1;
        mbar_wait(&tmem_empty[as], ((lt >> 1) & 1) ^ 1);  // the epilogue has drained this accumulator stage
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * BLOCK_N);
        for (int kb = 0; kb < num_kb; kb++, it++) {
          const int s = it % STAGES;
          mbar_wait(&full[s], (it / STAGES) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t da = umma_smem_desc(smem_a + s * TILE_A_BYTES), db = umma_smem_desc(smem_b + s * TILE_B_BYTES);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; k++)  // +32 bytes along K inside the swizzle atom = +2 in the (addr >> 4) field
            umma_f16_elect(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), instr_desc<BLOCK_N>(), (kb | k) ? 1u : 0u);
          umma_commit_elect(&empty[s]);  // frees the smem stage once these MMAs have read it
        }
        umma_commit_elect(&tmem_full[as]);  // accumulator of this tile complete
      }
    }
  } else {  // ===== epilogue (warps 2-5): TMEM -> registers -> bias + ReLU -> global =====
    const int quarter = warp & 3;  // TMEM lanes [32 * quarter, +32)
    const int et = threadIdx.x - 64;  // 0..127
    int lt = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, lt++) {
      const int m0 = (tile / tiles_n) * BLOCK_M, n0 = (tile % tiles_n) * BLOCK_N, as = lt & 1;
      // stage this n-block's bias while the main loop of the tile is still running
      asm volatile("bar.sync 1, 128;" ::: "memory");  // previous tile's readers of bias_s are done
      for (int c = et; c < BLOCK_N; c += 128) bias_s[c] = (bias && n0 + c < N) ? bias[n0 + c] : 0.0f;
      asm volatile("bar.sync 1, 128;" ::: "memory");
      mbar_wait(&tmem_full[as], (lt >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int row = m0 + quarter * 32 + lane;  // accumulator row m <-> TMEM lane m (M = 128, cta_group::1)
#pragma unroll 1
      for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
        uint32_t v[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * BLOCK_N + c0);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
              "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
              "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
              "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr)
            : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (c0 + 32 >= BLOCK_N) {  // last read of this accumulator stage: hand it back to the MMA warp
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          mbar_arrive(&tmem_empty[as]);
        }
        if (row < M) {
          // pad_w > 0: row m = (e, y, x) of an [n, pad_h, pad_w] image batch lands at (e, y + 1, x + 1) of the zero-bordered
          // [n, pad_h + 2, pad_w + 2] buffer the implicit-GEMM convolution reads (k_conv3x3_bf16)
          size_t orow = (size_t)row;
          if (pad_w > 0) {
            const int x = row % pad_w, t = row / pad_w, yy = t % pad_h, e = t / pad_h;
            orow = ((size_t)e * (pad_h + 2) + yy + 1) * (pad_w + 2) + x + 1;
          }
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; j++) {
            const float t = __uint_as_float(v[j]) + bias_s[c0 + j];
            f[j] = relu ? fmaxf(t, 0.0f) : t;
          }
          if (out_bf16) {  // bf16 activations for the next layer (N % 8 == 0): 16-byte stores of eight values
            __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(y_out) + orow * N + n0 + c0;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              if (n0 + c0 + j + 7 < N) {
                uint4 o;
                uint32_t* ow = &o.x;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                  const __nv_bfloat162 p2 = __floats2bfloat162_rn(f[j + 2 * q], f[j + 2 * q + 1]);
                  ow[q] = *reinterpret_cast<const uint32_t*>(&p2);
                }
                *reinterpret_cast<uint4*>(out + j) = o;
              } else {
                for (int q = 0; q < 8; q++) if (n0 + c0 + j + q < N) out[j + q] = __float2bfloat16_rn(f[j + q]);
              }
            }
          } else {
            float* out = reinterpret_cast<float*>(y_out) + orow * N + n0 + c0;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              if (n0 + c0 + j + 3 < N) *reinterpret_cast<float4*>(out + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
              else for (int q = 0; q < 4; q++) if (n0 + c0 + j + q < N) out[j + q] = f[j + q];
            }
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
}

// ------------------------------------------------------------------------------------------------------------------------
// 3 x 3, stride 1, SAME convolution + bias + ReLU as an IMPLICIT GEMM (the seven 64-channel layers and the n_tools head of the
// reference's fully convolutional policies, model.py:25-77).  Activations live in a zero-bordered NHWC buffer
// P[n][H + 2][W + 2][C] (bf16, C % 64 == 0).  Over the flattened padded positions p the input of filter tap (r, s) for 128
// consecutive outputs is the SAME 2-D tensor [n (H+2) (W+2), C] shifted by (r - 1)(W + 2) + (s - 1) rows, so every A tile is a
// plain TMA box load at a row offset (out-of-range rows are zero-filled by TMA) and no patch matrix is ever written; the
// three taps of one filter row differ by ONE row, so a single staged tile of 130 rows serves all three through UMMA
// descriptors that start 0, 1 and 2 rows into the swizzle atom (see umma_smem_desc) -- each activation tile is read from L2
// three times, not nine:
//     out[p][:] = relu(bias + sum_{tap} P[p + shift(tap)][:] . Wt[tap][:, :]^T)      for interior p; border p are stored as 0
// so the output is again a zero-bordered buffer for the next layer.  The whole weight matrix (9 C/64 blocks of BN x 64) is
// loaded once per CTA and stays in shared memory; the ring of stages carries A tiles only.  Same warp roles, pipelines and
// TMEM double buffering as k_linear_bf16.
// ------------------------------------------------------------------------------------------------------------------------
#ifndef CONV_ISSUERS
#define CONV_ISSUERS 2 /* MMA-issuing warps = TMEM accumulator stages (tile lt -> issuer / stage lt % CONV_ISSUERS) */
#endif
// Shared-memory stages: every issuer owns a private ring of CONV_RING stages that the TMA producer fills with the loads of
// that issuer's tiles only.  Each (producer, issuer) pair is then a single-producer / single-consumer ring, whose phase
// parities cannot alias; with ONE ring shared by several issuers an issuer may have to wait for the second phase of a stage
// whose first phase (another issuer's tile) has not completed yet, and mbarrier.try_wait.parity cannot tell "phase k + 1
// pending" from "phase k - 1 done" (found with four issuers: launch failure; latent with two).
#ifndef CONV_RING_STAGES
#define CONV_RING_STAGES 3
#endif
constexpr int CONV_RING = CONV_RING_STAGES;
constexpr int CONV_STAGES = CONV_ISSUERS * CONV_RING;
constexpr int CONV_ROWS = BLOCK_M + 2;                       // rows m0 - 1 .. m0 + 128 of one filter row: the taps kx = 0, 1, 2
constexpr int CONV_A_BYTES = ((CONV_ROWS * BLOCK_K * 2 + 1023) / 1024) * 1024;  // stage stride: 1024-byte aligned (17 KB)
constexpr int CONV_A_TX = CONV_ROWS * BLOCK_K * 2;           // bytes one TMA box delivers
// epilogue: BN / 32 groups of four warps (one warp per TMEM lane quarter and 32-column slice), 2 KB transpose buffer per warp
template <int BN> __host__ __device__ constexpr int conv_threads() { return 32 * (1 + CONV_ISSUERS) + 128 * (BN / 32); }  // TMA, MMA issuers, epilogue groups
template <int BN> __host__ __device__ constexpr int conv_smem_bytes(int kblocks) {
  return kblocks * tile_b_bytes<BN>() + CONV_STAGES * CONV_A_BYTES + 1024 + 2048 + 4 * (BN / 32) * 2048;
}

template <int BLOCK_N>
__global__ void __launch_bounds__(conv_threads<BLOCK_N>(), 1) k_conv3x3_bf16(const __grid_constant__ CUtensorMap map_x,
                                                             const __grid_constant__ CUtensorMap map_w,
                                                             const float* __restrict__ bias, __nv_bfloat16* __restrict__ y_out,
                                                             int Mp, int N, int C, int H, int W, int relu) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr int TILE_B_BYTES = tile_b_bytes<BLOCK_N>();
  constexpr uint32_t TMEM_COLS = CONV_ISSUERS * BLOCK_N;
  constexpr int EPI0 = 1 + CONV_ISSUERS;  // first epilogue warp
  static_assert(CONV_ISSUERS == 2 || CONV_ISSUERS == 4, "TMEM columns must be a power of two");
  const int cblocks = C / BLOCK_K, num_kb = 9 * cblocks;
  uint8_t* smem_b = smem;                                  // [num_kb] weight blocks, resident
  uint8_t* smem_a = smem + num_kb * TILE_B_BYTES;          // [CONV_STAGES] activation tiles of CONV_ROWS rows
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_a + CONV_STAGES * CONV_A_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + CONV_STAGES;
  uint64_t* tmem_full = bars + 2 * CONV_STAGES;
  uint64_t* tmem_empty = bars + 2 * CONV_STAGES + CONV_ISSUERS;
  uint64_t* b_full = bars + 2 * CONV_STAGES + 2 * CONV_ISSUERS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * CONV_STAGES + 2 * CONV_ISSUERS + 1);
  float* bias_s = reinterpret_cast<float*>(bars + 2 * CONV_STAGES + 2 * CONV_ISSUERS + 2);
  uint4* xpose = reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(bars) + 2048);  // [epilogue warp][32 rows][4 chunks]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = (Mp + BLOCK_M - 1) / BLOCK_M;
  const int PW = W + 2, PH = H + 2;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_x)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_w)) : "memory");
    for (int s = 0; s < CONV_STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int a = 0; a < CONV_ISSUERS; a++) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 128 * (BLOCK_N / 32)); }
    mbar_init(b_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x >= 32 * EPI0)
    for (int c = threadIdx.x - 32 * EPI0; c < BLOCK_N; c += blockDim.x - 32 * EPI0) bias_s[c] = (bias && c < N) ? bias[c] : 0.0f;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // (broadcast: the compiler then keeps the TMEM address in a uniform register instead of re-electing a lane per MMA)
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer =====
      mbar_expect_tx(b_full, (uint32_t)(num_kb * TILE_B_BYTES));
      for (int kb = 0; kb < num_kb; kb++) tma_load_2d(smem_b + kb * TILE_B_BYTES, &map_w, b_full, kb * BLOCK_K, 0);
      const int loads_per_tile = 3 * cblocks;
      int lt = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, lt++) {
        const int m0 = tile * BLOCK_M;
        const int ring = (lt % CONV_ISSUERS) * CONV_RING;       // the ring of the issuer that owns this tile
        int it = (lt / CONV_ISSUERS) * loads_per_tile;          // position in that ring's own load sequence
        for (int ky = 0; ky < 3; ky++)
          for (int cb = 0; cb < cblocks; cb++, it++) {
            const int s = ring + it % CONV_RING;
            mbar_wait(&empty[s], ((it / CONV_RING) & 1) ^ 1);
            mbar_expect_tx(&full[s], CONV_A_TX);
            // rows m0 + (ky - 1) PW - 1 .. + 129: serves kx = 0, 1, 2 at row offsets 0, 1, 2; rows < 0 or >= Mp: zero fill
            tma_load_2d(smem_a + s * CONV_A_BYTES, &map_x, &full[s], cb * BLOCK_K, m0 + (ky - 1) * PW - 1);
          }
      }
    }
  } else if (warp < EPI0) {
    // ===== CONV_ISSUERS MMA issuers: warp 1 + j drives the tiles lt = j (mod CONV_ISSUERS) of this CTA into TMEM stage j
    // from its own ring of shared-memory stages.  A 128 x 64 x 16 MMA lasts 32 cycles, less than one thread needs to
    // issue it, so two threads issue into the tensor pipe.
    {  // the whole warp runs the loop (convergent); elect.sync inside the MMA / commit helpers picks the issuing lane
      mbar_wait(b_full, 0);
      const int loads_per_tile = 3 * cblocks;
      const int issuer = __shfl_sync(0xffffffffu, warp - 1, 0);  // broadcast: loop state and descriptors stay in uniform registers
      const int ring = issuer * CONV_RING;
      int lt = issuer;
      for (int tile = blockIdx.x + issuer * gridDim.x; tile < num_tiles; tile += CONV_ISSUERS * gridDim.x, lt += CONV_ISSUERS) {
        int it = (lt / CONV_ISSUERS) * loads_per_tile;
        const int as = lt % CONV_ISSUERS;
        mbar_wait(&tmem_empty[as], ((lt / CONV_ISSUERS) & 1) ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * BLOCK_N);
        uint32_t first = 0u;  // the first MMA of a tile overwrites the accumulator
        for (int ky = 0; ky < 3; ky++)
          for (int cb = 0; cb < cblocks; cb++, it++) {
            const int s = ring + it % CONV_RING;
            mbar_wait(&full[s], (it / CONV_RING) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint64_t da0 = umma_smem_desc(smem_a + s * CONV_A_BYTES);            // kx = 1, 2: + 8, + 16 (one row = 128 B)
            const uint64_t db0 = umma_smem_desc(smem_b + (ky * 3 * cblocks + cb) * TILE_B_BYTES);
#pragma unroll
            for (int kx = 0; kx < 3; kx++) {
#pragma unroll
              for (int k = 0; k < BLOCK_K / UMMA_K; k++) {
                umma_f16_elect(tmem_d, da0 + (uint64_t)(8 * kx + 2 * k), db0 + (uint64_t)(kx * cblocks * (TILE_B_BYTES >> 4) + 2 * k),
                         instr_desc<BLOCK_N>(), first);
                first = 1u;
              }
            }
            umma_commit_elect(&empty[s]);
          }
        umma_commit_elect(&tmem_full[as]);
      }
    }
  } else {  // ===== epilogue: TMEM -> bias + ReLU, zero at border positions -> bf16 NHWC =====
    // warp (EPI0 + 4 g + q') owns TMEM lanes of quarter (warp & 3) and the 32 accumulator columns [32 g, 32 g + 32): one
    // tcgen05.ld per tile, packed to bf16 by the row's thread, transposed through 2 KB of shared memory so that the global
    // stores are 64-byte row segments (four lanes per row) instead of one 16-byte piece per row and instruction.
    const int quarter = warp & 3, c0 = ((warp - EPI0) >> 2) * 32;
    uint4* xp = xpose + (warp - EPI0) * 128;
    int lt = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, lt++) {
      const int m0 = tile * BLOCK_M, as = lt % CONV_ISSUERS;
      mbar_wait(&tmem_full[as], (lt / CONV_ISSUERS) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int row = m0 + quarter * 32 + lane;
      const int px = row % PW, py = (row / PW) % PH;
      const bool interior = px >= 1 && px <= W && py >= 1 && py <= H;
      uint32_t v[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * BLOCK_N + c0);
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
            "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
            "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
            "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(taddr)
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(&tmem_empty[as]);  // the accumulator slice is in registers: hand the stage back to the MMA warp
#pragma unroll
      for (int c = 0; c < 4; c++) {  // chunk c = columns [c0 + 8 c, +8) of this thread's row
        uint4 o = make_uint4(0u, 0u, 0u, 0u);
        if (interior) {
          uint32_t* ow = &o.x;
#pragma unroll
          for (int q = 0; q < 4; q++) {
            float a = __uint_as_float(v[8 * c + 2 * q]) + bias_s[c0 + 8 * c + 2 * q], b = __uint_as_float(v[8 * c + 2 * q + 1]) + bias_s[c0 + 8 * c + 2 * q + 1];
            if (relu) { a = fmaxf(a, 0.0f); b = fmaxf(b, 0.0f); }
            const __nv_bfloat162 p2 = __floats2bfloat162_rn(a, b);
            ow[q] = *reinterpret_cast<const uint32_t*>(&p2);
          }
        }
        xp[lane * 4 + (c ^ ((lane >> 1) & 3))] = o;
      }
      __syncwarp();
      const int cc = lane & 3;  // this lane stores chunk cc of rows (lane >> 2) + 8 j
      if (c0 + 8 * cc + 7 < N) {  // N % 8 == 0
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int r = (lane >> 2) + 8 * j, grow = m0 + quarter * 32 + r;
          if (grow < Mp) *reinterpret_cast<uint4*>(y_out + (size_t)grow * N + c0 + 8 * cc) = xp[r * 4 + (cc ^ ((r >> 1) & 3))];
        }
      }
      __syncwarp();
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
}

// im2col for NHWC activations: out[m][k] (bf16, k = (kh * KW + kw) * C + c, zero beyond 9C up to Kpad) for output pixel
// m = (n * Ho + oy) * Wo + ox; input uint8 (observations) or bf16 (activations).  One thread per 8 consecutive k (16-byte
// store); when C % 8 == 0 the eight values are one 16-byte load.  conv(x) = im2col(x) . W^T then runs on k_linear_bf16.
template <typename InT>
__global__ void __launch_bounds__(256) k_im2col(const InT* __restrict__ in, __nv_bfloat16* __restrict__ out, int n, int H, int W, int C,
                                                int KS, int stride, int pad, int Ho, int Wo, int Kpad) {
  const int chunks = Kpad >> 3;
  const size_t total = (size_t)n * Ho * Wo * chunks;
  const int Kreal = KS * KS * C;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int ch = (int)(t % chunks);
    const size_t m = t / chunks;
    const int ox = (int)(m % Wo), oy = (int)((m / Wo) % Ho), e = (int)(m / ((size_t)Wo * Ho));
    const int k0 = ch << 3;
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    __nv_bfloat16* ov = reinterpret_cast<__nv_bfloat16*>(&o);
    bool done = false;
    if (sizeof(InT) == 2 && (C & 7) == 0 && k0 < Kreal) {  // eight channels of one tap: a single 16-byte load
      const int tap = k0 / C, c = k0 - tap * C, ky = tap / KS, kx = tap - ky * KS;
      const int iy = oy * stride - pad + ky, ix = ox * stride - pad + kx;
      if (iy >= 0 && iy < H && ix >= 0 && ix < W)
        o = *reinterpret_cast<const uint4*>(in + (((size_t)e * H + iy) * W + ix) * C + c);
      done = true;
    }
    if (!done) {
#pragma unroll
      for (int q = 0; q < 8; q++) {
        const int k = k0 + q;
        float val = 0.0f;
        if (k < Kreal) {
          const int tap = k / C, c = k - tap * C, ky = tap / KS, kx = tap - ky * KS;
          const int iy = oy * stride - pad + ky, ix = ox * stride - pad + kx;
          if (iy >= 0 && iy < H && ix >= 0 && ix < W) val = (float)in[(((size_t)e * H + iy) * W + ix) * C + c];
        }
        ov[q] = __float2bfloat16_rn(val);
      }
    }
    *reinterpret_cast<uint4*>(out + m * Kpad + k0) = o;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// row-major bf16 [rows, K] -> 2-D tensor map with a {BLOCK_K, 128} box and the 128-byte swizzle (zero fill out of bounds)
static int make_map(CUtensorMap* map, const void* base, int rows, int K, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return -1;
  const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  const cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1u, 1u};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS ? 0 : -2;
}

}  // namespace pcgrl_linear

static thread_local char g_linear_err[256] = "";
extern "C" const char* pcgrl_linear_last_error(void) { return g_linear_err; }

// Y[M,N] (row-major, fp32 or bf16) = act(X[M,K] . W[N,K]^T + bias[N]); X, W bf16 row-major device pointers (16-byte aligned,
// K % 8 == 0; N % 4 == 0 for fp32 output, N % 8 == 0 for bf16 output); bias may be NULL; relu != 0 applies max(., 0).
// Enqueues on `stream`; 0 = OK.
static int linear_launch(const void* x_bf16, const void* w_bf16, const float* bias, void* y, int M, int N, int K, int relu,
                         int out_bf16, int pad_h, int pad_w, void* stream) {
  using namespace pcgrl_linear;
  if (!x_bf16 || !w_bf16 || !y) { snprintf(g_linear_err, sizeof(g_linear_err), "NULL argument"); return -1; }
  if (M <= 0 || N <= 0 || K <= 0 || (K & 7) || (N & (out_bf16 ? 7 : 3))) { snprintf(g_linear_err, sizeof(g_linear_err), "need M, N, K > 0, K %% 8 == 0, N %% 4 == 0 (fp32 out) or N %% 8 == 0 (bf16 out)"); return -1; }
  if (((uintptr_t)x_bf16 | (uintptr_t)w_bf16 | (uintptr_t)y) & 15) { snprintf(g_linear_err, sizeof(g_linear_err), "pointers must be 16-byte aligned"); return -1; }
  static thread_local int sm_count = 0;
  if (!sm_count) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev); if (sm_count < 1) sm_count = 148; }
  const int tiles_m = (M + BLOCK_M - 1) / BLOCK_M;
  const bool wide = (N >= 256) && tiles_m * ((N + 255) / 256) >= sm_count;  // enough 128 x 256 tiles to fill every SM
  const int bn = wide ? 256 : 128;
  CUtensorMap mx, mw;
  if (make_map(&mx, x_bf16, M, K, BLOCK_M) || make_map(&mw, w_bf16, N, K, bn)) { snprintf(g_linear_err, sizeof(g_linear_err), "cuTensorMapEncodeTiled failed"); return -1; }
  static thread_local bool configured = false;
  if (!configured) {
    cudaError_t c1 = cudaFuncSetAttribute(k_linear_bf16<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<128>());
    cudaError_t c2 = cudaFuncSetAttribute(k_linear_bf16<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<256>());
    if (c1 != cudaSuccess || c2 != cudaSuccess) { snprintf(g_linear_err, sizeof(g_linear_err), "smem opt-in: %s", cudaGetErrorString(c1 != cudaSuccess ? c1 : c2)); return (int)(c1 != cudaSuccess ? c1 : c2); }
    configured = true;
  }
  const int tiles = tiles_m * ((N + bn - 1) / bn), grid = tiles < sm_count ? tiles : sm_count;
  if (wide) k_linear_bf16<256><<<grid, THREADS, smem_bytes<256>(), (cudaStream_t)stream>>>(mx, mw, bias, y, M, N, K, relu, out_bf16, pad_h, pad_w);
  else k_linear_bf16<128><<<grid, THREADS, smem_bytes<128>(), (cudaStream_t)stream>>>(mx, mw, bias, y, M, N, K, relu, out_bf16, pad_h, pad_w);
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) { snprintf(g_linear_err, sizeof(g_linear_err), "launch: %s", cudaGetErrorString(ce)); return (int)ce; }
  return 0;
}

extern "C" int pcgrl_linear_bf16_ex(const void* x_bf16, const void* w_bf16, const float* bias, void* y, int M, int N, int K, int relu,
                                    int out_bf16, void* stream) {
  return linear_launch(x_bf16, w_bf16, bias, y, M, N, K, relu, out_bf16, 0, 0, stream);
}

// The same GEMM with bf16 output rows scattered into a zero-bordered NHWC buffer: row m = (e, y, x) of an [M / (H W), H, W]
// image batch is written at (e, y + 1, x + 1) of y_padded[M / (H W)][H + 2][W + 2][N] (the caller zeroes the buffer once; the
// border is never written).  This is how the first convolution of a fully convolutional policy (im2col: few input channels)
// hands its activations to pcgrl_conv3x3_bf16.
extern "C" int pcgrl_linear_bf16_pad(const void* x_bf16, const void* w_bf16, const float* bias, void* y_padded, int M, int N, int K,
                                     int relu, int H, int W, void* stream) {
  if (H <= 0 || W <= 0 || M % (H * W)) { snprintf(g_linear_err, sizeof(g_linear_err), "M must be a multiple of H * W"); return -1; }
  return linear_launch(x_bf16, w_bf16, bias, y_padded, M, N, K, relu, 1, H, W, stream);
}

// 3 x 3 / stride 1 / SAME convolution + bias (+ ReLU) on zero-bordered NHWC bf16 activations (k_conv3x3_bf16):
//   x_padded [n][H + 2][W + 2][C], C % 64 == 0;  w [Npad][9 C] bf16 with k = (ky * 3 + kx) * C + c, Npad % 8 == 0, Npad <= 64;
//   y_padded [n][H + 2][W + 2][Npad]: interior = act(conv + bias), border = 0.  Enqueues on `stream`; 0 = OK.
extern "C" int pcgrl_conv3x3_bf16(const void* x_padded, const void* w_bf16, const float* bias, void* y_padded, int n, int H, int W, int C,
                                  int Npad, int relu, void* stream) {
  using namespace pcgrl_linear;
  if (!x_padded || !w_bf16 || !y_padded) { snprintf(g_linear_err, sizeof(g_linear_err), "NULL argument"); return -1; }
  if (n <= 0 || H <= 0 || W <= 0 || C <= 0 || (C % BLOCK_K) || Npad <= 0 || (Npad & 7) || Npad > 64) {
    snprintf(g_linear_err, sizeof(g_linear_err), "need C %% 64 == 0 and Npad %% 8 == 0, Npad <= 64");
    return -1;
  }
  if (((uintptr_t)x_padded | (uintptr_t)w_bf16 | (uintptr_t)y_padded) & 15) { snprintf(g_linear_err, sizeof(g_linear_err), "pointers must be 16-byte aligned"); return -1; }
  const long long mp = (long long)n * (H + 2) * (W + 2);
  if (mp >= (1ll << 31) - 4096) { snprintf(g_linear_err, sizeof(g_linear_err), "too many padded positions: split the batch"); return -1; }
  const int Mp = (int)mp, kblocks = 9 * (C / BLOCK_K), bn = Npad <= 32 ? 32 : 64;
  const int smem = bn == 32 ? conv_smem_bytes<32>(kblocks) : conv_smem_bytes<64>(kblocks);
  if (smem > 227 * 1024) { snprintf(g_linear_err, sizeof(g_linear_err), "C too large for the resident-weight layout"); return -1; }
  static thread_local int sm_count = 0;
  if (!sm_count) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev); if (sm_count < 1) sm_count = 148; }
  CUtensorMap mx, mw;
  if (make_map(&mx, x_padded, Mp, C, CONV_ROWS) || make_map(&mw, w_bf16, Npad, 9 * C, bn)) { snprintf(g_linear_err, sizeof(g_linear_err), "cuTensorMapEncodeTiled failed"); return -1; }
  static thread_local int configured32 = 0, configured64 = 0;
  int& configured = bn == 32 ? configured32 : configured64;
  if (configured < smem) {
    const cudaError_t c1 = bn == 32 ? cudaFuncSetAttribute(k_conv3x3_bf16<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)
                                    : cudaFuncSetAttribute(k_conv3x3_bf16<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (c1 != cudaSuccess) { snprintf(g_linear_err, sizeof(g_linear_err), "smem opt-in: %s", cudaGetErrorString(c1)); return (int)c1; }
    configured = smem;
  }
  const int tiles = (Mp + BLOCK_M - 1) / BLOCK_M, grid = tiles < sm_count ? tiles : sm_count;
  if (bn == 32) k_conv3x3_bf16<32><<<grid, conv_threads<32>(), smem, (cudaStream_t)stream>>>(mx, mw, bias, (__nv_bfloat16*)y_padded, Mp, Npad, C, H, W, relu);
  else k_conv3x3_bf16<64><<<grid, conv_threads<64>(), smem, (cudaStream_t)stream>>>(mx, mw, bias, (__nv_bfloat16*)y_padded, Mp, Npad, C, H, W, relu);
  const cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) { snprintf(g_linear_err, sizeof(g_linear_err), "launch: %s", cudaGetErrorString(ce)); return (int)ce; }
  return 0;
}

extern "C" int pcgrl_linear_bf16(const void* x_bf16, const void* w_bf16, const float* bias, float* y, int M, int N, int K, int relu,
                                 void* stream) {
  return pcgrl_linear_bf16_ex(x_bf16, w_bf16, bias, y, M, N, K, relu, 0, stream);
}

// im2col of NHWC activations for a KS x KS convolution (stride, zero padding `pad`): in [n][H][W][C] uint8 (in_bf16 == 0) or
// bf16 -> out [n * Ho * Wo][Kpad] bf16 with k = (ky * KS + kx) * C + c, zero-filled up to Kpad (Kpad % 8 == 0, >= KS*KS*C).
// pad = -1 skips a one-cell border of the input: a VALID convolution over the interior of a zero-bordered buffer.
// conv + bias + ReLU = pcgrl_im2col followed by pcgrl_linear_bf16_ex on weights laid out [Cout][Kpad].
extern "C" int pcgrl_im2col(const void* in, int in_bf16, void* out_bf16, int n, int H, int W, int C, int ksize, int stride, int pad,
                            int Kpad, void* stream) {
  using namespace pcgrl_linear;
  if (!in || !out_bf16) { snprintf(g_linear_err, sizeof(g_linear_err), "NULL argument"); return -1; }
  if (n <= 0 || H <= 0 || W <= 0 || C <= 0 || ksize <= 0 || stride <= 0 || pad < -1 || (Kpad & 7) || Kpad < ksize * ksize * C) {
    snprintf(g_linear_err, sizeof(g_linear_err), "bad im2col arguments (Kpad %% 8 == 0, Kpad >= ksize^2 * C)");
    return -1;
  }
  const int Ho = (H + 2 * pad - ksize) / stride + 1, Wo = (W + 2 * pad - ksize) / stride + 1;
  if (Ho <= 0 || Wo <= 0) { snprintf(g_linear_err, sizeof(g_linear_err), "convolution output is empty"); return -1; }
  const size_t total = (size_t)n * Ho * Wo * (Kpad >> 3);
  const unsigned blocks = (unsigned)((total + 255) / 256 < 148u * 32u ? (total + 255) / 256 : 148u * 32u);
  if (in_bf16) k_im2col<__nv_bfloat16><<<blocks, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)in, (__nv_bfloat16*)out_bf16, n, H, W, C, ksize, stride, pad, Ho, Wo, Kpad);
  else k_im2col<uint8_t><<<blocks, 256, 0, (cudaStream_t)stream>>>((const uint8_t*)in, (__nv_bfloat16*)out_bf16, n, H, W, C, ksize, stride, pad, Ho, Wo, Kpad);
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) { snprintf(g_linear_err, sizeof(g_linear_err), "launch: %s", cudaGetErrorString(ce)); return (int)ce; }
  return 0;
}

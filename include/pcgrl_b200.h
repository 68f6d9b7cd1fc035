/*
 * pcgrl_b200.h -- C ABI of the B200-native batched PCGRL environment hot path.
 *
 * Drop-in boundary for the one data-parallel hot path of amidos2006/gym-pcgrl:
 * N lock-step PcgrlEnv.step()/reset() calls (reference: gym_pcgrl/envs/pcgrl_env.py:66-76,129-150),
 * i.e. Representation.update -> Problem.get_stats -> get_reward / get_episode_over.
 *
 * The reference is pure Python and has no FFI of its own; the entry points below are what a
 * ctypes binding inside the reference's PcgrlEnv / a VecEnv worker (utils.py:60-71) would call
 * (see INTEGRATION.md).  Conventions:
 *   - plain C, POD structs, raw pointers + sizes, no torch / C++ types;
 *   - every `pcgrl_*` device entry point only ENQUEUES work on the given CUDA stream
 *     (cudaStream_t passed as void*); it never synchronises, never allocates device memory and
 *     never throws.  The `*_host` entry point is the exception: it takes HOST buffers, copies
 *     in/out and synchronises the stream before returning;
 *   - all buffers are caller-owned (torch tensors in the Python host layer);
 *   - return value: 0 = OK, <0 = invalid argument (text in pcgrl_last_error()), >0 = cudaError_t.
 *
 * The same POD structs are used by the CPU oracle (oracle/pcgrl_oracle.c, test infrastructure).
 */
#ifndef PCGRL_B200_H
#define PCGRL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PCGRL_ABI_VERSION 2

/* PROBLEMS registry (reference: gym_pcgrl/envs/probs/__init__.py:9-16) */
enum { PCGRL_PROB_BINARY = 0, PCGRL_PROB_ZELDA = 1, PCGRL_PROB_SOKOBAN = 2, PCGRL_PROB_DDAVE = 3,
       PCGRL_PROB_MDUNGEON = 4, PCGRL_PROB_SMB = 5, PCGRL_NUM_PROBLEMS = 6 };
/* REPRESENTATIONS (reference: gym_pcgrl/envs/reps/__init__.py:9-16).  Action layout (int32 per env):
 *   narrow      [1]  0 = keep, a>0 writes tile a-1 at the cursor                  (narrow_rep.py:99-114)
 *   turtle      [1]  0..3 move, a>=4 writes tile a-4                               (turtle_rep.py:101-129)
 *   wide        [3]  x, y, tile                                                    (wide_rep.py:67-70)
 *   narrowcast  [2]  type (0 keep, 1 cursor cell, 2 3x3 block), tile               (narrow_cast_rep.py:36-59)
 *   narrowmulti [9]  one entry per cell of the 3x3 block, 0 = keep, a>0 tile a-1   (narrow_multi_rep.py:39-59)
 *   turtlecast  [2]  type (0..3 move, 4 cursor cell, 5 3x3 block), tile            (turtle_cast_rep.py:38-76) */
enum { PCGRL_REP_NARROW = 0, PCGRL_REP_TURTLE = 1, PCGRL_REP_WIDE = 2, PCGRL_REP_NARROWCAST = 3,
       PCGRL_REP_NARROWMULTI = 4, PCGRL_REP_TURTLECAST = 5, PCGRL_NUM_REPS = 6 };
#define PCGRL_MAX_ACTION_DIM 9

#define PCGRL_MAX_DIM 32     /* width, height <= 32: one bitboard row per warp lane (all problems but smb) */
#define PCGRL_SMB_MAX_W 122  /* smb keeps a byte map: width <= 122 (default 114), 3 <= height <= 16       */
#define PCGRL_SMB_MAX_H 16
#define PCGRL_MAX_TILES 8    /* zelda / mdungeon alphabets                                    */
#define PCGRL_MAX_STATS 16   /* stride of every stats row (one 64-byte line)                  */
#define PCGRL_MAX_REWARD_TERMS 12
/* info_stats columns written next to the statistics: the env counters at the END of the step, before any
 * auto-reset (info["iterations"], info["changes"], pcgrl_env.py:144-145) */
#define PCGRL_INFO_ITERATION 14
#define PCGRL_INFO_CHANGES 15
#define PCGRL_MT_WORDS 625   /* MT19937: 624 state words + word 624 = position (numpy `pos`)  */

/* flags */
#define PCGRL_FLAG_RANDOM_TILE 1u   /* narrow: random cursor (narrow_rep.py:104-106) vs raster scan (:107-113) */
#define PCGRL_FLAG_WARP 2u          /* turtle: wrap at the edges (turtle_rep.py:105-125)                       */
#define PCGRL_FLAG_RANDOM_START 4u  /* representation.py:41-45                                                */
#define PCGRL_FLAG_RANDOM_PROBS 8u  /* binary_prob.py:68-72: redraw tile probabilities at every reset          */
#define PCGRL_FLAG_AUTO_RESET 16u   /* VecEnv semantics: an env that is done is reset inside step()            */
#define PCGRL_FLAG_HEAT_U16 32u     /* heat map elements are uint16 (required when max_changes > 255)          */
#define PCGRL_FLAG_FULL_STATS 64u   /* binary / zelda: recompute every statistic of the whole map after every edit
                                       instead of the incremental updates (same results; A/B and test switch) */

/*
 * Stats row layout (int32, stride PCGRL_MAX_STATS), one order per problem == the key order of
 * each Problem.get_stats dict in the reference:
 *   binary   (binary_prob.py:81-86)     regions, path-length
 *   zelda    (zelda_prob.py:80-112)     player, key, door, enemies, regions, nearest-enemy, path-length
 *   sokoban  (sokoban_prob.py:133-145)  player, crate, target, regions, dist-win, len(solution)
 *   ddave    (ddave_prob.py:149-169)    player, dist-floor, exit, diamonds, key, spikes, regions,
 *                                       num-jumps, col-diamonds, dist-win, sol-length
 *   mdungeon (mdungeon_prob.py:151-171) player, exit, potions, treasures, enemies, regions,
 *                                       col-potions, col-treasures, col-enemies, dist-win, sol-length
 *   smb      (smb_prob.py:126-148)      dist-floor, disjoint-tubes, enemies, empty, noise, jumps, jumps-dist, dist-win
 *
 * iparam[] (integer thresholds set by Problem.adjust_param):
 *   binary   [0] target_path
 *   zelda    [0] max_enemies  [1] target_enemy_dist [2] target_path
 *   sokoban  [0] max_crates   [1] target_solution
 *   ddave    [0] max_diamonds [1] min_spikes [2] target_jumps [3] target_solution
 *   mdungeon [0] max_enemies  [1] max_potions [2] max_treasures [3] target_solution ; dparam[0] target_col_enemies
 *   smb      [0] min_empty    [1] min_enemies [2] max_enemies   [3] min_jumps
 *
 * reward_weight[] is in the order the terms are SUMMED in each Problem.get_reward (fp64, left to right):
 *   binary   regions, path-length                                               (binary_prob.py:105-106)
 *   zelda    player, key, door, enemies, regions, nearest-enemy, path-length    (zelda_prob.py:136-142)
 *   sokoban  player, crate, target, regions, ratio, dist-win, sol-length        (sokoban_prob.py:169-175)
 *   ddave    player, dist-floor, exit, spikes, diamonds, key, regions, num-jumps, dist-win, sol-length (ddave_prob.py:196-205)
 *   mdungeon player, exit, enemies, treasures, potions, regions, col-enemies, dist-win, sol-length     (mdungeon_prob.py:197-205)
 *   smb      dist-floor, disjoint-tubes, enemies, empty, noise, jumps, jumps-dist, dist-win            (smb_prob.py:163-170)
 */
typedef struct pcgrl_config {
  int32_t problem;         /* PCGRL_PROB_*                                                     */
  int32_t representation;  /* PCGRL_REP_*                                                      */
  int32_t width, height;   /* Problem._width/_height (problem.py:12-13)                        */
  int32_t num_tiles;       /* len(get_tile_types())                                            */
  int32_t max_changes;     /* PcgrlEnv._max_changes (pcgrl_env.py:33,109)                      */
  int32_t max_iterations;  /* PcgrlEnv._max_iterations (pcgrl_env.py:34,110)                   */
  uint32_t flags;          /* PCGRL_FLAG_*                                                     */
  int32_t solver_power;    /* ddave/mdungeon/sokoban _solver_power                             */
  int32_t iparam[7];
  double dparam[2];
  double reward_weight[PCGRL_MAX_REWARD_TERMS];
  double tile_prob[PCGRL_MAX_TILES]; /* Problem._prob values in tile order (un-normalised)     */
} pcgrl_config;

/*
 * Caller-owned state of n environments.  Device pointers for pcgrl_* (host pointers for the
 * oracle).  n is the batch dimension everywhere.
 */
typedef struct pcgrl_buffers {
  uint8_t* map;         /* [n][H][W]  Representation._map, tile indices                         */
  void* heatmap;        /* [n][H][W]  PcgrlEnv._heatmap counts (value-equal to the fp64 array): uint8, or
                                      uint16 when PCGRL_FLAG_HEAT_U16 is set                        */
  uint8_t* pos;         /* [n][2]     (x, y) cursor of narrow / turtle; unused for wide         */
  int32_t* iteration;   /* [n]        PcgrlEnv._iteration                                       */
  int32_t* changes;     /* [n]        PcgrlEnv._changes                                         */
  int32_t* stats;       /* [n][PCGRL_MAX_STATS]  PcgrlEnv._rep_stats (live state)               */
  int32_t* start_stats; /* [n][PCGRL_MAX_STATS]  Problem._start_stats                           */
  int32_t* info_stats;  /* [n][PCGRL_MAX_STATS]  stats at the end of the step, before any auto-reset (-> info) */
  double* reward;       /* [n]        step output                                              */
  uint8_t* done;        /* [n]        step output                                              */
  uint32_t* rng;        /* [n][2][PCGRL_MT_WORDS] MT19937 streams: [0] representation, [1] problem */
  double* tile_prob;    /* [n][PCGRL_MAX_TILES]  per-env Problem._prob values (binary redraws them) */
  uint8_t* start_map;   /* [n][H][W]  Representation._old_map (used when RANDOM_START is off)   */
  uint8_t* start_valid; /* [n]        1 once _old_map holds a map                               */
  void* scratch;        /* solver work space, pcgrl_scratch_bytes() bytes (may be NULL if that is 0) */
  size_t scratch_bytes;
  int32_t* status;      /* [4] device-side diagnostics: [0] != 0 -> a capacity limit was hit    */
} pcgrl_buffers;

int pcgrl_abi_version(void);
const char* pcgrl_last_error(void);

/* Checks sizes/ids/limits of a config; 0 if the CUDA path supports it. */
int pcgrl_config_validate(const pcgrl_config* cfg);

/* Bytes of pcgrl_buffers.scratch needed for n_envs environments (0 for binary / zelda). */
size_t pcgrl_scratch_bytes(const pcgrl_config* cfg, int n_envs);

/* PcgrlEnv.reset() for every env with mask[i] != 0 (all if mask == NULL).  pcgrl_env.py:66-76 */
int pcgrl_reset(const pcgrl_config* cfg, const pcgrl_buffers* bufs, const uint8_t* mask_or_null,
                int n, void* stream);

/* PcgrlEnv.step(action) for n envs.  actions: int32 [n][adim], adim per representation as listed at
 * the PCGRL_REP_* enum (1 narrow/turtle, 3 wide, 2 narrowcast/turtlecast, 9 narrowmulti).  pcgrl_env.py:129-150 */
int pcgrl_step(const pcgrl_config* cfg, const pcgrl_buffers* bufs, const int32_t* actions, int n,
               void* stream);

/* T consecutive steps in one launch sequence: actions [T][n][adim]; reward_out [T][n], done_out [T][n]
 * (both may be NULL).  Equivalent to T calls of pcgrl_step. */
int pcgrl_rollout(const pcgrl_config* cfg, const pcgrl_buffers* bufs, const int32_t* actions,
                  double* reward_out, uint8_t* done_out, int T, int n, void* stream);

/* Stand-alone Problem.get_stats on n maps [n][H][W] -> stats_out [n][PCGRL_MAX_STATS]. */
int pcgrl_get_stats(const pcgrl_config* cfg, const uint8_t* maps, int32_t* stats_out, int n,
                    void* scratch, size_t scratch_bytes, int32_t* status, void* stream);

/* Seeds both MT19937 streams of env i with numpy's RandomState(seeds[i]) (init_genrand). */
int pcgrl_seed(const pcgrl_buffers* bufs, const uint32_t* seeds, int n, void* stream);

/*
 * End-to-end step with HOST buffers (the call a reference-side VecEnv binding makes): host actions in, host
 * observation / reward / done out, stream synchronised on return.  Host pointers should be pinned.
 *
 * The actions are read by the kernel straight from the host buffer when it is pinned and device-mapped (any
 * cudaHostAlloc / torch pin_memory allocation); pageable buffers are copied H2D through d_actions first.
 *
 * mode 0 (full):  pcgrl_step, then D2H of map / heatmap / pos / reward / done (and info_stats) in full.
 * mode 1 (delta): the step kernels additionally write reward / done / cursor in their final layout, a compacted
 *                 list of 8-byte change records (env, changed cell, new tile) and the fresh maps of the envs that
 *                 were auto-reset into a small device staging buffer; ONE D2H copy of that buffer returns and the
 *                 library patches the caller's host arrays, which therefore always hold the complete current
 *                 observation.  The host arrays must persist between calls; `synced` = 0 (set it after
 *                 pcgrl_reset or any device-side step that bypassed this call) makes the next call fall back to
 *                 a full copy and re-arm.  d_staging / h_staging: device and pinned-host scratch of
 *                 pcgrl_host_staging_bytes(cfg, n) bytes each.
 */
typedef struct pcgrl_host_io {
  const int32_t* actions; /* in  [n][adim]            */
  uint8_t* map;           /* out [n][H][W] or NULL    */
  void* heatmap;          /* out [n][H][W] or NULL (uint8, or uint16 with PCGRL_FLAG_HEAT_U16) */
  uint8_t* pos;           /* out [n][2]    or NULL    */
  double* reward;         /* out [n]                  */
  uint8_t* done;          /* out [n]                  */
  int32_t* info_stats;    /* out [n][PCGRL_MAX_STATS] or NULL */
  void* d_staging;        /* mode 1: device scratch   */
  void* h_staging;        /* mode 1: pinned host scratch */
  size_t staging_bytes;
  int32_t mode;           /* 0 full copies, 1 delta records, 2 direct (the kernels store into the pinned host arrays) */
  int32_t synced;         /* in/out, modes 1 and 2: host arrays are in sync with the device state */
  int64_t reset_base;     /* in/out, mode 1: running counters of staged whole-map updates / change records */
  int64_t change_base;    /*                 (library-maintained)                                            */
  int32_t pending;        /* library-maintained: a pcgrl_step_host_begin of this block awaits its _end (init 0) */
  int32_t reserved;
} pcgrl_host_io;
size_t pcgrl_host_staging_bytes(const pcgrl_config* cfg, int n);
int pcgrl_step_host(const pcgrl_config* cfg, const pcgrl_buffers* bufs, int32_t* d_actions,
                    pcgrl_host_io* io, int n, void* stream);
/*
 * The same call split in two, for asynchronous vector-env bindings (the step_async / step_wait pair of the reference's
 * SubprocVecEnv, utils.py:60-71 -- here per ENV GROUP: a caller that shards its batch over several pcgrl_buffers +
 * streams keeps every group in flight and serves whichever finishes first, so one env stuck in a capped A* search
 * (solver problems, smb) delays its own group only):
 *   pcgrl_step_host_begin  enqueues everything on `stream` and returns at once;
 *   pcgrl_step_host_end    wait != 0: blocks until the step is complete; wait == 0: returns 1 if it is still running
 *                          (call again later), else completes it.  On return 0 the host arrays hold the step's results.
 * pcgrl_step_host(...) == begin + end(wait = 1).  One step per io block may be in flight.
 */
int pcgrl_step_host_begin(const pcgrl_config* cfg, const pcgrl_buffers* bufs, int32_t* d_actions,
                          pcgrl_host_io* io, int n, void* stream);
int pcgrl_step_host_end(const pcgrl_config* cfg, const pcgrl_buffers* bufs, pcgrl_host_io* io, int n,
                        void* stream, int wait);

/*
 * pcgrl_rollout_host: T consecutive PcgrlEnv.step calls on HOST buffers in one call -- the open-loop form of the
 * reference's rollout loop (README.md:59-72: `for _ in range(T): obs, r, d, info = env.step(action)`) for callers
 * whose actions do not depend on the observations (random-action rollouts, replays of recorded trajectories).
 *   actions  in  [T][n][adim] (copied to d_actions in one H2D transfer; pin the buffer for full PCIe speed),
 *   reward   out [T][n], done out [T][n]: every step's results,
 *   map / heatmap / pos / info_stats: the observation after the last step (NULL = not wanted).
 * d_actions / d_reward / d_done are caller-owned device staging buffers of the same shapes.  Synchronises the
 * stream before returning.  Solver problems run through the env-asynchronous rollout kernel.
 */
typedef struct pcgrl_host_rollout_io {
  const int32_t* actions; /* in  [T][n][adim]         */
  double* reward;         /* out [T][n]               */
  uint8_t* done;          /* out [T][n]               */
  uint8_t* map;           /* out [n][H][W] or NULL    */
  void* heatmap;          /* out [n][H][W] or NULL (uint8, or uint16 with PCGRL_FLAG_HEAT_U16) */
  uint8_t* pos;           /* out [n][2]    or NULL    */
  int32_t* info_stats;    /* out [n][PCGRL_MAX_STATS] or NULL */
} pcgrl_host_rollout_io;
int pcgrl_rollout_host(const pcgrl_config* cfg, const pcgrl_buffers* bufs, int32_t* d_actions, double* d_reward,
                       uint8_t* d_done, pcgrl_host_rollout_io* io, int T, int n, void* stream);

/*
 * Batched observation / action wrappers (reference: gym_pcgrl/wrappers.py -- "next" row f1 of the scope table).
 *
 * pcgrl_obs_image: Cropped (:163-206) + OneHotEncoding (:67-104) + ToImage (:18-60) in one pass.
 *   maps [n][H][W] u8, pos [n][2] (needed when crop_size > 0) -> out [n][S][S][C], S = crop_size (or H x W when
 *   crop_size == 0), C = num_tiles when one_hot else 1.  pad_value = border tile (Cropped pads with it).
 *   out_dtype: 0 = uint8, 1 = float32 (values are 0/1 or small tile indices: equal in any dtype).
 * pcgrl_action_map: ActionMap.step (:139-154): flat [n] indices over (H, W, num_tiles) -> actions_out in the
 *   layout the wrapped representation expects ([n][3] for wide, [n] for narrow / turtle).
 */
int pcgrl_obs_image(const pcgrl_config* cfg, const uint8_t* maps, const uint8_t* pos, void* out, int n,
                    int crop_size, int pad_value, int one_hot, int out_dtype, void* stream);
int pcgrl_action_map(const pcgrl_config* cfg, const pcgrl_buffers* bufs, const int32_t* flat_actions,
                     int32_t* actions_out, int n, void* stream);

/*
 * Batched PcgrlEnv.render(mode="rgb_array") (pcgrl_env.py:160-173 = Problem.render, probs/problem.py:134-156, + the cursor
 * frame of the cursor representations, reps/narrow_rep.py:126-140, + convert("RGB")) -- SURVEY.md 8f row f4, no PIL:
 *   maps [n][height][width] u8, pos [n][2] (x, y) or NULL (no cursor frame), atlas [num_tiles][ts][ts][4] RGBA u8 (16-byte
 *   aligned) -> out [n][(height + 2 border_h) ts][(width + 2 border_w) ts][3] RGB u8.  tile_size ts % 4 == 0.
 */
int pcgrl_render(const uint8_t* maps, const uint8_t* pos_or_null, const uint8_t* atlas, uint8_t* out, int n,
                 int height, int width, int num_tiles, int border_w, int border_h, int border_tile, int tile_size,
                 void* stream);

/*
 * smb (gym_pcgrl/envs/probs/smb_prob.py, probs/smb/engine.py).  The smb ENVIRONMENT runs through the
 * generic entry points above with cfg->problem = PCGRL_PROB_SMB (byte map, width <= 122, 3 <= height <= 16, solver_power
 * <= 16000; max_changes = 319 at the default size, hence PCGRL_FLAG_HEAT_U16).  The two functions below are the
 * stand-alone SMBProblem.get_stats operator without a config:
 *   maps [n][height][width] u8 (tiles: empty, solid, enemy, brick, question, coin, tube) ->
 *   stats_out [n][PCGRL_MAX_STATS] i32: dist-floor, disjoint-tubes, enemies, empty, noise, jumps, jumps-dist, dist-win.
 * scratch = pcgrl_smb_scratch_bytes(n, solver_power) bytes of caller-owned device memory (== pcgrl_scratch_bytes of an smb
 * config): per-env bitmaps of the level cells the last search read + the tails of the A* open lists of the resident warps.
 */
size_t pcgrl_smb_scratch_bytes(int n, int solver_power);
int pcgrl_smb_get_stats(const uint8_t* maps, int32_t* stats_out, int n, int width, int height, int solver_power,
                        void* scratch, size_t scratch_bytes, void* stream);

/*
 * Host twins: the same operations on HOST pointers, computed on the calling CPU thread by the very
 * scalar `__host__ __device__` functions the kernels run -- for plumbing / CI without a GPU (BASELINE config 1 style
 * single-env runs).  They are separate, explicitly named entry points: the CUDA entry points above never fall back to
 * them.  Available for every built-in problem: smb and the solver problems' game models are the kernels' own scalar
 * functions (under a plain scalar search loop), binary / zelda restate the bitboard algorithm over row arrays.
 * pcgrl_buffers.scratch must hold pcgrl_scratch_bytes() bytes of host memory.
 */
int pcgrl_reset_cpu(const pcgrl_config* cfg, const pcgrl_buffers* bufs, const uint8_t* mask_or_null, int n);
int pcgrl_step_cpu(const pcgrl_config* cfg, const pcgrl_buffers* bufs, const int32_t* actions, int n);
int pcgrl_get_stats_cpu(const pcgrl_config* cfg, const uint8_t* maps, int32_t* stats_out, int n);

/*
 * The dense contraction next to the path: the fully connected layer of the reference's policy
 * feature extractors (model.py:15,23 `linear(layer_3, 'fc1', n_hidden=512)` + ReLU) as a tcgen05 / TMEM / TMA kernel.
 *   y[M][N] (f32) = act(x[M][K] . w[N][K]^T + bias[N]);  x, w: bf16 row-major device pointers, 16-byte aligned, K % 8 == 0,
 *   N % 4 == 0; bias may be NULL; relu != 0 applies max(., 0).  Only enqueues on `stream`.  Text of a failure:
 *   pcgrl_linear_last_error().
 */
int pcgrl_linear_bf16(const void* x_bf16, const void* w_bf16, const float* bias, float* y, int M, int N, int K,
                      int relu, void* stream);
/* same, with the output type selectable: out_bf16 != 0 writes bf16 (N % 8 == 0) -- the activations of the next layer */
int pcgrl_linear_bf16_ex(const void* x_bf16, const void* w_bf16, const float* bias, void* y, int M, int N, int K,
                         int relu, int out_bf16, void* stream);
/*
 * im2col of NHWC activations for a ksize x ksize convolution (`conv(...)` of stable-baselines' a2c.utils as used by
 * model.py:9-77; pad = 0 VALID, pad = ksize / 2 SAME): in [n][H][W][C] uint8 (in_bf16 == 0: the wrapper's observation) or
 * bf16 -> out [n * Ho * Wo][Kpad] bf16, k = (ky * ksize + kx) * C + c, zero-filled up to Kpad (Kpad % 8 == 0).
 * conv + bias + ReLU = pcgrl_im2col, then pcgrl_linear_bf16_ex on weights laid out [Cout][Kpad]; its [n * Ho * Wo][Cout]
 * output IS the NHWC activation tensor of the next layer.
 */
int pcgrl_im2col(const void* in, int in_bf16, void* out_bf16, int n, int H, int W, int C, int ksize, int stride,
                 int pad, int Kpad, void* stream);
/*
 * The 3 x 3 / stride 1 / SAME convolutions of the fully convolutional policies (model.py:25-77: c2..c8) as an IMPLICIT
 * GEMM -- no patch matrix.  Activations live in zero-bordered NHWC buffers P[n][H + 2][W + 2][C] (bf16): over the flattened
 * padded positions the input of filter tap (ky, kx) is the same 2-D tensor shifted by (ky - 1)(W + 2) + (kx - 1) rows, i.e. a
 * plain TMA tile load at a row offset; the weights stay resident in shared memory.
 *   pcgrl_conv3x3_bf16: x_padded [n][H+2][W+2][C], C % 64 == 0; w [Npad][9 C] bf16, k = (ky * 3 + kx) * C + c, Npad % 8 == 0,
 *                       Npad <= 64; y_padded [n][H+2][W+2][Npad]: interior = act(conv + bias), border = 0.
 *   pcgrl_linear_bf16_pad: the GEMM of pcgrl_linear_bf16_ex with its bf16 output row m = (e, y, x) written at (e, y+1, x+1) of
 *                       such a buffer (zeroed once by the caller) -- the hand-over from the first, im2col-based, layer.
 *   pcgrl_im2col with pad = -1 reads the interior of such a buffer (the VALID convolutions of the value branch).
 */
int pcgrl_conv3x3_bf16(const void* x_padded, const void* w_bf16, const float* bias, void* y_padded, int n, int H, int W,
                       int C, int Npad, int relu, void* stream);
int pcgrl_linear_bf16_pad(const void* x_bf16, const void* w_bf16, const float* bias, void* y_padded, int M, int N, int K,
                          int relu, int H, int W, void* stream);
const char* pcgrl_linear_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* PCGRL_B200_H */

// pcgrl_problems.cuh -- Problem.get_stats / get_reward / get_episode_over on bitboards, one warp per env.
// Reference: gym_pcgrl/envs/probs/{binary,zelda,sokoban,ddave,mdungeon}_prob.py (cited per function).
#pragma once
#include <math.h>

#include "pcgrl_device.cuh"

namespace pcgrl {

template <int PROB> struct ProblemTraits;
template <> struct ProblemTraits<PCGRL_PROB_BINARY>   { static constexpr int NPLANES = 1, NSTATS = 2;  static constexpr bool SOLVER = false; };
template <> struct ProblemTraits<PCGRL_PROB_ZELDA>    { static constexpr int NPLANES = 3, NSTATS = 7;  static constexpr bool SOLVER = false; };
template <> struct ProblemTraits<PCGRL_PROB_SOKOBAN>  { static constexpr int NPLANES = 3, NSTATS = 6;  static constexpr bool SOLVER = true; };
template <> struct ProblemTraits<PCGRL_PROB_DDAVE>    { static constexpr int NPLANES = 3, NSTATS = 11; static constexpr bool SOLVER = true; };
template <> struct ProblemTraits<PCGRL_PROB_MDUNGEON> { static constexpr int NPLANES = 3, NSTATS = 11; static constexpr bool SOLVER = true; };

// Everything of Problem.get_stats that is a map scan / flood fill / BFS.  For the solver problems the
// play-through statistics keep their "not playable" defaults and *need_solver tells the caller that the
// reference would call _run_game on this map.  st[] is warp-uniform.
// known_regions >= 0 (zelda): the caller knows the region count of this map (a single-cell edit that did not change
// whether the cell belongs to the region board leaves calc_num_regions unchanged), so the floods are skipped.
template <int PROB>
__device__ __forceinline__ void map_stats(const Board& b, const pcgrl_config& cfg, int lane, int* st, bool& need_solver,
                                          int known_regions = -1) {
  const int W = cfg.width, H = cfg.height;
  const uint32_t rm = row_mask(W, H, lane);
  need_solver = false;
#pragma unroll
  for (int i = 0; i < ProblemTraits<PROB>::NSTATS; i++) st[i] = 0;

  if (PROB == PCGRL_PROB_BINARY) {  // binary_prob.py:81-86
    regions_and_longest_path(type_mask<0x01u>(b, rm), lane, st[0], st[1]);
  } else if (PROB == PCGRL_PROB_ZELDA) {  // zelda_prob.py:80-112
    const uint32_t player = type_mask<0x04u>(b, rm), key = type_mask<0x08u>(b, rm), door = type_mask<0x10u>(b, rm);
    const uint32_t enemies = type_mask<0xE0u>(b, rm);
    st[0] = popc_all(player);
    st[1] = popc_all(key);
    st[2] = popc_all(door);
    st[3] = popc_all(enemies);
    st[4] = known_regions >= 0 ? known_regions
                               : count_regions(type_mask<0xEDu>(b, rm), lane);  // empty, player, key, bat, spider, scorpion
    if (st[0] == 1 && st[4] == 1) {
      if (st[3] > 0) {  // nearest enemy: first BFS wave (d > 0) that touches an enemy; key and door block
        const uint32_t pass = type_mask<0xE5u>(b, rm);
        uint32_t f = player, visited = player;
        int d = 0, min_dist = W * H;
        while (true) {  // two waves per pair of votes
          const uint32_t n1 = dilate(f) & pass & ~visited;
          const uint32_t v1 = visited | n1;
          const uint32_t n2 = dilate(n1) & pass & ~v1;
          if (__any_sync(FULL_MASK, ((n1 | n2) & enemies) != 0u)) {
            min_dist = __any_sync(FULL_MASK, (n1 & enemies) != 0u) ? d + 1 : d + 2;
            break;
          }
          if (!__any_sync(FULL_MASK, n2 != 0u)) break;  // frontier ran out without touching an enemy
          visited = v1 | n2;
          f = n2;
          d += 2;
        }
        st[5] = min_dist;
      }
      if (st[1] == 1 && st[2] == 1) {  // player -> key (door blocks), key -> door; either leg may be -1
        st[6] += bfs_dist_to(player, key, type_mask<0xEDu>(b, rm));
        st[6] += bfs_dist_to(key, door, type_mask<0xFDu>(b, rm));
      }
    }
  } else if (PROB == PCGRL_PROB_SOKOBAN) {  // sokoban_prob.py:133-145
    st[0] = popc_all(type_mask<0x04u>(b, rm));
    st[1] = popc_all(type_mask<0x08u>(b, rm));
    st[2] = popc_all(type_mask<0x10u>(b, rm));
    st[3] = count_regions(type_mask<0x1Du>(b, rm), lane);
    st[4] = W * H * (W + H);
    st[5] = 0;
    need_solver = (st[0] == 1 && st[1] == st[2] && st[1] > 0 && st[3] == 1);
  } else if (PROB == PCGRL_PROB_DDAVE) {  // ddave_prob.py:149-169
    st[0] = popc_all(type_mask<0x04u>(b, rm));
    st[1] = floor_dist(type_mask<0x04u>(b, rm), type_mask<0x02u>(b, rm), H, lane);
    st[2] = popc_all(type_mask<0x08u>(b, rm));
    st[3] = popc_all(type_mask<0x10u>(b, rm));
    st[4] = popc_all(type_mask<0x20u>(b, rm));
    st[5] = popc_all(type_mask<0x40u>(b, rm));
    st[6] = count_regions(type_mask<0x3Du>(b, rm), lane);  // empty, player, diamond, key, exit
    st[9] = W * H;
    need_solver = (st[0] == 1 && st[2] == 1 && st[4] == 1 && st[6] == 1);
  } else {  // mdungeon_prob.py:151-171
    st[0] = popc_all(type_mask<0x04u>(b, rm));
    st[1] = popc_all(type_mask<0x08u>(b, rm));
    st[2] = popc_all(type_mask<0x10u>(b, rm));
    st[3] = popc_all(type_mask<0x20u>(b, rm));
    st[4] = popc_all(type_mask<0xC0u>(b, rm));
    st[5] = count_regions(type_mask<0xFDu>(b, rm), lane);
    st[9] = W * H;
    need_solver = (st[0] == 1 && st[1] == 1 && st[5] == 1);
  }
}

// Out-of-line copy of map_stats for the cold paths (reset): keeps the kernels small; the hot step path inlines
// map_stats so that its statistics stay in registers.
template <int PROB>
__device__ __noinline__ void map_stats_shared(const Board& b, const pcgrl_config& cfg, int lane, int* st, bool& need_solver) {
  map_stats<PROB>(b, cfg, lane, st, need_solver);
}

// Problem.get_reward: fp64, terms summed left to right exactly as the reference writes them
// (compiled with -fmad=false so no product is fused into the adds).
template <int PROB>
__host__ __device__ __forceinline__ double problem_reward(const pcgrl_config& cfg, const int* n, const int* o) {
  const double* w = cfg.reward_weight;
  const int32_t* ip = cfg.iparam;
  const double INF = HUGE_VAL;
  if (PROB == PCGRL_PROB_BINARY) {  // binary_prob.py:98-106
    // get_range_reward on integers is exact in int32: (regions, 1, 1) follows helper.py:366-376 case by case and
    // (path, inf, inf) always takes the second case, min(new, inf) - min(old, inf) = new - old; the two products and
    // the sum are then the same fp64 operations as in the reference.
    const int nv = n[0], ov = o[0];
    int r0;
    if (nv == 1 && ov == 1) r0 = 0;
    else if (ov <= 1 && nv <= 1) r0 = min(nv, 1) - min(ov, 1);
    else if (ov >= 1 && nv >= 1) r0 = max(ov, 1) - max(nv, 1);
    else if (nv > 1 && ov < 1) r0 = 1 - nv + ov - 1;
    else r0 = 1 - ov + nv - 1;  // nv < 1 && ov > 1
    return (double)r0 * w[0] + (double)(n[1] - o[1]) * w[1];
  }
  if (PROB == PCGRL_PROB_ZELDA)  // zelda_prob.py:124-142
    // The statistics and the finite bounds are integers, so each get_range_reward term is an exact small integer: it is
    // computed in int32 (the fp64 min / max / compare chain was 30 % of the zelda kernel's instructions) and converted
    // before the same fp64 products and left-to-right sum.  (inf, inf) always takes the second branch: new - old.
    return (double)range_reward_i(n[0], o[0], 1, 1) * w[0] + (double)range_reward_i(n[1], o[1], 1, 1) * w[1] +
           (double)range_reward_i(n[2], o[2], 1, 1) * w[2] + (double)range_reward_i(n[3], o[3], 2, ip[0]) * w[3] +
           (double)range_reward_i(n[4], o[4], 1, 1) * w[4] + (double)range_reward_i_hi_inf(n[5], o[5], ip[1]) * w[5] +
           (double)(n[6] - o[6]) * w[6];
  if (PROB == PCGRL_PROB_SOKOBAN)  // sokoban_prob.py:157-175
    return range_reward(n[0], o[0], 1, 1) * w[0] + range_reward(n[1], o[1], 1, ip[0]) * w[1] +
           range_reward(n[2], o[2], 1, ip[0]) * w[2] + range_reward(n[3], o[3], 1, 1) * w[3] +
           range_reward(abs(n[1] - n[2]), abs(o[1] - o[2]), -INF, -INF) * w[4] +
           range_reward(n[4], o[4], -INF, -INF) * w[5] + range_reward(n[5], o[5], INF, INF) * w[6];
  if (PROB == PCGRL_PROB_DDAVE)  // ddave_prob.py:181-205
    return range_reward(n[0], o[0], 1, 1) * w[0] + range_reward(n[1], o[1], 0, 0) * w[1] +
           range_reward(n[2], o[2], 1, 1) * w[2] + range_reward(n[5], o[5], ip[1], INF) * w[3] +
           range_reward(n[3], o[3], -INF, ip[0]) * w[4] + range_reward(n[4], o[4], 1, 1) * w[5] +
           range_reward(n[6], o[6], 1, 1) * w[6] + range_reward(n[7], o[7], INF, INF) * w[7] +
           range_reward(n[9], o[9], -INF, -INF) * w[8] + range_reward(n[10], o[10], INF, INF) * w[9];
  // mdungeon_prob.py:183-205
  return range_reward(n[0], o[0], 1, 1) * w[0] + range_reward(n[1], o[1], 1, 1) * w[1] +
         range_reward(n[4], o[4], 1, ip[0]) * w[2] + range_reward(n[3], o[3], -INF, ip[2]) * w[3] +
         range_reward(n[2], o[2], -INF, ip[1]) * w[4] + range_reward(n[5], o[5], 1, 1) * w[5] +
         range_reward(n[8], o[8], INF, INF) * w[6] + range_reward(n[9], o[9], -INF, -INF) * w[7] +
         range_reward(n[10], o[10], INF, INF) * w[8];
}

// Problem.get_episode_over
template <int PROB>
__host__ __device__ __forceinline__ bool problem_over(const pcgrl_config& cfg, const int* n, const int* start) {
  const int32_t* ip = cfg.iparam;
  if (PROB == PCGRL_PROB_BINARY) return n[0] == 1 && n[1] - start[1] >= ip[0];   // binary_prob.py:119-120
  if (PROB == PCGRL_PROB_ZELDA) return n[5] >= ip[1] && n[6] >= ip[2];            // zelda_prob.py:155-156
  if (PROB == PCGRL_PROB_SOKOBAN) return n[5] >= ip[1];                           // sokoban_prob.py:188-189
  if (PROB == PCGRL_PROB_DDAVE) return n[10] >= ip[3] && n[7] > ip[2];            // ddave_prob.py:218-220
  return n[10] >= ip[3] && n[4] > 0 &&                                            // mdungeon_prob.py:218-221
         (double)n[8] / (double)(n[4] > 1 ? n[4] : 1) > cfg.dparam[0];
}

}  // namespace pcgrl

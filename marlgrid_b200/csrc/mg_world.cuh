// mg_world.cuh -- a whole warp generates ONE fresh world: MultiGridEnv.reset (base.py:402-416) + _gen_grid (empty.py:9-16,
// cluttered.py:25-36, goalcycle.py:30-51) with lanes = consecutive placement tries of place_obj's rejection sampling
// (base.py:690-708).  Shared by the fused step kernel (mg_fused2.cuh: regenerating a finished env inside the step) and by the
// background world generator (mg_pregen.cu: the same worlds, produced ahead of time off the step's critical path).
#pragma once
#include "mg_common.cuh"

namespace mg {
namespace world {

// Pre-generated worlds (`MgState.pregen`, uint32 [B][PG_WORDS]): everything of an env's NEXT episode that does not depend on
// its trajectory -- the bit-plane words of the fresh world in natural order, the agents' spawn cells -- tagged with the seed
// and the episode number the draws were keyed with.  A fresh world is a pure function of (seed, global env index, episode),
// DESIGN.md "RNG contract".
constexpr int PG_WORDS = 64;
constexpr int PG_XY0 = BITS_WORDS;        // words 44 .. 44 + MG_MAX_AGENTS - 1: agent q's cell, x | y << 8
constexpr int PG_SEED_LO = 61, PG_SEED_HI = 62;
constexpr int PG_TAG = 63;                // episode number the world is for + 1 (0: none; bit 31: generation gave up)

__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(uint32_t* p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// The on-device policy of mg_rollout_policy: action = argmax_k (bias[k] + sum_i w[k][i] * obs[i]) over the agent's encoded
// observation (uint8) with int8 weights -- exact integer arithmetic (dp4a), lowest k wins ties -- or, with probability
// eps / 2^32, a uniform action from the Philox block (g, lifetime step, TAG_POLICY | agent) keyed by the policy's seed.
// `word(i)` returns observation bytes 4i .. 4i+3 (bytes beyond the observation must read as anything: their weights are 0).
template <int NW, class WordFn>
__device__ __forceinline__ int linear_policy_action(const KP& p, int a, unsigned long long g, uint32_t t_life, WordFn word) {
  int acc[8];
  const int4* const bias = reinterpret_cast<const int4*>(p.pol_b + a * 8);
  const int4 b0 = __ldg(bias), b1 = __ldg(bias + 1);
  acc[0] = b0.x; acc[1] = b0.y; acc[2] = b0.z; acc[3] = b0.w; acc[4] = b1.x; acc[5] = b1.y; acc[6] = b1.z; acc[7] = b1.w;
  const int4* const w = reinterpret_cast<const int4*>(p.pol_w + (size_t)a * NW * 8);
#pragma unroll 4
  for (int i = 0; i < NW; ++i) {
    const uint32_t o = word(i);
    const int4 w0 = __ldg(w + 2 * i), w1 = __ldg(w + 2 * i + 1);  // the same address in every lane of the warp (warp = agent)
    const int ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
    for (int k = 0; k < 8; ++k) asm("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(acc[k]) : "r"(o), "r"(ww[k]));
  }
  int best = 0;
#pragma unroll
  for (int k = 1; k < 8; ++k)
    if (k < p.pol_n && acc[k] > acc[best]) best = k;
  if (p.pol_eps != 0u) {
    const U4 r = philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), t_life, TAG_POLICY | (uint32_t)a, (uint32_t)p.pol_seed, (uint32_t)(p.pol_seed >> 32));
    if (r.x < p.pol_eps) best = (int)__umulhi(r.y, (uint32_t)p.pol_n);
  }
  return best;
}

// Transpose of two 16x16 bit matrices at once: lane r < 16 holds row r of matrix 0 in bits 0..15 and row r of matrix 1 in
// bits 16..31; returns, in lane c < 16, column c of both the same way.  Four butterfly stages (block swaps of 8, 4, 2, 1).
__device__ __forceinline__ uint32_t transpose16x16_pair(uint32_t v, int lane) {
#pragma unroll
  for (int j = 8; j >= 1; j >>= 1) {
    const uint32_t mask = j == 8 ? 0x00FF00FFu : j == 4 ? 0x0F0F0F0Fu : j == 2 ? 0x33333333u : 0x55555555u;
    const uint32_t other = __shfl_xor_sync(0xFFFFFFFFu, v, j);
    v = (lane & j) ? (((other >> j) & mask) | (v & ~mask)) : ((v & mask) | ((other & mask) << j));
  }
  return v;
}

// The placement of one episode by a whole warp.  `wk` = 36 words of scratch owned by the warp (wall_x[16]: bit y = canonical
// wall at (x, y); other_x[16]: Goal / BonusTile; list[4]: object list entries), left filled on success.
//   static objects (random goal, bonus tiles, clutter walls, in this order): scanning the tries in order, a try is accepted
//     iff its cell is free of walls / objects AND no earlier accepted try hit the same cell (the object placed there is what
//     the sequential code would find) -- within a batch of 32 tries that is "first valid lane of its cell" (match_any); the
//     j-th accepted try gets the j-th object.
//   agents (ghost mode: they may share cells): the tries after the last static object's, each non-wall try places the next
//     agent; lane q < A returns agent q's cell in a_xy (x | y << 8).
// Returns false when the run is not an ordinary one (a whole batch of 32 tries without a placement -- the only way max_tries,
// base.py:700-706, could come into play --, or more than MAXB batches): the caller then leaves the env to the sequential code.
template <int A>
__device__ __forceinline__ bool warp_sample(const KP& p, unsigned long long g, uint32_t ep, uint32_t* __restrict__ wk, int lane, uint32_t& a_xy) {
  constexpr int MAXB = 8;
  const int W = p.W, H = p.H;
  uint32_t* wall_x = wk;
  uint32_t* other_x = wk + 16;
  uint32_t* list = wk + 32;
  if (lane < 16) {
    const uint32_t fullr = (1u << H) - 1u, endsr = 1u | (1u << (H - 1));
    wall_x[lane] = (lane == 0 || lane == W - 1) ? fullr : (lane < W ? endsr : 0u);  // wall_rect base.py:172-176
    other_x[lane] = (p.goal_mode == MG_GOAL_FIXED && lane == W - 2) ? (1u << (H - 2)) : 0u;  // put_obj(Goal) base.py:655-662
  }
  if (lane < OBJ_SLOTS) list[lane] = (lane == 0 && p.goal_mode == MG_GOAL_FIXED) ? obj_entry(W - 2, H - 2, MG_T_GOAL, MG_C_GREEN, 0) : 0u;
  __syncwarp();
  const int n_goal = (p.goal_mode == MG_GOAL_RANDOM) ? 1 : 0, n_other = n_goal + p.n_bonus, n_static = n_other + p.n_clutter;
  const int list_base = (p.goal_mode == MG_GOAL_FIXED) ? 1 : 0;
  const uint32_t lt = (1u << lane) - 1u;
  int placed_static = 0, agents_done = 0;
  a_xy = 0;
  for (int batch = 0; batch < MAXB; ++batch) {
    const uint32_t k = (uint32_t)(batch * 32 + lane);
    const U4 r = philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), ep, TAG_RESET | (k >> 1), (uint32_t)p.seed, (uint32_t)(p.seed >> 32));
    const int x = (int)__umulhi((k & 1u) ? r.z : r.x, (uint32_t)W), y = (int)__umulhi((k & 1u) ? r.w : r.y, (uint32_t)H);
    int start_lane = 0;
    if (placed_static < n_static) {
      const bool valid = !(((wall_x[x] | other_x[x]) >> y) & 1u);
      const uint32_t vm = __ballot_sync(0xFFFFFFFFu, valid);
      bool acc = false;
      if (valid) acc = (__ffs(__match_any_sync(vm, x * 16 + y)) - 1) == lane;  // first valid try of this cell in the batch
      const uint32_t am = __ballot_sync(0xFFFFFFFFu, acc);
      if (am == 0u) return false;
      const int j = placed_static + __popc(am & lt);  // index of the object this try would place
      const bool take = acc && j < n_static;
      const uint32_t tm = __ballot_sync(0xFFFFFFFFu, take);
      if (take) {
        if (j < n_other) {
          atomicOr(&other_x[x], 1u << y);
          const uint32_t e = (j < n_goal) ? obj_entry(x, y, MG_T_GOAL, MG_C_GREEN, 0) : obj_entry(x, y, MG_T_BONUS, MG_C_YELLOW, j - n_goal);
          if (list_base + j < OBJ_SLOTS) list[list_base + j] = e;
        } else atomicOr(&wall_x[x], 1u << y);
      }
      placed_static += __popc(tm);
      __syncwarp();
      if (placed_static < n_static) continue;
      start_lane = 32 - __clz(tm);  // the agents' tries begin behind the last static object's
    }
    const bool valid_a = lane >= start_lane && !((wall_x[x] >> y) & 1u);
    const uint32_t vma = __ballot_sync(0xFFFFFFFFu, valid_a);
    if (vma == 0u) { if (start_lane == 0) return false; else continue; }
    const int q = agents_done + __popc(vma & lt);
    const uint32_t xy = (uint32_t)x | ((uint32_t)y << 8);
#pragma unroll
    for (int t = 0; t < A; ++t) {  // hand try "q == t" to lane t
      const uint32_t src = __ballot_sync(0xFFFFFFFFu, valid_a && q == t);
      if (src) { const uint32_t v = __shfl_sync(0xFFFFFFFFu, xy, __ffs(src) - 1); if (lane == t) a_xy = v; }
    }
    agents_done += __popc(vma);
    if (agents_done >= A) { __syncwarp(); return true; }
  }
  return false;
}

// The bit-plane words of the world warp_sample() left in `wk`, written with word stride STRIDE (BS: a tile's transposed chunk
// in shared memory; 1: a pre-generated world slot): x-lines as sampled, y-lines by transposition, object list, zero guards.
template <int STRIDE>
__device__ __forceinline__ void commit_lines(uint32_t* __restrict__ bits, const uint32_t* __restrict__ wk, int lane) {
  const uint32_t* wall_x = wk;
  const uint32_t* other_x = wk + 16;
  const uint32_t* list = wk + 32;
  // lane x < 16 holds its x-line (walls in the low half, Goal / BonusTiles in the high half); the y-lines are the transposed
  // 16x16 bit matrices: four butterfly stages on both halves at once
  const uint32_t xl = lane < 16 ? (wall_x[lane] | (other_x[lane] << 16)) : 0u;
  const uint32_t yl = transpose16x16_pair(xl, lane);
  if (lane < 16) {
    bits[(LINE_X0 + lane) * STRIDE] = xl;
    bits[(LINE_Y0 + lane) * STRIDE] = yl;
  }
  if (lane < 4) bits[(OBJ_WORD0 + lane) * STRIDE] = list[lane];
  if (lane >= 4 && lane < 8) bits[(OBJ_WORD0 + lane) * STRIDE] = 0u;             // words 40..43
  if (lane >= 8 && lane < 12) bits[((lane == 8) ? 0 : (lane == 9) ? 17 : (lane == 10) ? 18 : 35) * STRIDE] = 0u;  // guard lines
}

}  // namespace world
}  // namespace mg

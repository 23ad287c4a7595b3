// mg_fused2.cuh -- the hot path: env.step + auto-reset + egocentric observation (encoded or RGB) in ONE launch, specialised
// at compile time on the agent count A and view size V of the registered env shapes (MarlGrid-*: V = 7 or 5, A <= 6).
//
//   grid = persistent CTAs (as many as stay resident, trimmed to equal rounds); CTA c handles tiles c, c + gridDim.x, ...
//   tile = 32 consecutive envs, A warps:  lane = env, warp = agent  (thread = one agent = one view; no divisions, per-env
//          work is simply "warp 0", and every shared-memory access pattern is a fixed stride across lanes)
//   I/O  = bulk-async copies only (cp.async.bulk + mbarrier -- the TMA engine, SASS UBLKCP): the tile's bit-plane lines,
//          agent records, env records and actions come in as four contiguous chunks (into one of NST input stages: the next
//          tile is prefetched while this one is processed); observations, records, env records, rewards and done flags
//          leave the same way.  No thread touches global memory on the common path.
//   step = MultiGridEnv.step (base.py:501-649): in ghost mode an action that does not edit the planes is independent of
//          what the env's other agents do in the same step, so the A agents act in parallel; the reference's random
//          processing order (base.py:514-516, one Philox block per env) only ranks the arrival stamps of the movers.
//          Envs in which an action WOULD edit the planes (effective pickup / drop / toggle) are replayed with the
//          sequential code of mg_env.cuh; finished envs are regenerated the same way (env_reset), their byte planes
//          rebuilt from the bit-plane lines in shared memory and stored by bulk copies.
//   obs  = gen_obs_grid (base.py:418-451) + occlude_mask (agents.py:298-343), then
//          OBS 1: MultiGrid.encode (base.py:196-214).  A view row is ONE word load (line of the bit-planes: opaque |
//                 other<<16) + a window shift; the reference's rotation is a row-order flip and a bit reversal of the line;
//                 line of sight is carry propagation on 7-bit rows; visible canonical walls are written as constants, the
//                 few other objects come from the env's object list, agents from the records.
//          OBS 2: MultiGrid.render (base.py:301-331) at tile size 8.  The view threads write tile ids; the warps then copy
//                 tile rows from the atlas (shared memory) into chunk buffers that bulk copies stream to HBM.
// Everything here is integer work on the ALU/LSU pipes; the path has no dense contraction, hence no tensor cores.
// Launch-shape knobs for experiments: MG_F2_CTAS_PER_SM=n, MG_F2_PDL=0, MG_F2_RAGGED=1, MG_F2_VERBOSE=1.
#pragma once
#include <climits>
#include <cstdlib>

#include "mg_env.cuh"
#include "mg_world.cuh"

namespace mg {

namespace f2 {

constexpr uint32_t FL_SLOW = 1u, FL_RESET = 2u, FL_BITS_DIRTY = 4u, FL_NOTDONE = 8u, FL_IMAGE = 16u;  // s_flag bits; bits 8..15 movers, 16..31 error bits

// the general sequential code, kept out of line so the common path keeps its registers
static __device__ __noinline__ void seq_step(EnvCtx<32>& cref, unsigned long long g, const int32_t* act, double* rew) {
  EnvCtx<32> c = cref;
  env_step<32, true, MG_MAX_AGENTS>(c, g, act, rew);
  cref.sc = c.sc; cref.ep = c.ep; cref.tl = c.tl; cref.w3 = c.w3; cref.dirty = c.dirty;
}
template <bool PLANES>
__device__ __noinline__ void seq_reset(EnvCtx<32>& cref, unsigned long long g) {
  EnvCtx<32> c = cref;
  env_reset<32, true, PLANES>(c, g);
  cref.sc = c.sc; cref.ep = c.ep; cref.tl = c.tl; cref.w3 = c.w3; cref.dirty = c.dirty;
}

// shared memory of one persistent CTA: NST input stages (what the bulk loads fill and the state stores drain), one output
// tile, the small per-tile exchange arrays.  Offsets in bytes, every block a multiple of 16.
template <int OBS, int V, int A, int NST>
struct Smem {
  static constexpr int E = ENVS_PER_CTA;
  // OBS 1: the encoded observation tile.  OBS 2: per-view tile-id maps (V rows of 8 bytes) + per warp NBUF image chunks
  // (one row of V cells = 8 pixel rows of V*24 bytes) that bulk copies drain while the next chunk is being built.
  static constexpr int MAP_BYTES = E * A * V * 8;
  static constexpr int CHUNK = V * 8 * 24;
  static constexpr int NBUF = 6;  // OBS 2: chunk buffers per warp (NBUF - 1 bulk copies in flight while one is being filled)
  static constexpr int OUT_BYTES = OBS == 1 ? E * A * V * V * 3 : MAP_BYTES + A * NBUF * CHUNK;
  static constexpr int SCRATCH_BYTES = (A * 4 * 32 + 64 * 32) * 4;  // sequential path: transposed records + reset masks
  static constexpr int OUT_AREA = ((OUT_BYTES > SCRATCH_BYTES ? OUT_BYTES : SCRATCH_BYTES) + 15) / 16 * 16;
  // one input stage
  static constexpr int ST_BITS = 0;
  static constexpr int ST_REC = ST_BITS + E * BITS_WORDS * 4;
  static constexpr int ST_ENV = ST_REC + E * A * 16;
  static constexpr int ST_ACT = ST_ENV + E * 16;
  static constexpr int STAGE = ST_ACT + (E * A * 4 + 15) / 16 * 16;
  // after the stages
  static constexpr int REW = NST * STAGE;
  static constexpr int DONE = REW + E * A * 8;
  static constexpr int FLAG = DONE + E;          // two copies (tile parity)
  static constexpr int ORDER = FLAG + 2 * E * 4;
  static constexpr int BAR = ORDER + E * 4;      // NST mbarriers
  static constexpr int OUT = BAR + ((NST * 8 + 15) / 16) * 16;
  static constexpr int TOTAL = OUT + OUT_AREA;   // OBS 2: the tile atlas (run-time size) follows
};

// Lehmer / Fisher-Yates decode of the permutation number (oracle/philox.py shuffle_perm): nibble q of the result = agent
// processed q-th.  A is a compile-time constant: the divisions are multiplications.
template <int A>
__device__ __forceinline__ uint32_t decode_order_ct(uint32_t pidx) {
  uint32_t order = 0x76543210u;
#pragma unroll
  for (int i = A - 1; i >= 1; --i) {
    const uint32_t n = (uint32_t)(i + 1);
    const uint32_t qd = pidx / n;
    const uint32_t j = pidx - qd * n;
    pidx = qd;
    const uint32_t ni = (order >> (4 * i)) & 0xFu, nj = (order >> (4 * j)) & 0xFu;
    order = (order & ~(0xFu << (4 * i)) & ~(0xFu << (4 * j))) | (nj << (4 * i)) | (ni << (4 * j));
  }
  return order;
}

template <int N>
struct Fact { static constexpr uint32_t v = N * Fact<N - 1>::v; };
template <>
struct Fact<0> { static constexpr uint32_t v = 1; };

// occlude_mask (agents.py:298-343) for the agent at (V/2, V-1), i.e. view_offset 0: the upward pass visits every row;
// the downward pass then only re-sweeps the agent's own row V-1, which the upward pass has already closed under both
// sweeps (and row V does not exist), so it changes nothing and is skipped.
template <int V>
__device__ __forceinline__ void occlude_rows_vo0(const uint32_t (&T)[V], uint32_t (&M)[V]) {
  constexpr uint32_t RM = (1u << V) - 1u;
  constexpr int ax = V / 2;
  constexpr uint32_t ge_ax = RM & ~((1u << ax) - 1u);
  constexpr uint32_t left_src = ((1u << (ax + 2)) - 1u) & ~1u & RM;
#pragma unroll
  for (int j = 0; j < V; ++j) M[j] = (j == V - 1) ? (1u << ax) : 0u;
#pragma unroll
  for (int j = V - 1; j >= 1; --j) {
    uint32_t nxt = 0;
    sweep_row<V>(M[j], T[j], nxt, ge_ax, left_src);
    M[j - 1] |= nxt;
  }
  // (row 0 is only ever a sweep target: the reference's loop runs j = ay+1 .. 1)
}

// rows (one per register, V bits each) -> one byte per row of a 64-bit word: three byte permutes per four rows.  Bits of
// a row above bit 7 are dropped; bit 7 itself is kept (callers mask with a clean row set).
template <int V>
__device__ __forceinline__ uint64_t pack_rows8(const uint32_t (&r)[V]) {
  auto pack4 = [](uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3) {
    return __byte_perm(__byte_perm(r0, r1, 0x0040), __byte_perm(r2, r3, 0x0040), 0x5410);
  };
  const uint32_t lo = pack4(r[0], V > 1 ? r[1] : 0u, V > 2 ? r[2] : 0u, V > 3 ? r[3] : 0u);
  const uint32_t hi = V > 4 ? pack4(r[4 < V ? 4 : 0], V > 5 ? r[5 < V ? 5 : 0] : 0u, V > 6 ? r[6 < V ? 6 : 0] : 0u, V > 7 ? r[7 < V ? 7 : 0] : 0u) : 0u;
  return ((uint64_t)hi << 32) | lo;
}

// one-byte shared-memory store at a compile-time offset from a 32-bit shared address
__device__ __forceinline__ void sts_u8(uint32_t saddr, uint32_t v) {
  asm volatile("st.shared.u8 [%0], %1;" ::"r"(saddr), "r"(v) : "memory");
}

constexpr int WARP_RESET_MAX = 4;  // finished envs per warp up to which the warps regenerate them one by one

// Warp-cooperative reset, for tiles in which only a few envs finish -- the normal state of a long-running batch, whose episodes
// have drifted apart (tools/desync_probe.py): no table, no CTA-wide barriers, one env at a time per warp.  The placement itself
// is world::warp_sample (mg_world.cuh: lanes = consecutive placement tries).  Returns false WITHOUT having committed anything when
// the run is not an ordinary one: the caller then runs the sequential code.  On success bits / rec / envr (shared memory) hold
// the new episode.
template <int A>
__device__ __forceinline__ void commit_episode(uint32_t* __restrict__ rec, int32_t* __restrict__ envr, uint32_t a_xy, int lane) {
  if (lane < A) {  // agents.py:161-170 (dir survives), placement order = stamp order (base.py:409-412,686)
    const uint32_t old = rec[lane * 4];
    *reinterpret_cast<uint4*>(rec + lane * 4) = make_uint4((old & 0x00FF0000u) | a_xy | ((uint32_t)(MG_AF_PLACED | MG_AF_ACTIVE) << 24), 0xFF000000u, (uint32_t)lane, 0u);
  }
  if (lane == 0) {
    envr[0] = 0; envr[1] += 1;
    envr[3] = (int)(((uint32_t)envr[3] & 0xFFFF0000u) | (uint32_t)A);
  }
}
template <int A>
__device__ __forceinline__ bool warp_reset(const KP& p, unsigned long long g, uint32_t* __restrict__ bits, uint32_t* __restrict__ rec,
                                        int32_t* __restrict__ envr, uint32_t* __restrict__ wk /* 36 words of this warp */, int lane) {
  uint32_t a_xy;
  if (!world::warp_sample<A>(p, g, (uint32_t)envr[1], wk, lane, a_xy)) return false;  // (ends with __syncwarp: every lane has read envr[1])
  commit_episode<A>(rec, envr, a_xy, lane);
  world::commit_lines<BS>(bits, wk, lane);
  __syncwarp();
  return true;
}

// Pre-generated worlds (mg_pregen.cu, mg_world.cuh): the finished envs of `mine` (bit = env of the tile; this warp's share)
// whose slot holds the world of exactly their next episode are regenerated by copying it -- 46 words per env, the loads of up
// to four envs in flight together.  Two dependent round trips to L2: the tags (acquire: the generator puts a tag up, with
// release, after the slot is complete, and never touches a slot whose tag is valid for the env's current episode), then the
// words.  Returns the envs done; the others take the generating routes.
template <int A>
__device__ __forceinline__ uint32_t consume_pregen(const KP& p, uint32_t mine, long long env0, uint32_t* __restrict__ s_bits,
                                                   uint32_t* __restrict__ s_rec, int32_t* __restrict__ s_env, int lane) {
  using namespace world;
  if (p.pregen == nullptr || mine == 0u) return 0u;
  const uint32_t* const pg = p.pregen + env0 * PG_WORDS;
  bool ok = false;
  if ((mine >> lane) & 1u) ok = ld_acquire_gpu(pg + lane * PG_WORDS + PG_TAG) == (uint32_t)s_env[lane * 4 + 1] + 1u;  // lane == env
  uint32_t m = __ballot_sync(0xFFFFFFFFu, ok), done = 0u;
  while (m != 0u) {
    int e[4];
    uint32_t v0[4], v1[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { e[j] = __ffs(m) - 1; m &= m - 1u; }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v0[j] = 0u; v1[j] = 0u;
      if (e[j] >= 0) {
        v0[j] = __ldcg(pg + e[j] * PG_WORDS + lane);                                                    // words 0..31
        v1[j] = __ldcg(pg + e[j] * PG_WORDS + (lane < PG_XY0 - 32 + A ? 32 + lane : PG_SEED_LO + (lane & 1)));  // words 32..43, the agents' cells; else the seed
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (e[j] < 0) continue;
      // the world was drawn with this family's seed (lanes 30 / 31 hold the slot's seed words)
      const uint32_t slo = __shfl_sync(0xFFFFFFFFu, v1[j], 30), shi = __shfl_sync(0xFFFFFFFFu, v1[j], 31);
      if (slo != (uint32_t)p.seed || shi != (uint32_t)(p.seed >> 32)) continue;
      done |= 1u << e[j];
      uint32_t* const bits = s_bits + e[j];
      bits[lane * BS] = v0[j];
      if (lane < BITS_WORDS - 32) bits[(32 + lane) * BS] = v1[j];
      const uint32_t a_xy = __shfl_sync(0xFFFFFFFFu, v1[j], (PG_XY0 - 32 + lane) & 31);  // lane q < A: agent q's cell
      commit_episode<A>(s_rec + e[j] * (A * 4), s_env + e[j] * 4, a_xy, lane);
    }
  }
  __syncwarp();
  return done;
}

// The byte planes of freshly regenerated envs, written straight to global memory: ONE ENV PER HALF-WARP, lane hl = lane & 15
// builds cells [16 hl, 16 hl + 16) of all three planes from the env's bit-plane lines in shared memory (a fresh world holds
// canonical walls = OP & ~OT -> type 8 / colour 9, and the Goal / BonusTiles of the object list) and stores three 16-byte
// vectors -- coalesced plain stores, nothing to wait for, no staging image.  W, H <= 16: a plane has at most 16 chunks.
struct PlaneLane { int c0, x0, sh0; bool on; };
__device__ __forceinline__ PlaneLane plane_lane(const KP& p, int lane) {  // the per-lane constants (one integer division)
  PlaneLane q;
  q.c0 = (lane & 15) << 4;
  q.on = q.c0 < p.S;
  q.x0 = q.c0 / p.H;
  q.sh0 = q.x0 * p.H - q.c0;  // <= 0: where x-line x0 starts relative to the chunk
  return q;
}
__device__ __forceinline__ void patch_byte(uint4& v, int rel, uint32_t b) {
  const uint32_t sh = 8u * (rel & 3), keep = ~(0xFFu << sh), put = b << sh;
  const int j = rel >> 2;
  if (j == 0) v.x = (v.x & keep) | put; else if (j == 1) v.y = (v.y & keep) | put; else if (j == 2) v.z = (v.z & keep) | put; else v.w = (v.w & keep) | put;
}
__device__ __forceinline__ void store_planes(const KP& p, uint8_t* __restrict__ tp, const uint32_t* __restrict__ bits, const PlaneLane& q) {
  if (!q.on) return;
  const int W = p.W, H = p.H, S = p.S;
  uint32_t m = 0;  // bit k: cell c0 + k holds a canonical wall
  for (int x = q.x0, sh = q.sh0; sh < 16 && x < W; ++x, sh += H) {
    const uint32_t l = bits[(LINE_X0 + x) * BS], wl = l & 0xFFFFu & ~(l >> 16);
    m |= (sh >= 0) ? (wl << sh) : (wl >> (-sh));
  }
  uint32_t w[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) w[j] = (((m >> (4 * j)) & 15u) * 0x00204081u) & 0x01010101u;
  uint4 vt = make_uint4(w[0] * MG_T_WALL, w[1] * MG_T_WALL, w[2] * MG_T_WALL, w[3] * MG_T_WALL);
  uint4 vc = make_uint4(w[0] * MG_C_WORST, w[1] * MG_C_WORST, w[2] * MG_C_WORST, w[3] * MG_C_WORST);
  uint4 vs = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
  for (int k = 0; k < OBJ_SLOTS; ++k) {  // Goal / BonusTiles (always listed: the caller checked image_planes)
    const uint32_t e = bits[(OBJ_WORD0 + k) * BS];
    const int rel = (int)(e & 15u) * H + (int)((e >> 4) & 15u) - q.c0;
    if ((e >> 31) && (unsigned)rel < 16u) {
      patch_byte(vt, rel, (e >> 8) & 15u); patch_byte(vc, rel, (e >> 12) & 15u); patch_byte(vs, rel, (e >> 16) & 255u);
    }
  }
  *reinterpret_cast<uint4*>(tp + q.c0) = vt;
  *reinterpret_cast<uint4*>(tp + S + q.c0) = vc;
  *reinterpret_cast<uint4*>(tp + 2 * S + q.c0) = vs;
}
// the envs of `mask` (bit = env of the tile), two per pass: the lower half-warp takes one, the upper half-warp the next
__device__ __forceinline__ void store_planes_of(const KP& p, uint32_t mask, long long env0, const uint32_t* __restrict__ s_bits, int lane) {
  if (mask == 0u) return;
  const PlaneLane q = plane_lane(p, lane);
  while (mask != 0u) {
    const int e1 = __ffs(mask) - 1;
    mask &= mask - 1u;
    const int e2 = __ffs(mask) - 1;  // -1: none
    mask &= mask - 1u;               // (0 & -1 stays 0)
    const int e = (lane < 16) ? e1 : e2;
    if (e >= 0) store_planes(p, p.grid + (env0 + e) * 3 * p.S, s_bits + e, q);
  }
}

// Fast route of the rare path: only a few envs of the tile finished their episode (a long-running batch whose episodes have
// drifted apart: one or two per tile and step) and no env needs the sequential step replay.  Env e goes to warp e % A, which
// regenerates it (warp_reset: lanes = placement tries), stores its byte planes itself and hands its scratch words -- borrowed
// from the already zeroed output tile -- back clean: no table, no plane image in shared memory, no bulk copy to wait for, and
// a single CTA-wide barrier (the caller's).  Returns true if some env was left for the general route (nothing of it committed).
template <int A>
__device__ __forceinline__ bool fast_resets(const KP& p, uint32_t todo /* this warp's envs still to generate */, uint32_t ok /* its envs copied from pre-generated worlds */,
                                            uint32_t* s_flag, uint32_t* s_bits, uint32_t* s_rec, int32_t* s_env, uint32_t* wk, long long env0, int lane) {
  bool failed = false, used = false;
  while (todo != 0u) {
    const int e = __ffs(todo) - 1;
    todo &= todo - 1u;
    used = true;
    if (warp_reset<A>(p, (unsigned long long)(p.env_offset + env0 + e), s_bits + e, s_rec + e * (A * 4), s_env + e * 4, wk, lane)) {
      ok |= 1u << e;
      if (lane == 0) s_flag[e] &= ~FL_RESET;  // FL_BITS_DIRTY stays: the env's bit-plane words go back to global memory
    } else failed = true;
    __syncwarp();
  }
  store_planes_of(p, ok, env0, s_bits, lane);
  if (used) {  // the scratch lives in the output tile: zero again
    wk[lane] = 0u;
    if (lane < 4) wk[32 + lane] = 0u;
  }
  return failed;
}

// Reset of the finished envs of a tile: MultiGridEnv.reset (base.py:402-416) + _gen_grid (empty.py:9-16, cluttered.py:25-36,
// goalcycle.py:30-51), split so that the expensive part is parallel and the sequential part is cheap.  Per chunk of NT tries:
//   1. all threads of the CTA draw the Philox blocks of (finished env, block) pairs and leave the tries as cell bytes
//      (x | y << 4) in a table [try][env];
//   2. warp 0, lane = env, replays place_obj's rejection sampling (base.py:690-708, one try per iteration) from the table on
//      the env's fresh bit-plane lines: a handful of shared-memory accesses per try instead of ten Philox rounds, and every
//      lane runs the same loop.
// Objects in the reference's order: [random goal], bonus tiles, clutter walls (need an empty cell), then the agents (ghost
// mode: any cell that is not a wall; stamps = placement order, base.py:409-412,686).  An env that is not done after MAXCH
// chunks, or whose tries fail NT times in a row (the only way max_tries, base.py:700-706, could come into play), keeps its
// FL_RESET flag and nothing of it has been committed: the sequential code below then redoes it from scratch.
template <int A>
__device__ __forceinline__ void tile_resets(const KP& p, uint32_t* s_flag, uint32_t* s_bits, uint32_t* s_rec, int32_t* s_env,
                                            uint32_t* scratch, long long env0, int n_valid, int tid) {
  constexpr int NT = 32, MAXCH = 4;
  const int lane = tid & 31, a = tid >> 5;
  const int W = p.W, H = p.H;
  uint8_t* const table0 = reinterpret_cast<uint8_t*>(scratch);             // two tables [NT][32 envs]: while warp 0 replays one chunk,
  uint8_t* const table1 = reinterpret_cast<uint8_t*>(scratch + NT * 32 / 4);  // the other warps draw the next one
  uint32_t* const s_pending = scratch + 2 * (NT * 32 / 4);
  uint32_t* const s_ep = s_pending + 4;                                    // [32] the episode numbers that key the draws (before any commit)
  uint8_t* const s_list = reinterpret_cast<uint8_t*>(s_ep + 32) + a * 32;  // per warp: [32] the pending envs, compacted
  // lane == env in every warp: the same word everywhere
  const uint32_t todo = __ballot_sync(0xFFFFFFFFu, lane < n_valid && (s_flag[lane] & (FL_SLOW | FL_RESET)) == FL_RESET);
  if (todo == 0u) return;
  if (a == 0) s_ep[lane] = (uint32_t)s_env[lane * 4 + 1];
  // the Philox blocks of (pending env, block) pairs of one chunk, by the threads [t0, t0 + nthr) of the CTA
  auto fill = [&](int chunk, uint32_t pend, uint8_t* table, int ti, int nthr) {
    if ((pend >> lane) & 1u) s_list[__popc(pend & ((1u << lane) - 1u))] = (uint8_t)lane;  // k-th pending env (this warp's copy)
    __syncwarp();
    const int n_items = __popc(pend) * (NT / 2);
    for (int i = ti; i < n_items; i += nthr) {
      const int e = (int)s_list[i / (NT / 2)], j = i % (NT / 2);
      const unsigned long long g = (unsigned long long)(p.env_offset + env0 + e);
      const U4 r = philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), s_ep[e], TAG_RESET | (uint32_t)(chunk * (NT / 2) + j), (uint32_t)p.seed, (uint32_t)(p.seed >> 32));
      table[(2 * j) * 32 + e] = (uint8_t)(__umulhi(r.x, (uint32_t)W) | (__umulhi(r.y, (uint32_t)H) << 4));
      table[(2 * j + 1) * 32 + e] = (uint8_t)(__umulhi(r.z, (uint32_t)W) | (__umulhi(r.w, (uint32_t)H) << 4));
    }
    __syncwarp();
  };
  const bool fixed = p.goal_mode == MG_GOAL_FIXED;
  // fresh bit-plane chunks: wall_rect (base.py:172-176), the fixed goal (put_obj base.py:655-662), an empty object list
  for (int i = tid; i < ENVS_PER_CTA * BITS_WORDS; i += 32 * A) {  // i = w * 32 + e (tile-transposed words)
    const int e = i & 31, w = i >> 5;
    if (!((todo >> e) & 1u)) continue;
    uint32_t v = 0u;
    if (w >= LINE_X0 && w < LINE_X0 + 16) {
      const int x = w - LINE_X0;
      v = (x == 0 || x == W - 1) ? (1u << H) - 1u : (x < W ? (1u | (1u << (H - 1))) : 0u);
      if (fixed && x == W - 2) v |= 0x10000u << (H - 2);
    } else if (w >= LINE_Y0 && w < LINE_Y0 + 16) {
      const int y = w - LINE_Y0;
      v = (y == 0 || y == H - 1) ? (1u << W) - 1u : (y < H ? (1u | (1u << (W - 1))) : 0u);
      if (fixed && y == H - 2) v |= 0x10000u << (W - 2);
    } else if (w == OBJ_WORD0 && fixed) v = obj_entry(W - 2, H - 2, MG_T_GOAL, MG_C_GREEN, 0);
    s_bits[i] = v;
  }
  const int n_goal = (p.goal_mode == MG_GOAL_RANDOM) ? 1 : 0, n_other = n_goal + p.n_bonus, n_static = n_other + p.n_clutter;
  const int n_obj = n_static + A;
  int obj = 0, fails = 0, n_listed = fixed ? 1 : 0;  // warp 0: the placement loop's state of env `lane`
  uint32_t axy[A];
#pragma unroll
  for (int q = 0; q < A; ++q) axy[q] = 0u;
  uint32_t pending = todo;
  __syncthreads();  // s_ep is there
  fill(0, todo, table0, tid, 32 * A);
  __syncthreads();  // also: the fresh lines above are complete
  for (int chunk = 0; chunk < MAXCH; ++chunk) {
    const uint8_t* const table = (chunk & 1) ? table1 : table0;
    uint8_t* const next_table = (chunk & 1) ? table0 : table1;
    // the other warps draw the next chunk for every env still pending now (a superset of what will be needed)
    if (A > 1 && a != 0 && chunk + 1 < MAXCH) fill(chunk + 1, pending, next_table, tid - 32, 32 * (A - 1));
    if (a == 0) {
      bool open = (pending >> lane) & 1u;
      if (open) {
        uint32_t* const bits = s_bits + lane;
        for (int t = 0; t < NT; ++t) {
          const uint32_t cell = table[t * 32 + lane];
          const int x = (int)(cell & 15u), y = (int)(cell >> 4);
          const uint32_t line = bits[(LINE_X0 + x) * BS];
          const uint32_t cb = (line >> y) & 0x10001u;  // bit 0 wall, bit 16 Goal / BonusTile
          const bool agent = obj >= n_static;
          if (agent ? !(cb & 1u) : cb == 0u) {
            if (!agent) {
              const uint32_t bit = (obj < n_other) ? 0x10000u : 1u;
              bits[(LINE_X0 + x) * BS] = line | (bit << y);
              bits[(LINE_Y0 + y) * BS] |= bit << x;
              if (obj < n_other) {
                const uint32_t e = (obj < n_goal) ? obj_entry(x, y, MG_T_GOAL, MG_C_GREEN, 0) : obj_entry(x, y, MG_T_BONUS, MG_C_YELLOW, obj - n_goal);
                if (n_listed < OBJ_SLOTS) bits[(OBJ_WORD0 + n_listed++) * BS] = e;
              }
            } else {
              const int q = obj - n_static;
#pragma unroll
              for (int k = 0; k < A; ++k)
                if (k == q) axy[k] = (uint32_t)x | ((uint32_t)y << 8);
            }
            fails = 0;
            if (++obj == n_obj) break;
          } else if (++fails >= NT) break;
        }
        if (obj == n_obj) {  // commit: records (agents.py:161-170, dir survives), env record, flags
          uint32_t* const rec = s_rec + lane * (A * 4);
#pragma unroll
          for (int q = 0; q < A; ++q) {
            const uint32_t old = rec[q * 4];
            *reinterpret_cast<uint4*>(rec + q * 4) =
                make_uint4((old & 0x00FF0000u) | axy[q] | ((uint32_t)(MG_AF_PLACED | MG_AF_ACTIVE) << 24), 0xFF000000u, (uint32_t)q, 0u);
          }
          int32_t* const envr = s_env + lane * 4;
          envr[0] = 0; envr[1] += 1;
          envr[3] = (int)(((uint32_t)envr[3] & 0xFFFF0000u) | (uint32_t)A);
          s_flag[lane] = (s_flag[lane] & ~FL_RESET) | FL_IMAGE;
          open = false;
        } else if (fails >= NT) open = false;  // left to the sequential code
      }
      const uint32_t still = __ballot_sync(0xFFFFFFFFu, open);
      if (lane == 0) s_pending[chunk & 1] = still;  // two slots: a warp may read this chunk's word while warp 0 is already replaying the next chunk
      if (A == 1 && still != 0u && chunk + 1 < MAXCH) fill(chunk + 1, still, next_table, tid, 32);  // single-warp CTAs: no one to overlap with
    }
    __syncthreads();
    pending = s_pending[chunk & 1];
    if (pending == 0u) break;
  }
  __syncthreads();  // the table and s_pending alias the sequential code's scratch: nobody may still be reading them
}

template <int OBS, int A, class SM>
__device__ __forceinline__ void zero_out_tile(uint8_t* s_out, int tid) {  // invisible / empty cells encode as 0 (the RGB path writes every byte of its tile maps)
  if (OBS != 1) return;
  int4* z = reinterpret_cast<int4*>(s_out);
  constexpr int N16 = SM::OUT_BYTES / 16, ITERS = (N16 + 32 * A - 1) / (32 * A);
#pragma unroll
  for (int k = 0; k < ITERS; ++k) {
    const int i = tid + k * 32 * A;
    if (i < N16) z[i] = make_int4(0, 0, 0, 0);
  }
}

// The rare part of a tile's step (some env of the tile finished its episode, or an action edits the byte planes): the whole
// CTA comes here between the commit of the parallel envs and the observe phase.  Out of line and self-contained on purpose.
template <int OBS, int V, int A, int NST, bool KS>
__device__ __noinline__ void rare_path(const KP& p, const int tile, const int stage, const int it, const int step) {
  using SM = Smem<OBS, V, A, NST>;
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, a = tid >> 5;
  const int W = p.W, H = p.H, S = p.S;
  const bool image_planes = (p.goal_mode != MG_GOAL_NONE ? 1 : 0) + p.n_bonus <= OBJ_SLOTS;
  double* s_rew = reinterpret_cast<double*>(smem + SM::REW);
  uint8_t* s_done = smem + SM::DONE;
  uint8_t* s_out = smem + SM::OUT;
  uint32_t* s_trec = reinterpret_cast<uint32_t*>(s_out);  // sequential path scratch, aliased with the output tile
  uint32_t* s_scr = s_trec + A * 4 * 32;
  const int32_t* const act_g = p.actions + (KS ? (long long)step * p.B * A : 0);
  double* const rew_g = p.rewards + (KS ? (long long)step * p.B * A : 0);
  uint8_t* const done_g = p.done + (KS ? (long long)step * p.B : 0);
  unsigned char* const stg = smem + stage * SM::STAGE;
  uint32_t* const s_bits = reinterpret_cast<uint32_t*>(stg + SM::ST_BITS);
  uint32_t* const s_rec = reinterpret_cast<uint32_t*>(stg + SM::ST_REC);
  int32_t* const s_env = reinterpret_cast<int32_t*>(stg + SM::ST_ENV);
  uint32_t* const s_flag = reinterpret_cast<uint32_t*>(smem + SM::FLAG) + (it & 1) * 32;
  const long long env0 = (long long)tile * ENVS_PER_CTA;
  const int n_valid = (int)min((long long)ENVS_PER_CTA, p.B - env0);
  const bool full = n_valid == ENVS_PER_CTA;
  const bool mine = lane < n_valid;
  const long long env = env0 + lane;
  uint32_t* const rec = s_rec + lane * (A * 4);
  uint32_t* const bits = s_bits + lane;
  uint8_t* const tp = p.grid + env * 3 * S;
  {
    // lane == env in every warp: the envs are dealt out over the CTA's warps (env e goes to warp e % A), so that A warps
    // instead of one chew through the sequential code of mg_env.cuh; scratch lives in the (not yet used) output tile
    if (image_planes) {
      // lane == env in every warp: the same words everywhere; env e belongs to warp e % A
      const uint32_t todo = __ballot_sync(0xFFFFFFFFu, mine && (s_flag[lane] & (FL_SLOW | FL_RESET)) == FL_RESET);
      const uint32_t slow = __ballot_sync(0xFFFFFFFFu, mine && (s_flag[lane] & FL_SLOW));
      const uint32_t mine_a = todo & __ballot_sync(0xFFFFFFFFu, lane % A == a);
      // 1. worlds the background generator has ready: a copy
      const uint32_t got = consume_pregen<A>(p, mine_a, env0, s_bits, s_rec, s_env, lane);
      if (lane == 0 && p.stats != nullptr && mine_a != 0u) {
        if (got) atomicAdd(p.stats, (unsigned long long)__popc(got));
        if (mine_a & ~got) atomicAdd(p.stats + 1, (unsigned long long)__popc(mine_a & ~got));
      }
      if (slow == 0u && __popc(todo) <= WARP_RESET_MAX * A) {
        // 2. few finished envs (episodes drifted apart): the rest warp by warp, planes stored by the warps, ONE barrier
        if (lane == 0) for (uint32_t m = got; m != 0u; m &= m - 1u) s_flag[__ffs(m) - 1] &= ~FL_RESET;
        const bool failed = fast_resets<A>(p, mine_a & ~got, got, s_flag, s_bits, s_rec, s_env, s_scr + a * 36, env0, lane);
        if (!__syncthreads_or(failed ? 1 : 0)) return;  // the usual case: done, the barrier publishes the new episodes to the view threads
      } else {
        // 3. many (episodes in lock step): what was not pre-generated is drawn into a table by all threads and replayed
        if (lane == 0) for (uint32_t m = got; m != 0u; m &= m - 1u) { const int e = __ffs(m) - 1; s_flag[e] = (s_flag[e] & ~FL_RESET) | FL_IMAGE; }
        __syncthreads();
      }
      tile_resets<A>(p, s_flag, s_bits, s_rec, s_env, s_scr, env0, n_valid, tid);  // leftovers of 2. / the rest of 3.; then the sequential code below
    }
    if (a == lane % A && mine) {
      uint32_t fl = s_flag[lane];
      if (fl & (FL_SLOW | FL_RESET)) {
        EnvCtx<32> c{p, s_trec + lane, tp, bits, s_scr + lane, 0, 0, 0, 0u, false};
        for (int q = 0; q < A; ++q) { c.R(q, 0) = rec[q * 4]; c.R(q, 1) = rec[q * 4 + 1]; c.R(q, 2) = rec[q * 4 + 2]; }
        c.sc = s_env[lane * 4]; c.ep = s_env[lane * 4 + 1]; c.tl = s_env[lane * 4 + 2]; c.w3 = (uint32_t)s_env[lane * 4 + 3];
        const unsigned long long g = (unsigned long long)(p.env_offset + env);
        if (fl & FL_SLOW) {  // MultiGridEnv.step replayed in the reference's order (base.py:501-649)
          double rw[MG_MAX_AGENTS];
          seq_step(c, g, act_g + env * A, rw);
          bool nd = false;
          for (int q = 0; q < A; ++q) {
            nd = nd || !((c.R(q, 0) >> 24) & MG_AF_DONE);
            if (full) s_rew[lane * A + q] = rw[q]; else rew_g[env * A + q] = rw[q];
          }
          const bool dn = (c.sc >= p.max_steps) || !nd;
          if (full) s_done[lane] = dn ? 1 : 0; else done_g[env] = dn ? 1 : 0;
          fl = FL_SLOW | (c.dirty ? FL_BITS_DIRTY : 0u) | ((dn && p.autoreset) ? (FL_BITS_DIRTY | FL_RESET) : 0u);
        }
        if (fl & FL_RESET) {  // MultiGridEnv.reset (base.py:402-416), sequential code
          // an env whose replayed step edited the planes with plain stores keeps plain stores (no cross-proxy ordering games)
          if (image_planes && !(fl & FL_SLOW)) { seq_reset<false>(c, g); fl |= FL_IMAGE; } else seq_reset<true>(c, g);
          fl &= ~FL_RESET;
        }
        for (int q = 0; q < A; ++q) { rec[q * 4] = c.R(q, 0); rec[q * 4 + 1] = c.R(q, 1); rec[q * 4 + 2] = c.R(q, 2); rec[q * 4 + 3] = 0u; }
        s_env[lane * 4] = c.sc; s_env[lane * 4 + 1] = c.ep; s_env[lane * 4 + 2] = c.tl; s_env[lane * 4 + 3] = (int)c.w3;
        s_flag[lane] = fl;
      }
    }
    __syncthreads();
    if (image_planes) {
      // The byte planes of the regenerated envs (FL_IMAGE) from their bit-plane lines and object lists: env e by warp e % A,
      // coalesced 16-byte stores straight to global memory (round 1 staged plane images in the output area and sent them with
      // bulk copies it then had to wait for: four CTA-wide barriers per group of 19 envs, 2/3 of the all-reset step).
      // lane == env in every warp: the same word everywhere; warp a takes the envs e with e % A == a
      const uint32_t image_mask = __ballot_sync(0xFFFFFFFFu, mine && (s_flag[lane] & FL_IMAGE) && lane % A == a);
      store_planes_of(p, image_mask, env0, s_bits, lane);
    }
    zero_out_tile<OBS, A, SM>(s_out, tid);  // the sequential path borrowed the output area: clean it again
    __syncthreads();
  }
}

}  // namespace f2

// ---------------------------------------------------------------------------------------------
// on-device policy of the K-steps-per-launch kernel (mg_rollout_policy)
// ---------------------------------------------------------------------------------------------
// The policy is a small integer GEMM per warp and step: [32 envs of the tile] x [V*V*3 observation bytes] times
// [V*V*3] x [8 actions] (warp = agent).  It runs on the tensor cores as mma.sync m16n8k32 (u8 x s8 -> s32, exact): two row
// tiles of 16 envs, ceil(V*V*3 / 32) k-steps; the A fragments come straight from the observation tile in shared memory (a view
// starts at any byte: two aligned words + a funnel shift per fragment register), the B fragments (weights) from a
// per-launch table laid out fragment by fragment.  10 MMAs replace the 296 dp4a per THREAD of a thread-per-view dot product.
namespace f2 {
constexpr int POL_MAX_A = 4, POL_MAX_KSTEPS = 5;  // the K-steps-per-launch instantiations: A <= 4, V <= 7
// B fragments of the launch in flight: [A][k-step][register 0/1][lane] -- register j of k-step ks in lane (g = lane / 4, t = lane % 4)
// holds weight bytes k = 32 ks + 16 j + 4 t .. + 3 of action g.  Written by pack_policy_fragments, stream-ordered before the launch.
static __device__ uint32_t d_pol_frag[POL_MAX_A * POL_MAX_KSTEPS * 2 * 32];

// pol_w: the C ABI's layout [A][NW][8 actions] words (word i of action n = weight bytes 4 i .. 4 i + 3)
static __global__ void pack_policy_fragments(const int32_t* __restrict__ pol_w, int A, int NW, int ksteps) {
  for (int idx = threadIdx.x; idx < A * ksteps * 64; idx += blockDim.x) {
    const int lane = idx & 31, j = (idx >> 5) & 1, ks = (idx >> 6) % ksteps, a = (idx >> 6) / ksteps;
    const int i = ks * 8 + j * 4 + (lane & 3), n = lane >> 2;
    d_pol_frag[idx] = i < NW ? (uint32_t)pol_w[(a * NW + i) * 8 + n] : 0u;
  }
}

__device__ __forceinline__ void mma_u8s8(int (&c)[4], const uint32_t (&af)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(af[0]), "r"(af[1]), "r"(af[2]), "r"(af[3]), "r"(b0), "r"(b1));
}

// Actions of the tile's 32 envs for agent a (the whole warp takes part: mma.sync), written to `act_next` [env][A].
// s_out_u32: shared address of the observation tile [env][A][VV3]; last_s: its last aligned word (loads are clamped there: only
// zero-weight bytes lie beyond a view).
template <int VV3, int A>
__device__ __forceinline__ void policy_tile_mma(const KP& p, int a, int lane, uint32_t s_out_u32, uint32_t last_s, const int32_t* s_env,
                                                long long env0, int n_valid, int32_t* act_next) {
  constexpr int KSTEPS = (VV3 + 31) / 32;
  static_assert(KSTEPS <= POL_MAX_KSTEPS && A <= POL_MAX_A, "policy fragments");
  const int g = lane >> 2, t4 = lane & 3;
  int c[2][4];
  {
    const int2 b = __ldg(reinterpret_cast<const int2*>(p.pol_b + a * 8) + t4);  // accumulators start at the bias of columns 2 t, 2 t + 1
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) { c[mt][0] = b.x; c[mt][1] = b.y; c[mt][2] = b.x; c[mt][3] = b.y; }
  }
  uint32_t base[2][2], sh[2][2];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int h = 0; h < 2; ++h) {  // fragment rows g and g + 8 of row tile mt = envs 16 mt + 8 h + g
      const uint32_t b = s_out_u32 + (uint32_t)(((mt * 16 + h * 8 + g) * A + a) * VV3 + t4 * 4);
      base[mt][h] = b & ~3u; sh[mt][h] = (b & 3u) * 8u;
    }
  const uint32_t* const frag = d_pol_frag + a * (KSTEPS * 64) + lane;
#pragma unroll
  for (int ks = 0; ks < KSTEPS; ++ks) {
    const uint32_t b0 = frag[ks * 64], b1 = frag[ks * 64 + 32];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      uint32_t af[4];
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int h = 0; h < 2; ++h) {  // a0: (row g, k 0..15)  a1: (row g + 8, k 0..15)  a2: (row g, k 16..31)  a3: (row g + 8, k 16..31)
          uint32_t lo_a = base[mt][h] + (uint32_t)(ks * 32 + j * 16), hi_a = lo_a + 4u, lo, hi;
          if (ks == KSTEPS - 1) { lo_a = min(lo_a, last_s); hi_a = min(hi_a, last_s); }
          // (volatile + "memory": these loads read what the observe phase has just stored through ordinary pointers)
          asm volatile("ld.shared.u32 %0, [%1];" : "=r"(lo) : "r"(lo_a) : "memory");
          asm volatile("ld.shared.u32 %0, [%1];" : "=r"(hi) : "r"(hi_a) : "memory");
          af[j * 2 + h] = __funnelshift_r(lo, hi, sh[mt][h]);
        }
      mma_u8s8(c[mt], af, b0, b1);
    }
  }
  // argmax over the 8 columns of every row, lowest column on ties, columns >= pol_n excluded: key = 8 logit + (7 - column)
  long long best[2][2];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const long long k0 = (long long)c[mt][2 * h] * 8 + (7 - 2 * t4), k1 = (long long)c[mt][2 * h + 1] * 8 + (6 - 2 * t4);
      long long k = (2 * t4 < p.pol_n) ? k0 : LLONG_MIN;
      if (2 * t4 + 1 < p.pol_n && k1 > k) k = k1;
      k = max(k, __shfl_xor_sync(0xFFFFFFFFu, k, 1));
      k = max(k, __shfl_xor_sync(0xFFFFFFFFu, k, 2));
      best[mt][h] = k;
    }
  // the quad's four lanes share out its four rows
  const long long mine_k = (t4 & 2) ? ((t4 & 1) ? best[1][1] : best[1][0]) : ((t4 & 1) ? best[0][1] : best[0][0]);
  const int row = (t4 >> 1) * 16 + (t4 & 1) * 8 + g;
  if (row < n_valid) {
    int act = 7 - (int)(mine_k & 7);
    if (p.pol_eps != 0u) {
      const unsigned long long gidx = (unsigned long long)(p.env_offset + env0 + row);
      const U4 r = philox4x32_10((uint32_t)gidx, (uint32_t)(gidx >> 32), (uint32_t)s_env[row * 4 + 2], TAG_POLICY | (uint32_t)a, (uint32_t)p.pol_seed, (uint32_t)(p.pol_seed >> 32));
      if (r.x < p.pol_eps) act = (int)__umulhi(r.y, (uint32_t)p.pol_n);
    }
    act_next[(env0 + row) * A + a] = act;
  }
}
}  // namespace f2

template <int OBS, int V, int A, int NST>
constexpr int ctas_per_sm() {  // shared-memory / thread limited residency the register allocation should allow
  constexpr int by_smem = (227 * 1024) / (f2::Smem<OBS, V, A, NST>::TOTAL + (OBS == 2 ? 11 * 1024 : 0) + 1024), by_threads = 64 / A;
  constexpr int n = by_smem < by_threads ? by_smem : by_threads;
  return n > 32 ? 32 : n;
}

// Persistent CTAs: CTA c handles tiles c, c + gridDim.x, ...; while a tile is being processed the next tile's inputs are
// already on their way into the other input stage (NST = 2), and the previous tile's outputs drain in the background.
// KS (K steps per launch, mg_rollout_persistent): the CTA keeps the state of its (at most NST) tiles in its input stages and
// plays n_steps steps on them -- actions[step], observations[step], rewards[step], done[step] are per-step slices --, so that
// an open-loop rollout costs one launch, no state reloads and no grid-wide dependency between steps.
template <int OBS, int V, int A, bool VO0, int NST, bool KS = false, bool HIDE = false>
__global__ void __launch_bounds__(32 * A, ctas_per_sm<OBS, V, A, NST>()) fused2_kernel(const __grid_constant__ KP p, const int n_tiles, const int n_steps) {
  using namespace f2;
  using SM = Smem<OBS, V, A, NST>;
  constexpr int VV3 = V * V * 3;
  constexpr uint32_t RM = (1u << V) - 1u;
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, a = tid >> 5;  // lane = env within the tile, warp = agent
  const int W = p.W, H = p.H, S = p.S;
  // fresh worlds hold canonical walls plus (goal + bonus tiles) listed objects: if those always fit the object list, reset
  // leaves the byte planes to the cooperative image below
  const bool image_planes = (p.goal_mode != MG_GOAL_NONE ? 1 : 0) + p.n_bonus <= OBJ_SLOTS && 3 * S <= SM::OUT_AREA;

  double* s_rew = reinterpret_cast<double*>(smem + SM::REW);       // [env][a]
  uint8_t* s_done = smem + SM::DONE;                               // [env]
  uint32_t* s_order = reinterpret_cast<uint32_t*>(smem + SM::ORDER);
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + SM::BAR);
  uint8_t* s_out = smem + SM::OUT;
  uint8_t* const out = s_out + (lane * A + a) * VV3;
  const uint32_t out_s = smem_u32(out);

  auto zero_out = [&]() { f2::zero_out_tile<OBS, A, SM>(s_out, tid); };
  // one tile's inputs as four bulk copies on the stage's mbarrier.  Issuing a bulk copy costs its thread a few hundred
  // cycles, so the four go out from lane 0 of different warps; thread 0 arms the barrier with the byte count (the
  // transaction count may run ahead of it, the phase cannot complete before this arrival).
  auto issue_load = [&](int tile, int stage) {
    if (lane != 0) return;
    const long long e0 = (long long)tile * ENVS_PER_CTA;
    const int nv = (int)min((long long)ENVS_PER_CTA, p.B - e0);
    unsigned char* st = smem + stage * SM::STAGE;
    const uint32_t wbytes = (uint32_t)(ENVS_PER_CTA * BITS_WORDS * 4) /* always the whole tile-transposed chunk */, rbytes = (uint32_t)nv * (A * 16u), ebytes = (uint32_t)nv * 16u;
    const uint32_t abytes = (!KS && nv == ENVS_PER_CTA) ? (uint32_t)(ENVS_PER_CTA * A * 4) : 0u;  // KS: actions change every step, read directly
    if (a == 0) {
      mbar_expect_tx(s_bar + stage, wbytes + rbytes + ebytes + abytes);
      bulk_g2s(st + SM::ST_BITS, p.cellbits + e0 * BITS_WORDS, wbytes, s_bar + stage);
    }
    if (a == 1 % A) bulk_g2s(st + SM::ST_REC, p.agents + e0 * A * 16, rbytes, s_bar + stage);
    if (a == 2 % A) {
      bulk_g2s(st + SM::ST_ENV, p.envrec + e0 * 4, ebytes, s_bar + stage);
      if (abytes) bulk_g2s(st + SM::ST_ACT, p.actions + e0 * A, abytes, s_bar + stage);
    }
  };

  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < NST; ++i) mbar_init(s_bar + i, 1);
  }
  if (a == 0) { reinterpret_cast<uint32_t*>(smem + SM::FLAG)[lane] = 0u; reinterpret_cast<uint32_t*>(smem + SM::FLAG)[32 + lane] = 0u; }
  __syncthreads();
  uint8_t* const s_atlas = smem + SM::TOTAL;  // OBS 2: tiles [n_tiles] + the shadow tile, 192 bytes each (tile size 8)
  if (OBS == 2) {  // the atlas is constant data: copy it while the previous kernel may still be running
    const int4* src = reinterpret_cast<const int4*>(p.atlas);
    int4* dst = reinterpret_cast<int4*>(s_atlas);
    for (int i = tid; i < p.n_tiles * 12; i += 32 * A) dst[i] = __ldg(src + (i / 12) * 48 + (i % 12));  // global: [tile][4 orientations][192]; slot 0 only
    // COLORS['shadow'] = (35, 25, 30) (objects.py:25, base.py:305): the byte pattern repeats every three words
    for (int i = tid; i < 48; i += 32 * A)
      reinterpret_cast<uint32_t*>(s_atlas + p.n_tiles * 192)[i] = (i % 3 == 0) ? 0x231E1923u : (i % 3 == 1) ? 0x19231E19u : 0x1E19231Eu;
  }
  // programmatic dependent launch: this grid may have been started while the previous kernel of the stream was still
  // draining (its launch latency and this prologue overlap that tail); nothing before this line touches global memory
  // While this CTA waits for the previous kernel of the stream, its first tile's inputs can already travel from HBM to L2: a
  // prefetch reads nothing into the CTA, so it cannot observe stale data -- whatever the previous kernel still writes lands in
  // L2, the point of coherence, and the real loads below are issued after the dependency is resolved.  (A family that is
  // stepped back to back finds its state in L2 anyway; this is for batches / round-robin families larger than L2.)
  auto prefetch_tile = [&](int tile) {
    if (lane != 0) return;
    const long long e0 = (long long)tile * ENVS_PER_CTA;
    const int nv = (int)min((long long)ENVS_PER_CTA, p.B - e0);
    auto prefetch_l2 = [](const void* src, uint32_t bytes) { asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory"); };
    if (a == 0) prefetch_l2(p.cellbits + e0 * BITS_WORDS, (uint32_t)(ENVS_PER_CTA * BITS_WORDS * 4));
    if (a == 1 % A) prefetch_l2(p.agents + e0 * A * 16, (uint32_t)nv * (A * 16u));
    if (a == 2 % A) {
      prefetch_l2(p.envrec + e0 * 4, (uint32_t)nv * 16u);
      if (nv == ENVS_PER_CTA) prefetch_l2(p.actions + e0 * A, (uint32_t)(ENVS_PER_CTA * A * 4));
    }
  };
  if (!KS && OBS == 1 && p.f2_prefetch && (int)blockIdx.x < n_tiles) prefetch_tile((int)blockIdx.x);  // (OBS 2 is bound by its 38 KB of pixels per env: measured slightly worse with it)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if ((int)blockIdx.x < n_tiles) issue_load((int)blockIdx.x, 0);
  const int my_tiles = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // KS: <= NST, all resident
  if (KS) {
    for (int j = 1; j < my_tiles; ++j) issue_load((int)blockIdx.x + j * (int)gridDim.x, j);
  }
  uint32_t dirty_acc0 = 0, dirty_acc1 = 0;  // KS, warp 0: this env's bit-plane lines changed at some step (stored after the last one)

  int chunk_parity = 0;  // OBS 2: which of the warp's chunk buffers is filled next
  const int n_iter = KS ? my_tiles * n_steps : my_tiles;
  for (int it = 0; it < n_iter; ++it) {
  const int step = KS ? it / my_tiles : 0;
  const int tile = (int)blockIdx.x + (KS ? it % my_tiles : it) * (int)gridDim.x;
  const int stage = KS ? it % my_tiles : it % NST;
  const bool last_step = !KS || step == n_steps - 1;
  // this step's slices of the caller's arrays
  const int32_t* const act_g = p.actions + (KS ? (long long)step * p.B * A : 0);
  double* const rew_g = p.rewards + (KS ? (long long)step * p.B * A : 0);
  uint8_t* const done_g = p.done + (KS ? (long long)step * p.B : 0);
  uint8_t* const obs_g = p.obs + (KS ? (long long)step * p.B * (A * V * V * 3) : 0);
  unsigned char* const stg = smem + stage * SM::STAGE;
  uint32_t* const s_bits = reinterpret_cast<uint32_t*>(stg + SM::ST_BITS);
  uint32_t* const s_rec = reinterpret_cast<uint32_t*>(stg + SM::ST_REC);   // [env][a][4]
  int32_t* const s_env = reinterpret_cast<int32_t*>(stg + SM::ST_ENV);     // [env][4]
  const int32_t* const s_act = reinterpret_cast<const int32_t*>(stg + SM::ST_ACT);  // [env][a]
  uint32_t* const s_flag = reinterpret_cast<uint32_t*>(smem + SM::FLAG) + (it & 1) * 32;
  const long long env0 = (long long)tile * ENVS_PER_CTA;
  const int n_valid = (int)min((long long)ENVS_PER_CTA, p.B - env0);
  const bool full = n_valid == ENVS_PER_CTA;
  const bool mine = lane < n_valid;
  const long long env = env0 + lane;
  const int next_tile = tile + (int)gridDim.x;
  if (NST == 1 && it > 0) {  // single stage: the inputs can only be requested once the previous tile has drained
    if (lane == 0 || a == 0) bulk_wait_read0();  // every thread that issued bulk stores reading the stage (or the chunk buffers)
    __syncthreads();
    issue_load(tile, 0);
  }

  int action = MG_A_DONE;
  if ((KS || !full) && mine) action = KS ? __ldcg(act_g + env * A + a) : act_g[env * A + a];  // (KS: the policy hand-off below may have written it a step ago)
  if (!KS || step == 0) mbar_wait(s_bar + stage, KS ? 0u : (uint32_t)((it / NST) & 1));
  // single input stage: the next tile's inputs cannot be loaded before this tile has drained -- but they can come as far as L2
  if (!KS && OBS == 1 && NST == 1 && p.f2_prefetch && next_tile < n_tiles) prefetch_tile(next_tile);
  if (!KS && full) action = s_act[lane * A + a];

  // ---- warp 0: the step's agent order, base.py:514-516 -- one Philox block per env ----
  if (a == 0 && mine) {
    // an env about to time out (base.py:649) will want its pre-generated world a microsecond from now: start it on its way to L2
    if (p.pregen != nullptr && s_env[lane * 4] + 1 >= p.max_steps) {
      const uint32_t* const slot = p.pregen + env * world::PG_WORDS;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(slot));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(slot + 32));
    }
    const unsigned long long g = (unsigned long long)(p.env_offset + env);
    const U4 r = philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), (uint32_t)s_env[lane * 4 + 2], 0u, (uint32_t)p.seed, (uint32_t)(p.seed >> 32));
    s_order[lane] = decode_order_ct<A>(__umulhi(r.x, Fact<A>::v));
  }

  // ---- every agent plays its action on a private copy of its record (base.py:517-622) ----
  uint32_t* const rec = s_rec + lane * (A * 4);
  uint32_t* const bits = s_bits + lane;  // word w of this env at bits[w * BS]: bank == lane for every w
  uint8_t* const tp = p.grid + env * 3 * S;
  uint32_t w0 = 0, w1 = 0, errb = 0, base_stamp = 0;
  bool moved = false, slow = false;
  double reward = 0.0;
  if (mine) {
    const uint2 r01 = *reinterpret_cast<const uint2*>(rec + a * 4);
    w0 = r01.x; w1 = r01.y;
    base_stamp = (uint32_t)s_env[lane * 4 + 3] & 0xFFFFu;  // read before warp 0 advances it
    if ((w0 >> 24) & MG_AF_ACTIVE) {  // base.py:521
      const int cx = (int)(w0 & 0xFFu), cy = (int)((w0 >> 8) & 0xFFu), dir = (int)((w0 >> 16) & 3u);
      if (action == MG_A_LEFT) w0 = (w0 & 0xFF00FFFFu) | ((uint32_t)((dir + 3) & 3) << 16);        // base.py:530-531
      else if (action == MG_A_RIGHT) w0 = (w0 & 0xFF00FFFFu) | ((uint32_t)((dir + 1) & 3) << 16);  // base.py:534-535
      else if (action >= MG_A_FORWARD && action <= MG_A_TOGGLE) {
        const int fx = cx + ((dir == 0) ? 1 : (dir == 2) ? -1 : 0), fy = cy + ((dir == 1) ? 1 : (dir == 3) ? -1 : 0);  // agents.py:183
        const bool inb = (unsigned)fx < (unsigned)W && (unsigned)fy < (unsigned)H;
        if (!inb) errb |= MG_ERR_STACK;  // grid.get asserts in-bounds (base.py:154-156); never hit inside wall_rect
        const uint32_t fc = inb ? ((bits[(LINE_X0 + (fx & 15)) * BS] >> (fy & 15)) & 0x10001u) : 1u;  // 0 empty, 1 canonical wall, else "other"
        if (fc <= 1u) {  // empty or wall in front: only forward can do anything (pickup / toggle need an object, drop needs hands full)
          if (action == MG_A_FORWARD) {
            if (fc == 0u) {
              const uint32_t cc = (bits[(LINE_X0 + (cx & 15)) * BS] >> (cy & 15)) & 0x10001u;
              if (cc != 0u) {  // leaving a cell that holds a static object: it must be overlappable (base.py:558)
                const uint32_t ccell = cell_triple(bits, cx & 15, cy & 15, tp, H, S);
                if (!can_overlap_static((int)(ccell & 0xFFu), (int)(ccell >> 16))) errb |= MG_ERR_STACK;
              }
              w0 = (w0 & 0xFFFF0000u) | (uint32_t)fx | ((uint32_t)fy << 8);
              moved = true;
            }
          } else if (action == MG_A_DROP) {
            slow = inb && fc == 0u && (w1 & 0xFFu) != 0u;  // carrying and facing an empty cell (base.py:600-606)
          }
        } else {  // an object other than a canonical wall
          const uint32_t fcell = cell_triple(bits, fx & 15, fy & 15, tp, H, S);
          const int ftype = (int)(fcell & 0xFFu), fstate = (int)(fcell >> 16);
          if (action == MG_A_FORWARD) {  // base.py:538-585 (ghost mode: other agents never block)
            if (can_overlap_static(ftype, fstate)) {
              const uint32_t cc = (bits[(LINE_X0 + (cx & 15)) * BS] >> (cy & 15)) & 0x10001u;
              if (cc != 0u) {
                const uint32_t ccell = cell_triple(bits, cx & 15, cy & 15, tp, H, S);
                if (!can_overlap_static((int)(ccell & 0xFFu), (int)(ccell >> 16))) errb |= MG_ERR_STACK;
              }
              w0 = (w0 & 0xFFFF0000u) | (uint32_t)fx | ((uint32_t)fy << 8);
              moved = true;
              if (ftype == MG_T_GOAL || ftype == MG_T_BONUS) {  // base.py:576-581
                double rwd;
                if (ftype == MG_T_GOAL) rwd = p.goal_reward;
                else {  // BonusTile.get_reward objects.py:180-206 on the private copy of w1
                  const int n = p.n_bonus, bonus_id = fstate;
                  int bs = (int)(w1 >> 24);
                  bool first = false;
                  const double pen = p.bonus_penalty < 0 ? p.bonus_penalty : -p.bonus_penalty;
                  if (bs == 0xFF) { bs = ((bonus_id - 1) % n + n) % n; first = true; }
                  if (bs == bonus_id) rwd = pen;
                  else if ((bs + 1) % n == bonus_id) { bs = bonus_id; rwd = p.bonus_reward; }
                  else rwd = pen;
                  if (p.flags & MG_F_BONUS_RESET) bs = bonus_id;
                  w1 = (w1 & 0x00FFFFFFu) | ((uint32_t)bs << 24);
                  if (first && !(p.flags & MG_F_BONUS_INITIAL)) rwd = 0.0;
                }
                if (p.flags & MG_F_REWARD_DECAY) {  // base.py:579, every operation rounded on its own
                  const int sc = s_env[lane * 4] + 1;  // base.py:512
                  const double qd = __ddiv_rn((double)sc, (double)p.max_steps);
                  const double u = __dmul_rn(0.9, qd);
                  const double f = __dsub_rn(1.0, u);
                  rwd = __dmul_rn(rwd, f);
                }
                reward = __dadd_rn(0.0, rwd);  // step_rewards[agent_no] += rwd (base.py:580): 0.0 + (-0.0) is +0.0
              }
              if (ftype == MG_T_LAVA || ftype == MG_T_GOAL) {
                w0 = (w0 | ((uint32_t)MG_AF_DONE << 24)) & ~((uint32_t)MG_AF_ACTIVE << 24);  // base.py:584-585,646
                // respawn=True (base.py:626-644): the agent leaves its queue and is placed anew inside this step -- order-dependent:
                // the env is replayed by the sequential code.  (With respawn no agent ever starts a step `done`; worlds without
                // Goal / Lava -- the goal-cycle scenario of examples/human_player.py:35-55 -- never come here.)
                if (p.flags & MG_F_RESPAWN) slow = true;
              }
            }
          } else if (action == MG_A_PICKUP) {  // takes effect only on a pickable object with empty hands (base.py:590-597)
            slow = ((PICKUP_MASK >> ftype) & 1u) && (w1 & 0xFFu) == 0u;
          } else if (action == MG_A_TOGGLE) {  // only Door / Box react (base.py:609-613)
            slow = ftype == MG_T_DOOR || ftype == MG_T_BOX;
          }
        }
      } else if (action != MG_A_DONE) errb |= MG_ERR_BAD_ACTION;  // base.py:619-620
    }
    // one shared-memory atomic per agent: slow request / mover bit / "not done yet" bit / error bits
    const uint32_t add = (slow ? FL_SLOW : 0u) | (moved ? (0x100u << a) : 0u) | (((w0 >> 24) & MG_AF_DONE) ? 0u : FL_NOTDONE) | (errb << 16);
    if (add) atomicOr(&s_flag[lane], add);
  }
  if (lane == 0 || a == 0) bulk_wait_read0();  // the previous tile's stores (issued by these threads) have left shared memory
  __syncthreads();

  if (!KS && NST > 1 && next_tile < n_tiles) issue_load(next_tile, (it + 1) % NST);  // prefetch: lands during this tile's observe phase
  zero_out();
  if (a == 0) (reinterpret_cast<uint32_t*>(smem + SM::FLAG) + ((it + 1) & 1) * 32)[lane] = 0u;  // next tile's flag words

  // ---- commit the parallel envs; envs whose planes change (FL_SLOW) are replayed below ----
  const uint32_t fl1 = mine ? s_flag[lane] : 0u;
  bool rare = false;
  if (mine && !(fl1 & FL_SLOW)) {
    uint32_t stamp = rec[a * 4 + 2];
    if (moved) {  // arrival stamp: movers are numbered in the reference's processing order (base.py:547-552)
      const uint32_t order = s_order[lane], movers = (fl1 >> 8) & 0xFFu;
      uint32_t rank = 0;
      bool before = true;
#pragma unroll
      for (int q = 0; q < A; ++q) {
        const uint32_t b = (order >> (4 * q)) & 0xFu;
        before = before && (b != (uint32_t)a);
        rank += (before ? (movers >> b) & 1u : 0u);
      }
      stamp = (base_stamp + rank) & 0xFFFFu;
    }
    *reinterpret_cast<uint4*>(rec + a * 4) = make_uint4(w0, w1, stamp, 0u);
    if (full) s_rew[lane * A + a] = reward; else rew_g[env * A + a] = reward;
    if (a == 0) {  // env bookkeeping and done (base.py:512,649)
      int4 er = *reinterpret_cast<const int4*>(s_env + lane * 4);
      const uint32_t w3 = (uint32_t)er.w;
      er.x += 1;  // step_count
      er.z += 1;  // lifetime steps
      er.w = (int)((w3 & 0xFFFF0000u) | (((w3 & 0xFFFFu) + (uint32_t)__popc((fl1 >> 8) & 0xFFu)) & 0xFFFFu) | (fl1 & 0xFFFF0000u));
      *reinterpret_cast<int4*>(s_env + lane * 4) = er;
      const bool dn = (er.x >= p.max_steps) || !(fl1 & FL_NOTDONE);
      if (full) s_done[lane] = dn ? 1 : 0; else done_g[env] = dn ? 1 : 0;
      if (dn && p.autoreset) { rare = true; s_flag[lane] = fl1 | FL_BITS_DIRTY | FL_RESET; }
    }
  } else if (mine && a == 0) rare = true;
  // tiles with a finished env or an env whose planes change: everything rare lives in ONE out-of-line function that
  // recomputes its pointers, so that it costs the common path a call site and nothing else
  if (__syncthreads_or(rare ? 1 : 0)) {
    // Called through a pointer the compiler cannot see through: a direct call lets ptxas allocate registers across caller and
    // callee, and the values it then spills are stored where they are defined -- on the common path.  An opaque call follows
    // the ABI instead: the callee saves the registers it clobbers in its own prologue, and the common path has no local-memory
    // traffic at all (0 bytes of spills, against 24 / 44 before).
    void (*fn)(const KP&, int, int, int, int) = &f2::rare_path<OBS, V, A, NST, KS>;
    asm volatile("" : "+l"(fn));
    fn(p, tile, stage, it, step);
  }

  // ---- observe the post-step world: gen_obs_grid + occlude_mask + encode ----
  if (mine) {
    // All records of the env.  Queue heads (the agent with the smallest stamp on its cell is the cell's object or
    // `static_obj.agents[0]`, base.py:547-572) are recomputed by every view thread, branch-free: A is tiny, and it saves a
    // barrier.  Composite key = cell << 16 | stamp (an unplaced agent gets a cell nobody can stand on).
    uint32_t q0[A], ck[A];
    uint32_t heads = (1u << A) - 1u;
#pragma unroll
    for (int q = 0; q < A; ++q) {
      const uint4 r = *reinterpret_cast<const uint4*>(rec + q * 4);
      q0[q] = r.x;
      const bool placed = (r.x & ((uint32_t)MG_AF_PLACED << 24)) != 0u;
      ck[q] = placed ? __byte_perm(r.z, r.x, 0x5410) : (0xFF000000u | ((uint32_t)q << 16));
      if (!placed) heads &= ~(1u << q);
    }
#pragma unroll
    for (int q = 0; q < A; ++q)
#pragma unroll
      for (int r = q + 1; r < A; ++r) {
        const bool same = (ck[q] ^ ck[r]) < 0x10000u;  // same cell: the one that arrived later is not the head
        const uint32_t loser = (ck[q] < ck[r]) ? (1u << r) : (1u << q);
        if (same) heads &= ~loser;
      }
    const uint32_t me = rec[a * 4];
    uint8_t* const tmap = s_out + (lane * A + a) * (V * 8);  // OBS 2: this view's tile ids, V rows of 8 bytes
    if (OBS == 2 && !(me & ((uint32_t)MG_AF_ACTIVE << 24))) {  // inactive agent: every cell is shadow (base.py:305,420-425)
      const uint32_t sh4 = (uint32_t)p.n_tiles * 0x01010101u;
#pragma unroll
      for (int b = 0; b < V; ++b) *reinterpret_cast<uint2*>(tmap + b * 8) = make_uint2(sh4, sh4);
    }
    if (me & ((uint32_t)MG_AF_ACTIVE << 24)) {  // inactive agent: empty view, nothing visible (base.py:420-425)
      const int px = (int)(me & 0xFFu), py = (int)((me >> 8) & 0xFFu), dir = (int)((me >> 16) & 3u);
      constexpr int h = V / 2;
      const int vo = VO0 ? 0 : p.vo;
      // agents.py:237-266 get_view_exts; u = axis the agent faces along (view rows), v = axis across (bits of a row)
      const int topX = (dir == 0) ? px - vo : (dir == 2) ? px - V + 1 + vo : px - h;
      const int topY = (dir == 1) ? py - vo : (dir == 3) ? py - V + 1 + vo : py - h;
      const bool vertical = (dir & 1) != 0, flip = dir < 2, rev = ((dir + 1) & 2) != 0;
      const int u0 = vertical ? topY : topX, v0 = vertical ? topX : topY;
      // window of a line: the 16 line bits are parked in the top half of a word (zeros below), shifted down to bit 0.
      //   as stored:    OP = bits 0..15, OT = bits 16..31, view column a <-> line bit v0 + a
      //   rotated view: the line is bit-reversed and its halves swapped back, view column a <-> line bit v0 + V-1 - a
      const int shw = rev ? (32 - V - v0) : (v0 + 16);
      // line of view row b = u0 + b (or u0 + V-1 - b): slot index + 1, clamped as unsigned to the zero guard lines 0 and 17
      const uint32_t lines_s = smem_u32(bits + (vertical ? LINE_Y0 - 1 : LINE_X0 - 1) * BS);
      const int ustep = flip ? -1 : 1, ubase = (flip ? u0 + V - 1 : u0) + 1;
      uint32_t T[V], OT[V], OP[V], M[V];
#pragma unroll
      for (int b = 0; b < V; ++b) {
        uint32_t w;
        asm("ld.shared.u32 %0, [%1];" : "=r"(w) : "r"(lines_s + (4u * BS) * min((uint32_t)(ubase + b * ustep), 17u)));
        if (rev) w = __byte_perm(__brev(w), 0u, 0x1032);  // reversed line, OP back in the low half
        OP[b] = (w << 16) >> shw;                          // bits above the window (neighbouring cells) are masked by
        OT[b] = (w & 0xFFFF0000u) >> shw;                  // the visibility rows below
        T[b] = ~OP[b] & RM;
      }
      if (p.flags & MG_F_SEE_THROUGH) {  // agents.py:294-295
#pragma unroll
        for (int b = 0; b < V; ++b) M[b] = RM;
      } else if (VO0) {
        occlude_rows_vo0<V>(T, M);
      } else {
        occlude_rows<V>(T, V / 2, V - 1 - vo, M);  // agents.py:233-234,293
      }
      // rows packed one byte each (view cell (a, b) = bit 8*b + a): visible canonical walls / other objects / free cells
      const uint64_t m64 = pack_rows8<V>(M), op64 = pack_rows8<V>(OP), ot64 = pack_rows8<V>(OT);
      // HIDE (hide_item_types, agents.py:30, base.py:441-449; encoded observations): hidden canonical walls are not drawn, a hidden
      // object leaves its cell (hid64) to the head of the agents standing on it, and a head agent other than the observer gives
      // way to the second of its queue when agents are hidden -- every cell is replaced once, after the line of sight was
      // computed on the real grid (the same rules as obs_view_hidden_masks, mg_obs.cuh)
      const uint64_t wall64 = (HIDE && ((p.hide >> MG_T_WALL) & 1u)) ? 0ull : (m64 & op64 & ~ot64), free64 = m64 & ~(op64 | ot64);
      uint64_t oth64 = m64 & ot64, hid64 = 0ull;
      if (OBS == 1) {
      {  // visible canonical walls (8, 9, 0): constants at compile-time offsets of the staging tile.  The two constants
           // come in through a kernel parameter and the stores are spelled out, or ptxas re-materialises 8 / 9 around every store
          const uint32_t wlo = (uint32_t)wall64, whi = (uint32_t)(wall64 >> 32);
          const uint32_t c8 = p.wall_enc & 0xFFu, c9 = p.wall_enc >> 8;  // (MG_T_WALL, MG_C_WORST) through a kernel parameter
#pragma unroll
          for (int b = 0; b < V; ++b)
#pragma unroll
            for (int va = 0; va < V; ++va) {
              if ((b < 4 ? wlo : whi) & (1u << (8 * (b & 3) + va))) {
                sts_u8(out_s + (va * (V * 3) + b * 3 + 0), c8);
                sts_u8(out_s + (va * (V * 3) + b * 3 + 1), c9);
              }
            }
        }
      }
      if (OBS == 2) {
        // tile ids (render_tile base.py:275-299): shadow where invisible, 0 for a visible empty cell, the Wall tile for a
        // visible canonical wall -- per row, seven mask bits spread to seven bytes and scaled (no carries: one term per byte)
        const uint32_t shadow = (uint32_t)p.n_tiles, wall_tile = (uint32_t)p.kind_of_type[MG_T_WALL] * (1u + 4u * A);
#pragma unroll
        for (int b = 0; b < V; ++b) {
          const uint32_t vis = (uint32_t)(m64 >> (8 * b)) & 0xFFu, wl = (uint32_t)(wall64 >> (8 * b)) & 0xFFu;
          const uint32_t v_lo = ((vis & 15u) * 0x00204081u) & 0x01010101u, v_hi = ((vis >> 4) * 0x00204081u) & 0x01010101u;
          const uint32_t w_lo = ((wl & 15u) * 0x00204081u) & 0x01010101u, w_hi = ((wl >> 4) * 0x00204081u) & 0x01010101u;
          *reinterpret_cast<uint2*>(tmap + b * 8) = make_uint2((0x01010101u - v_lo) * shadow + w_lo * wall_tile, (0x01010101u - v_hi) * shadow + w_hi * wall_tile);
        }
      }
      // world cell -> view cell: vb = bu + su*cu, va = bv + sv*cv with (cu, cv) = (x, y) or (y, x)
      const int su = flip ? -1 : 1, bu = flip ? V - 1 + u0 : -u0, sv = rev ? -1 : 1, bv = rev ? V - 1 + v0 : -v0;
      bool bad_render = false;  // OBS 2: an object whose render() raises in the reference (objects.py:274-277,309-321,370)
      if (oth64 != 0ull) {  // visible Goal / BonusTile / Key ...: the object list answers for (almost) all of them
#pragma unroll
        for (int k = 0; k < OBJ_SLOTS; ++k) {
          const uint32_t e = bits[(OBJ_WORD0 + k) * BS];
          if (!(e >> 31)) continue;
          const int ex = (int)(e & 15u), ey = (int)((e >> 4) & 15u);
          const int vb = bu + su * (vertical ? ey : ex), va = bv + sv * (vertical ? ex : ey);
          if ((unsigned)vb >= (unsigned)V || (unsigned)va >= (unsigned)V) continue;
          const uint64_t m = 1ull << (8 * vb + va);
          if (!(oth64 & m)) continue;
          oth64 &= ~m;
          if (HIDE && ((p.hide >> ((e >> 8) & 15u)) & 1u)) { hid64 |= m; continue; }
          if (OBS == 1) {
            uint8_t* oo = out + va * (V * 3) + vb * 3;
            oo[0] = (uint8_t)((e >> 8) & 15u); oo[1] = (uint8_t)((e >> 12) & 15u); oo[2] = (uint8_t)((e >> 16) & 255u);
          } else {
            const uint32_t kind = p.kind_of_type[(e >> 8) & 15u];
            if (kind == 0xFFu) bad_render = true; else tmap[vb * 8 + va] = (uint8_t)(kind * (1u + 4u * A));
          }
        }
        // objects that did not fit the list: WorldObj.encode (objects.py:90-99) from the byte planes
        while (oth64 != 0ull) {
          const int bit = __ffsll((long long)oth64) - 1;
          oth64 &= oth64 - 1ull;
          const int va = bit & 7, vb = bit >> 3;
          const int uu = flip ? V - 1 - vb : vb, vv = rev ? V - 1 - va : va;
          const int wx = topX + (vertical ? vv : uu), wy = topY + (vertical ? uu : vv);
          const uint8_t* cp = p.grid + env * 3 * S + wx * H + wy;
          if (HIDE && ((p.hide >> cp[0]) & 1u)) { hid64 |= 1ull << bit; continue; }
          if (OBS == 1) {
            uint8_t* oo = out + va * (V * 3) + vb * 3;
            oo[0] = cp[0]; oo[1] = cp[S]; oo[2] = cp[2 * S];
          } else {
            const uint32_t kind = p.kind_of_type[cp[0] & 15u];
            if (kind == 0xFFu) bad_render = true; else tmap[vb * 8 + va] = (uint8_t)(kind * (1u + 4u * A));
          }
        }
      }
      // agents that are their cell's object: (13, colour, dir) where no static object stands (base.py:204-214)
      const uint32_t sel_u = vertical ? 0x4441u : 0x4440u, sel_v = vertical ? 0x4440u : 0x4441u;
#pragma unroll
      for (int q = 0; q < A; ++q) {
        const int vb = bu + su * (int)__byte_perm(q0[q], 0u, sel_u), va = bv + sv * (int)__byte_perm(q0[q], 0u, sel_v);
        const bool in_view = (unsigned)vb < (unsigned)V && (unsigned)va < (unsigned)V;
        if (OBS == 1) {
          const bool draw = in_view && ((heads >> q) & 1u) && (((HIDE ? (free64 | hid64) : free64) >> ((8 * vb + va) & 63)) & 1ull);
          if (draw) {
            uint32_t colour = p.agent_color[q], qdir = (q0[q] >> 16) & 3u;
            bool show = true;
            if (HIDE && ((p.hide >> MG_T_AGENT) & 1u) && q != a && !((hid64 >> ((8 * vb + va) & 63)) & 1ull)) {
              show = false;  // an agent as its cell's object, hidden: the second of the queue on that cell, or nothing
              uint32_t best = 0xFFFFFFFFu;
#pragma unroll
              for (int r = 0; r < A; ++r)
                if (r != q && (ck[r] ^ ck[q]) < 0x10000u && ck[r] < best) { best = ck[r]; colour = p.agent_color[r]; qdir = (q0[r] >> 16) & 3u; show = true; }
            }
            if (show) {
              uint8_t* oo = out + va * (V * 3) + vb * 3;
              oo[0] = MG_T_AGENT; oo[1] = (uint8_t)colour; oo[2] = (uint8_t)qdir;
            }
          }
        } else {
          // the cell's tile gets an agent on top: the observer itself if it stands there, else the queue head
          // (base.py:282-293); tiles of size <= 10 are rotation-equivariant, so the view orientation is a dir remap
          const bool draw = in_view && ((heads >> q) & 1u) && ((m64 >> ((8 * vb + va) & 63)) & 1ull);
          if (draw) {
            const bool own = ((q0[q] ^ me) & 0xFFFFu) == 0u;
            const uint32_t qq = own ? (uint32_t)a : (uint32_t)q, qd = ((own ? me : q0[q]) >> 16) & 3u;
            tmap[vb * 8 + va] = (uint8_t)(tmap[vb * 8 + va] + 1u + 4u * qq + ((qd + 3u - (uint32_t)dir) & 3u));
          }
        }
      }
      if (OBS == 2 && bad_render) atomicOr(reinterpret_cast<unsigned int*>(s_env) + lane * 4 + 3, (unsigned int)MG_ERR_RENDER << 16);
    }
  }
  fence_proxy_async_smem();  // writer side of the generic -> async proxy hand-over: every thread's shared-memory writes of this
  __syncthreads();           // tile (observations, records, rewards ...) are ordered before the bulk copies issued below

  // ---- policy hand-off (mg_rollout_policy): the actions of the NEXT step from the observation tile just written ----
  if constexpr (KS && OBS == 1 && A <= f2::POL_MAX_A) {
    if (p.pol_w != nullptr && step + 1 < n_steps)  // (uniform: every lane takes part in the MMAs, ragged tiles included)
      f2::policy_tile_mma<VV3, A>(p, a, lane, smem_u32(s_out), smem_u32(s_out) + (uint32_t)SM::OUT_BYTES - 4u, s_env, env0, n_valid,
                                  const_cast<int32_t*>(p.actions) + (long long)(step + 1) * p.B * A);
  }

  // ---- everything leaves as contiguous chunks; nobody waits for them here ----
  if (full) {
    if (lane == 0) {  // again spread over the warps' first lanes
      fence_proxy_async_smem();
      if (OBS == 1 && a == A - 1) bulk_s2g(obs_g + env0 * (A * VV3), s_out, (uint32_t)SM::OUT_BYTES);
      if (a == 0) { if (last_step) bulk_s2g(p.agents + env0 * A * 16, s_rec, ENVS_PER_CTA * A * 16u); bulk_s2g(done_g + env0, s_done, ENVS_PER_CTA); }
      if (a == 1 % A && last_step) bulk_s2g(p.envrec + env0 * 4, s_env, ENVS_PER_CTA * 16u);
      if (a == 2 % A) bulk_s2g(rew_g + env0 * A, s_rew, ENVS_PER_CTA * A * 8u);
      bulk_commit();
    }
  } else {  // ragged last tile: plain stores
    if (OBS == 1) {
      const int total = n_valid * A * VV3;
      uint8_t* dst = obs_g + env0 * (A * VV3);
      for (int i = tid; i < total; i += 32 * A) dst[i] = s_out[i];
    }
    if (tid == 0 && last_step) {
      fence_proxy_async_smem();
      bulk_s2g(p.agents + env0 * A * 16, s_rec, (uint32_t)n_valid * (A * 16u));
      bulk_s2g(p.envrec + env0 * 4, s_env, (uint32_t)n_valid * 16u);
      bulk_commit();
    }
  }
  if (KS && a == 0 && mine) { if (stage) dirty_acc1 |= s_flag[lane] & FL_BITS_DIRTY; else dirty_acc0 |= s_flag[lane] & FL_BITS_DIRTY; }
  if (a == 0) {
    // envs whose bit-plane lines changed (reset, plane edit): many -> the tile's chunk goes back whole; few -> their words
    // alone (an env's words are 128 bytes apart in the tile-transposed layout: lanes = words, plain stores)
    const uint32_t dm = __ballot_sync(0xFFFFFFFFu, mine && last_step && ((KS ? (stage ? dirty_acc1 : dirty_acc0) : s_flag[lane]) & FL_BITS_DIRTY));
    if (dm != 0u) {
      uint32_t* const gb = p.cellbits + env0 * BITS_WORDS;
      if (__popc(dm) > 6) {
        if (lane == 0) {
          fence_proxy_async_smem();
          bulk_s2g(gb, s_bits, (uint32_t)(ENVS_PER_CTA * BITS_WORDS * 4));
          bulk_commit();
        }
      } else {
        for (uint32_t m = dm; m != 0u; m &= m - 1u) {
          const int e = __ffs(m) - 1;
          gb[lane * BS + e] = s_bits[lane * BS + e];
          if (lane + 32 < BITS_WORDS) gb[(lane + 32) * BS + e] = s_bits[(lane + 32) * BS + e];
        }
      }
    }
  }
  if (OBS == 2) {
    // ---- RGB: MultiGrid.render (base.py:301-331) of the tile's 32*A views from their tile-id maps.  Warp w expands views
    // [32w, 32w+32), one row of V cells (8 pixel rows x V*24 bytes, contiguous in the image) at a time: every lane copies
    // 8-byte pieces of tile rows from the atlas into the chunk buffer -- piece u of the chunk belongs to pixel row u / (3V),
    // cell (u % 3V) / 3 --, then one bulk copy sends the chunk to HBM while the warp fills its other buffer.
    constexpr int PIECES = V * 8 * 3, ITERS = (PIECES + 31) / 32;
    uint32_t src_off[ITERS], cell_sel[ITERS];
#pragma unroll
    for (int i = 0; i < ITERS; ++i) {
      const int u = lane + 32 * i, py = u / (3 * V), rem = u % (3 * V);
      src_off[i] = (uint32_t)(py * 24 + (rem % 3) * 8);
      cell_sel[i] = (uint32_t)(rem / 3);
    }
    const uint32_t atlas_s = smem_u32(s_atlas);
    uint8_t* const bufs = s_out + SM::MAP_BYTES + a * (SM::NBUF * SM::CHUNK);
    const int n_views = n_valid * A;
    constexpr long long VIEW_BYTES = (long long)V * V * 192;
    for (int v = 0; v < 32; ++v) {
      const int vi = a * 32 + v;
      if (vi >= n_views) break;
      uint8_t* const dstv = p.obs + (env0 * A + vi) * VIEW_BYTES;
#pragma unroll 1
      for (int b = 0; b < V; ++b) {
        uint8_t* const buf = bufs + chunk_parity * SM::CHUNK;
        chunk_parity = (chunk_parity + 1) % SM::NBUF;
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(SM::NBUF - 1) : "memory");  // the copy that last read this buffer is done
        __syncwarp();
        const uint2 ids = *reinterpret_cast<const uint2*>(s_out + vi * (V * 8) + b * 8);  // the row's V tile ids
#pragma unroll
        for (int i = 0; i < ITERS; ++i) {
          const int u = lane + 32 * i;
          if (ITERS * 32 == PIECES || u < PIECES) {
            const uint32_t t = __byte_perm(ids.x, ids.y, cell_sel[i]) & 0xFFu;
            uint2 px;
            asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(px.x), "=r"(px.y) : "r"(atlas_s + t * 192u + src_off[i]));
            *reinterpret_cast<uint2*>(buf + 8 * u) = px;
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          bulk_s2g(dstv + b * SM::CHUNK, buf, (uint32_t)SM::CHUNK);
          bulk_commit();
        }
      }
    }
  }
  }  // tile loop
  bulk_wait_read0();  // the CTA's shared memory must outlive the reads
}

// ---------------------------------------------------------------------------------------------
// launcher
// ---------------------------------------------------------------------------------------------
static inline int sm_count(int dev) {
  static int cached[64] = {0};
  if (!cached[dev & 63]) cudaDeviceGetAttribute(&cached[dev & 63], cudaDevAttrMultiProcessorCount, dev);
  return cached[dev & 63];
}

static inline bool pdl_enabled() {  // MG_F2_PDL=0 turns programmatic dependent launch off (experiments)
  static int v = -1;
  if (v < 0) { const char* o = getenv("MG_F2_PDL"); v = (o && atoi(o) == 0) ? 0 : 1; }
  return v != 0;
}

template <int OBS, int V, int A, bool VO0, int NST, bool KS = false, bool HIDE = false>
static int launch_one(const KP& p, cudaStream_t s, int n_steps = 1) {
  using SM = f2::Smem<OBS, V, A, NST>;
  auto k = fused2_kernel<OBS, V, A, VO0, NST, KS, HIDE>;
  const int smem_bytes = SM::TOTAL + (OBS == 2 ? (p.n_tiles + 1) * 192 : 0);  // OBS 2: + the tile atlas and the shadow tile
  if (smem_bytes > 227 * 1024) return MG_E_UNSUPPORTED;
  static int resident[64] = {0}, configured_smem[64] = {0};  // CTAs per SM of this instantiation, per device
  int dev = 0;
  cudaGetDevice(&dev);
  if (!resident[dev & 63] || configured_smem[dev & 63] != smem_bytes) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);  // the kernel lives on shared memory, not on L1
    if (e != cudaSuccess) return (int)e;
    int n = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k, 32 * A, smem_bytes);
    if (e != cudaSuccess) return (int)e;
    if (const char* o = getenv("MG_F2_CTAS_PER_SM")) n = std::min(n, std::max(1, atoi(o)));  // experiments
    resident[dev & 63] = std::max(n, 1);
    configured_smem[dev & 63] = smem_bytes;
  }
  const long long tiles = (p.B + ENVS_PER_CTA - 1) / ENVS_PER_CTA;
  if (tiles <= 0) return 0;
  if (tiles > 0x7FFFFFFF) return MG_E_ARG;
  // as many CTAs as stay resident, trimmed so that every CTA gets the same number of tiles (no ragged last round)
  const long long slots = (long long)resident[dev & 63] * sm_count(dev);
  const long long rounds = (tiles + slots - 1) / slots;
  if (KS && rounds > NST) return MG_E_UNSUPPORTED;  // K steps per launch: every tile's state has to stay in a stage of its CTA
  long long grid = getenv("MG_F2_RAGGED") ? std::min(tiles, slots) : (tiles + rounds - 1) / rounds;
  if (!KS && getenv("MG_F2_ONE_TILE")) grid = tiles;  // experiment: one tile per CTA, the hardware scheduler refills the SMs
  if (getenv("MG_F2_VERBOSE")) fprintf(stderr, "fused2<OBS=%d,V=%d,A=%d,NST=%d,KS=%d>: %d CTAs/SM x %d SMs, %lld tiles in %lld rounds -> grid %lld, %d B shared, %d step(s)\n", OBS, V, A, NST, (int)KS, resident[dev & 63], sm_count(dev), tiles, rounds, grid, smem_bytes, n_steps);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(32 * A); cfg.dynamicSmemBytes = (size_t)smem_bytes; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr; cfg.numAttrs = 1;
  KP q = p;
  static const int prefetch = getenv("MG_F2_PREFETCH") ? atoi(getenv("MG_F2_PREFETCH")) : 1;  // MG_F2_PREFETCH=0: experiments
  q.f2_prefetch = prefetch;
  const cudaError_t e = cudaLaunchKernelEx(&cfg, k, q, (int)tiles, n_steps);
  count_launch();
  return (int)e;
}

template <int OBS, int V, bool VO0>
static int launch_a(const KP& p, cudaStream_t s) {
  // encoded observations: two input stages (the next tile is prefetched); RGB: the image expansion dwarfs everything else,
  // one stage leaves more shared memory for resident CTAs
  constexpr int NS = OBS == 1 ? 2 : 1;
  if (OBS == 1 && VO0 && p.A == 3) {  // experiment: single input stage (more resident CTAs) for the headline shape
    static const int nst1 = getenv("MG_F2_NST") ? atoi(getenv("MG_F2_NST")) : 0;
    if (nst1 == 1) return launch_one<OBS, V, 3, VO0, 1>(p, s);
  }
  switch (p.A) {
    case 1: return launch_one<OBS, V, 1, VO0, NS>(p, s);
    case 2: return launch_one<OBS, V, 2, VO0, NS>(p, s);
    case 3: return launch_one<OBS, V, 3, VO0, NS>(p, s);
    case 4: return launch_one<OBS, V, 4, VO0, NS>(p, s);
    case 5: return launch_one<OBS, V, 5, VO0, NS>(p, s);  // the reference's registry has colours for up to six agents
    case 6: return launch_one<OBS, V, 6, VO0, NS>(p, s);
  }
  return MG_E_UNSUPPORTED;
}

// hide_item_types (encoded observations): the HIDE instantiations, in translation units of their own (mg_fused2_enc{7,5}h.cu)
template <int V, bool VO0>
static int launch_a_hide(const KP& p, cudaStream_t s) {
  switch (p.A) {
    case 1: return launch_one<1, V, 1, VO0, 2, false, true>(p, s);
    case 2: return launch_one<1, V, 2, VO0, 2, false, true>(p, s);
    case 3: return launch_one<1, V, 3, VO0, 2, false, true>(p, s);
    case 4: return launch_one<1, V, 4, VO0, 2, false, true>(p, s);
    case 5: return launch_one<1, V, 5, VO0, 2, false, true>(p, s);
    case 6: return launch_one<1, V, 6, VO0, 2, false, true>(p, s);
  }
  return MG_E_UNSUPPORTED;
}
template <int V>
int launch_fused2_hide(const KP& p, cudaStream_t s) {
  return p.vo == 0 ? launch_a_hide<V, true>(p, s) : launch_a_hide<V, false>(p, s);
}

// one (OBS, V) family per translation unit (mg_fused2_*.cu), so that the instantiations compile in parallel
template <int OBS, int V>
int launch_fused2_ov(const KP& p, cudaStream_t s) {
  return p.vo == 0 ? launch_a<OBS, V, true>(p, s) : launch_a<OBS, V, false>(p, s);
}

// K steps per launch (encoded observations, view_offset 0, A <= 4: the registered MarlGrid-* shapes)
template <int V>
int launch_fused2_ks(const KP& p, int n_steps, cudaStream_t s) {
  if (p.vo != 0 || p.A > f2::POL_MAX_A) return MG_E_UNSUPPORTED;
  if (p.pol_w != nullptr) {  // the policy's weights as MMA B fragments, ordered on the stream before the launch
    constexpr int NW = (V * V * 3 + 3) / 4, KSTEPS = (V * V * 3 + 31) / 32;
    f2::pack_policy_fragments<<<1, 256, 0, s>>>(p.pol_w, p.A, NW, KSTEPS);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    count_launch();
  }
  switch (p.A) {
    case 1: return launch_one<1, V, 1, true, 2, true>(p, s, n_steps);
    case 2: return launch_one<1, V, 2, true, 2, true>(p, s, n_steps);
    case 3: return launch_one<1, V, 3, true, 2, true>(p, s, n_steps);
    case 4: return launch_one<1, V, 4, true, 2, true>(p, s, n_steps);
  }
  return MG_E_UNSUPPORTED;
}

}  // namespace mg

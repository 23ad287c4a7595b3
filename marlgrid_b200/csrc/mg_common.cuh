// mg_common.cuh -- kernel parameter block, derived bit-plane layout and launcher declarations shared by the
// translation units of libmarlgrid_b200.so (sm_100a).  Reference citations are relative to /root/reference.
#pragma once
#include <algorithm>
#include <cstdio>
#include <cstring>

#include "mg_device.cuh"

namespace mg {

constexpr int ENVS_PER_CTA = 32;
// Derived bit-planes `cellbits`, 44 words per env (grids up to 16x16).  Two bits per cell:
//   OP = opaque (Wall / Door that is not open: see_behind() false, objects.py:281-282,330-331)
//   OT = "other": non-empty and NOT a canonical wall (canonical = exactly Wall('worst', state 0), the only wall the
//        reference's generators create).   canonical wall == OP & ~OT,   non-empty == OP | OT.
// stored as LINES, one word each: bits 0..15 = OP along the line, bits 16..31 = OT along the line,
//   word LINE_X0 + x : x-line (cells (x, 0..15), bit = y)       words 0 and 17 stay zero (guard lines: a view
//   word LINE_Y0 + y : y-line (cells (0..15, y), bit = x)       words 18 and 35     row outside the world is empty)
//   word OBJ_WORD0 + k (k < 4): object list, x | y<<4 | type<<8 | colour<<12 | state<<16 | 1<<31 -- up to 4 of the OT objects
//        (Goal, BonusTiles, Keys ...).  An OT cell that is NOT listed is looked up in the byte planes, so the list may be
//        incomplete (more than 4 objects) but never wrong.
//   words 40..43: reserved (zero)
// MEMORY LAYOUT -- tile-transposed, in global AND shared memory: the 32 consecutive envs of a tile interleave their words,
//   word w of env e  at  cellbits[((e >> 5) * BITS_WORDS + w) * 32 + (e & 31)]
// so a tile is one contiguous 5 632-byte chunk (one bulk copy), the lanes of a warp (lane = env) hit 32 different banks
// whatever word each of them indexes (the [env][44] layout of round 1 put lanes l and l+8 on the same bank: 4-way conflicts
// on every view-row load), and a thread-per-env kernel reads / writes global memory coalesced.  Code holds a pointer to an
// env's word 0 (`env_bits`) and addresses word w as bits[w * BS].  The tensor is allocated for whole tiles.
constexpr int BITS_WORDS = 44;
constexpr int BS = 32;  // word stride between consecutive words of one env
constexpr int LINE_X0 = 1, LINE_Y0 = 19, OBJ_WORD0 = 36;
constexpr int OBJ_SLOTS = 4;
__host__ __device__ __forceinline__ long long bits_offset(long long env) { return (env >> 5) * (long long)(BITS_WORDS * BS) + (env & 31); }
template <class T>
__host__ __device__ __forceinline__ T* env_bits(T* cellbits, long long env) { return cellbits + bits_offset(env); }
constexpr uint32_t AF_HEAD = 0x80u;  // derived flag bit: agent is the head of its cell's queue

struct KP {
  int W, H, A, V, vo, ts, max_steps, n_clutter, n_bonus, goal_mode;
  uint32_t flags;
  int S;
  double goal_reward, bonus_reward, bonus_penalty;
  uint8_t agent_color[MG_MAX_AGENTS];
  int spawn_delay[MG_MAX_AGENTS];
  uint8_t kind_of_type[16];
  uint8_t* grid;
  uint8_t* agents;
  int32_t* envrec;
  uint32_t* cellbits;  // [ceil(B/32)][BITS_WORDS][32] (tile-transposed, see below) or nullptr (grid wider/taller than 16, or caller passed none): byte path
  long long B, env_offset;
  unsigned long long seed;
  const int32_t* actions;
  double* rewards;
  uint8_t* done;
  uint8_t* obs;
  const uint8_t* atlas;
  const uint8_t* reset_mask;
  int autoreset;
  int n_tiles;       // atlas tiles (without the appended shadow tile)
  int orient_slots;  // 1: atlas is rotation-equivariant (dir remap), 4: one slot per view orientation
  uint32_t hide;     // MgConfig.hide_types
  uint32_t wall_enc; // MG_T_WALL | MG_C_WORST << 8: the encoded canonical wall, as run-time data (see mg_fused2.cu)
  uint32_t* pregen;  // MgState.pregen: pre-generated next worlds [B][64] (mg_world.cuh) or nullptr
  int ax0, ay0, aw, ah, amax;  // agent spawn box [ax0, ax0 + aw) x [ay0, ay0 + ah) and max_tries (agent_spawn_kwargs, base.py:690-696)
  int scenario;                // MG_SCENARIO_*
  // on-device policy of the persistent rollout (mg_rollout_policy): int8 linear layer over the encoded observation
  const int32_t* pol_w;        // [A][NW][8] words: word i (4 observation bytes) of action k's weight row, NW = ceil(V*V*3 / 4)
  const int32_t* pol_b;        // [A][8]
  int pol_n;                   // number of actions the policy chooses from (1..7)
  uint32_t pol_eps;            // exploration probability * 2^32
  unsigned long long pol_seed;
  int f2_prefetch;             // specialised fused kernel: L2 prefetch of a CTA's first tile ahead of the grid dependency (launch_one sets it)
  double* prestige;            // MgState.prestige [B][A] or nullptr
  uint32_t prestige_mask, prestige_neg;  // agents coloured 'prestige' / with allow_negative_prestige
  double pbeta[MG_MAX_AGENTS], pscale[MG_MAX_AGENTS];
  unsigned long long* stats;  // per-device counters: [0] envs regenerated from a pre-generated world, [1] envs generated inside the step kernel
};

__device__ __forceinline__ bool cell_opaque(int type, int state) {  // objects.py:281-282,330-331
  return type == MG_T_WALL || (type == MG_T_DOOR && state != MG_DOOR_OPEN);
}
__device__ __forceinline__ bool cell_canon(int type, int colour, int state) {
  return type == MG_T_WALL && colour == MG_C_WORST && state == 0;
}
__device__ __forceinline__ uint32_t obj_entry(int x, int y, int type, int colour, int state) {
  return (uint32_t)x | ((uint32_t)y << 4) | ((uint32_t)type << 8) | ((uint32_t)(colour & 15) << 12) | ((uint32_t)(state & 255) << 16) | 0x80000000u;
}
__device__ __forceinline__ uint32_t obj_lookup(const uint32_t* bits, int x, int y) {
  const uint32_t key = 0x80000000u | (uint32_t)x | ((uint32_t)y << 4);
  uint32_t e = 0;
#pragma unroll
  for (int k = 0; k < OBJ_SLOTS; ++k) {
    const uint32_t w = bits[(OBJ_WORD0 + k) * BS];
    if ((w & 0x800000FFu) == key) e = w;
  }
  return e;
}
__device__ __forceinline__ void obj_update(uint32_t* bits, int x, int y, int type, int colour, int state) {
  const uint32_t key = 0x80000000u | (uint32_t)x | ((uint32_t)y << 4);
  const bool listable = type != MG_T_EMPTY && !cell_canon(type, colour, state) && colour < 16 && type < 16;
  int slot = -1;
  for (int k = 0; k < OBJ_SLOTS; ++k) {
    const uint32_t w = bits[(OBJ_WORD0 + k) * BS];
    if ((w & 0x800000FFu) == key) { bits[(OBJ_WORD0 + k) * BS] = 0u; if (slot < 0) slot = k; }
    else if (!(w >> 31) && slot < 0) slot = k;
  }
  if (listable && slot >= 0) bits[(OBJ_WORD0 + slot) * BS] = obj_entry(x, y, type, colour, state);
}
__device__ __forceinline__ void bits_update_cell(uint32_t* bits, int x, int y, int type, int colour, int state) {
  if (bits == nullptr) return;
  const uint32_t op = cell_opaque(type, state) ? 1u : 0u, ot = (type != MG_T_EMPTY && !cell_canon(type, colour, state)) ? 1u : 0u;
  bits[(LINE_X0 + x) * BS] = (bits[(LINE_X0 + x) * BS] & ~(0x10001u << y)) | (op << y) | (ot << (16 + y));
  bits[(LINE_Y0 + y) * BS] = (bits[(LINE_Y0 + y) * BS] & ~(0x10001u << x)) | (op << x) | (ot << (16 + x));
  obj_update(bits, x, y, type, colour, state);
}
// rebuild all words from the byte planes
__device__ inline void bits_rebuild(const uint8_t* tp, uint32_t* bits, int W, int H, int S) {
  if (bits == nullptr) return;
  for (int i = 0; i < BITS_WORDS; ++i) bits[i * BS] = 0u;
  for (int x = 0; x < W; ++x)
    for (int y = 0; y < H; ++y) {
      const int idx = x * H + y;
      const int t = tp[idx];
      if (t != MG_T_EMPTY) bits_update_cell(bits, x, y, t, tp[S + idx], tp[2 * S + idx]);
    }
}

// (type | colour<<8 | state<<16) of the static object at (x, y): bit-planes, then the object list, then -- for
// objects that did not fit the list -- the byte planes
__device__ __forceinline__ uint32_t cell_triple(const uint32_t* bits, int x, int y, const uint8_t* tp, int H, int S) {
  const uint32_t c = (bits[(LINE_X0 + x) * BS] >> y) & 0x10001u;
  if (c == 0u) return 0u;
  if (c == 1u) return (uint32_t)MG_T_WALL | ((uint32_t)MG_C_WORST << 8);  // opaque and not "other": canonical wall
  const uint32_t e = obj_lookup(bits, x, y);
  if (e) return ((e >> 8) & 0xFu) | (((e >> 12) & 0xFu) << 8) | (((e >> 16) & 0xFFu) << 16);
  const int idx = x * H + y;
  return (uint32_t)tp[idx] | ((uint32_t)tp[S + idx] << 8) | ((uint32_t)tp[2 * S + idx] << 16);
}

// ---- launchers (one translation unit per kernel family, so the library builds in parallel) ----
void count_launch();
int launch_env(int mode, const KP& p, cudaStream_t s);          // mg_env_kernels.cu: 0 step(+auto-reset), 1 reset, 2 sync derived
int launch_init(uint8_t* grid, uint8_t* agents, int32_t* envrec, uint32_t* cellbits, long long B, int A, int S, cudaStream_t s);
int launch_random_actions(int32_t* actions, long long n, int n_actions, unsigned long long seed, unsigned long long counter, cudaStream_t s);
int launch_los(const uint8_t* transparent, uint8_t* mask, long long n, int view_size, int ax, int ay, cudaStream_t s);
int launch_obs(const KP& p, int obs, cudaStream_t s);           // mg_obs_kernels.cu: obs 1 encoded / 2 rgb
int launch_fused(const KP& p, int obs, cudaStream_t s);         // mg_fused_kernels.cu: general one-launch step+observe
bool fused_eligible(const KP& p);
bool fused2_eligible(const KP& p);
constexpr int MG_E_UNSUPPORTED = -100;                          // internal: the specialised kernel has no instantiation for this shape
int launch_fused2(const KP& p, int obs, cudaStream_t s);        // mg_fused2.cu: specialised (compile-time A, V) one-launch step+observe
int launch_fused2_rollout(const KP& p, int n_steps, cudaStream_t s);
int launch_policy(const KP& p, const uint8_t* obs, int32_t* actions_next, cudaStream_t s);  // mg_env_kernels.cu: the linear policy as a kernel of its own
int launch_pregen(const KP& p, cudaStream_t s);                 // mg_pregen.cu: one pass of the background world generator over the family
cudaStream_t pregen_stream();                                   // its low-priority side stream on the current device  // the same kernel playing n_steps steps per launch (per-step output slices)

}  // namespace mg

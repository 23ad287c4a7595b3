// mg_kernels.cu -- batched MarlGrid hot path for B200 (sm_100a): step / reset / egocentric obs kernels.
//
// World state (include/marlgrid_b200.h): byte planes type/colour/state [B][3][S], agent records, env
// records -- plus DERIVED bit-planes `cellbits` [B][48] (grids up to 16x16): per cell one bit each for
// "opaque" (Wall / closed Door), "non-empty" and "canonical wall" (exactly Wall('worst', state 0), the only
// wall the reference's generators ever create), stored row-major AND column-major.  They are a lossless
// index of where things are: what is NOT a canonical wall or empty is looked up in the byte planes.
//
// env.step is two launches on one stream:
//   1. step_kernel -- one THREAD per env: MultiGridEnv.step (base.py:501-649): Philox agent order, per-agent
//      action application, stacking stamps, float64 reward, done, and -- for envs whose episode ended --
//      MultiGridEnv.reset (base.py:402-416) by Philox rejection sampling.  A step touches two cells per
//      agent: it works in place on global memory, answering "what is in that cell" from the bit-planes.
//   2. obs_kernel -- one CTA per 32 consecutive envs, one thread per agent-view: the envs' bit-planes and
//      agent records (two contiguous chunks) are staged into shared memory by bulk-async copies
//      (cp.async.bulk + mbarrier: the TMA engine, SASS UBLKCP); a view's transparency rows are one word
//      load + shifts each, the reference's rotation is a row-order / bit reversal, line of sight is carry
//      propagation on row masks, visible canonical walls are written as constants, the few other visible
//      objects are fetched from the byte planes; the CTA streams its staging tile to HBM as 16-byte stores.
//      Grids larger than 16x16 take the byte path: planes staged by TMA, crop gathered byte by byte.
// See DESIGN.md for the layout, the RNG contract and the roofline accounting.
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstring>

#include "mg_device.cuh"

namespace mg {

constexpr int ENVS_PER_CTA = 32;
constexpr int BITS_WORDS = 52;       // per env: 16 row words, 16 column words, 16 canonical-wall words, 4 object-list words
constexpr int OBJ_SLOTS = 4;
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
  uint32_t* cellbits;  // [B][48] or nullptr (grid wider/taller than 16, or caller passed none): byte path
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
};

// ---------------------------------------------------------------------------------------------
// bit-planes.  word x (0..15): row x, bit y = opaque(x,y), bit 16+y = non-empty(x,y)
//              word 16+y     : column y, bit x = opaque, bit 16+x = non-empty
//              word 32+i     : bit j = canonical wall at (i, j) [row i], bit 16+j = canonical wall at (j, i) [column i]
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool cell_opaque(int type, int state) {  // objects.py:281-282,330-331
  return type == MG_T_WALL || (type == MG_T_DOOR && state != MG_DOOR_OPEN);
}
__device__ __forceinline__ bool cell_canon(int type, int colour, int state) {
  return type == MG_T_WALL && colour == MG_C_WORST && state == 0;
}
//              word 48+k     : object list: up to 4 of the non-wall objects, x | y<<4 | type<<8 | colour<<12 | state<<16 | 1<<31.
//                              A non-empty, non-canonical cell that is NOT listed is looked up in the byte planes, so the
//                              list may be incomplete (more than 4 objects) but never wrong.
__device__ __forceinline__ uint32_t obj_entry(int x, int y, int type, int colour, int state) {
  return (uint32_t)x | ((uint32_t)y << 4) | ((uint32_t)type << 8) | ((uint32_t)(colour & 15) << 12) | ((uint32_t)(state & 255) << 16) | 0x80000000u;
}
__device__ __forceinline__ uint32_t obj_lookup(const uint32_t* bits, int x, int y) {
  const uint32_t key = 0x80000000u | (uint32_t)x | ((uint32_t)y << 4);
  uint32_t e = 0;
#pragma unroll
  for (int k = 0; k < OBJ_SLOTS; ++k) {
    const uint32_t w = bits[48 + k];
    if ((w & 0x800000FFu) == key) e = w;
  }
  return e;
}
__device__ __forceinline__ void obj_update(uint32_t* bits, int x, int y, int type, int colour, int state) {
  const uint32_t key = 0x80000000u | (uint32_t)x | ((uint32_t)y << 4);
  const bool listable = type != MG_T_EMPTY && !(type == MG_T_WALL && colour == MG_C_WORST && state == 0) && colour < 16;
  int slot = -1;
  for (int k = 0; k < OBJ_SLOTS; ++k) {
    const uint32_t w = bits[48 + k];
    if ((w & 0x800000FFu) == key) { bits[48 + k] = 0u; if (slot < 0) slot = k; }
    else if (!(w >> 31) && slot < 0) slot = k;
  }
  if (listable && slot >= 0) bits[48 + slot] = obj_entry(x, y, type, colour, state);
}
__device__ __forceinline__ void bits_update_cell(uint32_t* bits, int x, int y, int type, int colour, int state) {
  if (bits == nullptr) return;
  const uint32_t op = cell_opaque(type, state) ? 1u : 0u, ne = type != MG_T_EMPTY ? 1u : 0u, cn = cell_canon(type, colour, state) ? 1u : 0u;
  bits[x] = (bits[x] & ~((1u << y) | (1u << (16 + y)))) | (op << y) | (ne << (16 + y));
  bits[16 + y] = (bits[16 + y] & ~((1u << x) | (1u << (16 + x)))) | (op << x) | (ne << (16 + x));
  bits[32 + x] = (bits[32 + x] & ~(1u << y)) | (cn << y);
  bits[32 + y] = (bits[32 + y] & ~(1u << (16 + x))) | (cn << (16 + x));
  obj_update(bits, x, y, type, colour, state);
}
// rebuild all words from the byte planes
__device__ void bits_rebuild(const uint8_t* tp, uint32_t* bits, int W, int H, int S) {
  if (bits == nullptr) return;
  for (int i = 0; i < BITS_WORDS; ++i) bits[i] = 0u;
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
  if (!((bits[x] >> (16 + y)) & 1u)) return 0u;
  if ((bits[32 + x] >> y) & 1u) return (uint32_t)MG_T_WALL | ((uint32_t)MG_C_WORST << 8);
  const uint32_t e = obj_lookup(bits, x, y);
  if (e) return ((e >> 8) & 0xFu) | (((e >> 12) & 0xFu) << 8) | (((e >> 16) & 0xFFu) << 16);
  const int idx = x * H + y;
  return (uint32_t)tp[idx] | ((uint32_t)tp[S + idx] << 8) | ((uint32_t)tp[2 * S + idx] << 16);
}

// ---------------------------------------------------------------------------------------------
// per-env game state while a thread runs step()/reset(): agent records transposed in shared memory, word w
// of agent a at rec[(a*4+w)*RS] (RS = threads per CTA) -> bank == thread, conflict-free for any per-thread a.
//   w0 = x | y<<8 | dir<<16 | flags<<24     w1 = carry_type | carry_colour<<8 | carry_state<<16 | bonus<<24
//   w2 = stamp                               w3 = scratch (front-cell prefetch)
// `tp` = the env's type plane in global memory (colour at +S, state at +2S); `bits` = its 48 bit-plane words.
// ---------------------------------------------------------------------------------------------
template <int RS>
struct EnvCtx {
  const KP& p;
  uint32_t* rec;
  uint8_t* tp;
  uint32_t* bits;
  uint32_t* scratch;  // reset only: 64 transposed words (wall / other-object masks, by row and by column)
  int sc, ep, tl;     // step_count, episode, lifetime steps
  uint32_t w3;        // lo16 next stamp, hi16 error bits
  bool dirty;         // planes modified during this step
  __device__ __forceinline__ uint32_t& R(int a, int w) { return rec[(a * 4 + w) * RS]; }
  __device__ __forceinline__ void add_err(uint32_t bits_) { w3 |= bits_ << 16; }
  __device__ __forceinline__ uint32_t next_stamp() {
    const uint32_t s = w3 & 0xFFFFu;
    w3 = (w3 & 0xFFFF0000u) | ((s + 1u) & 0xFFFFu);
    return s;
  }
  // static object at (x, y) as type | colour<<8 | state<<16
  __device__ __forceinline__ uint32_t static_cell(int x, int y) {
    if (bits != nullptr) return cell_triple(bits, x, y, tp, p.H, p.S);
    const int idx = x * p.H + y;
    const uint32_t t = tp[idx];
    return t == 0u ? 0u : (t | ((uint32_t)tp[p.S + idx] << 8) | ((uint32_t)tp[2 * p.S + idx] << 16));
  }
  __device__ __forceinline__ int static_type(int x, int y) { return (int)(static_cell(x, y) & 0xFFu); }
  __device__ __forceinline__ void set_cell(int x, int y, int type, int colour, int state) {
    const int idx = x * p.H + y;
    tp[idx] = (uint8_t)type; tp[p.S + idx] = (uint8_t)colour; tp[2 * p.S + idx] = (uint8_t)state;
    bits_update_cell(bits, x, y, type, colour, state);
    dirty = true;
  }
};

// placed agent with the smallest stamp on (x, y), -1 if none: the reference's cell object when it is
// an agent, else `static_obj.agents[0]` (base.py:547-572)
template <int RS>
__device__ __forceinline__ int queue_head(EnvCtx<RS>& c, int x, int y) {
  int best = -1;
  uint32_t bs = 0;
  const uint32_t key = (uint32_t)x | ((uint32_t)y << 8);
  for (int a = 0; a < c.p.A; ++a) {
    const uint32_t w0 = c.R(a, 0);
    if (((w0 >> 24) & MG_AF_PLACED) && (w0 & 0xFFFFu) == key) {
      const uint32_t s = c.R(a, 2);
      if (best < 0 || s < bs) { best = a; bs = s; }
    }
  }
  return best;
}

template <int RS>
__device__ __forceinline__ void put_agent(EnvCtx<RS>& c, int agent, int x, int y) {
  const uint32_t w0 = c.R(agent, 0);
  c.R(agent, 0) = (w0 & 0xFFFF0000u) | (uint32_t)x | ((uint32_t)y << 8) | ((uint32_t)MG_AF_PLACED << 24);
  c.R(agent, 2) = c.next_stamp();
}

// base.py:664-688 try_place_obj for an AGENT in the live world (spawn delay / respawn inside step)
template <int RS>
__device__ __forceinline__ bool try_place_agent(EnvCtx<RS>& c, int x, int y, int agent) {
  const uint32_t cell = c.static_cell(x, y);
  const int st = (int)(cell & 0xFFu);
  if (st != MG_T_EMPTY && !can_overlap_static(st, (int)(cell >> 16))) return false;  // base.py:678-679
  if (!(c.p.flags & MG_F_GHOST) && queue_head(c, x, y) >= 0) return false;                          // base.py:683-684
  put_agent(c, agent, x, y);
  return true;
}

// base.py:690-708 place_obj(top=(0,0), size=None) for an agent in the live world
template <int RS>
__device__ __forceinline__ void place_agent(EnvCtx<RS>& c, Draws& d, int agent) {
  for (int t = 0; t < 100000; ++t) {
    int x, y;
    d.next(c.p.W, c.p.H, x, y);
    if (try_place_agent(c, x, y, agent)) return;
  }
  c.add_err(MG_ERR_PLACEMENT);  // RecursionError base.py:706
}

__device__ __forceinline__ Draws make_draws(const KP& p, unsigned long long g, uint32_t c2, uint32_t tag) {
  Draws d;
  d.g_lo = (uint32_t)g; d.g_hi = (uint32_t)(g >> 32); d.c2 = c2; d.tag = tag;
  d.k0 = (uint32_t)p.seed; d.k1 = (uint32_t)(p.seed >> 32); d.k = 0;
  d.r = U4{0, 0, 0, 0};
  return d;
}

// base.py:402-416 reset + _gen_grid (empty.py:9-16, cluttered.py:25-36, goalcycle.py:30-51), written straight to
// the global planes.  A fresh world only ever holds canonical walls, a Goal and BonusTiles, so with BITS the
// rejection sampling (base.py:690-708) runs on row/column mask sets kept in shared memory (walls / overlappable
// others) and never reads a plane; the bit-plane words are those masks.
// The placement list (goal?, bonus tiles, clutter walls, agents) is walked by ONE loop over the try index k, so
// that all lanes of a warp draw their Philox block on the same iteration (two tries per block).
template <int RS, bool BITS>
__device__ void env_reset(EnvCtx<RS>& c, unsigned long long g) {
  const KP& p = c.p;
  const int W = p.W, H = p.H, S = p.S, A = p.A;
  for (int a = 0; a < A; ++a) {  // agents.py:161-170 (dir survives)
    c.R(a, 0) = c.R(a, 0) & 0x00FF0000u;
    c.R(a, 1) = 0xFF000000u;
    c.R(a, 2) = 0;
  }
  int4* z = reinterpret_cast<int4*>(c.tp);
  for (int i = 0; i < 3 * S / 16; ++i) z[i] = make_int4(0, 0, 0, 0);
  c.w3 &= 0xFFFF0000u;
  uint32_t* wall = c.scratch;                // wall[x*RS]: bit y = canonical wall at (x, y)
  uint32_t* other = c.scratch + 16 * RS;     // other[x*RS]: bit y = Goal / BonusTile (both can_overlap)
  uint32_t* wallc = c.scratch + 32 * RS;     // the same two, column-major: wallc[y*RS] bit x
  uint32_t* otherc = c.scratch + 48 * RS;
  if (BITS) {
    const uint32_t fullr = (1u << H) - 1u, endsr = 1u | (1u << (H - 1)), fullc = (1u << W) - 1u, endsc = 1u | (1u << (W - 1));
    for (int i = 0; i < 16; ++i) {
      wall[i * RS] = (i == 0 || i == W - 1) ? fullr : (i < W ? endsr : 0u);
      wallc[i * RS] = (i == 0 || i == H - 1) ? fullc : (i < H ? endsc : 0u);
      other[i * RS] = 0u; otherc[i * RS] = 0u;
    }
    for (int k = 0; k < OBJ_SLOTS; ++k) c.bits[48 + k] = 0u;
  }
  for (int i = 0; i < W; ++i) {  // wall_rect base.py:172-176
    c.tp[i * H] = MG_T_WALL; c.tp[S + i * H] = MG_C_WORST;
    c.tp[i * H + H - 1] = MG_T_WALL; c.tp[S + i * H + H - 1] = MG_C_WORST;
  }
  for (int j = 0; j < H; ++j) {
    c.tp[j] = MG_T_WALL; c.tp[S + j] = MG_C_WORST;
    c.tp[(W - 1) * H + j] = MG_T_WALL; c.tp[S + (W - 1) * H + j] = MG_C_WORST;
  }
  int n_listed = 0;
  auto put_static = [&](int x, int y, int type, int colour, int state) {
    const int idx = x * H + y;
    c.tp[idx] = (uint8_t)type; c.tp[S + idx] = (uint8_t)colour; c.tp[2 * S + idx] = (uint8_t)state;
    if (BITS) {
      if (type == MG_T_WALL) { wall[x * RS] |= 1u << y; wallc[y * RS] |= 1u << x; }
      else {
        other[x * RS] |= 1u << y; otherc[y * RS] |= 1u << x;
        if (n_listed < OBJ_SLOTS) c.bits[48 + n_listed++] = obj_entry(x, y, type, colour, state);  // Goal / BonusTiles
      }
    }
  };
  if (p.goal_mode == MG_GOAL_FIXED) put_static(W - 2, H - 2, MG_T_GOAL, MG_C_GREEN, 0);  // put_obj base.py:655-662
  // placement list: [random goal] (cluttered.py:28-29), bonus tiles (goalcycle.py:34-46), clutter walls (cluttered.py:32-33), max_tries 100 each;
  // then the agents with spawn_delay 0 (base.py:409-412), max_tries 1e5
  const int n_goal = (p.goal_mode == MG_GOAL_RANDOM) ? 1 : 0;
  const int n_bonus = p.n_bonus;
  const int first_agent = n_goal + n_bonus + p.n_clutter, n_obj = first_agent + A;
  const bool ghost = (p.flags & MG_F_GHOST) != 0;
  uint32_t delayed = 0;  // agents that spawn later (agents.py:34): everything the loop needs lives in registers
  for (int a = 0; a < A; ++a) delayed |= (p.spawn_delay[a] != 0 ? 1u : 0u) << a;
  int obj = 0, tries = 0;
  Draws d = make_draws(p, g, (uint32_t)c.ep, TAG_RESET);
  while (obj < n_obj) {  // base.py:690-708 place_obj / :664-688 try_place_obj, one try per iteration
    const int agent = obj - first_agent;
    if (agent >= 0 && ((delayed >> agent) & 1u)) { ++obj; continue; }
    int x, y;
    d.next(W, H, x, y);
    int st;  // 0 empty, WALL, or GOAL standing for "overlappable other"
    if (BITS) st = ((wall[x * RS] >> y) & 1u) ? (int)MG_T_WALL : (((other[x * RS] >> y) & 1u) ? (int)MG_T_GOAL : (int)MG_T_EMPTY);
    else st = c.tp[x * H + y];
    bool ok;
    if (agent < 0) ok = (st == MG_T_EMPTY);  // statics are placed before any agent: empty cell <=> grid_obj is None
    else {
      const bool overlap = (st == MG_T_EMPTY) || (BITS ? st != MG_T_WALL : can_overlap_static(st, c.tp[2 * S + x * H + y]));
      ok = overlap && (ghost || queue_head(c, x, y) < 0);
    }
    if (ok) {
      if (agent >= 0) { put_agent(c, agent, x, y); c.R(agent, 0) |= (uint32_t)MG_AF_ACTIVE << 24; }
      else if (obj < n_goal) put_static(x, y, MG_T_GOAL, MG_C_GREEN, 0);
      else if (obj < n_goal + n_bonus) put_static(x, y, MG_T_BONUS, MG_C_YELLOW, obj - n_goal);
      else put_static(x, y, MG_T_WALL, MG_C_WORST, 0);
      ++obj; tries = 0;
    } else if (++tries >= (agent >= 0 ? 100000 : 100)) {
      c.add_err(MG_ERR_PLACEMENT);  // RecursionError base.py:706
      if (agent >= 0) c.R(agent, 0) |= (uint32_t)MG_AF_ACTIVE << 24;  // the reference would have raised before activate()
      ++obj; tries = 0;
    }
  }
  c.sc = 0;
  c.ep += 1;
  if (BITS) {  // the masks ARE the bit-plane words
    uint32_t* bits = c.bits;
    for (int i = 0; i < 16; ++i) {
      const uint32_t wl = wall[i * RS], ot = other[i * RS], wc = wallc[i * RS], oc = otherc[i * RS];
      bits[i] = wl | ((wl | ot) << 16);
      bits[16 + i] = wc | ((wc | oc) << 16);
      bits[32 + i] = wl | (wc << 16);
    }
  }
}

// BonusTile.get_reward objects.py:180-206
template <int RS>
__device__ __forceinline__ double bonus_get_reward(EnvCtx<RS>& c, int a, int bonus_id) {
  const KP& p = c.p;
  const int n = p.n_bonus;
  uint32_t w1 = c.R(a, 1);
  int bs = (int)(w1 >> 24);
  bool first = false;
  const double pen = p.bonus_penalty < 0 ? p.bonus_penalty : -p.bonus_penalty;
  double rew;
  if (bs == 0xFF) { bs = ((bonus_id - 1) % n + n) % n; first = true; }
  if (bs == bonus_id) rew = pen;
  else if ((bs + 1) % n == bonus_id) { bs = bonus_id; rew = p.bonus_reward; }
  else rew = pen;
  if (p.flags & MG_F_BONUS_RESET) bs = bonus_id;
  c.R(a, 1) = (w1 & 0x00FFFFFFu) | ((uint32_t)bs << 24);
  if (first && !(p.flags & MG_F_BONUS_INITIAL)) return 0.0;
  return rew;
}

// permutation number idx in [0, A!) -> processing order, nibble q of the result = order[q]
// (Fisher-Yates / Lehmer decode of the contract, oracle/philox.py shuffle_perm)
__constant__ uint32_t RECIP32[9] = {0u, 0u, 0x80000000u, 0x55555556u, 0x40000000u, 0x33333334u, 0x2AAAAAABu, 0x24924925u, 0x20000000u};  // ceil(2^32/n)
__device__ __forceinline__ uint32_t decode_order(uint32_t pidx, int A) {
  uint32_t order = 0x76543210u;
  for (int i = A - 1; i >= 1; --i) {
    const uint32_t n = (uint32_t)(i + 1);
    const uint32_t qd = __umulhi(pidx, RECIP32[n]);  // exact quotient: pidx < 8! = 40320, n <= 8
    const uint32_t j = pidx - qd * n;
    pidx = qd;
    const uint32_t ni = (order >> (4 * i)) & 0xFu, nj = (order >> (4 * j)) & 0xFu;
    order = (order & ~(0xFu << (4 * i)) & ~(0xFu << (4 * j))) | (nj << (4 * i)) | (ni << (4 * j));
  }
  return order;
}

// front-cell word of an agent: type of the cell it faces | its state << 8 | type of the cell it stands on << 16
// | state of that cell << 24 (only the low 7 bits matter: Door states)
template <int RS>
__device__ __forceinline__ uint32_t front_cells(EnvCtx<RS>& c, int cx, int cy, int fx, int fy, bool inb) {
  uint32_t pf = 0;
  if (inb) {
    const uint32_t f = c.static_cell(fx, fy);
    pf = (f & 0xFFu) | (((f >> 16) & 0xFFu) << 8);
  }
  const uint32_t u = c.static_cell(cx, cy);
  return pf | ((u & 0xFFu) << 16);
}

// base.py:501-649 step without the obs; returns done
template <int RS, bool BITS, int AMAX>
__device__ bool env_step(EnvCtx<RS>& c, unsigned long long g, const int32_t* __restrict__ act, double* __restrict__ rew) {
  const KP& p = c.p;
  const int W = p.W, H = p.H, A = p.A, S = p.S;
  const uint32_t t_life = (uint32_t)c.tl;
  Draws d = make_draws(p, g, t_life, TAG_INSTEP);
  for (int a = 0; a < A; ++a) {  // base.py:503-506
    const uint32_t fl = c.R(a, 0) >> 24;
    if (!(fl & MG_AF_ACTIVE) && !(fl & MG_AF_DONE) && c.sc >= p.spawn_delay[a]) {
      place_agent(c, d, a);
      c.R(a, 0) |= (uint32_t)MG_AF_ACTIVE << 24;
    }
  }
  c.sc += 1;  // base.py:512
  // Look up, for every agent at once, the cells its action can touch, together with its action: all loads of
  // this block are unconditional and independent (the bit-plane words of the row in front of / under each
  // agent), so they overlap into ONE memory round trip instead of one per agent and cell.  An agent's own
  // pos/dir only change when it is processed, so the addresses are final; the lookup is redone below if an
  // earlier agent of this step edited the planes (pickup / drop / toggle).  w3 = front word | action << 24.
  {
    int act_r[AMAX];
#pragma unroll
    for (int a = 0; a < AMAX; ++a) act_r[a] = (a < A) ? act[a] : 0;
#pragma unroll
    for (int a = 0; a < AMAX; ++a) {
      if (a < A) {
        const uint32_t w0 = c.R(a, 0);
        const int action = act_r[a];
        uint32_t pf = 0;
        if ((w0 >> 24) & MG_AF_ACTIVE) {
          const int cx = (int)(w0 & 0xFFu), cy = (int)((w0 >> 8) & 0xFFu), dir = (int)((w0 >> 16) & 3u);
          const int fx = cx + ((dir == 0) ? 1 : (dir == 2) ? -1 : 0), fy = cy + ((dir == 1) ? 1 : (dir == 3) ? -1 : 0);
          pf = front_cells(c, cx, cy, fx, fy, (unsigned)fx < (unsigned)W && (unsigned)fy < (unsigned)H);
        }
        c.R(a, 3) = (pf & 0x00FFFFFFu) | ((uint32_t)min(max(action, 0), 255) << 24) | ((action < 0) ? 0xFF000000u : 0u);
      }
    }
  }
  // base.py:514-516: one Philox word -> index of the permutation
  uint32_t fact = 1;
  for (int i = 2; i <= A; ++i) fact *= (uint32_t)i;
  const U4 r = philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), t_life, 0u, (uint32_t)p.seed, (uint32_t)(p.seed >> 32));
  const uint32_t order = decode_order(__umulhi(r.x, fact), A);
  c.tl += 1;
  for (int q = 0; q < A; ++q) {
    const int a = (int)((order >> (4 * q)) & 0xFu);
    uint32_t pf = c.R(a, 3);
    const int action = (int)(pf >> 24);  // out-of-range actions were clamped to 255: still invalid
    double reward = 0.0;
    uint32_t w0 = c.R(a, 0);
    if ((w0 >> 24) & MG_AF_ACTIVE) {  // base.py:521
      const int cx = (int)(w0 & 0xFFu), cy = (int)((w0 >> 8) & 0xFFu), dir = (int)((w0 >> 16) & 3u);
      if (action == MG_A_LEFT) {  // base.py:530-531
        c.R(a, 0) = (w0 & 0xFF00FFFFu) | ((uint32_t)((dir + 3) & 3) << 16);
      } else if (action == MG_A_RIGHT) {  // base.py:534-535
        c.R(a, 0) = (w0 & 0xFF00FFFFu) | ((uint32_t)((dir + 1) & 3) << 16);
      } else if (action >= MG_A_FORWARD && action <= MG_A_TOGGLE) {
        const int fx = cx + ((dir == 0) ? 1 : (dir == 2) ? -1 : 0);  // agents.py:183
        const int fy = cy + ((dir == 1) ? 1 : (dir == 3) ? -1 : 0);
        const bool inb = (unsigned)fx < (unsigned)W && (unsigned)fy < (unsigned)H;
        if (c.dirty) pf = front_cells(c, cx, cy, fx, fy, inb);
        const int ftype = inb ? (int)(pf & 0xFFu) : (int)MG_T_WALL;
        if (!inb) c.add_err(MG_ERR_STACK);  // grid.get asserts in-bounds (base.py:154-156); never hit with wall_rect
        if (action == MG_A_FORWARD) {  // base.py:538-585
          const int fstate = (int)((pf >> 8) & 0xFFu);
          bool can_move = (ftype == MG_T_EMPTY) || can_overlap_static(ftype, fstate);
          if (!(p.flags & MG_F_GHOST) && ftype == MG_T_EMPTY && queue_head(c, fx, fy) >= 0) can_move = false;  // fwd_cell is a GridAgent
          if (can_move) {
            const int ctype = (int)((pf >> 16) & 0xFFu);
            if (ctype != MG_T_EMPTY && !can_overlap_static(ctype, (int)(c.static_cell(cx, cy) >> 16))) c.add_err(MG_ERR_STACK);  // base.py:558
            w0 = (w0 & 0xFFFF0000u) | (uint32_t)fx | ((uint32_t)fy << 8);
            c.R(a, 2) = c.next_stamp();  // appended last to the target cell's queue (base.py:547-552)
            if (ftype == MG_T_GOAL || ftype == MG_T_BONUS) {  // hasattr(fwd_cell, 'get_reward') base.py:576
              double rwd = (ftype == MG_T_GOAL) ? p.goal_reward : bonus_get_reward(c, a, fstate);
              if (p.flags & MG_F_REWARD_DECAY) {  // base.py:579, every operation rounded on its own
                const double qd = __ddiv_rn((double)c.sc, (double)p.max_steps);
                const double u = __dmul_rn(0.9, qd);
                const double f = __dsub_rn(1.0, u);
                rwd = __dmul_rn(rwd, f);
              }
              reward = __dadd_rn(0.0, rwd);  // step_rewards[agent_no] += rwd (base.py:580): 0.0 + (-0.0) is +0.0
            }
            if (ftype == MG_T_LAVA || ftype == MG_T_GOAL) w0 |= (uint32_t)MG_AF_DONE << 24;  // base.py:584-585
            c.R(a, 0) = w0;
          }
        } else if (action == MG_A_PICKUP) {  // base.py:590-597
          const uint32_t w1 = c.R(a, 1);
          if (ftype != MG_T_EMPTY && ((PICKUP_MASK >> ftype) & 1u) && (w1 & 0xFFu) == 0u) {
            const uint32_t cell = c.static_cell(fx, fy);
            c.R(a, 1) = (w1 & 0xFF000000u) | (cell & 0x00FFFFFFu);
            c.set_cell(fx, fy, 0, 0, 0);
          }
        } else if (action == MG_A_DROP) {  // base.py:600-606
          const uint32_t w1 = c.R(a, 1);
          if (inb && ftype == MG_T_EMPTY && (w1 & 0xFFu) != 0u && queue_head(c, fx, fy) < 0) {
            c.set_cell(fx, fy, (int)(w1 & 0xFFu), (int)((w1 >> 8) & 0xFFu), (int)((w1 >> 16) & 0xFFu));
            c.R(a, 1) = w1 & 0xFF000000u;
          }
        } else {  // MG_A_TOGGLE base.py:609-613, Door.toggle objects.py:333-346
          if (ftype == MG_T_DOOR) {
            const uint32_t w1 = c.R(a, 1);
            const int fstate = (int)((pf >> 8) & 0xFFu), fcol = (int)((c.static_cell(fx, fy) >> 8) & 0xFFu);
            int ns = fstate;
            if (fstate == MG_DOOR_LOCKED) {
              if ((w1 & 0xFFu) == MG_T_KEY && (int)((w1 >> 8) & 0xFFu) == fcol) ns = MG_DOOR_CLOSED;
            } else if (fstate == MG_DOOR_CLOSED) ns = MG_DOOR_OPEN;
            else if (fstate == MG_DOOR_OPEN) ns = MG_DOOR_CLOSED;
            if (ns != fstate) c.set_cell(fx, fy, MG_T_DOOR, fcol, ns);
          } else if (ftype == MG_T_BOX) c.add_err(MG_ERR_TOGGLE);  // Box.toggle(self) objects.py:381
        }
      } else if (action != MG_A_DONE) {
        c.add_err(MG_ERR_BAD_ACTION);  // base.py:619-620
      }
    }
    rew[a] = reward;
  }
  bool all_done = true;
  for (int a = 0; a < A; ++a) {  // base.py:627-646
    uint32_t w0 = c.R(a, 0);
    if ((w0 >> 24) & MG_AF_DONE) {
      if (p.flags & MG_F_RESPAWN) {
        c.R(a, 0) = w0 & 0x00FF0000u;  // agent.reset(new_episode=False) agents.py:161-166
        c.R(a, 1) = c.R(a, 1) & 0xFF000000u;
        place_agent(c, d, a);
        c.R(a, 0) |= (uint32_t)MG_AF_ACTIVE << 24;
        all_done = false;
      } else {
        c.R(a, 0) = w0 & ~((uint32_t)MG_AF_ACTIVE << 24);
      }
    } else all_done = false;
  }
  return (c.sc >= p.max_steps) || all_done;  // base.py:649
}

// queue heads: the flag bit AF_HEAD (record byte +3, bit 7) is DERIVED state kept in the record so the
// observe kernel needs no per-view O(A^2) search; it is recomputed by whoever moves agents.
template <int RS>
__device__ __forceinline__ void mark_heads(EnvCtx<RS>& c) {
  for (int a = 0; a < c.p.A; ++a) {
    const uint32_t w0 = c.R(a, 0);
    bool head = ((w0 >> 24) & MG_AF_PLACED) != 0;
    if (head) {
      const uint32_t s = c.R(a, 2);
      for (int q = 0; q < c.p.A; ++q) {
        const uint32_t v0 = c.R(q, 0);
        if (q != a && ((v0 >> 24) & MG_AF_PLACED) && (v0 & 0xFFFFu) == (w0 & 0xFFFFu) && c.R(q, 2) < s) head = false;
      }
    }
    c.R(a, 0) = head ? (w0 | (AF_HEAD << 24)) : (w0 & ~(AF_HEAD << 24));
  }
}

// ---------------------------------------------------------------------------------------------
// per-env kernels (one THREAD per env, world state in place in global memory)
//   MODE 0: env.step (+ auto-reset of finished envs)   MODE 1: env.reset (mask or all)   MODE 2: sync derived state
// ---------------------------------------------------------------------------------------------
constexpr int ENV_THREADS = 128;

template <int MODE, bool BITS, int AMAX>
__global__ void __launch_bounds__(ENV_THREADS) env_kernel(const __grid_constant__ KP p) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint32_t* s_rec = reinterpret_cast<uint32_t*>(smem);
  uint32_t* s_scr = s_rec + p.A * 4 * ENV_THREADS;
  const long long env = (long long)blockIdx.x * ENV_THREADS + threadIdx.x;
  if (env >= p.B) return;
  if (MODE == 1 && p.reset_mask != nullptr && p.reset_mask[env] == 0) return;
  const int A = p.A;
  EnvCtx<ENV_THREADS> c{p, s_rec + threadIdx.x, p.grid + env * 3 * p.S, BITS ? p.cellbits + env * BITS_WORDS : nullptr,
                        s_scr + threadIdx.x, 0, 0, 0, 0u, false};
  int4* arec = reinterpret_cast<int4*>(p.agents) + env * A;
  const int4 er = reinterpret_cast<const int4*>(p.envrec)[env];
#pragma unroll
  for (int a = 0; a < AMAX; ++a) {
    if (a < A) {
      const int4 r = arec[a];
      c.R(a, 0) = (uint32_t)r.x; c.R(a, 1) = (uint32_t)r.y; c.R(a, 2) = (uint32_t)r.z;
    }
  }
  c.sc = er.x; c.ep = er.y; c.tl = er.z; c.w3 = (uint32_t)er.w;
  const unsigned long long g = (unsigned long long)(p.env_offset + env);
  if (MODE == 0) {
    const bool dn = env_step<ENV_THREADS, BITS, AMAX>(c, g, p.actions + env * A, p.rewards + env * A);
    p.done[env] = dn ? 1 : 0;
    if (dn && p.autoreset) env_reset<ENV_THREADS, BITS>(c, g);
  } else if (MODE == 1) {
    env_reset<ENV_THREADS, BITS>(c, g);
  } else {
    bits_rebuild(c.tp, c.bits, p.W, p.H, p.S);
  }
  mark_heads(c);
  for (int a = 0; a < A; ++a) arec[a] = make_int4((int)c.R(a, 0), (int)c.R(a, 1), (int)c.R(a, 2), 0);
  if (MODE != 2) reinterpret_cast<int4*>(p.envrec)[env] = make_int4(c.sc, c.ep, c.tl, (int)c.w3);
}

// ---------------------------------------------------------------------------------------------
// egocentric view of one agent (thread == view): gen_obs_grid (base.py:418-451)
//
// The VxV crop is described in WORLD orientation along per-thread axes: u walks the axis the agent faces
// along, v the axis across (so that a view row of the reference's rotated grid is a run of v at fixed u).
// The reference's rotation (rotate_grid, base.py:67-80, rot_k = dir+1) then reduces to an optional reversal
// of the row order (dir 0,1) and an optional bit reversal inside rows (dir 1,2):
//   dir 0: view[a][b] = sub[V-1-b][a]      rows flipped
//   dir 1: view[a][b] = sub[V-1-a][V-1-b]  rows flipped, bits reversed (u <-> y, v <-> x)
//   dir 2: view[a][b] = sub[b][V-1-a]      bits reversed
//   dir 3: view[a][b] = sub[a][b]          (u <-> y, v <-> x)
// ---------------------------------------------------------------------------------------------
struct ViewGeom {
  int topX, topY;  // agents.py:237-266 get_view_exts
  int su, sv;      // byte strides of u and v inside a plane
  int u0, v0;      // world coordinate of u = 0 / v = 0 along their axes
  int Lu, Lv;      // axis lengths
  bool vertical, flip, rev;
};

__device__ __forceinline__ ViewGeom view_geom(int px, int py, int dir, int V, int vo, int W, int H) {
  const int h = V / 2;
  ViewGeom g;
  g.topX = (dir == 0) ? px - vo : (dir == 2) ? px - V + 1 + vo : px - h;
  g.topY = (dir == 1) ? py - vo : (dir == 3) ? py - V + 1 + vo : py - h;
  g.vertical = (dir & 1) != 0;
  g.flip = dir < 2;
  g.rev = (dir == 1) || (dir == 2);
  g.su = g.vertical ? 1 : H; g.sv = g.vertical ? H : 1;
  g.u0 = g.vertical ? g.topY : g.topX; g.v0 = g.vertical ? g.topX : g.topY;
  g.Lu = g.vertical ? H : W; g.Lv = g.vertical ? W : H;
  return g;
}

// What the view thread needs after line of sight: visibility / non-empty / canonical-wall masks in VIEW
// orientation, packed with a row stride of 8 bits (bit 8*(b&3) + a of the lo word for rows 0..3, of the hi
// word for rows 4..7), and the plane offset of view cell (a, b): cell_idx = row0 + b*ustep + a*vstep.
struct PackedView {
  uint32_t vis_lo, vis_hi, ne_lo, ne_hi, cw_lo, cw_hi;
  int row0, ustep, vstep;
  __device__ __forceinline__ bool visible(int a, int b) const { return (((b < 4 ? vis_lo : vis_hi) >> (8 * (b & 3) + a)) & 1u) != 0; }
  __device__ __forceinline__ bool nonempty(int a, int b) const { return (((b < 4 ? ne_lo : ne_hi) >> (8 * (b & 3) + a)) & 1u) != 0; }
};

template <int V>
__device__ __forceinline__ void pack_rows(const uint32_t (&r)[V], uint32_t& lo, uint32_t& hi) {
  lo = 0; hi = 0;
#pragma unroll
  for (int b = 0; b < V; ++b) {
    if (b < 4) lo |= r[b] << (8 * b); else hi |= r[b] << (8 * (b - 4));
  }
}

// transparency / non-empty / canonical-wall rows from the bit-planes: two words per view row
template <int V>
__device__ __forceinline__ void rows_from_bits(const uint32_t* __restrict__ bits /* this env's 48 words */, const ViewGeom& g,
                                               uint32_t (&T)[V], uint32_t (&NE)[V], uint32_t (&CW)[V]) {
  constexpr uint32_t RM = (1u << V) - 1u;
  const uint32_t* bp = bits + (g.vertical ? 16 : 0);
  const int sh = g.v0 + 8;                  // >= 1: the 16 board bits are parked at bits 8..23 before shifting right
  const int csh = g.vertical ? 16 : 0;      // canonical walls: column half / row half of word 32+i
#pragma unroll
  for (int b = 0; b < V; ++b) {
    const int idx = g.u0 + (g.flip ? V - 1 - b : b);
    const bool in = (unsigned)idx < (unsigned)g.Lu;  // rows outside the world: empty, transparent
    const uint32_t w = in ? bp[idx] : 0u;
    const uint32_t cwd = in ? bits[32 + idx] : 0u;
    uint32_t opq = (((w & 0xFFFFu) << 8) >> sh) & RM;
    uint32_t ne = (((w >> 16) << 8) >> sh) & RM;
    uint32_t cw = ((((cwd >> csh) & 0xFFFFu) << 8) >> sh) & RM;
    if (g.rev) { opq = rev_bits<V>(opq); ne = rev_bits<V>(ne); cw = rev_bits<V>(cw); }
    T[b] = ~opq & RM;
    NE[b] = ne;
    CW[b] = cw;
  }
}

// the same rows gathered byte by byte from the type plane staged in shared memory (any grid size)
template <int V>
__device__ __forceinline__ void rows_from_planes(const KP& p, const uint8_t* __restrict__ tp, const ViewGeom& g, uint32_t (&T)[V],
                                                 uint32_t (&NE)[V]) {
  constexpr uint32_t RM = (1u << V) - 1u;
  const int S = p.S;
  // clamped per-axis offsets: every load is in range, out-of-world cells are masked afterwards
  // (MultiGrid.slice zero-pads: empty, transparent; base.py:132-141)
  uint32_t valid_u = 0, valid_v = 0;
  int voff[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const int vv = g.v0 + i, uu = g.u0 + i;
    valid_v |= ((unsigned)vv < (unsigned)g.Lv ? 1u : 0u) << i;
    valid_u |= ((unsigned)uu < (unsigned)g.Lu ? 1u : 0u) << i;
    voff[i] = min(max(vv, 0), g.Lv - 1) * g.sv;
  }
  uint32_t Tu[V], NEu[V];
#pragma unroll
  for (int u = 0; u < V; ++u) {
    const uint8_t* rowp = tp + min(max(g.u0 + u, 0), g.Lu - 1) * g.su;
    uint32_t opaque = 0, nonempty = 0, doors = 0;
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const uint32_t t = rowp[voff[v]];
      opaque |= (t == MG_T_WALL ? 1u : 0u) << v;
      nonempty |= (t != MG_T_EMPTY ? 1u : 0u) << v;
      doors |= (t == MG_T_DOOR ? 1u : 0u) << v;
    }
    doors &= valid_v;
    while (doors) {  // objects.py:330-331: a door hides what is behind it unless open (rare)
      const int v = __ffs(doors) - 1;
      doors &= doors - 1;
      if (rowp[2 * S + (g.v0 + v) * g.sv] != MG_DOOR_OPEN) opaque |= 1u << v;  // v is valid: no clamping needed
    }
    const bool urow = (valid_u >> u) & 1u;
    Tu[u] = urow ? (~opaque | ~valid_v) & RM : RM;
    NEu[u] = urow ? (nonempty & valid_v) : 0u;
  }
#pragma unroll
  for (int b = 0; b < V; ++b) {  // world rows -> view rows (rotate_grid as flip / bit reversal)
    uint32_t t = g.flip ? Tu[V - 1 - b] : Tu[b];
    uint32_t n = g.flip ? NEu[V - 1 - b] : NEu[b];
    if (g.rev) { t = rev_bits<V>(t); n = rev_bits<V>(n); }
    T[b] = t; NE[b] = n;
  }
}

template <int V, bool BITS>
__device__ __forceinline__ PackedView view_masks(const KP& p, const uint8_t* __restrict__ tp, const uint32_t* __restrict__ bits,
                                                 const ViewGeom& g) {
  constexpr uint32_t RM = (1u << V) - 1u;
  uint32_t T[V], NE[V], CW[V], M[V];
  if (BITS) rows_from_bits<V>(bits, g, T, NE, CW);
  else {
    rows_from_planes<V>(p, tp, g, T, NE);
#pragma unroll
    for (int b = 0; b < V; ++b) CW[b] = 0u;
  }
  if (p.flags & MG_F_SEE_THROUGH) {  // agents.py:294-295
#pragma unroll
    for (int b = 0; b < V; ++b) M[b] = RM;
  } else {
    occlude_rows<V>(T, V / 2, V - 1 - p.vo, M);  // agents.py:233-234,293
  }
  PackedView pv;
  pack_rows<V>(M, pv.vis_lo, pv.vis_hi);
  pack_rows<V>(NE, pv.ne_lo, pv.ne_hi);
  pack_rows<V>(CW, pv.cw_lo, pv.cw_hi);
  pv.ustep = g.flip ? -g.su : g.su;
  pv.vstep = g.rev ? -g.sv : g.sv;
  pv.row0 = g.topX * p.H + g.topY + (g.flip ? (V - 1) * g.su : 0) + (g.rev ? (V - 1) * g.sv : 0);
  return pv;
}

// world cell (qx, qy) -> view cell (a, b); false if outside the view
template <int V>
__device__ __forceinline__ bool world_to_view(const ViewGeom& g, int qx, int qy, int& a, int& b) {
  const int sx = qx - g.topX, sy = qy - g.topY;
  const int u = g.vertical ? sy : sx, v = g.vertical ? sx : sy;
  if ((unsigned)u >= (unsigned)V || (unsigned)v >= (unsigned)V) return false;
  b = g.flip ? V - 1 - u : u;
  a = g.rev ? V - 1 - v : v;
  return true;
}

// cells of rows [B0, B0+4) selected by m: WorldObj.encode (objects.py:90-99) from the byte planes `tp`
// (shared memory on the byte path, global memory for the rare non-wall objects on the bit-plane path)
template <int V, int B0>
__device__ __forceinline__ void encode_cells(uint32_t m, const PackedView& pv, const uint8_t* __restrict__ tp, int S, uint8_t* __restrict__ out) {
  while (m) {
    const int bit = __ffs(m) - 1;
    m &= m - 1;
    const int va = bit & 7, vb = (bit >> 3) + B0;
    const uint8_t* cp = tp + pv.row0 + vb * pv.ustep + va * pv.vstep;
    uint8_t* o = out + va * (V * 3) + vb * 3;
    o[0] = cp[0]; o[1] = cp[S]; o[2] = cp[2 * S];
  }
}

// visible canonical walls of rows [B0, B0+4): constants (8, 9, 0), no plane access, no loop: every store has a
// compile-time offset into the staging tile
template <int V, int B0>
__device__ __forceinline__ void encode_walls(uint32_t m, uint8_t* __restrict__ out) {
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    if (B0 + r < V) {
#pragma unroll
      for (int a = 0; a < V; ++a) {
        if ((m >> (8 * r + a)) & 1u) {
          out[a * (V * 3) + (B0 + r) * 3 + 0] = MG_T_WALL;
          out[a * (V * 3) + (B0 + r) * 3 + 1] = MG_C_WORST;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// building blocks shared by the observe kernel and the fused step+observe kernel (32 envs per CTA, one thread
// per agent view)
//   OBS : 1 = encoded (MultiGrid.encode base.py:196-214), 2 = RGB tiles (base.py:301-331)
//   TSC : RGB only: 8 = tile size 8 known at compile time (every registered env), 1 = run-time tile size that is a
//         multiple of 4 (tile rows are whole words -> 16-byte stores), 0 = any tile size (byte path)
//   BITS: world described by the bit-planes (W, H <= 16) / by the byte planes staged in shared memory
// ---------------------------------------------------------------------------------------------
template <int V>
struct ObsSmem {
  uint8_t* out;     // OBS 1: staging tile [32*A][V*V*3]
  uint8_t* tile;    // OBS 2: tile-id map [32*A][V*V]
  uint8_t* orient;  // OBS 2: view orientation [32*A]
  uint8_t* atlas;   // OBS 2: atlas copy + one shadow tile
};

template <int V>
__device__ __forceinline__ ObsSmem<V> obs_smem(uint8_t* s_out, int A) {
  ObsSmem<V> o;
  o.out = s_out; o.tile = s_out;
  o.orient = o.tile + ENVS_PER_CTA * A * V * V;
  o.atlas = o.orient + ((ENVS_PER_CTA * A + 15) / 16) * 16;
  return o;
}

// zero the staging tile (invisible / empty cells encode as 0) or copy the tile atlas (+ shadow tile)
template <int OBS, int V>
__device__ __forceinline__ void obs_prepare(const KP& p, const ObsSmem<V>& o, int tid, int nthreads) {
  const int A = p.A;
  if (OBS == 1) {
    int4* z = reinterpret_cast<int4*>(o.out);
    const int n16 = ENVS_PER_CTA * A * V * V * 3 / 16;
    constexpr int ITERS = (V * V * 3 + 15) / 16;  // n16 / (32*A) rounded up: the block has 32*A threads
#pragma unroll
    for (int k = 0; k < ITERS; ++k) {
      const int i = tid + k * nthreads;
      if (i < n16) z[i] = make_int4(0, 0, 0, 0);
    }
  } else {
    const int tile_bytes = p.ts * p.ts * 3;
    const int slots = p.n_tiles * p.orient_slots;
    if ((tile_bytes & 15) == 0) {  // whole 16-byte chunks: vector copy (the atlas pointer is 16-byte aligned, checked on the host)
      const int cpt = tile_bytes / 16;
      const int4* src = reinterpret_cast<const int4*>(p.atlas);
      int4* dst = reinterpret_cast<int4*>(o.atlas);
      for (int i = tid; i < slots * cpt; i += nthreads) {
        const int slot = i / cpt, ch = i - slot * cpt;
        const int tile = slot / p.orient_slots, orient = slot - tile * p.orient_slots;
        dst[i] = __ldg(src + (size_t)(tile * 4 + orient) * cpt + ch);
      }
    } else {
      for (int i = tid; i < slots * tile_bytes; i += nthreads) {
        const int slot = i / tile_bytes, off = i - slot * tile_bytes;
        const int tile = slot / p.orient_slots, orient = slot - tile * p.orient_slots;
        o.atlas[i] = p.atlas[(size_t)(tile * 4 + orient) * tile_bytes + off];
      }
    }
    for (int i = tid; i < tile_bytes; i += nthreads) {  // COLORS['shadow'] objects.py:25, base.py:305
      const int c = i % 3;
      o.atlas[slots * tile_bytes + i] = (c == 0) ? 35 : (c == 1) ? 25 : 30;
    }
  }
}

// one agent view: gen_obs_grid + encode / tile ids.  rec = the env's agent records [q*4 + w] in shared memory,
// tp = the env's byte planes (global memory on the bit-plane path, shared memory on the byte path)
template <int OBS, int V, bool BITS, bool HEADS = false>
__device__ __forceinline__ void obs_view(const KP& p, const ObsSmem<V>& o, int view, int a, long long env, const uint32_t* __restrict__ rec,
                                         const uint8_t* __restrict__ tp, const uint32_t* __restrict__ bits, const uint8_t* __restrict__ heads = nullptr) {
  constexpr int VV = V * V;
  const int A = p.A, S = p.S;
  const uint32_t w0 = rec[a * 4];
  const bool active = ((w0 >> 24) & MG_AF_ACTIVE) != 0;  // base.py:420-425
  const int px = (int)(w0 & 0xFFu), py = (int)((w0 >> 8) & 0xFFu), dir = (int)((w0 >> 16) & 3u);
  const int orient = (3 - dir) & 3;  // view orientation (0 - rot_k) % 4, base.py:130
  if (OBS == 2) {
    o.orient[view] = (uint8_t)((p.orient_slots == 4) ? orient : 0);
    if (!active) {
      const uint8_t shadow = (uint8_t)(p.n_tiles);  // one past the last tile: resolved to the shadow slot when expanding
      for (int i = 0; i < VV; ++i) o.tile[view * VV + i] = shadow;
    }
  }
  if (!active) return;
  const ViewGeom g = view_geom(px, py, dir, V, p.vo, p.W, p.H);
  const PackedView pv = view_masks<V, BITS>(p, tp, bits, g);
  if (OBS == 1) {
    uint8_t* out = o.out + view * (VV * 3);
    if (BITS) {
      encode_walls<V, 0>(pv.vis_lo & pv.cw_lo, out);
      if (V > 4) encode_walls<V, 4>(pv.vis_hi & pv.cw_hi, out);
    }
    uint32_t g_lo = pv.vis_lo & pv.ne_lo & ~pv.cw_lo, g_hi = pv.vis_hi & pv.ne_hi & ~pv.cw_hi;  // visible objects that are not canonical walls
    if (BITS) {
#pragma unroll
      for (int k = 0; k < OBJ_SLOTS; ++k) {  // the object list answers for (almost) all of them without touching the planes
        const uint32_t e = bits[48 + k];
        int va, vb;
        if (!(e >> 31) || !world_to_view<V>(g, (int)(e & 15u), (int)((e >> 4) & 15u), va, vb)) continue;
        const uint32_t bit = 1u << (8 * (vb & 3) + va);
        if (vb < 4) { if (!(g_lo & bit)) continue; g_lo &= ~bit; } else { if (!(g_hi & bit)) continue; g_hi &= ~bit; }
        uint8_t* oo = out + va * (V * 3) + vb * 3;
        oo[0] = (uint8_t)((e >> 8) & 15u); oo[1] = (uint8_t)((e >> 12) & 15u); oo[2] = (uint8_t)((e >> 16) & 255u);
      }
    }
    encode_cells<V, 0>(g_lo, pv, tp, S, out);  // whatever is left (byte path: everything) comes from the planes
    if (V > 4) encode_cells<V, 4>(g_hi, pv, tp, S, out);
    for (int q = 0; q < A; ++q) {  // agents that are their cell's object: (13, colour, dir)
      const uint32_t v0 = rec[q * 4];
      if (HEADS ? !heads[q] : !((v0 >> 24) & AF_HEAD)) continue;
      int va, vb;
      if (!world_to_view<V>(g, (int)(v0 & 0xFFu), (int)((v0 >> 8) & 0xFFu), va, vb)) continue;
      if (!pv.visible(va, vb) || pv.nonempty(va, vb)) continue;
      uint8_t* oo = out + va * (V * 3) + vb * 3;
      oo[0] = MG_T_AGENT; oo[1] = p.agent_color[q]; oo[2] = (uint8_t)((v0 >> 16) & 3u);
    }
  } else {  // OBS == 2: tile ids, render_tile base.py:275-299
    const int per_kind = 1 + 4 * A;
    const uint8_t wall_tile = (uint8_t)(p.kind_of_type[MG_T_WALL] * per_kind);
    uint8_t* tl = o.tile + view * VV;
    uint32_t bad = 0;
#pragma unroll
    for (int b = 0; b < V; ++b) {
      const uint8_t* rowp = tp + pv.row0 + b * pv.ustep;
      const uint32_t visr = ((b < 4 ? pv.vis_lo : pv.vis_hi) >> (8 * (b & 3))) & 0xFFu;
      const uint32_t ner = ((b < 4 ? pv.ne_lo : pv.ne_hi) >> (8 * (b & 3))) & 0xFFu;
      const uint32_t cwr = ((b < 4 ? pv.cw_lo : pv.cw_hi) >> (8 * (b & 3))) & 0xFFu;
#pragma unroll
      for (int va = 0; va < V; ++va) {
        uint8_t t = (uint8_t)p.n_tiles;  // shadow
        if ((visr >> va) & 1u) {
          t = 0;
          if ((cwr >> va) & 1u) t = wall_tile;
          else if ((ner >> va) & 1u) {
            int type;
            if (BITS) {  // object list first, byte plane for objects that did not fit
              const int cidx = pv.row0 + b * pv.ustep + va * pv.vstep;
              const uint32_t e = obj_lookup(bits, cidx / p.H, cidx % p.H);
              type = e ? (int)((e >> 8) & 15u) : (int)rowp[va * pv.vstep];
            } else type = rowp[va * pv.vstep];
            const int kind = p.kind_of_type[type];
            if (kind == 0xFF) bad = 1; else t = (uint8_t)(kind * per_kind);
          }
        }
        tl[b * V + va] = t;
      }
    }
    for (int q = 0; q < A; ++q) {
      const uint32_t v0 = rec[q * 4];
      if (HEADS ? !heads[q] : !((v0 >> 24) & AF_HEAD)) continue;
      int va, vb;
      if (!world_to_view<V>(g, (int)(v0 & 0xFFu), (int)((v0 >> 8) & 0xFFu), va, vb)) continue;
      if (!pv.visible(va, vb)) continue;
      // top_agent if it stands on this cell, else the queue head (base.py:282-293)
      const bool mine = ((v0 ^ w0) & 0xFFFFu) == 0u;
      const int qq = mine ? a : q;
      const int qd = (int)(((mine ? w0 : v0) >> 16) & 3u);
      const int slot_dir = (p.orient_slots == 4) ? qd : ((qd + orient) & 3);
      tl[vb * V + va] = (uint8_t)(tl[vb * V + va] + 1 + 4 * qq + slot_dir);
    }
    if (bad) atomicOr(reinterpret_cast<unsigned int*>(p.envrec) + env * 4 + 3, (unsigned int)MG_ERR_RENDER << 16);
  }
}

// stream the CTA's observations to HBM.  Must be called by all threads after a __syncthreads().
template <int OBS, int V, int TSC>
__device__ __forceinline__ void obs_emit(const KP& p, const ObsSmem<V>& o, long long env0, int n_valid, int tid, int nthreads) {
  constexpr int VV = V * V;
  const int A = p.A;
  if (OBS == 1) {
    const long long total = (long long)n_valid * A * VV * 3;
    uint8_t* dst = p.obs + env0 * A * VV * 3;
    if ((total & 15) == 0) {
      // the whole staging tile is one contiguous, 16-byte aligned run of the output tensor: a single
      // shared->global bulk copy (TMA) moves it; nobody spends an instruction on the 14 KB
      if (tid == 0) {
        fence_proxy_async_smem();
        bulk_s2g(dst, o.out, (uint32_t)total);
        bulk_commit();
      }
    } else {  // ragged last CTA
      const int n16 = (int)(total / 16);
      const int4* src = reinterpret_cast<const int4*>(o.out);
      for (int i = tid; i < n16; i += nthreads) st_stream_v4(reinterpret_cast<int4*>(dst) + i, src[i]);
      for (int i = n16 * 16 + tid; i < total; i += nthreads) dst[i] = o.out[i];
    }
  } else {
    const int ts = p.ts, n_views = n_valid * A;
    const int row_bytes = V * ts * 3;
    const long long view_bytes = (long long)row_bytes * V * ts;
    uint8_t* dst = p.obs + env0 * A * view_bytes;
    if (TSC != 0) {
      // one thread = one 16-byte store; with TSC == 8 every divisor below is a compile-time constant
      const int tsz = (TSC == 8) ? 8 : ts;
      const int wpt = tsz * 3 / 4;        // words per tile row
      const int wpr = V * wpt;            // words per image row
      const int v16 = (V * tsz * 3) * (V * tsz) / 16;
      const uint32_t* atlas_w = reinterpret_cast<const uint32_t*>(o.atlas);
      const int total16 = n_views * v16;
      const int shadow_slot = p.n_tiles * p.orient_slots;
      for (int i = tid; i < total16; i += nthreads) {
        const int view = i / v16, k = i - view * v16;
        const int os = o.orient[view];
        const uint8_t* tl = o.tile + view * VV;
        const int gw0 = 4 * k;
        int y = gw0 / wpr;
        const int xw = gw0 - y * wpr;
        int va = xw / wpt, r = xw - va * wpt;
        int vb = y / tsz, pyy = y - vb * tsz;
        // the four words of a store walk along a tile row and at most once into the next tile (or image row):
        // the tile is looked up again only then
        int t = tl[vb * V + va];
        const uint32_t* trow = atlas_w + (((t >= p.n_tiles) ? shadow_slot : t * p.orient_slots + os) * tsz + pyy) * wpt;
        uint32_t wv[4];
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          wv[w] = trow[r];
          if (++r == wpt && w < 3) {
            r = 0;
            if (++va == V) { va = 0; ++y; vb = y / tsz; pyy = y - vb * tsz; }
            t = tl[vb * V + va];
            trow = atlas_w + (((t >= p.n_tiles) ? shadow_slot : t * p.orient_slots + os) * tsz + pyy) * wpt;
          }
        }
        st_stream_v4(reinterpret_cast<int4*>(dst) + i, make_int4((int)wv[0], (int)wv[1], (int)wv[2], (int)wv[3]));
      }
    } else {
      const long long total = (long long)n_views * view_bytes;
      for (long long i = tid; i < total; i += nthreads) {
        const int view = (int)(i / view_bytes);
        const int k = (int)(i - view * view_bytes);
        const int y = k / row_bytes, xb = k - y * row_bytes;
        const int va = xb / (ts * 3), r = xb - va * ts * 3;
        const int vb = y / ts, pyy = y - vb * ts;
        const int t = o.tile[view * VV + vb * V + va];
        const int slot = (t >= p.n_tiles) ? p.n_tiles * p.orient_slots : t * p.orient_slots + o.orient[view];
        dst[i] = o.atlas[(slot * ts + pyy) * ts * 3 + r];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// observe kernel
// ---------------------------------------------------------------------------------------------
template <int OBS, int V, int TSC, bool BITS>
__global__ void __launch_bounds__(32 * MG_MAX_AGENTS) obs_kernel(const __grid_constant__ KP p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, nthreads = blockDim.x;
  const long long env0 = (long long)blockIdx.x * ENVS_PER_CTA;
  const int n_valid = (int)min((long long)ENVS_PER_CTA, p.B - env0);
  const int A = p.A, S = p.S;

  uint8_t* s_grid = smem;                                                                    // byte path only
  uint32_t* s_bits = reinterpret_cast<uint32_t*>(s_grid + (BITS ? 0 : ENVS_PER_CTA * 3 * S));  // bit-plane path only
  uint32_t* s_rec = s_bits + (BITS ? ENVS_PER_CTA * BITS_WORDS : 0);                         // agent records as stored: [env][a][4 words]
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_rec + ENVS_PER_CTA * A * 4);
  const ObsSmem<V> o = obs_smem<V>(reinterpret_cast<uint8_t*>(s_bar + 2), A);  // 16-byte aligned: every block above is a multiple of 16 bytes

  if (tid == 0) mbar_init(s_bar, 1);
  __syncthreads();
  if (tid == 0) {
    const uint32_t wbytes = BITS ? (uint32_t)n_valid * (BITS_WORDS * 4u) : (uint32_t)n_valid * 3u * (uint32_t)S;
    const uint32_t rbytes = (uint32_t)n_valid * (uint32_t)A * 16u;
    mbar_expect_tx(s_bar, wbytes + rbytes);
    if (BITS) bulk_g2s(s_bits, p.cellbits + env0 * BITS_WORDS, wbytes, s_bar);
    else bulk_g2s(s_grid, p.grid + env0 * 3 * S, wbytes, s_bar);
    bulk_g2s(s_rec, p.agents + env0 * A * 16, rbytes, s_bar);
  }
  obs_prepare<OBS, V>(p, o, tid, nthreads);  // while the copies are in flight
  mbar_wait(s_bar, 0);
  __syncthreads();
  if (tid < n_valid * A) {
    const int le = tid / A, a = tid - le * A;
    obs_view<OBS, V, BITS>(p, o, tid, a, env0 + le, s_rec + le * A * 4, BITS ? p.grid + (env0 + le) * 3 * S : s_grid + le * 3 * S,
                           BITS ? s_bits + le * BITS_WORDS : nullptr);
  }
  __syncthreads();
  obs_emit<OBS, V, TSC>(p, o, env0, n_valid, tid, nthreads);
  if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the CTA (and its shared memory) must outlive the read
}

// ---------------------------------------------------------------------------------------------
// fused env.step + observe kernel (bit-plane worlds, ghost mode, no respawn, no spawn delay): ONE launch per step.
//
// Thread (env, agent) first plays its own agent's action (MultiGridEnv.step, base.py:517-622): in ghost mode an
// action that does not edit the planes depends on nothing another agent does in the same step, so the A agents
// of an env act in parallel and the reference's random processing order (base.py:514-516) only decides the
// arrival stamps of the agents that moved.  Envs where some action WOULD edit the planes (a pickup / drop /
// toggle that takes effect) and envs whose episode just ended are handed to one lane that runs the general
// sequential code (env_step / env_reset above) -- rare, and exact.  Then the same threads observe.
// ---------------------------------------------------------------------------------------------
constexpr uint32_t FL_SLOW = 1u, FL_RESET = 2u, FL_BITS_DIRTY = 4u, FL_NOTDONE = 8u;  // s_flag bits; bits 8..15 movers, 16..31 error bits

// the general sequential code, kept out of line so the common path keeps its registers
__device__ __noinline__ void seq_step(EnvCtx<32>& cref, unsigned long long g, const int32_t* act, double* rew) {
  EnvCtx<32> c = cref;  // work on registers, not through the reference (local memory)
  env_step<32, true, MG_MAX_AGENTS>(c, g, act, rew);
  cref.sc = c.sc; cref.ep = c.ep; cref.tl = c.tl; cref.w3 = c.w3; cref.dirty = c.dirty;
}
__device__ __noinline__ void seq_reset(EnvCtx<32>& cref, unsigned long long g) {
  EnvCtx<32> c = cref;
  env_reset<32, true>(c, g);
  cref.sc = c.sc; cref.ep = c.ep; cref.tl = c.tl; cref.w3 = c.w3; cref.dirty = c.dirty;
}

template <int OBS, int V, int TSC>
__global__ void __launch_bounds__(32 * MG_MAX_AGENTS, 4) fused_kernel(const __grid_constant__ KP p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, nthreads = blockDim.x, lane = tid & 31, warp = tid >> 5;
  const long long env0 = (long long)blockIdx.x * ENVS_PER_CTA;
  const int n_valid = (int)min((long long)ENVS_PER_CTA, p.B - env0);
  const int A = p.A, S = p.S, W = p.W, H = p.H;

  uint32_t* s_bits = reinterpret_cast<uint32_t*>(smem);
  uint32_t* s_rec = s_bits + ENVS_PER_CTA * BITS_WORDS;                       // [env][a][4]
  int32_t* s_env = reinterpret_cast<int32_t*>(s_rec + ENVS_PER_CTA * A * 4);  // [env][4]
  uint32_t* s_flag = reinterpret_cast<uint32_t*>(s_env + ENVS_PER_CTA * 4);   // [env]
  uint32_t* s_order = s_flag + ENVS_PER_CTA;                                  // [env] processing order, nibble q = agent
  uint8_t* s_head = reinterpret_cast<uint8_t*>(s_order + ENVS_PER_CTA);       // [32*A] queue-head flag per agent (padded to 256)
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_head + 32 * MG_MAX_AGENTS);
  uint8_t* s_out = reinterpret_cast<uint8_t*>(s_bar + 2);
  const ObsSmem<V> o = obs_smem<V>(s_out, A);
  // scratch of the sequential path, aliased with the output area (which is re-zeroed if it was used)
  uint32_t* s_trec = reinterpret_cast<uint32_t*>(s_out);  // [A*4][32] transposed records
  uint32_t* s_scr = s_trec + A * 4 * 32;                  // [64][32] reset row / column masks

  if (tid == 0) mbar_init(s_bar, 1);
  if (tid < ENVS_PER_CTA) s_flag[tid] = 0u;
  __syncthreads();
  if (tid == 0) {
    const uint32_t wbytes = (uint32_t)n_valid * (BITS_WORDS * 4u), rbytes = (uint32_t)n_valid * (uint32_t)A * 16u, ebytes = (uint32_t)n_valid * 16u;
    mbar_expect_tx(s_bar, wbytes + rbytes + ebytes);
    bulk_g2s(s_bits, p.cellbits + env0 * BITS_WORDS, wbytes, s_bar);
    bulk_g2s(s_rec, p.agents + env0 * A * 16, rbytes, s_bar);
    bulk_g2s(s_env, p.envrec + env0 * 4, ebytes, s_bar);
  }
  const bool mine = tid < n_valid * A;
  const int le = mine ? tid / A : 0, a = mine ? tid - le * A : 0;
  const long long env = env0 + le;
  const int action = mine ? p.actions[env * A + a] : (int)MG_A_DONE;
  obs_prepare<OBS, V>(p, o, tid, nthreads);  // while the copies are in flight
  mbar_wait(s_bar, 0);

  // ---- phase 0 (warp 0, lane == env): the step's agent order, base.py:514-516 -- one Philox block per env ----
  if (warp == 0 && lane < n_valid) {
    const unsigned long long g = (unsigned long long)(p.env_offset + env0 + lane);
    uint32_t fact = 1;
    for (int i = 2; i <= A; ++i) fact *= (uint32_t)i;
    const U4 r = philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), (uint32_t)s_env[lane * 4 + 2], 0u, (uint32_t)p.seed, (uint32_t)(p.seed >> 32));
    s_order[lane] = decode_order(__umulhi(r.x, fact), A);
  }

  // ---- phase 1: every agent plays its action on a private copy of its record ----
  uint32_t* rec = s_rec + le * A * 4;
  const uint32_t* bits = s_bits + le * BITS_WORDS;
  uint8_t* tp = p.grid + env * 3 * S;
  uint32_t w0 = 0, w1 = 0, errb = 0, base_stamp = 0;
  bool moved = false, slow = false;
  double reward = 0.0;
  int sc = 0;
  if (mine) {
    w0 = rec[a * 4]; w1 = rec[a * 4 + 1];
    sc = s_env[le * 4] + 1;  // base.py:512
    base_stamp = (uint32_t)s_env[le * 4 + 3] & 0xFFFFu;
    if ((w0 >> 24) & MG_AF_ACTIVE) {  // base.py:521
      const int cx = (int)(w0 & 0xFFu), cy = (int)((w0 >> 8) & 0xFFu), dir = (int)((w0 >> 16) & 3u);
      if (action == MG_A_LEFT) w0 = (w0 & 0xFF00FFFFu) | ((uint32_t)((dir + 3) & 3) << 16);        // base.py:530-531
      else if (action == MG_A_RIGHT) w0 = (w0 & 0xFF00FFFFu) | ((uint32_t)((dir + 1) & 3) << 16);  // base.py:534-535
      else if (action >= MG_A_FORWARD && action <= MG_A_TOGGLE) {
        const int fx = cx + ((dir == 0) ? 1 : (dir == 2) ? -1 : 0), fy = cy + ((dir == 1) ? 1 : (dir == 3) ? -1 : 0);  // agents.py:183
        const bool inb = (unsigned)fx < (unsigned)W && (unsigned)fy < (unsigned)H;
        const uint32_t fcell = inb ? cell_triple(bits, fx & 15, fy & 15, tp, H, S) : (uint32_t)MG_T_WALL;
        const int ftype = (int)(fcell & 0xFFu);
        if (!inb) errb |= MG_ERR_STACK;
        if (action == MG_A_FORWARD) {  // base.py:538-585 (ghost mode: other agents never block)
          const int fstate = (int)(fcell >> 16);
          if (ftype == MG_T_EMPTY || can_overlap_static(ftype, fstate)) {
            const uint32_t ccell = cell_triple(bits, cx & 15, cy & 15, tp, H, S);
            if ((ccell & 0xFFu) != MG_T_EMPTY && !can_overlap_static((int)(ccell & 0xFFu), (int)(ccell >> 16))) errb |= MG_ERR_STACK;  // base.py:558
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
                const double qd = __ddiv_rn((double)sc, (double)p.max_steps);
                const double u = __dmul_rn(0.9, qd);
                const double f = __dsub_rn(1.0, u);
                rwd = __dmul_rn(rwd, f);
              }
              reward = __dadd_rn(0.0, rwd);
            }
            if (ftype == MG_T_LAVA || ftype == MG_T_GOAL) w0 = (w0 | ((uint32_t)MG_AF_DONE << 24)) & ~((uint32_t)MG_AF_ACTIVE << 24);  // base.py:584-585,646
          }
        } else if (action == MG_A_PICKUP) {  // takes effect only on a pickable object with empty hands (base.py:590-597)
          slow = ftype != MG_T_EMPTY && ((PICKUP_MASK >> ftype) & 1u) && (w1 & 0xFFu) == 0u;
        } else if (action == MG_A_DROP) {    // takes effect only when carrying and facing an empty cell (base.py:600-606)
          slow = inb && ftype == MG_T_EMPTY && (w1 & 0xFFu) != 0u;
        } else {                             // toggle: only Door / Box react (base.py:609-613)
          slow = ftype == MG_T_DOOR || ftype == MG_T_BOX;
        }
      } else if (action != MG_A_DONE) errb |= MG_ERR_BAD_ACTION;  // base.py:619-620
    }
    // one smem atomic per agent: slow request / mover bit / "not done yet" bit / error bits
    const uint32_t add = (slow ? FL_SLOW : 0u) | (moved ? (0x100u << a) : 0u) | (((w0 >> 24) & MG_AF_DONE) ? 0u : FL_NOTDONE) | (errb << 16);
    if (add) atomicOr(&s_flag[le], add);
  }
  __syncthreads();

  // ---- phase 2: commit (parallel envs) or replay sequentially (envs whose planes change) ----
  const uint32_t fl1 = mine ? s_flag[le] : 0u;
  const bool slow_env = (fl1 & FL_SLOW) != 0;
  bool used_scratch = false;
  if (mine && !slow_env) {
    rec[a * 4] = w0; rec[a * 4 + 1] = w1;
    p.rewards[env * A + a] = reward;
    if (moved) {  // arrival stamp: movers are numbered in the reference's processing order (base.py:547-552)
      const uint32_t order = s_order[le], movers = (fl1 >> 8) & 0xFFu;
      int rank = 0;
      for (int q = 0; q < A; ++q) {
        const int b = (int)((order >> (4 * q)) & 0xFu);
        if (b == a) break;
        rank += (int)((movers >> b) & 1u);
      }
      rec[a * 4 + 2] = (base_stamp + (uint32_t)rank) & 0xFFFFu;
    }
  } else if (slow_env && a == 0) {
    used_scratch = true;
    EnvCtx<32> c{p, s_trec + le, tp, s_bits + le * BITS_WORDS, s_scr + le, 0, 0, 0, 0u, false};
    for (int q = 0; q < A; ++q) { c.R(q, 0) = rec[q * 4]; c.R(q, 1) = rec[q * 4 + 1]; c.R(q, 2) = rec[q * 4 + 2]; }
    c.sc = s_env[le * 4]; c.ep = s_env[le * 4 + 1]; c.tl = s_env[le * 4 + 2]; c.w3 = (uint32_t)s_env[le * 4 + 3];
    seq_step(c, (unsigned long long)(p.env_offset + env), p.actions + env * A, p.rewards + env * A);
    bool nd = false;
    for (int q = 0; q < A; ++q) {
      rec[q * 4] = c.R(q, 0); rec[q * 4 + 1] = c.R(q, 1); rec[q * 4 + 2] = c.R(q, 2); rec[q * 4 + 3] = 0u;
      nd = nd || !((c.R(q, 0) >> 24) & MG_AF_DONE);
    }
    s_env[le * 4] = c.sc; s_env[le * 4 + 2] = c.tl; s_env[le * 4 + 3] = (int)c.w3;
    // the parallel pass left its own mover / not-done / error bits in the flag word: replace them by the replay's
    s_flag[le] = FL_SLOW | (c.dirty ? FL_BITS_DIRTY : 0u) | (nd ? FL_NOTDONE : 0u);
  }
  __syncthreads();

  // ---- phase 3 (warp 0, lane == env): env bookkeeping and done (base.py:649) ----
  bool want_reset = false;
  if (warp == 0 && lane < n_valid) {
    const int e = lane;
    const uint32_t fl = s_flag[e];
    if (!(fl & FL_SLOW)) {
      const uint32_t w3 = (uint32_t)s_env[e * 4 + 3];
      s_env[e * 4] += 1;      // step_count, base.py:512
      s_env[e * 4 + 2] += 1;  // lifetime steps
      s_env[e * 4 + 3] = (int)((w3 & 0xFFFF0000u) | (((w3 & 0xFFFFu) + (uint32_t)__popc((fl >> 8) & 0xFFu)) & 0xFFFFu) | (fl & 0xFFFF0000u));
    }
    const bool dn = (s_env[e * 4] >= p.max_steps) || !(fl & FL_NOTDONE);
    p.done[env0 + e] = dn ? 1 : 0;
    if (dn && p.autoreset) { want_reset = true; s_flag[e] = fl | FL_BITS_DIRTY | FL_RESET; }
  }
  const int scratch_state = __syncthreads_or((used_scratch ? 1 : 0) | (want_reset ? 2 : 0));
  if (scratch_state) {
    // finished envs: MultiGridEnv.reset (base.py:402-416), one lane per env spread over all warps of the CTA
    if (mine && a == 0 && (s_flag[le] & FL_RESET)) {
      EnvCtx<32> c{p, s_trec + le, tp, s_bits + le * BITS_WORDS, s_scr + le, 0, 0, 0, 0u, false};
      for (int q = 0; q < A; ++q) { c.R(q, 0) = rec[q * 4]; c.R(q, 1) = rec[q * 4 + 1]; c.R(q, 2) = rec[q * 4 + 2]; }
      c.sc = s_env[le * 4]; c.ep = s_env[le * 4 + 1]; c.tl = s_env[le * 4 + 2]; c.w3 = (uint32_t)s_env[le * 4 + 3];
      seq_reset(c, (unsigned long long)(p.env_offset + env));
      for (int q = 0; q < A; ++q) { rec[q * 4] = c.R(q, 0); rec[q * 4 + 1] = c.R(q, 1); rec[q * 4 + 2] = c.R(q, 2); rec[q * 4 + 3] = 0u; }
      s_env[le * 4] = c.sc; s_env[le * 4 + 1] = c.ep; s_env[le * 4 + 3] = (int)c.w3;
    }
    __syncthreads();
    obs_prepare<OBS, V>(p, o, tid, nthreads);  // the sequential path borrowed the output area: clean it again
    __syncthreads();
  }

  // ---- phase 4: queue heads (derived flag) from the final positions and stamps ----
  if (mine) {
    const uint32_t v0 = rec[a * 4];
    bool head = ((v0 >> 24) & MG_AF_PLACED) != 0;
    if (head) {
      const uint32_t st = rec[a * 4 + 2];
      for (int q = 0; q < A; ++q) {
        const uint32_t u0 = rec[q * 4];
        if (q != a && ((u0 >> 24) & MG_AF_PLACED) && ((u0 ^ v0) & 0xFFFFu) == 0u && rec[q * 4 + 2] < st) head = false;
      }
    }
    s_head[tid] = head ? 1 : 0;
  }
  __syncthreads();

  // ---- phase 5: observe the post-step world ----
  if (mine) {
    // own record: publish the head flag (other threads take head flags from s_head and ignore this bit)
    rec[a * 4] = s_head[tid] ? (rec[a * 4] | (AF_HEAD << 24)) : (rec[a * 4] & ~(AF_HEAD << 24));
    obs_view<OBS, V, true, true>(p, o, tid, a, env, rec, tp, bits, s_head + le * A);
  }
  __syncthreads();
  obs_emit<OBS, V, TSC>(p, o, env0, n_valid, tid, nthreads);
  if (tid == 0) {  // state goes back as it came: contiguous chunks, bulk copies
    fence_proxy_async_smem();
    bulk_s2g(p.agents + env0 * A * 16, s_rec, (uint32_t)n_valid * (uint32_t)A * 16u);
    bulk_s2g(p.envrec + env0 * 4, s_env, (uint32_t)n_valid * 16u);
    bulk_commit();
  }
  if (warp == 0 && lane < n_valid && (s_flag[lane] & FL_BITS_DIRTY)) {
    fence_proxy_async_smem();
    bulk_s2g(p.cellbits + (env0 + lane) * BITS_WORDS, s_bits + lane * BITS_WORDS, BITS_WORDS * 4u);
    bulk_commit();
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// zero-initialised family of freshly constructed envs (bonus_state = None)
__global__ void init_kernel(uint8_t* grid, uint8_t* agents, int32_t* envrec, uint32_t* cellbits, long long B, int A, int S) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n_grid = B * 3 * S / 16, n_ag = B * A, n_er = B;
  if (i < n_grid) reinterpret_cast<int4*>(grid)[i] = make_int4(0, 0, 0, 0);
  if (cellbits != nullptr && i < B * (BITS_WORDS / 4)) reinterpret_cast<int4*>(cellbits)[i] = make_int4(0, 0, 0, 0);
  if (i < n_ag) reinterpret_cast<int4*>(agents)[i] = make_int4(0, (int)0xFF000000u, 0, 0);
  if (i < n_er) reinterpret_cast<int4*>(envrec)[i] = make_int4(0, 0, 0, 0);
}

// synthetic uniform policy: 4 actions per Philox call
__global__ void random_actions_kernel(int32_t* actions, long long n, int n_actions, unsigned long long seed, unsigned long long counter) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i * 4 >= n) return;
  const U4 r = philox4x32_10((uint32_t)i, (uint32_t)(i >> 32), (uint32_t)counter, (uint32_t)(counter >> 32) ^ 0xAC710000u, (uint32_t)seed,
                             (uint32_t)(seed >> 32));
  const uint32_t v[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (i * 4 + k < n) actions[i * 4 + k] = (int32_t)__umulhi(v[k], (uint32_t)n_actions);
}

// occlude_mask (agents.py:298-343) known-answer kernel: one thread per VxV grid, layout [i][j]
template <int V>
__global__ void los_kernel(const uint8_t* __restrict__ transparent, uint8_t* __restrict__ mask, long long n, int ax, int ay) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint8_t* t = transparent + i * V * V;
  uint32_t T[V], M[V];
#pragma unroll
  for (int j = 0; j < V; ++j) {
    uint32_t row = 0;
#pragma unroll
    for (int x = 0; x < V; ++x) row |= (t[x * V + j] ? 1u : 0u) << x;
    T[j] = row;
  }
  // generic agent position: the row-mask routine is specialised on (ax, ay) being runtime values
  occlude_rows<V>(T, ax, ay, M);
  uint8_t* m = mask + i * V * V;
#pragma unroll
  for (int j = 0; j < V; ++j)
#pragma unroll
    for (int x = 0; x < V; ++x) m[x * V + j] = (uint8_t)((M[j] >> x) & 1u);
}

}  // namespace mg

// =============================================================================================
// host side: C ABI (include/marlgrid_b200.h)
// =============================================================================================
using namespace mg;

static std::atomic<long long> g_launches{0};

static int check_cfg(const MgConfig* c) {
  if (!c) return MG_E_CONFIG;
  if (c->n_agents < 1 || c->n_agents > MG_MAX_AGENTS) return MG_E_CONFIG;
  if (c->view_size < 3 || c->view_size > MG_MAX_VIEW) return MG_E_CONFIG;
  if (c->width < 3 || c->height < 3 || c->width > 255 || c->height > 255) return MG_E_CONFIG;
  if (c->plane_stride % 16 != 0 || c->plane_stride < c->width * c->height) return MG_E_CONFIG;
  if (c->view_offset < 0 || c->view_offset >= c->view_size) return MG_E_CONFIG;
  if (c->max_steps < 1 || c->n_clutter < 0 || c->n_bonus_tiles < 0 || c->n_bonus_tiles > 250) return MG_E_CONFIG;
  return 0;
}

static KP make_kp(const MgConfig* c, const MgState* st) {
  KP p;
  memset(&p, 0, sizeof p);
  p.W = c->width; p.H = c->height; p.A = c->n_agents; p.V = c->view_size; p.vo = c->view_offset; p.ts = c->view_tile_size;
  p.max_steps = c->max_steps; p.n_clutter = c->n_clutter; p.n_bonus = c->n_bonus_tiles; p.goal_mode = c->goal_mode;
  p.flags = c->flags; p.S = c->plane_stride;
  p.goal_reward = c->goal_reward; p.bonus_reward = c->bonus_reward; p.bonus_penalty = c->bonus_penalty;
  for (int i = 0; i < MG_MAX_AGENTS; ++i) { p.agent_color[i] = c->agent_color[i]; p.spawn_delay[i] = c->spawn_delay[i]; }
  for (int i = 0; i < 15; ++i) p.kind_of_type[i] = c->kind_of_type[i];
  p.kind_of_type[15] = 0xFF;
  p.grid = st->grid; p.agents = st->agents; p.envrec = st->envrec; p.B = st->n_envs;
  p.cellbits = (c->width <= 16 && c->height <= 16) ? st->cellbits : nullptr; p.env_offset = st->env_offset; p.seed = st->seed;
  p.n_tiles = (c->n_static_kinds + 1) * (1 + 4 * c->n_agents);
  p.orient_slots = 4;
  return p;
}

static size_t obs_smem_bytes(const KP& p, int obs) {
  size_t b = (p.cellbits ? (size_t)ENVS_PER_CTA * BITS_WORDS * 4 : (size_t)ENVS_PER_CTA * 3 * p.S) + (size_t)ENVS_PER_CTA * p.A * 16 + 16;
  if (obs == 1) b += (size_t)ENVS_PER_CTA * p.A * p.V * p.V * 3;
  if (obs == 2) {
    b += (size_t)ENVS_PER_CTA * p.A * p.V * p.V + (size_t)((ENVS_PER_CTA * p.A + 15) / 16) * 16;
    b += (size_t)(p.n_tiles * p.orient_slots + 1) * p.ts * p.ts * 3;
  }
  return (b + 15) / 16 * 16;
}

template <int OBS, int V, int TSC, bool BITS>
static int launch_obs_one(const KP& p, cudaStream_t s) {
  const size_t sm = obs_smem_bytes(p, OBS);
  auto k = obs_kernel<OBS, V, TSC, BITS>;
  static size_t configured[64] = {0};  // per instantiation and device
  int dev = 0;
  cudaGetDevice(&dev);
  if (sm > 48 * 1024 && sm > configured[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (e != cudaSuccess) return (int)e;
    configured[dev & 63] = sm;
  }
  const long long blocks = (p.B + ENVS_PER_CTA - 1) / ENVS_PER_CTA;
  if (blocks <= 0) return 0;
  k<<<(unsigned)blocks, 32 * p.A, sm, s>>>(p);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return (int)cudaGetLastError();
}

template <int OBS, int TSC, bool BITS>
static int launch_obs_v(const KP& p, cudaStream_t s) {
  switch (p.V) {
    case 3: return launch_obs_one<OBS, 3, TSC, BITS>(p, s);
    case 4: return launch_obs_one<OBS, 4, TSC, BITS>(p, s);
    case 5: return launch_obs_one<OBS, 5, TSC, BITS>(p, s);
    case 6: return launch_obs_one<OBS, 6, TSC, BITS>(p, s);
    case 7: return launch_obs_one<OBS, 7, TSC, BITS>(p, s);
    case 8: return launch_obs_one<OBS, 8, TSC, BITS>(p, s);
  }
  return MG_E_CONFIG;
}

// obs: 1 encoded / 2 rgb
static int launch_obs(const KP& p, int obs, cudaStream_t s) {
  const bool bits = p.cellbits != nullptr;
  if (obs == 1) return bits ? launch_obs_v<1, 0, true>(p, s) : launch_obs_v<1, 0, false>(p, s);
  if (p.ts == 8) return bits ? launch_obs_v<2, 8, true>(p, s) : launch_obs_v<2, 8, false>(p, s);
  if (p.ts % 4 == 0) return bits ? launch_obs_v<2, 1, true>(p, s) : launch_obs_v<2, 1, false>(p, s);
  return bits ? launch_obs_v<2, 0, true>(p, s) : launch_obs_v<2, 0, false>(p, s);
}

static size_t fused_smem_bytes(const KP& p, int obs) {
  size_t out = 0;
  if (obs == 1) out = (size_t)ENVS_PER_CTA * p.A * p.V * p.V * 3;
  else out = (size_t)ENVS_PER_CTA * p.A * p.V * p.V + (size_t)((ENVS_PER_CTA * p.A + 15) / 16) * 16 + (size_t)(p.n_tiles * p.orient_slots + 1) * p.ts * p.ts * 3;
  const size_t scratch = (size_t)(p.A * 4 * 32 + 64 * 32) * 4;
  const size_t b = (size_t)ENVS_PER_CTA * BITS_WORDS * 4 + (size_t)ENVS_PER_CTA * p.A * 16 + (size_t)ENVS_PER_CTA * 16 + (size_t)ENVS_PER_CTA * 8 +
                   32 * MG_MAX_AGENTS + 16 + std::max(out, scratch);
  return (b + 15) / 16 * 16;
}

template <int OBS, int V, int TSC>
static int launch_fused_one(const KP& p, cudaStream_t s) {
  const size_t sm = fused_smem_bytes(p, OBS);
  auto k = fused_kernel<OBS, V, TSC>;
  static size_t configured[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (sm > 48 * 1024 && sm > configured[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (e != cudaSuccess) return (int)e;
    configured[dev & 63] = sm;
  }
  const long long blocks = (p.B + ENVS_PER_CTA - 1) / ENVS_PER_CTA;
  if (blocks <= 0) return 0;
  k<<<(unsigned)blocks, 32 * p.A, sm, s>>>(p);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return (int)cudaGetLastError();
}

template <int OBS, int TSC>
static int launch_fused_v(const KP& p, cudaStream_t s) {
  switch (p.V) {
    case 3: return launch_fused_one<OBS, 3, TSC>(p, s);
    case 4: return launch_fused_one<OBS, 4, TSC>(p, s);
    case 5: return launch_fused_one<OBS, 5, TSC>(p, s);
    case 6: return launch_fused_one<OBS, 6, TSC>(p, s);
    case 7: return launch_fused_one<OBS, 7, TSC>(p, s);
    case 8: return launch_fused_one<OBS, 8, TSC>(p, s);
  }
  return MG_E_CONFIG;
}

// the one-launch path exists for bit-plane worlds in ghost mode without respawn / spawn delay (every registered env)
static bool fused_eligible(const KP& p) {
  if (p.cellbits == nullptr || !(p.flags & MG_F_GHOST) || (p.flags & MG_F_RESPAWN)) return false;
  for (int a = 0; a < p.A; ++a)
    if (p.spawn_delay[a] != 0) return false;
  return true;
}

static int g_force_two_kernels = 0;  // test hook: exercise the per-env step kernel + observe kernel pair

// per-env kernels: MODE 0 step (+auto-reset), 1 reset, 2 sync derived state
template <int MODE>
static int launch_env(const KP& p, cudaStream_t s) {
  const long long blocks = (p.B + ENV_THREADS - 1) / ENV_THREADS;
  if (blocks <= 0) return 0;
  const size_t sm = (size_t)ENV_THREADS * p.A * 16 + (size_t)ENV_THREADS * 64 * 4;
  if (p.A <= 4) {
    if (p.cellbits) env_kernel<MODE, true, 4><<<(unsigned)blocks, ENV_THREADS, sm, s>>>(p);
    else env_kernel<MODE, false, 4><<<(unsigned)blocks, ENV_THREADS, sm, s>>>(p);
  } else {
    if (p.cellbits) env_kernel<MODE, true, MG_MAX_AGENTS><<<(unsigned)blocks, ENV_THREADS, sm, s>>>(p);
    else env_kernel<MODE, false, MG_MAX_AGENTS><<<(unsigned)blocks, ENV_THREADS, sm, s>>>(p);
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return (int)cudaGetLastError();
}

static cudaEvent_t g_mid_event = nullptr;  // profiling hook: recorded between the two launches of a step

// env.step: one fused launch when eligible; else the step kernel (incl. auto-reset), then the observation
static int launch_step_obs(const KP& p, int obs, cudaStream_t s) {
  if (obs != 0 && !g_force_two_kernels && fused_eligible(p)) {
    if (obs == 1) return launch_fused_v<1, 0>(p, s);
    if (p.ts == 8) return launch_fused_v<2, 8>(p, s);
    return (p.ts % 4 == 0) ? launch_fused_v<2, 1>(p, s) : launch_fused_v<2, 0>(p, s);
  }
  int e = launch_env<0>(p, s);
  if (e) return e;
  if (g_mid_event) cudaEventRecord(g_mid_event, s);
  if (obs != 0) return launch_obs(p, obs, s);
  return 0;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static int check_state(const MgConfig* c, const MgState* st) {
  int e = check_cfg(c);
  if (e) return e;
  if (!st || !st->grid || !st->agents || !st->envrec || st->n_envs < 0) return MG_E_ARG;
  if (!aligned16(st->grid) || !aligned16(st->agents) || !aligned16(st->envrec) || !aligned16(st->cellbits)) return MG_E_ARG;
  return 0;
}

extern "C" {

int mg_version(void) { return 1; }
const char* mg_build_info(void) { return "marlgrid_b200 sm_100a: per-env step kernel + bit-plane observe kernel (cp.async.bulk + mbarrier staging, 32 envs/CTA)"; }
int mg_sizeof_config(void) { return (int)sizeof(MgConfig); }
int mg_config_validate(const MgConfig* cfg) { return check_cfg(cfg); }
int64_t mg_obs_bytes_per_env(const MgConfig* c, int rgb) {
  if (check_cfg(c)) return MG_E_CONFIG;
  const int64_t v = c->view_size;
  return rgb ? (int64_t)c->n_agents * v * c->view_tile_size * v * c->view_tile_size * 3 : (int64_t)c->n_agents * v * v * 3;
}
int64_t mg_launch_count(void) { return g_launches.load(); }
void mg_debug_set_mid_event(void* cuda_event) { g_mid_event = (cudaEvent_t)cuda_event; }
void mg_debug_force_two_kernels(int on) { g_force_two_kernels = on; }

int mg_init(const MgConfig* cfg, const MgState* st, mg_stream_t stream) {
  int e = check_state(cfg, st);
  if (e) return e;
  if (st->n_envs == 0) return 0;
  const long long n = std::max<long long>(std::max<long long>(st->n_envs * 3 * cfg->plane_stride / 16, st->n_envs * cfg->n_agents), st->n_envs * (BITS_WORDS / 4));
  init_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(st->grid, st->agents, st->envrec, st->cellbits, st->n_envs, cfg->n_agents, cfg->plane_stride);
  g_launches.fetch_add(1);
  return (int)cudaGetLastError();
}

int mg_sync_derived(const MgConfig* cfg, const MgState* st, mg_stream_t stream) {
  int e = check_state(cfg, st);
  if (e) return e;
  if (st->n_envs == 0) return 0;
  KP p = make_kp(cfg, st);
  return launch_env<2>(p, (cudaStream_t)stream);
}

int mg_reset(const MgConfig* cfg, const MgState* st, const uint8_t* reset_mask, mg_stream_t stream) {
  int e = check_state(cfg, st);
  if (e) return e;
  KP p = make_kp(cfg, st);
  p.reset_mask = reset_mask;
  return launch_env<1>(p, (cudaStream_t)stream);
}

int mg_step(const MgConfig* cfg, const MgState* st, const int32_t* actions, double* rewards, uint8_t* done, int autoreset,
            mg_stream_t stream) {
  int e = check_state(cfg, st);
  if (e) return e;
  if (!actions || !rewards || !done) return MG_E_ARG;
  KP p = make_kp(cfg, st);
  p.actions = actions; p.rewards = rewards; p.done = done; p.autoreset = autoreset;
  return launch_step_obs(p, 0, (cudaStream_t)stream);
}

int mg_obs_encode(const MgConfig* cfg, const MgState* st, uint8_t* obs, mg_stream_t stream) {
  int e = check_state(cfg, st);
  if (e) return e;
  if (!obs || !aligned16(obs)) return MG_E_ARG;
  KP p = make_kp(cfg, st);
  p.obs = obs;
  return launch_obs(p, 1, (cudaStream_t)stream);
}

static int atlas_mode(const MgConfig* cfg) { return (cfg->view_tile_size <= 10) ? 1 : 4; }  // empty_tile alpha == 0 (base.py:247): rotation-equivariant

int mg_obs_rgb(const MgConfig* cfg, const MgState* st, const uint8_t* atlas, uint8_t* obs, mg_stream_t stream) {
  int e = check_state(cfg, st);
  if (e) return e;
  if (!obs || !atlas || !aligned16(obs) || !aligned16(atlas) || cfg->view_tile_size < 1) return MG_E_ARG;
  KP p = make_kp(cfg, st);
  p.obs = obs; p.atlas = atlas; p.orient_slots = atlas_mode(cfg);
  return launch_obs(p, 2, (cudaStream_t)stream);
}

int mg_step_fused(const MgConfig* cfg, const MgState* st, const int32_t* actions, double* rewards, uint8_t* done, uint8_t* obs,
                  int autoreset, mg_stream_t stream) {
  int e = check_state(cfg, st);
  if (e) return e;
  if (!actions || !rewards || !done || !obs || !aligned16(obs)) return MG_E_ARG;
  KP p = make_kp(cfg, st);
  p.actions = actions; p.rewards = rewards; p.done = done; p.obs = obs; p.autoreset = autoreset;
  return launch_step_obs(p, 1, (cudaStream_t)stream);
}

int mg_step_fused_rgb(const MgConfig* cfg, const MgState* st, const int32_t* actions, double* rewards, uint8_t* done,
                      const uint8_t* atlas, uint8_t* obs, int autoreset, mg_stream_t stream) {
  int e = check_state(cfg, st);
  if (e) return e;
  if (!actions || !rewards || !done || !obs || !atlas || !aligned16(obs) || !aligned16(atlas) || cfg->view_tile_size < 1) return MG_E_ARG;
  KP p = make_kp(cfg, st);
  p.actions = actions; p.rewards = rewards; p.done = done; p.obs = obs; p.atlas = atlas; p.autoreset = autoreset;
  p.orient_slots = atlas_mode(cfg);
  return launch_step_obs(p, 2, (cudaStream_t)stream);
}

int mg_rollout_fused(const MgConfig* cfg, const MgState* st, const int32_t* actions, int64_t n_steps, double* rewards, uint8_t* done,
                     uint8_t* obs, int autoreset, mg_stream_t stream) {
  int e = check_state(cfg, st);
  if (e) return e;
  if (!actions || !rewards || !done || !obs || !aligned16(obs)) return MG_E_ARG;
  KP p = make_kp(cfg, st);
  p.rewards = rewards; p.done = done; p.obs = obs; p.autoreset = autoreset;
  for (int64_t t = 0; t < n_steps; ++t) {
    p.actions = actions + t * st->n_envs * cfg->n_agents;
    e = launch_step_obs(p, 1, (cudaStream_t)stream);
    if (e) return e;
  }
  return 0;
}

int mg_random_actions(int32_t* actions, int64_t n, int n_actions, uint64_t seed, uint64_t counter, mg_stream_t stream) {
  if (!actions || n < 0 || n_actions < 1) return MG_E_ARG;
  if (n == 0) return 0;
  const long long threads = (n + 3) / 4;
  random_actions_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(actions, n, n_actions, seed, counter);
  g_launches.fetch_add(1);
  return (int)cudaGetLastError();
}

int mg_los_batch(const uint8_t* transparent, uint8_t* mask, int64_t n, int view_size, int ax, int ay, mg_stream_t stream) {
  if (!transparent || !mask || n < 0 || ax < 0 || ay < 0 || ax >= view_size || ay >= view_size) return MG_E_ARG;
  if (n == 0) return 0;
  const unsigned blocks = (unsigned)((n + 127) / 128);
  cudaStream_t s = (cudaStream_t)stream;
  switch (view_size) {
    case 3: los_kernel<3><<<blocks, 128, 0, s>>>(transparent, mask, n, ax, ay); break;
    case 4: los_kernel<4><<<blocks, 128, 0, s>>>(transparent, mask, n, ax, ay); break;
    case 5: los_kernel<5><<<blocks, 128, 0, s>>>(transparent, mask, n, ax, ay); break;
    case 6: los_kernel<6><<<blocks, 128, 0, s>>>(transparent, mask, n, ax, ay); break;
    case 7: los_kernel<7><<<blocks, 128, 0, s>>>(transparent, mask, n, ax, ay); break;
    case 8: los_kernel<8><<<blocks, 128, 0, s>>>(transparent, mask, n, ax, ay); break;
    default: return MG_E_CONFIG;
  }
  g_launches.fetch_add(1);
  return (int)cudaGetLastError();
}

// ---- host-buffer engine -----------------------------------------------------------------------
struct MgEngine {
  MgConfig cfg;
  MgState st;
  int device, rgb;
  cudaStream_t stream;
  int32_t* d_actions;
  double* d_rewards;
  uint8_t* d_done;
  uint8_t* d_obs;
  uint8_t* d_atlas;
  int64_t obs_bytes;
};

#define MG_CUDA(x)                                \
  do {                                            \
    cudaError_t _e = (x);                         \
    if (_e != cudaSuccess) return (int)_e;        \
  } while (0)

int mg_engine_create(MgEngine** out, const MgConfig* cfg, int64_t n_envs, int64_t env_offset, uint64_t seed, int device, int rgb,
                     const uint8_t* atlas_host, int64_t atlas_bytes) {
  if (!out || n_envs < 1) return MG_E_ARG;
  int e = check_cfg(cfg);
  if (e) return e;
  if (rgb && (!atlas_host || atlas_bytes <= 0)) return MG_E_ARG;
  MG_CUDA(cudaSetDevice(device));
  MgEngine* en = new MgEngine();
  memset(en, 0, sizeof *en);
  en->cfg = *cfg; en->device = device; en->rgb = rgb;
  en->st.n_envs = n_envs; en->st.env_offset = env_offset; en->st.seed = seed;
  en->obs_bytes = n_envs * mg_obs_bytes_per_env(cfg, rgb);
  MG_CUDA(cudaStreamCreateWithFlags(&en->stream, cudaStreamNonBlocking));
  MG_CUDA(cudaMalloc(&en->st.grid, (size_t)n_envs * 3 * cfg->plane_stride));
  MG_CUDA(cudaMalloc(&en->st.agents, (size_t)n_envs * cfg->n_agents * MG_AGENT_REC));
  MG_CUDA(cudaMalloc(&en->st.envrec, (size_t)n_envs * MG_ENV_REC));
  MG_CUDA(cudaMalloc(&en->st.cellbits, (size_t)n_envs * BITS_WORDS * 4));
  MG_CUDA(cudaMalloc(&en->d_actions, (size_t)n_envs * cfg->n_agents * sizeof(int32_t)));
  MG_CUDA(cudaMalloc(&en->d_rewards, (size_t)n_envs * cfg->n_agents * sizeof(double)));
  MG_CUDA(cudaMalloc(&en->d_done, (size_t)n_envs));
  MG_CUDA(cudaMalloc(&en->d_obs, (size_t)en->obs_bytes));
  if (rgb) {
    MG_CUDA(cudaMalloc(&en->d_atlas, (size_t)atlas_bytes));
    MG_CUDA(cudaMemcpy(en->d_atlas, atlas_host, (size_t)atlas_bytes, cudaMemcpyHostToDevice));
  }
  e = mg_init(&en->cfg, &en->st, en->stream);
  if (e) return e;
  MG_CUDA(cudaStreamSynchronize(en->stream));
  *out = en;
  return 0;
}

void mg_engine_destroy(MgEngine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  cudaStreamSynchronize(e->stream);
  cudaFree(e->st.grid); cudaFree(e->st.agents); cudaFree(e->st.envrec); cudaFree(e->st.cellbits);
  cudaFree(e->d_actions); cudaFree(e->d_rewards); cudaFree(e->d_done); cudaFree(e->d_obs); cudaFree(e->d_atlas);
  cudaStreamDestroy(e->stream);
  delete e;
}

int mg_engine_reset(MgEngine* e, uint8_t* obs_host) {
  if (!e) return MG_E_ARG;
  MG_CUDA(cudaSetDevice(e->device));
  int r = mg_reset(&e->cfg, &e->st, nullptr, e->stream);
  if (r) return r;
  r = e->rgb ? mg_obs_rgb(&e->cfg, &e->st, e->d_atlas, e->d_obs, e->stream) : mg_obs_encode(&e->cfg, &e->st, e->d_obs, e->stream);
  if (r) return r;
  if (obs_host) MG_CUDA(cudaMemcpyAsync(obs_host, e->d_obs, (size_t)e->obs_bytes, cudaMemcpyDeviceToHost, e->stream));
  MG_CUDA(cudaStreamSynchronize(e->stream));
  return 0;
}

int mg_engine_step(MgEngine* e, const int32_t* actions_host, uint8_t* obs_host, double* rewards_host, uint8_t* done_host, int autoreset) {
  if (!e || !actions_host) return MG_E_ARG;
  MG_CUDA(cudaSetDevice(e->device));
  const size_t na = (size_t)e->st.n_envs * e->cfg.n_agents;
  MG_CUDA(cudaMemcpyAsync(e->d_actions, actions_host, na * sizeof(int32_t), cudaMemcpyHostToDevice, e->stream));
  int r = e->rgb ? mg_step_fused_rgb(&e->cfg, &e->st, e->d_actions, e->d_rewards, e->d_done, e->d_atlas, e->d_obs, autoreset, e->stream)
                 : mg_step_fused(&e->cfg, &e->st, e->d_actions, e->d_rewards, e->d_done, e->d_obs, autoreset, e->stream);
  if (r) return r;
  if (obs_host) MG_CUDA(cudaMemcpyAsync(obs_host, e->d_obs, (size_t)e->obs_bytes, cudaMemcpyDeviceToHost, e->stream));
  if (rewards_host) MG_CUDA(cudaMemcpyAsync(rewards_host, e->d_rewards, na * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
  if (done_host) MG_CUDA(cudaMemcpyAsync(done_host, e->d_done, (size_t)e->st.n_envs, cudaMemcpyDeviceToHost, e->stream));
  MG_CUDA(cudaStreamSynchronize(e->stream));
  return 0;
}

void* mg_host_alloc(int64_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, (size_t)bytes) != cudaSuccess) return nullptr;
  return p;
}
void mg_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

}  // extern "C"

// mg_env.cuh -- MultiGridEnv.step / reset for ONE env (sequential restatement of base.py:402-416,501-649 on the SoA
// state), used by the per-env kernel and as the exact slow path of the fused kernels.
#pragma once
#include "mg_common.cuh"

namespace mg {

// ---------------------------------------------------------------------------------------------
// per-env game state while a thread runs step()/reset(): agent records transposed in shared memory, word w
// of agent a at rec[(a*4+w)*RS] (RS = threads per CTA) -> bank == thread, conflict-free for any per-thread a.
//   w0 = x | y<<8 | dir<<16 | flags<<24     w1 = carry_type | carry_colour<<8 | carry_state<<16 | bonus<<24
//   w2 = stamp                               w3 = scratch (front-cell prefetch)
// `tp` = the env's type plane in global memory (colour at +S, state at +2S); `bits` = its bit-plane words (word w at bits[w * BS], mg_common.cuh).
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
  double* prest = nullptr;  // the env's GridAgentInterface.prestige values [A] (agents.py:141-153) or nullptr
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
  for (int t = 0; t < c.p.amax; ++t) {  // place_obj(agent, **agent_spawn_kwargs), base.py:505,642
    int x, y;
    d.next_box(c.p.ax0, c.p.ay0, c.p.aw, c.p.ah, x, y);
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

// base.py:402-416 reset + DoorKeyEnv._gen_grid (doorkey.py:15-41, with `_rand_int(lo, hi)` = np_random.randint(lo, hi): the
// reference's class calls a method it does not have): border walls, goal at (W-2, H-2), a vertical wall at a random column
// with a locked yellow door at a random row, a yellow key somewhere left of the wall, agents anywhere (the generator resets
// agent_spawn_kwargs to {}, doorkey.py:40).  Sequential, on the byte planes; the bit-plane lines are rebuilt from them.
template <int RS>
__device__ void env_reset_doorkey(EnvCtx<RS>& c, unsigned long long g) {
  const KP& p = c.p;
  const int W = p.W, H = p.H, S = p.S, A = p.A;
  for (int a = 0; a < A; ++a) {  // agents.py:161-170 (dir survives)
    c.R(a, 0) = c.R(a, 0) & 0x00FF0000u;
    c.R(a, 1) = 0xFF000000u;
    c.R(a, 2) = 0;
    if (c.prest != nullptr) c.prest[a] = 0.0;  // new_episode: agents.py:167-168
  }
  int4* z = reinterpret_cast<int4*>(c.tp);
  for (int i = 0; i < 3 * S / 16; ++i) z[i] = make_int4(0, 0, 0, 0);
  c.w3 &= 0xFFFF0000u;
  auto put = [&](int x, int y, int type, int colour, int state) {
    const int idx = x * H + y;
    c.tp[idx] = (uint8_t)type; c.tp[S + idx] = (uint8_t)colour; c.tp[2 * S + idx] = (uint8_t)state;
  };
  for (int i = 0; i < W; ++i) { put(i, 0, MG_T_WALL, MG_C_WORST, 0); put(i, H - 1, MG_T_WALL, MG_C_WORST, 0); }  // wall_rect base.py:172-176
  for (int j = 0; j < H; ++j) { put(0, j, MG_T_WALL, MG_C_WORST, 0); put(W - 1, j, MG_T_WALL, MG_C_WORST, 0); }
  put(W - 2, H - 2, MG_T_GOAL, MG_C_GREEN, 0);                         // doorkey.py:23
  Draws d = make_draws(p, g, (uint32_t)c.ep, TAG_RESET);
  const int split = d.next_int(2, W - 2);                              // doorkey.py:26
  for (int j = 0; j < H; ++j) put(split, j, MG_T_WALL, MG_C_WORST, 0); // vert_wall(splitIdx, 0) base.py:166-170
  const int door = d.next_int(1, W - 2);                               // doorkey.py:34 (sic: width)
  if (door < H) put(split, door, MG_T_DOOR, MG_C_YELLOW, MG_DOOR_LOCKED); else c.add_err(MG_ERR_STACK);  // grid.set asserts j < height
  {  // place_obj(Key('yellow'), top=(0, 0), size=(splitIdx, height)), doorkey.py:37: an empty cell (no agent is placed yet)
    int t = 0;
    for (; t < 100000; ++t) {
      int x, y;
      d.next_box(0, 0, split, H, x, y);
      if (c.tp[x * H + y] == MG_T_EMPTY) { put(x, y, MG_T_KEY, MG_C_YELLOW, 0); break; }
    }
    if (t == 100000) c.add_err(MG_ERR_PLACEMENT);
  }
  bits_rebuild(c.tp, c.bits, W, H, S);
  for (int a = 0; a < A; ++a)  // base.py:409-412 with agent_spawn_kwargs = {}
    if (p.spawn_delay[a] == 0) {
      bool placed = false;
      for (int t = 0; t < 100000 && !placed; ++t) {
        int x, y;
        d.next(W, H, x, y);
        placed = try_place_agent(c, x, y, a);
      }
      if (!placed) c.add_err(MG_ERR_PLACEMENT);
      c.R(a, 0) |= (uint32_t)MG_AF_ACTIVE << 24;
    }
  c.sc = 0;
  c.ep += 1;
  c.dirty = true;
}

// base.py:402-416 reset + _gen_grid (empty.py:9-16, cluttered.py:25-36, goalcycle.py:30-51), written straight to
// the global planes.  A fresh world only ever holds canonical walls, a Goal and BonusTiles, so with BITS the
// rejection sampling (base.py:690-708) runs on row/column mask sets kept in shared memory (walls / overlappable
// others) and never reads a plane; the bit-plane words are those masks.
// The placement list (goal?, bonus tiles, clutter walls, agents) is walked by ONE loop over the try index k, so
// that all lanes of a warp draw their Philox block on the same iteration (two tries per block).
// PLANES = false (needs BITS): the byte planes are NOT written here; the caller rebuilds them from the bit-plane lines and
// the object list (mg_fused2.cu stages the image in shared memory and stores it with one bulk copy per group of envs).
template <int RS, bool BITS, bool PLANES = true>
__device__ void env_reset(EnvCtx<RS>& c, unsigned long long g) {
  const KP& p = c.p;
  if (p.scenario == MG_SCENARIO_DOORKEY) { env_reset_doorkey(c, g); return; }
  const int W = p.W, H = p.H, S = p.S, A = p.A;
  for (int a = 0; a < A; ++a) {  // agents.py:161-170 (dir survives)
    c.R(a, 0) = c.R(a, 0) & 0x00FF0000u;
    c.R(a, 1) = 0xFF000000u;
    c.R(a, 2) = 0;
    if (c.prest != nullptr) c.prest[a] = 0.0;  // new_episode: agents.py:167-168
  }
  if (PLANES) {
    int4* z = reinterpret_cast<int4*>(c.tp);
    for (int i = 0; i < 3 * S / 16; ++i) z[i] = make_int4(0, 0, 0, 0);
  }
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
    for (int k = 0; k < OBJ_SLOTS; ++k) c.bits[(OBJ_WORD0 + k) * BS] = 0u;
  }
  if (PLANES) {
    for (int i = 0; i < W; ++i) {  // wall_rect base.py:172-176
      c.tp[i * H] = MG_T_WALL; c.tp[S + i * H] = MG_C_WORST;
      c.tp[i * H + H - 1] = MG_T_WALL; c.tp[S + i * H + H - 1] = MG_C_WORST;
    }
    for (int j = 0; j < H; ++j) {
      c.tp[j] = MG_T_WALL; c.tp[S + j] = MG_C_WORST;
      c.tp[(W - 1) * H + j] = MG_T_WALL; c.tp[S + (W - 1) * H + j] = MG_C_WORST;
    }
  }
  int n_listed = 0;
  auto put_static = [&](int x, int y, int type, int colour, int state) {
    if (PLANES) {
      const int idx = x * H + y;
      c.tp[idx] = (uint8_t)type; c.tp[S + idx] = (uint8_t)colour; c.tp[2 * S + idx] = (uint8_t)state;
    }
    if (BITS) {
      if (type == MG_T_WALL) { wall[x * RS] |= 1u << y; wallc[y * RS] |= 1u << x; }
      else {
        other[x * RS] |= 1u << y; otherc[y * RS] |= 1u << x;
        if (n_listed < OBJ_SLOTS) c.bits[(OBJ_WORD0 + n_listed++) * BS] = obj_entry(x, y, type, colour, state);  // Goal / BonusTiles
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
    if (agent >= 0) d.next_box(p.ax0, p.ay0, p.aw, p.ah, x, y);  // place_obj(agent, **agent_spawn_kwargs), base.py:409-412
    else d.next(W, H, x, y);
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
    } else if (++tries >= (agent >= 0 ? p.amax : 100)) {
      c.add_err(MG_ERR_PLACEMENT);  // RecursionError base.py:706
      if (agent >= 0) c.R(agent, 0) |= (uint32_t)MG_AF_ACTIVE << 24;  // the reference would have raised before activate()
      ++obj; tries = 0;
    }
  }
  c.sc = 0;
  c.ep += 1;
  if (BITS) {  // the masks ARE the bit-plane lines (OP = walls, OT = Goal / BonusTiles)
    uint32_t* bits = c.bits;
    for (int i = 0; i < 16; ++i) {
      bits[(LINE_X0 + i) * BS] = wall[i * RS] | (other[i * RS] << 16);
      bits[(LINE_Y0 + i) * BS] = wallc[i * RS] | (otherc[i * RS] << 16);
    }
    bits[0] = 0u; bits[17 * BS] = 0u; bits[18 * BS] = 0u; bits[35 * BS] = 0u;
    for (int i = OBJ_WORD0 + OBJ_SLOTS; i < BITS_WORDS; ++i) bits[i * BS] = 0u;
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
static __constant__ uint32_t RECIP32[9] = {0u, 0u, 0x80000000u, 0x55555556u, 0x40000000u, 0x33333334u, 0x2AAAAAABu, 0x24924925u, 0x20000000u};  // ceil(2^32/n)
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
              if (c.prest != nullptr) {  // agent.reward(rwd) base.py:581, agents.py:146-153
                if ((p.prestige_neg >> a) & 1u) c.add_err(MG_ERR_PRESTIGE);  // `self.rew += rew`: no such attribute
                else c.prest[a] = (rwd >= 0.0) ? __dadd_rn(c.prest[a], rwd) : 0.0;
              }
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
      if (c.prest != nullptr && !(action > MG_A_DONE || action < 0)) c.prest[a] = __dmul_rn(c.prest[a], p.pbeta[a]);  // agent.on_step base.py:622, agents.py:141-144
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

}  // namespace mg

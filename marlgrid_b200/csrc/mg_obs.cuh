// mg_obs.cuh -- egocentric observation building blocks (gen_obs_grid base.py:418-451, occlude_mask agents.py:298-343,
// MultiGrid.encode base.py:196-214, MultiGrid.render base.py:301-331) shared by the observe and fused kernels.
#pragma once
#include "mg_common.cuh"

namespace mg {

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

// transparency / non-empty / canonical-wall rows from the bit-plane lines: one word per view row
template <int V>
__device__ __forceinline__ void rows_from_bits(const uint32_t* __restrict__ bits /* this env's words */, const ViewGeom& g,
                                               uint32_t (&T)[V], uint32_t (&NE)[V], uint32_t (&CW)[V]) {
  constexpr uint32_t RM = (1u << V) - 1u;
  const uint32_t* bp = bits + (g.vertical ? LINE_Y0 : LINE_X0) * BS;
  const int sh = g.v0 + 8;                  // >= 1: the 16 board bits are parked at bits 8..23 before shifting right
#pragma unroll
  for (int b = 0; b < V; ++b) {
    const int idx = g.u0 + (g.flip ? V - 1 - b : b);
    const bool in = (unsigned)idx < (unsigned)g.Lu;  // rows outside the world: empty, transparent
    const uint32_t w = in ? bp[idx * BS] : 0u;
    uint32_t opq = (((w & 0xFFFFu) << 8) >> sh) & RM;
    uint32_t ot = (((w >> 16) << 8) >> sh) & RM;
    if (g.rev) { opq = rev_bits<V>(opq); ot = rev_bits<V>(ot); }
    T[b] = ~opq & RM;
    NE[b] = opq | ot;
    CW[b] = opq & ~ot;
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
  uint8_t* amax;    // OBS 2, 'prestige' agents: [1 + 4A] largest triangle alpha of every agent tile slot
  uint8_t* pcol;    // OBS 2, 'prestige' agents: [32*A][4] = (red, blue, active, -) of every agent of the tile's envs
};
constexpr int PRESTIGE_SMEM = 64 + ENVS_PER_CTA * MG_MAX_AGENTS * 4;  // bytes behind the atlas when MgConfig.prestige_mask != 0

template <int V>
__device__ __forceinline__ ObsSmem<V> obs_smem(uint8_t* s_out, int A) {
  ObsSmem<V> o;
  o.out = s_out; o.tile = s_out;
  o.orient = o.tile + ENVS_PER_CTA * A * V * V;
  o.atlas = o.orient + ((ENVS_PER_CTA * A + 15) / 16) * 16;
  o.amax = nullptr; o.pcol = nullptr;
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
    if (o.amax != nullptr) {  // 'prestige' agents (agents.py:92-119): the largest alpha of every (white) agent tile, from the global atlas
      for (int slot = tid; slot < 1 + 4 * A; slot += nthreads) {
        int m = 0;
        for (int px = 0; px < p.ts * p.ts; ++px)
          m = max(m, (int)p.atlas[(size_t)(slot * 4) * tile_bytes + px * 3] - (int)p.atlas[px * 3]);  // minus the empty tile's border (base.py:245-250,296-298)
        o.amax[slot] = (uint8_t)m;
      }
    }
  }
}

// GridAgentInterface.render_post (agents.py:92-119): the colour a 'prestige' agent's tile is multiplied with,
// new_color = (prestige_scaled * blue + (1 - prestige_scaled) * red).astype(int) = ((1 - s) * 255, 0, s * 255) truncated,
// s = tanh(prestige / scale) or the logistic function when negative prestige is allowed.  (tanh / exp are libdevice's: within
// 1 ulp of numpy's, which can only show where s * 255 lies within ~1e-13 of an integer.)
__device__ __forceinline__ void prestige_colour(const KP& p, int q, double prestige, uint8_t* __restrict__ out4, bool active) {
  const double x = prestige / p.pscale[q];
  const double s = ((p.prestige_neg >> q) & 1u) ? 1.0 / (1.0 + exp(-x)) : tanh(x);
  out4[0] = (uint8_t)(int)(s * 0.0 + (1.0 - s) * 255.0);
  out4[1] = (uint8_t)(int)(s * 255.0 + (1.0 - s) * 0.0);
  out4[2] = active ? 1 : 0;
  out4[3] = 0;
}

// hide_item_types (agents.py:30, base.py:441-449): after the line of sight has been computed on the real grid, every cell
// of the view whose object `item` (the static object, else the head of the agent queue) is not the observer itself and has
// a hidden type shows `item.agents[0]` instead -- for a static object the queue head, for a head agent the second in the
// queue -- or nothing.  The replacement is a bare agent: no blending with the object underneath, no "own tile on top"
// (render_tile base.py:282-293 looks at the replaced object's own, empty, `agents`).  Rarely used, so written as a plain
// loop over the view cells that follows the reference cell by cell; the fast paths never come here.
template <int OBS, int V, bool BITS>
__device__ __noinline__ void obs_view_hidden(const KP& p, const ObsSmem<V>& o, int view, int a, long long env, const uint32_t* __restrict__ rec,
                                             const uint8_t* __restrict__ tp, const ViewGeom& g_in, const PackedView& pv_in, int orient) {
  // private copies: the byte stores below may alias anything the references point to, and every mask would be reloaded from
  // the caller's stack frame after each of them
  const ViewGeom g = g_in;
  const PackedView pv = pv_in;
  constexpr int VV = V * V;
  const int A = p.A, S = p.S, W = p.W, H = p.H, per_kind = 1 + 4 * A;
  const uint32_t w0 = rec[a * 4];
  bool bad = false;
  unsigned long long agent_cells = 0ull;  // view cells some placed agent stands on (bit 8 vb + va): only those need the queue scan
  for (int q = 0; q < A; ++q) {
    const uint32_t v0 = rec[q * 4];
    int va, vb;
    if (((v0 >> 24) & MG_AF_PLACED) && world_to_view<V>(g, (int)(v0 & 0xFFu), (int)((v0 >> 8) & 0xFFu), va, vb)) agent_cells |= 1ull << (8 * vb + va);
  }
  for (int vb = 0; vb < V; ++vb)
    for (int va = 0; va < V; ++va) {
      const bool vis = pv.visible(va, vb);
      const bool has_agent = ((agent_cells >> (8 * vb + va)) & 1ull) != 0ull;
      if (BITS && (!vis || (!pv.nonempty(va, vb) && !has_agent))) {  // nothing to decide: invisible, or an empty cell without agents
        if (OBS == 2) o.tile[view * VV + vb * V + va] = vis ? (uint8_t)0 : (uint8_t)p.n_tiles;
        continue;
      }
      const int u = g.flip ? V - 1 - vb : vb, v = g.rev ? V - 1 - va : va;
      const int wx = g.topX + (g.vertical ? v : u), wy = g.topY + (g.vertical ? u : v);
      int type = 0, colour = 0, state = 0, head = -1, second = -1;
      if (vis && (unsigned)wx < (unsigned)W && (unsigned)wy < (unsigned)H) {
        const int idx = wx * H + wy;
        // bit-plane worlds: the masks answer for empty cells and canonical walls; only the few other objects are read from
        // the byte planes (global memory on this path)
        if (BITS && !pv.nonempty(va, vb)) type = MG_T_EMPTY;
        else if (BITS && ((((vb < 4 ? pv.cw_lo : pv.cw_hi) >> (8 * (vb & 3) + va)) & 1u) != 0)) { type = MG_T_WALL; colour = MG_C_WORST; state = 0; }
        else { type = tp[idx]; colour = tp[S + idx]; state = tp[2 * S + idx]; }
        uint32_t hs = 0, ss = 0;
        if (has_agent)
        for (int q = 0; q < A; ++q) {  // the cell's queue: placed agents in stamp order
          const uint32_t v0 = rec[q * 4];
          if (!((v0 >> 24) & MG_AF_PLACED) || (int)(v0 & 0xFFu) != wx || (int)((v0 >> 8) & 0xFFu) != wy) continue;
          const uint32_t st = rec[q * 4 + 2];
          if (head < 0 || st < hs) { second = head; ss = hs; head = q; hs = st; }
          else if (second < 0 || st < ss) { second = q; ss = st; }
        }
      }
      const bool on_cell = ((w0 & 0xFFu) == (uint32_t)wx) && (((w0 >> 8) & 0xFFu) == (uint32_t)wy);
      // what the cell shows after hiding: the static object, an agent as the cell's object, or nothing
      bool show_static = false, replaced = false;
      int ag = -1;
      if (type != MG_T_EMPTY) {
        if ((p.hide >> type) & 1u) { replaced = true; ag = head; }             // grid.set(i, j, item.agents[0]) / None, base.py:446-449
        else show_static = true;
      } else if (head >= 0) {
        if (((p.hide >> MG_T_AGENT) & 1u) && head != a) { replaced = true; ag = second; }  // `item is not agent`: the observer never hides itself
        else ag = head;
      }
      if (OBS == 1) {
        if (!vis) continue;  // the staging tile is zero-filled
        uint8_t* oo = o.out + view * (VV * 3) + va * (V * 3) + vb * 3;
        if (show_static) { oo[0] = (uint8_t)type; oo[1] = (uint8_t)colour; oo[2] = (uint8_t)state; }
        else if (ag >= 0) { oo[0] = MG_T_AGENT; oo[1] = p.agent_color[ag]; oo[2] = (uint8_t)((rec[ag * 4] >> 16) & 3u); }
      } else {
        uint8_t t = (uint8_t)p.n_tiles;  // shadow
        if (vis) {
          t = 0;
          int q = -1;
          if (show_static) {
            const int kind = p.kind_of_type[type];
            if (kind == 0xFF) bad = true; else t = (uint8_t)(kind * per_kind);
            if (head >= 0) q = on_cell ? a : head;             // the observer's own tile if it stands there, else agents[0] (base.py:289-293)
          } else if (ag >= 0) q = (!replaced && on_cell) ? a : ag;  // base.py:282-285; a replacement has no agents of its own
          if (q >= 0) {
            const int qd = (int)((rec[q * 4] >> 16) & 3u);
            t = (uint8_t)(t + 1 + 4 * q + ((p.orient_slots == 4) ? qd : ((qd + orient) & 3)));
          }
        }
        o.tile[view * VV + vb * V + va] = t;
      }
    }
  if (OBS == 2 && bad) atomicOr(reinterpret_cast<unsigned int*>(p.envrec) + env * 4 + 3, (unsigned int)MG_ERR_RENDER << 16);
}

// hide_item_types for encoded observations of bit-plane worlds, on the masks (same rules as obs_view_hidden above, base.py:441-449):
// a hidden static object leaves its cell to the head of the agents standing on it (whatever that agent's type says: the cell is
// replaced once); a head agent other than the observer is replaced by the second of its queue when agents are hidden.
template <int V, bool HEADS>
__device__ __forceinline__ void obs_view_hidden_masks(const KP& p, const ObsSmem<V>& o, int view, int a, const uint32_t* __restrict__ rec,
                                                      const uint8_t* __restrict__ tp, const uint32_t* __restrict__ bits, const ViewGeom& g,
                                                      const PackedView& pv, const uint8_t* __restrict__ heads, uint32_t headmask) {
  const int A = p.A, S = p.S;
  uint8_t* out = o.out + view * (V * V * 3);
  const bool hide_agents = ((p.hide >> MG_T_AGENT) & 1u) != 0u;
  uint32_t ne_lo = pv.ne_lo, ne_hi = pv.ne_hi;  // cells that still show a static object after hiding
  if ((p.hide >> MG_T_WALL) & 1u) { ne_lo &= ~pv.cw_lo; ne_hi &= ~pv.cw_hi; }
  else {
    encode_walls<V, 0>(pv.vis_lo & pv.cw_lo, out);
    if (V > 4) encode_walls<V, 4>(pv.vis_hi & pv.cw_hi, out);
  }
  uint32_t g_lo = pv.vis_lo & pv.ne_lo & ~pv.cw_lo, g_hi = pv.vis_hi & pv.ne_hi & ~pv.cw_hi;  // visible objects that are not canonical walls
#pragma unroll
  for (int k = 0; k < OBJ_SLOTS; ++k) {
    const uint32_t e = bits[(OBJ_WORD0 + k) * BS];
    int va, vb;
    if (!(e >> 31) || !world_to_view<V>(g, (int)(e & 15u), (int)((e >> 4) & 15u), va, vb)) continue;
    const uint32_t bit = 1u << (8 * (vb & 3) + va);
    if (vb < 4) { if (!(g_lo & bit)) continue; g_lo &= ~bit; } else { if (!(g_hi & bit)) continue; g_hi &= ~bit; }
    if ((p.hide >> ((e >> 8) & 15u)) & 1u) { if (vb < 4) ne_lo &= ~bit; else ne_hi &= ~bit; continue; }
    uint8_t* oo = out + va * (V * 3) + vb * 3;
    oo[0] = (uint8_t)((e >> 8) & 15u); oo[1] = (uint8_t)((e >> 12) & 15u); oo[2] = (uint8_t)((e >> 16) & 255u);
  }
#pragma unroll
  for (int half = 0; half < 2; ++half) {  // objects that did not fit the list: from the byte planes
    uint32_t m = half ? g_hi : g_lo;
    while (m) {
      const int bitn = __ffs(m) - 1;
      m &= m - 1;
      const int va = bitn & 7, vb = (bitn >> 3) + 4 * half;
      const uint8_t* cp = tp + pv.row0 + vb * pv.ustep + va * pv.vstep;
      const uint32_t type = cp[0];
      if ((p.hide >> type) & 1u) { if (half) ne_hi &= ~(1u << bitn); else ne_lo &= ~(1u << bitn); continue; }
      uint8_t* oo = out + va * (V * 3) + vb * 3;
      oo[0] = (uint8_t)type; oo[1] = cp[S]; oo[2] = cp[2 * S];
    }
  }
  for (int q = 0; q < A; ++q) {  // queue heads: the cell's object where no static object shows
    const uint32_t v0 = rec[q * 4];
    if (HEADS ? !heads[q] : !((headmask >> q) & 1u)) continue;
    int va, vb;
    if (!world_to_view<V>(g, (int)(v0 & 0xFFu), (int)((v0 >> 8) & 0xFFu), va, vb)) continue;
    const uint32_t bit = 1u << (8 * (vb & 3) + va);
    if (!pv.visible(va, vb) || ((vb < 4 ? ne_lo : ne_hi) & bit)) continue;
    int show = q;
    if (hide_agents && q != a && !pv.nonempty(va, vb)) {  // an agent as the cell's object, hidden: the second of its queue, or nothing
      show = -1;
      uint32_t best = 0xFFFFFFFFu;
      for (int r = 0; r < A; ++r) {
        const uint32_t u0 = rec[r * 4];
        if (r == q || !((u0 >> 24) & MG_AF_PLACED) || ((u0 ^ v0) & 0xFFFFu) != 0u) continue;
        if (rec[r * 4 + 2] < best) { best = rec[r * 4 + 2]; show = r; }
      }
      if (show < 0) continue;
    }
    uint8_t* oo = out + va * (V * 3) + vb * 3;
    oo[0] = MG_T_AGENT; oo[1] = p.agent_color[show]; oo[2] = (uint8_t)((rec[show * 4] >> 16) & 3u);
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
  // queue heads: the placed agent with the smallest stamp on a cell is the cell's object or `static_obj.agents[0]`
  // (base.py:547-572).  Taken from the caller (fused kernel) or worked out here from the records.
  uint32_t headmask = 0;
  if (!HEADS && active) {
    for (int q = 0; q < A; ++q) {
      const uint32_t v0 = rec[q * 4];
      bool head = ((v0 >> 24) & MG_AF_PLACED) != 0;
      for (int r = 0; r < A; ++r) {
        const uint32_t u0 = rec[r * 4];
        if (r != q && ((u0 >> 24) & MG_AF_PLACED) && ((u0 ^ v0) & 0xFFFFu) == 0u && rec[r * 4 + 2] < rec[q * 4 + 2]) head = false;
      }
      headmask |= (head ? 1u : 0u) << q;
    }
  }
  const int px = (int)(w0 & 0xFFu), py = (int)((w0 >> 8) & 0xFFu), dir = (int)((w0 >> 16) & 3u);
  const int orient = (3 - dir) & 3;  // view orientation (0 - rot_k) % 4, base.py:130
  if (OBS == 2) {
    if (o.pcol != nullptr && ((p.prestige_mask >> a) & 1u)) prestige_colour(p, a, p.prestige[env * A + a], o.pcol + view * 4, active);  // view == local env * A + a
    o.orient[view] = (uint8_t)((p.orient_slots == 4) ? orient : 0);
    if (!active) {
      const uint8_t shadow = (uint8_t)(p.n_tiles);  // one past the last tile: resolved to the shadow slot when expanding
      for (int i = 0; i < VV; ++i) o.tile[view * VV + i] = shadow;
    }
  }
  if (!active) return;
  const ViewGeom g = view_geom(px, py, dir, V, p.vo, p.W, p.H);
  const PackedView pv = view_masks<V, BITS>(p, tp, bits, g);
  if (p.hide != 0u) {
    if (OBS == 1 && BITS) {  // hide_item_types on the mask path: the same rules as obs_view_hidden, whole rows at a time
      obs_view_hidden_masks<V, HEADS>(p, o, view, a, rec, tp, bits, g, pv, heads, headmask);
      return;
    }
    obs_view_hidden<OBS, V, BITS>(p, o, view, a, env, rec, tp, g, pv, orient);  // the cell-by-cell variant
    return;
  }
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
        const uint32_t e = bits[(OBJ_WORD0 + k) * BS];
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
      if (HEADS ? !heads[q] : !((headmask >> q) & 1u)) continue;
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
      if (HEADS ? !heads[q] : !((headmask >> q) & 1u)) continue;
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
        uint8_t val = o.atlas[(slot * ts + pyy) * ts * 3 + r];
        if (o.pcol != nullptr && t < p.n_tiles) {
          // a tile with a 'prestige'-coloured agent on it (agents.py:92-119): the atlas holds that agent in white, i.e. its
          // triangle's alpha; the pixels are alpha * new_color >> 8, blended over the cell's object like any agent tile
          // (blend_tiles base.py:260-273 with the recoloured tile) -- an inactive agent keeps the cached white tile
          const int per_kind = 1 + 4 * A, kind = t / per_kind, as = t - kind * per_kind;
          if (as > 0) {
            const int q = (as - 1) >> 2;
            const uint8_t* pc = o.pcol + ((view / A) * A + q) * 4;
            if (((p.prestige_mask >> q) & 1u) && pc[2]) {
              const int os = o.orient[view], tb = ts * ts * 3, pix = (pyy * ts) * 3 + (r / 3) * 3, ch = r % 3;
              const uint8_t* white = o.atlas + (as * p.orient_slots + os) * tb + pix;
              const uint8_t* empty = o.atlas + os * tb + pix;  // empty_tile: only a border, for tile sizes > 10 (base.py:245-250)
              const int alpha = (int)white[0] - (int)empty[0];
              const int col[3] = {pc[0], 0, pc[1]};
              const int ag = (alpha * col[ch]) >> 8;
              if (kind == 0) val = (uint8_t)(ag + empty[ch]);
              else {
                const int base = o.atlas[((kind * per_kind) * p.orient_slots + os) * tb + pix + ch];
                const int sa = ((alpha * col[0]) >> 8) + ((alpha * col[2]) >> 8);
                const int am = o.amax[as], M = ((am * col[0]) >> 8) + ((am * col[2]) >> 8);
                val = (uint8_t)(M == 0 ? base : (base * (M - sa) + ag * sa) / M);
              }
            }
          }
        }
        dst[i] = val;
      }
    }
  }
}

}  // namespace mg

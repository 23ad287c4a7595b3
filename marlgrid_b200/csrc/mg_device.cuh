// mg_device.cuh -- device-side building blocks of the batched MarlGrid kernels (sm_100a).
//
// Everything here is integer gather/scatter work: no tensor cores, no floating point except the
// float64 reward (marlgrid/base.py:510,578-580).  Reference citations are relative to
// /root/reference (kandouss/marlgrid).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/marlgrid_b200.h"

namespace mg {

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 counter RNG.  Draw schedule: DESIGN.md "RNG contract" (host statement oracle/philox.py).
// Replaces self.np_random (base.py:373) at base.py:516 (agent order) and base.py:699 (placement).
// ---------------------------------------------------------------------------------------------
struct U4 { uint32_t x, y, z, w; };

__device__ __forceinline__ U4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return U4{c0, c1, c2, c3};
}

constexpr uint32_t TAG_POLICY = 0x20000000u;  // exploration draws of the on-device policy (mg_rollout_policy)
constexpr uint32_t TAG_RESET = 0x80000000u;
constexpr uint32_t TAG_INSTEP = 0x40000000u;

// Sequential placement-try stream of one reset()/step() call (np_random.randint(top, bottom), base.py:699).
// One Philox call yields two tries: (r.x, r.y) then (r.z, r.w).
struct Draws {
  uint32_t g_lo, g_hi, c2, tag, k0, k1;
  uint32_t k;  // tries drawn so far
  U4 r;
  __device__ __forceinline__ void next(int W, int H, int& x, int& y) {
    if ((k & 1u) == 0u) {
      r = philox4x32_10(g_lo, g_hi, c2, tag | (k >> 1), k0, k1);
      x = (int)__umulhi(r.x, (uint32_t)W);
      y = (int)__umulhi(r.y, (uint32_t)H);
    } else {
      x = (int)__umulhi(r.z, (uint32_t)W);
      y = (int)__umulhi(r.w, (uint32_t)H);
    }
    ++k;
  }
  // np_random.randint(top, bottom) on a box (place_obj(top, size), base.py:690-699): the same try, mapped to [x0, x0 + w) x [y0, y0 + h)
  __device__ __forceinline__ void next_box(int x0, int y0, int w, int h, int& x, int& y) {
    next(w, h, x, y);
    x += x0; y += y0;
  }
  // a scalar np_random.randint(lo, hi) (doorkey.py:26,34 `_rand_int`): one try slot, first word
  __device__ __forceinline__ int next_int(int lo, int hi) {
    int x, y;
    next(hi - lo, 1, x, y);
    return lo + x;
  }
};

// ---------------------------------------------------------------------------------------------
// Per-type behaviour tables (objects.py predicates) as bit masks over the type index.
// ---------------------------------------------------------------------------------------------
// can_overlap(): BonusTile, Goal, Floor, Lava always; Door only when open (objects.py:174,216,230,258,327-328).
constexpr uint32_t OVERLAP_ALWAYS = (1u << MG_T_BONUS) | (1u << MG_T_GOAL) | (1u << MG_T_FLOOR) | (1u << MG_T_LAVA);
// can_pickup(): Key, Ball, Box (objects.py:292,314,378)
constexpr uint32_t PICKUP_MASK = (1u << MG_T_KEY) | (1u << MG_T_BALL) | (1u << MG_T_BOX);

__device__ __forceinline__ bool can_overlap_static(int type, int state) {
  return ((OVERLAP_ALWAYS >> type) & 1u) || (type == MG_T_DOOR && state == MG_DOOR_OPEN);
}

// ---------------------------------------------------------------------------------------------
// Line of sight: occlude_mask (agents.py:298-343) on V-bit row masks.
//   T[j] bit i = cell (i, j) of the egocentric view is transparent, M[j] bit i = visible.
// Canonical out-of-bounds semantics: the reference's first upward row j = ay+1 == V
// (view_offset 0) lies outside the array; it reads as zeros, i.e. that row is a no-op.
// The in-place sweeps of the reference are closed-form carry propagations:
//   right sweep (agents.py:305-312): sources = set & transparent cells at i >= ax; a source lights
//     i+1, and keeps doing so through the run of transparent cells: (T + G) ^ T ripples a carry from
//     each generator G through the run and stops on (and lights) the first opaque cell.
//   left sweep (agents.py:314-321): same towards lower i, sources restricted to i in [1, ax+1]
//     (column 0 is never a source -- the reference's left/right asymmetry), done on bit-reversed rows.
// ---------------------------------------------------------------------------------------------
template <int V>
__device__ __forceinline__ uint32_t rev_bits(uint32_t x) { return __brev(x) >> (32 - V); }

template <int V>
__device__ __forceinline__ void sweep_row(uint32_t& row, uint32_t t, uint32_t& next_row, uint32_t ge_ax, uint32_t left_src) {
  constexpr uint32_t RM = (1u << V) - 1u;
  // right sweep
  uint32_t g = row & t & ge_ax;
  row |= ((t + g) ^ t) & RM;
  const uint32_t src_r = row & t & ge_ax;
  next_row |= (src_r | (src_r << 1)) & RM;
  // left sweep
  g = row & t & left_src;
  const uint32_t tr = rev_bits<V>(t), gr = rev_bits<V>(g);
  row |= rev_bits<V>(((tr + gr) ^ tr) & RM);
  const uint32_t src_l = row & t & left_src;
  next_row |= src_l | (src_l >> 1);
}

template <int V>
__device__ __forceinline__ void occlude_rows(const uint32_t (&T)[V], int ax, int ay, uint32_t (&M)[V]) {
  constexpr uint32_t RM = (1u << V) - 1u;
#pragma unroll
  for (int j = 0; j < V; ++j) M[j] = (j == ay) ? (1u << ax) : 0u;
  const uint32_t ge_ax = RM & ~((1u << ax) - 1u);
  const uint32_t left_src = ((1u << (ax + 2)) - 1u) & ~1u & RM;
  // upward pass: j = ay+1 .. 1 writes into row j-1 (agents.py:304-321)
#pragma unroll
  for (int j = V - 1; j >= 1; --j) {
    if (j <= ay + 1) {
      uint32_t nxt = 0;
      sweep_row<V>(M[j], T[j], nxt, ge_ax, left_src);
      M[j - 1] |= nxt;
    }
  }
  // downward pass: j = ay .. V-1 writes into row j+1 when it exists (agents.py:324-341)
#pragma unroll
  for (int j = 0; j < V; ++j) {
    if (j >= ay) {
      uint32_t nxt = 0;
      sweep_row<V>(M[j], T[j], nxt, ge_ax, left_src);
      if (j + 1 < V) M[j + 1 < V ? j + 1 : j] |= nxt;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// TMA-style bulk copy helpers (cp.async.bulk + mbarrier; SASS: UBLKCP / SYNCS)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared bulk copy, completion counted on the mbarrier (bytes, src, dst multiples of 16)
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// shared -> global bulk copy (bulk async-group completion)
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// streaming 16-byte global store (obs are write-once: do not allocate in L1)
__device__ __forceinline__ void st_stream_v4(void* p, const int4& v) {
  asm volatile("st.global.L1::no_allocate.v4.s32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

}  // namespace mg

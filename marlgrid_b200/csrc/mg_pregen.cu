// mg_pregen.cu -- background world generator.  A fresh world (clutter, goal, bonus tiles, agent spawn cells) is a pure function
// of (seed, global env index, episode): it does not depend on what happens in the running episode.  This kernel produces the
// NEXT episode's world of every env ahead of time into `MgState.pregen`; the fused step kernel then regenerates a finished env
// by copying 46 words instead of running Philox + rejection sampling on the step's critical path (which made a step of a
// long-running batch -- ~1 % of the envs finishing in every step -- twice as slow as a step of envs in lock step).
//
// A pass is launched after every 8th fused step of a family (mg_abi.cu) on a low-priority side stream, behind an event edge
// "that step has finished", and runs CONCURRENTLY with the following step kernels (which use half of the SMs' issue slots and
// leave registers / warp slots free); nothing ever waits for a pass:
//   generator:  ep = envrec[e].episode (L2);  if slot.tag != ep + 1:  tag = 0; write world(seed, g, ep), seed;  st.release tag = ep + 1
//   step:       tag = ld.acquire slot.tag;  if tag == ep + 1 and slot.seed == seed: copy the world;  else generate it in place
// Both sides derive the same world, so results never depend on timing.  The generator never overwrites a slot that a step
// kernel may still use: it only writes when the tag is stale, i.e. after the consuming step kernel has published the env's new
// episode number (which it does after reading the slot); episode numbers only grow between two mg_init / load calls, and
// a change of seed drains the side stream first (mg_pregen_drain; env.seed() does it).
#include <mutex>

#include "mg_world.cuh"

namespace mg {

namespace world {
// A world by ONE THREAD: place_obj's rejection sampling try by try, base.py:690-708, on x-line masks kept in a column of shared
// memory (xs[x * 32 + lane]: conflict-free); every lane of the warp works on its own env.  Instruction-wise this is the
// cheapest way to generate worlds (~70 warp instructions per env against ~800 for the latency-oriented warp_sample), at a
// latency nobody waits for.  Writes the slot words (not the tag); returns false if the run was not an ordinary one.
static __device__ __forceinline__ bool lane_world(const KP& p, unsigned long long g, uint32_t ep, uint32_t* __restrict__ slot, uint32_t* __restrict__ xs) {
  const int W = p.W, H = p.H, A = p.A;
  // bit y: canonical wall at (x, y); bit 16 + y: Goal / BonusTile
  const uint32_t fullr = (1u << H) - 1u, endsr = 1u | (1u << (H - 1));
#pragma unroll
  for (int x = 0; x < 16; ++x) xs[x * 32] = (x == 0 || x == W - 1) ? fullr : (x < W ? endsr : 0u);  // wall_rect base.py:172-176
  uint32_t list[OBJ_SLOTS] = {0u, 0u, 0u, 0u};
  int n_listed = 0;
  if (p.goal_mode == MG_GOAL_FIXED) {  // put_obj(Goal) base.py:655-662
    xs[(W - 2) * 32] |= 0x10000u << (H - 2);
    list[0] = obj_entry(W - 2, H - 2, MG_T_GOAL, MG_C_GREEN, 0);
    n_listed = 1;
  }
  const int n_goal = (p.goal_mode == MG_GOAL_RANDOM) ? 1 : 0, n_other = n_goal + p.n_bonus, n_static = n_other + p.n_clutter, n_obj = n_static + A;
  int obj = 0, fails = 0;
  U4 r = U4{0, 0, 0, 0};
  for (uint32_t k = 0; obj < n_obj; ++k) {
    if (k >= 4096u) return false;
    if ((k & 1u) == 0u) r = philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), ep, TAG_RESET | (k >> 1), (uint32_t)p.seed, (uint32_t)(p.seed >> 32));
    const int x = (int)__umulhi((k & 1u) ? r.z : r.x, (uint32_t)W), y = (int)__umulhi((k & 1u) ? r.w : r.y, (uint32_t)H);
    const uint32_t line = xs[x * 32], cb = (line >> y) & 0x10001u;
    const bool agent = obj >= n_static;
    if (agent ? !(cb & 1u) : cb == 0u) {
      if (!agent) {
        xs[x * 32] = line | (((obj < n_other) ? 0x10000u : 1u) << y);
        if (obj < n_other && n_listed < OBJ_SLOTS) {
          const uint32_t e = (obj < n_goal) ? obj_entry(x, y, MG_T_GOAL, MG_C_GREEN, 0) : obj_entry(x, y, MG_T_BONUS, MG_C_YELLOW, obj - n_goal);
#pragma unroll
          for (int q = 0; q < OBJ_SLOTS; ++q) if (q == n_listed) list[q] = e;
          ++n_listed;
        }
      } else slot[PG_XY0 + obj - n_static] = (uint32_t)x | ((uint32_t)y << 8);  // (the tag is down: nobody reads the slot now)
      fails = 0;
      ++obj;
    } else if (++fails >= 32) return false;  // the warp routes give up on such runs too: left to the step kernel's sequential code
  }
  slot[0] = 0u; slot[17] = 0u; slot[18] = 0u; slot[35] = 0u;
  uint32_t xl[16];
#pragma unroll
  for (int x = 0; x < 16; ++x) { xl[x] = xs[x * 32]; slot[LINE_X0 + x] = xl[x]; }
#pragma unroll
  for (int y = 0; y < 16; ++y) {  // the y-lines: the transposed bit matrices (all indices static: registers)
    uint32_t v = 0;
#pragma unroll
    for (int x = 0; x < 16; ++x) v |= ((xl[x] >> y) & 0x10001u) << x;
    slot[LINE_Y0 + y] = v;
  }
#pragma unroll
  for (int k = 0; k < OBJ_SLOTS; ++k) { slot[OBJ_WORD0 + k] = list[k]; slot[OBJ_WORD0 + OBJ_SLOTS + k] = 0u; }
  return true;
}
}  // namespace world

// One pass = two small kernels on the side stream:
//   scan:     thread per env; the envs whose slot is stale (tag != episode + 1, or another seed) are appended to a work list
//             (kept in word 60 of the slots; one warp-aggregated atomic per 32 envs);
//   generate: thread per work-list entry, every lane of a warp on its own env (world::lane_world) -- ~125 warp instructions
//             per world, and a pass in the steady state of a long-running batch (a few thousand stale slots) is a few
//             hundred warps: next to the step kernel's 3 072 resident warps it is invisible.
constexpr int PG_WORK = 60;  // slot word holding work-list entry i (of slot i)

__global__ void __launch_bounds__(256) pregen_scan_kernel(const __grid_constant__ KP p, unsigned int* __restrict__ counter) {
  using namespace world;
  const long long env = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  bool need = false;
  if (env < p.B) {
    const uint32_t ep = (uint32_t)__ldcg(p.envrec + env * 4 + 1);
    const uint4 t = __ldcg(reinterpret_cast<const uint4*>(p.pregen + env * PG_WORDS + PG_WORDS - 4));  // [work, seed lo, seed hi, tag]
    need = (t.w & 0x7FFFFFFFu) != ep + 1u || t.y != (uint32_t)p.seed || t.z != (uint32_t)(p.seed >> 32);
  }
  const uint32_t m = __ballot_sync(0xFFFFFFFFu, need);
  if (m == 0u) return;
  unsigned int base = 0;
  if (lane == 0) base = atomicAdd(counter, (unsigned int)__popc(m));
  base = __shfl_sync(0xFFFFFFFFu, base, 0);
  if (need) p.pregen[(long long)(base + __popc(m & ((1u << lane) - 1u))) * PG_WORDS + PG_WORK] = (uint32_t)env;
}

__global__ void __launch_bounds__(32) pregen_generate_kernel(const __grid_constant__ KP p, const unsigned int* __restrict__ counter) {
  using namespace world;
  __shared__ uint32_t s_x[16 * 32];
  const unsigned int n = *counter;
  for (unsigned int i = blockIdx.x * 32u + threadIdx.x; i < n; i += gridDim.x * 32u) {
    const long long env = (long long)p.pregen[(long long)i * PG_WORDS + PG_WORK];
    const uint32_t ep = (uint32_t)__ldcg(p.envrec + env * 4 + 1);
    uint32_t* const slot = p.pregen + env * PG_WORDS;
    slot[PG_TAG] = 0u;  // the tag comes down before the slot changes ...
    __threadfence();
    const bool ok = lane_world(p, (unsigned long long)(p.env_offset + env), ep, slot, s_x + threadIdx.x);
    slot[PG_SEED_LO] = (uint32_t)p.seed; slot[PG_SEED_HI] = (uint32_t)(p.seed >> 32);
    __threadfence();
    st_release_gpu(slot + PG_TAG, ok ? ep + 1u : (0x80000000u | (ep + 1u)));  // ... and goes up (release) once it is complete; gave up: the step kernel's own routes
  }
}

static unsigned int* pregen_counter() {  // one work-list counter per device
  static unsigned int* ptr[64] = {nullptr};
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  int dev = 0;
  cudaGetDevice(&dev);
  if (ptr[dev & 63] == nullptr && cudaMalloc(&ptr[dev & 63], 256) != cudaSuccess) ptr[dev & 63] = nullptr;
  return ptr[dev & 63];
}

// one pass over the family on `s` (passes of one device must be issued to ONE stream: they share the work-list counter)
int launch_pregen(const KP& p, cudaStream_t s) {
  if (p.pregen == nullptr || p.cellbits == nullptr || p.B <= 0) return 0;
  if (p.scenario != 0 || p.ax0 != 0 || p.ay0 != 0 || p.aw != p.W || p.ah != p.H || p.amax != 100000) return 0;  // only the specialised step kernel consumes slots
  if ((p.goal_mode != MG_GOAL_NONE ? 1 : 0) + p.n_bonus > OBJ_SLOTS) return 0;  // worlds whose objects do not fit the list: the step kernel's sequential route
  if (p.B > 0x7FFFFFFFll) return MG_E_ARG;
  unsigned int* const counter = pregen_counter();
  if (counter == nullptr) return (int)cudaErrorMemoryAllocation;
  int dev = 0;
  cudaGetDevice(&dev);
  static int sms[64] = {0};
  if (!sms[dev & 63]) cudaDeviceGetAttribute(&sms[dev & 63], cudaDevAttrMultiProcessorCount, dev);
  // The generator's CTAs share SMs with the resident CTAs of the step kernel, which runs with the maximum shared-memory
  // carve-out: an SM can only host both if they agree on the carve-out (changing it needs the SM drained).
  static bool configured[64] = {false};
  if (!configured[dev & 63]) {
    cudaFuncSetAttribute(pregen_scan_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(pregen_generate_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    configured[dev & 63] = true;
  }
  cudaError_t e = cudaMemsetAsync(counter, 0, sizeof(unsigned int), s);
  if (e != cudaSuccess) return (int)e;
  pregen_scan_kernel<<<(unsigned)((p.B + 255) / 256), 256, 0, s>>>(p, counter);
  const unsigned grid = (unsigned)std::min<long long>((p.B + 31) / 32, 8ll * sms[dev & 63]);
  pregen_generate_kernel<<<grid, 32, 0, s>>>(p, counter);
  count_launch(); count_launch();
  return (int)cudaGetLastError();
}

// the side stream of the calling thread's current device (lowest priority, non-blocking), created on first use
cudaStream_t pregen_stream() {
  static cudaStream_t streams[64] = {nullptr};
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  int dev = 0;
  cudaGetDevice(&dev);
  cudaStream_t& s = streams[dev & 63];
  if (s == nullptr) {
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);  // lo = numerically greatest = least priority
    if (cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, lo) != cudaSuccess) s = nullptr;
  }
  return s;
}

}  // namespace mg

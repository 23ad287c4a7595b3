// mg_obs_kernels.cu -- observe kernel: one CTA per 32 consecutive envs, one thread per agent view; world state staged
// into shared memory by bulk-async copies (cp.async.bulk + mbarrier: the TMA engine, SASS UBLKCP).
#include "mg_obs.cuh"

namespace mg {

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
  ObsSmem<V> o = obs_smem<V>(reinterpret_cast<uint8_t*>(s_bar + 2), A);  // 16-byte aligned: every block above is a multiple of 16 bytes
  if (OBS == 2 && p.prestige_mask != 0u) {  // behind the atlas and the shadow tile
    o.amax = o.atlas + ((p.n_tiles * p.orient_slots + 1) * p.ts * p.ts * 3 + 15) / 16 * 16;
    o.pcol = o.amax + 64;
  }

  if (tid == 0) mbar_init(s_bar, 1);
  __syncthreads();
  if (tid == 0) {
    const uint32_t wbytes = BITS ? (uint32_t)(ENVS_PER_CTA * BITS_WORDS * 4) : (uint32_t)n_valid * 3u * (uint32_t)S;  // bit-planes: always the whole (transposed) tile
    const uint32_t rbytes = (uint32_t)n_valid * (uint32_t)A * 16u;
    mbar_expect_tx(s_bar, wbytes + rbytes);
    if (BITS) bulk_g2s(s_bits, p.cellbits + env0 * BITS_WORDS, wbytes, s_bar);  // env0 is a multiple of 32: the tile's chunk
    else bulk_g2s(s_grid, p.grid + env0 * 3 * S, wbytes, s_bar);
    bulk_g2s(s_rec, p.agents + env0 * A * 16, rbytes, s_bar);
  }
  obs_prepare<OBS, V>(p, o, tid, nthreads);  // while the copies are in flight
  mbar_wait(s_bar, 0);
  __syncthreads();
  if (tid < n_valid * A) {
    const int le = tid / A, a = tid - le * A;
    obs_view<OBS, V, BITS>(p, o, tid, a, env0 + le, s_rec + le * A * 4, BITS ? p.grid + (env0 + le) * 3 * S : s_grid + le * 3 * S,
                           BITS ? s_bits + le : nullptr);
  }
  fence_proxy_async_smem();  // writer side of the generic -> async proxy hand-over for the bulk copies issued after the barrier
  __syncthreads();
  obs_emit<OBS, V, TSC>(p, o, env0, n_valid, tid, nthreads);
  if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the CTA (and its shared memory) must outlive the read
}

static size_t obs_smem_bytes(const KP& p, int obs) {
  size_t b = (p.cellbits ? (size_t)ENVS_PER_CTA * BITS_WORDS * 4 : (size_t)ENVS_PER_CTA * 3 * p.S) + (size_t)ENVS_PER_CTA * p.A * 16 + 16;
  if (obs == 1) b += (size_t)ENVS_PER_CTA * p.A * p.V * p.V * 3;
  if (obs == 2) {
    b += (size_t)ENVS_PER_CTA * p.A * p.V * p.V + (size_t)((ENVS_PER_CTA * p.A + 15) / 16) * 16;
    b += (size_t)(p.n_tiles * p.orient_slots + 1) * p.ts * p.ts * 3;
    if (p.prestige_mask != 0u) b += 16 + PRESTIGE_SMEM;
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
  if (sm > 227 * 1024) return MG_E_CONFIG;  // tile size / agent count / grid size beyond what one CTA can stage
  if (sm > 48 * 1024 && sm > configured[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (e != cudaSuccess) return (int)e;
    configured[dev & 63] = sm;
  }
  const long long blocks = (p.B + ENVS_PER_CTA - 1) / ENVS_PER_CTA;
  if (blocks <= 0) return 0;
  k<<<(unsigned)blocks, 32 * p.A, sm, s>>>(p);
  count_launch();
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
int launch_obs(const KP& p, int obs, cudaStream_t s) {
  const bool bits = p.cellbits != nullptr;
  if (obs == 1) return bits ? launch_obs_v<1, 0, true>(p, s) : launch_obs_v<1, 0, false>(p, s);
  if (p.prestige_mask != 0u) return bits ? launch_obs_v<2, 0, true>(p, s) : launch_obs_v<2, 0, false>(p, s);  // per-pixel recolouring: the byte path
  if (p.ts == 8) return bits ? launch_obs_v<2, 8, true>(p, s) : launch_obs_v<2, 8, false>(p, s);
  if (p.ts % 4 == 0) return bits ? launch_obs_v<2, 1, true>(p, s) : launch_obs_v<2, 1, false>(p, s);
  return bits ? launch_obs_v<2, 0, true>(p, s) : launch_obs_v<2, 0, false>(p, s);
}

}  // namespace mg

// mg_env_kernels.cu -- per-env kernels (one THREAD per env, world state in place in global memory) and the small
// utility kernels (init, synthetic actions, line-of-sight known-answer).
#include "mg_env.cuh"
#include "mg_world.cuh"

namespace mg {

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
  EnvCtx<ENV_THREADS> c{p, s_rec + threadIdx.x, p.grid + env * 3 * p.S, BITS ? env_bits(p.cellbits, env) : nullptr,
                        s_scr + threadIdx.x, 0, 0, 0, 0u, false};
  c.prest = p.prestige != nullptr ? p.prestige + env * p.A : nullptr;
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

// zero-initialised family of freshly constructed envs (bonus_state = None)
__global__ void init_kernel(uint8_t* grid, uint8_t* agents, int32_t* envrec, uint32_t* cellbits, long long B, int A, int S) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n_grid = B * 3 * S / 16, n_ag = B * A, n_er = B;
  if (i < n_grid) reinterpret_cast<int4*>(grid)[i] = make_int4(0, 0, 0, 0);
  if (cellbits != nullptr && i < (B + 31) / 32 * (BITS_WORDS * BS / 4)) reinterpret_cast<int4*>(cellbits)[i] = make_int4(0, 0, 0, 0);  // whole tiles
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

template <int MODE>
static int launch_env_mode(const KP& p, cudaStream_t s) {
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
  count_launch();
  return (int)cudaGetLastError();
}

// The policy of mg_rollout_policy as a kernel of its own (batches whose tiles do not all fit the resident CTAs of the
// persistent rollout kernel are stepped launch by launch): thread = (env, agent); observations of the step just played ->
// actions of the next step.
template <int NW>
__global__ void policy_kernel(const __grid_constant__ KP p, const uint8_t* __restrict__ obs, int32_t* __restrict__ actions_next) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.B * p.A) return;
  const long long env = i / p.A;
  const int a = (int)(i - env * p.A);
  const uint8_t* o = obs + i * (p.V * p.V * 3);
  const int n_bytes = p.V * p.V * 3;
  actions_next[i] = world::linear_policy_action<NW>(p, a, (unsigned long long)(p.env_offset + env), (uint32_t)p.envrec[env * 4 + 2], [&](int w) {
    uint32_t v = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b)
      if (4 * w + b < n_bytes) v |= (uint32_t)o[4 * w + b] << (8 * b);
    return v;
  });
}

int launch_policy(const KP& p, const uint8_t* obs, int32_t* actions_next, cudaStream_t s) {
  const long long n = p.B * p.A;
  if (n <= 0) return 0;
  const unsigned blocks = (unsigned)((n + 127) / 128);
  switch (p.V) {
    case 3: policy_kernel<7><<<blocks, 128, 0, s>>>(p, obs, actions_next); break;
    case 4: policy_kernel<12><<<blocks, 128, 0, s>>>(p, obs, actions_next); break;
    case 5: policy_kernel<19><<<blocks, 128, 0, s>>>(p, obs, actions_next); break;
    case 6: policy_kernel<27><<<blocks, 128, 0, s>>>(p, obs, actions_next); break;
    case 7: policy_kernel<37><<<blocks, 128, 0, s>>>(p, obs, actions_next); break;
    case 8: policy_kernel<48><<<blocks, 128, 0, s>>>(p, obs, actions_next); break;
    default: return MG_E_CONFIG;
  }
  count_launch();
  return (int)cudaGetLastError();
}

int launch_env(int mode, const KP& p, cudaStream_t s) {
  switch (mode) {
    case 0: return launch_env_mode<0>(p, s);
    case 1: return launch_env_mode<1>(p, s);
    default: return launch_env_mode<2>(p, s);
  }
}

int launch_init(uint8_t* grid, uint8_t* agents, int32_t* envrec, uint32_t* cellbits, long long B, int A, int S, cudaStream_t s) {
  const long long n = std::max<long long>(std::max<long long>(B * 3 * S / 16, B * A), (B + 31) / 32 * (BITS_WORDS * BS / 4));
  init_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(grid, agents, envrec, cellbits, B, A, S);
  count_launch();
  return (int)cudaGetLastError();
}

int launch_random_actions(int32_t* actions, long long n, int n_actions, unsigned long long seed, unsigned long long counter, cudaStream_t s) {
  const long long threads = (n + 3) / 4;
  random_actions_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(actions, n, n_actions, seed, counter);
  count_launch();
  return (int)cudaGetLastError();
}

int launch_los(const uint8_t* transparent, uint8_t* mask, long long n, int view_size, int ax, int ay, cudaStream_t s) {
  const unsigned blocks = (unsigned)((n + 127) / 128);
  switch (view_size) {
    case 3: los_kernel<3><<<blocks, 128, 0, s>>>(transparent, mask, n, ax, ay); break;
    case 4: los_kernel<4><<<blocks, 128, 0, s>>>(transparent, mask, n, ax, ay); break;
    case 5: los_kernel<5><<<blocks, 128, 0, s>>>(transparent, mask, n, ax, ay); break;
    case 6: los_kernel<6><<<blocks, 128, 0, s>>>(transparent, mask, n, ax, ay); break;
    case 7: los_kernel<7><<<blocks, 128, 0, s>>>(transparent, mask, n, ax, ay); break;
    case 8: los_kernel<8><<<blocks, 128, 0, s>>>(transparent, mask, n, ax, ay); break;
    default: return MG_E_CONFIG;
  }
  count_launch();
  return (int)cudaGetLastError();
}

}  // namespace mg

// mg_fused_kernels.cu -- general fused env.step + observe kernel (any agent count / view size, encoded or RGB).
#include "mg_env.cuh"
#include "mg_obs.cuh"

namespace mg {

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
    const uint32_t wbytes = (uint32_t)(ENVS_PER_CTA * BITS_WORDS * 4) /* the whole transposed tile */, rbytes = (uint32_t)n_valid * (uint32_t)A * 16u, ebytes = (uint32_t)n_valid * 16u;
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
  const uint32_t* bits = s_bits + le;
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
    EnvCtx<32> c{p, s_trec + le, tp, s_bits + le, s_scr + le, 0, 0, 0, 0u, false};
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
      EnvCtx<32> c{p, s_trec + le, tp, s_bits + le, s_scr + le, 0, 0, 0, 0u, false};
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

  // own record: publish the head flag (a barrier of its own: the neighbours' observe phase reads this word -- they take the
  // head flags from s_head and ignore the bit, but racecheck rightly calls an unordered write/read pair a hazard)
  if (mine) rec[a * 4] = s_head[tid] ? (rec[a * 4] | (AF_HEAD << 24)) : (rec[a * 4] & ~(AF_HEAD << 24));
  __syncthreads();

  // ---- phase 5: observe the post-step world ----
  if (mine) obs_view<OBS, V, true, true>(p, o, tid, a, env, rec, tp, bits, s_head + le * A);
  fence_proxy_async_smem();  // writer side of the generic -> async proxy hand-over for the bulk copies issued after the barrier
  __syncthreads();
  obs_emit<OBS, V, TSC>(p, o, env0, n_valid, tid, nthreads);
  if (tid == 0) {  // state goes back as it came: contiguous chunks, bulk copies
    fence_proxy_async_smem();
    bulk_s2g(p.agents + env0 * A * 16, s_rec, (uint32_t)n_valid * (uint32_t)A * 16u);
    bulk_s2g(p.envrec + env0 * 4, s_env, (uint32_t)n_valid * 16u);
    bulk_commit();
  }
  if (warp == 0) {  // bit-plane lines changed (reset, plane edit): the tile's (transposed) chunk goes back whole
    const uint32_t dirty = __ballot_sync(0xFFFFFFFFu, lane < n_valid && (s_flag[lane] & FL_BITS_DIRTY));
    if (dirty != 0u && lane == 0) {
      fence_proxy_async_smem();
      bulk_s2g(p.cellbits + env0 * BITS_WORDS, s_bits, (uint32_t)(ENVS_PER_CTA * BITS_WORDS * 4));
      bulk_commit();
    }
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
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
// the specialised kernel (mg_fused2.cuh): bit-plane worlds in ghost mode without spawn delays; respawn is fine (an env in which
// an agent finishes is replayed sequentially)
bool fused2_eligible(const KP& p) {
  if (p.cellbits == nullptr || !(p.flags & MG_F_GHOST)) return false;
  if (p.prestige != nullptr) return false;  // the running reward of 'prestige' agents is kept by the per-env step kernel only
  for (int a = 0; a < p.A; ++a)
    if (p.spawn_delay[a] != 0) return false;
  return true;
}
// the general fused kernel: the same without respawn
bool fused_eligible(const KP& p) { return fused2_eligible(p) && !(p.flags & MG_F_RESPAWN); }

int launch_fused(const KP& p, int obs, cudaStream_t s) {
  if (obs == 1) return launch_fused_v<1, 0>(p, s);
  if (p.ts == 8) return launch_fused_v<2, 8>(p, s);
  return (p.ts % 4 == 0) ? launch_fused_v<2, 1>(p, s) : launch_fused_v<2, 0>(p, s);
}

}  // namespace mg

// mg_abi.cu -- host side of libmarlgrid_b200.so: the C ABI of include/marlgrid_b200.h on top of the kernel launchers.
#include <atomic>
#include <mutex>
#include <unordered_map>

#include "mg_common.cuh"

using namespace mg;

static std::atomic<long long> g_launches{0};
namespace mg {
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
}

#define MG_CUDA(x)                                \
  do {                                            \
    cudaError_t _e = (x);                         \
    if (_e != cudaSuccess) return (int)_e;        \
  } while (0)

static int check_cfg(const MgConfig* c) {
  if (!c) return MG_E_CONFIG;
  if (c->n_agents < 1 || c->n_agents > MG_MAX_AGENTS) return MG_E_CONFIG;
  if (c->view_size < 3 || c->view_size > MG_MAX_VIEW) return MG_E_CONFIG;
  if (c->width < 3 || c->height < 3 || c->width > 255 || c->height > 255) return MG_E_CONFIG;
  if (c->plane_stride % 16 != 0 || c->plane_stride < c->width * c->height) return MG_E_CONFIG;
  if (c->view_offset < 0 || c->view_offset >= c->view_size) return MG_E_CONFIG;
  if (c->max_steps < 1 || c->n_clutter < 0 || c->n_bonus_tiles < 0 || c->n_bonus_tiles > 250) return MG_E_CONFIG;
  // arrival stamps are 16 bits wide and compared raw: an episode issues at most A placement stamps + A per step (twice
  // that with respawn, base.py:629-644), which must not wrap
  if ((long long)(2 * c->max_steps + 1) * c->n_agents >= 65536) return MG_E_CONFIG;
  {  // agent_spawn_kwargs: the sampled box must not be empty (numpy raises on randint(low >= high))
    const int tx = std::max(c->spawn_top[0], 0), ty = std::max(c->spawn_top[1], 0);
    const bool whole = c->spawn_size[0] == 0 && c->spawn_size[1] == 0;
    if (std::min(tx + (whole ? c->width : c->spawn_size[0]), c->width) <= tx || std::min(ty + (whole ? c->height : c->spawn_size[1]), c->height) <= ty) return MG_E_CONFIG;
    if (c->spawn_max_tries < 0 || c->scenario < 0 || c->scenario > MG_SCENARIO_DOORKEY) return MG_E_CONFIG;
    if (c->scenario == MG_SCENARIO_DOORKEY && (c->width < 5 || c->height < 5)) return MG_E_CONFIG;
  }
  // grids wider / taller than 16 cells take the byte-plane kernels, which stage 32 envs' planes in shared memory
  if (c->width > 16 || c->height > 16) {
    const long long sm = 32ll * 3 * c->plane_stride + 32ll * c->n_agents * 16 + 16 + 32ll * c->n_agents * c->view_size * c->view_size * 3;
    if (sm > 227 * 1024) return MG_E_CONFIG;
  }
  return 0;
}

static unsigned long long* device_stats() {  // [2] counters per device, allocated on first use
  static unsigned long long* ptr[64] = {nullptr};
  int dev = 0;
  cudaGetDevice(&dev);
  if (ptr[dev & 63] == nullptr) {
    if (cudaMalloc(&ptr[dev & 63], 2 * sizeof(unsigned long long)) != cudaSuccess) return nullptr;
    cudaMemset(ptr[dev & 63], 0, 2 * sizeof(unsigned long long));
  }
  return ptr[dev & 63];
}

static KP make_kp(const MgConfig* c, const MgState* st) {
  KP p;
  memset(&p, 0, sizeof p);
  p.W = c->width; p.H = c->height; p.A = c->n_agents; p.V = c->view_size; p.vo = c->view_offset; p.ts = c->view_tile_size;
  p.max_steps = c->max_steps; p.n_clutter = c->n_clutter; p.n_bonus = c->n_bonus_tiles; p.goal_mode = c->goal_mode;
  p.flags = c->flags; p.S = c->plane_stride; p.hide = c->hide_types;
  p.goal_reward = c->goal_reward; p.bonus_reward = c->bonus_reward; p.bonus_penalty = c->bonus_penalty;
  for (int i = 0; i < MG_MAX_AGENTS; ++i) { p.agent_color[i] = c->agent_color[i]; p.spawn_delay[i] = c->spawn_delay[i]; }
  for (int i = 0; i < 15; ++i) p.kind_of_type[i] = c->kind_of_type[i];
  p.kind_of_type[15] = 0xFF;
  p.grid = st->grid; p.agents = st->agents; p.envrec = st->envrec; p.B = st->n_envs;
  p.cellbits = (c->width <= 16 && c->height <= 16) ? st->cellbits : nullptr; p.env_offset = st->env_offset; p.seed = st->seed;
  p.n_tiles = (c->n_static_kinds + 1) * (1 + 4 * c->n_agents);
  p.orient_slots = 4;
  p.wall_enc = (uint32_t)MG_T_WALL | ((uint32_t)MG_C_WORST << 8);
  p.ax0 = std::max(c->spawn_top[0], 0); p.ay0 = std::max(c->spawn_top[1], 0);
  const bool whole = c->spawn_size[0] == 0 && c->spawn_size[1] == 0;  // size=None
  p.aw = std::min(p.ax0 + (whole ? c->width : c->spawn_size[0]), c->width) - p.ax0;
  p.ah = std::min(p.ay0 + (whole ? c->height : c->spawn_size[1]), c->height) - p.ay0;
  p.amax = c->spawn_max_tries > 0 ? std::min(c->spawn_max_tries, 100000) : 100000;
  p.scenario = c->scenario;
  p.prestige = st->prestige; p.prestige_mask = c->prestige_mask; p.prestige_neg = c->prestige_neg_mask;
  for (int i = 0; i < MG_MAX_AGENTS; ++i) { p.pbeta[i] = c->prestige_beta[i]; p.pscale[i] = c->prestige_scale[i]; }
  p.pregen = p.cellbits ? st->pregen : nullptr;
  p.stats = p.pregen ? device_stats() : nullptr;
  return p;
}

static int g_force_two_kernels = 0;    // test hook: exercise the per-env step kernel + observe kernel pair
static int g_force_general_fused = 0;  // test hook: exercise the general fused kernel where the specialised one applies

// The generator is launched after every PREGEN_EVERY-th fused step of a family (a pass refills every stale slot it finds; an
// env needs its next world one episode later, so a lag of a few steps costs nothing): per-family step counters, keyed by the
// address of the family's slots.
constexpr int PREGEN_EVERY = 8;
static bool pregen_due(const void* key) {
  static std::mutex mu;
  static std::unordered_map<const void*, unsigned> counters;
  std::lock_guard<std::mutex> lock(mu);
  if (counters.size() > 4096) counters.clear();
  static const int every = getenv("MG_PREGEN_EVERY") ? std::max(1, atoi(getenv("MG_PREGEN_EVERY"))) : PREGEN_EVERY;  // experiments
  return (counters[key]++ % (unsigned)every) == 0;
}
// a ring of events for the "generator pass after this step" edges (a wait captures the record it was called after: an event
// can be recorded again as soon as the wait has been enqueued)
static cudaEvent_t pregen_event() {
  static std::mutex mu;
  static cudaEvent_t ring[64][32] = {{nullptr}};
  static unsigned next[64] = {0};
  std::lock_guard<std::mutex> lock(mu);
  int dev = 0;
  cudaGetDevice(&dev);
  cudaEvent_t& e = ring[dev & 63][next[dev & 63]++ & 31];
  if (e == nullptr && cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) e = nullptr;
  return e;
}
static int g_pregen_auto = 1;           // 0: mg_step* do not launch the background world generator (tests; callers that drive mg_pregen_run themselves)
static cudaEvent_t g_mid_event = nullptr;  // profiling hook: recorded between the two launches of a step

// env.step: one fused launch when eligible; else the step kernel (incl. auto-reset), then the observation
// The background world generator (mg_pregen.cu): one pass on the low-priority side stream, concurrent with the NEXT steps.  It
// starts once the work enqueued on `s` so far has finished (an event edge): the host may be a thousand launches ahead of the
// device, and a pass that ran when it was enqueued would look at the family long before the steps it is meant to follow.
// Nothing ever waits for the pass.
static int pregen_pass_after(const KP& p, cudaStream_t s) {
  cudaStream_t ps = pregen_stream();
  cudaEvent_t ev = pregen_event();
  if (ps == nullptr || ev == nullptr) return 0;
  MG_CUDA(cudaEventRecord(ev, s));
  MG_CUDA(cudaStreamWaitEvent(ps, ev, 0));
  return launch_pregen(p, ps);
}

static int launch_step_obs(const KP& p, int obs, cudaStream_t s) {
  if (obs != 0 && !g_force_two_kernels && fused2_eligible(p)) {
    if (!g_force_general_fused) {  // specialised kernel for the common shapes (mg_fused2.cu); MG_E_UNSUPPORTED = not one of them
      const int e2 = launch_fused2(p, obs, s);
      if (e2 == 0 && p.pregen != nullptr && p.autoreset && g_pregen_auto && pregen_due(p.pregen)) {
        const int e3 = pregen_pass_after(p, s);
        if (e3) return e3;
      }
      if (e2 != MG_E_UNSUPPORTED) return e2;
    }
    if (fused_eligible(p)) return launch_fused(p, obs, s);
  }
  int e = launch_env(0, p, s);
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
  if (c->prestige_mask != 0u && st->prestige == nullptr) return MG_E_ARG;  // a 'prestige'-coloured agent needs its running reward
  if (!aligned16(st->grid) || !aligned16(st->agents) || !aligned16(st->envrec) || !aligned16(st->cellbits) || !aligned16(st->pregen)) return MG_E_ARG;
  return 0;
}

// The K-steps-per-launch entries outside the persistent kernel's reach: one step launch (+ one policy launch) per step on the
// per-step slices.  The kernels store observations as 16-byte vectors; a slice that starts off that grid (batches that are
// not a multiple of 16 envs) is produced in a stream-ordered scratch buffer and copied into place.
static int rollout_step_by_step(const MgConfig* cfg, const MgState* st, const KP& p, int64_t n_steps, bool closed_loop, cudaStream_t s) {
  const int64_t na = st->n_envs * cfg->n_agents, no = mg_obs_bytes_per_env(cfg, 0) * st->n_envs;
  uint8_t* scratch = nullptr;
  if (no % 16 != 0) MG_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&scratch), (size_t)no, s));
  int e = 0;
  for (int64_t t = 0; t < n_steps && !e; ++t) {
    KP q = p;
    uint8_t* const slice = p.obs + t * no;
    const bool direct = (reinterpret_cast<uintptr_t>(slice) & 15u) == 0;
    q.actions = p.actions + t * na; q.rewards = p.rewards + t * na; q.done = p.done + t * st->n_envs; q.obs = direct ? slice : scratch;
    q.pol_w = nullptr;
    e = launch_step_obs(q, 1, s);
    if (!e && !direct) e = (int)cudaMemcpyAsync(slice, scratch, (size_t)no, cudaMemcpyDeviceToDevice, s);
    if (!e && closed_loop && t + 1 < n_steps) e = launch_policy(p, slice, const_cast<int32_t*>(p.actions) + (t + 1) * na, s);
  }
  if (scratch) cudaFreeAsync(scratch, s);
  return e;
}

// A K-steps-per-launch rollout consumes pre-generated worlds like K single steps do (each env's slot holds its NEXT world: one
// reset per env and launch is served, further ones are generated in the kernel), so a generator pass follows every launch of
// eight steps or more; shorter launches count towards the per-family period like single steps.
static int rollout_pregen_pass(const KP& p, int64_t n_steps, cudaStream_t s) {
  if (p.pregen == nullptr || !p.autoreset || !g_pregen_auto) return 0;
  if (n_steps < PREGEN_EVERY && !pregen_due(p.pregen)) return 0;
  return pregen_pass_after(p, s);
}

extern "C" {

int mg_version(void) { return 1; }
const char* mg_build_info(void) { return "marlgrid_b200 sm_100a: persistent fused step+observe kernel (cp.async.bulk + mbarrier staging, 32 envs per tile, PDL), general fused / per-env / observe kernels"; }
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
void mg_debug_force_general_fused(int on) { g_force_general_fused = on; }

int mg_pregen_words_per_env(void) { return 64; }
void mg_pregen_set_auto(int on) { g_pregen_auto = on; }

int mg_pregen_run(const MgConfig* cfg, const MgState* st, mg_stream_t stream) {
  int e = check_state(cfg, st);
  if (e) return e;
  if (st->n_envs == 0 || st->pregen == nullptr) return 0;
  KP p = make_kp(cfg, st);
  return launch_pregen(p, stream == MG_PREGEN_STREAM ? pregen_stream() : (cudaStream_t)stream);
}

int mg_pregen_stats(uint64_t* hits_misses, int reset) {
  unsigned long long* d = device_stats();
  if (d == nullptr || hits_misses == nullptr) return MG_E_ARG;
  MG_CUDA(cudaMemcpy(hits_misses, d, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));  // (synchronises with the device)
  if (reset) MG_CUDA(cudaMemset(d, 0, 2 * sizeof(unsigned long long)));
  return 0;
}

int mg_pregen_drain(void) {
  cudaStream_t ps = pregen_stream();
  if (ps == nullptr) return 0;
  return (int)cudaStreamSynchronize(ps);
}

int mg_init(const MgConfig* cfg, const MgState* st, mg_stream_t stream) {
  int e = check_state(cfg, st);
  if (e) return e;
  if (st->n_envs == 0) return 0;
  if (st->prestige != nullptr) MG_CUDA(cudaMemsetAsync(st->prestige, 0, (size_t)st->n_envs * cfg->n_agents * sizeof(double), (cudaStream_t)stream));
  if (st->pregen != nullptr) {  // no world is ready; a generator pass still in flight on a reused buffer must be over first
    mg_pregen_drain();
    MG_CUDA(cudaMemsetAsync(st->pregen, 0, (size_t)st->n_envs * 64 * sizeof(uint32_t), (cudaStream_t)stream));
  }
  return launch_init(st->grid, st->agents, st->envrec, st->cellbits, st->n_envs, cfg->n_agents, cfg->plane_stride, (cudaStream_t)stream);
}

int mg_sync_derived(const MgConfig* cfg, const MgState* st, mg_stream_t stream) {
  int e = check_state(cfg, st);
  if (e) return e;
  if (st->n_envs == 0) return 0;
  KP p = make_kp(cfg, st);
  return launch_env(2, p, (cudaStream_t)stream);
}

int mg_reset(const MgConfig* cfg, const MgState* st, const uint8_t* reset_mask, mg_stream_t stream) {
  int e = check_state(cfg, st);
  if (e) return e;
  if (st->n_envs == 0) return 0;
  KP p = make_kp(cfg, st);
  p.reset_mask = reset_mask;
  return launch_env(1, p, (cudaStream_t)stream);
}

int mg_step(const MgConfig* cfg, const MgState* st, const int32_t* actions, double* rewards, uint8_t* done, int autoreset,
            mg_stream_t stream) {
  int e = check_state(cfg, st);
  if (e) return e;
  if (!actions || !rewards || !done) return MG_E_ARG;
  if (st->n_envs == 0) return 0;
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

int mg_rollout_persistent(const MgConfig* cfg, const MgState* st, const int32_t* actions, int64_t n_steps, double* rewards, uint8_t* done,
                          uint8_t* obs, int autoreset, mg_stream_t stream) {
  int e = check_state(cfg, st);
  if (e) return e;
  if (!actions || !rewards || !done || !obs || !aligned16(obs) || n_steps < 0 || n_steps > 0x7FFFFFFF) return MG_E_ARG;
  if (n_steps == 0 || st->n_envs == 0) return 0;
  KP p = make_kp(cfg, st);
  p.actions = actions; p.rewards = rewards; p.done = done; p.obs = obs; p.autoreset = autoreset;
  if (!g_force_two_kernels && !g_force_general_fused) {
    e = launch_fused2_rollout(p, (int)n_steps, (cudaStream_t)stream);  // ONE launch, the tiles' state stays in shared memory
    if (e == 0) return rollout_pregen_pass(p, n_steps, (cudaStream_t)stream);
    if (e != MG_E_UNSUPPORTED) return e;
  }
  return rollout_step_by_step(cfg, st, p, n_steps, false, (cudaStream_t)stream);  // shapes / batch sizes outside the persistent kernel's reach
}

// Closed-loop rollout on the device: step t + 1 plays the actions the policy chose from step t's observations (see the header).
int mg_rollout_policy(const MgConfig* cfg, const MgState* st, const MgLinearPolicy* pol, int64_t n_steps, int32_t* actions, double* rewards,
                      uint8_t* done, uint8_t* obs, int autoreset, mg_stream_t stream) {
  int e = check_state(cfg, st);
  if (e) return e;
  if (!pol || !pol->weights || !pol->bias || pol->n_actions < 1 || pol->n_actions > 7) return MG_E_ARG;
  if (!actions || !rewards || !done || !obs || !aligned16(obs) || !aligned16(pol->weights) || !aligned16(pol->bias) || n_steps < 0 || n_steps > 0x7FFFFFFF) return MG_E_ARG;
  if (n_steps == 0 || st->n_envs == 0) return 0;
  KP p = make_kp(cfg, st);
  p.actions = actions; p.rewards = rewards; p.done = done; p.obs = obs; p.autoreset = autoreset;
  p.pol_w = reinterpret_cast<const int32_t*>(pol->weights); p.pol_b = pol->bias; p.pol_n = pol->n_actions; p.pol_eps = pol->epsilon; p.pol_seed = pol->seed;
  if (!g_force_two_kernels && !g_force_general_fused) {
    e = launch_fused2_rollout(p, (int)n_steps, (cudaStream_t)stream);  // ONE launch: state and policy hand-off stay on the SMs
    if (e == 0) return rollout_pregen_pass(p, n_steps, (cudaStream_t)stream);
    if (e != MG_E_UNSUPPORTED) return e;
  }
  return rollout_step_by_step(cfg, st, p, n_steps, true, (cudaStream_t)stream);  // a step launch + a policy launch per step
}

// The policy alone: actions the policy chooses from `obs` (the observations of the step just played) -- what the host loop
// `act = agents.action_step(obs)` (README.md:43-57) calls between two steps.
int mg_policy_act(const MgConfig* cfg, const MgState* st, const MgLinearPolicy* pol, const uint8_t* obs, int32_t* actions, mg_stream_t stream) {
  int e = check_state(cfg, st);
  if (e) return e;
  if (!pol || !pol->weights || !pol->bias || pol->n_actions < 1 || pol->n_actions > 7 || !aligned16(pol->weights) || !aligned16(pol->bias) || !obs || !actions) return MG_E_ARG;
  if (st->n_envs == 0) return 0;
  KP p = make_kp(cfg, st);
  p.pol_w = reinterpret_cast<const int32_t*>(pol->weights); p.pol_b = pol->bias; p.pol_n = pol->n_actions; p.pol_eps = pol->epsilon; p.pol_seed = pol->seed;
  return launch_policy(p, obs, actions, (cudaStream_t)stream);
}

int mg_rollout_fused_rr(const MgConfig* cfg, const MgState* states, int n_states, const int32_t* actions, int64_t n_steps,
                         double* const* rewards, uint8_t* const* done, uint8_t* const* obs, int autoreset, mg_stream_t stream) {
  if (!states || n_states < 1 || !actions || !rewards || !done || !obs) return MG_E_ARG;
  for (int r = 0; r < n_states; ++r) {
    int e = check_state(cfg, &states[r]);
    if (e) return e;
    if (!rewards[r] || !done[r] || !obs[r] || !aligned16(obs[r]) || states[r].n_envs != states[0].n_envs) return MG_E_ARG;
  }
  const int64_t per_step = states[0].n_envs * cfg->n_agents;
  for (int64_t t = 0; t < n_steps; ++t) {
    const int r = (int)(t % n_states);
    KP p = make_kp(cfg, &states[r]);
    p.actions = actions + t * per_step; p.rewards = rewards[r]; p.done = done[r]; p.obs = obs[r]; p.autoreset = autoreset;
    int e = launch_step_obs(p, 1, (cudaStream_t)stream);
    if (e) return e;
  }
  return 0;
}

int mg_random_actions(int32_t* actions, int64_t n, int n_actions, uint64_t seed, uint64_t counter, mg_stream_t stream) {
  if (!actions || n < 0 || n_actions < 1) return MG_E_ARG;
  if (n == 0) return 0;
  return launch_random_actions(actions, n, n_actions, seed, counter, (cudaStream_t)stream);
}

int mg_los_batch(const uint8_t* transparent, uint8_t* mask, int64_t n, int view_size, int ax, int ay, mg_stream_t stream) {
  if (!transparent || !mask || n < 0 || ax < 0 || ay < 0 || ax >= view_size || ay >= view_size) return MG_E_ARG;
  if (n == 0) return 0;
  return launch_los(transparent, mask, n, view_size, ax, ay, (cudaStream_t)stream);
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
  cudaStream_t copy_stream;    // device -> host copies of a finished slice run here, under the next slice's kernel
  cudaEvent_t slice_done[8];
};

constexpr int ENGINE_SLICES = 4;  // mg_engine_step cuts the batch into this many env ranges (multiples of 32 envs)


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
  MG_CUDA(cudaStreamCreateWithFlags(&en->copy_stream, cudaStreamNonBlocking));
  for (int i = 0; i < 8; ++i) MG_CUDA(cudaEventCreateWithFlags(&en->slice_done[i], cudaEventDisableTiming));
  MG_CUDA(cudaMalloc(&en->st.grid, (size_t)n_envs * 3 * cfg->plane_stride));
  MG_CUDA(cudaMalloc(&en->st.agents, (size_t)n_envs * cfg->n_agents * MG_AGENT_REC));
  MG_CUDA(cudaMalloc(&en->st.envrec, (size_t)n_envs * MG_ENV_REC));
  MG_CUDA(cudaMalloc(&en->st.cellbits, (size_t)((n_envs + 31) / 32) * (BITS_WORDS * BS * 4)));  // whole tiles of 32 envs
  MG_CUDA(cudaMalloc(&en->st.pregen, (size_t)n_envs * 64 * sizeof(uint32_t)));
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
  // MultiGridEnv.__init__ ends with self.reset() (base.py:368): the engine can be stepped right away; the episode counter
  // (Philox counter of the placement draws) is rewound so that mg_engine_reset() as first call regenerates the same world
  e = mg_reset(&en->cfg, &en->st, nullptr, en->stream);
  if (e) return e;
  MG_CUDA(cudaMemset2DAsync(en->st.envrec + 1, MG_ENV_REC, 0, sizeof(int32_t), (size_t)n_envs, en->stream));
  MG_CUDA(cudaStreamSynchronize(en->stream));
  *out = en;
  return 0;
}

void mg_engine_destroy(MgEngine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  cudaStreamSynchronize(e->stream);
  mg_pregen_drain();  // a generator pass may still be reading / writing this engine's buffers
  cudaFree(e->st.grid); cudaFree(e->st.agents); cudaFree(e->st.envrec); cudaFree(e->st.cellbits); cudaFree(e->st.pregen);
  cudaFree(e->d_actions); cudaFree(e->d_rewards); cudaFree(e->d_done); cudaFree(e->d_obs); cudaFree(e->d_atlas);
  cudaStreamSynchronize(e->copy_stream);
  for (int i = 0; i < 8; ++i) cudaEventDestroy(e->slice_done[i]);
  cudaStreamDestroy(e->copy_stream);
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
  // The step is PCIe-bound (the observations leaving the device): the batch is cut into env ranges, and while range k's
  // results cross the bus on the copy stream, range k+1's actions come in and its kernel runs on the compute stream.
  const int64_t B = e->st.n_envs;
  const int A = e->cfg.n_agents;
  const int64_t per_env_obs = e->obs_bytes / B;
  const int n_slices = B >= 4096 ? ENGINE_SLICES : 1;
  const int64_t slice = ((B + n_slices - 1) / n_slices + 31) / 32 * 32;
  int k = 0;
  for (int64_t b0 = 0; b0 < B; b0 += slice, ++k) {
    const int64_t nb = std::min(slice, B - b0);
    MgState st = e->st;
    st.grid += b0 * 3 * e->cfg.plane_stride; st.agents += b0 * A * MG_AGENT_REC; st.envrec += b0 * 4; st.cellbits += b0 * BITS_WORDS; st.pregen += b0 * 64;
    st.n_envs = nb; st.env_offset = e->st.env_offset + b0;
    int32_t* d_act = e->d_actions + b0 * A;
    double* d_rew = e->d_rewards + b0 * A;
    uint8_t* d_done = e->d_done + b0;
    uint8_t* d_obs = e->d_obs + b0 * per_env_obs;
    MG_CUDA(cudaMemcpyAsync(d_act, actions_host + b0 * A, (size_t)nb * A * sizeof(int32_t), cudaMemcpyHostToDevice, e->stream));
    int r = e->rgb ? mg_step_fused_rgb(&e->cfg, &st, d_act, d_rew, d_done, e->d_atlas, d_obs, autoreset, e->stream)
                   : mg_step_fused(&e->cfg, &st, d_act, d_rew, d_done, d_obs, autoreset, e->stream);
    if (r) return r;
    MG_CUDA(cudaEventRecord(e->slice_done[k], e->stream));
    MG_CUDA(cudaStreamWaitEvent(e->copy_stream, e->slice_done[k], 0));
    if (obs_host) MG_CUDA(cudaMemcpyAsync(obs_host + b0 * per_env_obs, d_obs, (size_t)(nb * per_env_obs), cudaMemcpyDeviceToHost, e->copy_stream));
    if (rewards_host) MG_CUDA(cudaMemcpyAsync(rewards_host + b0 * A, d_rew, (size_t)nb * A * sizeof(double), cudaMemcpyDeviceToHost, e->copy_stream));
    if (done_host) MG_CUDA(cudaMemcpyAsync(done_host + b0, d_done, (size_t)nb, cudaMemcpyDeviceToHost, e->copy_stream));
  }
  MG_CUDA(cudaStreamSynchronize(e->copy_stream));
  MG_CUDA(cudaStreamSynchronize(e->stream));
  return 0;
}

// The transfers of mg_engine_step without the kernel: the copy ceiling of the box for this batch (bench.py e2e.copy_ceiling).
int mg_engine_copy_only(MgEngine* e, const int32_t* actions_host, uint8_t* obs_host, double* rewards_host, uint8_t* done_host) {
  if (!e || !actions_host) return MG_E_ARG;
  MG_CUDA(cudaSetDevice(e->device));
  const int64_t B = e->st.n_envs;
  const int A = e->cfg.n_agents;
  const int64_t per_env_obs = e->obs_bytes / B;
  const int n_slices = B >= 4096 ? ENGINE_SLICES : 1;
  const int64_t slice = ((B + n_slices - 1) / n_slices + 31) / 32 * 32;
  int k = 0;
  for (int64_t b0 = 0; b0 < B; b0 += slice, ++k) {
    const int64_t nb = std::min(slice, B - b0);
    MG_CUDA(cudaMemcpyAsync(e->d_actions + b0 * A, actions_host + b0 * A, (size_t)nb * A * sizeof(int32_t), cudaMemcpyHostToDevice, e->stream));
    MG_CUDA(cudaEventRecord(e->slice_done[k], e->stream));
    MG_CUDA(cudaStreamWaitEvent(e->copy_stream, e->slice_done[k], 0));
    if (obs_host) MG_CUDA(cudaMemcpyAsync(obs_host + b0 * per_env_obs, e->d_obs + b0 * per_env_obs, (size_t)(nb * per_env_obs), cudaMemcpyDeviceToHost, e->copy_stream));
    if (rewards_host) MG_CUDA(cudaMemcpyAsync(rewards_host + b0 * A, e->d_rewards + b0 * A, (size_t)nb * A * sizeof(double), cudaMemcpyDeviceToHost, e->copy_stream));
    if (done_host) MG_CUDA(cudaMemcpyAsync(done_host + b0, e->d_done + b0, (size_t)nb, cudaMemcpyDeviceToHost, e->copy_stream));
  }
  MG_CUDA(cudaStreamSynchronize(e->copy_stream));
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

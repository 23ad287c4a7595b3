/*
 * marlgrid_b200.h -- C ABI of the B200-native batched MarlGrid hot path.
 *
 * The reference (kandouss/marlgrid, pure Python) has NO FFI; its boundary for this path is the
 * gym surface of MultiGridEnv (marlgrid/base.py:334-653).  This header is the boundary a
 * maintainer would bind instead (ctypes stub: INTEGRATION.md).  Every entry point names the
 * reference code it replaces.  All functions are `extern "C"`, take plain pointers and sizes,
 * never throw, never allocate behind the caller's back (except the mg_engine_* family, which
 * owns its device buffers by design), and return 0 on success or a CUDA error code (>0) /
 * MG_E_* (<0).
 *
 * World state for B independent env instances is structure-of-arrays in device memory:
 *
 *   grid    uint8 [B][3][S]     three planes per env: object TYPE index, COLOUR index, STATE
 *                               (marlgrid/objects.py:31-43 OBJECT_TYPES order, :11-29 COLORS order,
 *                               WorldObj.encode objects.py:90-99).  Cell (x, y) of a plane lives at
 *                               x*H + y, i.e. the reference's `MultiGrid.grid[i, j]` (base.py:91).
 *                               S = plane_stride >= W*H, a multiple of 16 B (bulk-copy granularity).
 *                               Agents are NOT stored in the planes; they are an overlay:
 *   agents  uint8 [B][A][16]    per-agent record (GridAgentInterface state, marlgrid/agents.py:155-170):
 *                                 +0 x  +1 y  +2 dir  +3 flags(MG_AF_*)  +4 carry_type  +5 carry_colour
 *                                 +6 carry_state  +7 bonus_state (0xFF = None)  +8 int32 stamp (arrival
 *                                 order inside the episode: queue position of stacked agents,
 *                                 base.py:547-572)  +12 reserved
 *   envrec  int32 [B][4]        +0 step_count (base.py:414,512)  +1 episode (resets so far)
 *                               +2 lifetime step() calls  +3 lo16 = next stamp, hi16 = MG_ERR_* bits
 *
 * Randomness is a counter-based Philox4x32-10 stream keyed by (seed, global env index); the exact
 * draw schedule is stated in oracle/philox.py and DESIGN.md.  It replaces `self.np_random`
 * (base.py:373) at its two call sites, base.py:516 (per-step agent order) and :699 (placement).
 */
#ifndef MARLGRID_B200_H
#define MARLGRID_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MG_MAX_AGENTS 8
#define MG_MAX_VIEW 8
#define MG_AGENT_REC 16
#define MG_ENV_REC 16

/* object type indices: position in OBJECT_TYPES (marlgrid/objects.py:31-43; agents.py:9) */
enum {
  MG_T_EMPTY = 0, MG_T_GRIDAGENT = 1, MG_T_BULK = 2, MG_T_BONUS = 3, MG_T_GOAL = 4, MG_T_FLOOR = 5,
  MG_T_EMPTYSPACE = 6, MG_T_LAVA = 7, MG_T_WALL = 8, MG_T_KEY = 9, MG_T_BALL = 10, MG_T_DOOR = 11,
  MG_T_BOX = 12, MG_T_AGENT = 13, MG_N_TYPES = 14
};
/* colour indices: key order of COLORS (marlgrid/objects.py:11-29) */
enum {
  MG_C_RED = 0, MG_C_ORANGE = 1, MG_C_GREEN = 2, MG_C_BLUE = 3, MG_C_CYAN = 4, MG_C_PURPLE = 5,
  MG_C_YELLOW = 6, MG_C_OLIVE = 7, MG_C_GREY = 8, MG_C_WORST = 9, MG_C_PINK = 10, MG_C_WHITE = 11,
  MG_C_PRESTIGE = 12, MG_C_SHADOW = 13
};
/* Door.states (marlgrid/objects.py:325) */
enum { MG_DOOR_OPEN = 1, MG_DOOR_CLOSED = 2, MG_DOOR_LOCKED = 3 };
/* GridAgentInterface.actions (marlgrid/agents.py:10-17) */
enum { MG_A_LEFT = 0, MG_A_RIGHT = 1, MG_A_FORWARD = 2, MG_A_PICKUP = 3, MG_A_DROP = 4, MG_A_TOGGLE = 5, MG_A_DONE = 6 };

/* agent flags (record byte +3) */
#define MG_AF_PLACED 1u  /* pos is not None: the agent is somewhere in the grid */
#define MG_AF_ACTIVE 2u  /* GridAgentInterface.active (agents.py:155-159) */
#define MG_AF_DONE 4u    /* GridAgentInterface.done (base.py:584-585) */
#define MG_AF_HEAD 128u  /* scratch bit: some kernels leave "head of its cell's queue" here; readers must ignore it (the queue
                            order is defined by the stamps: the placed agent with the smallest stamp on a cell is the cell's
                            object or `static_obj.agents[0]` of the reference, base.py:547-572) */

/* MgConfig.flags */
#define MG_F_GHOST 1u           /* ghost_mode (base.py:345,541-542,683-684) */
#define MG_F_RESPAWN 2u         /* respawn (base.py:344,629-644) */
#define MG_F_REWARD_DECAY 4u    /* reward_decay (base.py:342,578-579) */
#define MG_F_SEE_THROUGH 8u     /* see_through_walls (agents.py:292-295) */
#define MG_F_BONUS_INITIAL 16u  /* BonusTile.initial_reward (objects.py:203) */
#define MG_F_BONUS_RESET 32u    /* BonusTile.reset_on_mistake (objects.py:200) */

/* goal_mode */
enum { MG_GOAL_NONE = 0, MG_GOAL_FIXED = 1, MG_GOAL_RANDOM = 2 };

/* error bits accumulated in envrec word 3 (hi16): what the reference would have raised */
#define MG_ERR_BAD_ACTION 1u  /* ValueError("Environment can't handle action") base.py:619-620 */
#define MG_ERR_PLACEMENT 2u   /* RecursionError("Rejection sampling failed") base.py:706 */
#define MG_ERR_STACK 4u       /* AssertionError / ValueError("?!?!?!") base.py:558,568-569 */
#define MG_ERR_TOGGLE 8u      /* TypeError from Box.toggle(self) objects.py:381 */
#define MG_ERR_RENDER 16u     /* object whose render() raises in the reference (objects.py:274-277,309-321,370) */
#define MG_ERR_PRESTIGE 32u   /* AttributeError: GridAgentInterface.reward() with allow_negative_prestige does `self.rew += rew` on an
                                 attribute that is never created (agents.py:146-148) */

/* negative return codes */
#define MG_E_CONFIG (-1)
#define MG_E_ARG (-2)

/* Static description of one env family: MultiGridEnv ctor kwargs (base.py:335-347), the scenario
 * generator (envs/empty.py:9-16, envs/cluttered.py:9-36, envs/goalcycle.py:9-51) and the agent
 * interface geometry (agents.py:19-35).  POD, passed by pointer, copied by the callee. */
typedef struct MgConfig {
  int32_t width, height;     /* grid_size | width/height (base.py:349-358) */
  int32_t n_agents;          /* len(agents) <= MG_MAX_AGENTS */
  int32_t view_size;         /* GridAgentInterface.view_size (agents.py:21), 3..MG_MAX_VIEW */
  int32_t view_offset;       /* agents.py:23 */
  int32_t view_tile_size;    /* agents.py:22; RGB path only */
  int32_t max_steps;         /* base.py:341 */
  int32_t n_clutter;         /* cluttered.py:15-18 (already resolved from clutter_density) */
  int32_t n_bonus_tiles;     /* goalcycle.py:9 */
  int32_t goal_mode;         /* MG_GOAL_* : empty.py:12 / cluttered.py:28-31 */
  uint32_t flags;            /* MG_F_* */
  int32_t plane_stride;      /* S, bytes per plane per env, multiple of 16, >= width*height */
  double goal_reward;        /* Goal(reward=1) empty.py:12, cluttered.py:29,31 */
  double bonus_reward;       /* goalcycle.py:9 reward */
  double bonus_penalty;      /* goalcycle.py:9 penalty */
  uint8_t agent_color[MG_MAX_AGENTS];  /* colour index per agent (envs/__init__.py:31) */
  int32_t spawn_delay[MG_MAX_AGENTS];  /* agents.py:34 */
  uint8_t n_static_kinds;    /* RGB: number of distinct static tile kinds in the atlas (excl. empty) */
  uint8_t kind_of_type[15];  /* RGB: type index -> atlas kind (0 = none/empty, 0xFF = undefined render) */
  uint32_t hide_types;       /* GridAgentInterface.hide_item_types (agents.py:30, base.py:441-449) as a bit set over the type
                                indices (bit MG_T_AGENT = 'Agent'): objects of these types are masked out of every agent's
                                observation (after the line of sight has been computed with them in place) */
  int32_t spawn_top[2];      /* agent_spawn_kwargs (base.py:346,409-412,505,642): place_obj(agent, top=..., size=..., max_tries=...) */
  int32_t spawn_size[2];     /*   (base.py:690-696); size {0, 0} = None = the whole grid.  Agents are then sampled in the box  */
  int32_t spawn_max_tries;   /*   [max(top,0), min(top+size, grid)); 0 = the default 1e5 (`reject_fn`, a Python callable, is not supported) */
  int32_t scenario;          /* MG_SCENARIO_*: which _gen_grid builds the world */
  uint32_t prestige_mask;    /* bit a: agent a is coloured 'prestige' (agents.py:99): its tile is recoloured from its running reward */
  uint32_t prestige_neg_mask;/* bit a: allow_negative_prestige (agents.py:103-106,146-153) */
  double prestige_beta[MG_MAX_AGENTS];   /* agents.py:49,144 */
  double prestige_scale[MG_MAX_AGENTS];  /* agents.py:50,104-106 */
} MgConfig;

/* scenario generators */
enum { MG_SCENARIO_STANDARD = 0 /* empty.py / cluttered.py / goalcycle.py, selected by goal_mode, n_clutter, n_bonus_tiles */,
       MG_SCENARIO_DOORKEY = 1  /* doorkey.py:15-41 with `_rand_int` = np_random.randint (the reference's class cannot be built: the
                                   method is missing there) */ };

/* Device pointers of the SoA world state (caller-owned, e.g. torch tensors). */
typedef struct MgState {
  uint8_t* grid;    /* [B][3][S] */
  uint8_t* agents;  /* [B][A][16] */
  int32_t* envrec;  /* [B][4] */
  uint32_t* cellbits; /* [ceil(B/32)][44][32] DERIVED bit-planes, or NULL (then, and for grids wider/taller than 16, the
                         observe kernel stages and gathers the byte planes instead).  Two bits per cell: OP = opaque (Wall /
                         Door not open), OT = "other" = non-empty and not a canonical wall Wall('worst', 0); canonical wall ==
                         OP & ~OT, non-empty == OP | OT.  Per env 44 words, one word per line (bits 0..15 OP, bits 16..31 OT):
                           word 1+x   (x = 0..15): cells (x, 0..15), bit = y        words 0, 17, 18, 35 are zero guard lines
                           word 19+y  (y = 0..15): cells (0..15, y), bit = x
                           word 36+k  (k = 0..3) : object list, x | y<<4 | type<<8 | colour<<12 | state<<16 | 1<<31: the first OT
                                                   objects (Goal, BonusTiles, Keys ...); unlisted ones are read from the byte planes
                           word 40..43: reserved (zero)
                         TILE-TRANSPOSED in memory: the 32 consecutive envs of a tile interleave their words -- word w of env e
                         lives at cellbits[((e / 32) * 44 + w) * 32 + e % 32] -- so that a tile is one contiguous 5 632-byte
                         chunk whose shared-memory image is bank-conflict-free for lane = env, and a thread-per-env kernel
                         accesses it coalesced.  Allocate whole tiles (ceil(B/32) * 5 632 bytes); sub-ranges of a batch handed
                         to the library must start at a multiple of 32 envs (or pass NULL).
                         Maintained by mg_reset / mg_step*; after editing planes or agent records by hand call
                         mg_sync_derived. */
  int64_t n_envs;   /* B (envs on THIS device) */
  int64_t env_offset; /* global index of local env 0 (RNG is keyed by the global index) */
  uint64_t seed;
  uint32_t* pregen; /* [B][64] or NULL: PRE-GENERATED NEXT WORLDS.  A fresh world (walls, goal, bonus tiles, spawn cells) is a pure
                       function of (seed, global env index, episode number), so the library produces every env's next one ahead
                       of time -- a background kernel on a low-priority side stream, launched after each mg_step_fused* /
                       mg_rollout_* call, running concurrently with the following steps (no ordering edge; a tag per slot + a
                       seqlock decide whether a step kernel copies the slot or generates the world itself: both give the same
                       world, results never depend on timing).  It takes reset()'s Philox + rejection sampling off the step's
                       critical path.  Zeroed by mg_init.  Before freeing or re-purposing the buffer call mg_pregen_drain(). */
  double* prestige; /* [B][A] or NULL: GridAgentInterface.prestige (agents.py:141-153,168), the running discounted reward that colours
                       a color='prestige' agent's tile (agents.py:92-119).  Required (non-NULL) when MgConfig.prestige_mask != 0;
                       such families take the per-env step kernel + observe kernel (the fused kernels do not carry the state). */
} MgState;

typedef void* mg_stream_t; /* cudaStream_t */

/* ---- library / build info ------------------------------------------------------------------ */
int mg_version(void);                 /* ABI version */
const char* mg_build_info(void);      /* "sm_100a ..." */
int mg_sizeof_config(void);           /* sizeof(MgConfig) as compiled: binding self-check */
int mg_config_validate(const MgConfig* cfg);
int64_t mg_obs_bytes_per_env(const MgConfig* cfg, int rgb);

/* ---- device-pointer API (all pointers are DEVICE pointers; async on `stream`) -------------- */

/* Zero-initialise state as a freshly constructed env family (all agents unplaced, dir 0,
 * episode 0).  Replaces MultiGridEnv.__init__ state setup (base.py:353-367, agents.py:90). */
int mg_init(const MgConfig* cfg, const MgState* st, mg_stream_t stream);

/* Pre-generated worlds (MgState.pregen): words per env; one explicit generator pass on `stream` (MG_PREGEN_STREAM = the
 * library's own side stream); wait for the side stream (before freeing the buffers a pass may still touch); switch the
 * automatic launch after every fused step off / on (default on). */
#define MG_PREGEN_STREAM ((mg_stream_t)(intptr_t)-1)
int mg_pregen_words_per_env(void);
int mg_pregen_run(const MgConfig* cfg, const MgState* st, mg_stream_t stream);
int mg_pregen_drain(void);
void mg_pregen_set_auto(int on);
/* counters of the current device since load / the last reset: [0] envs a step kernel regenerated by copying a pre-generated
 * world, [1] envs it had to generate itself (synchronises with the device) */
int mg_pregen_stats(uint64_t* hits_misses, int reset);

/* Recompute the derived state (MgState.cellbits) from the planes and agent records. */
int mg_sync_derived(const MgConfig* cfg, const MgState* st, mg_stream_t stream);

/* Start a new episode in every env (reset_mask == NULL) or in envs whose mask byte != 0.
 * Replaces MultiGridEnv.reset (base.py:402-416) + _gen_grid (empty.py:9-16, cluttered.py:25-36,
 * goalcycle.py:30-51) + place_obj/try_place_obj (base.py:664-708). */
int mg_reset(const MgConfig* cfg, const MgState* st, const uint8_t* reset_mask, mg_stream_t stream);

/* One env.step() for every env: MultiGridEnv.step (base.py:501-649) without the obs
 * (actions int32 [B][A]; rewards float64 [B][A]; done uint8 [B]).  autoreset != 0 starts a new
 * episode inside the kernel for envs that finished (the reference has no such mode; it is
 * `obs,r,d,_ = env.step(a); if d: obs = env.reset()` of the caller loop). */
int mg_step(const MgConfig* cfg, const MgState* st, const int32_t* actions, double* rewards,
            uint8_t* done, int autoreset, mg_stream_t stream);

/* Encoded egocentric observation uint8 [B][A][V][V][3]:
 * gen_obs_grid (base.py:418-451) + occlude_mask (agents.py:298-343) + MultiGrid.encode (base.py:196-214). */
int mg_obs_encode(const MgConfig* cfg, const MgState* st, uint8_t* obs, mg_stream_t stream);

/* RGB egocentric observation uint8 [B][A][V*ts][V*ts][3]:
 * gen_agent_obs (base.py:453-460) + MultiGrid.render/render_tile (base.py:275-331).
 * atlas: uint8 [n_tiles][4][ts][ts][3], n_tiles = (n_static_kinds+1)*(1+4A). */
int mg_obs_rgb(const MgConfig* cfg, const MgState* st, const uint8_t* atlas, uint8_t* obs, mg_stream_t stream);

/* step + autoreset + encoded obs in one call (the benchmarked hot path).  ONE launch (fused step+observe kernel)
 * for bit-plane worlds in ghost mode without respawn / spawn delay; otherwise two launches on `stream`: the
 * per-env step kernel, then the observe kernel. */
int mg_step_fused(const MgConfig* cfg, const MgState* st, const int32_t* actions, double* rewards,
                  uint8_t* done, uint8_t* obs, int autoreset, mg_stream_t stream);

/* step + autoreset + RGB obs in one call (step kernel, then reset+render kernel). */
int mg_step_fused_rgb(const MgConfig* cfg, const MgState* st, const int32_t* actions, double* rewards,
                      uint8_t* done, const uint8_t* atlas, uint8_t* obs, int autoreset, mg_stream_t stream);

/* n_steps fused steps back to back; actions int32 [n_steps][B][A]; rewards/done/obs hold the LAST
 * step's outputs.  Rollout driver for benchmarking (launch overhead amortised by the C loop). */
int mg_rollout_fused(const MgConfig* cfg, const MgState* st, const int32_t* actions, int64_t n_steps,
                     double* rewards, uint8_t* done, uint8_t* obs, int autoreset, mg_stream_t stream);

/* On-device rollout loop: n_steps env.steps on a fixed action tape with EVERY step's outputs kept -- actions int32
 * [n_steps][B][A], rewards float64 [n_steps][B][A], done uint8 [n_steps][B], obs uint8 [n_steps][B][A][V][V][3].  For the
 * registered shapes and batches whose tiles all fit the resident CTAs (65 536 envs of 3 agents on a B200) this is ONE launch:
 * every CTA keeps its tiles' state in shared memory between the steps; otherwise it launches step by step. */
int mg_rollout_persistent(const MgConfig* cfg, const MgState* st, const int32_t* actions, int64_t n_steps,
                          double* rewards, uint8_t* done, uint8_t* obs, int autoreset, mg_stream_t stream);

/* ON-DEVICE POLICY HAND-OFF (the caller loop of README.md:43-57 `act = agents.action_step(obs);
 * obs, rew, done, _ = env.step(act)` without leaving the GPU): a closed-loop rollout of n_steps steps in which step t + 1 plays the
 * actions a built-in policy chose from the observations of step t.  The policy is one int8 linear layer per agent over the
 * encoded observation (exact integer arithmetic): action = argmax_k (bias[a][k] + sum_i weights[a][k][i] * obs[i]), lowest k on
 * ties, k < n_actions; with probability epsilon / 2^32 a uniform action instead (Philox block keyed by `seed`, counter =
 * (global env index, lifetime step, agent)).  For the registered shapes and batches whose tiles fit the resident CTAs this is
 * ONE launch (the persistent rollout kernel evaluates the policy on the observation tile while it is still in shared memory);
 * otherwise one step launch + one policy launch per step.
 *   weights : device, int8, packed for the kernels as [A][NW][8][4]: byte b of word i of action k's row = w[a][k][4 i + b],
 *             NW = ceil(V*V*3 / 4), zero beyond the observation and for k >= n_actions (marlgrid_b200.policy packs it)
 *   bias    : device, int32 [A][8]
 *   actions : int32 [n_steps][B][A]: row 0 = the first step's actions, given by the caller; rows 1.. are WRITTEN (the actions
 *             played at every step, for the learner); rewards / done / obs: per-step slices as for mg_rollout_persistent
 * Batches that are not a multiple of 16 envs (per-step slices off the 16-byte grid) take the launch-per-step route.
 * The one-launch route evaluates the layer on the tensor cores (mma.sync m16n8k32, u8 x s8 -> s32: exact) from a per-process table
 * of weight fragments that a small kernel fills, stream-ordered, before the launch: rollouts with DIFFERENT policies must not be
 * in flight on two streams of one process at the same time (the same policy, or one stream, is fine). */
typedef struct MgLinearPolicy {
  const int8_t* weights;
  const int32_t* bias;
  int32_t n_actions;
  uint32_t epsilon;
  uint64_t seed;
} MgLinearPolicy;
int mg_rollout_policy(const MgConfig* cfg, const MgState* st, const MgLinearPolicy* policy, int64_t n_steps, int32_t* actions,
                      double* rewards, uint8_t* done, uint8_t* obs, int autoreset, mg_stream_t stream);

/* The policy alone, for a host loop: actions int32 [B][A] <- what the policy chooses from obs uint8 [B][A][V][V][3] (the encoded
 * observations of the step just played; exploration draws are keyed by the envs' lifetime step counters in `st`). */
int mg_policy_act(const MgConfig* cfg, const MgState* st, const MgLinearPolicy* policy, const uint8_t* obs, int32_t* actions, mg_stream_t stream);

/* The same driver over n_states independent env families of equal size, visited round robin: step t advances family
 * t % n_states with actions[t] and writes that family's rewards[r] / done[r] / obs[r].  With enough families the working
 * set exceeds the L2 cache, which is how bench.py times cold steps back to back (no flush kernel in between). */
int mg_rollout_fused_rr(const MgConfig* cfg, const MgState* states, int n_states, const int32_t* actions, int64_t n_steps,
                        double* const* rewards, uint8_t* const* done, uint8_t* const* obs, int autoreset, mg_stream_t stream);

/* Uniform random actions in {0..n_actions-1}, Philox stream keyed (seed, call counter):
 * the synthetic policy of the benchmark. */
int mg_random_actions(int32_t* actions, int64_t n, int n_actions, uint64_t seed, uint64_t counter,
                      mg_stream_t stream);

/* Line-of-sight known-answer entry: n independent VxV transparency grids (uint8, [n][V][V],
 * index [i][j] like the reference) -> visibility masks (same layout).  Replaces occlude_mask
 * (agents.py:298-343) for one agent position (ax, ay). */
int mg_los_batch(const uint8_t* transparent, uint8_t* mask, int64_t n, int view_size, int ax, int ay,
                 mg_stream_t stream);

/* ---- host-buffer engine API (owns device memory; HOST pointers; synchronous) ---------------- */
typedef struct MgEngine MgEngine;

/* Allocates device state for n_envs on CUDA device `device`.  rgb != 0 selects the RGB path
 * (atlas_host: host pointer to the atlas, copied once). */
int mg_engine_create(MgEngine** out, const MgConfig* cfg, int64_t n_envs, int64_t env_offset, uint64_t seed,
                     int device, int rgb, const uint8_t* atlas_host, int64_t atlas_bytes);
void mg_engine_destroy(MgEngine* e);
/* env.reset(): obs_host (pageable or pinned) receives the first observations. */
int mg_engine_reset(MgEngine* e, uint8_t* obs_host);
/* env.step(actions): H2D actions, fused step kernel, D2H obs/rewards/done; returns after the copies. */
int mg_engine_step(MgEngine* e, const int32_t* actions_host, uint8_t* obs_host, double* rewards_host,
                   uint8_t* done_host, int autoreset);
/* The host<->device transfers of mg_engine_step WITHOUT the kernel (same buffers, slices and streams): measures the copy
 * ceiling of the box that the end-to-end figure is bound by (bench.py e2e.copy_ceiling). */
int mg_engine_copy_only(MgEngine* e, const int32_t* actions_host, uint8_t* obs_host, double* rewards_host, uint8_t* done_host);
/* Pinned host allocation helpers so callers without a CUDA runtime can get page-locked buffers. */
void* mg_host_alloc(int64_t bytes);
void mg_host_free(void* p);
/* Counters: kernels launched by this library since load (bench.py's gpu_launches). */
int64_t mg_launch_count(void);
/* Profiling hook (bench.py): a cudaEvent_t recorded between the step kernel and the reset+observe kernel
 * of every following mg_step* call, so each kernel's duration can be read live; NULL disables it. */
void mg_debug_set_mid_event(void* cuda_event);
/* Test hook: != 0 makes mg_step_fused* use the two-launch path (per-env step kernel, then observe kernel) even
 * where the single fused kernel applies, so both implementations stay covered. */
void mg_debug_force_two_kernels(int on);
/* Test hook: != 0 makes mg_step_fused* use the general fused kernel (run-time agent count / view size) even where a
 * specialised instantiation (compile-time A and V: the registered env shapes) exists. */
void mg_debug_force_general_fused(int on);

#ifdef __cplusplus
}
#endif
#endif /* MARLGRID_B200_H */

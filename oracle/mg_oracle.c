/*
 * mg_oracle.c -- CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file restates, in plain sequential C, the algorithm of kandouss/marlgrid's
 * MultiGridEnv.reset / step / gen_obs_grid / MultiGrid.encode / MultiGrid.render on the same
 * structure-of-arrays layout the device uses (include/marlgrid_b200.h), so outputs can be compared
 * byte for byte.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load it; the product path never does.
 *
 * Parity status: PINNED.  oracle/validate_against_reference.py runs the unmodified reference
 * (under the import shims of oracle/shims) in lock step with this file on random and scripted
 * trajectories -- obs (encoded + RGB), float64 rewards, done and full state -- and
 * oracle/gen_golden.py freezes reference outputs as tests/golden/ (.npz files), which
 * tests/test_oracle_golden.py replays against this file where /root/reference is absent.
 * The reference itself has no tests or golden vectors.
 *
 * Each function cites the reference lines it follows (paths relative to /root/reference).
 * Python object lists (`obj.agents`, base.py:547-572) are represented by arrival stamps: the
 * agents standing on one cell, sorted by stamp, ARE the reference's queue (cell object first,
 * then its `.agents` in append order) -- see DESIGN.md "stacking".
 */
#include <stdint.h>
#include <stdlib.h>
#include <math.h>
#include <string.h>

#include "../include/marlgrid_b200.h"

#include <pthread.h>
#include <unistd.h>

/* ------------------------------------------------------------------------------------------ */
/* Philox4x32-10 (contract: oracle/philox.py)                                                  */
/* ------------------------------------------------------------------------------------------ */
static void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
static uint32_t mulhi32(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }

#define TAG_RESET 0x80000000u
#define TAG_INSTEP 0x40000000u

typedef struct {
  uint64_t seed, g;
  uint32_t c2, tag, k; /* k = placement tries drawn so far in this reset()/step() */
} Draws;

static void draw_pos(Draws* d, int W, int H, int* x, int* y) { /* base.py:699 np_random.randint(top, bottom) */
  uint32_t ctr[4] = {(uint32_t)d->g, (uint32_t)(d->g >> 32), d->c2, d->tag | (d->k >> 1)};
  uint32_t key[2] = {(uint32_t)d->seed, (uint32_t)(d->seed >> 32)}, r[4];
  philox4x32_10(ctr, key, r);
  int o = 2 * (int)(d->k & 1u);
  *x = (int)mulhi32(r[o], (uint32_t)W);
  *y = (int)mulhi32(r[o + 1], (uint32_t)H);
  d->k++;
}

static void draw_order(uint64_t seed, uint64_t g, uint32_t t, int A, int* order) { /* base.py:514-516 */
  uint32_t ctr[4] = {(uint32_t)g, (uint32_t)(g >> 32), t, 0u};
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)}, r[4];
  philox4x32_10(ctr, key, r);
  uint32_t fact = 1;
  for (int i = 2; i <= A; ++i) fact *= (uint32_t)i;
  uint32_t idx = mulhi32(r[0], fact);
  for (int i = 0; i < A; ++i) order[i] = i;
  for (int i = A - 1; i >= 1; --i) {
    int j = (int)(idx % (uint32_t)(i + 1));
    idx /= (uint32_t)(i + 1);
    int tmp = order[i]; order[i] = order[j]; order[j] = tmp;
  }
}

/* ------------------------------------------------------------------------------------------ */
/* one env's view of the SoA buffers                                                           */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
  const MgConfig* c;
  uint8_t* type; uint8_t* colour; uint8_t* state; /* planes, cell (x,y) at x*H+y (base.py:91) */
  uint8_t* ag;   /* [A][16] */
  int32_t* er;   /* [4] */
  double* pr;    /* [A] GridAgentInterface.prestige (agents.py:141-153) or NULL */
} Env;

/* the batch's prestige values [B][A], set by mgo_set_prestige (NULL: not tracked) */
static double* g_prestige = 0;
void mgo_set_prestige(double* p) { g_prestige = p; }

static Env env_at(const MgConfig* c, uint8_t* grid, uint8_t* agents, int32_t* envrec, int64_t e) {
  Env v; v.c = c;
  v.type = grid + (size_t)e * 3 * c->plane_stride;
  v.colour = v.type + c->plane_stride;
  v.state = v.colour + c->plane_stride;
  v.ag = agents + (size_t)e * c->n_agents * MG_AGENT_REC;
  v.er = envrec + (size_t)e * 4;
  v.pr = g_prestige ? g_prestige + (size_t)e * c->n_agents : 0;
  return v;
}
#define AX(v, a) ((v)->ag[(a) * 16 + 0])
#define AY(v, a) ((v)->ag[(a) * 16 + 1])
#define ADIR(v, a) ((v)->ag[(a) * 16 + 2])
#define AFL(v, a) ((v)->ag[(a) * 16 + 3])
#define ACT(v, a) ((v)->ag[(a) * 16 + 4])
#define ACC(v, a) ((v)->ag[(a) * 16 + 5])
#define ACS(v, a) ((v)->ag[(a) * 16 + 6])
#define ABONUS(v, a) ((v)->ag[(a) * 16 + 7])
static int32_t get_stamp(const Env* v, int a) { int32_t s; memcpy(&s, v->ag + a * 16 + 8, 4); return s; }
static void set_stamp(Env* v, int a, int32_t s) { memcpy(v->ag + a * 16 + 8, &s, 4); }
static void add_err(Env* v, uint32_t bits) { v->er[3] = (int32_t)((uint32_t)v->er[3] | (bits << 16)); }
static int32_t next_stamp(Env* v) {
  uint32_t w = (uint32_t)v->er[3];
  uint32_t s = w & 0xFFFFu;
  v->er[3] = (int32_t)((w & 0xFFFF0000u) | ((s + 1) & 0xFFFFu));
  return (int32_t)s;
}

/* objects.py predicates */
static int can_overlap_static(int type, int state) { /* objects.py:75-76,147-148,174,216,230,258,327-328 */
  switch (type) {
    case MG_T_BONUS: case MG_T_GOAL: case MG_T_FLOOR: case MG_T_LAVA: return 1;
    case MG_T_DOOR: return state == MG_DOOR_OPEN;
    default: return 0; /* incl. EmptySpace: `can_verlap` typo objects.py:250 */
  }
}
static int can_pickup(int type) { return type == MG_T_KEY || type == MG_T_BALL || type == MG_T_BOX; } /* objects.py:292,314,378 */
static int see_behind(int type, int state) { /* objects.py:84-85,281-282,330-331 */
  if (type == MG_T_WALL) return 0;
  if (type == MG_T_DOOR) return state == MG_DOOR_OPEN;
  return 1;
}

/* queue head = the placed agent with the smallest stamp on (x,y); -1 if none.
 * (reference: cell object if it is an agent, else static_obj.agents[0]) */
static int queue_head(const Env* v, int x, int y) {
  int best = -1; int32_t bs = 0;
  for (int a = 0; a < v->c->n_agents; ++a)
    if ((AFL(v, a) & MG_AF_PLACED) && AX(v, a) == x && AY(v, a) == y) {
      int32_t s = get_stamp(v, a);
      if (best < 0 || s < bs) { best = a; bs = s; }
    }
  return best;
}

/* the agent right behind `head` in the queue of (x,y): the reference's head.agents[0] / static_obj.agents[1]; -1 if none */
static int queue_second(const Env* v, int x, int y, int head) {
  int best = -1; int32_t bs = 0;
  for (int a = 0; a < v->c->n_agents; ++a)
    if (a != head && (AFL(v, a) & MG_AF_PLACED) && AX(v, a) == x && AY(v, a) == y) {
      int32_t s = get_stamp(v, a);
      if (best < 0 || s < bs) { best = a; bs = s; }
    }
  return best;
}

/* base.py:664-688 try_place_obj.  agent >= 0: placing that agent; else placing a static triple. */
static int try_place(Env* v, int x, int y, int agent, int type, int colour, int state) {
  const MgConfig* c = v->c;
  int idx = x * c->height + y;
  int st = v->type[idx];
  int head = queue_head(v, x, y);
  if (st == MG_T_EMPTY && head < 0) { /* grid_obj is None, base.py:672-675 */
    if (agent >= 0) {
      AX(v, agent) = (uint8_t)x; AY(v, agent) = (uint8_t)y; AFL(v, agent) |= MG_AF_PLACED;
      set_stamp(v, agent, next_stamp(v));
    } else {
      v->type[idx] = (uint8_t)type; v->colour[idx] = (uint8_t)colour; v->state[idx] = (uint8_t)state;
    }
    return 1;
  }
  int overlap = (st != MG_T_EMPTY) ? can_overlap_static(st, v->state[idx]) : 1 /* base is an agent */;
  if (!(overlap && agent >= 0)) return 0; /* base.py:678-679 */
  if (!(c->flags & MG_F_GHOST) && head >= 0) return 0; /* base.py:683-684 */
  AX(v, agent) = (uint8_t)x; AY(v, agent) = (uint8_t)y; AFL(v, agent) |= MG_AF_PLACED; /* base.py:686-687 */
  set_stamp(v, agent, next_stamp(v));
  return 1;
}

/* base.py:690-708 place_obj(obj, top, size, max_tries): pos = np_random.randint(top, bottom) in the box
 * top = max(top, 0), bottom = min(top + size, grid); size NULL = the whole grid */
static void place_obj_box(Env* v, Draws* d, int agent, int type, int colour, int state, int max_tries, int tx, int ty, int bx, int by) {
  if (max_tries > 100000) max_tries = 100000;
  if (max_tries < 1) max_tries = 1;
  for (int t = 0; t < max_tries; ++t) {
    int x, y;
    draw_pos(d, bx - tx, by - ty, &x, &y);
    if (try_place(v, x + tx, y + ty, agent, type, colour, state)) return;
  }
  add_err(v, MG_ERR_PLACEMENT); /* RecursionError base.py:706 */
}
static void place_obj(Env* v, Draws* d, int agent, int type, int colour, int state, int max_tries) {
  place_obj_box(v, d, agent, type, colour, state, max_tries, 0, 0, v->c->width, v->c->height);
}
/* place_obj(agent, **self.agent_spawn_kwargs): base.py:409-412 (reset), :505 (spawn delay), :642 (respawn) */
static void place_agent(Env* v, Draws* d, int agent) {
  const MgConfig* c = v->c;
  int tx = c->spawn_top[0] > 0 ? c->spawn_top[0] : 0, ty = c->spawn_top[1] > 0 ? c->spawn_top[1] : 0;
  int whole = c->spawn_size[0] == 0 && c->spawn_size[1] == 0;
  int bx = tx + (whole ? c->width : c->spawn_size[0]), by = ty + (whole ? c->height : c->spawn_size[1]);
  if (bx > c->width) bx = c->width;
  if (by > c->height) by = c->height;
  place_obj_box(v, d, agent, 0, 0, 0, c->spawn_max_tries > 0 ? c->spawn_max_tries : 100000, tx, ty, bx, by);
}
/* a scalar np_random.randint(lo, hi): one try slot of the stream, its first word */
static int draw_int(Draws* d, int lo, int hi) {
  int x, y;
  draw_pos(d, hi - lo, 1, &x, &y);
  return lo + x;
}

/* DoorKeyEnv._gen_grid, doorkey.py:15-41 (with `_rand_int(lo, hi)` = np_random.randint(lo, hi)) */
static void gen_doorkey(Env* v, Draws* d) {
  const MgConfig* c = v->c;
  int W = c->width, H = c->height;
  int idx = (W - 2) * H + (H - 2);
  v->type[idx] = MG_T_GOAL; v->colour[idx] = MG_C_GREEN; v->state[idx] = 0;             /* doorkey.py:23 */
  int split = draw_int(d, 2, W - 2);                                                      /* doorkey.py:26 */
  for (int j = 0; j < H; ++j) {                                                           /* vert_wall base.py:166-170 */
    v->type[split * H + j] = MG_T_WALL; v->colour[split * H + j] = MG_C_WORST; v->state[split * H + j] = 0;
  }
  int door = draw_int(d, 1, W - 2);                                                       /* doorkey.py:34 */
  if (door < H) {
    v->type[split * H + door] = MG_T_DOOR; v->colour[split * H + door] = MG_C_YELLOW; v->state[split * H + door] = MG_DOOR_LOCKED;
  } else add_err(v, MG_ERR_STACK);                                                        /* grid.set asserts j < height */
  place_obj_box(v, d, -1, MG_T_KEY, MG_C_YELLOW, 0, 100000, 0, 0, split < W ? split : W, H); /* doorkey.py:37 */
}

/* base.py:402-416 reset + empty.py:9-16 / cluttered.py:25-36 / goalcycle.py:30-51 _gen_grid */
static void env_reset(Env* v, uint64_t seed, uint64_t g) {
  const MgConfig* c = v->c;
  int W = c->width, H = c->height, A = c->n_agents;
  for (int a = 0; a < A; ++a) { /* agents.py:161-170: dir/state untouched */
    AFL(v, a) = 0; AX(v, a) = 0; AY(v, a) = 0; ACT(v, a) = 0; ACC(v, a) = 0; ACS(v, a) = 0; ABONUS(v, a) = 0xFF;
    set_stamp(v, a, 0);
    if (v->pr) v->pr[a] = 0.0; /* agent.reset(new_episode=True) agents.py:167-168 */
  }
  memset(v->type, 0, (size_t)c->plane_stride * 3);
  v->er[3] = (int32_t)((uint32_t)v->er[3] & 0xFFFF0000u); /* next stamp = 0, keep error bits */
  for (int i = 0; i < W; ++i) for (int j = 0; j < H; ++j) /* wall_rect base.py:172-176 */
    if (i == 0 || j == 0 || i == W - 1 || j == H - 1) {
      v->type[i * H + j] = MG_T_WALL; v->colour[i * H + j] = MG_C_WORST; v->state[i * H + j] = 0;
    }
  Draws d = {seed, g, (uint32_t)v->er[1], TAG_RESET, 0};
  if (c->scenario == MG_SCENARIO_DOORKEY) {
    gen_doorkey(v, &d);
    for (int a = 0; a < A; ++a) /* base.py:409-412; the generator has reset agent_spawn_kwargs to {} (doorkey.py:40) */
      if (c->spawn_delay[a] == 0) {
        place_obj(v, &d, a, 0, 0, 0, 100000);
        AFL(v, a) |= MG_AF_ACTIVE;
      }
    v->er[0] = 0;
    v->er[1] += 1;
    return;
  }
  if (c->goal_mode == MG_GOAL_FIXED) { /* put_obj base.py:655-662 replaces whatever is there */
    int idx = (W - 2) * H + (H - 2);
    v->type[idx] = MG_T_GOAL; v->colour[idx] = MG_C_GREEN; v->state[idx] = 0;
  } else if (c->goal_mode == MG_GOAL_RANDOM) {
    place_obj(v, &d, -1, MG_T_GOAL, MG_C_GREEN, 0, 100); /* cluttered.py:28-29 */
  }
  for (int b = 0; b < c->n_bonus_tiles; ++b) place_obj(v, &d, -1, MG_T_BONUS, MG_C_YELLOW, b, 100); /* goalcycle.py:34-46 */
  for (int k = 0; k < c->n_clutter; ++k) place_obj(v, &d, -1, MG_T_WALL, MG_C_WORST, 0, 100);        /* cluttered.py:32-33 */
  for (int a = 0; a < A; ++a) /* base.py:409-412 */
    if (c->spawn_delay[a] == 0) {
      place_agent(v, &d, a);
      AFL(v, a) |= MG_AF_ACTIVE;
    }
  v->er[0] = 0; /* step_count base.py:414 */
  v->er[1] += 1;
}

/* BonusTile.get_reward objects.py:180-206 */
static double bonus_get_reward(Env* v, int a, int bonus_id) {
  const MgConfig* c = v->c;
  int n = c->n_bonus_tiles, first = 0;
  double pen = c->bonus_penalty < 0 ? c->bonus_penalty : -c->bonus_penalty; /* -abs(penalty) */
  double rew;
  if (ABONUS(v, a) == 0xFF) { ABONUS(v, a) = (uint8_t)(((bonus_id - 1) % n + n) % n); first = 1; }
  if (ABONUS(v, a) == bonus_id) rew = pen;
  else if ((ABONUS(v, a) + 1) % n == bonus_id) { ABONUS(v, a) = (uint8_t)bonus_id; rew = c->bonus_reward; }
  else rew = pen;
  if (c->flags & MG_F_BONUS_RESET) ABONUS(v, a) = (uint8_t)bonus_id;
  if (first && !(c->flags & MG_F_BONUS_INITIAL)) return 0.0;
  return rew;
}

/* base.py:501-649 step (no obs).  returns done. */
static int env_step(Env* v, const int32_t* actions, double* rewards, uint64_t seed, uint64_t g) {
  const MgConfig* c = v->c;
  int W = c->width, H = c->height, A = c->n_agents;
  static const int DX[4] = {1, 0, -1, 0}, DY[4] = {0, 1, 0, -1}; /* agents.py:183 */
  uint32_t t_life = (uint32_t)v->er[2];
  Draws d = {seed, g, t_life, TAG_INSTEP, 0};
  for (int a = 0; a < A; ++a) /* base.py:503-506 */
    if (!(AFL(v, a) & MG_AF_ACTIVE) && !(AFL(v, a) & MG_AF_DONE) && v->er[0] >= c->spawn_delay[a]) {
      place_agent(v, &d, a);
      AFL(v, a) |= MG_AF_ACTIVE;
    }
  for (int a = 0; a < A; ++a) rewards[a] = 0.0; /* base.py:510 */
  v->er[0] += 1;                                 /* base.py:512 */
  int order[MG_MAX_AGENTS];
  draw_order(seed, g, t_life, A, order);
  v->er[2] += 1;
  for (int p = 0; p < A; ++p) {
    int a = order[p];
    if (!(AFL(v, a) & MG_AF_ACTIVE)) continue; /* base.py:521 */
    int act = actions[a];
    int cx = AX(v, a), cy = AY(v, a), dir = ADIR(v, a) & 3;
    int fx = cx + DX[dir], fy = cy + DY[dir];
    int inb = fx >= 0 && fy >= 0 && fx < W && fy < H;
    int fidx = inb ? fx * H + fy : 0;
    int ftype = inb ? v->type[fidx] : MG_T_WALL; /* grid.get asserts in-bounds (base.py:154-156); never hit with wall_rect */
    int fstate = inb ? v->state[fidx] : 0;
    int fhead = inb ? queue_head(v, fx, fy) : -1;
    int f_none = (ftype == MG_T_EMPTY && fhead < 0);      /* fwd_cell is None */
    int f_is_agent = (ftype == MG_T_EMPTY && fhead >= 0); /* fwd_cell is a GridAgent */
    if (!inb) add_err(v, MG_ERR_STACK);
    if (act == MG_A_LEFT) ADIR(v, a) = (uint8_t)((dir + 3) & 3);       /* base.py:530-531 */
    else if (act == MG_A_RIGHT) ADIR(v, a) = (uint8_t)((dir + 1) & 3); /* base.py:534-535 */
    else if (act == MG_A_FORWARD) {                                    /* base.py:538-585 */
      int can_move = f_none || f_is_agent || can_overlap_static(ftype, fstate);
      if (!(c->flags & MG_F_GHOST) && f_is_agent) can_move = 0;
      if (can_move) {
        int cidx = cx * H + cy;
        /* leaving a non-overlappable static cell trips `assert cur_cell.can_overlap()` base.py:558 */
        if (v->type[cidx] != MG_T_EMPTY && !can_overlap_static(v->type[cidx], v->state[cidx])) add_err(v, MG_ERR_STACK);
        AX(v, a) = (uint8_t)fx; AY(v, a) = (uint8_t)fy;
        set_stamp(v, a, next_stamp(v)); /* appended last to the target cell's queue base.py:547-552 */
        if (ftype == MG_T_GOAL || ftype == MG_T_BONUS) { /* hasattr(fwd_cell,'get_reward') base.py:576 */
          double rwd = (ftype == MG_T_GOAL) ? c->goal_reward : bonus_get_reward(v, a, fstate);
          if (c->flags & MG_F_REWARD_DECAY) {
            volatile double q = (double)v->er[0] / (double)c->max_steps; /* base.py:579, no contraction */
            volatile double u = 0.9 * q;
            volatile double f = 1.0 - u;
            rwd = rwd * f;
          }
          rewards[a] += rwd;
          if (v->pr) { /* agent.reward(rwd) base.py:581, agents.py:146-153 */
            if ((c->prestige_neg_mask >> a) & 1u) add_err(v, MG_ERR_PRESTIGE); /* `self.rew += rew`: AttributeError */
            else if (rwd >= 0) { volatile double s = v->pr[a] + rwd; v->pr[a] = s; }
            else v->pr[a] = 0.0;
          }
        }
        if (ftype == MG_T_LAVA || ftype == MG_T_GOAL) AFL(v, a) |= MG_AF_DONE; /* base.py:584-585 */
      }
    } else if (act == MG_A_PICKUP) { /* base.py:590-597 */
      if (ftype != MG_T_EMPTY && can_pickup(ftype) && ACT(v, a) == 0) {
        ACT(v, a) = (uint8_t)ftype; ACC(v, a) = v->colour[fidx]; ACS(v, a) = (uint8_t)fstate;
        v->type[fidx] = 0; v->colour[fidx] = 0; v->state[fidx] = 0;
      }
    } else if (act == MG_A_DROP) { /* base.py:600-606 */
      if (f_none && inb && ACT(v, a) != 0) {
        v->type[fidx] = ACT(v, a); v->colour[fidx] = ACC(v, a); v->state[fidx] = ACS(v, a);
        ACT(v, a) = 0; ACC(v, a) = 0; ACS(v, a) = 0;
      }
    } else if (act == MG_A_TOGGLE) { /* base.py:609-613, Door.toggle objects.py:333-346 */
      if (ftype == MG_T_DOOR) {
        if (fstate == MG_DOOR_LOCKED) {
          if (ACT(v, a) == MG_T_KEY && ACC(v, a) == v->colour[fidx]) v->state[fidx] = MG_DOOR_CLOSED;
        } else if (fstate == MG_DOOR_CLOSED) v->state[fidx] = MG_DOOR_OPEN;
        else if (fstate == MG_DOOR_OPEN) v->state[fidx] = MG_DOOR_CLOSED;
      } else if (ftype == MG_T_BOX) add_err(v, MG_ERR_TOGGLE); /* Box.toggle(self) objects.py:381 */
    } else if (act == MG_A_DONE) { /* base.py:616-617 */
    } else { add_err(v, MG_ERR_BAD_ACTION); continue; } /* base.py:619-620: raises before on_step */
    if (v->pr) { volatile double s = v->pr[a] * c->prestige_beta[a]; v->pr[a] = s; } /* agent.on_step base.py:622, agents.py:141-144 */
  }
  for (int a = 0; a < A; ++a) /* base.py:627-646 */
    if (AFL(v, a) & MG_AF_DONE) {
      if (c->flags & MG_F_RESPAWN) {
        AFL(v, a) = 0; ACT(v, a) = 0; ACC(v, a) = 0; ACS(v, a) = 0; /* agent.reset(new_episode=False) agents.py:161-166 */
        place_agent(v, &d, a);
        AFL(v, a) |= MG_AF_ACTIVE;
      } else AFL(v, a) &= (uint8_t)~MG_AF_ACTIVE;
    }
  int all_done = 1;
  for (int a = 0; a < A; ++a) if (!(AFL(v, a) & MG_AF_DONE)) all_done = 0;
  return (v->er[0] >= c->max_steps) || all_done; /* base.py:649 */
}

/* ------------------------------------------------------------------------------------------ */
/* observation                                                                                 */
/* ------------------------------------------------------------------------------------------ */
/* agents.py:298-343 occlude_mask, canonical OOB semantics: row j == V is a no-op. */
static void occlude_mask(const uint8_t* grid /* [V][V] transparent, [i][j] */, int V, int ax, int ay, uint8_t* mask) {
  memset(mask, 0, (size_t)V * V);
  mask[ax * V + ay] = 1;
  for (int j = ay + 1; j > 0; --j) {
    if (j >= V) continue; /* out-of-bounds row reads as zeros */
    for (int i = ax; i < V; ++i)
      if (mask[i * V + j] && grid[i * V + j]) {
        if (i < V - 1) mask[(i + 1) * V + j] = 1;
        if (j > 0) { mask[i * V + j - 1] = 1; if (i < V - 1) mask[(i + 1) * V + j - 1] = 1; }
      }
    for (int i = ax + 1; i > 0; --i) {
      if (i >= V) continue;
      if (mask[i * V + j] && grid[i * V + j]) {
        if (i > 0) mask[(i - 1) * V + j] = 1;
        if (j > 0) { mask[i * V + j - 1] = 1; if (i > 0) mask[(i - 1) * V + j - 1] = 1; }
      }
    }
  }
  for (int j = ay; j < V; ++j) {
    for (int i = ax; i < V; ++i)
      if (mask[i * V + j] && grid[i * V + j]) {
        if (i < V - 1) mask[(i + 1) * V + j] = 1;
        if (j < V - 1) { mask[i * V + j + 1] = 1; if (i < V - 1) mask[(i + 1) * V + j + 1] = 1; }
      }
    for (int i = ax + 1; i > 0; --i) {
      if (i >= V) continue;
      if (mask[i * V + j] && grid[i * V + j]) {
        if (i > 0) mask[(i - 1) * V + j] = 1;
        if (j < V - 1) { mask[i * V + j + 1] = 1; if (i > 0) mask[(i - 1) * V + j + 1] = 1; }
      }
    }
  }
}

typedef struct { uint8_t type, colour, state; int8_t head; /* queue head agent or -1 */ uint8_t has_obs;
                 int8_t second; /* agent behind the head or -1 */ uint8_t replaced; /* hide_item_types put agents[0] here */ } ViewCell;

/* base.py:418-451 gen_obs_grid: slice (base.py:123-147) + rotate_grid (base.py:67-80) + opacity (103-106)
 * + process_vis (agents.py:290-295).  Fills cells[V*V] ([a][b]) and vis[V*V]; returns 0 if inactive. */
static int gen_obs_grid(const Env* v, int a, ViewCell* cells, uint8_t* vis) {
  const MgConfig* c = v->c;
  int V = c->view_size, W = c->width, H = c->height, o = c->view_offset, h = V / 2;
  if (!(AFL(v, a) & MG_AF_ACTIVE)) return 0; /* base.py:420-425 */
  int px = AX(v, a), py = AY(v, a), dir = ADIR(v, a) & 3, topX, topY;
  if (dir == 0) { topX = px - o; topY = py - h; }              /* agents.py:245-247 */
  else if (dir == 1) { topX = px - h; topY = py - o; }         /* agents.py:249-251 */
  else if (dir == 2) { topX = px - V + 1 + o; topY = py - h; } /* agents.py:253-255 */
  else { topX = px - h; topY = py - V + 1 + o; }               /* agents.py:257-259 */
  ViewCell sub[MG_MAX_VIEW * MG_MAX_VIEW];
  for (int sx = 0; sx < V; ++sx) for (int sy = 0; sy < V; ++sy) { /* slice: zero padded */
    ViewCell cell = {0, 0, 0, -1, 0, -1, 0};
    int wx = topX + sx, wy = topY + sy;
    if (wx >= 0 && wy >= 0 && wx < W && wy < H) {
      int idx = wx * H + wy;
      cell.type = v->type[idx]; cell.colour = v->colour[idx]; cell.state = v->state[idx];
      cell.head = (int8_t)queue_head(v, wx, wy);
      cell.second = (int8_t)(cell.head >= 0 ? queue_second(v, wx, wy, cell.head) : -1);
      cell.has_obs = (uint8_t)(AX(v, a) == wx && AY(v, a) == wy); /* the observer itself stands on this cell */
    }
    sub[sx * V + sy] = cell;
  }
  int k = (dir + 1) & 3; /* rot_k base.py:429-431 */
  uint8_t transp[MG_MAX_VIEW * MG_MAX_VIEW];
  for (int va = 0; va < V; ++va) for (int vb = 0; vb < V; ++vb) {
    int sx, sy;
    if (k == 0) { sx = va; sy = vb; }
    else if (k == 1) { sx = V - 1 - vb; sy = va; }         /* moveaxis(grid[::-1,:],0,1) base.py:75-76 */
    else if (k == 2) { sx = V - 1 - va; sy = V - 1 - vb; } /* base.py:77-78 */
    else { sx = vb; sy = V - 1 - va; }                     /* moveaxis(grid[:,::-1],0,1) base.py:73-74 */
    cells[va * V + vb] = sub[sx * V + sy];
    const ViewCell* cc = &cells[va * V + vb];
    transp[va * V + vb] = (uint8_t)(cc->type == MG_T_EMPTY ? 1 : see_behind(cc->type, cc->state)); /* agents transparent */
  }
  if (c->flags & MG_F_SEE_THROUGH) memset(vis, 1, (size_t)V * V); /* agents.py:294-295 */
  else occlude_mask(transp, V, V / 2, V - 1 - o, vis);           /* agents.py:233-234,293 */
  if (c->hide_types) { /* base.py:441-449: objects of hidden types are replaced by item.agents[0] (or nothing) */
    for (int i = 0; i < V * V; ++i) {
      ViewCell* cc = &cells[i];
      if (cc->type != MG_T_EMPTY) { /* item = the static object; item.agents = the whole queue */
        if ((c->hide_types >> cc->type) & 1u) {
          cc->type = MG_T_EMPTY; cc->colour = 0; cc->state = 0; cc->replaced = 1; /* head (if any) becomes the cell object */
          cc->second = -1;
        }
      } else if (cc->head >= 0 && cc->head != a && ((c->hide_types >> MG_T_AGENT) & 1u)) { /* item = the head agent, not the observer */
        cc->head = cc->second; cc->second = -1; cc->replaced = 1;
      }
    }
  }
  return 1;
}

/* base.py:196-214 MultiGrid.encode + objects.py:90-99 */
static void obs_encode_env(const Env* v, uint8_t* obs) {
  const MgConfig* c = v->c;
  int V = c->view_size;
  ViewCell cells[MG_MAX_VIEW * MG_MAX_VIEW]; uint8_t vis[MG_MAX_VIEW * MG_MAX_VIEW];
  for (int a = 0; a < c->n_agents; ++a) {
    uint8_t* out = obs + (size_t)a * V * V * 3;
    memset(out, 0, (size_t)V * V * 3);
    if (!gen_obs_grid(v, a, cells, vis)) continue;
    for (int i = 0; i < V * V; ++i) {
      if (!vis[i]) continue;
      const ViewCell* cc = &cells[i];
      if (cc->type != MG_T_EMPTY) { out[i * 3] = cc->type; out[i * 3 + 1] = cc->colour; out[i * 3 + 2] = cc->state; }
      else if (cc->head >= 0) { /* the cell object is the head agent: (13, colour, state==dir) */
        out[i * 3] = MG_T_AGENT; out[i * 3 + 1] = c->agent_color[cc->head]; out[i * 3 + 2] = ADIR(v, cc->head);
      }
    }
  }
}

/* atlas tile index: kind*(1+4A) + (no agent ? 0 : 1 + 4*q + dir_q); see DESIGN.md */
static int tile_index(const MgConfig* c, int kind, int q, int qdir) {
  return kind * (1 + 4 * c->n_agents) + (q < 0 ? 0 : 1 + 4 * q + qdir);
}

/* base.py:453-460 gen_agent_obs + base.py:301-331 render + base.py:275-299 render_tile.
 * atlas[tile][orientation] holds rotate_grid(tile, orientation) (base.py:324) precomputed on the host. */
static void obs_rgb_env(Env* v, const uint8_t* atlas, uint8_t* obs) {
  const MgConfig* c = v->c;
  int V = c->view_size, ts = c->view_tile_size, row = V * ts * 3;
  size_t tile_bytes = (size_t)ts * ts * 3;
  ViewCell cells[MG_MAX_VIEW * MG_MAX_VIEW]; uint8_t vis[MG_MAX_VIEW * MG_MAX_VIEW];
  static const uint8_t shadow[3] = {35, 25, 30}; /* COLORS['shadow'] objects.py:25, base.py:305 */
  for (int a = 0; a < c->n_agents; ++a) {
    uint8_t* img = obs + (size_t)a * V * ts * row;
    for (int p = 0; p < V * ts * V * ts; ++p) memcpy(img + p * 3, shadow, 3);
    if (!gen_obs_grid(v, a, cells, vis)) continue;
    int orient = (3 - (ADIR(v, a) & 3)) & 3; /* (0 - rot_k) % 4, base.py:130 */
    for (int vb = 0; vb < V; ++vb) for (int va = 0; va < V; ++va) { /* base.py:307-324 */
      if (!vis[va * V + vb]) continue;
      const ViewCell* cc = &cells[va * V + vb];
      int kind = 0;
      if (cc->type != MG_T_EMPTY) {
        kind = c->kind_of_type[cc->type];
        if (kind == 0xFF) { add_err(v, MG_ERR_RENDER); kind = 0; }
      }
      int q = -1;
      if (cc->head >= 0) q = (cc->has_obs && !cc->replaced) ? a : cc->head; /* base.py:282-293: top_agent if it is on this cell, else queue
                                                                               head; a hide_item_types replacement has no agents of its own */
      int t = tile_index(c, kind, q, q >= 0 ? (ADIR(v, q) & 3) : 0);
      const uint8_t* tile = atlas + ((size_t)t * 4 + orient) * tile_bytes;
      for (int y = 0; y < ts; ++y) memcpy(img + (size_t)(vb * ts + y) * row + (size_t)va * ts * 3, tile + (size_t)y * ts * 3, (size_t)ts * 3);
      if (q >= 0 && v->pr && ((c->prestige_mask >> q) & 1u) && (AFL(v, q) & MG_AF_ACTIVE)) {
        /* GridAgentInterface.render_post (agents.py:92-119) on the cached WHITE agent tile, then render_tile's blend over the
         * cell's object (base.py:260-273,289-293) and border rule (:296-298): the atlas tile of (no object, agent q) minus the
         * empty tile's border is the triangle's alpha; an inactive agent keeps the cached tile (agents.py:93-94) */
        double x = v->pr[q] / c->prestige_scale[q];
        double sc = ((c->prestige_neg_mask >> q) & 1u) ? 1.0 / (1.0 + exp(-x)) : tanh(x);
        int col[3] = {(int)(sc * 0.0 + (1.0 - sc) * 255.0), 0, (int)(sc * 255.0 + (1.0 - sc) * 0.0)};
        const uint8_t* white = atlas + ((size_t)tile_index(c, 0, q, ADIR(v, q) & 3) * 4 + orient) * tile_bytes;
        const uint8_t* empty = atlas + (size_t)orient * tile_bytes;
        const uint8_t* base = atlas + ((size_t)tile_index(c, kind, -1, 0) * 4 + orient) * tile_bytes;
        int amax = 0;
        for (int p = 0; p < ts * ts; ++p) { int al = white[p * 3] - empty[p * 3]; if (al > amax) amax = al; }
        int M = ((amax * col[0]) >> 8) + ((amax * col[2]) >> 8);
        for (int y = 0; y < ts; ++y) for (int xx = 0; xx < ts; ++xx) {
          int p = y * ts + xx, al = white[p * 3] - empty[p * 3];
          int ag[3] = {(al * col[0]) >> 8, 0, (al * col[2]) >> 8}, sa = ag[0] + ag[2];
          uint8_t* o = img + (size_t)(vb * ts + y) * row + (size_t)(va * ts + xx) * 3;
          for (int ch = 0; ch < 3; ++ch)
            o[ch] = (uint8_t)(kind == 0 ? ag[ch] + empty[p * 3 + ch] : (M == 0 ? base[p * 3 + ch] : (base[p * 3 + ch] * (M - sa) + ag[ch] * sa) / M));
        }
      }
    }
  }
}

/* ------------------------------------------------------------------------------------------ */
/* batch entry points (ctypes); env ranges are split over pthreads (no OpenMP runtime in image) */
/* ------------------------------------------------------------------------------------------ */
static int g_threads = 1;
void mgo_set_threads(int n) { g_threads = n > 0 ? n : 1; }
int mgo_get_threads(void) { return g_threads; }
int mgo_hw_threads(void) { long n = sysconf(_SC_NPROCESSORS_ONLN); return n > 0 ? (int)n : 1; }

typedef struct Job {
  int kind; /* 0 reset, 1 step, 2 obs_encode, 3 obs_rgb, 4 rollout (n_steps fused steps per env) */
  int64_t n_steps, B, pool; /* rollout: actions[(t % pool)][B][A] */
  const MgConfig* c; uint8_t* grid; uint8_t* agents; int32_t* envrec;
  uint64_t seed; int64_t env_offset; const uint8_t* mask; const int32_t* actions; double* rewards; uint8_t* done;
  int autoreset; const uint8_t* atlas; uint8_t* obs;
  int64_t lo, hi;
} Job;

static void* run_job(void* arg) {
  Job* j = (Job*)arg;
  const MgConfig* c = j->c;
  size_t enc_env = (size_t)c->n_agents * c->view_size * c->view_size * 3;
  size_t rgb_env = (size_t)c->n_agents * c->view_size * c->view_tile_size * c->view_size * c->view_tile_size * 3;
  for (int64_t e = j->lo; e < j->hi; ++e) {
    Env v = env_at(c, j->grid, j->agents, j->envrec, e);
    uint64_t g = (uint64_t)(j->env_offset + e);
    switch (j->kind) {
      case 0: if (!j->mask || j->mask[e]) env_reset(&v, j->seed, g); break;
      case 1: {
        int d = env_step(&v, j->actions + e * c->n_agents, j->rewards + e * c->n_agents, j->seed, g);
        j->done[e] = (uint8_t)d;
        if (d && j->autoreset) env_reset(&v, j->seed, g);
        if (j->obs) obs_encode_env(&v, j->obs + (size_t)e * enc_env);
      } break;
      case 4: /* time-major actions [n_steps][B][A]; every env is independent, so each thread runs its envs to the end */
        for (int64_t t = 0; t < j->n_steps; ++t) {
          int d = env_step(&v, j->actions + ((size_t)(t % j->pool) * j->B + e) * c->n_agents, j->rewards + e * c->n_agents, j->seed, g);
          j->done[e] = (uint8_t)d;
          if (d && j->autoreset) env_reset(&v, j->seed, g);
          obs_encode_env(&v, j->obs + (size_t)e * enc_env);
        }
        break;
      case 2: obs_encode_env(&v, j->obs + (size_t)e * enc_env); break;
      case 3: obs_rgb_env(&v, j->atlas, j->obs + (size_t)e * rgb_env); break;
    }
  }
  return 0;
}

static void run_parallel(Job* proto, int64_t B) {
  int n = g_threads;
  if (n > B) n = (int)(B > 0 ? B : 1);
  if (n <= 1) { proto->lo = 0; proto->hi = B; run_job(proto); return; }
  Job* jobs = (Job*)malloc(sizeof(Job) * (size_t)n);
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)n);
  for (int i = 0; i < n; ++i) {
    jobs[i] = *proto; jobs[i].lo = B * i / n; jobs[i].hi = B * (i + 1) / n;
    if (i > 0) pthread_create(&th[i], 0, run_job, &jobs[i]);
  }
  run_job(&jobs[0]);
  for (int i = 1; i < n; ++i) pthread_join(th[i], 0);
  free(jobs); free(th);
}

void mgo_init(const MgConfig* c, uint8_t* grid, uint8_t* agents, int32_t* envrec, int64_t B) {
  memset(grid, 0, (size_t)B * 3 * c->plane_stride);
  memset(agents, 0, (size_t)B * c->n_agents * MG_AGENT_REC);
  memset(envrec, 0, (size_t)B * MG_ENV_REC);
  for (int64_t e = 0; e < B; ++e)
    for (int a = 0; a < c->n_agents; ++a) agents[((size_t)e * c->n_agents + a) * 16 + 7] = 0xFF;
}

void mgo_reset(const MgConfig* c, uint8_t* grid, uint8_t* agents, int32_t* envrec, int64_t B, uint64_t seed,
               int64_t env_offset, const uint8_t* mask) {
  Job j; memset(&j, 0, sizeof j);
  j.kind = 0; j.c = c; j.grid = grid; j.agents = agents; j.envrec = envrec; j.seed = seed; j.env_offset = env_offset; j.mask = mask;
  run_parallel(&j, B);
}

/* obs may be NULL (step only) or receives the encoded obs of the post-step (post-autoreset) world */
void mgo_step(const MgConfig* c, uint8_t* grid, uint8_t* agents, int32_t* envrec, int64_t B, uint64_t seed,
              int64_t env_offset, const int32_t* actions, double* rewards, uint8_t* done, int autoreset, uint8_t* obs) {
  Job j; memset(&j, 0, sizeof j);
  j.kind = 1; j.c = c; j.grid = grid; j.agents = agents; j.envrec = envrec; j.seed = seed; j.env_offset = env_offset;
  j.actions = actions; j.rewards = rewards; j.done = done; j.autoreset = autoreset; j.obs = obs;
  run_parallel(&j, B);
}

/* n_steps x (step + auto-reset + encoded obs) for every env; outputs hold the last step's values;
 * actions is a pool of `pool` time-major action batches used cyclically */
void mgo_rollout(const MgConfig* c, uint8_t* grid, uint8_t* agents, int32_t* envrec, int64_t B, uint64_t seed, int64_t env_offset,
                 const int32_t* actions, int64_t n_steps, int64_t pool, double* rewards, uint8_t* done, int autoreset, uint8_t* obs) {
  Job j; memset(&j, 0, sizeof j);
  j.kind = 4; j.c = c; j.grid = grid; j.agents = agents; j.envrec = envrec; j.seed = seed; j.env_offset = env_offset;
  j.actions = actions; j.rewards = rewards; j.done = done; j.autoreset = autoreset; j.obs = obs; j.n_steps = n_steps; j.B = B; j.pool = pool > 0 ? pool : n_steps;
  run_parallel(&j, B);
}

void mgo_obs_encode(const MgConfig* c, uint8_t* grid, uint8_t* agents, int32_t* envrec, int64_t B, uint8_t* obs) {
  Job j; memset(&j, 0, sizeof j);
  j.kind = 2; j.c = c; j.grid = grid; j.agents = agents; j.envrec = envrec; j.obs = obs;
  run_parallel(&j, B);
}

void mgo_obs_rgb(const MgConfig* c, uint8_t* grid, uint8_t* agents, int32_t* envrec, int64_t B, const uint8_t* atlas,
                 uint8_t* obs) {
  Job j; memset(&j, 0, sizeof j);
  j.kind = 3; j.c = c; j.grid = grid; j.agents = agents; j.envrec = envrec; j.atlas = atlas; j.obs = obs;
  run_parallel(&j, B);
}

/* visibility masks only ([B][A][V][V], for debugging / LOS parity) */
void mgo_vis(const MgConfig* c, uint8_t* grid, uint8_t* agents, int32_t* envrec, int64_t B, uint8_t* out) {
  int V = c->view_size;
  for (int64_t e = 0; e < B; ++e) {
    Env v = env_at(c, grid, agents, envrec, e);
    ViewCell cells[MG_MAX_VIEW * MG_MAX_VIEW];
    for (int a = 0; a < c->n_agents; ++a) {
      uint8_t* m = out + ((size_t)e * c->n_agents + a) * V * V;
      memset(m, 0, (size_t)V * V);
      if (!gen_obs_grid(&v, a, cells, m)) memset(m, 0, (size_t)V * V);
    }
  }
}

void mgo_los_batch(const uint8_t* transparent, uint8_t* mask, int64_t n, int V, int ax, int ay) {
  for (int64_t i = 0; i < n; ++i) occlude_mask(transparent + (size_t)i * V * V, V, ax, ay, mask + (size_t)i * V * V);
}

void mgo_philox(const uint32_t* ctr, const uint32_t* key, uint32_t* out) { philox4x32_10(ctr, key, out); }
void mgo_order(uint64_t seed, uint64_t g, uint32_t t, int A, int32_t* order) {
  int o[MG_MAX_AGENTS]; draw_order(seed, g, t, A, o);
  for (int i = 0; i < A; ++i) order[i] = o[i];
}

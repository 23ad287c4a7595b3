"""CPU statement of the on-device policy hand-off (TEST INFRASTRUCTURE ONLY -- see oracle/mg_oracle.py).

The reference has no built-in policy: its README.md:43-57 loop calls `agents.action_step(obs)` on the host between two
`env.step` calls.  marlgrid_b200 closes that loop on the device for one policy family (include/marlgrid_b200.h:
mg_rollout_policy): an int8 linear layer per agent over the encoded observation + epsilon-greedy exploration.  This file
restates that contract in numpy and plays the closed loop on the C oracle (oracle/mg_oracle.c), so the GPU rollout can be
compared step by step.

Contract (g = global env index, t = lifetime steps of env g AFTER the step whose observation is used, a = agent, s = policy seed):
  logits[k] = bias[a][k] + sum_i int8 w[a][k][i] * uint8 obs[i]   (exact in int32), action = lowest k of the maximum, k < n_actions
  r = philox4x32_10(ctr = (lo32(g), hi32(g), t, 0x20000000 | a), key = (lo32(s), hi32(s)))
  if r[0] < epsilon_u32: action = mulhi32(r[1], n_actions)
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)
TAG_POLICY = 0x20000000


def philox4x32_10_np(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10 (Salmon et al., SC'11): counters are equal-shape arrays, key two python ints -> 4 uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(c).astype(np.uint64) & MASK for c in (c0, c1, c2, c3))
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    s32 = np.uint64(32)
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        c0, c1, c2, c3 = ((p1 >> s32) ^ c1 ^ np.uint64(k0)) & MASK, p1 & MASK, ((p0 >> s32) ^ c3 ^ np.uint64(k1)) & MASK, p0 & MASK
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def linear_policy_actions(obs, weights, bias, n_actions, epsilon_u32, seed, g, t_life):
    """obs uint8 [B][A][V][V][3], weights int8 [A][K][V*V*3], bias int32 [A][K], g int64 [B], t_life int [B] -> int32 [B][A]."""
    B, A = obs.shape[:2]
    o = obs.reshape(B, A, -1).astype(np.float64)  # |sum| <= V*V*3 * 255 * 128 < 2^53: the float64 matmul is exact
    logits = np.stack([o[:, a] @ weights[a].astype(np.float64).T for a in range(A)], axis=1).astype(np.int64) + bias.astype(np.int64)[None]
    assert np.all(np.abs(logits) < 2**31)
    act = np.argmax(logits[:, :, :n_actions], axis=2).astype(np.int32)  # first maximum = lowest k
    if epsilon_u32:
        g = np.asarray(g, np.uint64)
        gg = np.broadcast_to(g[:, None], (B, A))
        tt = np.broadcast_to(np.asarray(t_life).astype(np.uint64)[:, None], (B, A))
        aa = np.broadcast_to(np.arange(A, dtype=np.uint64)[None, :] | np.uint64(TAG_POLICY), (B, A))
        r0, r1, _, _ = philox4x32_10_np(gg & MASK, gg >> np.uint64(32), tt, aa, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
        explore = r0 < np.uint32(epsilon_u32)
        uni = ((r1.astype(np.uint64) * np.uint64(n_actions)) >> np.uint64(32)).astype(np.int32)
        act = np.where(explore, uni, act).astype(np.int32)
    return act


def closed_loop(ob, policy, first_actions, n_steps, autoreset=True):
    """Play n_steps steps on OracleBatch `ob`: step 0 = first_actions, then the policy's choices.  Returns per-step
    (obs, rewards, done, actions) arrays like env.rollout_policy."""
    B, A, V = ob.B, ob.A, ob.V
    obs = np.zeros((n_steps, B, A, V, V, 3), np.uint8)
    rew = np.zeros((n_steps, B, A), np.float64)
    done = np.zeros((n_steps, B), np.uint8)
    acts = np.zeros((n_steps, B, A), np.int32)
    g = np.arange(B, dtype=np.int64) + ob.env_offset
    a = np.ascontiguousarray(first_actions, np.int32).reshape(B, A)
    for t in range(n_steps):
        acts[t] = a
        obs[t], rew[t], done[t] = ob.step(a, autoreset=autoreset, with_obs=True)
        a = linear_policy_actions(obs[t], policy.weights, policy.bias, policy.n_actions, policy.epsilon_u32, policy.seed, g, ob.envrec[:, 2])
    return obs, rew, done, acts

"""Torch learners for the batched env -- the caller side of the hot path (the reference leaves learners
to the user: README.md:21-64 shows the protocol `action_step / save_step / start_episode / end_episode`, examples/human_player.py:8-32
one implementation).

`LinearQLearner` is an independent Q-learner per agent with a linear Q-function over the encoded view,
Q(o)[k] = W[k] . o / 16 + b[k].  It plugs into `IndependentLearners` like any learner (host loop: `action_step(obs)` between two
`env.step` calls), and -- because a linear layer with epsilon-greedy exploration is exactly what the rollout kernel can evaluate
itself -- `LinearQTrainer` runs the whole actor side on the device: weights quantised to int8 (`quantize()`), n steps played in ONE
kernel launch (`env.rollout_policy`, include/marlgrid_b200.h: mg_rollout_policy), then one TD(0) update on the returned batch.  With
`torch.distributed` initialised (one process per GPU, NCCL over NVLink) the gradients are averaged across ranks: the only collective
of the whole system, off the env's data path.
"""
import numpy as np
import torch

from .agents import LearningAgent
from .policy import LinearPolicy

OBS_SCALE = 1.0 / 16.0  # encoded bytes are small integers (type <= 13, colour <= 5, state <= 3): keeps the features O(1)


class LinearQLearner(LearningAgent):
    """One agent's linear Q-function.  kwargs beyond the learner's own go to GridAgentInterface (marlgrid/agents.py:19-35)."""

    def __init__(self, n_actions=7, gamma=0.95, lr=2e-3, epsilon=0.1, device="cpu", seed=0, target_period=0, **interface_kwargs):
        super().__init__(**interface_kwargs)
        self.n_actions, self.gamma, self.epsilon = int(n_actions), float(gamma), float(epsilon)
        n = self.view_size * self.view_size * 3
        g = torch.Generator().manual_seed(seed)
        self.W = (0.01 * torch.randn((self.n_actions, n), generator=g)).to(device).requires_grad_()
        self.b = torch.zeros((self.n_actions,), device=device, requires_grad=True)
        self.opt = torch.optim.Adam([self.W, self.b], lr=lr)
        # target_period > 0: bootstrap from a copy of the weights refreshed every target_period updates (steadier TD targets)
        self.target_period, self.updates = int(target_period), 0
        self.W_target, self.b_target = self.W.detach().clone(), self.b.detach().clone()
        self._gen = torch.Generator(device=device).manual_seed(seed + 1)
        self.buffer = []

    # ---- the float model ------------------------------------------------------------------------------------------
    def q_values(self, obs, target=False):
        """obs uint8 [..., V, V, 3] -> float32 [..., n_actions]."""
        x = obs.reshape(*obs.shape[:-3], -1).to(self.W.device, torch.float32) * OBS_SCALE
        if target and self.target_period > 0:
            return x @ self.W_target.t() + self.b_target
        return x @ self.W.t() + self.b

    # ---- learner protocol (README.md:21-25), batched: obs [B, V, V, 3] ----------------------------------------------
    def action_step(self, obs):
        with torch.no_grad():
            q = self.q_values(torch.as_tensor(obs))
            act = q.argmax(-1)
            if self.epsilon > 0:
                explore = torch.rand(act.shape, generator=self._gen, device=act.device) < self.epsilon
                uni = torch.randint(0, self.n_actions, act.shape, generator=self._gen, device=act.device)
                act = torch.where(explore, uni, act)
            return act.to(torch.int32)

    def save_step(self, obs, act, next_obs, rew, done):
        self.buffer.append((torch.as_tensor(obs), torch.as_tensor(act), torch.as_tensor(next_obs), torch.as_tensor(rew), torch.as_tensor(done)))

    def end_episode(self):
        if self.buffer:
            o, a, n, r, d = (torch.stack(x) for x in zip(*self.buffer))
            self.update(o, a, n, r, d)
            self.buffer = []

    # ---- TD(0) --------------------------------------------------------------------------------------------------------
    def td_loss(self, obs, act, next_obs, rew, done):
        """Mean squared TD error of a batch of transitions (any leading shape; done broadcasts over it)."""
        dev = self.W.device
        q = self.q_values(obs).gather(-1, act.to(dev, torch.int64).unsqueeze(-1)).squeeze(-1)
        with torch.no_grad():
            nxt = self.q_values(next_obs, target=True).max(-1).values
            target = rew.to(dev, torch.float32) + self.gamma * nxt * (1.0 - done.to(dev, torch.float32))
        return torch.mean((q - target) ** 2)

    def update(self, obs, act, next_obs, rew, done, process_group=None):
        """One optimiser step on the batch; gradients averaged over the ranks of torch.distributed when it is initialised."""
        self.opt.zero_grad(set_to_none=False)
        loss = self.td_loss(obs, act, next_obs, rew, done)
        loss.backward()
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            world = torch.distributed.get_world_size(process_group)
            if world > 1:
                flat = torch.cat([self.W.grad.reshape(-1), self.b.grad.reshape(-1)])
                torch.distributed.all_reduce(flat, group=process_group)  # NCCL over NVLink on the GPU box, gloo in the CPU tests
                flat /= world
                self.W.grad.copy_(flat[: self.W.numel()].view_as(self.W))
                self.b.grad.copy_(flat[self.W.numel():])
        self.opt.step()
        self.updates += 1
        if self.target_period > 0 and self.updates % self.target_period == 0:
            self.W_target.copy_(self.W.detach()); self.b_target.copy_(self.b.detach())
        return float(loss.detach())

    # ---- hand-off to the rollout kernel ---------------------------------------------------------------------------
    def quantize(self):
        """(int8 [n_actions][V*V*3], int32 [n_actions]): the Q-function scaled so that the largest weight is +-127.  argmax is
        invariant under the positive scale; rounding moves a logit by at most 0.5 * sum(obs) quantisation steps."""
        with torch.no_grad():
            w = (self.W * OBS_SCALE).detach().cpu().numpy().astype(np.float64)
            b = self.b.detach().cpu().numpy().astype(np.float64)
        s = 127.0 / max(float(np.abs(w).max()), 1e-12)
        return np.clip(np.rint(w * s), -127, 127).astype(np.int8), np.clip(np.rint(b * s), -(2**30), 2**30).astype(np.int32)


def quantized_policy(learners, epsilon, seed=0):
    """LinearPolicy (marlgrid_b200.policy) of a list of LinearQLearner, one per agent."""
    wq, bq = zip(*(l.quantize() for l in learners))
    return LinearPolicy(np.stack(wq), np.stack(bq), epsilon=epsilon, seed=seed, view_size=learners[0].view_size)


class LinearQTrainer:
    """Actor on the device, learner in torch: every iteration plays `horizon` steps of all envs in one launch with the current
    quantised policy, then updates each agent's Q-function on the (obs[t], act[t+1], rew[t+1], obs[t+1], done[t+1]) transitions."""

    def __init__(self, env, learners, horizon=32, epsilon=0.1, seed=0, process_group=None):
        assert len(learners) == env.num_agents
        self.env, self.learners, self.horizon, self.epsilon = env, list(learners), int(horizon), float(epsilon)
        self.seed, self.process_group, self.iteration = int(seed), process_group, 0
        B, A, V, T = env.num_envs, env.num_agents, env.cfg.view_size, self.horizon
        dev = env.device
        self.out = (torch.empty((T, B, A, V, V, 3), dtype=torch.uint8, device=dev), torch.empty((T, B, A), dtype=torch.float64, device=dev),
                    torch.empty((T, B), dtype=torch.bool, device=dev), torch.empty((T, B, A), dtype=torch.int32, device=dev))
        self.next_actions = None

    def iterate(self):
        """One rollout + one update per agent; returns {'reward_per_env_step', 'loss': [per agent], 'episodes'}."""
        env = self.env
        pol = quantized_policy(self.learners, self.epsilon, seed=self.seed + self.iteration)
        first = self.next_actions if self.next_actions is not None else env.policy_act(pol)
        obs, rew, done, act = env.rollout_policy(pol, first, self.horizon, out=self.out)
        self.next_actions = env.policy_act(pol)  # from the last observations: the next rollout continues the trajectory
        losses = []
        for k, l in enumerate(self.learners):
            losses.append(l.update(obs[:-1, :, k], act[1:, :, k], obs[1:, :, k], rew[1:, :, k], done[1:], process_group=self.process_group))
        self.iteration += 1
        return {"reward_per_env_step": float(rew.sum(-1).mean()), "loss": losses, "episodes": int(done.sum())}

"""GPU-resident rollout storage for batched PPO rollouts.

Mirrors the reference's single-episode ``RolloutBuffer`` (``agents/our_agents/ppo/rollout.py:14-198``: states, actions,
action_mask_history, log_prob_actions, values, rewards, dones; ``remember`` / ``clear`` / ``calculate_gae`` / ``unpack``)
for N environments stepping in lock-step for T steps:

  * every field is a ``[T, N, ...]`` tensor on the env's device, written in place - no host round trip per decision
    (the reference moves a (1,265) + (1,3510) float64 pair to the device and an action back, ``training.py:96-116``);
  * legal-action masks are kept BIT-PACKED (``[T, N, 110]`` int32, 440 B per decision instead of 28,080 B of float64) and
    expanded on demand in chunks (``action_masks``);
  * returns are the reference's discounted reward sums (``common.py:11-36``), restarted wherever ``dones`` marks the end
    of an episode (auto-reset puts several episodes in one column), normalised over the whole batch like the reference
    normalises over its episode; advantages = returns - values, normalised (``common.py:38-57``).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import torch

VERY_SMALL_EPSILON = 1e-8  # agents/our_agents/ppo/constants.py
MASK_WORDS, NUM_ACTIONS = 110, 3510


@dataclass
class BatchedRolloutBuffer:
    size: int                      # T: decisions per environment
    num_envs: int                  # N
    input_dim: int = 265
    device: torch.device | None = None
    dtype: torch.dtype = torch.float32
    index: int = field(default=0, init=False)
    full: bool = field(default=False, init=False)

    def __post_init__(self) -> None:
        if self.device is None:
            self.device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        T, N, d = self.size, self.num_envs, self.device
        self.states = torch.zeros((T, N, self.input_dim), dtype=self.dtype, device=d)
        self.actions = torch.zeros((T, N), dtype=torch.int16, device=d)
        self.action_mask_history = torch.zeros((T, N, MASK_WORDS), dtype=torch.int32, device=d)
        self.log_prob_actions = torch.zeros((T, N), dtype=self.dtype, device=d)
        self.values = torch.zeros((T, N), dtype=self.dtype, device=d)
        self.rewards = torch.zeros((T, N), dtype=self.dtype, device=d)
        self.dones = torch.zeros((T, N), dtype=torch.bool, device=d)

    def remember(self, state, action, action_mask, log_prob_action, value, reward, done) -> None:
        """Store one decision of every env (all arguments are [N, ...] tensors; action_mask is the bit-packed mask)."""
        if self.full:
            return
        t = self.index
        with torch.no_grad():
            self.states[t].copy_(state)
            self.actions[t].copy_(action)
            self.action_mask_history[t].copy_(action_mask)
            self.log_prob_actions[t].copy_(log_prob_action)
            self.values[t].copy_(value.reshape(self.num_envs))
            self.rewards[t].copy_(reward)
            self.dones[t].copy_(done)
        self.index += 1
        self.full = self.index >= self.size

    def clear(self) -> None:
        self.index, self.full = 0, False

    def discounted_returns(self, discount_factor: float) -> torch.Tensor:
        """[index, N] un-normalised returns: G_t = r_t + gamma * G_{t+1}, with G restarting after a done."""
        T = self.index
        out = torch.empty((T, self.num_envs), dtype=self.dtype, device=self.device)
        running = torch.zeros(self.num_envs, dtype=self.dtype, device=self.device)
        for t in range(T - 1, -1, -1):
            running = self.rewards[t] + discount_factor * torch.where(self.dones[t], torch.zeros_like(running), running)
            out[t] = running
        return out

    def calculate_gae(self, discount_factor: float, normalize: bool = True):
        """(advantages, returns), each [index, N], computed as the reference does (normalisation over the batch)."""
        with torch.no_grad():
            returns = self.discounted_returns(discount_factor)
            if normalize:
                returns = (returns - returns.mean()) / (returns.std() + VERY_SMALL_EPSILON)
            advantages = returns - self.values[: self.index]
            if normalize:
                advantages = (advantages - advantages.mean()) / (advantages.std() + VERY_SMALL_EPSILON)
        return advantages, returns

    def action_masks(self, lo: int, hi: int, dtype=torch.bool) -> torch.Tensor:
        """Dense 0/1 masks of flattened decisions lo..hi-1 ([hi-lo, 3510]) - expand only what a minibatch needs."""
        words = self.action_mask_history[: self.index].reshape(-1, MASK_WORDS)[lo:hi]
        shifts = torch.arange(32, device=words.device, dtype=torch.int32)
        return ((words.unsqueeze(-1) >> shifts) & 1).reshape(words.shape[0], -1)[:, :NUM_ACTIONS].to(dtype)

    def unpack(self, discount_factor: float):
        """Flattened (decision-major) views in the reference's order: states, actions, packed action masks,
        log_prob_actions, advantages, returns, dones."""
        T = self.index
        advantages, returns = self.calculate_gae(discount_factor)
        flat = lambda x: x[:T].reshape(T * self.num_envs, *x.shape[2:])
        return (flat(self.states), flat(self.actions), flat(self.action_mask_history), flat(self.log_prob_actions),
                advantages.reshape(-1), returns.reshape(-1), flat(self.dones))

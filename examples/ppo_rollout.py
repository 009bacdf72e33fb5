"""BASELINE.json config 4 end to end: 262,144 Splendor envs stepped on the GPU, feeding a policy with the reference PPO
actor's shape (MLP 265 -> 4 x 128 -> 3510 with a legal-action mask, agents/our_agents/ppo/network.py:77-116; random
weights here - the learner itself is out of scope).  Nothing leaves the device: observations, bit-packed masks, the
masked categorical sample and the env step are all CUDA tensors.

    python examples/ppo_rollout.py [--envs 262144] [--steps 32]
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from splendor_ai_b200.env import SplendorVecEnv, unpack_mask  # noqa: E402
from splendor_ai_b200.rollout import BatchedRolloutBuffer  # noqa: E402
from splendor_ai_b200.sampling import masked_sample  # noqa: E402


class Actor(nn.Module):
    """Same layer sizes as the reference's PPO actor; masking as the reference does (huge negative logit on illegal)."""

    def __init__(self, obs=265, hidden=128, actions=3510):
        super().__init__()
        layers, d = [], obs
        for _ in range(4):
            layers += [nn.Linear(d, hidden), nn.ReLU()]
            d = hidden
        self.body, self.head = nn.Sequential(*layers), nn.Linear(d, actions)

    def forward(self, obs, mask_bits=None):
        logits = self.head(self.body(obs))
        if mask_bits is None:  # raw logits: spl_masked_sample applies the bit-packed mask itself
            return logits
        legal = unpack_mask(mask_bits, torch.bool)
        return logits.masked_fill(~legal, float("-inf"))


def rollout(num_envs=262_144, steps=32, num_agents=2, seed=1234, chunk=65_536, fused_sampling=True, store=False):
    env = SplendorVecEnv(num_envs, num_agents, seed=seed)
    policy = Actor().to(env.device).eval()
    obs, info = env.reset()
    mask = info["legal_mask"]
    buffer = BatchedRolloutBuffer(steps, num_envs, device=env.device) if store else None
    finished = torch.zeros((), dtype=torch.int64, device=env.device)
    returns = torch.zeros(num_envs, device=env.device)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.no_grad():
        for t in range(steps):
            actions = torch.empty(num_envs, dtype=torch.int16, device=env.device)
            log_probs = torch.zeros(num_envs, device=env.device)
            for lo in range(0, num_envs, chunk):  # the dense [N, 3510] logits are produced chunk by chunk
                if fused_sampling:  # masked softmax + Categorical sample + log-prob in one kernel, from the mask bits
                    a, lp = masked_sample(policy(obs[lo:lo + chunk]), mask[lo:lo + chunk], seed, t, first_id=lo)
                    actions[lo:lo + chunk], log_probs[lo:lo + chunk] = a, lp
                else:               # the plain PyTorch way: dense boolean mask, masked_fill, Categorical
                    logits = policy(obs[lo:lo + chunk], mask[lo:lo + chunk])
                    actions[lo:lo + chunk] = torch.distributions.Categorical(logits=logits).sample().to(torch.int16)
            prev_obs, prev_mask = obs, mask
            obs, reward, terminated, truncated, info = env.step(actions)
            if buffer is not None:  # everything stays on the device; masks stay bit-packed (440 B per decision)
                buffer.remember(prev_obs, actions, prev_mask, log_probs, torch.zeros(num_envs, device=env.device), reward, terminated)
            mask = info["legal_mask"]
            assert not bool(info["illegal"].any())
            returns += reward
            finished += terminated.sum()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if buffer is not None:
        advantages, norm_returns = buffer.calculate_gae(0.99)
        assert advantages.shape == (steps, num_envs) and bool(torch.isfinite(norm_returns).all())
    return {"envs": num_envs, "steps": steps, "fused_sampling": fused_sampling, "stored": store, "env_step_calls_per_s": num_envs * steps / dt, "seconds": dt,
            "episodes_finished": int(finished.item()), "mean_reward_per_call": float(returns.mean().item()) / steps}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=262_144)
    ap.add_argument("--steps", type=int, default=32)
    ap.add_argument("--torch-sampling", action="store_true", help="sample with torch.distributions instead of spl_masked_sample")
    ap.add_argument("--store", action="store_true", help="also fill a BatchedRolloutBuffer and compute returns/advantages")
    a = ap.parse_args()
    print(json.dumps(rollout(a.envs, a.steps, fused_sampling=not a.torch_sampling, store=a.store)))

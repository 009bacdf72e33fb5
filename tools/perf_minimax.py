"""Time the depth-2 minimax search (spl_minimax2) on mid-game states.  Usage: python tools/perf_minimax.py [--envs 16384]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import splendor_ai_b200 as pkg

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, nargs="+", default=[16384])
ap.add_argument("--advance", type=int, default=30)
ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()
for N in a.envs:
    eng = pkg.BatchedSplendor(N, 2, seed=1234)
    skip = torch.full((N,), -1, dtype=torch.int16, device=eng.device)
    obs, rew, term, ill, mask = eng.env_step(None, skip, self_play=True)
    for t in range(a.advance):  # random self-play to a mid-game position
        logits = torch.zeros((N, 3510), device=eng.device)
        act, _ = pkg.sampling.masked_sample(logits, mask, seed=7, step=t)
        obs, rew, term, ill, mask = eng.env_step(None, act.to(torch.int16), self_play=True)
    for _ in range(2):
        out = eng.minimax2()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        best_action, best_value, child_action, child_value, offsets = eng.minimax2()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.reps
    M = int(offsets[-1].item())
    counts = eng.count_legal().double()
    # leaves = sum over root actions of the opponent's reply count; estimate it from the mean branching
    print(json.dumps({"kernel": "minimax2", "roots": N, "root_actions": M, "mean_branching": M / N, "ms": ms,
                      "roots_per_s": N / ms * 1e3, "root_actions_per_s": M / ms * 1e3,
                      "approx_leaf_evals_per_s": M * (M / N) / ms * 1e3}), flush=True)

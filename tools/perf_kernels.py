"""Time the single-call kernels (legal mask, step, observe, fused PPO env-step) on mid-game states and report achieved
HBM GB/s against algorithmic bytes: per-launch times on cold data, tools/cold_timing.py.
Usage: python tools/perf_kernels.py [--envs 262144] [--players 2] [--reps 5] [--only legal_mask observe step env_step masked_sample]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import splendor_ai_b200 as pkg
from splendor_ai_b200.env import unpack_mask

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from cold_timing import per_launch_ms  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, nargs="+", default=[262144])
ap.add_argument("--players", type=int, default=2)
ap.add_argument("--reps", type=int, default=5, help="timed graph replays (each replay is 8-16 launches on cold data)")
ap.add_argument("--advance", type=int, default=30, help="env-steps played before timing (mid-game states)")
ap.add_argument("--only", nargs="+", default=None, help="time only these kernels (legal_mask observe step env_step masked_sample)")
a = ap.parse_args()
PEAK = 6533.8


for N in a.envs:
    P = a.players
    W = (8 + 4 * P) * 4
    eng = pkg.BatchedSplendor(N, P, seed=1234, limit_rounds=True)
    gen = torch.Generator(device=eng.device).manual_seed(0)
    learner = torch.zeros(N, dtype=torch.int8, device=eng.device)
    skip = torch.full((N,), -1, dtype=torch.int16, device=eng.device)
    obs, rew, term, ill, mask = eng.env_step(learner, skip)

    def sample(mask):
        return torch.multinomial(unpack_mask(mask, torch.float32), 1, generator=gen).squeeze(1).to(torch.int16)

    for _ in range(a.advance // 2):
        obs, rew, term, ill, mask = eng.env_step(learner, sample(mask))
    saved = eng.state.clone()
    act = sample(mask)
    res = {}
    want = lambda k: a.only is None or k in a.only  # noqa: E731
    if want("legal_mask"):
        res["legal_mask"] = (per_launch_ms(torch, eng, saved, lambda o: eng.legal_mask(out=o), ((N, 110), torch.int32), replays=a.reps), W + 440)
    if want("observe"):
        res["observe"] = (per_launch_ms(torch, eng, saved, lambda o: eng.observe(out=o), ((N, 265), torch.float32), replays=a.reps), W + 1060)
    if want("step"):
        res["step"] = (per_launch_ms(torch, eng, saved, lambda o: eng.step(act), restore=True, replays=a.reps), 2 * W + 2 + 3)
    if want("env_step"):
        res["env_step"] = (per_launch_ms(torch, eng, saved, lambda o: eng.env_step(learner, act), restore=True, replays=a.reps), 2 * W + 440 + 1060 + 3 + 3)
    for k, (ms, bytes_per_env) in res.items():
        gbs = N * bytes_per_env / (ms * 1e-3) / 1e9
        print(json.dumps({"kernel": k, "envs": N, "players": P, "ms": ms, "envs_per_s": N / (ms * 1e-3), "bytes_per_env": bytes_per_env,
                          "GBps": gbs, "frac_of_hbm_peak": gbs / PEAK}))

# masked action sampling (spl_masked_sample): 3510 float32 logits + 110 mask words per row
from splendor_ai_b200.sampling import masked_sample  # noqa: E402

for N in (a.envs if a.only is None or "masked_sample" in a.only else []):
    rows = min(N, 262144)
    logits = torch.randn((rows, 3510), device="cuda")
    eng = pkg.BatchedSplendor(rows, a.players, seed=1)
    mask = eng.legal_mask()
    def timed(fn, reps=20):  # 3.7 GB of logits per call: the inputs exceed L2 by themselves
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    ms = timed(lambda: masked_sample(logits, mask, 1, 0))
    n_legal = float(torch.ops.aten.bitwise_and(mask.view(torch.uint8).unsqueeze(-1) >> torch.arange(8, device="cuda", dtype=torch.uint8), 1).sum()) / rows
    b = 440 + 32 * n_legal + 6  # mask bits + one 32-byte sector per legal logit (gather) + action/log-prob out
    print(json.dumps({"kernel": "masked_sample", "envs": rows, "ms": ms, "rows_per_s": rows / (ms * 1e-3), "legal_per_row": n_legal,
                      "bytes_per_env": b, "GBps": rows * b / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": rows * b / (ms * 1e-3) / 1e9 / PEAK,
                      "dense_equivalent_GBps": rows * (3510 * 4 + 446) / (ms * 1e-3) / 1e9}))
    break

"""Per-warp-step latency of the playout kernel against resident warps per scheduler: launches of 592*w warps (one
32-thread block per warp, w = 1..5 warps per scheduler on 148 SMs) and the time per step of the longest game.
Usage: SPL_PLAYOUT_BLOCK=32 python tools/perf_latency_curve.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import splendor_ai_b200 as pkg

lib = pkg._native.lib()
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream(dev).cuda_stream
for w in (1, 2, 3, 4, 5):
    G = 148 * 4 * 32 * w
    steps = torch.zeros(G, dtype=torch.int16, device=dev)
    ctr = torch.zeros(16, dtype=torch.int64, device=dev)
    for _ in range(3):
        lib.spl_playout(1234, 0, G, 2, 0, 1000, None, steps.data_ptr(), None, torch.zeros_like(ctr).data_ptr(), None, 0, None, stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for r in range(reps):
        lib.spl_playout(1234, 0, G, 2, 0, 1000, None, steps.data_ptr(), None, ctr.data_ptr(), None, 0, None, stream)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    s = steps.cpu().numpy().astype(int)
    wm = s.reshape(-1, 32).max(1)
    print(json.dumps({"warps_per_scheduler": w, "games": G, "ms": ms, "max_steps": int(s.max()), "warp_max_mean": float(wm.mean()),
                      "us_per_step_of_longest": 1e3 * ms / s.max(), "env_steps_per_s": int(ctr[1]) / reps / (ms * 1e-3)}))

"""Time per round of the playout kernel against the number of games in a launch: 148 * 32 * w games for w = 1 .. 19 (w stepping
warps per SM; up to 4-5 per scheduler).  Usage: python tools/perf_latency_curve.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import splendor_ai_b200 as pkg

lib = pkg._native.lib()
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream(dev).cuda_stream
for w in (1, 2, 4, 8, 12, 14, 16, 19):
    G = 148 * 32 * w
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
    print(json.dumps({"stepping_warps_per_sm": w, "games": G, "ms": ms, "max_steps": int(s.max()),
                      "us_per_step_of_longest": 1e3 * ms / s.max(), "env_steps_per_s": int(ctr[1]) / reps / (ms * 1e-3)}))

"""Time spl_playout (fused random playouts) for a few batch sizes; prints one JSON line per configuration.
Usage: [SPL_LIB=csrc/variants/x.so] python tools/perf_playout.py [--games 65536 1048576] [--players 2] [--reps 20]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import splendor_ai_b200 as pkg

ap = argparse.ArgumentParser()
ap.add_argument("--games", type=int, nargs="+", default=[65536])
ap.add_argument("--players", type=int, nargs="+", default=[2])
ap.add_argument("--reps", type=int, default=20)
a = ap.parse_args()
lib = pkg._native.lib()
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream(dev).cuda_stream
for P in a.players:
    for G in a.games:
        ctr = torch.zeros(16, dtype=torch.int64, device=dev)
        for w in range(3):
            lib.spl_playout(1234, w * G, G, P, 0, 1000, None, None, None, torch.zeros_like(ctr).data_ptr(), None, 0, None, stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for r in range(a.reps):
            lib.spl_playout(1234, (3 + r) * G, G, P, 0, 1000, None, None, None, ctr.data_ptr(), None, 0, None, stream)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        steps = int(ctr[1].item())
        print(json.dumps({"lib": os.path.basename(os.environ.get("SPL_LIB", "default")), "players": P, "games": G, "ms_per_launch": ms / a.reps,
                          "env_steps_per_s": steps / (ms * 1e-3), "frac_of_hbm_roofline": steps / (ms * 1e-3) * (160 if P == 2 else 224) / 6533.8e9}))

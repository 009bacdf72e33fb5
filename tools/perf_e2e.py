"""Time spl_playout_host (host deck orders in, host results out) for one batch size; prints one JSON line.
Usage: python tools/perf_e2e.py [--games 65536] [--players 2] [--reps 50]"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import splendor_ai_b200 as pkg

ap = argparse.ArgumentParser()
ap.add_argument("--games", type=int, default=65536)
ap.add_argument("--players", type=int, default=2)
ap.add_argument("--reps", type=int, default=50)
a = ap.parse_args()
lib = pkg._native.lib()
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream(dev).cuda_stream
G, P = a.games, a.players
sets = []
for s in range(4):
    lists = torch.empty((G, 90), dtype=torch.uint8, device=dev)
    nobles = torch.empty((G, 5), dtype=torch.uint8, device=dev)
    st = torch.empty(pkg.state.state_shape(G, P), dtype=torch.int32, device=dev)
    pkg._native.check(lib.spl_reset(st.data_ptr(), 1234, s * G, None, None, G, P, lists.data_ptr(), nobles.data_ptr(), stream), "spl_reset")
    sets.append((lists.cpu().pin_memory(), nobles.cpu().pin_memory()))
h_score = torch.empty((G, P), dtype=torch.int8).pin_memory()
h_steps = torch.empty(G, dtype=torch.int16).pin_memory()
h_kind = torch.empty(G, dtype=torch.uint8).pin_memory()
h_ctr = np.zeros(16, np.int64)


def call(s):
    lists, nobles = sets[s % 4]
    pkg._native.check(lib.spl_playout_host(lists.data_ptr(), nobles.data_ptr(), 1234, (s % 4) * G, G, P, 0, 1000, h_score.data_ptr(),
                                           h_steps.data_ptr(), h_kind.data_ptr(), h_ctr.ctypes.data_as(C.c_void_p), stream), "spl_playout_host")


for s in range(5):
    call(s)
h_ctr[:] = 0
torch.cuda.synchronize(dev)
t0 = time.perf_counter()
for s in range(a.reps):
    call(s)
torch.cuda.synchronize(dev)
dt = time.perf_counter() - t0
print(json.dumps({"players": P, "games": G, "ms_per_call": dt / a.reps * 1e3,
                  "env_steps_per_s": int(h_ctr[1]) / dt}))

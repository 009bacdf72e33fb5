"""The oracle's counters for bench.py's 2^24-game sweep (seed 1234, game ids 2^44 ..): ~1.5e9 env-steps of the C restatement,
about a quarter of an hour on 8 host threads.  Writes tests/golden/sweep_counters.json (compared by bench.py's `sweep` leg and
by tests/test_gpu_parity.py).  Usage: python tools/sweep_oracle.py [--games 16777216]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle as O

ap = argparse.ArgumentParser()
ap.add_argument("--games", type=int, default=1 << 24)
a = ap.parse_args()
NAMES = ("games", "steps", "wins0", "wins1", "wins2", "wins3", "ties0", "ties1", "ties2", "ties3", "end_score", "end_deadlock",
         "end_roundcap", "end_truncated", "score2_sum", "nobles")
total, t0 = np.zeros(16, np.int64), time.time()
for lo in range(0, a.games, 1 << 18):
    _, _, ctr = O.playout_many(1234, (1 << 44) + lo, min(1 << 18, a.games - lo), 2, False, 1000, nthreads=os.cpu_count() or 1, want_outputs=False)
    total += ctr
out = {"seed": 1234, "first_game": 1 << 44, "games": a.games, "players": 2, "source": "oracle/splendor_oracle.c (tools/sweep_oracle.py)",
       "seconds": time.time() - t0, "counters": dict(zip(NAMES, (int(x) for x in total)))}
if a.games == 1 << 24:
    json.dump(out, open(os.path.join(ROOT, "tests", "golden", "sweep_counters.json"), "w"), indent=1)
print(json.dumps(out))

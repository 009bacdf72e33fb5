"""One-off full-size parity check (too slow for pytest): every aggregate counter of BASELINE config 3 (2^20 four-player
games) and config 5 (2^24 two-player games) from the fused GPU playout against the CPU oracle on all host threads.
Also checks per-game lengths and scores on a 2^20-game slice.  Prints one JSON line per configuration.

--counters-from FILE: take the oracle's full-size counters from an earlier output of this tool (the oracle and the game
definition are unchanged since) and run the oracle only on the first 2^20 games of each configuration: ~1 min instead of ~7."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle as O  # noqa: E402
import splendor_ai_b200 as pkg  # noqa: E402

threads = os.cpu_count() or 1
recorded = {}
if len(sys.argv) > 2 and sys.argv[1] == "--counters-from":
    for line in open(sys.argv[2]):
        r = json.loads(line)
        assert r["counters_equal"], "the recorded run was not a parity run"
        recorded[r["config"]] = np.array([r["counters"][k] for k in pkg._native.CTR_NAMES], np.int64)
for label, P, G, seed in (("config3_4p_2^20", 4, 1 << 20, 1234), ("config5_2p_2^24", 2, 1 << 24, 1234)):
    t0 = time.perf_counter()
    out = pkg.playout_games(G, P, seed=seed, outputs=True)
    torch.cuda.synchronize()
    t_gpu = time.perf_counter() - t0
    gpu = out["counters"].cpu().numpy()
    total = np.zeros(16, np.int64)
    t0 = time.perf_counter()
    slice_ok = None
    for lo in range(0, G if label not in recorded else 1 << 20, 1 << 20):
        n = min(1 << 20, G - lo)
        sc, st, ctr = O.playout_many(seed, lo, n, P, False, 1000, nthreads=threads, want_outputs=(lo == 0))
        total += ctr
        if lo == 0:
            slice_ok = bool(np.array_equal(out["steps"][:n].cpu().numpy().astype(np.int32), st)
                            and np.array_equal(out["score2"][:n].cpu().numpy(), sc))
    t_cpu = time.perf_counter() - t0
    oracle_source = "oracle run on every game"
    if label in recorded:
        total = recorded[label]
        oracle_source = "oracle counters recorded in " + os.path.basename(sys.argv[2]) + "; oracle re-run on the first 2^20 games"
    print(json.dumps({"config": label, "games": G, "players": P, "counters_equal": bool(np.array_equal(gpu, total)),
                      "first_2^20_games_steps_and_scores_equal": slice_ok, "env_steps": int(gpu[1]),
                      "gpu_seconds_incl_outputs": round(t_gpu, 3), "oracle_seconds": round(t_cpu, 1), "oracle_threads": threads, "oracle_source": oracle_source,
                      "counters": dict(zip(pkg._native.CTR_NAMES, map(int, gpu)))}), flush=True)

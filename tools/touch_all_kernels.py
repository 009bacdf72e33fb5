"""Exercise every kernel once on small, awkwardly sized batches (for compute-sanitizer memcheck runs)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import splendor_ai_b200 as pkg
from splendor_ai_b200.env import SplendorVecEnv, unpack_mask
from splendor_ai_b200.sampling import masked_sample

for P in (2, 3, 4):
    for N in (1, 37, 131):
        eng = pkg.BatchedSplendor(N, P, seed=3, limit_rounds=True)
        lists, nobles = eng.reset(want_lists=True)
        mask = eng.legal_mask()
        obs = eng.observe()
        act, _ = masked_sample(torch.randn((N, 3510), device="cuda"), mask, 1, 0)
        r, d, ill = eng.step(act)
        assert int(ill.sum()) == 0
        child, a2, rew, par, off = eng.expand()
        child.legal_mask()
        eng.final_scores()
        eng.minimax_eval()
        if P == 2:
            eng.minimax2()
            menv = SplendorVecEnv(N, 2, seed=4, opponent="minimax")
            mo, minfo = menv.reset()
            for t in range(3):
                ma, _ = masked_sample(torch.randn((N, 3510), device="cuda"), minfo["legal_mask"], 3, t)
                mo, _, _, _, minfo = menv.step(ma)
        seats = ((eng.state[1, :, 3] >> 20) & 3).to(torch.int8)
        eng.env_step(torch.zeros(N, dtype=torch.int8), torch.full((N,), -1, dtype=torch.int16))
        eng.env_step(None, torch.full((N,), -1, dtype=torch.int16), self_play=True)
        eng.inject(lists.cpu().numpy(), nobles.cpu().numpy())
        out = eng.playout(max_steps=300, trace_len=64)
        pkg.playout_games(N, P, seed=5, trace_len=32, want_final_state=True)
        env = SplendorVecEnv(N, P, seed=9)
        o, info = env.reset()
        for t in range(12):
            a, lp = masked_sample(torch.randn((N, 3510), device="cuda"), info["legal_mask"], 2, t)
            o, rw, term, tr, info = env.step(a)
torch.cuda.synchronize()
print("touched all kernels")

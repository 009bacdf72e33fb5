"""generatePredecessor (splendor_model.py:156-220): oracle and device code vs tests/golden/undo.npz, recorded from the
reference by oracle/gen_golden_undo.py (successor -> predecessor -> snapshot at every step of 8 games)."""
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT


def load_undo():
    z = np.load(os.path.join(GOLDEN, "undo.npz"))
    off = np.concatenate([[0], np.cumsum(z["n_steps"])])
    for g in range(len(z["P"])):
        lo, hi = off[g], off[g + 1]
        yield dict(P=int(z["P"][g]), decks=z["decks"][g], nobles=z["nobles"][g][:int(z["P"][g]) + 1], action=z["action"][lo:hi],
                   card=z["card"][lo:hi], noble=z["noble"][lo:hi], returned=z["returned"][lo:hi],
                   after_undo=z["after_undo"][lo:hi], after_step=z["after_step"][lo:hi])


def test_oracle_undo_matches_reference(oracle):
    n = 0
    for g in load_undo():
        s = oracle.init(g["P"], g["decks"], g["nobles"])
        for t, a in enumerate(g["action"]):
            seat = s.cur
            trial = s.copy()
            assert oracle.step(trial, int(a)) >= 0
            trial.cur = seat  # generateSuccessor/Predecessor do not move the seat; GameRule.update does
            assert oracle.undo(trial, seat, int(a), int(g["card"][t]), int(g["noble"][t]), g["returned"][t]) == 0
            assert np.array_equal(trial.snapshot(), g["after_undo"][t]), (t, int(a))
            oracle.step(s, int(a))
            assert np.array_equal(s.snapshot(), g["after_step"][t]), t
            n += 1
    assert n > 500


def test_device_undo_matches_reference_hostsim(oracle):
    sys.path.insert(0, os.path.join(ROOT, "tests", "hostsim"))
    import sim as S

    from splendor_ai_b200 import state

    for g in load_undo():
        sim = S.Sim(g["P"])
        sim.inject(g["decks"], g["nobles"])
        for t, a in enumerate(g["action"]):
            trial = S.Sim(g["P"])
            trial.w[:], trial.lists = sim.w, sim.lists
            seat = int(state.unpack(sim.words, g["P"])["cur"][0])
            assert trial.step(int(a))[0] == 0
            trial.undo(seat, int(a), int(g["card"][t]), int(g["noble"][t]), g["returned"][t])
            assert np.array_equal(state.snapshots(trial.words, g["P"])[0], g["after_undo"][t]), (t, int(a))
            sim.step(int(a))
            assert np.array_equal(state.snapshots(sim.words, g["P"])[0], g["after_step"][t]), t


@pytest.mark.gpu
def test_gpu_undo_matches_reference():
    """spl_undo through the C ABI, all recorded games of one player count as one batch."""
    import torch

    import splendor_ai_b200 as pkg

    games = list(load_undo())
    for P in (2, 3, 4):
        gs = [g for g in games if g["P"] == P]
        if not gs:
            continue
        N, T = len(gs), max(len(g["action"]) for g in gs)
        eng = pkg.BatchedSplendor(N, P)
        eng.inject(np.stack([g["decks"] for g in gs]), np.stack([np.concatenate([g["nobles"], [255] * (5 - len(g["nobles"]))]) for g in gs]))
        for t in range(T):
            live = [i for i, g in enumerate(gs) if t < len(g["action"])]
            pick = lambda key, fill: np.array([g[key][t] if t < len(g["action"]) else fill for g in gs])
            act = torch.from_numpy(pick("action", -1).astype(np.int16))
            before = eng.state.clone()
            seats = pkg.state.unpack(eng.state_numpy(), P)["cur"]
            _, _, illegal = eng.step(act)
            assert int(illegal.sum()) == 0
            after = eng.state.clone()
            eng.undo(torch.from_numpy(seats.astype(np.int8)), act, torch.from_numpy(pick("card", 255).astype(np.uint8)),
                     torch.from_numpy(pick("noble", 255).astype(np.uint8)),
                     torch.from_numpy(np.stack([g["returned"][t] if t < len(g["action"]) else np.zeros(6, np.uint8) for g in gs])))
            sn = pkg.state.snapshots(eng.state_numpy(), P)
            for i in live:
                assert np.array_equal(sn[i], gs[i]["after_undo"][t]), (P, i, t)
            for i in set(range(N)) - set(live):
                assert torch.equal(eng.state[:, i], before[:, i])
            eng.state.copy_(after)


@pytest.mark.gpu
def test_facade_successor_predecessor_roundtrip():
    """SplendorGameRule.generateSuccessor / generatePredecessor as a depth-2 search uses them: state restored each time."""
    import random

    from splendor_ai_b200 import game_rule as G
    from splendor_ai_b200.state import snapshots

    random.seed(3)
    rule = G.SplendorGameRule(2)
    st = rule.current_game_state
    for ply in range(25):
        seat = rule.getCurrentAgentIndex()
        acts = rule.getLegalActions(st, seat)
        before = snapshots(st._words.reshape(-1, 1, 4), 2)[0].copy()
        for act in random.sample(acts, min(3, len(acts))):
            rule.generateSuccessor(st, act, seat)
            inner = rule.getLegalActions(st, 1 - seat)
            rule.generateSuccessor(st, inner[0], 1 - seat)
            rule.generatePredecessor(st, inner[0], 1 - seat)
            rule.generatePredecessor(st, act, seat)
            after = snapshots(st._words.reshape(-1, 1, 4), 2)[0]
            same = after == before
            same[18:23] = True  # a re-appended noble may sit at a different list position (reference quirk, :210)
            for p in range(2):
                same[28 + 20 * p + 11:28 + 20 * p + 14] = True  # and so may an un-bought reserved card (:201)
            assert same.all(), (ply, act["type"], np.nonzero(~same))
            assert sorted(after[18:23]) == sorted(before[18:23])
            assert all(sorted(after[28 + 20 * p + 11:28 + 20 * p + 14]) == sorted(before[28 + 20 * p + 11:28 + 20 * p + 14]) for p in range(2))
            if not np.array_equal(after, before):
                break  # list order changed: positions (hence action indices) shift; go on from the equivalent state
        rule.update(random.choice(rule.getLegalActions(st, seat)))
        st = rule.current_game_state
        if rule.gameEnds():
            break

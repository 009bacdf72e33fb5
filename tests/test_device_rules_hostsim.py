"""The DEVICE rule code (splendor-ai_b200/csrc/engine.cuh), compiled for the host by tests/hostsim, against the CPU oracle
and the reference recordings.  This covers every device function except the kernels' launch/synchronisation glue, which
the `-m gpu` tests cover on the B200."""
import importlib.util
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, load_replays

sys.path.insert(0, os.path.join(ROOT, "tests", "hostsim"))
import sim as S  # noqa: E402

_spec = importlib.util.spec_from_file_location("spl_state", os.path.join(ROOT, "splendor-ai_b200", "state.py"))
state = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(state)


def _check_state(sim, orc, where, lazy):
    snap = state.snapshots(sim.words, sim.P)[0]
    assert np.array_equal(snap, orc.snapshot()), (where, np.nonzero(snap != orc.snapshot()))
    if lazy:
        u = state.unpack(sim.words, sim.P)
        assert state.remaining_deck_sets(u["deck_masks"][0]) == orc.remaining_deck_sets(), where


def _check_decision(sim, orc, oracle, where):
    want = oracle.legal(orc)
    assert np.array_equal(sim.legal_select(), want), (where, "count/select")
    assert np.array_equal(sim.legal_member(), want), (where, "membership")
    assert np.array_equal(S.mask_to_indices(sim.mask()), want), (where, "mask row")
    assert np.array_equal(sim.mask(), oracle.legal_mask(orc)), where
    for seat in range(sim.P):
        a, b = sim.observe(seat), oracle.observe(orc, seat)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (where, seat, np.nonzero(a != b))


@pytest.mark.parametrize("name", ["replays_2p", "replays_3p", "replays_4p", "replays_2p_b", "replays_3p_b", "replays_4p_b"])
def test_reference_recordings_through_device_code(oracle, name):
    """Injected reference deck orders + recorded action traces: every state, legal set and observation identical."""
    for gi, g in enumerate(load_replays(name)):
        sim = S.Sim(g["P"])
        sim.inject(g["decks"], g["nobles"])
        orc = oracle.init(g["P"], g["decks"], g["nobles"])
        assert np.array_equal(state.snapshots(sim.words, sim.P)[0], g["snap0"])
        for t, a in enumerate(g["action"]):
            where = (name, gi, t)
            assert np.array_equal(sim.legal_select(), g["legal"][t]), where
            assert np.array_equal(sim.legal_member(), g["legal"][t]), where
            assert np.array_equal(S.mask_to_indices(sim.mask()), g["legal"][t]), where
            assert np.array_equal(sim.observe(int(g["seat"][t])).view(np.uint32), g["obs"][t].view(np.uint32)), where
            bad, rew, done = sim.step(int(a), flags=1 if g["limit"] else 0)
            assert (bad, rew) == (0, g["reward"][t]), where
            assert np.array_equal(state.snapshots(sim.words, sim.P)[0], g["snap"][t]), where
            assert (done != 0) == bool(g["ends_limit"][t] if g["limit"] else g["ends_plain"][t]), where
            oracle.step(orc, int(a))
        sc, oc, _ = sim.scores()
        assert list(sc) == list(g["score2"]), (name, gi)


@pytest.mark.parametrize("P,flags,n_games", [(2, 0, 60), (3, 0, 20), (4, 0, 30), (2, 1, 20), (4, 1, 10)])
def test_synthetic_playouts_match_oracle(oracle, P, flags, n_games):
    """Philox games: same deck orders, same action trace, same final state and scores as the oracle; and per step the
    three device formulations of the legal set agree with the oracle's enumeration."""
    seed = 1234
    for game in range(n_games):
        lists_o, nobles_o = oracle.synth_game(seed, game, P)
        sim = S.Sim(P)
        lists_s, nobles_s = sim.reset(seed, game, want_lists=True)
        assert np.array_equal(lists_s, lists_o) and np.array_equal(nobles_s[:P + 1], nobles_o), (P, game)
        steps_o, fin, trace_o, kind_o = oracle.playout(seed, game, P, bool(flags), 1000)
        steps_s, trace_s, sc, kind_s = sim.playout(seed, game, flags, 1000)
        assert (steps_s, kind_s) == (steps_o, kind_o), (P, game)
        assert np.array_equal(trace_s, trace_o), (P, game, np.nonzero(trace_s != trace_o)[0][:3])
        assert list(sc) == [oracle.cal_score2(fin, i) for i in range(P)]
        _check_state(sim, fin, (P, game, "final"), lazy=True)
        if game % 4 == 0:  # step-by-step with all cross-checks (lazy decks)
            sim.reset(seed, game)
            orc = oracle.init(P, lists_o, nobles_o)
            for t in range(steps_o):
                _check_decision(sim, orc, oracle, (P, game, t))
                bad, rew, done = sim.step(int(trace_o[t]), flags)
                assert bad == 0 and rew == oracle.step(orc, int(trace_o[t]))
                assert done == oracle.game_ends(orc, bool(flags))
                _check_state(sim, orc, (P, game, t), lazy=True)


def test_illegal_and_out_of_range_actions_leave_state_untouched(oracle):
    sim = S.Sim(2)
    sim.reset(7, 3)
    legal = set(sim.legal_member().tolist())
    before = sim.w.copy()
    for idx in [0, 1, 5, 3509, 3420, 3437, 6, -1, 3510, 20000]:
        if idx in legal:
            continue
        bad, rew, _ = sim.step(idx)
        assert bad == 1 and rew == 0
        after = sim.w.copy()
        after[7] &= ~np.uint32(state.META_ILLEGAL)
        assert np.array_equal(after, before), idx


def test_philox_matches_oracle(oracle):
    out = np.zeros(4, np.uint32)
    for seed, game, idx, stream in [(0, 0, 0, 0), (1234, 5, 17, 2), (2**63 + 11, 2**40 + 3, 4000000000, 5)]:
        S.lib().sim_philox(seed, game, idx, stream, out.ctypes.data)
        assert np.array_equal(out, oracle.philox(seed, game, idx, stream))
        S.lib().sim_philox_rk(seed, game, idx, stream, out.ctypes.data)  # the precomputed-key-schedule form of the playout kernel
        assert np.array_equal(out, oracle.philox(seed, game, idx, stream))
    # known-answer vectors of Philox4x32-10 from the Random123 distribution (kat_vectors)
    S.lib().sim_philox(0, 0, 0, 0, out.ctypes.data)
    assert [hex(x) for x in out] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    S.lib().sim_philox(0xFFFFFFFFFFFFFFFF, 0xFFFFFFFFFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, out.ctypes.data)
    assert [hex(x) for x in out] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]


def test_small_div_is_exact_on_its_domain():
    assert S.lib().sim_small_div_mismatches() == 0

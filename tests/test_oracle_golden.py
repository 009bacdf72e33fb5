"""Pins the CPU oracle (oracle/splendor_oracle.c) against recordings of the UNMODIFIED reference engine
(tests/golden/replays_*.npz, produced by oracle/gen_golden.py from splendor_model.py / features.py / gym/envs/utils.py).

Every step of every recorded game is compared bit-exactly: legal ALL_ACTIONS index set, observation vector of the seat
to move, reward, resulting state, both gameEnds variants, and calScore at the end."""
import numpy as np
import pytest

from conftest import load_replays


@pytest.mark.parametrize("name", ["replays_2p", "replays_3p", "replays_4p", "replays_2p_b", "replays_3p_b", "replays_4p_b"])
def test_oracle_replays_reference_games(oracle, name):
    games = load_replays(name)
    n_steps = 0
    for gi, g in enumerate(games):
        s = oracle.init(g["P"], g["decks"], g["nobles"])
        assert np.array_equal(s.snapshot(), g["snap0"]), (name, gi, "initial deal")
        for t in range(len(g["action"])):
            where = (name, gi, t, g["policy"])
            assert s.cur == g["seat"][t], where
            assert np.array_equal(oracle.legal(s), g["legal"][t]), where
            mask = oracle.legal_mask(s)
            assert int(sum(bin(int(w)).count("1") for w in mask)) == len(g["legal"][t]), where
            obs = oracle.observe(s, s.cur)
            assert np.array_equal(obs.view(np.uint32), g["obs"][t].view(np.uint32)), (where, np.nonzero(obs != g["obs"][t]))
            assert oracle.step(s, int(g["action"][t])) == g["reward"][t], where
            assert np.array_equal(s.snapshot(), g["snap"][t]), where
            assert (oracle.game_ends(s, False) != 0) == bool(g["ends_plain"][t]), where
            assert (oracle.game_ends(s, True) != 0) == bool(g["ends_limit"][t]), where
            n_steps += 1
        assert [oracle.cal_score2(s, i) for i in range(g["P"])] == list(g["score2"]), (name, gi)
        for i in range(g["P"]):
            assert np.array_equal(oracle.observe(s, i).view(np.uint32), g["final_obs"][i].view(np.uint32)), (name, gi, i)
    assert n_steps > 1000


def test_oracle_rejects_illegal_actions(oracle):
    g = load_replays("replays_2p")[0]
    s = oracle.init(g["P"], g["decks"], g["nobles"])
    legal = set(oracle.legal(s).tolist())
    before = s.snapshot()
    for idx in (0, 5, 3509, 3420, 7, -1, 3510, 99999):
        if idx not in legal:
            assert oracle.step(s, idx) == -1
            assert np.array_equal(s.snapshot(), before)

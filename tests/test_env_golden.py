"""Pins the ENV semantics (SURVEY.md 8a row "SplendorEnv.reset/step/_simulate_opponents/get_legal_actions_mask") to
recordings of the reference's own `SplendorEnv` (tests/golden/env_*.npz, oracle/gen_golden_env.py).

CPU part: the oracle, driven the way spl_env_step is specified (learner move, opponents until the learner's seat or the
end, then the learner's observation + legal set, 100-round cap), reproduces every recorded call bit for bit.  The GPU
part (tests/test_gpu_env_golden.py) replays the same files through spl_env_step, SplendorVecEnv and SplendorEnv."""
import numpy as np
import pytest

from conftest import load_env_episodes, scripted_choice


def opponents_until_learner(oracle, s, ep, log=None):
    """_simulate_opponents (splendor_env.py:219-231)."""
    P, me = ep["P"], ep["my_id"]
    while s.cur != me and not oracle.game_ends(s, True):
        legal = oracle.legal(s)
        t = sum(s.pl[i].turns for i in range(P))
        if ep["opponents"] == "philox":
            r = int(oracle.philox(ep["seed"], ep["game"], t, 0)[0])
            a = int(legal[(r * len(legal)) >> 32])
        else:
            a = scripted_choice([int(i) for i in legal], t)
        if log is not None:
            log.append(a)
        oracle.step(s, a)


@pytest.mark.parametrize("name", ["env_2p", "env_3p", "env_4p", "env_cap", "env_rand"])
def test_oracle_env_emulation_matches_reference_env(oracle, name):
    eps = load_env_episodes(name)
    calls = 0
    for gi, ep in enumerate(eps):
        s = oracle.init(ep["P"], ep["decks"], ep["nobles"])
        d, n = oracle.synth_game(ep["seed"], ep["game"], ep["P"])
        assert np.array_equal(d, ep["decks"]) and np.array_equal(n, ep["nobles"]), "recorded games are this repo's synthetic games"
        log = []
        for c in range(len(ep["action"])):
            where = (name, gi, c)
            if c > 0:
                log.append(int(ep["action"][c]))
                assert oracle.step(s, int(ep["action"][c])) == ep["reward"][c], where
            else:
                assert ep["action"][c] == -1 and ep["reward"][c] == 0
            opponents_until_learner(oracle, s, ep, log)
            if c == 0 and ep["pad"]:  # turn counters preset right after reset (see gen_golden_env.py)
                for i in range(ep["P"]):
                    s.pl[i].turns += ep["pad"]
            assert (oracle.game_ends(s, True) != 0) == bool(ep["terminated"][c]), where
            assert np.array_equal(s.snapshot(), ep["snap"][c]), where
            assert np.array_equal(oracle.legal(s, ep["my_id"]), ep["legal"][c]), where
            assert np.array_equal(oracle.observe(s, ep["my_id"]).view(np.uint32), ep["obs"][c].view(np.uint32)), where
            calls += 1
        assert ep["terminated"][-1] and not ep["terminated"][:-1].any()
        assert log == [int(a) for a in ep["all_actions"]], (name, gi)
    assert calls > 100


def test_env_recordings_cover_the_corners():
    eps = [e for n in ("env_2p", "env_3p", "env_4p", "env_cap") for e in load_env_episodes(n)]
    assert {e["P"] for e in eps} == {2, 3, 4}
    for P in (2, 3, 4):
        assert {e["my_id"] for e in eps if e["P"] == P and e["opponents"] == "philox"} == set(range(P))
    # games that ended after an OPPONENT's move (the seat to move is 0 again and the learner is not the last seat)
    assert sum(1 for e in eps if e["snap"][-1][27] == 0 and e["my_id"] != e["P"] - 1 and e["opponents"] == "philox") >= 5
    # the 100-round cap of LimitRoundsGameRule: every agent has made exactly 100 turns
    capped = [e for e in eps if all(e["snap"][-1][28 + 20 * i + 18] == 100 for i in range(e["P"]))]
    assert {e["P"] for e in capped} == {2, 3, 4}

"""State fuzzing of the device rule code (host build) against the oracle on arbitrary hand-built states."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tests", "hostsim"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import sim as S  # noqa: E402
from fuzz_states import random_state  # noqa: E402


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_fuzzed_states_legal_sets_observations_and_successors(oracle, seed):
    from splendor_ai_b200 import state

    rng = np.random.default_rng(seed)
    seen = dict(pass_only=0, hold10=0, multi_noble=0, cap7=0, max_legal=0)
    for it in range(1500):
        orc, f, lists = random_state(rng, oracle)
        P = orc.P
        sim = S.Sim(P)
        sim.w[:] = state.pack(f, P).reshape(-1)
        sim.lists = lists
        assert np.array_equal(state.snapshots(sim.words, P)[0], orc.snapshot()), (seed, it, "pack/unpack")
        want = oracle.legal(orc)
        assert np.array_equal(sim.legal_select(), want), (seed, it, "count/select")
        assert np.array_equal(sim.legal_member(), want), (seed, it, "membership")
        assert np.array_equal(sim.mask(), oracle.legal_mask(orc)), (seed, it, "mask row")
        for seat in range(P):
            assert np.array_equal(sim.observe(seat).view(np.uint32), oracle.observe(orc, seat).view(np.uint32)), (seed, it, seat)
        for flags in (0, 1):
            sc, oc, kind = sim.scores(flags)
            assert kind == oracle.game_ends(orc, bool(flags))
            assert list(sc) == [oracle.cal_score2(orc, i) for i in range(P)]
        seen["pass_only"] += int(want[0] < 6)
        seen["hold10"] += int(sum(orc.pl[orc.cur].gems) == 10)
        seen["multi_noble"] += int(len({(i - 3438) % 6 for i in want if i >= 3438}) > 2)
        seen["cap7"] += int(7 in list(orc.pl[orc.cur].cards))
        seen["max_legal"] = max(seen["max_legal"], len(want))
        for a in rng.choice(want, size=min(4, len(want)), replace=False):  # successors of a few legal actions
            o2, s2 = orc.copy(), S.Sim(P)
            s2.w[:], s2.lists = sim.w, lists
            bad, rew, done = s2.step(int(a))
            assert bad == 0 and rew == oracle.step(o2, int(a)), (seed, it, int(a))
            assert np.array_equal(state.snapshots(s2.words, P)[0], o2.snapshot()), (seed, it, int(a))
            assert done == oracle.game_ends(o2, False)
    assert seen["hold10"] > 100 and seen["cap7"] > 100 and seen["pass_only"] > 0 and seen["multi_noble"] > 0, seen

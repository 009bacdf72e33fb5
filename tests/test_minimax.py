"""MiniMaxAgent (agents/our_agents/minmax.py): the oracle's restatement against values recorded from the reference
(tests/golden/minimax.npz, oracle/gen_golden_minimax.py), then the CUDA search against the oracle (-m gpu).

Tolerances: the evaluation is float64 arithmetic in the reference's operation order, so evaluation and per-action
depth-2 values are compared BIT-EXACTLY.  The action the reference agent returns is checked as a member of the arg-max
set within 1e-9: inside its own walk generatePredecessor re-orders the reserved list, so its float sums carry last-bit
noise that depends on its random.shuffle; the tie-break order (type name, then shuffle) is restated in select_best."""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
TYPE_RANK = {"buy_available": 0, "buy_reserve": 1, "collect_diff": 2, "collect_same": 3, "pass": 4, "reserve": 5}


def type_rank(idx: int) -> int:
    """Rank of the action's type string in the reference's sort (minmax.py:59 sorts by the "type" value)."""
    for base, name in ((3438, "buy_available"), (3420, "buy_reserve"), (1140, "collect_diff"), (510, "collect_same"),
                       (6, "reserve"), (0, "pass")):
        if idx >= base:
            return TYPE_RANK[name]
    raise ValueError(idx)


def golden_roots(oracle):
    g = np.load(os.path.join(HERE, "golden", "minimax.npz"))
    a_off = np.concatenate([[0], np.cumsum(g["n_steps"])])
    v_off = np.concatenate([[0], np.cumsum(g["s_n"])])
    for gi in range(len(g["n_steps"])):
        s = oracle.init(2, g["decks"][gi], g["nobles"][gi][:3])
        acts = g["action"][a_off[gi]:a_off[gi + 1]]
        samples = {int(g["s_step"][k]): k for k in np.nonzero(g["s_game"] == gi)[0]}
        for t, a in enumerate(acts):
            if t in samples:
                k = samples[t]
                yield s, dict(evals=g["s_evals"][k], actions=g["s_actions"][v_off[k]:v_off[k + 1]],
                              values=g["s_values"][v_off[k]:v_off[k + 1]], chosen=int(g["s_chosen"][k]))
            assert oracle.step(s, int(a)) >= 0
        yield s, dict(evals=g["final_evals"][gi], actions=None)


def bits(x):
    return np.asarray(x, np.float64).view(np.uint64)


def test_oracle_evaluation_and_depth2_values_match_the_reference(oracle):
    n_roots = n_vals = n_final = 0
    for s, rec in golden_roots(oracle):
        got = [oracle.minimax_eval(s, i) for i in range(2)]
        assert np.array_equal(bits(got), bits(rec["evals"])), (got, rec["evals"])
        if rec["actions"] is None:
            assert abs(got[0]) > 99999 and got[0] == -got[1]  # finished games take the max_score >= 15 branch
            n_final += 1
            continue
        acts, vals = oracle.minimax2(s)
        assert np.array_equal(acts, rec["actions"].astype(np.uint16))
        assert np.array_equal(bits(vals), bits(rec["values"])), np.abs(vals - rec["values"]).max()
        best = vals.max()
        top = [int(a) for a, v in zip(acts, vals) if v >= best - 1e-9]
        first_type = min(type_rank(a) for a in top)
        assert rec["chosen"] in top and type_rank(rec["chosen"]) == first_type, (rec["chosen"], top)
        n_roots += 1
        n_vals += len(acts)
    assert n_roots >= 70 and n_vals >= 1000 and n_final == 6


# ------------------------------------------------------------------------------------------------ CUDA (through the C ABI)
@pytest.fixture(scope="module")
def pkg():
    import splendor_ai_b200 as p

    p.build()
    return p


def select_best(acts, vals):
    """The engine's deterministic tie-break: maximum value, then lowest type rank, then lowest index."""
    best = max(vals)
    return min((type_rank(int(a)), int(a)) for a, v in zip(acts, vals) if v == best)[1]


@pytest.mark.gpu
def test_gpu_minimax_on_the_reference_recordings(pkg, oracle):
    """Every sampled root of tests/golden/minimax.npz on the GPU (deck orders injected, recorded actions replayed up to
    the sample): the evaluation for both seats and all 1,072 depth-2 values bit-exact with what the reference computed."""
    import torch

    g = np.load(os.path.join(HERE, "golden", "minimax.npz"))
    a_off = np.concatenate([[0], np.cumsum(g["n_steps"])])
    v_off = np.concatenate([[0], np.cumsum(g["s_n"])])
    n_games, S = len(g["n_steps"]), len(g["s_game"])
    game = np.concatenate([g["s_game"], np.arange(n_games)])            # the sampled roots, then every final position
    stop = np.concatenate([g["s_step"], g["n_steps"]]).astype(np.int64)  # actions to replay before looking
    N = len(game)
    eng = pkg.BatchedSplendor(N, 2)
    eng.inject(g["decks"][game], g["nobles"][game])
    for t in range(int(stop.max())):
        act = np.array([g["action"][a_off[game[e]] + t] if t < stop[e] else -1 for e in range(N)], np.int16)
        _, _, illegal = eng.step(torch.from_numpy(act))
        assert not illegal.any()
    ev = [eng.minimax_eval(torch.full((N,), s, dtype=torch.int8)).cpu().numpy() for s in range(2)]
    want = np.concatenate([g["s_evals"], g["final_evals"]])
    for s in range(2):
        assert np.array_equal(bits(ev[s]), bits(want[:, s])), s
    eng.minimax_eval()  # seat to move: just has to run
    best_action, best_value, child_action, child_value, offsets = (x.cpu().numpy() for x in eng.minimax2())
    for k in range(S):
        lo, hi = offsets[k], offsets[k + 1]
        acts, vals = g["s_actions"][v_off[k]:v_off[k + 1]], g["s_values"][v_off[k]:v_off[k + 1]]
        assert np.array_equal(child_action[lo:hi], acts), k
        assert np.array_equal(bits(child_value[lo:hi]), bits(vals)), (k, np.abs(child_value[lo:hi] - vals).max())
        assert best_action[k] == select_best(acts, vals) and best_value[k] == vals.max(), k
        top = [int(a) for a, v in zip(acts, vals) if v >= vals.max() - 1e-9]
        assert int(g["s_chosen"][k]) in top and type_rank(int(g["s_chosen"][k])) == type_rank(int(best_action[k])), k


@pytest.mark.gpu
def test_gpu_minimax_matches_oracle_on_counter_based_games(pkg, oracle):
    """Mid- and late-game states of 256 synthetic games (Philox decks: the replies deal replacement cards from the
    counter-based deck), all root actions: child values, chosen action and value vs the oracle, bit-exact."""
    import torch

    N, seed = 256, 2024
    eng = pkg.BatchedSplendor(N, 2, seed=seed, first_game=900)
    orcs = [oracle.init(2, *oracle.synth_game(seed, 900 + e, 2)) for e in range(N)]
    rng = np.random.default_rng(5)
    n_checked = 0
    for rounds in (10, 30, 26):  # look after 10, 40 and 66 plies
        for _ in range(rounds):
            act = np.array([rng.choice(oracle.legal(o)) if not oracle.game_ends(o, False) else -1 for o in orcs], np.int16)
            eng.step(torch.from_numpy(act))
            for o, a in zip(orcs, act):
                if a >= 0:
                    oracle.step(o, int(a))
        ev = [eng.minimax_eval(torch.full((N,), s, dtype=torch.int8)).cpu().numpy() for s in range(2)]
        best_action, best_value, child_action, child_value, offsets = (x.cpu().numpy() for x in eng.minimax2())
        for e, o in enumerate(orcs):
            assert [ev[0][e], ev[1][e]] == [oracle.minimax_eval(o, 0), oracle.minimax_eval(o, 1)], e
            acts, vals = oracle.minimax2(o)
            lo, hi = offsets[e], offsets[e + 1]
            assert np.array_equal(child_action[lo:hi], acts.astype(np.int16)), e
            assert np.array_equal(bits(child_value[lo:hi]), bits(vals)), e
            assert best_action[e] == select_best(acts, vals) and best_value[e] == vals.max(), e
            n_checked += len(acts)
    assert n_checked > 10000


@pytest.mark.gpu
def test_gpu_minimax_on_fuzzed_states(pkg, oracle):
    """Arbitrary hand-built 2-player states (tests/fuzz_states.py) with explicit deck lists: evaluation and search."""
    import torch
    from fuzz_states import random_state

    N, rng = 400, np.random.default_rng(77)
    orcs, fields, lists = [], [], []
    for _ in range(N):
        o, f, l = random_state(rng, oracle, 2)
        orcs.append(o); fields.append(f); lists.append(l)
    merged = {k: np.concatenate([f[k] for f in fields]) for k in fields[0]}
    eng = pkg.BatchedSplendor(N, 2)
    eng.state.copy_(torch.from_numpy(pkg.state.pack(merged, 2).view(np.int32)))
    eng.decks = torch.from_numpy(np.stack(lists)).to(eng.device)
    ev = [eng.minimax_eval(torch.full((N,), s, dtype=torch.int8)).cpu().numpy() for s in range(2)]
    best_action, best_value, child_action, child_value, offsets = (x.cpu().numpy() for x in eng.minimax2())
    for e, o in enumerate(orcs):
        assert [ev[0][e], ev[1][e]] == [oracle.minimax_eval(o, 0), oracle.minimax_eval(o, 1)], e
        acts, vals = oracle.minimax2(o)
        lo, hi = offsets[e], offsets[e + 1]
        assert np.array_equal(child_action[lo:hi], acts.astype(np.int16)), e
        assert np.array_equal(bits(child_value[lo:hi]), bits(vals)), e
        assert best_action[e] == select_best(acts, vals), e


@pytest.mark.gpu
def test_minimax_agent_class_plays_a_whole_game_like_the_oracle(pkg, oracle):
    """The drop-in MiniMaxAgent (splendor-ai_b200/minimax.py) inside the reference-shaped Game loop against a RandomAgent:
    every decision equals the oracle's depth-2 search on the replayed position, and the search agent wins."""
    from splendor_ai_b200 import game_rule as G
    from splendor_ai_b200 import minimax, template
    from splendor_ai_b200.game import Game

    assert minimax.myAgent is minimax.MiniMaxAgent
    log = []

    class LoggedSearch(minimax.MiniMaxAgent):
        def SelectAction(self, actions, game_state, game_rule):
            a = super().SelectAction(actions, game_state, game_rule)
            log.append((self.id, G.action_index(a, game_state, self.id), self.last_value, self.evaluate(game_state)))
            return a

    class LoggedRandom(template.RandomAgent):
        def SelectAction(self, actions, game_state, game_rule):
            a = super().SelectAction(actions, game_state, game_rule)
            log.append((self.id, G.action_index(a, game_state, self.id), None, None))
            return a

    g = Game(G.SplendorGameRule, [LoggedSearch(0), LoggedRandom(1)], 2, seed=5, agents_namelist=["minimax", "random"], time_limit=60)
    st = g.game_rule.current_game_state
    s = oracle.init(2, st._deck_lists.copy(), [G.NOBLE_ID[n[0]] for n in st.board.nobles])
    h = g.Run()
    assert len(log) == len(h["actions"]) > 20
    n_search = 0
    for seat, idx, value, ev in log:
        assert seat == s.cur
        if seat == 0:
            acts, vals = oracle.minimax2(s)
            assert idx == select_best(acts, vals) and value == vals.max() and ev == oracle.minimax_eval(s, 0), n_search
            n_search += 1
        assert oracle.step(s, idx) >= 0
    assert n_search >= 10 and h["scores"][0] > h["scores"][1]


@pytest.mark.gpu
def test_vec_env_with_minimax_opponent_matches_oracle_emulation(pkg, oracle):
    """SplendorVecEnv(opponent="minimax"): learner move, the opponent's depth-2 search reply, the learner's next rows -
    against the same loop written with the oracle (100-round cap, next-step auto-reset left out: 40 calls per env)."""
    import torch

    from splendor_ai_b200.env import SplendorVecEnv, unpack_mask

    N, seed = 48, 99
    env = SplendorVecEnv(N, 2, seed=seed, opponent="minimax")
    obs, info = env.reset(seed=3)
    me = info["my_id"].cpu().numpy()
    orcs = [oracle.init(2, *oracle.synth_game(seed, e, 2)) for e in range(N)]

    def reply(o, e):
        if o.cur != me[e] and not oracle.game_ends(o, True):
            acts, vals = oracle.minimax2(o)
            oracle.step(o, select_best(acts, vals))

    for e, o in enumerate(orcs):
        reply(o, e)
    rng = np.random.default_rng(0)
    alive = np.ones(N, bool)
    for t in range(40):
        mask = unpack_mask(info["legal_mask"], torch.uint8).cpu().numpy()
        obs_np = obs.cpu().numpy()
        act = np.zeros(N, np.int16)
        for e, o in enumerate(orcs):
            if not alive[e]:
                continue
            legal = oracle.legal(o, int(me[e]))
            assert np.array_equal(np.nonzero(mask[e])[0], legal), (e, t)
            assert np.array_equal(obs_np[e].view(np.uint32), oracle.observe(o, int(me[e])).view(np.uint32)), (e, t)
            act[e] = rng.choice(legal)
        obs, reward, terminated, truncated, info = env.step(torch.from_numpy(act))
        reward, terminated = reward.cpu().numpy(), terminated.cpu().numpy()
        for e, o in enumerate(orcs):
            if not alive[e]:
                continue
            assert reward[e] == oracle.step(o, int(act[e])), (e, t)
            reply(o, e)
            assert bool(terminated[e]) == bool(oracle.game_ends(o, True)), (e, t)
            if terminated[e]:
                alive[e] = False  # the next call resets this env; the emulation stops following it
    assert alive.sum() < N  # the searched opponent finished some games inside 40 calls

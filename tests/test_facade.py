"""Host-side mirror of the reference interface (splendor-ai_b200/{game_rule,env,template,dist}.py)."""
import os
import random
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, load_replays


def test_same_seed_gives_the_reference_game():
    """random.seed(s) + SplendorState draws == the reference's BoardState.__init__ draws (needs /root/reference)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_harness

    if not ref_harness.available():
        pytest.skip("reference tree not present")
    ref_harness.install()
    from splendor.splendor.splendor_model import SplendorGameRule as RefRule

    from splendor_ai_b200 import game_rule as G

    for P, seed in [(2, 1), (3, 99), (4, 12345)]:
        random.seed(seed)
        ref = RefRule(P)
        b = ref.current_game_state.board
        random.seed(seed)
        lists, nobles = G.draw_initial_decks(P)
        assert [G.NOBLES[i][0] for i in nobles] == [n[0] for n in b.nobles]
        for t in range(3):
            mine = lists[G.TIER_BASE[t]:G.TIER_BASE[t] + G.TIER_SIZE[t]]
            ref_list = [c.code for c in b.decks[t]] + [b.dealt[t][j].code for j in (3, 2, 1, 0)]
            assert [G.card_of(i).code for i in mine] == ref_list
        # and the card objects carry the reference's attributes
        c = b.dealt[2][1]
        m = G.card_of(G.CARD_ID[c.code])
        assert (m.colour, m.code, m.cost, m.deck_id, m.points) == (c.colour, c.code, c.cost, c.deck_id, c.points)


def _worker(rank, size, port, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(size), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    from splendor_ai_b200 import dist as D

    r, _, s = D.init("gloo")
    lo, hi = D.shard_range(1001, r, s)
    ctr = torch.zeros(16, dtype=torch.int64)
    ctr[0], ctr[1], ctr[2 + r] = hi - lo, sum(range(lo, hi)), 1
    D.allreduce_counters(ctr)
    q.put((r, lo, hi, ctr.tolist(), D.allreduce_max(float(r + 1), "cpu")))
    D.barrier()


def test_sharding_and_counter_allreduce_world_size_2():
    """The N>1 path on CPU (gloo): contiguous shards cover every game once; the int64[16] counter block sums."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, c0, m0), (r1, lo1, hi1, c1, m1) = out
    assert (lo0, hi0, lo1, hi1) == (0, 501, 501, 1001)
    assert c0 == c1 and c0[0] == 1001 and c0[1] == sum(range(1001)) and c0[2] == 1 and c0[3] == 1
    assert m0 == m1 == 2.0


def test_shard_range_partitions():
    from splendor_ai_b200.dist import shard_range

    for total in (0, 1, 7, 65536, 2**24 + 3):
        for size in (1, 2, 4, 8):
            edges = [shard_range(total, r, size) for r in range(size)]
            assert edges[0][0] == 0 and edges[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
            assert max(h - l for l, h in edges) - min(h - l for l, h in edges) <= 1


@pytest.mark.gpu
def test_game_rule_facade_replays_reference_recordings():
    """SplendorGameRule facade (dict actions in, reference attributes out) over the recorded reference games."""
    from splendor_ai_b200 import game_rule as G
    from splendor_ai_b200.state import snapshots

    for name in ("replays_2p", "replays_4p"):
        for g in load_replays(name)[:3]:
            P = g["P"]
            rule = (G.LimitRoundsGameRule if g["limit"] else G.SplendorGameRule)(P)
            rule.current_game_state = G.SplendorState(P, _deck_lists=g["decks"], _nobles=list(g["nobles"]))
            st = rule.current_game_state
            assert np.array_equal(snapshots(st._words.reshape(-1, 1, 4), P)[0], g["snap0"])
            for t, a in enumerate(g["action"]):
                seat = rule.getCurrentAgentIndex()
                assert seat == g["seat"][t]
                acts = rule.getLegalActions(st, seat)
                idxs = [G.action_index(x, st, seat) for x in acts]
                assert idxs == list(g["legal"][t])
                chosen = acts[idxs.index(int(a))]
                before = st.agents[seat].score
                rule.update(chosen)
                st = rule.current_game_state
                assert st.agents[seat].score - before == g["reward"][t]
                assert np.array_equal(snapshots(st._words.reshape(-1, 1, 4), P)[0], g["snap"][t]), (name, t)
                assert rule.gameEnds() == bool(g["ends_limit"][t] if g["limit"] else g["ends_plain"][t])
                assert len(st.agents[seat].agent_trace.action_reward) == int(g["snap"][t][28 + 20 * seat + 18])
            assert [int(round(2 * rule.calScore(st, i))) for i in range(P)] == list(g["score2"])
            bought = sum(len(v) for c, v in st.agents[0].cards.items() if c != "yellow")
            assert bought == int(g["snap"][-1][28 + 6:28 + 11].sum())


@pytest.mark.gpu
def test_game_rule_facade_rejects_illegal_actions():
    from splendor_ai_b200 import game_rule as G

    random.seed(5)
    rule = G.SplendorGameRule(2)
    st = rule.current_game_state
    with pytest.raises(ValueError):
        rule.generateSuccessor(st, 3437, 0)  # buy a reserved card nobody holds
    with pytest.raises(ValueError):
        rule.generateSuccessor(st, {"type": "collect_same", "collected_gems": {"red": 2}, "returned_gems": {"blue": 1}, "noble": None}, 0)


@pytest.mark.gpu
def test_single_env_episode_with_host_agents():
    """SplendorEnv (splendor_env.py contract): reset/step tuples, reward = my score delta, ValueErrors, termination."""
    from splendor_ai_b200 import template
    from splendor_ai_b200.env import SplendorEnv

    random.seed(11)
    np.random.seed(11)
    env = SplendorEnv(agents=[template.RandomAgent(0)], fixed_turn=1)
    obs, info = env.reset(seed=11)
    assert obs.shape == (265,) and obs.dtype == np.float32 and info == {"my_id": 1}
    assert env.action_space.n == 3510 and env.observation_space.shape == (265,)
    with pytest.raises(ValueError):
        env.step(3510)
    mask = env.get_legal_actions_mask()
    assert mask.shape == (3510,) and mask.dtype == np.float64
    with pytest.raises(ValueError):
        env.step(int(np.nonzero(mask == 0)[0][0]))
    total, steps, terminated = 0, 0, False
    while not terminated and steps < 120:
        mask = env.get_legal_actions_mask()
        a = int(np.random.choice(np.nonzero(mask)[0]))
        obs, reward, terminated, truncated, _ = env.step(a)
        assert truncated is False and obs[0] == 1.0
        total += reward
        steps += 1
    assert terminated and total == env.state.agents[1].score == int(obs[2])


@pytest.mark.gpu
def test_vec_env_rollout_with_auto_reset(oracle):
    """SplendorVecEnv: masked sampling on the GPU, no illegal moves, next-step auto-reset onto fresh game ids, and one
    env's second episode replayed through the oracle from its (seed, game id)."""
    from splendor_ai_b200.env import SplendorVecEnv, unpack_mask

    N, P, seed = 2048, 2, 77
    env = SplendorVecEnv(N, P, seed=seed, first_game=0, fixed_turn=0)
    obs, info = env.reset()
    assert obs.shape == (N, 265) and info["my_id"].shape == (N,)
    gen = torch.Generator(device=env.device).manual_seed(0)
    episodes = torch.zeros(N, dtype=torch.int64, device=env.device)
    returns = torch.zeros(N, device=env.device)
    track, trace = 5, []
    for it in range(160):
        dense = unpack_mask(env.legal_mask, torch.float32)
        assert bool((dense.sum(1) > 0).all())
        a = torch.multinomial(dense, 1, generator=gen).squeeze(1)
        was_done = env._done.clone()
        obs, reward, term, trunc, info = env.step(a)
        assert not bool(info["illegal"].any()) and not bool(trunc.any())
        if int(episodes[track]) == 1 and not bool(was_done[track]):
            trace.append(int(a[track]))
        fresh = was_done  # these envs were reset by this call: first observation of a new game
        if bool(fresh.any()):
            assert bool((obs[fresh][:, 1] <= 1).all()) and bool((reward[fresh] == 0).all())
        episodes += was_done.to(torch.int64)
        returns += reward
    assert int(episodes.min()) >= 1  # every env finished at least one game (100-round cap => <= 100 learner moves)
    assert bool((env._episode == episodes).all())
    # replay env `track`'s second episode: game id = first_game + track + 1*N, learner seat 0, random opponent from Philox
    s = oracle.init(P, *oracle.synth_game(seed, track + N, P))
    for a in trace:
        assert s.cur == 0 and a in set(oracle.legal(s).tolist())
        oracle.step(s, a)
        while s.cur != 0 and not oracle.game_ends(s, True):
            legal = oracle.legal(s)
            r = int(oracle.philox(seed, track + N, sum(s.pl[i].turns for i in range(P)), 0)[0])
            oracle.step(s, int(legal[(r * len(legal)) >> 32]))
    assert len(trace) > 10


def test_tally_matches_the_reference_runner_rules():
    """win = unique best score, tie = shared best, else loss (general_game_runner.py:416-430)."""
    from splendor_ai_b200.game import tally

    t = tally([[15, 12], [16.5, 16], [15, 15], [3, 18], [15.5, 15.5]])
    assert t["wins"] == [2, 1] and t["ties"] == [2, 2] and t["loses"] == [1, 2]
    assert t["total_scores"] == [65.0, 76.5] and t["win_rates"] == [40.0, 20.0]


@pytest.mark.gpu
def test_game_loop_history_and_replayer():
    """Game.Run with host RandomAgents: the reference's history dict, then GameReplayer reproduces the final state."""
    from splendor_ai_b200 import game_rule as G
    from splendor_ai_b200 import template
    from splendor_ai_b200.game import Game, GameReplayer
    from splendor_ai_b200.state import snapshots

    g = Game(G.SplendorGameRule, [template.RandomAgent(0), template.RandomAgent(1)], 2, seed=21, agents_namelist=["a", "b"])
    h = g.Run()
    assert set(h) >= {"actions", "seed", "num_of_agent", "agents_namelist", "warning_positions", "warning_limit", "scores"}
    assert h["seed"] == 21 and len(h["actions"]) > 20 and set(h["scores"]) == {0, 1}
    for i, item in enumerate(h["actions"]):
        ((idx, info),) = item.items()
        assert idx == i and info["agent_id"] == i % 2 and "type" in info["action"]
    final = g.game_rule.current_game_state
    assert max(a.score for a in final.agents) >= 15 or all(a.passed for a in final.agents)
    assert h["scores"][0] == g.game_rule.calScore(final, 0)
    r = GameReplayer(G.SplendorGameRule, h)
    replayed = r.Run()
    assert np.array_equal(snapshots(replayed._words.reshape(-1, 1, 4), 2), snapshots(final._words.reshape(-1, 1, 4), 2))


@pytest.mark.gpu
def test_random_tournament_statistics(oracle):
    """The batched `splendor -m N` for random agents: the runner's matches dict, equal to the oracle's tallies."""
    from splendor_ai_b200.game import run_random_tournament

    N, seed = 5000, 4
    m = run_random_tournament(N, 2, seed=seed)
    sc, st, ctr = oracle.playout_many(seed, 0, N, 2, False, 1000, nthreads=os.cpu_count() or 1)
    assert m["wins"] == [int(ctr[2]), int(ctr[3])] and m["ties"] == [int(ctr[6]), int(ctr[7])]
    assert m["loses"] == [N - m["wins"][i] - m["ties"][i] for i in range(2)]
    assert m["total_scores"] == (sc.astype(np.float64).sum(0) / 2).tolist()
    assert abs(sum(m["win_rates"]) + 100.0 * m["ties"][0] / N - 100.0) < 1e-9 and m["succ"]


@pytest.mark.gpu
def test_ppo_shaped_policy_rollout_example():
    """examples/ppo_rollout.py on a small batch: a reference-shaped actor (265 -> 4x128 -> 3510, masked) drives the
    vector env for two episodes' worth of steps without ever choosing an illegal action."""
    sys.path.insert(0, os.path.join(ROOT, "examples"))
    import ppo_rollout

    out = ppo_rollout.rollout(num_envs=4096, steps=120, chunk=2048, store=True)
    assert out["episodes_finished"] >= 4096 and out["env_step_calls_per_s"] > 0 and out["stored"]

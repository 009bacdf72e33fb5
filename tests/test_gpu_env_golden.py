"""`-m gpu`: the recordings of the reference's own SplendorEnv (tests/golden/env_*.npz, oracle/gen_golden_env.py)
replayed through the C ABI's spl_env_step, through SplendorVecEnv and through the drop-in SplendorEnv class - every
observation (float32 bit pattern), reward, terminated flag, legal index set and state snapshot identical, call by call."""
import numpy as np
import pytest
import torch

from conftest import load_env_episodes, scripted_choice

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg():
    import splendor_ai_b200 as p

    p.build()
    return p


def unpack_mask(mask_row):
    bits = np.unpackbits(np.ascontiguousarray(mask_row).view(np.uint8), bitorder="little")
    return np.nonzero(bits)[0].astype(np.uint16)


def snaps(eng, pkg):
    return pkg.state.snapshots(eng.state_numpy(), eng.P)


def add_turns(eng, P, pad):
    """Preset every seat's turn counter (+pad): bits 0..15 of word 3 of each player's chunk (include/splendor_b200.h)."""
    for p in range(P):
        eng.state[2 + p, :, 3] += pad


def nobles5(eps):
    return np.stack([np.concatenate([e["nobles"], [255] * (5 - len(e["nobles"]))]) for e in eps]).astype(np.uint8)


@pytest.mark.parametrize("name", ["env_2p", "env_3p", "env_4p", "env_rand"])
def test_spl_env_step_replays_reference_env(pkg, name):
    """spl_env_step with injected deck orders (explicit-list decks) and per-env game ids: learner move, in-kernel
    Philox opponents, learner's mask + observation - against what the reference env returned."""
    eps = load_env_episodes(name)
    N, P, seed = len(eps), eps[0]["P"], eps[0]["seed"]
    eng = pkg.BatchedSplendor(N, P, seed=seed, limit_rounds=True)
    eng.inject(np.stack([e["decks"] for e in eps]), nobles5(eps), game_ids=torch.tensor([e["game"] for e in eps]))
    learner = torch.tensor([e["my_id"] for e in eps], dtype=torch.int8)
    for c in range(max(len(e["action"]) for e in eps)):
        act = np.array([int(e["action"][c]) if c < len(e["action"]) else -1 for e in eps], np.int16)
        obs, reward, term, illegal, mask = (x.cpu().numpy() for x in eng.env_step(learner, torch.from_numpy(act)))
        sn = snaps(eng, pkg)
        for i, e in enumerate(eps):
            k = min(c, len(e["action"]) - 1)  # a finished episode stays where it ended (its action is -1 from then on)
            where = (name, i, c)
            assert illegal[i] == 0, where
            if c < len(e["action"]):
                assert reward[i] == e["reward"][c], where
            assert bool(term[i]) == bool(e["terminated"][k]), where
            assert np.array_equal(sn[i], e["snap"][k]), where
            assert np.array_equal(unpack_mask(mask[i]), e["legal"][k]), where
            assert np.array_equal(obs[i].view(np.uint32), e["obs"][k].view(np.uint32)), where


@pytest.mark.parametrize("name", ["env_2p", "env_3p", "env_4p"])
def test_vec_env_replays_reference_env(pkg, name):
    """SplendorVecEnv (counter-based Philox decks, auto-reset) plays the recorded games from their (seed, game id) alone:
    the recorded games ARE this repo's synthetic games, so nothing is injected."""
    eps = load_env_episodes(name)
    P, seed = eps[0]["P"], eps[0]["seed"]
    for seat in range(P):
        group = [e for e in eps if e["my_id"] == seat]
        ids = [e["game"] for e in group]
        assert ids == list(range(ids[0], ids[0] + len(ids)))
        venv = pkg.SplendorVecEnv(len(group), P, seed=seed, first_game=ids[0], fixed_turn=seat)
        obs, info = venv.reset()
        assert info["my_id"].tolist() == [seat] * len(group)
        obs, mask = obs.cpu().numpy(), info["legal_mask"].cpu().numpy()
        for i, e in enumerate(group):
            assert np.array_equal(obs[i].view(np.uint32), e["obs"][0].view(np.uint32)), (name, seat, i)
            assert np.array_equal(unpack_mask(mask[i]), e["legal"][0]), (name, seat, i)
        for c in range(1, max(len(e["action"]) for e in group)):
            live = [c < len(e["action"]) for e in group]
            # an env whose recorded episode is over has been auto-reset to a new game: play any legal action there
            m = venv.get_legal_actions_mask(packed=True).cpu().numpy()
            act = [int(e["action"][c]) if live[i] else int(unpack_mask(m[i])[0]) for i, e in enumerate(group)]
            obs, reward, term, trunc, info = venv.step(torch.tensor(act, dtype=torch.int16))
            obs, reward, term, mask = obs.cpu().numpy(), reward.cpu().numpy(), term.cpu().numpy(), info["legal_mask"].cpu().numpy()
            assert not bool(trunc.any()) and not bool(info["illegal"].any())
            for i, e in enumerate(group):
                if not live[i]:
                    continue
                where = (name, seat, i, c)
                assert reward[i] == e["reward"][c] and bool(term[i]) == bool(e["terminated"][c]), where
                assert np.array_equal(obs[i].view(np.uint32), e["obs"][c].view(np.uint32)), where
                assert np.array_equal(unpack_mask(mask[i]), e["legal"][c]), where


def test_self_play_replay_of_scripted_episodes_reaches_the_round_cap(pkg):
    """Every action of every seat of the scripted (hoarding) episodes, one per spl_env_step(SELF_PLAY) call: the state
    after each learner-facing call, the terminated flag at every move and the 100-round cap (LimitRoundsGameRule)."""
    for gi, e in enumerate(load_env_episodes("env_cap")):
        P, me = e["P"], e["my_id"]
        eng = pkg.BatchedSplendor(1, P, seed=e["seed"], limit_rounds=True)
        eng.inject(e["decks"][None], nobles5([e]), game_ids=torch.tensor([e["game"]]))
        call = 0

        def at_call_boundary(obs, mask, term):
            nonlocal call
            sn = snaps(eng, pkg)[0]
            seat = int(sn[27])
            if seat != me and not term:
                return
            assert np.array_equal(sn, e["snap"][call]), (gi, call)
            assert bool(term) == bool(e["terminated"][call]), (gi, call)
            if seat == me:  # the rows are the next mover's
                assert np.array_equal(unpack_mask(mask), e["legal"][call]), (gi, call)
                assert np.array_equal(obs.view(np.uint32), e["obs"][call].view(np.uint32)), (gi, call)
            call += 1

        acts = [int(a) for a in e["all_actions"]]
        if e["pad"]:  # opponents seated before the learner move first, then the turn counters are preset
            for a in acts[:me]:
                eng.env_step(None, torch.tensor([a], dtype=torch.int16), self_play=True, want_mask=False, want_obs=False)
            acts = acts[me:]
            add_turns(eng, P, e["pad"])
        obs, _, term, _, mask = (x.cpu().numpy() for x in eng.env_step(None, torch.tensor([-1], dtype=torch.int16), self_play=True))
        at_call_boundary(obs[0], mask[0], term[0])
        for a in acts:
            assert not term[0]
            obs, _, term, illegal, mask = (x.cpu().numpy() for x in eng.env_step(None, torch.tensor([int(a)], dtype=torch.int16), self_play=True))
            assert illegal[0] == 0
            at_call_boundary(obs[0], mask[0], term[0])
        assert term[0] and call == len(e["action"]), (gi, call)


def test_dropin_env_class_replays_scripted_episodes(pkg, monkeypatch):
    """The drop-in SplendorEnv class (host-side Agent opponents, GPU rule object) against the reference env: same
    scripted opponents, same injected decks, same learner actions -> same observations, rewards, masks, terminations."""
    from splendor_ai_b200 import game_rule as GR

    class ScriptedAgent(pkg.template.Agent):
        def SelectAction(self, actions, game_state, game_rule):
            by_index = {GR.action_index(a, game_state, self.id): a for a in actions}
            t = sum(len(ag.agent_trace.action_reward) for ag in game_state.agents)
            return by_index[scripted_choice(sorted(by_index), t)]

    for gi, e in enumerate(load_env_episodes("env_cap")):
        P = e["P"]
        monkeypatch.setattr(GR, "draw_initial_decks", lambda n, e=e: ([int(c) for c in e["decks"]], [int(x) for x in e["nobles"]]))
        env = pkg.SplendorEnv([ScriptedAgent(0) for _ in range(P - 1)], shuffle_turns=False, fixed_turn=e["my_id"])
        obs, info = env.reset()
        assert info == {"my_id": e["my_id"]}
        if e["pad"]:
            st = env.state
            for p in range(P):
                st._words[8 + 4 * p + 3] += np.uint32(e["pad"])
                st._traces[p].action_reward.extend([({"type": "pass", "noble": None}, 0)] * e["pad"])
            st._refresh()
            obs = env._vectorize(st, env.my_turn)
        for c in range(len(e["action"])):
            if c > 0:
                obs, reward, terminated, truncated, _ = env.step(int(e["action"][c]))
                assert reward == e["reward"][c] and bool(terminated) == bool(e["terminated"][c]) and truncated is False, (gi, c)
            assert np.array_equal(np.asarray(obs, np.float32).view(np.uint32), e["obs"][c].view(np.uint32)), (gi, c)
            assert np.array_equal(np.nonzero(env.get_legal_actions_mask())[0], e["legal"][c]), (gi, c)
        with pytest.raises(ValueError):
            env.step(3510)


def test_policy_opponents_match_oracle_emulation(pkg, oracle):
    """SplendorVecEnv(opponent=policy_fn): the seats that are not the learner's are moved by a frozen policy on the device
    (reference: ppo/ppo.py:176-196 + ppo_agent.py:32-64 - own-seat observation, masked network, arg-max).  The policy here
    has exactly reproducible integer logits, a function of the agent's turn count (observation entry 1) and the action
    index; the whole vector env - auto-reset included - is followed by an oracle emulation for three players."""
    N, P, seed, seat = 48, 3, 77, 1
    idx = torch.arange(3510, device="cuda", dtype=torch.int64)

    def policy(obs, mask, seats):
        turn = obs[:, 1].to(torch.int64)
        return ((idx[None, :] * 7919 + turn[:, None] * 104729) % 1009).to(torch.float32)

    def opponent_choice(s):
        legal = oracle.legal(s).astype(np.int64)
        vals = (legal * 7919 + int(s.pl[s.cur].turns) * 104729) % 1009
        return int(legal[int(np.argmax(vals))])  # first maximum = lowest index, as the arg-max of spl_masked_sample

    venv = pkg.SplendorVecEnv(N, P, seed=seed, first_game=500, fixed_turn=seat, opponent=policy)
    episode = np.zeros(N, np.int64)
    orcs = [oracle.init(P, *oracle.synth_game(seed, 500 + e, P)) for e in range(N)]

    def opponents(e):
        s = orcs[e]
        while s.cur != seat and not oracle.game_ends(s, True):
            oracle.step(s, opponent_choice(s))

    obs, info = venv.reset()
    for e in range(N):
        opponents(e)
    rng = np.random.default_rng(3)
    done = np.zeros(N, bool)  # the game of env e ended at the previous step: the next step auto-resets it and ignores the action
    finished = 0
    for it in range(75):
        o, m = obs.cpu().numpy(), info["legal_mask"].cpu().numpy()
        act = np.zeros(N, np.int16)
        for e in range(N):
            legal = oracle.legal(orcs[e], seat)
            assert np.array_equal(unpack_mask(m[e]), legal), (e, it)
            assert np.array_equal(o[e].view(np.uint32), oracle.observe(orcs[e], seat).view(np.uint32)), (e, it)
            act[e] = rng.choice(legal)
        obs, reward, term, trunc, info = venv.step(torch.from_numpy(act))
        reward, term = reward.cpu().numpy(), term.cpu().numpy()
        for e in range(N):
            if done[e]:  # auto-reset: a new game id, the opponents seated before the learner move, no reward
                episode[e] += 1
                orcs[e] = oracle.init(P, *oracle.synth_game(seed, 500 + e + episode[e] * N, P))
                opponents(e)
                assert reward[e] == 0 and not term[e], (e, it)
                done[e] = False
                continue
            assert reward[e] == oracle.step(orcs[e], int(act[e])), (e, it)
            opponents(e)
            assert bool(term[e]) == (oracle.game_ends(orcs[e], True) != 0), (e, it)
            if term[e]:
                done[e] = True
                finished += 1
    assert finished >= N // 2


def test_vec_env_raises_before_it_changes_anything(pkg):
    """raise_on_illegal=True: the reference raises ValueError before it mutates (splendor_env.py:139-153); the vector env
    tests the action bits of its current masks first, so a caught error leaves states, masks and bookkeeping untouched."""
    N = 64
    venv = pkg.SplendorVecEnv(N, 2, seed=5, raise_on_illegal=True)
    obs, info = venv.reset()
    before = venv.state.clone()
    mask = info["legal_mask"].clone()
    legal = [int(unpack_mask(m)[0]) for m in mask.cpu().numpy()]
    bad = torch.tensor(legal, dtype=torch.int16)
    bad[17] = 3509 if 3509 not in unpack_mask(mask[17].cpu().numpy()).tolist() else 0
    with pytest.raises(ValueError):
        venv.step(bad)
    assert torch.equal(venv.state, before) and torch.equal(venv.legal_mask, mask) and not bool(venv._done.any())
    with pytest.raises(ValueError):
        venv.step(torch.full((N,), 3510, dtype=torch.int16))
    assert torch.equal(venv.state, before)
    obs2, reward, term, trunc, info2 = venv.step(torch.tensor(legal, dtype=torch.int16))  # and it still works afterwards
    assert not bool(info2["illegal"].any())

"""`-m gpu` parity tests: the CUDA kernels, called through the C ABI, against the CPU oracle and the reference
recordings.  Integer / index work: bit-exact.  The observation's two non-integer features (np.var, log1p) are compared
by float32 bit pattern as well (tolerance 0)."""
import os

import numpy as np
import pytest
import torch

from conftest import load_replays

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


@pytest.mark.parametrize("P", [2, 3, 4])
def test_reset_matches_oracle(pkg, oracle, P):
    N = 300
    eng = pkg.BatchedSplendor(N, P, seed=77, first_game=5)
    lists, nobles = eng.reset(want_lists=True)
    lists, nobles, sn = lists.cpu().numpy(), nobles.cpu().numpy(), snaps(eng, pkg)
    u = pkg.state.unpack(eng.state_numpy(), P)
    for e in range(N):
        lo, no = oracle.synth_game(77, 5 + e, P)
        assert np.array_equal(lists[e], lo) and np.array_equal(nobles[e][:P + 1], no), e
        s = oracle.init(P, lo, no)
        assert np.array_equal(sn[e], s.snapshot()), e
        assert pkg.state.remaining_deck_sets(u["deck_masks"][e]) == s.remaining_deck_sets(), e


@pytest.mark.parametrize("name", ["replays_2p", "replays_3p", "replays_4p", "replays_2p_b", "replays_3p_b", "replays_4p_b"])
def test_reference_recordings_replayed_on_gpu(pkg, name):
    """The REFERENCE's own recorded games (deck orders injected, recorded actions replayed): every legal mask,
    observation, reward, terminal flag, state and final score identical to what the Python engine produced."""
    games = load_replays(name)
    P = games[0]["P"]
    for limit in (False, True):
        gs = [g for g in games if g["limit"] == limit]
        if not gs:
            continue
        N, T = len(gs), max(len(g["action"]) for g in gs)
        eng = pkg.BatchedSplendor(N, P, limit_rounds=limit)
        eng.inject(np.stack([g["decks"] for g in gs]), np.stack([np.concatenate([g["nobles"], [255] * (5 - len(g["nobles"]))]) for g in gs]))
        assert np.array_equal(snaps(eng, pkg), np.stack([g["snap0"] for g in gs]))
        for t in range(T):
            live = [i for i, g in enumerate(gs) if t < len(g["action"])]
            mask = eng.legal_mask().cpu().numpy()
            seat = torch.tensor([int(g["seat"][t]) if t < len(g["action"]) else 0 for g in gs], dtype=torch.int8)
            obs = eng.observe(seat).cpu().numpy()
            for i in live:
                assert np.array_equal(unpack_mask(mask[i]), gs[i]["legal"][t]), (name, i, t)
                assert np.array_equal(obs[i].view(np.uint32), gs[i]["obs"][t].view(np.uint32)), (name, i, t)
            act = torch.tensor([int(g["action"][t]) if t < len(g["action"]) else -1 for g in gs], dtype=torch.int16)
            reward, done, illegal = (x.cpu().numpy() for x in eng.step(act))
            sn = snaps(eng, pkg)
            for i in live:
                g = gs[i]
                assert illegal[i] == 0 and reward[i] == g["reward"][t], (name, i, t)
                assert (done[i] != 0) == bool(g["ends_limit"][t] if limit else g["ends_plain"][t]), (name, i, t)
                assert np.array_equal(sn[i], g["snap"][t]), (name, i, t)
        score2, _, _ = eng.final_scores()
        score2 = score2.cpu().numpy()
        for i, g in enumerate(gs):
            assert list(score2[i]) == list(g["score2"]), (name, i)
            fo = eng.observe(torch.full((N,), 0, dtype=torch.int8)).cpu().numpy()
            assert np.array_equal(fo[i].view(np.uint32), g["final_obs"][0].view(np.uint32)), (name, i)


@pytest.mark.parametrize("P,limit", [(2, False), (3, False), (4, False), (2, True)])
def test_step_mask_observe_against_oracle_on_random_trajectories(pkg, oracle, P, limit):
    """Batched mask -> choose -> step on counter-based decks vs the oracle stepping the same games (explicit lists)."""
    N, seed, rng = 192, 4242, np.random.default_rng(P * 10 + limit)
    eng = pkg.BatchedSplendor(N, P, seed=seed, first_game=1000, limit_rounds=limit)
    orcs = [oracle.init(P, *oracle.synth_game(seed, 1000 + e, P)) for e in range(N)]
    alive = np.ones(N, bool)
    for t in range(400):
        if not alive.any():
            break
        mask = eng.legal_mask().cpu().numpy()
        obs = eng.observe().cpu().numpy()
        act = np.full(N, -1, np.int16)
        for e in np.nonzero(alive)[0]:
            legal = oracle.legal(orcs[e])
            assert np.array_equal(unpack_mask(mask[e]), legal), (e, t)
            assert np.array_equal(obs[e].view(np.uint32), oracle.observe(orcs[e], orcs[e].cur).view(np.uint32)), (e, t)
            act[e] = rng.choice(legal)
        reward, done, illegal = (x.cpu().numpy() for x in eng.step(torch.from_numpy(act)))
        sn = snaps(eng, pkg)
        for e in np.nonzero(alive)[0]:
            assert illegal[e] == 0 and reward[e] == oracle.step(orcs[e], int(act[e])), (e, t)
            assert done[e] == oracle.game_ends(orcs[e], limit), (e, t)
            assert np.array_equal(sn[e], orcs[e].snapshot()), (e, t)
            if done[e]:
                alive[e] = False
    assert not alive.any()
    score2, outcome, kind = (x.cpu().numpy() for x in eng.final_scores())
    for e in range(N):
        assert list(score2[e]) == [oracle.cal_score2(orcs[e], i) for i in range(P)]


@pytest.mark.parametrize("P,limit,G", [(2, False, 4096), (3, False, 1024), (4, False, 2048), (2, True, 1024), (4, True, 512)])
def test_fused_playout_traces_match_oracle(pkg, oracle, P, limit, G):
    """Config 2 of BASELINE.json: whole random playouts on the GPU; every game's action trace, length, ending, final
    state and calScore identical to the oracle's playout of the same (seed, game)."""
    seed, first = 1234, 10_000
    out = pkg.playout_games(G, P, seed=seed, first_game=first, limit_rounds=limit, max_steps=1000, trace_len=420, want_final_state=True)
    trace = out["trace"].cpu().numpy()
    steps, kind, score2 = out["steps"].cpu().numpy(), out["end_kind"].cpu().numpy(), out["score2"].cpu().numpy()
    sn = pkg.state.snapshots(out["final_state"].cpu().numpy().view(np.uint32), P)
    for g in range(G):
        st, fin, tr, k = oracle.playout(seed, first + g, P, limit, 1000)
        assert (steps[g], kind[g]) == (st, k), g
        assert np.array_equal(trace[:st, g], tr[:st]) and (trace[st:, g] == -1).all(), g
        assert list(score2[g]) == [oracle.cal_score2(fin, i) for i in range(P)], g
        assert np.array_equal(sn[g], fin.snapshot()), g
    _, _, ctr = oracle.playout_many(seed, first, G, P, limit, 1000, nthreads=os.cpu_count() or 1)
    assert np.array_equal(out["counters"].cpu().numpy(), ctr)


def test_playout_full_size_counters_match_oracle(pkg, oracle):
    """BASELINE config 2 at full size (65,536 two-player games): all 16 aggregate counters (wins/ties per seat, endings,
    total steps, score sum, nobles) equal the oracle's over the same games; per-game lengths and scores equal too."""
    G, seed = 65_536, 1234
    out = pkg.playout_games(G, 2, seed=seed)
    sc, st, ctr = oracle.playout_many(seed, 0, G, 2, False, 1000, nthreads=os.cpu_count() or 1)
    assert np.array_equal(out["counters"].cpu().numpy(), ctr)
    assert np.array_equal(out["steps"].cpu().numpy().astype(np.int32), st)
    assert np.array_equal(out["score2"].cpu().numpy(), sc)


def test_explicit_deck_playout_equals_counter_based(pkg):
    """spl_inject + spl_playout_from on the materialised deck orders == the fused counter-based playout."""
    G, P, seed = 3000, 2, 99
    eng = pkg.BatchedSplendor(G, P, seed=seed, first_game=0)
    lists, nobles = eng.reset(want_lists=True)
    eng.inject(lists.cpu().numpy(), nobles.cpu().numpy())
    a = eng.playout(max_steps=1000, trace_len=300)
    b = pkg.playout_games(G, P, seed=seed, first_game=0, trace_len=300)
    for key in ("score2", "steps", "end_kind", "counters", "trace"):
        assert torch.equal(a[key], b[key]), key


def test_properties_at_full_size_4p(pkg):
    """Config 3 (4 players, 2^20 envs): size-independent invariants + determinism + sharding invariance."""
    G, P = 1 << 20, 4
    a = pkg.playout_games(G, P, seed=5)
    c = pkg.counters_dict(a["counters"])
    steps = a["steps"].to(torch.int64)
    assert c["games"] == G and c["steps"] == int(steps.sum())
    assert c["end_score"] + c["end_deadlock"] + c["end_roundcap"] + c["end_truncated"] == G
    assert c["end_roundcap"] == 0 and c["end_truncated"] == 0
    assert c["score2_sum"] == int(a["score2"].to(torch.int64).sum())
    wins = sum(c[f"wins{i}"] for i in range(4))
    assert wins <= G and wins + sum(c[f"ties{i}"] for i in range(4)) >= G
    kind = a["end_kind"]
    best = a["score2"].max(dim=1).values
    assert bool(((best >= 30) | (kind != 1)).all())           # a score ending means someone reached 15 points
    assert bool((steps[kind == 1] % P == 0).all())            # ... and only at the end of a round (seat 0 to move)
    b = pkg.playout_games(G, P, seed=5)                        # determinism
    assert torch.equal(a["score2"], b["score2"]) and torch.equal(a["steps"], b["steps"])
    h0 = pkg.playout_games(G // 2, P, seed=5, first_game=0)    # sharding by game id gives identical per-game results
    h1 = pkg.playout_games(G // 2, P, seed=5, first_game=G // 2)
    assert torch.equal(torch.cat([h0["score2"], h1["score2"]]), a["score2"])
    assert torch.equal(h0["counters"] + h1["counters"], a["counters"])


def sampled_blocks(total, n_blocks, block):
    """n_blocks contiguous runs of `block` game indices spread evenly over [0, total) (the first and the last run included)."""
    return [int(round(i * (total - block) / (n_blocks - 1))) for i in range(n_blocks)]


def test_full_size_4p_sample_is_exact(pkg, oracle):
    """Config 3 at full size (2^20 four-player games in ONE launch): 16,384 games sampled in 64 runs across the launch have
    exactly the oracle's length, scores and ending (about 2.7 M oracle steps)."""
    G, P, seed, first = 1 << 20, 4, 5, 0
    a = pkg.playout_games(G, P, seed=seed, first_game=first)
    steps, score2, kind = a["steps"].cpu().numpy(), a["score2"].cpu().numpy(), a["end_kind"].cpu().numpy()
    checked = 0
    for lo in sampled_blocks(G, 64, 256):
        sc, st, ctr = oracle.playout_many(seed, first + lo, 256, P, False, 1000, nthreads=os.cpu_count() or 1)
        assert np.array_equal(steps[lo:lo + 256].astype(np.int32), st) and np.array_equal(score2[lo:lo + 256], sc), lo
        assert int((kind[lo:lo + 256] == 1).sum()) == int(ctr[10]) and int((kind[lo:lo + 256] == 2).sum()) == int(ctr[11]), lo
        checked += 256
    assert checked == 16_384


def test_sweep_2p_2_24_games_counters_and_sample_are_exact(pkg, oracle):
    """Config 5 at full size: the 2^24-game two-player sweep of bench.py (same seed and game ids, 16 launches of 2^20):
    all 16 counters equal tests/golden/sweep_counters.json - computed by the ORACLE over all 2^24 games
    (tools/sweep_oracle.py) - and 32,768 games sampled in 128 runs have exactly the oracle's length, scores and ending."""
    import json

    from conftest import GOLDEN

    G, P, seed, first, per = 1 << 24, 2, 1234, 1 << 44, 1 << 20
    ctr = torch.zeros(16, dtype=torch.int64, device="cuda")
    steps = torch.empty(G, dtype=torch.int16, device="cuda")
    score2 = torch.empty((G, P), dtype=torch.int8, device="cuda")
    kind = torch.empty(G, dtype=torch.uint8, device="cuda")
    lib = pkg._native.lib()
    stream = torch.cuda.current_stream().cuda_stream
    for lo in range(0, G, per):
        rc = lib.spl_playout(seed, first + lo, per, P, 0, 1000, score2[lo:].data_ptr(), steps[lo:].data_ptr(), kind[lo:].data_ptr(),
                             ctr.data_ptr(), None, 0, None, stream)
        assert rc == 0
    c = pkg.counters_dict(ctr)
    assert c["games"] == G and c["steps"] == int(steps.to(torch.int64).sum()) and c["score2_sum"] == int(score2.to(torch.int64).sum())
    want = json.load(open(os.path.join(GOLDEN, "sweep_counters.json")))
    assert (want["seed"], want["first_game"], want["games"]) == (seed, first, G)
    assert c == want["counters"]
    steps, score2, kind = steps.cpu().numpy(), score2.cpu().numpy(), kind.cpu().numpy()
    for lo in sampled_blocks(G, 128, 256):
        sc, st, octr = oracle.playout_many(seed, first + lo, 256, P, False, 1000, nthreads=os.cpu_count() or 1)
        assert np.array_equal(steps[lo:lo + 256].astype(np.int32), st) and np.array_equal(score2[lo:lo + 256], sc), lo
        assert int((kind[lo:lo + 256] == 1).sum()) == int(octr[10]), lo


def test_env_step_at_full_size_sample_is_exact(pkg, oracle):
    """Config 4 at full size: spl_env_step on 262,144 two-player envs for 32 calls (reset call + 31 learner moves chosen on
    the device from the returned masks); 4,096 envs sampled in 32 runs are followed by the oracle's emulation of
    SplendorEnv (learner move, Philox opponents, 100-round cap) and compared EVERY call: reward, terminated, legal set,
    observation bits - and the full state after the last call."""
    N, P, seed = 262_144, 2, 4321
    rng = np.random.default_rng(9)
    learner_np = rng.integers(0, P, N).astype(np.int8)
    learner = torch.from_numpy(learner_np).cuda()
    eng = pkg.BatchedSplendor(N, P, seed=seed, first_game=0, limit_rounds=True)
    sample = np.concatenate([np.arange(lo, lo + 128) for lo in sampled_blocks(N, 32, 128)])
    orcs = {int(e): oracle.init(P, *oracle.synth_game(seed, int(e), P)) for e in sample}

    def opponents(e):
        s = orcs[e]
        while s.cur != learner_np[e] and not oracle.game_ends(s, True):
            legal = oracle.legal(s)
            t = sum(s.pl[i].turns for i in range(P))
            r = int(oracle.philox(seed, e, t, 0)[0])
            oracle.step(s, int(legal[(r * len(legal)) >> 32]))

    gen = torch.Generator(device="cuda").manual_seed(1)
    act = torch.full((N,), -1, dtype=torch.int16, device="cuda")
    sample_t = torch.from_numpy(sample).cuda()
    for call in range(32):
        obs, reward, term, illegal, mask = eng.env_step(learner, act)
        assert int(illegal.sum()) == 0
        a_np = act[sample_t].cpu().numpy()
        o_np, r_np, t_np, m_np = obs[sample_t].cpu().numpy(), reward[sample_t].cpu().numpy(), term[sample_t].cpu().numpy(), mask[sample_t].cpu().numpy()
        for j, e in enumerate(sample):
            e = int(e)
            if a_np[j] >= 0:
                assert r_np[j] == oracle.step(orcs[e], int(a_np[j])), (e, call)
            opponents(e)
            assert t_np[j] == oracle.game_ends(orcs[e], True), (e, call)
            assert np.array_equal(unpack_mask(m_np[j]), oracle.legal(orcs[e], int(learner_np[e]))), (e, call)
            assert np.array_equal(o_np[j].view(np.uint32), oracle.observe(orcs[e], int(learner_np[e])).view(np.uint32)), (e, call)
        # next actions, on the device: a uniformly random legal action where the game goes on, "skip" where it is over
        bits = ((mask.unsqueeze(-1) >> torch.arange(32, device="cuda", dtype=torch.int32)) & 1).reshape(N, -1)[:, :3510].to(torch.float32)
        pick = torch.multinomial(bits, 1, generator=gen).squeeze(1).to(torch.int16)
        act = torch.where(term != 0, torch.full_like(pick, -1), pick)
    sn = pkg.state.snapshots(eng.state[:, sample_t].cpu().numpy().view(np.uint32), P)
    for j, e in enumerate(sample):
        assert np.array_equal(sn[j], orcs[int(e)].snapshot()), e


def test_illegal_actions_are_flagged_and_ignored(pkg, oracle):
    N, P = 256, 2
    eng = pkg.BatchedSplendor(N, P, seed=3)
    before = eng.state.clone()
    mask = eng.legal_mask().cpu().numpy()
    rng = np.random.default_rng(0)
    act = rng.integers(0, 3510, N).astype(np.int16)
    act[:4] = [3510, 32000, 0, 5]
    legal = np.array([a < 3510 and a in set(unpack_mask(mask[e]).tolist()) for e, a in enumerate(act)])
    reward, done, illegal = (x.cpu().numpy() for x in eng.step(torch.from_numpy(act)))
    assert np.array_equal(illegal == 0, legal)
    after = eng.state.clone()
    after[1, :, 3] &= ~(1 << 26)
    bad = torch.from_numpy(~legal).to(after.device)
    assert torch.equal(after[:, bad], before[:, bad])
    assert (reward[~legal] == 0).all()


@pytest.mark.parametrize("N", [1, 33, 257])
def test_ragged_batches_take_the_plain_store_path(pkg, oracle, N):
    """Batch sizes whose last 32-env group has an ODD row count: the mask rows of that group cannot leave through the
    16-byte bulk copy and use the plain coalesced store; rows, observations and the fused env call must still match."""
    P, seed = 2, 77
    eng = pkg.BatchedSplendor(N, P, seed=seed, first_game=5, limit_rounds=True)
    orcs = [oracle.init(P, *oracle.synth_game(seed, 5 + e, P)) for e in range(N)]
    rng = np.random.default_rng(N)
    guard = torch.full((N + 1, 110), 0x5A5A5A5A, dtype=torch.int32, device=eng.device)  # one spare row: no write past the end
    for t in range(14):
        mask = eng.legal_mask(out=guard[:N]).cpu().numpy()
        assert (guard[N] == 0x5A5A5A5A).all()
        obs = eng.observe().cpu().numpy()
        act = np.zeros(N, np.int16)
        for e in range(N):
            legal = oracle.legal(orcs[e])
            assert np.array_equal(unpack_mask(mask[e]), legal), (e, t)
            assert np.array_equal(obs[e].view(np.uint32), oracle.observe(orcs[e], orcs[e].cur).view(np.uint32)), (e, t)
            act[e] = rng.choice(legal)
            oracle.step(orcs[e], int(act[e]))
        eng.step(torch.from_numpy(act))
    # the fused call (self-play form: acts for the seat to move, returns the next mover's rows)
    act = np.array([rng.choice(oracle.legal(o)) for o in orcs], np.int16)
    obs, rew, term, ill, mask = eng.env_step(None, torch.from_numpy(act), self_play=True)
    mask, obs = mask.cpu().numpy(), obs.cpu().numpy()
    for e in range(N):
        oracle.step(orcs[e], int(act[e]))
        assert np.array_equal(unpack_mask(mask[e]), oracle.legal(orcs[e])), e
        assert np.array_equal(obs[e].view(np.uint32), oracle.observe(orcs[e], orcs[e].cur).view(np.uint32)), e


@pytest.mark.parametrize("N,P", [(70, 2), (131, 3), (96, 4)])
def test_observation_buffer_that_is_only_float_aligned(pkg, oracle, N, P):
    """The observation kernel writes a warp's 32 rows as 128-bit stores when the caller's buffer is 16-byte aligned and falls
    back to scalar stores when it is not: both must give the same bits, and neither may write outside the N rows."""
    eng = pkg.BatchedSplendor(N, P, seed=5, first_game=900, limit_rounds=True)
    learner = torch.zeros(N, dtype=torch.int8, device=eng.device)
    act = torch.full((N,), -1, dtype=torch.int16, device=eng.device)
    gen = torch.Generator(device=eng.device).manual_seed(3)
    for _ in range(12):  # mid-game states
        _, _, _, _, mask = eng.env_step(learner, act)
        act = torch.multinomial(pkg.env.unpack_mask(mask, torch.float32), 1, generator=gen).squeeze(1).to(torch.int16)
    want = eng.observe()
    flat = torch.full((N * 265 + 8,), float("nan"), dtype=torch.float32, device=eng.device)
    for off in (1, 2, 3, 4):
        flat.fill_(float("nan"))
        view = flat[off:off + N * 265].view(N, 265)
        assert view.data_ptr() % 16 == (4 * off) % 16
        got = eng.observe(out=view)
        assert torch.equal(got.view(torch.int32), want.view(torch.int32)), off
        assert torch.isnan(flat[:off]).all() and torch.isnan(flat[off + N * 265:]).all(), off
    seats = torch.from_numpy(np.random.default_rng(1).integers(0, P, N).astype(np.int8)).cuda()
    assert torch.equal(eng.observe(seat=seats, out=flat[1:1 + N * 265].view(N, 265)).view(torch.int32), eng.observe(seat=seats).view(torch.int32))


@pytest.mark.parametrize("P", [2, 3, 4])
def test_env_step_matches_oracle_emulation_of_splendor_env(pkg, oracle, P):
    """spl_env_step vs SplendorEnv.step semantics (splendor_env.py:127-231) emulated with the oracle: learner move,
    Philox-random opponents (one per other seat) until the learner's seat, then the learner's mask + observation,
    100-round cap."""
    N, seed = (160 if P == 2 else 64), 31
    rng = np.random.default_rng(1)
    learner = rng.integers(0, P, N).astype(np.int8)
    eng = pkg.BatchedSplendor(N, P, seed=seed, first_game=0, limit_rounds=True)
    orcs = [oracle.init(P, *oracle.synth_game(seed, e, P)) for e in range(N)]

    def simulate_opponents(e):
        s = orcs[e]
        while s.cur != learner[e] and not oracle.game_ends(s, True):
            legal = oracle.legal(s)
            t = sum(s.pl[i].turns for i in range(P))
            r = int(oracle.philox(seed, e, t, 0)[0])
            oracle.step(s, int(legal[(r * len(legal)) >> 32]))

    act = np.full(N, -1, np.int16)  # reset step: opponents play up to the learner's first turn
    alive = np.ones(N, bool)
    for it in range(130):
        obs, reward, term, illegal, mask = (x.cpu().numpy() for x in eng.env_step(torch.from_numpy(learner), torch.from_numpy(act)))
        for e in np.nonzero(alive)[0]:
            if act[e] >= 0:
                assert reward[e] == oracle.step(orcs[e], int(act[e])), (e, it)
            simulate_opponents(e)
            assert illegal[e] == 0 and term[e] == oracle.game_ends(orcs[e], True), (e, it)
            assert np.array_equal(unpack_mask(mask[e]), oracle.legal(orcs[e], int(learner[e]))), (e, it)
            assert np.array_equal(obs[e].view(np.uint32), oracle.observe(orcs[e], int(learner[e])).view(np.uint32)), (e, it)
            if term[e]:
                alive[e] = False
                act[e] = -1
            else:
                act[e] = rng.choice(oracle.legal(orcs[e]))
        sn = snaps(eng, pkg)
        for e in range(N):
            assert np.array_equal(sn[e], orcs[e].snapshot()), (e, it)
        if not alive.any():
            break
    assert not alive.any()


def test_host_buffer_abi_call(pkg, oracle):
    """spl_playout_host: host deck orders in, host results out (the end-to-end path bench.py times)."""
    import ctypes as C

    N, P, seed = 2048, 2, 8
    lists = np.stack([oracle.synth_game(seed, g, P)[0] for g in range(N)])
    nobles = np.full((N, 5), 255, np.uint8)
    for g in range(N):
        nobles[g, :P + 1] = oracle.synth_game(seed, g, P)[1]
    score2, steps, kind = np.zeros((N, P), np.int8), np.zeros(N, np.int16), np.zeros(N, np.uint8)
    ctr = np.zeros(16, np.int64)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = pkg._native.lib().spl_playout_host(p(lists), p(nobles), seed, 0, N, P, 0, 1000, p(score2), p(steps), p(kind), p(ctr), None)
    assert rc == 0, pkg._native.lib().spl_error_string(rc)
    sc, st, c = oracle.playout_many(seed, 0, N, P, False, 1000, nthreads=os.cpu_count() or 1)
    assert np.array_equal(score2, sc) and np.array_equal(steps.astype(np.int32), st) and np.array_equal(ctr, c)


def test_host_buffer_abi_call_with_a_ragged_batch(pkg, oracle):
    """spl_playout_host with a batch size that is no multiple of the warp or pool size, and a non-zero first game id."""
    import ctypes as C

    N, P, seed, first = 20_011, 2, 77, 5_000
    eng = pkg.BatchedSplendor(N, P, seed=seed, first_game=first)
    lists_d, nobles_d = eng.reset(want_lists=True)
    lists, nobles = lists_d.cpu().numpy(), nobles_d.cpu().numpy()
    score2, steps, kind = np.zeros((N, P), np.int8), np.zeros(N, np.int16), np.zeros(N, np.uint8)
    ctr = np.zeros(16, np.int64)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = pkg._native.lib().spl_playout_host(p(lists), p(nobles), seed, first, N, P, 0, 1000, p(score2), p(steps), p(kind), p(ctr), None)
    assert rc == 0, pkg._native.lib().spl_error_string(rc)
    sc, st, c = oracle.playout_many(seed, first, N, P, False, 1000, nthreads=os.cpu_count() or 1)
    assert np.array_equal(score2, sc) and np.array_equal(steps.astype(np.int32), st) and np.array_equal(ctr, c)


def test_async_host_buffer_calls_pipeline_on_two_slots(pkg, oracle):
    """spl_playout_host_async: six batches in flight pairwise on two slots / two streams with pinned buffers; every
    batch's outputs and counters equal the oracle's (and so the synchronous call's)."""
    N, P, seed = 18_000, 2, 31
    lib = pkg._native.lib()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    batches = []
    for b in range(6):
        eng = pkg.BatchedSplendor(N, P, seed=seed, first_game=b * N)
        lists_d, nobles_d = eng.reset(want_lists=True)
        batches.append((lists_d.cpu().pin_memory(), nobles_d.cpu().pin_memory()))
    torch.cuda.synchronize()
    outs = [(torch.zeros((N, P), dtype=torch.int8).pin_memory(), torch.zeros(N, dtype=torch.int16).pin_memory(),
             torch.zeros(N, dtype=torch.uint8).pin_memory(), torch.zeros(16, dtype=torch.int64).pin_memory()) for _ in range(6)]
    for b in range(6):
        slot = b & 1
        lists, nobles = batches[b]
        sc, st, kd, ctr = outs[b]
        rc = lib.spl_playout_host_async(lists.data_ptr(), nobles.data_ptr(), seed, b * N, N, P, 0, 1000, sc.data_ptr(), st.data_ptr(),
                                        kd.data_ptr(), ctr.data_ptr(), slot, streams[slot].cuda_stream)
        assert rc == 0, lib.spl_error_string(rc)
    for s in streams:
        s.synchronize()
    for b in range(6):
        sc, st, kd, ctr = outs[b]
        osc, ost, octr = oracle.playout_many(seed, b * N, N, P, False, 1000, nthreads=os.cpu_count() or 1)
        assert np.array_equal(sc.numpy(), osc) and np.array_equal(st.numpy().astype(np.int32), ost), b
        assert np.array_equal(ctr.numpy(), octr), b
    assert lib.spl_playout_host_async(batches[0][0].data_ptr(), batches[0][1].data_ptr(), seed, 0, N, P, 0, 1000, None, None, None, None,
                                      2, None) == pkg._native.SPL_E_BADARG  # only slots 0 and 1 exist


def test_results_do_not_depend_on_the_launch_geometry(pkg, oracle):
    """A game's trajectory is a function of (seed, game id) only: the same 5,000 three-player games played in one launch,
    and cut into launches of very different sizes (so that they land in other pools, other warps and other rounds of the
    block-pool playout kernel), give identical per-game outputs, traces and final states."""
    G, P, seed, first = 5000, 3, 5, 123
    whole = pkg.playout_games(G, P, seed=seed, first_game=first, trace_len=64, want_final_state=True)
    cuts = [0, 1, 33, 700, 2748, 2749, 4999, 5000]
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        part = pkg.playout_games(hi - lo, P, seed=seed, first_game=first + lo, trace_len=64, want_final_state=True)
        for key in ("score2", "steps", "end_kind"):
            assert torch.equal(part[key], whole[key][lo:hi]), (key, lo, hi)
        assert torch.equal(part["trace"], whole["trace"][:, lo:hi]), (lo, hi)
        assert torch.equal(part["final_state"], whole["final_state"][:, lo:hi]), (lo, hi)
    assert int(whole["counters"][0]) == G


def test_row_kernels_do_not_depend_on_the_launch_shape(pkg):
    """The mask and observation kernels are persistent: the host picks (warps per block) x (blocks per SM) per batch size, and
    SPL_MASK_SHAPE / SPL_OBS_SHAPE override it.  Every shape - one warp per block, blocks that split a tile row between
    neighbours differently, more blocks than tiles - must write the same bits."""
    N, P = 5003, 3  # ragged: the last tile has 11 rows
    eng = pkg.BatchedSplendor(N, P, seed=21, first_game=77, limit_rounds=True)
    learner = torch.zeros(N, dtype=torch.int8, device=eng.device)
    act = torch.full((N,), -1, dtype=torch.int16, device=eng.device)
    gen = torch.Generator(device=eng.device).manual_seed(5)
    for _ in range(10):
        _, _, _, _, mask = eng.env_step(learner, act)
        act = torch.multinomial(pkg.env.unpack_mask(mask, torch.float32), 1, generator=gen).squeeze(1).to(torch.int16)
    want_mask, want_obs = eng.legal_mask().clone(), eng.observe().clone()
    try:
        for shape in ("1x1", "1x4", "2x3", "3x7", "4x5", "5x4", "7x2", "13x1", "16x1"):
            os.environ["SPL_MASK_SHAPE"] = os.environ["SPL_OBS_SHAPE"] = shape
            assert torch.equal(eng.legal_mask(), want_mask), shape
            assert torch.equal(eng.observe().view(torch.int32), want_obs.view(torch.int32)), shape
    finally:
        os.environ.pop("SPL_MASK_SHAPE", None)
        os.environ.pop("SPL_OBS_SHAPE", None)


def test_two_devices_from_one_process(pkg, oracle):
    """The library keeps its staging memory, events and SM count PER DEVICE: one process drives two GPUs through the
    host-buffer calls and the device-pointer calls alternately (skipped on a one-GPU box)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two visible GPUs")
    N, P, seed = 6000, 2, 12
    lib = pkg._native.lib()
    want = oracle.playout_many(seed, 0, N, P, False, 1000, nthreads=os.cpu_count() or 1)
    lists = np.stack([oracle.synth_game(seed, g, P)[0] for g in range(N)])
    nobles = np.full((N, 5), 255, np.uint8)
    for g in range(N):
        nobles[g, :P + 1] = oracle.synth_game(seed, g, P)[1]
    import ctypes as C

    p = lambda a: a.ctypes.data_as(C.c_void_p)
    for rep in range(2):
        for d in (0, 1):
            with torch.cuda.device(d):
                score2, steps, kind, ctr = np.zeros((N, P), np.int8), np.zeros(N, np.int16), np.zeros(N, np.uint8), np.zeros(16, np.int64)
                assert lib.spl_playout_host(p(lists), p(nobles), seed, 0, N, P, 0, 1000, p(score2), p(steps), p(kind), p(ctr), None) == 0
                assert np.array_equal(score2, want[0]) and np.array_equal(steps.astype(np.int32), want[1]) and np.array_equal(ctr, want[2]), (rep, d)
                out = pkg.playout_games(N, P, seed=seed, device=f"cuda:{d}")
                assert np.array_equal(out["score2"].cpu().numpy(), want[0]), (rep, d)
                eng = pkg.BatchedSplendor(256, P, seed=seed, device=f"cuda:{d}")
                assert eng.legal_mask().device.index == d and eng.observe().device.index == d


@pytest.mark.parametrize("P", [2, 3, 4])
def test_fuzzed_states_on_gpu(pkg, oracle, P):
    """Arbitrary hand-built states (tests/fuzz_states.py: 10-gem holders, drained banks, empty slots/decks, 7-card cap,
    3 reserved, several nobles earned, pass-only, 100-turn cap) packed into the device layout: mask, observation for
    every seat, gameEnds/calScore, and the successor of one sampled legal action per env - all vs the oracle."""
    from fuzz_states import random_state

    N, rng = 3000, np.random.default_rng(100 + P)
    orcs, fields, lists = [], [], []
    for _ in range(N):
        o, f, l = random_state(rng, oracle, P)
        orcs.append(o); fields.append(f); lists.append(l)
    merged = {k: np.concatenate([f[k] for f in fields]) for k in fields[0]}
    for limit in (False, True):
        eng = pkg.BatchedSplendor(N, P, limit_rounds=limit)
        eng.state.copy_(torch.from_numpy(pkg.state.pack(merged, P).view(np.int32)))
        eng.decks = torch.from_numpy(np.stack(lists)).to(eng.device)
        assert np.array_equal(snaps(eng, pkg), np.stack([o.snapshot() for o in orcs]))
        mask = eng.legal_mask().cpu().numpy()
        score2, outcome, kind = (x.cpu().numpy() for x in eng.final_scores())
        obs = [eng.observe(torch.full((N,), s, dtype=torch.int8)).cpu().numpy() for s in range(P)]
        act = np.zeros(N, np.int16)
        for e, o in enumerate(orcs):
            legal = oracle.legal(o)
            assert np.array_equal(unpack_mask(mask[e]), legal), (P, e)
            assert kind[e] == oracle.game_ends(o, limit) and list(score2[e]) == [oracle.cal_score2(o, i) for i in range(P)], (P, e)
            for s in range(P):
                assert np.array_equal(obs[s][e].view(np.uint32), oracle.observe(o, s).view(np.uint32)), (P, e, s)
            act[e] = rng.choice(legal)
        reward, done, illegal = (x.cpu().numpy() for x in eng.step(torch.from_numpy(act)))
        sn = snaps(eng, pkg)
        for e, o in enumerate(orcs):
            o2 = o.copy()
            assert illegal[e] == 0 and reward[e] == oracle.step(o2, int(act[e])), (P, e)
            assert done[e] == oracle.game_ends(o2, limit) and np.array_equal(sn[e], o2.snapshot()), (P, e)


@pytest.mark.parametrize("P", [2, 4])
def test_one_and_two_ply_expansion(pkg, oracle, P):
    """spl_count_legal + spl_expand: every successor of mid-game states (counter-based decks), then the successors'
    successors, against the oracle stepping copies - the batched form of minimax's successor/predecessor walk."""
    N, seed = 96, 11
    eng = pkg.BatchedSplendor(N, P, seed=seed, first_game=50)
    orcs = [oracle.init(P, *oracle.synth_game(seed, 50 + e, P)) for e in range(N)]
    rng = np.random.default_rng(P)
    for _ in range(24):  # advance to mid-game with random legal actions
        act = np.array([rng.choice(oracle.legal(o)) if not oracle.game_ends(o, False) else -1 for o in orcs], np.int16)
        eng.step(torch.from_numpy(act))
        for o, a in zip(orcs, act):
            if a >= 0:
                oracle.step(o, int(a))
    child, action, reward, parent, offsets = eng.expand()
    action, reward, parent, offsets = (x.cpu().numpy() for x in (action, reward, parent, offsets))
    counts = eng.count_legal().cpu().numpy()
    sn = pkg.state.snapshots(child.state_numpy(), P)
    kids = []
    for e, o in enumerate(orcs):
        legal = oracle.legal(o)
        assert counts[e] == len(legal) == offsets[e + 1] - offsets[e]
        for j, a in enumerate(legal):
            m = offsets[e] + j
            c = o.copy()
            assert action[m] == a and parent[m] == e and reward[m] == oracle.step(c, int(a)), (e, j)
            assert np.array_equal(sn[m], c.snapshot()), (e, j)
            kids.append(c)
    # second ply on the children (their lazily dealt cards must continue the parent's deck order)
    grand, action2, _, parent2, offsets2 = child.expand()
    action2, parent2, offsets2 = (x.cpu().numpy() for x in (action2, parent2, offsets2))
    sn2 = pkg.state.snapshots(grand.state_numpy(), P)
    for m in rng.choice(len(kids), size=min(300, len(kids)), replace=False):
        legal = oracle.legal(kids[m])
        assert offsets2[m + 1] - offsets2[m] == len(legal)
        for j in (0, len(legal) - 1):
            c = kids[m].copy()
            oracle.step(c, int(legal[j]))
            assert action2[offsets2[m] + j] == legal[j] and np.array_equal(sn2[offsets2[m] + j], c.snapshot()), (m, j)


def test_truncated_playouts_resume_to_the_same_results(pkg, oracle):
    """max_steps cuts games short (end kind 4, states saved); continuing from the saved states with spl_playout_from gives
    exactly the results of the uninterrupted playout - and the truncated prefix equals the oracle's."""
    G, P, seed = 2048, 2, 21
    full = pkg.playout_games(G, P, seed=seed, trace_len=300)
    cut = pkg.playout_games(G, P, seed=seed, max_steps=40, trace_len=300, want_final_state=True)
    kinds = cut["end_kind"].cpu().numpy()
    assert (kinds[full["steps"].cpu().numpy() > 40] == 4).all() and (cut["steps"].cpu().numpy() <= 40).all()
    c = pkg.counters_dict(cut["counters"])
    assert c["end_truncated"] == int((kinds == 4).sum()) > 0 and c["steps"] == int(cut["steps"].to(torch.int64).sum())
    for g in range(0, G, 64):
        st, fin, tr, k = oracle.playout(seed, g, P, False, 40)
        assert (int(cut["steps"][g]), int(kinds[g])) == (st, k)
        assert np.array_equal(cut["trace"][:st, g].cpu().numpy(), tr[:st])
    eng = pkg.BatchedSplendor(G, P, seed=seed)
    eng.state.copy_(cut["final_state"])
    rest = eng.playout(max_steps=1000, trace_len=300)
    # `steps` of a playout from handed-in states counts the moves of THIS call (not the turns already on the clock)
    assert torch.equal(rest["score2"], full["score2"]) and torch.equal(rest["steps"] + cut["steps"], full["steps"])
    assert torch.equal(rest["end_kind"], full["end_kind"])
    joined = torch.where(cut["trace"] >= 0, cut["trace"], rest["trace"])
    assert torch.equal(joined, full["trace"])


@pytest.mark.parametrize("P", [2, 3])
def test_self_play_env_step_matches_oracle(pkg, oracle, P):
    """SPL_F_SELF_PLAY: the caller moves every seat; reward / end kind / next mover's mask and observation vs the oracle."""
    from splendor_ai_b200.env import SplendorVecEnv

    N, seed = 96, 17
    env = SplendorVecEnv(N, P, seed=seed, self_play=True)
    obs, info = env.reset()
    orcs = [oracle.init(P, *oracle.synth_game(seed, e, P)) for e in range(N)]
    rng = np.random.default_rng(P)
    episode = np.zeros(N, np.int64)
    for it in range(140):
        o, m, who = obs.cpu().numpy(), info["legal_mask"].cpu().numpy(), info["my_id"].cpu().numpy()
        act = np.zeros(N, np.int16)
        for e in range(N):
            s = orcs[e]
            assert who[e] == s.cur
            legal = oracle.legal(s)
            assert np.array_equal(unpack_mask(m[e]), legal), (e, it)
            assert np.array_equal(o[e].view(np.uint32), oracle.observe(s, s.cur).view(np.uint32)), (e, it)
            act[e] = rng.choice(legal)
        was_done = env._done.cpu().numpy()
        obs, reward, term, _, info = env.step(torch.from_numpy(act))
        reward, kind = reward.cpu().numpy(), info["end_kind"].cpu().numpy()
        for e in range(N):
            if was_done[e]:  # auto-reset: a fresh game, the action was ignored
                episode[e] += 1
                orcs[e] = oracle.init(P, *oracle.synth_game(seed, e + episode[e] * N, P))
                assert reward[e] == 0
            else:
                assert reward[e] == oracle.step(orcs[e], int(act[e])), (e, it)
            assert kind[e] == oracle.game_ends(orcs[e], True), (e, it)
    assert (episode >= 1).mean() > 0.7  # one call = one move here, so a few long games are still running


def test_partial_reset_of_an_injected_batch_keeps_both_deck_forms(pkg, oracle):
    """inject() then reset(where=...): the reset envs draw from their counter-based Philox decks, the untouched ones keep
    popping their explicit lists (the deck form is chosen per env from META_EXPLICIT, engine.cuh deal())."""
    N, P, seed = 64, 2, 91
    rng = np.random.default_rng(5)
    lists = np.stack([np.concatenate([rng.permutation(40), 40 + rng.permutation(30), 70 + rng.permutation(20)]) for _ in range(N)]).astype(np.uint8)
    nobles = np.stack([rng.permutation(10)[:P + 1] for _ in range(N)]).astype(np.uint8)
    eng = pkg.BatchedSplendor(N, P, seed=seed)
    eng.inject(lists, nobles)
    orcs = [oracle.init(P, lists[e], nobles[e]) for e in range(N)]

    def play(rounds):
        for _ in range(rounds):
            act = np.full(N, -1, np.int16)
            for e in range(N):
                if not oracle.game_ends(orcs[e], False):
                    act[e] = rng.choice(oracle.legal(orcs[e]))
                    oracle.step(orcs[e], int(act[e]))
            _, _, illegal = eng.step(torch.from_numpy(act))
            assert int(illegal.sum()) == 0
            assert np.array_equal(snaps(eng, pkg), np.stack([o.snapshot() for o in orcs]))

    play(25)
    where = torch.from_numpy((np.arange(N) % 2 == 1).astype(np.uint8))
    ids = torch.arange(1000, 1000 + N)
    eng.reset(where=where, game_ids=ids)
    for e in range(1, N, 2):
        orcs[e] = oracle.init(P, *oracle.synth_game(seed, 1000 + e, P))
    assert np.array_equal(snaps(eng, pkg), np.stack([o.snapshot() for o in orcs]))
    play(60)  # reserve / buy moves of both kinds of env deal from their decks

"""BatchedRolloutBuffer (splendor-ai_b200/rollout.py) against the reference's return / advantage formulas
(agents/our_agents/ppo/common.py:11-57), restated inline so the test also runs without the reference tree."""
import numpy as np
import torch


def reference_returns(rewards, gamma):  # common.py:24-30, un-normalised
    out, running = [], 0.0
    for r in reversed(rewards):
        running = r + running * gamma
        out.insert(0, running)
    return out


def test_returns_restart_at_episode_boundaries_and_match_the_reference_formula():
    from splendor_ai_b200.rollout import BatchedRolloutBuffer

    T, N, gamma = 40, 7, 0.99
    rng = np.random.default_rng(0)
    buf = BatchedRolloutBuffer(T, N, device=torch.device("cpu"), dtype=torch.float64)
    rewards = rng.integers(0, 4, (T, N)).astype(np.float64)
    dones = rng.random((T, N)) < 0.08
    for t in range(T):
        buf.remember(torch.zeros(N, 265, dtype=torch.float64), torch.full((N,), t, dtype=torch.int16),
                     torch.full((N, 110), t, dtype=torch.int32), torch.zeros(N, dtype=torch.float64),
                     torch.full((N, 1), 0.5, dtype=torch.float64), torch.from_numpy(rewards[t]), torch.from_numpy(dones[t]))
    assert buf.full and buf.index == T
    got = buf.discounted_returns(gamma).numpy()
    for e in range(N):
        start = 0
        for t in range(T):
            if dones[t, e] or t == T - 1:  # an episode (or the unfinished tail) covers start..t
                want = reference_returns(list(rewards[start:t + 1, e]), gamma)
                assert np.allclose(got[start:t + 1, e], want, rtol=0, atol=1e-12), (e, start, t)
                start = t + 1
    adv, ret = buf.calculate_gae(gamma)
    assert abs(float(ret.mean())) < 1e-9 and abs(float(ret.std()) - 1) < 1e-6   # normalised like common.py:33-36
    raw = torch.from_numpy(got)
    want_ret = (raw - raw.mean()) / (raw.std() + 1e-8)
    want_adv = want_ret - 0.5
    want_adv = (want_adv - want_adv.mean()) / (want_adv.std() + 1e-8)
    assert torch.allclose(ret, want_ret) and torch.allclose(adv, want_adv)
    states, actions, masks, logp, advantages, returns, d = buf.unpack(gamma)
    assert states.shape == (T * N, 265) and masks.shape == (T * N, 110) and advantages.shape == (T * N,)
    assert int(actions[3 * N + 2]) == 3 and bool(d[5 * N + 1]) == bool(dones[5, 1])
    buf.clear()
    assert buf.index == 0 and not buf.full


def test_packed_masks_expand_to_the_dense_reference_form():
    from splendor_ai_b200.rollout import BatchedRolloutBuffer

    buf = BatchedRolloutBuffer(2, 3, device=torch.device("cpu"))
    words = torch.zeros((3, 110), dtype=torch.int32)
    words[0, 0] = 0b1000001          # actions 0 and 6
    words[1, 109] = 1 << 21          # action 3509
    words[2, 35] = -2**31            # bit 31 of word 35 = action 1151
    z = torch.zeros(3)
    buf.remember(torch.zeros(3, 265), torch.zeros(3, dtype=torch.int16), words, z, z, z, torch.zeros(3, dtype=torch.bool))
    dense = buf.action_masks(0, 3, torch.float64)
    assert dense.shape == (3, 3510) and dense.dtype == torch.float64
    assert dense[0].nonzero().flatten().tolist() == [0, 6]
    assert dense[1].nonzero().flatten().tolist() == [3509] and dense[2].nonzero().flatten().tolist() == [1151]

"""spl_masked_sample (masked softmax + categorical sample + log-prob from bit-packed masks) against plain PyTorch.
Float kernel: log-probabilities are compared with a float64 torch reference, tolerance 1e-5 (stated in the header)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HUGE_NEG = -1e8  # agents/our_agents/ppo/constants.py:6


def setup(N, seed=0, p_legal=0.01):
    g = torch.Generator(device="cuda").manual_seed(seed)
    logits = torch.randn((N, 3510), generator=g, device="cuda") * 3.0
    dense = torch.rand((N, 3510), generator=g, device="cuda") < p_legal
    dense[:, 7] |= ~dense.any(dim=1)  # every row needs one legal action
    bits = torch.zeros((N, 110 * 32), dtype=torch.int64, device="cuda")
    bits[:, :3510] = dense.to(torch.int64)
    weights = (1 << torch.arange(32, device="cuda", dtype=torch.int64))
    words = (bits.view(N, 110, 32) * weights).sum(-1)
    words = torch.where(words >= 2**31, words - 2**32, words).to(torch.int32)
    return logits, dense, words


def test_samples_are_legal_and_log_probs_match_torch():
    from splendor_ai_b200.sampling import masked_sample

    N = 20_000
    logits, dense, words = setup(N)
    action, logp = masked_sample(logits, words, seed=5, step=3)
    a = action.to(torch.int64)
    assert bool(dense[torch.arange(N, device="cuda"), a].all())
    ref = torch.log_softmax(torch.where(dense, logits.double(), torch.full_like(logits, HUGE_NEG, dtype=torch.float64)), dim=1)
    want = ref[torch.arange(N, device="cuda"), a]
    assert float((logp.double() - want).abs().max()) < 1e-5
    a2, _ = masked_sample(logits, words, seed=5, step=3)            # deterministic in (seed, id, step)
    a3, _ = masked_sample(logits, words, seed=5, step=4)
    assert torch.equal(action, a2) and not torch.equal(action, a3)
    g, lg = masked_sample(logits, words, seed=5, step=3, greedy=True)
    best = torch.where(dense, logits, torch.full_like(logits, -float("inf"))).argmax(dim=1)
    assert torch.equal(g.to(torch.int64), best)
    assert float((lg.double() - ref[torch.arange(N, device="cuda"), best]).abs().max()) < 1e-5


def test_single_legal_action_and_strided_rows():
    from splendor_ai_b200.sampling import masked_sample

    N = 512
    logits, dense, words = setup(N, seed=1, p_legal=0.0)  # only action 7 legal everywhere
    wide = torch.zeros((N, 3584), device="cuda")
    wide[:, :3510] = logits
    action, logp = masked_sample(wide, words, seed=1, step=0)
    assert bool((action == 7).all()) and float(logp.abs().max()) == 0.0


def test_inverse_cdf_reproduced_on_the_host(oracle):
    """The kernel's fixed-point inverse CDF restated with numpy + the oracle's Philox: same draw, same index (the float32
    exp may differ by an ulp between libraries, so a handful of boundary rows are tolerated)."""
    from splendor_ai_b200.sampling import masked_sample

    N, seed, step, first = 3000, 99, 12, 1000
    logits, dense, words = setup(N, seed=2, p_legal=0.02)
    action, _ = masked_sample(logits, words, seed=seed, step=step, first_id=first)
    L, D = logits.cpu().numpy(), dense.cpu().numpy()
    same = 0
    for r in range(N):
        l = np.where(D[r], L[r], -np.inf).astype(np.float32)
        p = np.exp((l - l.max()).astype(np.float32)).astype(np.float32)
        w = np.floor(p.astype(np.float64) * 16777216.0).astype(np.uint64)
        total = int(w.sum())
        u = int(oracle.philox(seed, first + r, step, 6)[0])
        target = (total * u) >> 32
        idx = int(np.searchsorted(np.cumsum(w), target, side="right"))
        same += int(idx == int(action[r]))
    assert same >= N - 3, same


def test_empirical_distribution_matches_softmax():
    from splendor_ai_b200.sampling import masked_sample

    n_rep = 400_000
    logits, dense, words = setup(1, seed=3, p_legal=0.004)
    legal = torch.nonzero(dense[0]).squeeze(1)
    probs = torch.softmax(logits[0, legal].double(), dim=0)
    rep_logits = logits.expand(n_rep, -1).contiguous()
    rep_words = words.expand(n_rep, -1).contiguous()
    action, _ = masked_sample(rep_logits, rep_words, seed=8, step=0)
    counts = torch.bincount(action.to(torch.int64), minlength=3510)[legal].double()
    assert int(counts.sum()) == n_rep
    sigma = torch.sqrt(n_rep * probs * (1 - probs)).clamp_min(1.0)
    assert float(((counts - n_rep * probs).abs() / sigma).max()) < 5.0

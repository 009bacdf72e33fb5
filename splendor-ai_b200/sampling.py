"""Action choice under the legal mask for batched rollouts (reference: ``agents/our_agents/ppo/network.py:113-115`` masks the
actor's logits with ``HUGE_NEG`` and soft-maxes them, ``training.py:110-116`` samples a ``Categorical`` and keeps its
log-probability).  ``masked_sample`` does the three steps in one kernel (``spl_masked_sample``) straight from the
bit-packed masks the env kernels emit - no dense [N, 3510] mask, no extra passes over the logits."""
from __future__ import annotations

import torch

from . import _native as nat

NUM_ACTIONS = 3510


def masked_sample(logits: torch.Tensor, mask_bits: torch.Tensor, seed: int, step: int, ids: torch.Tensor | None = None,
                  first_id: int = 0, greedy: bool = False, want_log_prob: bool = True):
    """logits float32 [N, >=3510] (row-major, last dim contiguous), mask_bits int32 [N, 110] -> (action int16 [N],
    log_prob float32 [N] | None).  The draw of row e is a function of (seed, ids[e] or first_id + e, step) only."""
    if logits.dtype != torch.float32 or logits.stride(-1) != 1 or logits.shape[-1] < NUM_ACTIONS:
        raise ValueError("logits must be float32 [N, >=3510] with a contiguous last dimension")
    if not logits.is_cuda:
        raise nat.SplendorNativeError("masked_sample needs CUDA tensors: there is no CPU fallback")
    N, dev = logits.shape[0], logits.device
    m = mask_bits.contiguous()
    assert m.shape == (N, 110) and m.dtype == torch.int32 and m.device == dev
    action = torch.empty(N, dtype=torch.int16, device=dev)
    log_prob = torch.empty(N, dtype=torch.float32, device=dev) if want_log_prob else None
    i = None if ids is None else ids.to(device=dev, dtype=torch.int64).contiguous()
    with torch.cuda.device(dev):
        nat.check(nat.lib().spl_masked_sample(logits.data_ptr(), logits.stride(0), m.data_ptr(), int(seed), int(first_id),
                                              None if i is None else i.data_ptr(), int(step), int(greedy), action.data_ptr(),
                                              None if log_prob is None else log_prob.data_ptr(), N,
                                              torch.cuda.current_stream(dev).cuda_stream), "spl_masked_sample")
    return action, log_prob

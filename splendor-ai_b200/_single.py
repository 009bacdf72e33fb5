"""One-env device helper behind the single-game drop-in classes (game_rule.py): uploads a packed state, runs the CUDA
kernels through the C ABI, downloads the result.  Per-call latency is dominated by the launch + the two tiny copies;
the batched classes (engine.py, env.py) are the fast path."""
from __future__ import annotations

import numpy as np
import torch

from . import _native as nat
from .state import state_words

_bufs = {}


def _dev():
    if not torch.cuda.is_available():
        raise nat.SplendorNativeError("no CUDA device: the B200 engine has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _buf(P):
    key = (P, torch.cuda.current_device())
    if key not in _bufs:
        d = _dev()
        _bufs[key] = dict(
            state=torch.zeros((state_words(P) // 4, 1, 4), dtype=torch.int32, device=d),
            decks=torch.zeros((1, 90), dtype=torch.uint8, device=d),
            nobles=torch.zeros((1, 5), dtype=torch.uint8, device=d),
            mask=torch.zeros((1, 110), dtype=torch.int32, device=d),
            action=torch.zeros(1, dtype=torch.int16, device=d),
            out=torch.zeros(8, dtype=torch.uint8, device=d),
            score=torch.zeros((1, P), dtype=torch.int8, device=d),
        )
    return _bufs[key]


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _up(b, words, seat=None):
    w = np.array(words, dtype=np.uint32, copy=True)
    if seat is not None:
        w[7] = (w[7] & ~np.uint32(3 << 20)) | np.uint32(seat << 20)
    b["state"].copy_(torch.from_numpy(w.view(np.int32).reshape(-1, 1, 4)))


def inject(P, deck_lists, nobles) -> np.ndarray:
    b = _buf(P)
    b["decks"].copy_(torch.from_numpy(np.asarray(deck_lists, np.uint8).reshape(1, 90)))
    n = np.full((1, 5), 0xFF, np.uint8)
    n[0, :len(nobles)] = nobles
    b["nobles"].copy_(torch.from_numpy(n))
    nat.check(nat.lib().spl_inject(b["state"].data_ptr(), b["decks"].data_ptr(), b["nobles"].data_ptr(), 1, P, _stream()), "spl_inject")
    return b["state"].cpu().numpy().view(np.uint32).reshape(-1).copy()


def legal_indices(words, P, seat) -> np.ndarray:
    b = _buf(P)
    _up(b, words, seat)
    nat.check(nat.lib().spl_legal_mask(b["state"].data_ptr(), b["mask"].data_ptr(), 1, P, _stream()), "spl_legal_mask")
    bits = np.unpackbits(b["mask"].cpu().numpy().view(np.uint8).reshape(-1), bitorder="little")
    return np.nonzero(bits)[0]


def step(words, deck_lists, P, seat, index):
    """Applies ALL_ACTIONS[index] for `seat` IN PLACE on words; returns (reward, illegal)."""
    b = _buf(P)
    _up(b, words, seat)
    b["decks"].copy_(torch.from_numpy(np.asarray(deck_lists, np.uint8).reshape(1, 90)))
    b["action"].fill_(int(index))
    o = b["out"]
    nat.check(nat.lib().spl_step(b["state"].data_ptr(), b["decks"].data_ptr(), 0, 0, None, b["action"].data_ptr(), o[0:1].data_ptr(),
                                 o[1:2].data_ptr(), o[2:3].data_ptr(), 1, P, 0, _stream()), "spl_step")
    res = o.cpu().numpy()
    if not res[2]:
        words[:] = b["state"].cpu().numpy().view(np.uint32).reshape(-1)
    return int(res[0].astype(np.int8)), bool(res[2])


def game_ends(words, P, current_agent_index, limit_rounds=False) -> int:
    b = _buf(P)
    _up(b, words, current_agent_index)
    o = b["out"]
    nat.check(nat.lib().spl_final_scores(b["state"].data_ptr(), None, None, o[0:1].data_ptr(), 1, P,
                                         nat.F_LIMIT_ROUNDS if limit_rounds else 0, _stream()), "spl_final_scores")
    return int(o[0].item())


def cal_score2(words, P):
    b = _buf(P)
    _up(b, words)
    nat.check(nat.lib().spl_final_scores(b["state"].data_ptr(), b["score"].data_ptr(), None, None, 1, P, 0, _stream()), "spl_final_scores")
    return b["score"].cpu().numpy()[0].astype(int)


def observe(words, P, seat) -> np.ndarray:
    b = _buf(P)
    _up(b, words, seat)
    obs = torch.empty((1, 265), dtype=torch.float32, device=b["state"].device)
    nat.check(nat.lib().spl_observe(b["state"].data_ptr(), None, obs.data_ptr(), 1, P, _stream()), "spl_observe")
    return obs.cpu().numpy()[0]


def undo(words, P, seat, index, card, noble, returned):
    """generatePredecessor IN PLACE on words (spl_undo)."""
    b = _buf(P)
    _up(b, words)
    d = b["state"].device
    t = lambda v, dt: torch.tensor(v, dtype=dt, device=d)
    seat_t, act_t, card_t, noble_t = t([seat], torch.int8), t([index], torch.int16), t([card], torch.uint8), t([noble], torch.uint8)
    ret_t = t([list(returned)], torch.uint8)
    nat.check(nat.lib().spl_undo(b["state"].data_ptr(), seat_t.data_ptr(), act_t.data_ptr(), card_t.data_ptr(), noble_t.data_ptr(),
                                 ret_t.data_ptr(), 1, P, _stream()), "spl_undo")
    words[:] = b["state"].cpu().numpy().view(np.uint32).reshape(-1)


def minimax2(words, deck_lists, P, seat):
    """MiniMaxAgent.SelectAction for `seat` on one state: (best ALL_ACTIONS index, best value, actions, values)."""
    if P != 2:
        raise ValueError("the minimax agent supports only a game of 2 players")
    b = _buf(P)
    _up(b, words, seat)
    b["decks"].copy_(torch.from_numpy(np.asarray(deck_lists, np.uint8).reshape(1, 90)))
    d = b["state"].device
    counts = torch.zeros(1, dtype=torch.int32, device=d)
    nat.check(nat.lib().spl_count_legal(b["state"].data_ptr(), counts.data_ptr(), 1, P, _stream()), "spl_count_legal")
    M = int(counts.item())
    offsets = torch.tensor([0, M], dtype=torch.int64, device=d)
    child_value = torch.empty(M, dtype=torch.float64, device=d)
    child_action = torch.empty(M, dtype=torch.int16, device=d)
    best_action = torch.empty(1, dtype=torch.int16, device=d)
    best_value = torch.empty(1, dtype=torch.float64, device=d)
    nat.check(nat.lib().spl_minimax2(b["state"].data_ptr(), b["decks"].data_ptr(), 0, 0, None, offsets.data_ptr(), child_value.data_ptr(),
                                     child_action.data_ptr(), best_action.data_ptr(), best_value.data_ptr(), 1, M, P, _stream()),
              "spl_minimax2")
    return int(best_action.item()), float(best_value.item()), child_action.cpu().numpy(), child_value.cpu().numpy()


def minimax_eval(words, P, seat) -> float:
    b = _buf(P)
    _up(b, words, seat)
    out = torch.empty(1, dtype=torch.float64, device=b["state"].device)
    nat.check(nat.lib().spl_minimax_eval(b["state"].data_ptr(), None, out.data_ptr(), 1, P, _stream()), "spl_minimax_eval")
    return float(out.item())

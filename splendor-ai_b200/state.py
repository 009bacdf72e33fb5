"""Packed HBM state layout (include/splendor_b200.h) <-> plain numpy field arrays.

One env = W = 8 + 4*P little-endian u32 words = C = W/4 chunks of 16 bytes; a batch is stored chunk-major,
``words[c, e, 0:4]`` = chunk c of env e (uint4[C][N] on the device).  The fields restate the reference's object graph
(splendor_model.py:52-138): board gems, dealt grid, decks, noble list, and per agent gems / bought cards / reserved
list / score / nobles / passed / turn count.

Gem and card counts are six 5-bit fields per word in the reference's COLOURS order black, red, yellow, green, blue,
white (splendor_utils.py:124-131); the unpacked arrays use colour ids black0 red1 green2 blue3 white4 yellow5.
"""
from __future__ import annotations

import numpy as np

NONE = 0xFF
FIELD_OF_COLOUR_ID = (0, 1, 3, 4, 5, 2)  # colour id -> 5-bit field
TIER_BASE = (0, 40, 70)
TIER_SIZE = (40, 30, 20)
META_EXPLICIT = 1 << 24
META_ILLEGAL = 1 << 26


def state_words(P: int) -> int:
    if P not in (2, 3, 4):
        raise ValueError("number of agents must be 2, 3 or 4")
    return 8 + 4 * P


def state_shape(N: int, P: int):
    """Shape of the uint32 array/tensor holding N packed envs."""
    return (state_words(P) // 4, N, 4)


def _g6(word: np.ndarray) -> np.ndarray:
    """[..., 6] counts by colour id from a g6 word."""
    return np.stack([(word >> (5 * f)) & 31 for f in FIELD_OF_COLOUR_ID], axis=-1).astype(np.uint8)


def _bytes4(word: np.ndarray) -> np.ndarray:
    return np.stack([(word >> (8 * i)) & 0xFF for i in range(4)], axis=-1).astype(np.uint8)


def _popcount(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint64)
    c = np.zeros(x.shape, np.uint8)
    for i in range(40):
        c += ((x >> np.uint64(i)) & np.uint64(1)).astype(np.uint8)
    return c


def unpack(words: np.ndarray, P: int) -> dict:
    """Decode packed states (uint32 [C, N, 4]) into field arrays (all leading dimension N)."""
    words = np.asarray(words, dtype=np.uint32)
    C, N, _ = words.shape
    assert C == state_words(P) // 4
    b0, b1 = words[0], words[1]
    meta = b1[:, 3]
    explicit = (meta & META_EXPLICIT) != 0
    t1_mask = b1[:, 0].astype(np.uint64) | (((b1[:, 2] >> 20) & 0xFF).astype(np.uint64) << np.uint64(32))
    t2_mask = (b1[:, 1] & 0x3FFFFFFF).astype(np.uint64)
    t3_mask = (b1[:, 2] & 0xFFFFF).astype(np.uint64)
    lazy_n = np.stack([_popcount(t1_mask), _popcount(t2_mask), _popcount(t3_mask)], axis=-1)
    expl_n = _bytes4(b1[:, 0])[:, :3]
    nobles = np.stack([(meta >> (4 * k)) & 15 for k in range(5)], axis=-1).astype(np.uint8)
    nobles[nobles == 15] = NONE
    out = dict(
        P=P,
        bank=_g6(b0[:, 0]),
        dealt=np.concatenate([_bytes4(b0[:, 1]), _bytes4(b0[:, 2]), _bytes4(b0[:, 3])], axis=-1),
        nobles=nobles,
        n_nobles=(nobles != NONE).sum(-1).astype(np.uint8),
        deck_n=np.where(explicit[:, None], expl_n, lazy_n).astype(np.uint8),
        deck_masks=np.stack([t1_mask, t2_mask, t3_mask], axis=-1),
        explicit=explicit,
        illegal_flag=(meta & META_ILLEGAL) != 0,
        cur=((meta >> 20) & 3).astype(np.uint8),
        gems=np.zeros((N, P, 6), np.uint8), cards=np.zeros((N, P, 5), np.uint8), resv=np.zeros((N, P, 3), np.uint8),
        n_resv=np.zeros((N, P), np.uint8), score=np.zeros((N, P), np.uint8), passed=np.zeros((N, P), np.uint8),
        turns=np.zeros((N, P), np.uint16), nobles_mask=np.zeros((N, P), np.uint16), n_nobles_won=np.zeros((N, P), np.uint8),
    )
    for p in range(P):
        w = words[2 + p]
        out["gems"][:, p] = _g6(w[:, 0])
        out["cards"][:, p] = _g6(w[:, 1])[:, :5]
        rb = _bytes4(w[:, 2])
        out["resv"][:, p] = rb[:, :3]
        out["n_resv"][:, p] = (rb[:, :3] != NONE).sum(-1)
        out["score"][:, p] = rb[:, 3]
        out["turns"][:, p] = w[:, 3] & 0xFFFF
        out["nobles_mask"][:, p] = (w[:, 3] >> 16) & 0x3FF
        out["n_nobles_won"][:, p] = _popcount((w[:, 3] >> 16) & 0x3FF)
        out["passed"][:, p] = (w[:, 3] >> 26) & 1
    return out


def snapshots(words: np.ndarray, P: int) -> np.ndarray:
    """The canonical 108-byte field vector per env used by the golden files and parity tests ([N, 108] uint8):
    bank6 | dealt12 | nobles5 | n_nobles | deck_n3 | cur | 4 x (gems6 cards5 resv3 n_resv score passed n_nobles turns_lo turns_hi)."""
    u = unpack(words, P)
    N = u["bank"].shape[0]
    out = np.zeros((N, 108), np.uint8)
    out[:, 0:6] = u["bank"]
    out[:, 6:18] = u["dealt"]
    out[:, 18:23] = u["nobles"]
    out[:, 23] = u["n_nobles"]
    out[:, 24:27] = u["deck_n"]
    out[:, 27] = u["cur"]
    for p in range(P):
        o = 28 + 20 * p
        out[:, o:o + 6] = u["gems"][:, p]
        out[:, o + 6:o + 11] = u["cards"][:, p]
        out[:, o + 11:o + 14] = u["resv"][:, p]
        out[:, o + 14] = u["n_resv"][:, p]
        out[:, o + 15] = u["score"][:, p]
        out[:, o + 16] = u["passed"][:, p]
        out[:, o + 17] = u["n_nobles_won"][:, p]
        out[:, o + 18] = u["turns"][:, p] & 0xFF
        out[:, o + 19] = u["turns"][:, p] >> 8
    return out


def remaining_deck_sets(deck_masks_row) -> list:
    """Counter-based decks: per tier, the sorted card ids still in the deck (from one env's deck_masks)."""
    return [[TIER_BASE[t] + i for i in range(TIER_SIZE[t]) if (int(deck_masks_row[t]) >> i) & 1] for t in range(3)]


def _g6_word(counts: np.ndarray) -> np.ndarray:
    """[..., 6] (or 5) counts by colour id -> g6 word."""
    w = np.zeros(counts.shape[:-1], np.uint32)
    for cid in range(counts.shape[-1]):
        w |= counts[..., cid].astype(np.uint32) << np.uint32(5 * FIELD_OF_COLOUR_ID[cid])
    return w


def pack(fields: dict, P: int, deck_lists: np.ndarray | None = None):
    """Inverse of `unpack` for EXPLICIT-deck states: field arrays (leading dimension N; keys bank, dealt, nobles, deck_n,
    cur, gems, cards, resv, score, passed, turns, nobles_mask) -> packed words [C, N, 4] uint32.  The deck lists ([N,90],
    reference list order) are not part of the words; they are passed to the kernels separately."""
    N = np.asarray(fields["bank"]).shape[0]
    words = np.zeros((state_words(P) // 4, N, 4), np.uint32)
    words[0, :, 0] = _g6_word(np.asarray(fields["bank"]))
    dealt = np.asarray(fields["dealt"], np.uint32)
    for t in range(3):
        words[0, :, 1 + t] = dealt[:, 4 * t] | dealt[:, 4 * t + 1] << 8 | dealt[:, 4 * t + 2] << 16 | dealt[:, 4 * t + 3] << 24
    dn = np.asarray(fields["deck_n"], np.uint32)
    words[1, :, 0] = dn[:, 0] | dn[:, 1] << 8 | dn[:, 2] << 16
    nob = np.asarray(fields["nobles"], np.uint32).copy()
    nob[nob == NONE] = 15
    meta = np.zeros(N, np.uint32)
    for k in range(5):
        meta |= (nob[:, k] & 15) << np.uint32(4 * k)
    meta |= np.asarray(fields["cur"], np.uint32) << 20
    meta |= np.uint32((P - 2) << 22) | np.uint32(META_EXPLICIT)
    words[1, :, 3] = meta
    for p in range(P):
        words[2 + p, :, 0] = _g6_word(np.asarray(fields["gems"])[:, p])
        words[2 + p, :, 1] = _g6_word(np.asarray(fields["cards"])[:, p])
        rv = np.asarray(fields["resv"], np.uint32)[:, p]
        words[2 + p, :, 2] = rv[:, 0] | rv[:, 1] << 8 | rv[:, 2] << 16 | np.asarray(fields["score"], np.uint32)[:, p] << 24
        words[2 + p, :, 3] = (np.asarray(fields["turns"], np.uint32)[:, p] & 0xFFFF) | \
            (np.asarray(fields["nobles_mask"], np.uint32)[:, p] & 0x3FF) << 16 | (np.asarray(fields["passed"], np.uint32)[:, p] & 1) << 26
    return words


def unpack_one(words, P: int) -> dict:
    """Scalar fast path of `unpack` for ONE env given as a flat sequence of 8+4P words (plain Python ints: the
    single-game classes decode a state after every move, and numpy on 1-element arrays is mostly overhead)."""
    w = [int(x) for x in words]
    g6 = lambda x: [(x >> (5 * f)) & 31 for f in FIELD_OF_COLOUR_ID]
    b4 = lambda x: [(x >> (8 * i)) & 0xFF for i in range(4)]
    meta = w[7]
    explicit = bool(meta & META_EXPLICIT)
    if explicit:
        deck_n = b4(w[4])[:3]
    else:
        deck_n = [bin(w[4] | ((w[6] >> 20) & 0xFF) << 32).count("1"), bin(w[5] & 0x3FFFFFFF).count("1"), bin(w[6] & 0xFFFFF).count("1")]
    nobles = [(meta >> (4 * k)) & 15 for k in range(5)]
    out = dict(P=P, bank=g6(w[0]), dealt=b4(w[1]) + b4(w[2]) + b4(w[3]), nobles=[NONE if n == 15 else n for n in nobles],
               deck_n=deck_n, explicit=explicit, cur=(meta >> 20) & 3, gems=[], cards=[], resv=[], score=[], passed=[], turns=[],
               nobles_mask=[])
    for p in range(P):
        q = w[8 + 4 * p:12 + 4 * p]
        rb = b4(q[2])
        out["gems"].append(g6(q[0]))
        out["cards"].append(g6(q[1])[:5])
        out["resv"].append(rb[:3])
        out["score"].append(rb[3])
        out["turns"].append(q[3] & 0xFFFF)
        out["nobles_mask"].append((q[3] >> 16) & 0x3FF)
        out["passed"].append((q[3] >> 26) & 1)
    return out

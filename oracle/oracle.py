"""ctypes front-end of the CPU oracle (oracle/splendor_oracle.c).  TEST INFRASTRUCTURE - see that file's header.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

NONE = 0xFF
NUM_ACTIONS = 3510
MASK_WORDS = 110
OBS_SIZE = 265
T_PASS, T_RESERVE, T_SAME, T_DIFF, T_BUY_RESERVE, T_BUY_AVAILABLE = range(6)
END_NONE, END_SCORE, END_DEADLOCK, END_ROUNDCAP, END_TRUNCATED = range(5)


class Player(C.Structure):
    _fields_ = [
        ("gems", C.c_uint8 * 6), ("cards", C.c_uint8 * 5), ("resv", C.c_uint8 * 3), ("n_resv", C.c_uint8),
        ("score", C.c_uint8), ("passed", C.c_uint8), ("n_nobles", C.c_uint8), ("pad", C.c_uint8),
        ("nobles_mask", C.c_uint16), ("turns", C.c_uint16),
    ]


class State(C.Structure):
    _fields_ = [
        ("P", C.c_uint8), ("cur", C.c_uint8), ("bank", C.c_uint8 * 6), ("dealt", C.c_uint8 * 12),
        ("nobles", C.c_uint8 * 5), ("n_nobles", C.c_uint8), ("deck", (C.c_uint8 * 40) * 3),
        ("deck_n", C.c_uint8 * 3), ("pad", C.c_uint8), ("pl", Player * 4),
    ]

    def copy(self) -> "State":
        other = State()
        C.memmove(C.byref(other), C.byref(self), C.sizeof(State))
        return other

    def snapshot(self) -> np.ndarray:
        """Canonical 108-byte field vector used by the golden files and the GPU parity tests."""
        out = np.zeros(108, np.uint8)
        out[0:6] = self.bank
        out[6:18] = self.dealt
        out[18:23] = self.nobles
        out[23] = self.n_nobles
        out[24:27] = self.deck_n
        out[27] = self.cur
        for i in range(4):
            p, o = self.pl[i], 28 + 20 * i
            if i >= self.P:
                continue
            out[o:o + 6] = p.gems
            out[o + 6:o + 11] = p.cards
            out[o + 11:o + 14] = p.resv
            out[o + 14] = p.n_resv
            out[o + 15] = p.score
            out[o + 16] = p.passed
            out[o + 17] = p.n_nobles
            out[o + 18] = p.turns & 0xFF
            out[o + 19] = p.turns >> 8
        return out

    def remaining_deck_sets(self):
        """Per tier, the sorted card ids still in the deck."""
        return [sorted(self.deck[t][i] for i in range(self.deck_n[t])) for t in range(3)]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "splendor_oracle.c")
    hdr = os.path.join(_HERE, "..", "include", "spl_tables.h")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(
        os.path.getmtime(src), os.path.getmtime(hdr) if os.path.exists(hdr) else 0
    ):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        u8p, u16p, u32p, i8p, i16p, i32p, i64p, f32p = (C.POINTER(t) for t in (
            C.c_uint8, C.c_uint16, C.c_uint32, C.c_int8, C.c_int16, C.c_int32, C.c_int64, C.c_float))
        sp = C.POINTER(State)
        L.orc_sizeof_state.restype = C.c_int
        assert L.orc_sizeof_state() == C.sizeof(State), (L.orc_sizeof_state(), C.sizeof(State))
        L.orc_init.argtypes = [sp, C.c_int, u8p, u8p]
        L.orc_legal.argtypes = [sp, C.c_int, u16p]; L.orc_legal.restype = C.c_int
        L.orc_legal_mask.argtypes = [sp, C.c_int, u32p]; L.orc_legal_mask.restype = C.c_int
        L.orc_step.argtypes = [sp, C.c_int]; L.orc_step.restype = C.c_int
        L.orc_game_ends.argtypes = [sp, C.c_int]; L.orc_game_ends.restype = C.c_int
        L.orc_cal_score2.argtypes = [sp, C.c_int]; L.orc_cal_score2.restype = C.c_int
        L.orc_observe.argtypes = [sp, C.c_int, f32p]
        L.orc_philox.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, u32p]
        L.orc_synth_game.argtypes = [C.c_uint64, C.c_uint64, C.c_int, u8p, u8p]
        L.orc_playout.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int, sp, i16p, C.POINTER(C.c_int)]
        L.orc_playout.restype = C.c_int
        L.orc_playout_many.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, i8p, i32p, i64p, C.c_int]
        L.orc_undo.argtypes = [sp, C.c_int, C.c_int, C.c_int, C.c_int, u8p]
        L.orc_undo.restype = C.c_int
        L.orc_describe_action.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), u8p, u8p]
        L.orc_describe_action.restype = C.c_int
        L.orc_minimax_eval.argtypes = [sp, C.c_int]; L.orc_minimax_eval.restype = C.c_double
        L.orc_minimax2.argtypes = [sp, u16p, C.POINTER(C.c_double)]; L.orc_minimax2.restype = C.c_int
        _lib = L
    return _lib


def _ptr(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def init(P: int, deck_lists, nobles) -> State:
    """deck_lists: 90 card ids (tier lists concatenated, reference list order); nobles: P+1 noble ids."""
    s = State()
    d = np.ascontiguousarray(deck_lists, np.uint8)
    n = np.ascontiguousarray(list(nobles) + [NONE] * (5 - len(nobles)), np.uint8)
    assert d.size == 90
    lib().orc_init(C.byref(s), P, _ptr(d, C.c_uint8), _ptr(n, C.c_uint8))
    return s


def legal(s: State, seat=None) -> np.ndarray:
    out = np.zeros(1024, np.uint16)
    n = lib().orc_legal(C.byref(s), s.cur if seat is None else seat, _ptr(out, C.c_uint16))
    return out[:n].copy()


def legal_mask(s: State, seat=None) -> np.ndarray:
    out = np.zeros(MASK_WORDS, np.uint32)
    lib().orc_legal_mask(C.byref(s), s.cur if seat is None else seat, _ptr(out, C.c_uint32))
    return out


def step(s: State, idx: int) -> int:
    return lib().orc_step(C.byref(s), int(idx))


def game_ends(s: State, limit_rounds: bool = False) -> int:
    return lib().orc_game_ends(C.byref(s), int(limit_rounds))


def cal_score2(s: State, agent: int) -> int:
    return lib().orc_cal_score2(C.byref(s), agent)


def observe(s: State, seat: int) -> np.ndarray:
    out = np.zeros(OBS_SIZE, np.float32)
    lib().orc_observe(C.byref(s), seat, _ptr(out, C.c_float))
    return out


def philox(seed: int, game: int, index: int, stream: int) -> np.ndarray:
    out = np.zeros(4, np.uint32)
    lib().orc_philox(seed, game, index, stream, _ptr(out, C.c_uint32))
    return out


def synth_game(seed: int, game: int, P: int):
    d = np.zeros(90, np.uint8)
    n = np.zeros(5, np.uint8)
    lib().orc_synth_game(seed, game, P, _ptr(d, C.c_uint8), _ptr(n, C.c_uint8))
    return d, n[:P + 1].copy()


def playout(seed: int, game: int, P: int, limit_rounds: bool = False, max_steps: int = 1000, want_trace: bool = True):
    fin = State()
    trace = np.full(max_steps, -1, np.int16) if want_trace else None
    kind = C.c_int(0)
    steps = lib().orc_playout(seed, game, P, int(limit_rounds), max_steps, C.byref(fin),
                              _ptr(trace, C.c_int16) if want_trace else None, C.byref(kind))
    return steps, fin, trace, kind.value


def playout_many(seed: int, first_game: int, n: int, P: int, limit_rounds: bool = False, max_steps: int = 1000,
                 nthreads: int = 1, want_outputs: bool = True):
    score2 = np.zeros((n, P), np.int8) if want_outputs else None
    steps = np.zeros(n, np.int32) if want_outputs else None
    ctr = np.zeros(16, np.int64)
    lib().orc_playout_many(seed, first_game, n, P, int(limit_rounds), max_steps,
                           _ptr(score2, C.c_int8) if want_outputs else None,
                           _ptr(steps, C.c_int32) if want_outputs else None, _ptr(ctr, C.c_int64), nthreads)
    return score2, steps, ctr


def describe_action(idx: int):
    t, p, n = C.c_int(), C.c_int(), C.c_int()
    col, ret = np.zeros(6, np.uint8), np.zeros(6, np.uint8)
    rc = lib().orc_describe_action(idx, C.byref(t), C.byref(p), C.byref(n), _ptr(col, C.c_uint8), _ptr(ret, C.c_uint8))
    if rc:
        raise ValueError(idx)
    return t.value, p.value, n.value, col, ret


def undo(s: State, seat: int, idx: int, card: int, noble: int, returned) -> int:
    r = np.ascontiguousarray(returned, np.uint8)
    return lib().orc_undo(C.byref(s), seat, int(idx), int(card), int(noble), _ptr(r, C.c_uint8))


def minimax_eval(s: State, me: int) -> float:
    """MiniMaxAgent._evaluation_function (minmax.py:90-136) for agent `me`, float64."""
    return float(lib().orc_minimax_eval(C.byref(s), me))


def minimax2(s: State):
    """Depth-2 minimax values of every legal action of the seat to move: (actions uint16 [n], values float64 [n])."""
    acts = np.zeros(1024, np.uint16)
    vals = np.zeros(1024, np.float64)
    n = lib().orc_minimax2(C.byref(s), _ptr(acts, C.c_uint16), _ptr(vals, C.c_double))
    return acts[:n].copy(), vals[:n].copy()

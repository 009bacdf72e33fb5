"""ctypes front-end of tests/hostsim/hostsim.cpp (TEST INFRASTRUCTURE: the device rule code compiled for the host)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
_LIB = os.path.join(_HERE, "libhostsim.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        srcs = [os.path.join(_HERE, "hostsim.cpp"), os.path.join(_ROOT, "splendor-ai_b200", "csrc", "engine.cuh"),
                os.path.join(_ROOT, "include", "spl_tables.h")]
        if not os.path.exists(_LIB) or os.path.getmtime(_LIB) < max(map(os.path.getmtime, srcs)):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-DSPL_DEBUG", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", _LIB, srcs[0]])  # SPL_CHECK asserts on
        _lib = C.CDLL(_LIB)
        _lib.sim_reset.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p]
        _lib.sim_inject.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        _lib.sim_legal_select.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        _lib.sim_legal_member.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        _lib.sim_mask.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        _lib.sim_step.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_uint32,
                                  C.POINTER(C.c_int), C.POINTER(C.c_int)]
        _lib.sim_undo.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        _lib.sim_observe.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        _lib.sim_scores.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
        _lib.sim_playout.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_uint32, C.c_int,
                                     C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_int)]
        _lib.sim_philox.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p]
        _lib.sim_philox_rk.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p]
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Sim:
    """One env in the packed layout, driven by the host build of the device code."""

    def __init__(self, P):
        self.P = P
        self.w = np.zeros(8 + 4 * P, np.uint32)
        self.lists = None
        self.seed = 0
        self.game = 0

    @property
    def words(self):
        return self.w.reshape(-1, 1, 4)

    def reset(self, seed, game, want_lists=False):
        self.seed, self.game, self.lists = seed, game, None
        lists = np.zeros(90, np.uint8) if want_lists else None
        nobles = np.zeros(5, np.uint8)
        lib().sim_reset(_p(self.w), seed, game, self.P, _p(lists), _p(nobles))
        return lists, nobles

    def inject(self, lists, nobles):
        self.lists = np.ascontiguousarray(lists, np.uint8)
        n = np.full(5, 0xFF, np.uint8)
        n[:len(nobles)] = nobles
        lib().sim_inject(_p(self.w), _p(self.lists), _p(n), self.P)

    def legal_select(self):
        out = np.zeros(4096, np.uint16)
        n = lib().sim_legal_select(_p(self.w), self.P, _p(out))
        return out[:n].copy()

    def legal_member(self):
        out = np.zeros(4096, np.uint16)
        n = lib().sim_legal_member(_p(self.w), self.P, _p(out))
        return out[:n].copy()

    def mask(self, seat=-1):
        out = np.zeros(110, np.uint32)
        lib().sim_mask(_p(self.w), self.P, seat, _p(out))
        return out

    def step(self, idx, flags=0):
        r, d = C.c_int(), C.c_int()
        bad = lib().sim_step(_p(self.w), _p(self.lists), self.seed, self.game, self.P, int(idx), flags, C.byref(r), C.byref(d))
        return bad, r.value, d.value

    def undo(self, seat, idx, card, noble, returned):
        r = np.ascontiguousarray(returned, np.uint8)
        lib().sim_undo(_p(self.w), self.P, int(seat), int(idx), int(card), int(noble), _p(r))

    def observe(self, seat):
        out = np.zeros(265, np.float32)
        lib().sim_observe(_p(self.w), self.P, seat, _p(out))
        return out

    def scores(self, flags=0):
        sc, oc, k = np.zeros(4, np.int32), np.zeros(4, np.int32), C.c_int()
        lib().sim_scores(_p(self.w), self.P, flags, _p(sc), _p(oc), C.byref(k))
        return sc[:self.P], oc[:self.P], k.value

    def playout(self, seed, game, flags=0, max_steps=1000, fresh=True):
        self.seed, self.game = seed, game
        trace = np.full(max_steps, -1, np.int16)
        sc, k = np.zeros(4, np.int32), C.c_int()
        steps = lib().sim_playout(_p(self.w), int(fresh), _p(self.lists), seed, game, self.P, flags, max_steps, _p(trace),
                                  max_steps, _p(sc), C.byref(k))
        return steps, trace, sc[:self.P], k.value


def mask_to_indices(mask):
    bits = np.unpackbits(mask.view(np.uint8), bitorder="little")
    return np.nonzero(bits)[0].astype(np.uint16)

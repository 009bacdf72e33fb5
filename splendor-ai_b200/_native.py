"""Builds and loads libsplendor_b200.so (the C ABI of include/splendor_b200.h) with ctypes.

There is NO CPU fallback: if the library is missing or no CUDA device is present, every compute entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.environ.get("SPL_LIB") or os.path.join(CSRC, "libsplendor_b200.so")  # SPL_LIB: A/B a variant build
INCLUDE = os.path.join(os.path.dirname(_HERE), "include")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC"]

SPL_OK, SPL_E_BADARG, SPL_E_CUDA, SPL_E_NODEVICE, SPL_E_RANGE = 0, -1, -2, -3, -4
F_LIMIT_ROUNDS = 1
F_SELF_PLAY = 2
END_NONE, END_SCORE, END_DEADLOCK, END_ROUNDCAP, END_TRUNCATED = range(5)
CTR_NAMES = ("games", "steps", "wins0", "wins1", "wins2", "wins3", "ties0", "ties1", "ties2", "ties3",
             "end_score", "end_deadlock", "end_roundcap", "end_truncated", "score2_sum", "nobles")

EXPORTS = ("spl_abi_version", "spl_state_words", "spl_last_cuda_error", "spl_error_string", "spl_warmup", "spl_reset",
           "spl_inject", "spl_legal_mask", "spl_step", "spl_final_scores", "spl_observe", "spl_playout", "spl_playout_from",
           "spl_env_step", "spl_playout_host", "spl_playout_host_async", "spl_undo", "spl_count_legal", "spl_expand", "spl_masked_sample",
           "spl_minimax_eval", "spl_minimax2")


class SplendorNativeError(RuntimeError):
    pass


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise SplendorNativeError("nvcc not found: cannot build libsplendor_b200.so")


def _sources():
    return [os.path.join(CSRC, "kernels.cu"), os.path.join(CSRC, "engine.cuh"), os.path.join(CSRC, "minimax.cuh"),
            os.path.join(INCLUDE, "spl_tables.h"), os.path.join(INCLUDE, "splendor_b200.h")]


def build(force: bool = False, verbose: bool = False) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> csrc/libsplendor_b200.so (in-tree; cross-compiles without a GPU)."""
    stale = not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < max(map(os.path.getmtime, _sources()))
    if force or stale:
        cmd = [_nvcc(), *NVCC_FLAGS, "-o", LIB_PATH, os.path.join(CSRC, "kernels.cu")]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        subprocess.check_call(cmd, cwd=CSRC)
    return LIB_PATH


_lib = None


def lib():
    """The loaded library.  Raises (never falls back) if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SplendorNativeError(
                f"{LIB_PATH} is missing - run `python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc). "
                "There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        vp, u64, i64, i32, u32 = C.c_void_p, C.c_uint64, C.c_int64, C.c_int, C.c_uint32
        L.spl_abi_version.restype = i32
        L.spl_state_words.argtypes = [i32]
        L.spl_error_string.argtypes = [i32]
        L.spl_error_string.restype = C.c_char_p
        L.spl_reset.argtypes = [vp, u64, u64, vp, vp, i64, i32, vp, vp, vp]
        L.spl_inject.argtypes = [vp, vp, vp, i64, i32, vp]
        L.spl_legal_mask.argtypes = [vp, vp, i64, i32, vp]
        L.spl_step.argtypes = [vp, vp, u64, u64, vp, vp, vp, vp, vp, i64, i32, u32, vp]
        L.spl_count_legal.argtypes = [vp, vp, i64, i32, vp]
        L.spl_expand.argtypes = [vp, vp, u64, u64, vp, vp, vp, vp, vp, vp, i64, i64, i32, vp]
        L.spl_minimax_eval.argtypes = [vp, vp, vp, i64, i32, vp]
        L.spl_minimax2.argtypes = [vp, vp, u64, u64, vp, vp, vp, vp, vp, vp, i64, i64, i32, vp]
        L.spl_masked_sample.argtypes = [vp, i64, vp, u64, u64, vp, u32, i32, vp, vp, i64, vp]
        L.spl_undo.argtypes = [vp, vp, vp, vp, vp, vp, i64, i32, vp]
        L.spl_final_scores.argtypes = [vp, vp, vp, vp, i64, i32, u32, vp]
        L.spl_observe.argtypes = [vp, vp, vp, i64, i32, vp]
        L.spl_playout.argtypes = [u64, u64, i64, i32, u32, i32, vp, vp, vp, vp, vp, i32, vp, vp]
        L.spl_playout_from.argtypes = [vp, vp, u64, u64, vp, i64, i32, u32, i32, vp, vp, vp, vp, vp, i32, vp]
        L.spl_env_step.argtypes = [vp, vp, u64, u64, vp, vp, vp, vp, vp, vp, vp, vp, i64, i32, u32, vp]
        L.spl_playout_host.argtypes = [vp, vp, u64, u64, i64, i32, u32, i32, vp, vp, vp, vp, vp]
        L.spl_playout_host_async.argtypes = [vp, vp, u64, u64, i64, i32, u32, i32, vp, vp, vp, vp, i32, vp]
        for name in EXPORTS:
            getattr(L, name)  # AttributeError here = the .so does not match include/splendor_b200.h
        _lib = L
    return _lib


def check(rc: int, what: str) -> None:
    if rc != SPL_OK:
        msg = lib().spl_error_string(rc).decode()
        raise SplendorNativeError(f"{what} failed: {msg} (code {rc})")

"""CPU-side checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads, and exports exactly the symbols
include/splendor_b200.h declares.  No compute calls (no GPU here)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def native():
    import splendor_ai_b200 as pkg

    pkg.build()
    return pkg._native


def test_library_exports_every_declared_symbol(native):
    header = open(os.path.join(ROOT, "include", "splendor_b200.h")).read()
    declared = set(re.findall(r"^(?:int|const char \*)\s*(spl_\w+)\(", header, flags=re.M))
    assert len(declared) >= 14
    lib = ctypes.CDLL(native.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(native.EXPORTS)


def test_abi_constants(native):
    lib = native.lib()
    assert lib.spl_abi_version() == 1
    assert [lib.spl_state_words(p) for p in (2, 3, 4)] == [16, 20, 24]
    assert lib.spl_state_words(5) == native.SPL_E_BADARG
    assert b"no CPU fallback" in lib.spl_error_string(native.SPL_E_NODEVICE)


def test_no_cpu_fallback_without_a_gpu(native):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from splendor_ai_b200 import BatchedSplendor, playout_games

    with pytest.raises(native.SplendorNativeError):
        BatchedSplendor(8)
    with pytest.raises(native.SplendorNativeError):
        playout_games(8)
    # the raw ABI refuses as well instead of computing on the host
    assert native.lib().spl_playout(0, 0, 8, 2, 0, 100, None, None, None, None, None, 0, None, None) == native.SPL_E_NODEVICE


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "splendor-ai_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.lower(), os.path.join(dirpath, f)

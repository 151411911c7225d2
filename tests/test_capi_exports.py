"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU and exports
every symbol include/gamma_b200.h declares; creation fails loudly (no fallback) when no device."""
import ctypes
import os
import re

import pytest

from gamma_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "gamma_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gb200_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_documented_surface():
    syms = declared_symbols()
    assert "gb200_ivfpq_search" in syms and "gb200_flat_search" in syms and len(syms) >= 25
    assert set(syms) == set(api.EXPORTS), set(syms) ^ set(api.EXPORTS)


def test_library_exports_every_declared_symbol():
    L = ctypes.CDLL(api.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(L, s)]
    assert not missing, missing


def test_no_silent_fallback_without_gpu():
    if api.lib().gb200_device_count() > 0:
        pytest.skip("GPU present")
    ix = api.B200IVFPQ(0)
    rc = ix.Init('{"ncentroids": 16, "nsubvector": 8, "metric_type": "L2"}', 32)
    assert rc != 0  # creation must fail, never fall back to a CPU path
    assert api.lib().gb200_last_error()

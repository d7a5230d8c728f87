"""CPU tier: the C-ABI library builds for sm_100a, loads without a GPU and exports every symbol that
include/jtb_check.h declares; the ctypes struct images match the header's layout."""
import ctypes
import os
import re

import pytest

from jepsen_tigerbeetle_b200 import abi, history, native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "jtb_check.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(jtb_[a-z_]+)\s*\(", src)))


def test_header_and_binding_agree_on_exports():
    assert declared_functions() == sorted(native.EXPORTS)


def test_library_builds_loads_and_exports_every_symbol():
    native.build()
    lib = native.lib()
    for name in declared_functions():
        assert hasattr(lib, name), name
    assert lib.jtb_abi_version() == abi.ABI_VERSION


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    assert native.device_count() <= 0
    with pytest.raises(native.NativeError):
        native.Context(device=0)


def test_struct_sizes_match_header():
    # sizes computed from the C declarations (LP64, natural alignment)
    assert ctypes.sizeof(history.CHistory) == 8 + 12 * 8 + 8 + 8 + 2 * 8  # n_events, 12 ptrs, n_payload, n_shards(+pad), 2 ptrs
    assert ctypes.sizeof(history.CModel) == 4 * 3 + 4 * 8 + 4 * 8 + 4
    assert ctypes.sizeof(abi.COpts) == 32
    assert ctypes.sizeof(abi.CLinShard) == 32
    assert ctypes.sizeof(abi.CLinResult) == 56
    assert ctypes.sizeof(abi.CSetFullShard) == 48
    assert ctypes.sizeof(abi.CSetFullOut) == 8 * 7 + 8 + 16
    assert ctypes.sizeof(abi.CBankResult) == 8 + 16 + 8 + 40 + 60 + 4 + 16 + 8 + 16


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "jepsen_tigerbeetle_b200")
    for dirpath, _dirs, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "oracle/" not in text or f in ("jtb_prep.h",), f

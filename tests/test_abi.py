"""The C-ABI shared library loads and exports every symbol include/atropos_b200.h declares; the ctypes
mirror agrees with the header; and without a GPU the product fails loudly instead of falling back."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "atropos_b200.h")


@pytest.fixture(scope="module")
def lib():
    from atropos_b200 import build, _lib
    build.build()
    return _lib.load()


def _declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(atr_[a-z0-9_]+)\s*\(", text)))


def test_exports_every_declared_symbol(lib):
    from atropos_b200 import _lib
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), "library does not export %s" % name
    assert sorted(_lib.SYMBOLS) == declared


def test_argtypes_match_header_arity(lib):
    """every ctypes prototype has as many arguments as the C declaration"""
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    for name, params in re.findall(r"\b(atr_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        params = params.strip()
        arity = 0 if params in ("", "void") else params.count(",") + 1
        fn = getattr(lib, name)
        if fn.argtypes is not None:
            assert len(fn.argtypes) == arity, (name, len(fn.argtypes), arity)
        else:
            assert arity == 0, name


def test_struct_layout(lib):
    from atropos_b200 import _abi
    assert lib.atr_abi_version() == _abi.ATR_ABI_VERSION
    assert ctypes.sizeof(_abi.AtrMatch) == 16
    assert _abi.INSERT_DTYPE.itemsize == 48
    # 8-byte aligned double after an int32 -> the C compiler pads exactly like ctypes
    assert _abi.AtrAdapterDesc.max_error_rate.offset == 16
    assert ctypes.sizeof(_abi.AtrAdapterDesc) == 64


def test_packed_words_host_helper(lib):
    import numpy as np
    offs = np.array([0, 150, 150, 158, 159], dtype=np.int64)
    assert lib.atr_packed_words(offs.ctypes.data, 4) == 19 + 0 + 1 + 1


def test_no_silent_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from atropos_b200 import _lib
    from atropos_b200.align import Aligner
    with pytest.raises(_lib.EngineError):
        Aligner("ACGT", 0.1).locate("TTACGTTT")


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "atropos_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "liboracle" not in text, f

"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/bls12381_b200.h declares, and fails loudly (no CPU fallback) without a GPU."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "bls12381_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported():
    from bls12_381_b200 import _lib
    lib = _lib.load()
    syms = declared_symbols()
    assert len(syms) >= 40
    for s in syms:
        assert hasattr(lib, s), "libbls12381_b200.so does not export %s" % s
    assert set(_lib.SIGNATURES) == set(syms)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import bls12_381_b200
    with pytest.raises(bls12_381_b200.B200Error):
        bls12_381_b200.Engine()


def test_product_does_not_import_oracle():
    """the product package must never reach into oracle/ (parity would be void)"""
    pkg = os.path.join(ROOT, "bls12_381_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no CPU fallback", ""), os.path.join(dirpath, f)

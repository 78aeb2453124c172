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


def test_rust_sys_binding_covers_the_header():
    """bindings/rust/bls12381-b200-sys/src/lib.rs is generated from the header (tools/gen_rust_sys.py): every
    exported symbol must be bound, with the same number of parameters"""
    rs = open(os.path.join(ROOT, "bindings", "rust", "bls12381-b200-sys", "src", "lib.rs")).read()
    hdr = open(os.path.join(ROOT, "include", "bls12381_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    for sym in declared_symbols():
        m = re.search(r"pub fn %s\(([^)]*)\)" % sym, rs)
        assert m, "Rust binding misses %s" % sym
        c = re.search(r"\b%s\s*\(([^;{]*?)\)\s*;" % sym, hdr)
        nargs_c = len([a for a in c.group(1).split(",") if a.strip() and a.strip() != "void"])
        nargs_rs = len([a for a in m.group(1).split(",") if a.strip()])
        assert nargs_c == nargs_rs, sym

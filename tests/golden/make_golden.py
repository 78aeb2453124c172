#!/usr/bin/env python
"""Regenerates tests/golden/* from the reference's OWN test vectors (run in the build container,
where /root/reference exists; the outputs are committed so the GPU box never needs the reference).

 * dat_vectors.npz  : the four src/tests/*.dat golden files ([i]G for i=0..999, G1/G2,
                      compressed/uncompressed; src/tests/mod.rs:3-76), stored as uint8 arrays.
 * kat.json         : every 64-bit hex literal of the known-answer tests listed in SURVEY.md §8(c),
                      in source order, grouped 6-per-Fp (4-per-Scalar), keyed "file::function".
                      tests/test_oracle_golden.py documents how each list is interpreted.  Also the scalar-field
                      constants of src/scalar.rs:76-222 ("scalar.rs::const_*"), the byte vectors of its
                      test_to_bytes / test_from_bytes, and src/hash_to_curve/map_g1.rs::test_simple_swu_expected.
 * h2c_vectors.json : the RFC 9380 (draft-16) vectors of the reference's integration tests — tests/expand_msg.rs
                      (expand_message_xmd, SHA-256, short and long DST), tests/hash_to_curve_g1.rs and _g2.rs
                      (hash_to_curve / encode_to_curve, uncompressed encodings) — and
                      src/hash_to_curve/map_scalar.rs::test_hash_to_scalar, all as hex strings.
"""
import json, re, sys, os
import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
HEX64 = re.compile(r"0x([0-9a-fA-F]{4}(?:_[0-9a-fA-F]{4}){3})\b")


def fn_body(path, name):
    lines = open(path).read().split("\n")
    start = next(i for i, l in enumerate(lines) if re.match(r"\s*(pub )?(const )?fn %s\(" % re.escape(name), l))
    indent = len(lines[start]) - len(lines[start].lstrip())
    out = []
    for l in lines[start + 1:]:
        if l.startswith(" " * indent + "}") and len(l.rstrip()) == indent + 1:
            break
        out.append(l)
    return "\n".join(out)


def limbs(text, per):
    toks = [int(m.replace("_", ""), 16) for m in HEX64.findall(text)]
    assert len(toks) % per == 0, (len(toks), per)
    return [["%016x" % v for v in toks[i:i + per]] for i in range(0, len(toks), per)]


KATS = {
    "fp.rs": (6, ["test_squaring", "test_multiplication", "test_addition", "test_subtraction", "test_negation",
                  "test_sqrt", "test_inversion", "test_lexicographic_largest"]),
    "fp2.rs": (6, ["test_squaring", "test_multiplication", "test_addition", "test_subtraction", "test_negation",
                   "test_sqrt", "test_inversion", "test_lexicographic_largest"]),
    "fp6.rs": (6, ["test_arithmetic"]),
    "fp12.rs": (6, ["test_arithmetic"]),
    "g1.rs": (6, ["test_doubling", "test_projective_addition", "test_mixed_addition", "test_beta", "test_is_torsion_free"]),
    "hash_to_curve/map_g1.rs": (6, ["test_simple_swu_expected"]),
    "g2.rs": (6, ["test_doubling", "test_projective_addition", "test_mixed_addition", "test_is_torsion_free"]),
    "pairings.rs": (6, ["generator"]),           # Gt::generator(), src/pairings.rs:359-475
    "scalar.rs": (4, ["test_from_bytes_wide_maximum", "test_addition", "test_double"]),
    "tests/mod.rs": (6, ["test_pairing_result_against_relic"]),   # expected Gt, Montgomery limbs (:114-231)
}


def rust_bytes(lit):
    """body of a Rust b"..." literal -> bytes (handles the backslash-newline continuation and \\x escapes)"""
    lit = re.sub(r"\\\n\s*", "", lit)
    return lit.encode("latin1").decode("unicode_escape").encode("latin1")


def h2c_vectors():
    """RFC 9380 (draft-16) vectors the reference's integration tests hold: tests/expand_msg.rs (XMD SHA-256, short and
    long DST), tests/hash_to_curve_g1.rs, tests/hash_to_curve_g2.rs.  Everything as hex strings."""
    out = {}
    src = open(os.path.join(REF, "tests/expand_msg.rs")).read()
    for fn in ["expand_msg_xmd_works_for_draft16_testvectors_sha256", "expand_msg_xmd_works_for_draft16_testvectors_sha256_long_dst"]:
        body = fn_body(os.path.join(REF, "tests/expand_msg.rs"), fn)
        dst = rust_bytes(re.search(r'let dst = b"(.*?)";', body, re.S).group(1))
        cases = []
        for m in re.finditer(r'msg: b"(.*?)",\s*dst,\s*len_in_bytes: (0x[0-9a-f]+),\s*uniform_bytes: &hex!\(\s*"(.*?)"\s*\)', body, re.S):
            cases.append({"msg": rust_bytes(m.group(1)).hex(), "len_in_bytes": int(m.group(2), 16),
                          "uniform_bytes": re.sub(r"\s+", "", m.group(3))})
        assert len(cases) == 10, (fn, len(cases))
        out["expand_msg.rs::" + fn] = {"dst": dst.hex(), "cases": cases}
    for f, fns in (("hash_to_curve_g1.rs", ["encode_to_curve_works_for_draft16_testvectors_g1_sha256_nu",
                                            "hash_to_curve_works_for_draft16_testvectors_g1_sha256_ro"]),
                   ("hash_to_curve_g2.rs", ["encode_to_curve_works_for_draft16_testvectors_g2_sha256_nu",
                                            "hash_to_curve_works_for_draft16_testvectors_g2_sha256_ro"])):
        for fn in fns:
            body = fn_body(os.path.join(REF, "tests", f), fn)
            dst = rust_bytes(re.search(r'let dst = b"(.*?)";', body, re.S).group(1))
            cases = []
            for m in re.finditer(r'msg: b"(.*?)",\s*dst,\s*expected: &hex!\(\s*"(.*?)"\s*\)', body, re.S):
                cases.append({"msg": rust_bytes(m.group(1)).hex(), "expected": re.sub(r"\s+", "", m.group(2))})
            assert len(cases) == 5, (fn, len(cases))
            out[f + "::" + fn] = {"dst": dst.hex(), "cases": cases}
    # src/hash_to_curve/map_scalar.rs:25-45 (hash_to_field for Scalar)
    body = fn_body(os.path.join(REF, "src/hash_to_curve/map_scalar.rs"), "test_hash_to_scalar")
    out["map_scalar.rs::test_hash_to_scalar"] = [
        {"okm": rust_bytes(a).hex(), "expected": b} for a, b in re.findall(r'b"(.*?)",\s*"0x([0-9a-f]{64})"', body, re.S)]
    return out


def main():
    kat = {}
    for f, (per, names) in KATS.items():
        for n in names:
            kat["%s::%s" % (f, n)] = limbs(fn_body(os.path.join(REF, "src", f), n), per)
    # scalar-field constants (src/scalar.rs:76-222, LARGEST :1051) and the byte encodings of test_to_bytes /
    # test_from_bytes (:864-968): 32-byte decimal arrays in source order
    sc = open(os.path.join(REF, "src/scalar.rs")).read()
    for name in ["MODULUS", "GENERATOR", "R", "R2", "R3", "TWO_INV", "ROOT_OF_UNITY", "ROOT_OF_UNITY_INV", "DELTA", "LARGEST"]:
        m = re.search(r"const %s: Scalar = Scalar\(\[(.*?)\]\);" % name, sc, re.S)
        kat["scalar.rs::const_%s" % name] = limbs(m.group(1), 4)[0]
    for fn in ["test_to_bytes", "test_from_bytes"]:
        arrs = re.findall(r"\[\s*((?:\d+,\s*){31}\d+)\s*\]", fn_body(os.path.join(REF, "src/scalar.rs"), fn))
        kat["scalar.rs::%s_bytes" % fn] = [[int(x) for x in a.replace("\n", " ").split(",")] for a in arrs]
    # RELIC-derived pairing KAT bytes (src/tests/mod.rs:78-231): the expected Gt as raw byte list
    # the same value as sent by the RELIC author, canonical (non-Montgomery) big-endian hex (src/tests/mod.rs:81-95)
    body = fn_body(os.path.join(REF, "src/tests/mod.rs"), "test_pairing_result_against_relic")
    words = re.findall(r"\b([0-9A-F]{16})\b", body.split("*/")[0])
    assert len(words) == 72
    kat["tests/mod.rs::relic_canonical_hex"] = ["".join(words[i:i + 6]) for i in range(0, 72, 6)]
    json.dump(kat, open(os.path.join(HERE, "kat.json"), "w"), indent=0)
    json.dump(h2c_vectors(), open(os.path.join(HERE, "h2c_vectors.json"), "w"), indent=0)
    d = {}
    for name in ["g1_compressed", "g1_uncompressed", "g2_compressed", "g2_uncompressed"]:
        d[name] = np.fromfile(os.path.join(REF, "src/tests/%s_valid_test_vectors.dat" % name), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "dat_vectors.npz"), **d)
    print({k: len(v) for k, v in kat.items()}, {k: v.shape for k, v in d.items()})


if __name__ == "__main__":
    main()

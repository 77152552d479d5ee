"""The reference's golden vectors for the requirement algebra (tests/golden/requirement_golden.json, extracted from
pkg/scheduling/requirement_test.go and requirements_test.go by tests/golden/extract_golden.py), checked against
  (1) the CPU oracle's string/std::set implementation, and
  (2) the product's dictionary-bitmask implementation (csrc/reqmask.cuh — the same code the CUDA kernels compile),
      evaluated on the host."""
import ctypes as C
import json
from pathlib import Path

import pytest

GOLD = json.loads((Path(__file__).parent / "golden" / "requirement_golden.json").read_text())
DICT = ["1", "2", "9", "A", "B"]  # kh_mask_* test dictionary (scheduler.cc TestDict)


def split(spec):
    op, _, vals = spec.partition(":")
    return op.encode(), vals.encode()


def test_counts_match_the_reference_tables():
    assert len(GOLD["intersection"]) == 196 and len(GOLD["has"]) == 70 and len(GOLD["compatible"]) == 225
    assert sum(1 for c in GOLD["compatible"] if not c["want"]) == 94  # "the 94 failing pairs"


# ------------------------------------------------------------------ oracle
def test_oracle_intersection(oracle):
    for c in GOLD["intersection"]:
        assert oracle.intersection(c["a"], c["b"]) == c["want"], c


def test_oracle_has_operator_len_string(oracle):
    for c in GOLD["has"]:
        assert oracle.has(c["a"], c["value"]) == c["want"], c
    for c in GOLD["operator"]:
        assert oracle.operator(c["a"]) == c["want"], c
    for c in GOLD["len"]:
        assert oracle.length(c["a"]) == c["want"], c
    for c in GOLD["string"]:
        assert oracle.string(c["a"], c["b"]) == c["want"], c


def test_oracle_compatible(oracle):
    for c in GOLD["compatible"]:
        assert oracle.compatible(GOLD["compatible_key"], c["a"], c["b"]) == c["want"], c


def test_oracle_normalises_aliased_labels(oracle):
    for k, v in GOLD["normalize"].items():
        assert oracle.normalize(k) == v
    assert oracle.normalize("custom/label") == "custom/label"


# ------------------------------------------------------------------ product bitmask algebra (host build of reqmask.cuh)
def render_mask(out):
    present, comp, values, has_gt, gt, has_lt, lt = out[:7]
    vals = sorted(DICT[b] for b in range(len(DICT)) if (values >> b) & 1)
    return "complement=%d;values=%s;gt=%s;lt=%s" % (comp, ",".join(vals), gt if has_gt else "", lt if has_lt else "")


def test_mask_intersection(pkg):
    L = pkg.lib()
    out = (C.c_longlong * 9)()
    for c in GOLD["intersection"]:
        L.kh_mask_intersection(*split(c["a"]), *split(c["b"]), out)
        assert render_mask(list(out)) == c["want"], c


def test_mask_has(pkg):
    L = pkg.lib()
    for c in GOLD["has"]:
        allowed = L.kh_mask_allowed(*split(c["a"]))
        got = bool((allowed >> DICT.index(c["value"])) & 1)
        assert got == c["want"], c


def test_mask_operator_classes(pkg):
    """Operator() only matters to the kernels as negative / non-negative and Len()==0"""
    L = pkg.lib()
    out = (C.c_longlong * 9)()
    for c in GOLD["operator"]:
        L.kh_mask_intersection(*split(c["a"]), b"Exists", b"", out)
        assert bool(out[8]) == (c["want"] in ("NotIn", "DoesNotExist")), c
    for c in GOLD["len"]:
        L.kh_mask_intersection(*split(c["a"]), b"Exists", b"", out)
        assert bool(out[7]) == (c["want"] == 0), c


def test_mask_compatible(pkg):
    L = pkg.lib()
    for c in GOLD["compatible"]:
        a_op, a_vals = split(c["a"]) if c["a"] else (b"Exists", b"")
        b_op, b_vals = split(c["b"]) if c["b"] else (b"Exists", b"")
        got = bool(L.kh_mask_compatible(a_op, a_vals, int(bool(c["a"])), b_op, b_vals, int(bool(c["b"])), 1))
        assert got == c["want"], c


def test_mask_custom_label_rule(pkg, oracle):
    """Compatible on a key that is NOT well known: undefined on the node is denied unless the operator is negative
    (requirements.go:125-130, suite_test.go:401-434)."""
    L = pkg.lib()
    for spec, want in (("In:A", False), ("NotIn:A", True), ("Exists", False), ("DoesNotExist", True)):
        op, vals = split(spec)
        assert bool(L.kh_mask_compatible(b"Exists", b"", 0, op, vals, 1, 0)) == want
        assert oracle.compatible("custom-key", "", spec, well_known=False) == want

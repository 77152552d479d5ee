#!/usr/bin/env python3
"""Extract the reference's own golden vectors for the requirement algebra into JSON fixtures.

Run HERE (the container that has /root/reference); the GPU box never sees the reference, only
the committed fixtures. Sources (expectations are parsed, no reference code is copied):

  pkg/scheduling/requirement_test.go   Intersection (196), Has (70), Operator (14), Len (14), String (16)
  pkg/scheduling/requirements_test.go  Compatible (15 x 15 = 225 on the well-known zone key)

Output: tests/golden/requirement_golden.json
"""
import json
import re
import sys
from pathlib import Path

REF = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
OUT = Path(__file__).resolve().parent / "requirement_golden.json"

OPS = {
    "NodeSelectorOpExists": "Exists", "NodeSelectorOpDoesNotExist": "DoesNotExist", "NodeSelectorOpIn": "In",
    "NodeSelectorOpNotIn": "NotIn", "NodeSelectorOpGt": "Gt", "NodeSelectorOpLt": "Lt",
}


def parse_defs(text, ctor):
    """name := NewRequirement("key", v1.NodeSelectorOpIn, "A", "B") -> {name: "In:A,B"}"""
    defs = {}
    pat = re.compile(r"(\w+)\s*:=\s*" + ctor + r"\((?:NewRequirement\()?[^,\n]+,\s*v1\.(\w+)((?:,\s*\"[^\"]*\")*)\)")
    for m in pat.finditer(text):
        name, op, vals = m.group(1), OPS[m.group(2)], re.findall(r'"([^"]*)"', m.group(3))
        defs[name] = op + (":" + ",".join(vals) if vals else "")
    return defs


def render(complement, values, gt, lt):
    return "complement=%d;values=%s;gt=%s;lt=%s" % (complement, ",".join(sorted(values)), gt, lt)


def render_spec(spec):
    op, _, vals = spec.partition(":")
    vals = vals.split(",") if vals else []
    if op == "In":
        return render(0, vals, "", "")
    if op == "NotIn":
        return render(1, vals, "", "")
    if op == "Exists":
        return render(1, [], "", "")
    if op == "DoesNotExist":
        return render(0, [], "", "")
    if op == "Gt":
        return render(1, [], vals[0], "")
    if op == "Lt":
        return render(1, [], "", vals[0])
    raise ValueError(spec)


def main():
    req = (REF / "pkg/scheduling/requirement_test.go").read_text()
    defs = parse_defs(req, "NewRequirement")
    assert len(defs) == 14, defs
    golden = {"defs": defs, "intersection": [], "has": [], "operator": [], "len": [], "string": []}

    # Expect(a.Intersection(b)).To(Equal(c))  or  Equal(&Requirement{...})
    for m in re.finditer(r"Expect\((\w+)\.Intersection\((\w+)\)\)\.To\(Equal\((.*)\)\)\s*$", req, re.M):
        a, b, exp = m.group(1), m.group(2), m.group(3).strip()
        if exp in defs:
            want = render_spec(defs[exp])
        else:
            lit = re.match(r"&Requirement\{(.*)\}$", exp)
            assert lit, exp
            body = lit.group(1)
            comp = 1 if "complement: true" in body else 0
            vals = re.findall(r'"([^"]*)"', re.search(r"values:\s*sets\.NewString\(([^)]*)\)", body).group(1))
            gt = re.search(r"greaterThan:\s*(\w+)\.greaterThan", body)
            lt = re.search(r"lessThan:\s*(\w+)\.lessThan", body)
            gtv = defs[gt.group(1)].split(":")[1] if gt else ""
            ltv = defs[lt.group(1)].split(":")[1] if lt else ""
            want = render(comp, vals, gtv, ltv)
        golden["intersection"].append({"a": defs[a], "b": defs[b], "want": want})
    assert len(golden["intersection"]) == 196, len(golden["intersection"])

    for m in re.finditer(r"Expect\((\w+)\.Has\(\"([^\"]*)\"\)\)\.To\((BeTrue|BeFalse)\(\)\)", req):
        golden["has"].append({"a": defs[m.group(1)], "value": m.group(2), "want": m.group(3) == "BeTrue"})
    assert len(golden["has"]) == 70, len(golden["has"])

    for m in re.finditer(r"Expect\((\w+)\.Operator\(\)\)\.To\(Equal\(v1\.(\w+)\)\)", req):
        golden["operator"].append({"a": defs[m.group(1)], "want": OPS[m.group(2)]})
    assert len(golden["operator"]) == 14

    for m in re.finditer(r"Expect\((\w+)\.Len\(\)\)\.To\(Equal\(([^)]*)\)\)", req):
        expr = m.group(2).replace("math.MaxInt64", str(2**63 - 1))
        golden["len"].append({"a": defs[m.group(1)], "want": eval(expr)})  # noqa: S307 - arithmetic on literals only
    assert len(golden["len"]) == 14

    for m in re.finditer(r"Expect\((\w+)(?:\.Intersection\((\w+)\))?\.String\(\)\)\.To\(Equal\(\"([^\"]*)\"\)\)", req):
        golden["string"].append({"a": defs[m.group(1)], "b": defs[m.group(2)] if m.group(2) else "", "want": m.group(3)})
    assert len(golden["string"]) == 16, len(golden["string"])

    reqs = (REF / "pkg/scheduling/requirements_test.go").read_text()
    rdefs = parse_defs(reqs, "NewRequirements")
    rdefs = {k: v for k, v in rdefs.items() if k in defs}
    rdefs["unconstrained"] = ""
    assert len(rdefs) == 15, rdefs
    golden["compatible_key"] = "topology.kubernetes.io/zone"
    golden["compatible"] = []
    for m in re.finditer(r"Expect\((\w+)\.Compatible\((\w+)\)\)\.(To|ToNot)\(Succeed\(\)\)", reqs):
        golden["compatible"].append({"a": rdefs[m.group(1)], "b": rdefs[m.group(2)], "want": m.group(3) == "To"})
    assert len(golden["compatible"]) == 225, len(golden["compatible"])

    # alias normalisation (requirement_test.go:45-79, requirements_test.go:27-31)
    golden["normalize"] = {
        "failure-domain.beta.kubernetes.io/zone": "topology.kubernetes.io/zone",
        "failure-domain.beta.kubernetes.io/region": "topology.kubernetes.io/region",
        "beta.kubernetes.io/arch": "kubernetes.io/arch",
        "beta.kubernetes.io/os": "kubernetes.io/os",
        "beta.kubernetes.io/instance-type": "node.kubernetes.io/instance-type",
    }
    OUT.write_text(json.dumps(golden, indent=1, sort_keys=True) + "\n")
    print("wrote", OUT, {k: len(v) for k, v in golden.items() if isinstance(v, list)})


if __name__ == "__main__":
    main()

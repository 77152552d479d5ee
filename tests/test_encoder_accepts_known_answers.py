"""Every known-answer problem (including the ones that have only run against the oracle so far) must be ENCODABLE by the
product's host layer: no "unsupported" refusal, and the encoded launch-choice table must reproduce the oracle's launch
choices through the shared host/device core. CPU only - the solve itself needs a GPU."""
import pytest

import consolidation_answers as ca
import known_answers as ka

ALL = ka.CASES + ka.CPU_ONLY_CASES + ca.CASES + ca.CPU_ONLY_CASES


@pytest.mark.parametrize("name,ref,build", ALL, ids=[c[0] for c in ALL])
def test_problem_is_encodable_and_launch_table_is_consistent(pkg, oracle, name, ref, build):
    prob, _ = build()
    for pd in (prob["multi"] if "multi" in prob else [prob]):
        problem = pkg.Problem.from_dict(pd)
        rs = pkg.ResidentSolve(problem)  # kh_encode: dictionary, classes, topology groups, existing nodes
        assert rs.dims["pods"] >= len(pd["pods"])  # plus the pods of nodes that are being deleted (provisioner.go:119-144)
        res = pkg.Result()
        assert oracle.solve(problem, res) == 0, res.error
        assert pkg.lib().kh_launch_table_selfcheck(problem.ptr, res.ptr) == 0

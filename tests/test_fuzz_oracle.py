"""CPU: the oracle runs every fuzz problem without error and its result satisfies basic conservation invariants."""
import pytest

from fuzz_problems import random_problem, random_problem_with_bounds


@pytest.mark.parametrize("seed", range(40))
def test_oracle_invariants_on_random_problems(pkg, oracle, seed):
    prob = random_problem(seed)
    problem = pkg.Problem.from_dict(prob)
    res = pkg.Result()
    assert oracle.solve(problem, res) == 0, res.error
    r = res.to_dict()
    ne = len(r["existing"])
    placed = [i for i, a in enumerate(r["assign"]) if a >= 0]
    on_nodes = sorted(p for n in r["newNodes"] for p in n["pods"]) + sorted(p for e in r["existing"] for p in e["pods"])
    assert sorted(on_nodes) == placed
    for n in r["newNodes"]:
        assert n["options"], "a committed node always has at least one instance type left"
        assert n["requests"].get("pods", 0) >= 1000 * len(n["pods"])
    assert all(a < ne + len(r["newNodes"]) for a in r["assign"])


@pytest.mark.parametrize("seed", range(60))
def test_encoder_accepts_bounds_fuzz_and_oracle_runs(pkg, oracle, seed):
    """Gt / Lt / complement-type problems: the oracle runs them and the product's encoder hands them over (region form) or refuses loudly"""
    prob = random_problem_with_bounds(seed)
    problem = pkg.Problem.from_dict(prob)
    res = pkg.Result()
    assert oracle.solve(problem, res) == 0, res.error
    try:
        pkg.ResidentSolve(problem)
    except pkg.KschedError as e:
        assert e.code == pkg.KSCHED_ERR_UNSUPPORTED, str(e)

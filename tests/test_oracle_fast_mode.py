"""CPU: the oracle's FAST MODE (oracle.cc header; used only to produce the full-size golden digests under
tests/golden/fullsize/) must be result-neutral: digest-identical to the literal restatement on every known answer, the
fuzz corpus, consolidation cases and C2/C3/C4/C5-shaped samples."""
import pytest

from consolidation_answers import CASES as CONS_CASES, CPU_ONLY_CASES as CONS_CPU
from fuzz_problems import random_problem
from known_answers import CASES, CPU_ONLY_CASES


@pytest.fixture()
def both(pkg, oracle):
    def run(problem, candidates=()):
        out = []
        for fast in (0, 1):
            oracle.lib.oracle_set_fast(fast)
            try:
                res = pkg.Result()
                assert oracle.solve(problem, res, candidates) == 0, res.error
            finally:
                oracle.lib.oracle_set_fast(0)
            out.append(res)
        a, b = out
        assert a.digest() == b.digest()
        assert (a.assign == b.assign).all() and (a.relax_level == b.relax_level).all()
        assert a.nodes_visited == b.nodes_visited and a.add_calls == b.add_calls
        assert a.to_dict() == b.to_dict()
    return run


ALL = CASES + CPU_ONLY_CASES


@pytest.mark.parametrize("name,ref,build", ALL, ids=[c[0] for c in ALL])
def test_fast_mode_equals_literal_on_known_answers(pkg, both, name, ref, build):
    prob, _ = build()
    for p in prob.get("multi", [prob]):
        both(pkg.Problem.from_dict(p))


@pytest.mark.parametrize("seed", range(120))
def test_fast_mode_equals_literal_on_fuzz(pkg, both, seed):
    both(pkg.Problem.from_dict(random_problem(seed)))


@pytest.mark.parametrize("cfg,pods,types,seed", [(2, 2000, 500, 3), (3, 1200, 1000, 5), (4, 1000, 1000, 9), (4, 1500, 1000, 42)])
def test_fast_mode_equals_literal_on_baseline_shapes(pkg, both, cfg, pods, types, seed):
    both(pkg.Problem.synth(cfg, pods, types, seed, 0))


def test_fast_mode_equals_literal_on_simulations(pkg, both):
    problem = pkg.Problem.synth(5, 600, 1000, 4, 60)
    both(problem, candidates=list(range(10)))
    both(problem, candidates=list(range(0, 60, 3)))


ALLC = CONS_CASES + CONS_CPU


@pytest.mark.parametrize("name,ref,build", ALLC, ids=[c[0] for c in ALLC])
def test_fast_mode_equals_literal_on_consolidation(pkg, oracle, name, ref, build):
    prob, _ = build()
    problem = pkg.Problem.from_dict(prob)
    got = []
    for fast in (0, 1):
        oracle.lib.oracle_set_fast(fast)
        try:
            got.append(oracle.consolidate(problem))
        finally:
            oracle.lib.oracle_set_fast(0)
    assert got[0] == got[1]

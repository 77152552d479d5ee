"""Consolidation known answers against the oracle (CPU)."""
import pytest

from consolidation_answers import CASES, CPU_ONLY_CASES


ALL = CASES + CPU_ONLY_CASES


@pytest.mark.parametrize("name,ref,build", ALL, ids=[c[0] for c in ALL])
def test_oracle_consolidation_known_answer(pkg, oracle, name, ref, build):
    prob, check = build()
    problem = pkg.Problem.from_dict(prob)
    check(lambda count: oracle.consolidate_probe(problem, count), lambda: oracle.consolidate(problem))

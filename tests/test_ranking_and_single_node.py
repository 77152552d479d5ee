"""Candidate ranking and single-node consolidation known answers against the oracle (CPU)."""
import pytest

from ranking_answers import CASES, SINGLE


@pytest.mark.parametrize("name,ref,build", CASES, ids=[c[0] for c in CASES])
def test_oracle_ranking_known_answer(pkg, oracle, name, ref, build):
    prob, check = build()
    order, cost = oracle.rank_candidates(pkg.Problem.from_dict(prob))
    check(order, cost)


@pytest.mark.parametrize("name,ref,build", SINGLE, ids=[c[0] for c in SINGLE])
def test_oracle_single_node_consolidation_known_answer(pkg, oracle, name, ref, build):
    prob, check = build()
    check(oracle.consolidate_single(pkg.Problem.from_dict(prob)))

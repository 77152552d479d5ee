"""CPU: the oracle must reproduce the reference's own known answers — this is what pins it."""
import pytest

from known_answers import CASES, CPU_ONLY_CASES


def run_oracle(pkg, oracle, prob_dict):
    problem = pkg.Problem.from_dict(prob_dict)
    res = pkg.Result()
    assert oracle.solve(problem, res) == 0, res.error
    return res.to_dict()


ALL = CASES + CPU_ONLY_CASES


@pytest.mark.parametrize("name,ref,build", ALL, ids=[c[0] for c in ALL])
def test_oracle_known_answer(pkg, oracle, name, ref, build):
    prob, check = build()
    if "multi" in prob:
        check([run_oracle(pkg, oracle, p) for p in prob["multi"]])
    else:
        check(run_oracle(pkg, oracle, prob))

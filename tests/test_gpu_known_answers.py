"""GPU: every known-answer case through the CUDA path — identical to the oracle AND satisfying the reference's invariant."""
import pytest

from known_answers import CASES as _CASES, CPU_ONLY_CASES

CASES = _CASES + CPU_ONLY_CASES  # every restated known answer runs on the GPU (the second list was added after round 1's GPU budget)

pytestmark = pytest.mark.gpu


def both(pkg, oracle, prob_dict):
    problem = pkg.Problem.from_dict(prob_dict)
    want = pkg.Result()
    assert oracle.solve(problem, want) == 0, want.error
    got = pkg.Scheduler(problem).solve()
    g, w = got.to_dict(), want.to_dict()
    assert g["assign"] == w["assign"]
    assert g["relax"] == w["relax"]
    assert g["existing"] == w["existing"]
    assert len(g["newNodes"]) == len(w["newNodes"])
    for a, b in zip(g["newNodes"], w["newNodes"]):
        assert a["provisioner"] == b["provisioner"] and a["pods"] == b["pods"] and a["options"] == b["options"] and a["requests"] == b["requests"]
        assert a["requirements"] == {k: v for k, v in b["requirements"].items() if k != "node.kubernetes.io/instance-type"}
        assert a.get("launch") == b.get("launch") and a.get("launch") is not None
    fast = pkg.Scheduler(problem).solve(count_visited=False).to_dict()  # production setting: steady-state kernel paths
    assert fast["assign"] == g["assign"] and fast["newNodes"] == g["newNodes"] and fast["existing"] == g["existing"]
    return g


@pytest.mark.parametrize("name,ref,build", CASES, ids=[c[0] for c in CASES])
def test_gpu_known_answer(pkg, oracle, name, ref, build):
    prob, check = build()
    if "multi" in prob:
        check([both(pkg, oracle, p) for p in prob["multi"]])
    else:
        check(both(pkg, oracle, prob))

"""The independent placement checker (tests/placement_invariants.py): accepted on oracle results, and it must notice a
corrupted result. The GPU suite applies the same checker to the full-size BASELINE configurations."""
import json

import numpy as np
import pytest

import placement_invariants as pi


@pytest.mark.parametrize("cfg,pods,types,seed", [(4, 1500, 1000, 1), (4, 1200, 1000, 3), (3, 1200, 1000, 42), (2, 2000, 500, 42)])
def test_oracle_results_satisfy_the_invariants(pkg, oracle, cfg, pods, types, seed):
    problem = pkg.Problem.synth(cfg, pods, types, seed, 0)
    result = pkg.Result()
    oracle.solve(problem, result)
    out = pi.check(problem, result)
    assert out["new_nodes"] > 0
    if cfg == 4:
        assert out["zone"] + out["host"] + out["anti"] > 0


class _Tampered:
    def __init__(self, base, mutate):
        self._d = base.to_dict(brief=True)
        mutate(self._d)

    def to_dict(self, brief=False):
        return json.loads(json.dumps(self._d))


def test_checker_rejects_corrupted_results(pkg, oracle):
    problem = pkg.Problem.synth(4, 1500, 1000, 1, 0)
    result = pkg.Result()
    oracle.solve(problem, result)
    summ = problem.pod_summary()
    d = result.to_dict(brief=True)
    assign = np.asarray(d["assign"])
    anti_pods = [p for p in range(len(assign)) if summ[p, 3]]
    assert len(anti_pods) >= 2

    def same_node(dd):  # two pods of an anti-affinity app forced onto one node
        a, b = anti_pods[0], anti_pods[1]
        na, nb = dd["assign"][a], dd["assign"][b]
        dd["newNodes"][nb]["pods"].remove(b)
        dd["newNodes"][na]["pods"].append(b)
        dd["assign"][b] = na
        for node in (na, nb):
            pods = dd["newNodes"][node]["pods"]
            if not pods:
                dd["newNodes"][node]["pods"] = [b]  # keep the structural checks quiet; the request check will fire instead
            dd["newNodes"][node]["requests"]["cpu"] = int(summ[pods, 0].sum()) if pods else 0
            dd["newNodes"][node]["requests"]["memory"] = int(summ[pods, 1].sum()) if pods else 0
            dd["newNodes"][node]["requests"]["pods"] = 1000 * len(pods)

    with pytest.raises(AssertionError):
        pi.check(problem, _Tampered(result, same_node))

    def lost_options(dd):
        dd["newNodes"][0]["nOptions"] = 0

    with pytest.raises(AssertionError):
        pi.check(problem, _Tampered(result, lost_options))

    def wrong_requests(dd):
        dd["newNodes"][0]["requests"]["cpu"] += 1

    with pytest.raises(AssertionError):
        pi.check(problem, _Tampered(result, wrong_requests))

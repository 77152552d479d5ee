"""BASELINE.json's configurations at FULL size on the GPU. C1/C2 are compared with the oracle bit for bit (it finishes in
under a second); C3/C4/C5-shaped problems are too slow for the CPU oracle at full size, so they go through the
size-independent placement invariants (tests/placement_invariants.py) plus exact parity on a prefix-sized sample."""
import numpy as np
import pytest

import placement_invariants as pi

pytestmark = pytest.mark.gpu


def test_c2_full_size_matches_oracle(pkg, oracle):
    problem = pkg.Problem.synth(2, 10000, 500, 42, 0)
    got = pkg.Scheduler(problem).solve()
    want = pkg.Result()
    oracle.solve(problem, want)
    assert np.array_equal(got.assign, want.assign)
    assert got.digest() == want.digest()
    pi.check(problem, got)


def test_c3_full_size_invariants(pkg):
    problem = pkg.Problem.synth(3, 50000, 1000, 42, 0)
    got = pkg.Scheduler(problem).solve(count_visited=False)
    out = pi.check(problem, got)
    assert out["scheduled"] == 46250 and out["new_nodes"] == 9250  # the round-1 reference run of this seed


def test_c4_full_size_invariants(pkg):
    problem = pkg.Problem.synth(4, 100000, 1000, 42, 0)
    got = pkg.Scheduler(problem).solve(count_visited=False)
    out = pi.check(problem, got)
    assert out["scheduled"] == 100000 and out["new_nodes"] == 20000
    assert out["anti"] > 50 and out["zone"] > 100 and out["host"] > 50


@pytest.mark.parametrize("cfg,pods,types,seed", [(3, 3000, 1000, 7), (4, 3000, 1000, 7), (4, 2500, 1000, 11)])
def test_c3_c4_samples_match_oracle(pkg, oracle, cfg, pods, types, seed):
    problem = pkg.Problem.synth(cfg, pods, types, seed, 0)
    got = pkg.Scheduler(problem).solve()
    want = pkg.Result()
    oracle.solve(problem, want)
    assert np.array_equal(got.assign, want.assign)
    assert got.digest() == want.digest()

"""Regenerates tests/golden/encoding_digests.json: a digest (kh_encoded_digest) of everything the host encoder hands to the
C-ABI, for a corpus of problems. The encoder is the one piece of the product that is pure host code on the GPU path, so a
refactor / speed-up of it can be proven result-neutral on the CPU: same bytes in, same kernels, same results.
Run after any INTENDED change of the encoding (new catalog field, different dictionary order, ...):
    python tests/golden/make_encoding_digests.py
"""
import ctypes as C
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def corpus(pkg):
    import consolidation_answers as ca
    import known_answers as ka
    import itertools

    import fixtures as fx
    from fuzz_problems import random_problem

    def fresh():  # pod names / uids come from a process-wide counter in the fixtures: restart it so the corpus does not depend on test order
        fx._uid = itertools.count()
    for seed in range(300):
        fresh()
        yield f"fuzz-{seed}", pkg.Problem.from_dict(random_problem(seed)), ()
    for name, _, build in ka.CASES + ka.CPU_ONLY_CASES + ca.CASES + ca.CPU_ONLY_CASES:
        fresh()
        prob, _ = build()
        for i, pd in enumerate(prob["multi"] if "multi" in prob else [prob]):
            problem = pkg.Problem.from_dict(pd)
            yield f"ka-{name}-{i}", problem, ()
            n_cand = sum(1 for n in pd.get("nodes", []) if n.get("candidate"))
            for c in range(1, n_cand + 1):
                yield f"ka-{name}-{i}-cand{c}", problem, tuple(j for j, n in enumerate(pd["nodes"]) if n.get("candidate"))[:c]
    for cfg, pods, types, nodes in ((1, 100, 10, 0), (2, 3000, 500, 0), (3, 3000, 1000, 0), (4, 3000, 1000, 0), (5, 2000, 1000, 200)):
        problem = pkg.Problem.synth(cfg, pods, types, 42, nodes)
        yield f"synth-{cfg}", problem, ()
        if cfg == 5:
            for c in (1, 2, 17, 100, 199):
                yield f"synth-5-cand{c}", problem, tuple(range(c))


def digest(pkg, problem, candidates):
    arr = (C.c_int * max(1, len(candidates)))(*candidates)
    enc = pkg.lib().kh_encode(problem.ptr, arr, len(candidates))
    if not enc:
        return "refused: " + pkg.lib().kh_scheduler_error().decode()[:80]
    d = pkg.lib().kh_encoded_digest(enc)
    pkg.lib().kh_encoded_free(enc)
    return f"{d:016x}"


def main():
    import __graft_entry__ as g
    pkg = g.load_pkg()
    out = {name: digest(pkg, problem, cands) for name, problem, cands in corpus(pkg)}
    (ROOT / "tests" / "golden" / "encoding_digests.json").write_text(json.dumps(out, indent=0, sort_keys=True) + "\n")
    print(len(out), "digests written")


if __name__ == "__main__":
    main()

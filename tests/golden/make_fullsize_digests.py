"""Runs the CPU oracle ONCE on BASELINE.json's C3 / C4 at full size (seed 42) and writes the committed fixtures the GPU
tests compare against: tests/golden/fullsize/c<N>.json (digest, counts, wall time) + c<N>_assign.npz (the assignment vector,
so a mismatch can be located). Test infrastructure: only tests/ reads these.

  python tests/golden/make_fullsize_digests.py 3      # minutes
  python tests/golden/make_fullsize_digests.py 4      # minutes in fast mode
  python tests/golden/make_fullsize_digests.py 5      # C5: 5 000 nodes / 50 000 bound pods: candidate order, the multi-node binary
                                                      # search (command + probe trace) and single-node consolidation; ~7 minutes
"""
import hashlib, json, sys, time
from pathlib import Path
HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent)); sys.path.insert(0, str(HERE.parent.parent))
import numpy as np
from conftest import load_pkg
import oracle_lib

SIZES = {2: (2, 10000, 500), 3: (3, 50000, 1000), 4: (4, 100000, 1000)}

def main():
    k = load_pkg(); oracle = oracle_lib.load()
    oracle.lib.oracle_set_fast(1)  # result-neutral shortcuts (oracle.cc: FAST MODE); equality with the literal path: tests/test_oracle_fast_mode.py
    for cfg in [int(a) for a in sys.argv[1:]]:
        if cfg == 5:
            nodes, T = 5000, 1000
            problem = k.Problem.synth(5, nodes * 10, T, 42, nodes)
            t0 = time.time()
            multi = oracle.consolidate(problem)
            t1 = time.time()
            single = oracle.consolidate_single(problem)
            t2 = time.time()
            order, cost = oracle.rank_candidates(problem)
            out = {"config": 5, "nodes": nodes, "pods": nodes * 10, "types": T, "seed": 42,
                   "multi_node": {kk: multi[kk] for kk in ("action", "nodes_removed", "simulations", "options", "probes", "probe_actions")},
                   "single_node": {kk: single[kk] for kk in ("action", "node", "options", "simulations")},
                   "candidate_order_sha256": hashlib.sha256(np.asarray(order, dtype=np.int32).tobytes()).hexdigest(),
                   "candidate_cost_sha256": hashlib.sha256(np.asarray(cost, dtype=np.float64).tobytes()).hexdigest(),
                   "pods_repacked_multi_node": int(sum(10 * c for c in multi["probes"])),
                   "oracle_seconds_multi_node": round(t1 - t0, 1), "oracle_seconds_single_node": round(t2 - t1, 1)}
            (HERE / "fullsize" / "c5.json").write_text(json.dumps(out, indent=1) + "\n")
            print(json.dumps(out), flush=True)
            continue
        c, P, T = SIZES[cfg]
        problem = k.Problem.synth(c, P, T, 42, 0)
        want = k.Result()
        t0 = time.time()
        assert oracle.solve(problem, want) == 0, want.error
        dt = time.time() - t0
        a = np.asarray(want.assign, dtype=np.int32)
        info = want.new_node_info()
        out = {"config": cfg, "pods": P, "types": T, "seed": 42, "digest": int(want.digest()),
               "scheduled": int((a >= 0).sum()), "new_nodes": int(want.num_new_nodes),
               "assign_sha256": hashlib.sha256(a.tobytes()).hexdigest(),
               "nodes_visited": int(want.nodes_visited), "oracle_seconds": round(dt, 1),
               "oracle_pods_per_s": round(P / dt, 1)}
        (HERE / "fullsize" / f"c{cfg}.json").write_text(json.dumps(out, indent=1) + "\n")
        np.savez_compressed(HERE / "fullsize" / f"c{cfg}_assign.npz", assign=a, relax=np.asarray(want.relax_level, dtype=np.int32),
                            node_info=np.asarray(info))
        print(json.dumps(out), flush=True)

if __name__ == "__main__":
    main()

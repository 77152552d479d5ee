"""Ad-hoc: fast kernel paths (count_visited=False) against the generic step (count_visited=True) on the same problem."""
import sys
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from conftest import load_pkg
import numpy as np
k = load_pkg()
cases = [(2, 3000, 500), (3, 3000, 1000), (4, 3000, 1000), (3, 50000, 1000)]
for cfg, P, T in cases:
    p = k.Problem.synth(cfg, P, T, 42, 0)
    ref = k.Scheduler(p).solve(count_visited=True)
    try:
        got = k.Scheduler(p).solve(count_visited=False)
    except Exception as e:
        print(cfg, P, "ERROR", e); continue
    a, b = got.assign, ref.assign
    print(f"C{cfg} P={P}: equal={np.array_equal(a, b)} new {got.num_new_nodes} vs {ref.num_new_nodes} digest_equal={got.digest() == ref.digest()}", flush=True)
    if not np.array_equal(a, b):
        s = p.pod_summary()
        order = np.lexsort((np.arange(P), -s[:, 1], -s[:, 0]))
        for t, pod in enumerate(order):
            if a[pod] != b[pod]:
                print("  first diff at queue pos", t, "pod", pod, "fast", a[pod], "generic", b[pod], "summary", s[pod].tolist())
                lo = max(0, t - 3)
                print("  around:", [(int(q), int(a[q]), int(b[q])) for q in order[lo:t + 4]])
                break
        cg = np.bincount(a[a >= 0]); cr = np.bincount(b[b >= 0])
        print("  pods/node fast max", cg.max(), "generic max", cr.max(), "nodes fast", len(cg), "generic", len(cr))
    else:
        gi, ri = got.new_node_info(), ref.new_node_info()
        if not np.array_equal(gi, ri):
            bad = np.nonzero((gi != ri).any(axis=1))[0]
            print("  node info differs at", bad[:5], gi[bad[:3]].tolist(), ri[bad[:3]].tolist())

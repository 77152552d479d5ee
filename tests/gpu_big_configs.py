"""Ad-hoc: full-size BASELINE configs on the GPU (timings + basic invariants)."""
import sys, time
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from conftest import load_pkg
import numpy as np
k = load_pkg()
ALL = {2: (2, 10000, 500), 3: (3, 50000, 1000), 4: (4, 100000, 1000)}
which = [int(x) for x in sys.argv[1:]] or [3, 4]
for cfg, P, T in [ALL[w] for w in which]:
    t0 = time.time(); p = k.Problem.synth(cfg, P, T, 42, 0); t1 = time.time()
    rs = k.ResidentSolve(p); t2 = time.time()
    rs.set_count_visited(False); rs.load(); t3 = time.time()
    rs.run(flush_l2=True); tm = rs.timings(); res = rs.download(); t4 = time.time()
    a = res.assign
    info = res.new_node_info()
    print(f"C{cfg}: synth {t1-t0:.2f}s encode {t2-t1:.2f}s load {t3-t2:.2f}s run+dl {t4-t3:.2f}s dims {rs.dims}")
    print("   timings", {kk: round(v, 1) for kk, v in tm.items() if kk.endswith('_us')}, "scheduled", int((a >= 0).sum()), "new nodes", res.num_new_nodes,
          "pods/node max", int(info[:,1].max()) if len(info) else 0, "min options", int(info[:,2].min()) if len(info) else 0)
    us, _ = rs.solve_e2e(); print("   e2e ms", us/1000, "pods/s", P/(us*1e-6))

import sys, os, time
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from conftest import load_pkg
k = load_pkg()
cfg, P, T = int(sys.argv[1]), int(sys.argv[2]), 1000
p = k.Problem.synth(cfg, P, T, 42, 0)
rs = k.ResidentSolve(p); rs.set_count_visited(False); rs.load()
rs.run(flush_l2=True); tm = rs.timings(); res = rs.download()
print("threads", os.environ.get("KSCHED_PACK_THREADS"), "pack_ms", tm["pack_us"]/1000, "nodes", res.num_new_nodes)

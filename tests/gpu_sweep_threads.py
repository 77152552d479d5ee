"""Scratch driver (not a test): pack-kernel time per configuration x block size x level-run on/off, resident problem,
L2 flushed. Usage: python tests/gpu_sweep_threads.py [configs, default 2,3,4]"""
import os, sys
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from conftest import load_pkg
k = load_pkg()
SIZES = {2: (10000, 500), 3: (50000, 1000), 4: (100000, 1000)}
cfgs = [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else "2,3,4".split(","))]
for cfg in cfgs:
    P, T = SIZES[cfg]
    p = k.Problem.synth(cfg, P, T, 42, 0)
    rs = k.ResidentSolve(p); rs.set_count_visited(False); rs.load()
    base = None
    for lvl in (0, 1):
        for threads in (128, 256, 512):
            os.environ["KSCHED_PACK_THREADS"] = str(threads)
            if lvl: os.environ.pop("KSCHED_NO_LEVELRUN", None)
            else: os.environ["KSCHED_NO_LEVELRUN"] = "1"
            best = 1e30
            for _ in range(3):
                rs.run(flush_l2=True); best = min(best, rs.timings()["pack_us"])
            res = rs.download()
            d = res.digest() if hasattr(res, "digest") else None
            if base is None: base = d
            print(f"C{cfg} level_run={lvl} threads={threads} pack_ms={best/1000:.3f} pods/s={P/best*1e6:.0f} nodes={res.num_new_nodes} same_digest={d == base}", flush=True)

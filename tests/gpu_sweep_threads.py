"""Scratch driver (not a test): pack-kernel time per configuration with the class-run modes switched on and off
(KSCHED_NO_LEVELRUN / KSCHED_NO_MASKRUN) and per block size, resident problem, L2 flushed.
Usage: python tests/gpu_sweep_threads.py [configs, default 2,3,4] [threads, default 128] [final]"""
import os, sys
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from conftest import load_pkg
k = load_pkg()
SIZES = {2: (10000, 500), 3: (50000, 1000), 4: (100000, 1000)}
cfgs = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "2,3,4").split(",")]
threads_list = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "128").split(",")]
for cfg in cfgs:
    P, T = SIZES[cfg]
    p = k.Problem.synth(cfg, P, T, 42, 0)
    rs = k.ResidentSolve(p); rs.set_count_visited(False); rs.load()
    base = None
    modes = ("KSCHED_NO_LEVELWARP", "") if len(sys.argv) > 3 and sys.argv[3] == "final" else ("KSCHED_NO_LEVELRUN,KSCHED_NO_MASKRUN", "KSCHED_NO_MASKRUN", "KSCHED_NO_LEVELWARP", "")
    for off in modes:
        for threads in threads_list:
            os.environ["KSCHED_PACK_THREADS"] = str(threads)
            for v in ("KSCHED_NO_LEVELRUN", "KSCHED_NO_MASKRUN", "KSCHED_NO_LEVELWARP"):
                os.environ.pop(v, None)
            for v in off.split(","):
                if v: os.environ[v] = "1"
            best = 1e30
            for _ in range(3):
                rs.run(flush_l2=True); best = min(best, rs.timings()["pack_us"])
            res = rs.download()
            d = res.digest()
            if base is None: base = d
            print(f"C{cfg} off=[{off}] threads={threads} pack_ms={best/1000:.3f} pods/s={P/best*1e6:.0f} nodes={res.num_new_nodes} same_digest={d == base}", flush=True)

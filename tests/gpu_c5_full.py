"""Ad-hoc: BASELINE config 5 at full size on one GPU (5 000 nodes, 50 000 bound pods, 1 000 types): the multi-node binary
search through kh_consolidate, per-probe wall time and the last probe's device phases."""
import json, os, sys, time
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from conftest import load_pkg
k = load_pkg()
nodes = int(os.environ.get("KS_NODES", 5000)); pods = nodes * 10
t0 = time.time(); p = k.Problem.synth(5, pods, 1000, 42, nodes); print("synth s", round(time.time() - t0, 2), flush=True)
mnc = k.MultiNodeConsolidation(p)
mnc.probe(2)
out = []
for count in [int(x) for x in os.environ.get("KS_PROBES", "2501,1251,626,313,4000").split(",")]:
    if count > nodes: continue
    t0 = time.perf_counter(); r = mnc.probe(count); dt = time.perf_counter() - t0
    _t = k.Timings(); k.lib().kh_gpu_timings(__import__('ctypes').byref(_t)); tm = _t.as_dict()
    out.append({"count": count, "action": r[0], "n_options": len(r[1]), "wall_ms": round(dt * 1e3, 2),
                **{kk: round(v / 1e3, 3) for kk, v in tm.items() if kk.endswith("_us")}})
    print(json.dumps(out[-1]), flush=True)
t0 = time.perf_counter(); sn = k.SingleNodeConsolidation(p).compute_command(batch=64); dt = time.perf_counter() - t0
print(json.dumps({"single_node_s": round(dt, 3), **{kk: v for kk, v in sn.items() if kk != "options"}}), flush=True)
cs = mnc.session()
t0 = time.perf_counter(); sw = cs.probe_sets([[i] for i in range(min(512, cs.n_candidates))], False); dt = time.perf_counter() - t0
print(json.dumps({"single_node_sweep_512_s": round(dt, 3), "hits": sum(1 for a, _ in sw if a)}), flush=True)
t0 = time.perf_counter(); full = mnc.first_n_node_consolidation_option(); dt = time.perf_counter() - t0
print(json.dumps({"search_s": round(dt, 3), **{kk: v for kk, v in full.items() if kk != "options"}, "n_options": len(full["options"])}), flush=True)

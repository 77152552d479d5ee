"""Ad-hoc GPU debugging driver (not a pytest file): compares oracle vs GPU on small configs and prints diffs."""
import sys, time
sys.path.insert(0, "tests")
from conftest import load_pkg
import oracle_lib

k = load_pkg(); o = oracle_lib.load()
print("devices", k.device_count())
cases = [(1,100,10,0,42),(2,1000,500,0,42),(3,1500,1000,0,42),(4,1500,1000,0,42),(5,600,1000,60,42)]
for cfg,P,T,N,seed in cases:
    p = k.Problem.synth(cfg,P,T,seed,N)
    want = k.Result(); o.solve(p, want, [0,1,2] if cfg==5 else [])
    t=time.time()
    try:
        got = k.Scheduler(p).solve([0,1,2] if cfg==5 else [])
    except Exception as e:
        print(cfg, "GPU ERROR", e); continue
    dt=time.time()-t
    w,g = want.to_dict(), got.to_dict()
    same = g["assign"]==w["assign"]
    print(cfg, "assign_equal", same, "new", len(g["newNodes"]), len(w["newNodes"]), "visited", got.nodes_visited, want.nodes_visited,
          "adds", got.add_calls, want.add_calls, "%.3fs"%dt)
    if not same:
        for i,(a,b) in enumerate(zip(g["assign"],w["assign"])):
            if a!=b: print("  first diff pod",i,"gpu",a,"oracle",b); break
    for i,(a,b) in enumerate(zip(g["newNodes"],w["newNodes"])):
        if a["options"]!=b["options"] or a["requests"]!=b["requests"] or a["pods"]!=b["pods"]:
            print("  node",i,"differs: opts",len(a["options"]),len(b["options"]),"req",a["requests"],b["requests"],"pods",a["pods"][:5],b["pods"][:5]); break
        breq = {kk:v for kk,v in b["requirements"].items() if kk!="node.kubernetes.io/instance-type"}
        if a["requirements"]!=breq:
            print("  node",i,"reqs differ",a["requirements"],breq); break

"""Ad-hoc: categorise fuzz divergences (not a pytest file)."""
import sys, json, collections
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from conftest import load_pkg
import oracle_lib
from fuzz_problems import random_problem
k = load_pkg(); o = oracle_lib.load()
seeds = [int(x) for x in sys.argv[1:]] or range(300)
skips = collections.Counter(); fails = []
for seed in seeds:
    prob = random_problem(seed)
    p = k.Problem.from_dict(prob)
    want = k.Result(); o.solve(p, want)
    try:
        got = k.Scheduler(p).solve()
    except k.KschedError as e:
        skips[str(e)[:90]] += 1; continue
    g, w = got.to_dict(), want.to_dict()
    why = None
    if g["assign"] != w["assign"]:
        i = next(i for i,(a,b) in enumerate(zip(g["assign"], w["assign"])) if a != b)
        why = f"assign pod {i}: gpu {g['assign'][i]} oracle {w['assign'][i]}"
        pd = prob["pods"][i] if i < len(prob["pods"]) else None
        feats = [kk for kk in ("nodeSelector","nodeAffinity","tolerations","ports","topologySpreadConstraints","podAntiAffinity","podAffinity") if pd and kk in pd]
        why += f" feats={feats} detail={ {kk: pd[kk] for kk in feats} if pd else None}"
    elif g["relax"] != w["relax"]: why = "relax"
    elif g["existing"] != w["existing"]: why = "existing pods order"
    else:
        for n,(a,b) in enumerate(zip(g["newNodes"], w["newNodes"])):
            for f in ("provisioner","pods","options","requests"):
                if a[f] != b[f]: why = f"node {n} {f}: {str(a[f])[:80]} vs {str(b[f])[:80]}"; break
            if why: break
            br = {kk: v for kk, v in b["requirements"].items() if kk != "node.kubernetes.io/instance-type"}
            if a["requirements"] != br: why = f"node {n} reqs: {a['requirements']} vs {br}"; break
        if not why and got.nodes_visited != want.nodes_visited: why = f"nodes_visited {got.nodes_visited} vs {want.nodes_visited}"
    if why: fails.append((seed, why))
print("SKIPS"); [print(" ", n, s) for s, n in skips.most_common()]
print("FAILS", len(fails))
for s, w in fails: print(" ", s, w[:400])

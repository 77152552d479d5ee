"""Size-independent properties every valid Scheduler.Solve result has, checked from the inputs alone (no oracle, no
kernel code): used on the full-size BASELINE configurations where the CPU oracle would take minutes.

 * a pod is assigned to at most one node, every new node holds >= 1 pod, `pods` lists and `assign` agree;
 * Node.Add keeps at least one instance type that fits (node.go:92-101): nOptions >= 1, and the summed cpu / memory /
   pod count of a node equals its recorded requests (no daemonsets in the synthetic configs);
 * required hostname anti-affinity on the own app: at most one pod of the app per node (topologygroup.go:221-243);
 * DoNotSchedule topology spread: at the end of the solve the per-domain counts of the app differ by at most maxSkew
   over the zones (all three zones are always eligible in the synthetic configs), and a node holds at most maxSkew pods
   of the app for hostname spread (min over hostnames is 0, topologygroup.go:186-188).
"""
import re

import numpy as np


def check(problem, result, n_zones=3):
    summ = problem.pod_summary()
    d = result.to_dict(brief=True)
    assign = np.asarray(d["assign"], dtype=np.int64)
    n_existing = len(d["existing"])
    nodes = d["newNodes"]
    P = len(assign)
    # pods lists <-> assign
    seen = np.full(P, -1, dtype=np.int64)
    for i, n in enumerate(nodes):
        assert len(n["pods"]) >= 1, f"new node {i} is empty"
        for p in n["pods"]:
            assert seen[p] == -1, f"pod {p} on two nodes"
            seen[p] = n_existing + i
        assert n["nOptions"] >= 1, f"new node {i} has no instance type left"
    for e in d["existing"]:
        for p in e["pods"]:
            assert seen[p] == -1
            seen[p] = e["node"] if "node" in e else -2
    new_mask = assign >= n_existing
    assert np.array_equal(seen[new_mask], assign[new_mask])
    assert np.all((seen == -1) == (assign == -1))
    # requests
    for i, n in enumerate(nodes):
        pods = np.asarray(n["pods"])
        assert n["requests"].get("cpu", 0) == int(summ[pods, 0].sum()), f"node {i} cpu"
        assert n["requests"].get("memory", 0) == int(summ[pods, 1].sum()), f"node {i} memory"
        assert n["requests"].get("pods", 0) == 1000 * len(pods), f"node {i} pods"
    # topology
    zone_of = {}
    for i, n in enumerate(nodes):
        z = n["requirements"].get("topology.kubernetes.io/zone", "")
        m = re.fullmatch(r"In \[([^\],]+)\]", z)
        zone_of[n_existing + i] = m.group(1) if m else None
    apps = {}
    for p in np.nonzero(assign >= 0)[0]:
        a = int(summ[p, 2])
        if a >= 0 and (summ[p, 3] or summ[p, 4] or summ[p, 5]):
            apps.setdefault(a, []).append(int(p))
    checked = {"anti": 0, "zone": 0, "host": 0}
    for a, pods in apps.items():
        pods = np.asarray(pods)
        node_ids = assign[pods]
        per_node = np.unique(node_ids, return_counts=True)[1]
        if summ[pods[0], 3]:
            assert per_node.max() == 1, f"app {a}: anti-affinity violated"
            checked["anti"] += 1
        if summ[pods[0], 5]:
            assert per_node.max() <= summ[pods[0], 5], f"app {a}: hostname spread violated"
            checked["host"] += 1
        if summ[pods[0], 4]:
            zones = [zone_of.get(int(n)) for n in node_ids]
            assert all(z is not None for z in zones), f"app {a}: a zone-spread pod sits on a node without a single zone"
            counts = {}
            for z in zones:
                counts[z] = counts.get(z, 0) + 1
            cs = list(counts.values()) + [0] * (n_zones - len(counts))
            assert max(cs) - min(cs) <= summ[pods[0], 4], f"app {a}: zone skew {cs}"
            checked["zone"] += 1
    return {"scheduled": int((assign >= 0).sum()), "new_nodes": len(nodes), **checked}

"""Seeded random problem generator for differential testing (GPU vs oracle): mixes every feature of the hot path —
requests, node selectors, required / preferred node affinity (relaxation), taints / tolerations, host ports, provisioner
limits, several provisioners, daemonsets, existing nodes with bound pods, zonal / hostname / capacity-type spread,
pod affinity and anti-affinity (required and preferred)."""
import random

import fixtures as fx
from fixtures import ARCH, CAPACITY_TYPE, HOSTNAME, OS, ZONE, ZONES

CPUS = ["100m", "250m", "500m", "1", "1500m", "3"]
MEMS = ["100Mi", "256Mi", "512Mi", "1Gi", "2Gi", "4Gi"]


def random_problem(seed):
    rng = random.Random(seed)
    kind = rng.choice(["default", "fake", "assorted"])
    if kind == "default":
        its = fx.default_instance_types()
    elif kind == "fake":
        its = fx.fake_instance_types(rng.choice([5, 12, 70]))
    else:
        its = []
        for cpu in (1, 2, 4, 8, 16):
            for mem in (2, 8, 32):
                for zone in ZONES:
                    for ct in ("spot", "on-demand"):
                        for arch in ("amd64", "arm64"):
                            name = f"{cpu}c-{mem}g-{arch}-{zone}-{ct}"
                            res = {"cpu": str(cpu), "memory": f"{mem}Gi", "pods": str(rng.choice([5, 20]))}
                            its.append(fx.instance_type(name, res, offerings=[{"capacityType": ct, "zone": zone, "price": fx.price_from_resources(res),
                                                                                 "available": rng.random() > 0.05}], arch=arch, oses=("linux",)))
        rng.shuffle(its)
        its = its[: rng.choice([40, 90, 180])]
    provisioners = []
    n_prov = rng.choice([1, 1, 2, 3])
    for i in range(n_prov):
        pr = {"name": f"prov-{i}", "weight": rng.choice([0, 10, 10, 50])}
        if rng.random() < 0.4:
            pr["requirements"] = [{"key": ZONE, "operator": "In", "values": rng.sample(ZONES, rng.choice([1, 2, 3]))}]
        if rng.random() < 0.3:
            pr.setdefault("requirements", []).append({"key": ARCH, "operator": rng.choice(["In", "NotIn"]), "values": [rng.choice(["amd64", "arm64"])]})
        if rng.random() < 0.3:
            pr["labels"] = {"team": rng.choice(["a", "b"])}
        if rng.random() < 0.3:
            pr["taints"] = [{"key": "dedicated", "value": rng.choice(["x", "y"]), "effect": rng.choice(["NoSchedule", "PreferNoSchedule"])}]
        r = rng.random()
        if r < 0.3:
            pr["limits"] = {"cpu": str(rng.choice([4, 16, 64]))}
        elif r < 0.5:
            pr["limits"] = {}
        provisioners.append(pr)
    apps = [f"app-{i}" for i in range(rng.choice([1, 2, 4]))]
    pods = []
    n_pods = rng.choice([3, 8, 20, 45])
    for i in range(n_pods):
        app = rng.choice(apps)
        labels = {"app": app}
        p = fx.pod({"cpu": rng.choice(CPUS), "memory": rng.choice(MEMS)} if rng.random() < 0.9 else {"cpu": rng.choice(CPUS)}, labels=labels)
        p["creationTimestamp"] = rng.choice([0, 0, 5])
        r = rng.random()
        if r < 0.15:
            p["nodeSelector"] = {ZONE: rng.choice(ZONES + ["test-zone-9"])}
        elif r < 0.25:
            p["nodeSelector"] = {ARCH: rng.choice(["amd64", "arm64"])}
        elif r < 0.32:
            p["nodeSelector"] = {CAPACITY_TYPE: rng.choice(["spot", "on-demand"])}
        elif r < 0.38:
            p["nodeSelector"] = {"team": rng.choice(["a", "b", "c"])}
        r = rng.random()
        if r < 0.12:
            p["nodeAffinity"] = {"required": [[{"key": ZONE, "operator": rng.choice(["In", "NotIn"]), "values": rng.sample(ZONES, rng.choice([1, 2]))}]
                                              for _ in range(rng.choice([1, 2, 3]))]}
        elif r < 0.22:
            p["nodeAffinity"] = {"preferred": [{"weight": rng.choice([1, 50, 100]), "terms": [{"key": rng.choice([ZONE, ARCH]), "operator": "In",
                                                                                              "values": [rng.choice(ZONES + ["amd64", "nope"])]}]}
                                               for _ in range(rng.choice([1, 2]))]}
        if rng.random() < 0.35:
            p["tolerations"] = [rng.choice([{"key": "dedicated", "operator": "Exists"}, {"key": "dedicated", "operator": "Equal", "value": "x", "effect": "NoSchedule"},
                                            {"operator": "Exists"}])]
        if rng.random() < 0.08:
            p["ports"] = [{"hostPort": rng.choice([80, 443]), "hostIP": rng.choice(["", "1.2.3.4", "0.0.0.0"]), "protocol": rng.choice(["TCP", "UDP"])}]
        r = rng.random()
        sel = {"app": rng.choice(apps)} if rng.random() < 0.3 else labels
        if r < 0.2:
            p["topologySpreadConstraints"] = [fx.spread(rng.choice([ZONE, ZONE, HOSTNAME, CAPACITY_TYPE]), sel, max_skew=rng.choice([1, 1, 2]),
                                                        when=rng.choice(["DoNotSchedule", "DoNotSchedule", "ScheduleAnyway"]))]
        elif r < 0.3:
            p["podAntiAffinity"] = {"required": [fx.affinity_term(rng.choice([HOSTNAME, HOSTNAME, ZONE]), sel)]}
        elif r < 0.38:
            p["podAffinity"] = {"required": [fx.affinity_term(rng.choice([HOSTNAME, ZONE]), sel)]}
        elif r < 0.44:
            p["podAntiAffinity"] = {"preferred": [{"weight": 10, "term": fx.affinity_term(rng.choice([HOSTNAME, ZONE]), sel)}]}
        elif r < 0.5:
            p["podAffinity"] = {"preferred": [{"weight": 10, "term": fx.affinity_term(rng.choice([HOSTNAME, ZONE]), sel)}]}
        pods.append(p)
    nodes = []
    for i in range(rng.choice([0, 0, 1, 3, 6])):
        it = rng.choice(its)
        bound = [fx.pod({"cpu": rng.choice(CPUS)}, labels={"app": rng.choice(apps)}, nodeName=f"node-{i}") for _ in range(rng.choice([0, 1, 2]))]
        if bound and rng.random() < 0.2:
            bound[0]["podAntiAffinity"] = {"required": [fx.affinity_term(rng.choice([HOSTNAME, ZONE]), {"app": rng.choice(apps)})]}
        alloc = {k: v for k, v in it["capacity"].items()}
        nodes.append(fx.state_node(f"node-{i}", it["name"], zone=rng.choice(ZONES), capacity_type=rng.choice(["spot", "on-demand"]), allocatable=alloc,
                                   pods_=bound, initialized=rng.random() < 0.8, provisioner_name=rng.choice(provisioners)["name"] if rng.random() < 0.9 else "",
                                   taints=[{"key": "dedicated", "value": "x", "effect": "NoSchedule"}] if rng.random() < 0.15 else [],
                                   markedForDeletion=rng.random() < 0.1))
    daemons = []
    if rng.random() < 0.3:
        daemons.append({"name": "ds", "uid": "ds", "requests": {"cpu": "100m", "memory": "64Mi"},
                        "tolerations": [{"operator": "Exists"}] if rng.random() < 0.5 else []})
    return fx.problem(pods, instance_types=its, provisioners=provisioners, nodes=nodes, daemonSetPods=daemons)


def random_problem_with_bounds(seed):
    """random_problem(seed) plus the requirement forms the device carries in region form (include/ksched.h: ksched_key_regions):
    Gt / Lt on provisioners and pods (node affinity, required and preferred), NotIn / Exists / DoesNotExist / Gt / Lt requirements
    on instance types. A separate random stream, so random_problem's corpus stays what it was."""
    prob = random_problem(seed)
    rng = random.Random(seed * 7919 + 13)
    its = prob["instanceTypes"]
    ints = sorted({int(r["values"][0]) for it in its for r in it["requirements"] if r["key"] == "integer"})
    thresholds = [str(rng.choice(ints) + rng.choice([-1, 0, 0, 1])) for _ in range(rng.choice([1, 2, 3]))] if ints else ["1"]
    def bound():
        return {"key": "integer", "operator": rng.choice(["Gt", "Lt"]), "values": [rng.choice(thresholds)]}
    for pr in prob["provisioners"]:
        if rng.random() < 0.5:
            pr.setdefault("requirements", []).append(bound())
        if rng.random() < 0.15:
            pr.setdefault("requirements", []).append({"key": "integer", "operator": "NotIn", "values": [str(rng.choice(ints or [1]))]})
    for p in prob["pods"]:
        r = rng.random()
        if r < 0.25:
            term = [bound()] + ([bound()] if rng.random() < 0.3 else [])
            na = p.setdefault("nodeAffinity", {})
            if "required" in na:
                for t in na["required"]:
                    t.extend(term)
            else:
                na["required"] = [term]
        elif r < 0.35:
            p.setdefault("nodeAffinity", {}).setdefault("preferred", []).append({"weight": rng.choice([1, 70]), "terms": [bound()]})
        elif r < 0.42:
            p.setdefault("nodeAffinity", {}).setdefault("required", [[]])[0].append(
                {"key": "tier", "operator": rng.choice(["In", "NotIn", "Exists", "DoesNotExist"]), "values": [rng.choice(["gold", "silver"])]})
    for it in its:
        r = rng.random()
        if r < 0.10:
            it["requirements"].append({"key": "tier", "operator": "NotIn", "values": [rng.choice(["gold", "silver"])]})
        elif r < 0.18:
            it["requirements"].append({"key": "tier", "operator": "Exists", "values": []})
        elif r < 0.26:
            it["requirements"].append({"key": "tier", "operator": "In", "values": [rng.choice(["gold", "silver"])]})
        elif r < 0.30:
            it["requirements"].append({"key": "tier", "operator": "DoesNotExist", "values": []})
        elif r < 0.34:
            it["requirements"] = [q for q in it["requirements"] if q["key"] != "integer"] + [bound()]
    for t in [q for p in prob["pods"] for q in p.get("nodeAffinity", {}).get("required", [])]:
        for q in t:
            if q["operator"] in ("Exists", "DoesNotExist"):
                q["values"] = []
    prob["wellKnownLabels"] = list(prob["wellKnownLabels"]) + (["tier"] if rng.random() < 0.5 else [])
    # CSI volume limits (existingnode.go:88-96): limits on some nodes, private and shared claims on pending and bound pods
    if prob.get("nodes") and rng.random() < 0.6:
        drivers = ["ebs", "efs"]
        def claims(owner):
            out = []
            for _ in range(rng.choice([0, 0, 1, 2, 3])):
                d = rng.choice(drivers + [""])
                out.append({"driver": d, "pvc": rng.choice([f"default/shared-{rng.randrange(4)}", f"default/{owner}-{len(out)}"])})
            return out
        for n in prob["nodes"]:
            if rng.random() < 0.7:
                n["volumeLimits"] = {d: rng.choice([0, 1, 2, 4]) for d in drivers if rng.random() < 0.7}
            for bp in n["pods"]:
                bp["volumes"] = claims(bp["name"])
        for p in prob["pods"]:
            if rng.random() < 0.5:
                p["volumes"] = claims(p["name"])
    return prob

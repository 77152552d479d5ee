"""Problem builders that restate the reference's test fixtures (semantics only):
   pkg/cloudprovider/fake/cloudprovider.go:116-158  default instance types
   pkg/cloudprovider/fake/instancetype.go:48-187   NewInstanceType / InstanceTypes(n) / prices
   pkg/test/pods.go:62-111, pkg/test/provisioner.go:62-64  pod / provisioner defaults
and a `launch` helper restating how the fake provider picks the instance type of a solved node
(fake/cloudprovider.go:74-84: cheapest available offering compatible with the node requirements)."""
import itertools

ZONES = ["test-zone-1", "test-zone-2", "test-zone-3"]
GI = 1024 ** 3
WELL_KNOWN_EXTRA = ["size", "special", "integer"]
HOSTNAME = "kubernetes.io/hostname"
ZONE = "topology.kubernetes.io/zone"
ARCH = "kubernetes.io/arch"
OS = "kubernetes.io/os"
CAPACITY_TYPE = "karpenter.sh/capacity-type"
INSTANCE_TYPE = "node.kubernetes.io/instance-type"
PROVISIONER_NAME = "karpenter.sh/provisioner-name"
INITIALIZED = "karpenter.sh/initialized"


def _qty_float(q):
    """approximate float of a quantity string, for priceFromResources only"""
    suffixes = {"Ki": 2 ** 10, "Mi": 2 ** 20, "Gi": 2 ** 30, "Ti": 2 ** 40, "m": 1e-3, "k": 1e3, "M": 1e6, "G": 1e9, "T": 1e12}
    q = str(q)
    for s in sorted(suffixes, key=len, reverse=True):
        if q.endswith(s):
            return float(q[: -len(s)]) * suffixes[s]
    return float(q)


def price_from_resources(resources):
    price = 0.0
    for k, v in resources.items():
        if k == "cpu":
            price += 0.1 * _qty_float(v)
        elif k == "memory":
            price += 0.1 * _qty_float(v) / 1e9
        elif k in ("fake.com/vendor-a", "fake.com/vendor-b"):
            price += 1.0
    return price


def instance_type(name, resources=None, offerings=None, arch="amd64", oses=("darwin", "linux", "windows")):
    res = dict(resources or {})
    res.setdefault("cpu", "4")
    res.setdefault("memory", "4Gi")
    res.setdefault("pods", "5")
    if offerings is None:
        p = price_from_resources(res)
        offerings = [{"capacityType": ct, "zone": z, "price": p, "available": True}
                     for ct, z in (("spot", ZONES[0]), ("spot", ZONES[1]), ("on-demand", ZONES[0]), ("on-demand", ZONES[1]), ("on-demand", ZONES[2]))]
    avail = [o for o in offerings if o.get("available", True)]
    large = _qty_float(res["cpu"]) > 4 and _qty_float(res["memory"]) > 8 * GI
    reqs = [
        {"key": INSTANCE_TYPE, "operator": "In", "values": [name]},
        {"key": ARCH, "operator": "In", "values": [arch]},
        {"key": OS, "operator": "In", "values": sorted(oses)},
        {"key": ZONE, "operator": "In", "values": sorted({o["zone"] for o in avail})},
        {"key": CAPACITY_TYPE, "operator": "In", "values": sorted({o["capacityType"] for o in avail})},
        {"key": "size", "operator": "In", "values": ["large" if large else "small"]},
        {"key": "special", "operator": "In", "values": ["optional"]} if large else {"key": "special", "operator": "DoesNotExist"},
        {"key": "integer", "operator": "In", "values": [str(int(_qty_float(res["cpu"])))]},
    ]
    return {"name": name, "requirements": reqs, "offerings": offerings, "capacity": res,
            "overhead": {"kubeReserved": {"cpu": "100m", "memory": "10Mi"}}}


def default_instance_types():
    return [
        instance_type("default-instance-type"),
        instance_type("small-instance-type", {"cpu": "2", "memory": "2Gi"}),
        instance_type("gpu-vendor-instance-type", {"fake.com/vendor-a": "2"}),
        instance_type("gpu-vendor-b-instance-type", {"fake.com/vendor-b": "2"}),
        instance_type("arm-instance-type", {"cpu": "16", "memory": "128Gi"}, arch="arm64", oses=("ios", "linux", "windows", "darwin")),
        instance_type("single-pod-instance-type", {"pods": "1"}),
    ]


def fake_instance_types(total):
    return [instance_type(f"fake-it-{i}", {"cpu": str(i + 1), "memory": f"{(i + 1) * 2}Gi", "pods": str((i + 1) * 10)}) for i in range(total)]


_uid = itertools.count()


def pod(requests=None, **kw):
    i = next(_uid)
    p = {"name": f"pod-{i:05d}", "uid": f"uid-{i:05d}"}
    if requests is not None:
        p["requests"] = requests
    p.update(kw)
    return p


def pods(n, **kw):
    return [pod(**kw) for _ in range(n)]


def provisioner(name="default", **kw):
    pr = {"name": name, "limits": {"cpu": "2000"}}
    pr.update(kw)
    return pr


def problem(pods_, instance_types=None, provisioners=None, **kw):
    d = {"wellKnownLabels": WELL_KNOWN_EXTRA,
         "instanceTypes": instance_types if instance_types is not None else default_instance_types(),
         "provisioners": provisioners if provisioners is not None else [provisioner()],
         "pods": pods_}
    d.update(kw)
    return d


def spread(key, labels, max_skew=1, when="DoNotSchedule"):
    return {"maxSkew": max_skew, "topologyKey": key, "whenUnsatisfiable": when, "labelSelector": {"matchLabels": labels}}


def affinity_term(key, labels, namespaces=None):
    t = {"topologyKey": key, "labelSelector": {"matchLabels": labels}}
    if namespaces:
        t["namespaces"] = namespaces
    return t


def state_node(name, instance_type_name="default-instance-type", zone=ZONES[0], capacity_type="on-demand", allocatable=None, pods_=None,
               initialized=True, provisioner_name="default", labels=None, **kw):
    lab = {PROVISIONER_NAME: provisioner_name, INSTANCE_TYPE: instance_type_name, ZONE: zone, CAPACITY_TYPE: capacity_type, HOSTNAME: name,
           ARCH: "amd64", OS: "linux"}
    if initialized:
        lab[INITIALIZED] = "true"
    lab.update(labels or {})
    n = {"name": name, "labels": lab, "allocatable": allocatable or {"cpu": "3900m", "memory": "4086Mi", "pods": "5"}, "pods": pods_ or []}
    n.update(kw)
    return n


# ---------------------------------------------------------------- result helpers
def _req_has(req_str, value):
    """requirement rendered as 'In [a b]' / 'NotIn [a]' / 'Exists []' / 'DoesNotExist []'"""
    if req_str is None:
        return True
    op, _, rest = req_str.partition(" ")
    vals = rest.strip()[1:-1].split() if rest.strip() != "[]" else []
    if op == "In":
        return value in vals
    if op == "NotIn":
        return value not in vals
    return op == "Exists"


def launch(problem_dict, new_node):
    """fake CloudProvider.Create: the cheapest surviving option under the node's zone / capacity-type requirements."""
    its = problem_dict["instanceTypes"]
    zr, cr = new_node["requirements"].get(ZONE), new_node["requirements"].get(CAPACITY_TYPE)
    best = None
    for idx in new_node["options"]:
        it = its[idx]
        offers = [o for o in it["offerings"] if o.get("available", True) and _req_has(zr, o["zone"]) and _req_has(cr, o["capacityType"])]
        if not offers:
            continue
        cheapest = min(offers, key=lambda o: o["price"])
        if best is None or cheapest["price"] < best[0]:
            best = (cheapest["price"], it["name"], cheapest["zone"])
    return best


def node_of(result_dict, pod_index):
    return result_dict["assign"][pod_index]


def skew(problem_dict, result_dict, key, pod_indices=None):
    """multiset of per-domain pod counts (ExpectSkew, pkg/test/expectations/expectations.go:335-359) over new nodes"""
    counts = {}
    n_existing = len(result_dict["existing"])
    for i, a in enumerate(result_dict["assign"]):
        if a < 0 or (pod_indices is not None and i not in pod_indices):
            continue
        if key == HOSTNAME:
            dom = a
        elif a < n_existing:
            dom = problem_dict["nodes"][result_dict["existing"][a]["node"]]["labels"].get(key)
        else:
            node = result_dict["newNodes"][a - n_existing]
            req = node["requirements"].get(key, "")
            op, _, rest = req.partition(" ")
            vals = rest.strip()[1:-1].split()
            if op == "In" and len(vals) == 1:
                dom = vals[0]
            elif key == ZONE:
                dom = launch(problem_dict, node)[2]
            else:
                dom = tuple(vals)
        counts[dom] = counts.get(dom, 0) + 1
    return sorted(counts.values())

"""Known answers for deprovisioning candidate ranking (candidateNodes / sortAndFilterCandidates,
deprovisioning/helpers.go:125-165,171-249,275-287,339-366, consolidation.go:85-118, pdblimits.go:55-68) and single-node
consolidation (singlenodeconsolidation.go:43-84), restated from pkg/controllers/deprovisioning/suite_test.go."""
import consolidation_answers as ca
import fixtures as fx

CASES = []


def case(ref):
    def deco(fn):
        CASES.append((fn.__name__, ref, fn))
        return fn
    return deco


def _node(name, pods_, **kw):
    its = kw.pop("its")
    it = ca.on_demand_by_price(its)[-1]
    of = it["offerings"][0]
    return fx.state_node(name, it["name"], zone=of["zone"], capacity_type=of["capacityType"], allocatable={"cpu": "32", "memory": "64Gi", "pods": "100"},
                         pods_=pods_, **kw)


def _problem(nodes, its, **kw):
    prov = fx.provisioner(consolidationEnabled=True, **kw.pop("provisioner", {}))
    return fx.problem([], instance_types=its, provisioners=[prov], nodes=nodes, deriveCandidates=True, **kw)


@case("suite_test.go:818-871 (GetPodEvictionCost)")
def pod_eviction_cost_table():
    its = ca.assorted()
    mk = lambda i, **kw: _node(f"n{i}", [fx.pod({"cpu": "1"}, nodeName=f"n{i}", **kw)], its=its)
    nodes = [mk(0), mk(1, deletionCost=100), mk(2, deletionCost=-100), mk(3, deletionCost=101), mk(4, deletionCost=99), mk(5, priority=1), mk(6, priority=-1),
             mk(7, deletionCost=2147483647, priority=1000000000), mk(8, deletionCost=-2147483647, priority=-2147483648)]
    prob = _problem(nodes, its)

    def check(order, cost):
        c = dict(zip(order, cost))
        assert c[0] == 1.0
        assert c[1] > 1.0 > c[2]
        assert c[3] > c[1] > c[4]
        assert c[5] > 1.0 > c[6]
        assert c[7] == 10.0 and c[8] == -10.0            # clamped to [-10, 10]
        assert c[1] == 1.0 + 100 / 2 ** 27 and c[5] == 1.0 + 1 / 2 ** 25
        assert cost == sorted(cost)
    return prob, check


@case("suite_test.go:1746-1823 (lifetime remaining scales the disruption cost)")
def lifetime_remaining_scales_the_cost():
    its = ca.assorted()
    n1 = _node("node1", [fx.pod({"cpu": "1"}, nodeName="node1"), fx.pod({"cpu": "1"}, nodeName="node1")], its=its, creationTimestamp=100.0)
    n2 = _node("node2", [fx.pod({"cpu": "1"}, nodeName="node2")], its=its, creationTimestamp=102.0)
    prob = _problem([n1, n2], its, provisioner={"ttlSecondsUntilExpired": 3}, now=102.5)

    def check(order, cost):
        # node1 has more pods but only 1/6 of its lifetime left: 2 * (0.5 / 3) < 1 * (2.5 / 3)
        assert order == [0, 1]
        assert cost[0] == 2.0 * ((3.0 - 2.5) / 3.0) and cost[1] == 1.0 * ((3.0 - 0.5) / 3.0)
    return prob, check


@case("suite_test.go:930-1074,1075-1154,1497-1745 (PDBs, do-not-evict, do-not-consolidate, filters of candidateNodes)")
def filters():
    its = ca.assorted()
    app = {"app": "test"}
    nodes = [
        _node("ok", [fx.pod({"cpu": "1"}, nodeName="ok")], its=its),
        _node("pdb", [fx.pod({"cpu": "1"}, nodeName="pdb", labels=app)], its=its),                               # PDB with 0 disruptions allowed
        _node("pdb-other-ns", [fx.pod({"cpu": "1"}, nodeName="pdb-other-ns", labels=app, namespace="other")], its=its),  # namespace must match
        _node("dne", [fx.pod({"cpu": "1"}, nodeName="dne", doNotEvict=True)], its=its),
        _node("dne-static", [fx.pod({"cpu": "1"}, nodeName="dne-static", doNotEvict=True, ownedByNode=True)], its=its),  # owned by the node: ignored
        _node("dnc", [fx.pod({"cpu": "1"}, nodeName="dnc")], its=its, doNotConsolidate="true"),
        _node("dnc-false", [fx.pod({"cpu": "1"}, nodeName="dnc-false")], its=its, doNotConsolidate="false"),
        _node("nominated", [fx.pod({"cpu": "1"}, nodeName="nominated")], its=its, nominated=True),
        _node("deleting", [fx.pod({"cpu": "1"}, nodeName="deleting")], its=its, deleting=True),
        _node("marked", [fx.pod({"cpu": "1"}, nodeName="marked")], its=its, markedForDeletion=True),
        _node("uninit", [fx.pod({"cpu": "1"}, nodeName="uninit")], its=its, initialized=False),
        _node("foreign", [fx.pod({"cpu": "1"}, nodeName="foreign")], its=its, provisioner_name="someone-else"),
        _node("empty", [], its=its),
    ]
    nodes.append(fx.state_node("unknown-type", "no-such-type", allocatable={"cpu": "32", "memory": "64Gi", "pods": "100"},
                               pods_=[fx.pod({"cpu": "1"}, nodeName="unknown-type")]))
    prob = _problem(nodes, its, pdbs=[{"namespace": "default", "selector": {"matchLabels": app}, "disruptionsAllowed": 0}])

    def check(order, cost):
        names = [nodes[i]["name"] for i in order]
        assert names == ["empty", "ok", "pdb-other-ns", "dne-static", "dnc-false"]   # cost 0 first, then stable among the cost-1 nodes
    return prob, check


SINGLE = []


def single(ref):
    def deco(fn):
        SINGLE.append((fn.__name__, ref, fn))
        return fn
    return deco


@single("singlenodeconsolidation.go:43-84 with suite_test.go:1422-1496 (delete when the pods fit elsewhere)")
def first_deletable_candidate_wins():
    its = ca.assorted()
    big = ca.on_demand_by_price(its)[-1]
    # node0 is full of pods that fit nowhere else (cost 4: tried last); node1's single pod fits on node2
    full = [fx.pod({"cpu": "7"}, nodeName="n0") for _ in range(4)]
    n0 = ca.node_of_type("n0", big, full, 4.0)
    n1 = ca.node_of_type("n1", big, [fx.pod({"cpu": "1"}, nodeName="n1")], 1.0)
    n2 = ca.node_of_type("n2", big, [fx.pod({"cpu": "1"}, nodeName="n2")], 2.0)
    prob = fx.problem([], instance_types=its, nodes=[n0, n1, n2])

    def check(cmd):
        assert cmd["action"] == 1 and cmd["node"] == 1 and cmd["simulations"] == 1
    return prob, check


@single("singlenodeconsolidation.go:43-84 with suite_test.go:874-929 (replace with a cheaper node; no filterOutSameType for a single node)")
def replace_when_nothing_can_be_deleted():
    its = ca.assorted()
    worst = ca.on_demand_by_price(its)[-1]
    n0 = ca.node_of_type("n0", worst, [fx.pod({"cpu": "30"}, nodeName="n0")], 1.0)      # needs a 32-cpu node: nothing cheaper on-demand... spot is
    n1 = ca.node_of_type("n1", worst, [fx.pod({"cpu": "31"}, nodeName="n1")], 2.0)
    prob = fx.problem([], instance_types=its, nodes=[n0, n1])

    def check(cmd):
        assert cmd["action"] == 2 and cmd["node"] == 0 and cmd["simulations"] == 1
        names = {its[i]["name"] for i in cmd["options"]}
        assert names and all(n.startswith("32c-") for n in names)
        price = worst["offerings"][0]["price"]
        assert all(its[i]["offerings"][0]["price"] < price for i in cmd["options"])
    return prob, check


@single("singlenodeconsolidation.go:54-77 (every candidate is tried before giving up)")
def nothing_to_do():
    its = ca.assorted()
    cheapest = ca.on_demand_by_price(its)[0]
    spot = [it for it in its if it["offerings"][0]["capacityType"] == "spot"]
    cheapest_spot = sorted(spot, key=lambda it: it["offerings"][0]["price"])[0]
    n0 = ca.node_of_type("n0", cheapest_spot, [fx.pod({"cpu": "900m"}, nodeName="n0")], 1.0)
    n0["allocatable"] = {"cpu": "1", "memory": "2Gi", "pods": "100"}
    n1 = ca.node_of_type("n1", cheapest_spot, [fx.pod({"cpu": "900m"}, nodeName="n1")], 2.0)
    n1["allocatable"] = {"cpu": "1", "memory": "2Gi", "pods": "100"}
    prob = fx.problem([], instance_types=its, nodes=[n0, n1])

    def check(cmd):
        assert cmd["action"] == 0 and cmd["node"] == -1 and cmd["simulations"] == 2
    return prob, check

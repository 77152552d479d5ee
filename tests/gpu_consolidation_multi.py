"""Multi-GPU consolidation search (BASELINE config 5: 5k pods on 1k nodes x 1k types), one rank per GPU.
Run: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/gpu_consolidation_multi.py
Each rank evaluates its share of every round's probe frontier on its own GPU; the command equals the sequential one."""
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import torch
import torch.distributed as dist

import __graft_entry__ as g


def main():
    pkg = g.load_pkg()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    pkg.lib().kh_set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pods = int(os.environ.get("KS_PODS", 5000))
    nodes = int(os.environ.get("KS_NODES", 1000))
    problem = pkg.Problem.synth(5, pods, 1000, 42, nodes)
    mnc = pkg.MultiNodeConsolidation(problem)

    def all_gather(obj):
        parts = [None] * world
        dist.all_gather_object(parts, obj)
        return parts

    mnc.probe(2)  # warm-up: context, catalog tables
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sharded = mnc.first_n_node_consolidation_option_sharded(rank, world, all_gather if world > 1 else None)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    out = {"n_gpus": world, "pods": pods, "nodes": nodes, "sharded_s": round(t1 - t0, 4), "rounds": sharded["rounds"],
           "probes_on_path": len(sharded["probes"]), "action": sharded["action"], "nodes_removed": sharded["nodes_removed"],
           "n_options": len(sharded["options"])}
    if rank == 0:
        t0 = time.perf_counter()
        seq = mnc.first_n_node_consolidation_option()
        out["sequential_s"] = round(time.perf_counter() - t0, 4)
        out["same_command"] = (seq["action"], seq["nodes_removed"], seq["options"]) == (sharded["action"], sharded["nodes_removed"], sharded["options"])
        if os.environ.get("KS_ORACLE", "1") == "1":
            import oracle_lib
            t0 = time.perf_counter()
            want = oracle_lib.load().consolidate(problem)
            out["oracle_s"] = round(time.perf_counter() - t0, 3)
            out["parity_vs_oracle"] = (want["action"], want["nodes_removed"], want["options"]) == (sharded["action"], sharded["nodes_removed"], sharded["options"])
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

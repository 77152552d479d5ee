"""bench.py's reference arm runs on the CPU here; the N>1 host-side plumbing (unique-id broadcast, shard ranges,
max-over-ranks timing) is exercised with world_size-2 gloo processes."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1", "--pods", "400"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "pods/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] == 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["higher_is_better"] is True and line["steps"] == 2


WORKER = r'''
import os, sys, json
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import torch, torch.distributed as dist
from __graft_entry__ import load_pkg
pkg = load_pkg()
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
# 1) the 128-byte communicator id travels from rank 0 to everyone (NCCL itself needs GPUs; the plumbing does not)
uid = torch.arange(128, dtype=torch.uint8) if rank == 0 else torch.zeros(128, dtype=torch.uint8)
dist.broadcast(uid, 0)
assert uid.tolist() == list(range(128))
# 2) every rank encodes the same problem and owns a disjoint column-word range
problem = pkg.Problem.synth(3, 500, 1000, 42, 0)
rs = pkg.ResidentSolve(problem)
w32 = rs.dims["type_words"] * 2
b, e = pkg.shard_range(w32, rank, world)
spans = [None] * world
dist.all_gather_object(spans, (b, e))
covered = [w for bb, ee in spans for w in range(bb, ee)]
assert covered == list(range(w32)), spans
# 3) timing is the max over ranks
t = torch.tensor([1.0 + rank], dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
assert t.item() == float(world)
dist.destroy_process_group()
print("ok", rank)
'''


def test_world_size_two_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29517")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29517", str(script), str(ROOT)], capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("ok") == 2

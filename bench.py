#!/usr/bin/env python3
"""bench.py — pods scheduled/sec of Scheduler.Solve on a synthetic cluster (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA path through the C-ABI)
  python bench.py --impl reference --gpus N --steps K ...   the reference's algorithm on the host CPU (oracle port)

A "step" is one Solve over one batch of synthetic pods. Workload at N=1: BASELINE.json configs[1]
(C2: 10 000 pods x 500 instance types, cpu/memory requests only, one provisioner).
  value     pods/s with the encoded problem already resident in HBM (sort + feasibility + pack kernels), L2 flushed
            before every step, timed with CUDA events on the library's stream, max over ranks.
  e2e       the same metric through ksched_solve with HOST buffers: pod/node/topology upload, kernels and result
            download inside the timed region (the instance-type catalog stays resident, as across reconciles).
  roofline  the dominant kernel (pack_kernel); roofline_feasibility the kernel BASELINE.json's target names.
N>1: K1's columns are sharded over the ranks, one ncclAllReduce(min) on the per-pod best-column vector, the pack
chain runs replicated — total work is fixed, so scaling is "strong".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

CONFIGS = {
    2: dict(pods=10_000, types=500, nodes=0, name="C2: 10k pods x 500 instance types, cpu/mem requests only, 1 provisioner"),
    3: dict(pods=50_000, types=1000, nodes=0, name="C3: 50k pods nodeSelector+taints x 1k instance types, 3 provisioners"),
    4: dict(pods=100_000, types=1000, nodes=0, name="C4: 100k pods anti-affinity+topology-spread x 1k instance types"),
    1: dict(pods=100, types=10, nodes=0, name="C1: 100 pods x 10 instance types (reference benchmark shape)"),
}


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(kernel):
    p = ROOT / "profiles" / "traffic.json"
    if p.exists():
        try:
            return json.loads(p.read_text()).get(kernel)
        except Exception:
            return None
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.proc = None

    def start(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx = max(mx, float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def run_reference(args, rank, world):
    """The reference's algorithm on the host CPU: the C++ oracle (the Go code cannot be built in this image)."""
    if rank != 0:
        return
    from __graft_entry__ import load_pkg
    import oracle_lib
    pkg, oracle = load_pkg(), oracle_lib.load()
    cfg = CONFIGS[args.config]
    pods = args.pods or cfg["pods"]
    types = args.types or cfg["types"]
    # bound the sample so that steps+warmup finish within a few minutes
    budget_s = 150.0
    sample = pods
    problem = pkg.Problem.synth(args.config, sample, types, 42, cfg["nodes"])
    res = pkg.Result()
    t0 = time.perf_counter()
    oracle.solve(problem, res)
    first = time.perf_counter() - t0
    total_runs = args.steps + args.warmup
    while first * total_runs > budget_s and sample > 500:
        sample = max(500, sample // 2)
        problem = pkg.Problem.synth(args.config, sample, types, 42, cfg["nodes"])
        t0 = time.perf_counter()
        oracle.solve(problem, res)
        first = time.perf_counter() - t0
    for _ in range(max(0, args.warmup - 1)):
        oracle.solve(problem, res)
    times = []
    scheduled = 0
    for _ in range(args.steps):
        t0 = time.perf_counter()
        oracle.solve(problem, res)
        times.append(time.perf_counter() - t0)
        scheduled = int((res.assign >= 0).sum())
    total = sum(times)
    value = scheduled * args.steps / total
    sample_desc = f"{sample} of {pods} pods of the workload, {types} instance types, {args.steps} full Solve runs" if sample != pods \
        else f"full workload ({pods} pods, {types} instance types), {args.steps} Solve runs"
    line = {
        "impl": "reference", "metric": "pods scheduled/sec (Scheduler.Solve)", "value": value, "unit": "pods/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000 * total / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": {"workload": cfg["name"], "pods": sample, "instance_types": types, "note": "C++ oracle port of the Go reference, 1 thread "
                   "(Scheduler.Solve is a single goroutine); the Go toolchain is absent so the reference itself cannot run"},
        "cpu_baseline": {"value": value, "unit": "pods/s", "cores": 1, "kind": "port", "sample": sample_desc},
        "e2e": {"value": value, "unit": "pods/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS))
    ap.add_argument("--pods", type=int, default=0)
    ap.add_argument("--types", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import ctypes as C
    import numpy as np
    import torch
    from __graft_entry__ import load_pkg
    pkg = load_pkg()
    if not torch.cuda.is_available() or pkg.device_count() < 1:
        raise RuntimeError("bench.py needs a CUDA device: the solver has no CPU path")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    L = pkg.lib()
    L.kh_set_device(local_rank)

    cfg = CONFIGS[args.config]
    pods = args.pods or cfg["pods"]
    types = args.types or cfg["types"]
    problem = pkg.Problem.synth(args.config, pods, types, 42, cfg["nodes"])
    t0 = time.perf_counter()
    rs = pkg.ResidentSolve(problem)
    encode_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    rs.load_catalog()
    catalog_s = time.perf_counter() - t0
    rs.load()
    rs.run()  # one untimed run with the exact nodes_visited statistic (feeds the pack kernel's algorithmic bytes)
    nodes_visited = int(rs.download().nodes_visited)
    rs.set_count_visited(False)
    rs.load()
    handle = L.kh_handle()
    if world > 1:
        L.ksched_set_shard(C.c_void_p(handle), rank, world)
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = (C.c_ubyte * 128)()
            assert L.ksched_nccl_unique_id(buf) == 0
            uid = torch.tensor(list(buf), dtype=torch.uint8)
        uid = uid.cuda()
        dist.broadcast(uid, 0)
        raw = bytes(uid.cpu().tolist())
        rc = L.ksched_nccl_init(C.c_void_p(handle), raw, rank, world)
        if rc != 0:
            raise RuntimeError("ksched_nccl_init failed")

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        rs.run(flush_l2=True)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    phase = {"sort_us": 0.0, "feasibility_us": 0.0, "pack_us": 0.0, "allreduce_us": 0.0, "total_us": 0.0}
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        rs.run(flush_l2=True)
        t = rs.timings()
        for k in phase:
            phase[k] += t[k]
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop()
    dev_s = phase["total_us"] * 1e-6
    if dist is not None:
        tt = torch.tensor([dev_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dev_s = float(tt.item())
    res = rs.download()
    scheduled = int((res.assign >= 0).sum())
    tm = rs.timings()
    value = scheduled * args.steps / dev_s

    # ---- e2e: ksched_solve with host buffers (upload + kernels + download inside the timed region)
    for _ in range(2):
        rs.solve_e2e()
    barrier()
    e2e_us = 0.0
    for _ in range(args.steps):
        us, _ = rs.solve_e2e()
        e2e_us += us
    barrier()
    tm_e2e = rs.timings()
    e2e_s = e2e_us * 1e-6
    if dist is not None:
        tt = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())
    e2e_value = scheduled * args.steps / e2e_s

    # ---- feasibility kernel alone, cold L2, CUDA events on the launching stream
    rs.load()
    rs.run(flush_l2=False)
    k1_us = []
    for _ in range(max(5, args.steps)):
        k1_us.append(rs.run_feasibility(flush_l2=True))
    k1_us.sort()
    k1_avg = sum(k1_us) / len(k1_us)
    peak, peak_src = measured_peak()
    k1_bytes = tm["feasibility_bytes"]
    k1_gbs = k1_bytes / (k1_avg * 1e-6) / 1e9
    d = rs.dims
    phys_k1 = d["pods"] * 256 + d["templates"] * d["types"] * 0 + d["pods"] * d["templates"] * d["type_words"] * 8 + d["pods"] * 8
    pack_avg_us = phase["pack_us"] / args.steps
    pack_bytes = nodes_visited * 128 + d["pods"] * 256
    pack_gbs = pack_bytes / (pack_avg_us * 1e-6) / 1e9

    # ---- the same kernel on a workload large enough not to be launch-bound: C4's shape (100k pods x 1k types, 38.4 MB)
    k1_big = None
    c4_line = None
    if rank == 0 and not args.no_cpu_baseline:
        big = pkg.Problem.synth(4, 100_000, 1000, 42, 0)
        rb = pkg.ResidentSolve(big)
        rb.load()
        tb = sorted(rb.run_feasibility(flush_l2=True) for _ in range(8))
        bbytes = rb.timings()["feasibility_bytes"]
        bavg = sum(tb) / len(tb)
        k1_big = {"kernel": "feasibility_kernel", "workload": "C4 shape: 100k pod rows x 1 provisioner x 1000 instance types", "bound": "hbm",
                  "achieved": bbytes / (bavg * 1e-6) / 1e9, "peak": peak, "unit": "GB/s", "frac": bbytes / (bavg * 1e-6) / 1e9 / peak,
                  "algorithmic_bytes": int(bbytes), "us_per_launch": bavg, "us_min": tb[0], "traffic": ncu_traffic("feasibility_kernel_c4")}
        if world == 1:  # a sharded handle would wait for the other ranks' allreduce
            # ... and the whole Solve on that configuration (BASELINE.json's target shape), resident, L2 flushed: 1 warm-up + 2 runs
            rb.set_count_visited(False)
            rb.load()
            rb.run(flush_l2=True)
            c4_ph = {"sort_us": 0.0, "feasibility_us": 0.0, "pack_us": 0.0, "total_us": 0.0}
            c4_runs = 2
            for _ in range(c4_runs):
                rb.run(flush_l2=True)
                t4 = rb.timings()
                for k in c4_ph:
                    c4_ph[k] += t4[k]
            r4 = rb.download()
            s4 = int((r4.assign >= 0).sum())
            c4_line = {"workload": CONFIGS[4]["name"], "pods": 100_000, "instance_types": 1000, "scheduled": s4, "new_nodes": int(r4.num_new_nodes),
                       "value": s4 * c4_runs / (c4_ph["total_us"] * 1e-6), "unit": "pods/s", "ms_per_step": c4_ph["total_us"] / c4_runs / 1000,
                       "phases_ms_per_step": {k[:-3]: v / c4_runs / 1000 for k, v in c4_ph.items()}, "steps": c4_runs, "warmup": 1,
                       "note": "secondary measurement on one GPU; parity for this shape is covered by tests/test_gpu_fullsize.py"}
            del r4
        del rb, big
        rs.load()  # make the benchmarked problem resident again

    line = {
        "metric": "pods scheduled/sec (Scheduler.Solve)", "value": value, "unit": "pods/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000 * dev_s / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "int64", "data": "synthetic",
        "config": {"workload": cfg["name"], "pods": pods, "instance_types": types, "scheduled": scheduled, "new_nodes": int(res.num_new_nodes),
                   "l2": "flushed before every timed step (256 MiB write)", "parallelism": f"K1 columns sharded x{world}, pack replicated",
                   "phases_ms_per_step": {k[:-3]: v / args.steps / 1000 for k, v in phase.items()},
                   "encode_host_ms": 1000 * encode_s, "catalog_load_ms": 1000 * catalog_s, "wall_ms_per_step": 1000 * wall / args.steps,
                   "library_kernels_per_step": "3x cub::DeviceRadixSort::SortPairs (not counted in gpu_launches)"},
        "e2e": {"value": e2e_value, "unit": "pods/s", "h2d_bytes_per_step": int(tm_e2e["h2d_bytes"]), "d2h_bytes_per_step": int(tm_e2e["d2h_bytes"]),
                "ms_per_step": 1000 * e2e_s / args.steps, "path": "ksched_solve(handle, problem*, result*) with host buffers"},
        # own kernels per resident Solve: sort_keys, 2x gather_u64, gather_rows, feasibility, pack, finalize_options
        "gpu_launches": 7 * args.steps,
        "roofline": {"kernel": "pack_kernel", "bound": "hbm", "achieved": pack_gbs, "peak": peak, "unit": "GB/s", "frac": pack_gbs / peak,
                     "traffic": ncu_traffic("pack_kernel"), "peak_source": peak_src, "algorithmic_bytes": int(pack_bytes),
                     "us_per_launch": pack_avg_us, "note": "latency-bound sequential first-fit chain; bytes = nodes_visited*128 + P*256 (SURVEY 8d K2)"},
        "roofline_feasibility": {"kernel": "feasibility_kernel", "bound": "hbm", "achieved": k1_gbs, "peak": peak, "unit": "GB/s", "frac": k1_gbs / peak,
                                 "traffic": ncu_traffic("feasibility_kernel"), "peak_source": peak_src, "algorithmic_bytes": int(k1_bytes),
                                 "physical_bytes": int(phys_k1), "us_per_launch": k1_avg, "us_min": k1_us[0],
                                 "note": "bytes = P*256 + C*256 + P*C/8 (SURVEY 8d K1), cold L2"},
        "clocks": clocks,
    }
    if k1_big:
        line["roofline_feasibility_c4"] = k1_big
    if c4_line:
        line["config_c4"] = c4_line
    if rank == 0 and world > 1 and not args.no_cpu_baseline:
        # parity of the sharded path (K1 column shards + allreduce, pack replicated) against the oracle on the same inputs
        import oracle_lib
        want = pkg.Result()
        oracle_lib.load().solve(problem, want)
        line["config"]["parity_vs_oracle"] = bool((want.assign == res.assign).all() and want.num_new_nodes == res.num_new_nodes)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import oracle_lib
        oracle = oracle_lib.load()
        sample = min(pods, 10_000) if args.config == 2 else min(pods, 2_000)
        sp = problem if sample == pods else pkg.Problem.synth(args.config, sample, types, 42, cfg["nodes"])
        want = pkg.Result()
        t0 = time.perf_counter()
        oracle.solve(sp, want)
        cpu_s = time.perf_counter() - t0
        cpu_sched = int((want.assign >= 0).sum())
        line["cpu_baseline"] = {"value": cpu_sched / cpu_s, "unit": "pods/s", "cores": 1, "kind": "port",
                                "sample": f"oracle Solve of {sample} of {pods} pods x {types} instance types, 1 run, {cpu_s:.2f} s; host has {os.cpu_count()} cores"}
        if sample == pods:
            line["config"]["parity_vs_oracle"] = bool((want.assign == res.assign).all() and want.num_new_nodes == res.num_new_nodes)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

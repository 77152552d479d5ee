"""Turn the raw ncu outputs of one gpurun call (gpurun_out/) into the small summaries kept under profiles/.

  python profiles/summarize.py r02        # reads gpurun_out/<tag>_launches.csv (or launches.csv), gpurun_out/<tag>_full.ncu-rep

launch list  : ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \\
               python bench.py --steps 2 --warmup 3 --no-cpu-baseline
full capture : ncu --set full --clock-control none --import-source on -k regex:"feasibility_kernel|pack_kernel" -s 6 -c 2 \\
               -o gpurun_out/r01_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline
Per-launch times under ncu are cold-cache and serialised: the SHARES are what agrees with bench.py, not the absolutes.
"""
import collections
import csv
import io
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
OUT = ROOT / "gpurun_out"
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"


def launch_shares():
    src = OUT / f"{tag}_launches.csv"
    if not src.exists():
        src = OUT / "launches.csv"
    if not src.exists():
        return
    rows = [l for l in src.read_text().splitlines() if l.startswith('"')]
    per = collections.OrderedDict()
    for r in csv.DictReader(io.StringIO("\n".join(rows))):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
        k = r["Kernel Name"]
        n, t = per.get(k, (0, 0.0))
        per[k] = (n + 1, t + us)
    total = sum(t for _, t in per.values())
    lines = ["ncu --metrics gpu__time_duration.sum --clock-control none (python bench.py --steps 2 --warmup 3 --no-cpu-baseline);",
             "per-launch times are cold and serialised: compare SHARES", "kernel | launches | total_us | share"]
    for k, (n, t) in sorted(per.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"{k[:58]} | {n} | {t:.1f} | {100 * t / total:.2f}%")
    (ROOT / "profiles" / f"{tag}_launch_shares.txt").write_text("\n".join(lines) + "\n")
    (ROOT / "profiles" / f"{tag}_launches.csv").write_text("\n".join(rows) + "\n")


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__cycles_active.avg", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warp_latency_per_inst_issued.ratio"]


def full_capture():
    rep = OUT / f"{tag}_full.ncu-rep"
    if not rep.exists():
        return
    suffix = "" if tag == "r01" else "_c4"  # round 2 on: the bench's workload is C4
    old = ROOT / "profiles" / "traffic.json"
    traffic_prev = json.loads(old.read_text()) if old.exists() else {}
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    kernels, traffic = [], {}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for r in data:
        name = r[ix["Kernel Name"]]
        rec = {"Kernel Name": name}
        for w in WANT:
            if w in ix:
                rec[w] = r[ix[w]] + " " + units[ix[w]]
        kernels.append(rec)
        short = "pack_kernel" if "pack_kernel" in name else ("class_feasibility_kernel" if "class_feasibility" in name else
                                                             ("feasibility_kernel" if "feasibility_kernel" in name else name.split("(")[0]))
        short += suffix
        rd = float(r[ix["dram__bytes_read.sum"]].replace(",", "")) * scale.get(units[ix["dram__bytes_read.sum"]], 1.0)
        wr = float(r[ix["dram__bytes_write.sum"]].replace(",", "")) * scale.get(units[ix["dram__bytes_write.sum"]], 1.0)
        traffic[short] = rd + wr
    (ROOT / "profiles" / f"{tag}_ncu_full_summary.json").write_text(json.dumps({
        "command": "ncu --set full --clock-control none --import-source on -k regex:feasibility_kernel|pack_kernel python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-c5 (the step after the warm-ups)",
        "kernels": kernels}, indent=1) + "\n")
    traffic_prev.update(traffic)
    (ROOT / "profiles" / "traffic.json").write_text(json.dumps(traffic_prev, indent=1) + "\n")


launch_shares()
full_capture()

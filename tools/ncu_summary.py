#!/usr/bin/env python
"""Condense ncu outputs into the small text files committed under profiles/.

  python tools/ncu_summary.py list  gpurun_out/launches.csv   > profiles/rNN_launches.txt
  python tools/ncu_summary.py full  gpurun_out/prof.ncu-rep   > profiles/rNN_<kernel>_full.txt
  python tools/ncu_summary.py traffic gpurun_out/gemm_traffic.csv "<command>" > profiles/rNN_gemm_traffic.json
"""
import collections
import csv
import re
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor", "sm__cycles_elapsed.max",
    "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
]


def clean(name):
    name = re.sub(r"\(.*", "", name)
    return name.replace("<unnamed>::", "").replace("void ", "").strip()


def launch_list(path):
    lines = open(path).readlines()
    start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
    agg = collections.OrderedDict()
    total = 0.0
    n = 0
    for r in csv.DictReader(lines[start:]):
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        v = v / 1e3 if r["Metric Unit"] == "ns" else (v * 1e3 if r["Metric Unit"] == "ms" else v)
        k = clean(r["Kernel Name"])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
        total += v
        n += 1
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised): {n} launches, {total/1e3:.2f} ms")
    print(f"{'us total':>12} {'share':>6} {'n':>6} {'us avg':>9}  kernel")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{t:12.1f} {100*t/total:5.1f}% {c:6d} {t/c:9.1f}  {k[:100]}")


def traffic(path, command, only="gemm_bf16_tn"):
    """Mean DRAM bytes (read + write) and duration per launch, overall and per kernel variant (bench.py roofline.traffic);
    launches whose kernel name does not match `only` are skipped (the csv may hold the whole step)."""
    import json

    lines = open(path).readlines()
    start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
    per_id = collections.OrderedDict()
    for r in csv.DictReader(lines[start:]):
        if only and not re.search(only, r["Kernel Name"]):
            continue
        d = per_id.setdefault(r["ID"], {"k": clean(r["Kernel Name"]), "bytes": 0.0, "us": 0.0})
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        if r["Metric Name"].startswith("dram__bytes"):
            d["bytes"] += v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        elif r["Metric Name"] == "gpu__time_duration.sum":
            d["us"] += v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
    var = collections.OrderedDict()
    for d in per_id.values():
        a = var.setdefault(d["k"], [0, 0.0, 0.0])
        a[0] += 1
        a[1] += d["bytes"]
        a[2] += d["us"]
    n = len(per_id)
    out = {"source": command, "launches": n,
           "dram_bytes_per_launch": sum(d["bytes"] for d in per_id.values()) / max(1, n),
           "per_variant": {k: {"launches": c, "dram_bytes_per_launch": b / c, "us_per_launch": t / c}
                           for k, (c, b, t) in var.items()}}
    print(json.dumps(out, indent=1))


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for row in rows[2:]:
        d = dict(zip(hdr, row))
        print("kernel:", clean(d.get("Kernel Name", "?")), "grid", d.get("Grid Size"), "block", d.get("Block Size"))
        for k in KEEP:
            if k in d:
                print(f"  {k:75s} {d[k]:>16} {units[hdr.index(k)]}")
        print()


if __name__ == "__main__" and len(sys.argv) > 2 and sys.argv[1] == "traffic":
    traffic(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "", sys.argv[4] if len(sys.argv) > 4 else "gemm_bf16_tn")
    sys.exit(0)

if __name__ == "__main__":
    {"list": launch_list, "full": full}[sys.argv[1]](sys.argv[2])

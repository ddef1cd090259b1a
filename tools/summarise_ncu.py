#!/usr/bin/env python
"""Summarise ncu output into the small text files committed under profiles/.

  python tools/summarise_ncu.py launches gpurun_out/r01c_launches.csv  > profiles/r01c_launch_shares.md
  python tools/summarise_ncu.py full gpurun_out/r01c_gemm_conv.ncu-rep > profiles/r01c_gemm_conv_full.md

`launches`: per-kernel launch count, total and share of gpu__time_duration.sum (cold-cache, serialised: compare shares).
`full`    : the roofline-relevant raw metrics of every captured launch (needs the `ncu` CLI to read the report).
"""
import csv
import re
import subprocess
import sys
from collections import defaultdict

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "lts__t_bytes.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg"]


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = name.replace("ssg::", "").replace("tc::", "").replace("(anonymous namespace)::", "")
    return name.strip()[:90]


def launches(path):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.reader(lines)
    header = None
    for r in rd:
        if header is None:
            if "Kernel Name" in r:
                header = r
            continue
        rows.append(dict(zip(header, r)))
    agg = defaultdict(lambda: [0, 0.0])
    total = 0.0
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9, "nsecond": 1, "usecond": 1e3, "msecond": 1e6, "second": 1e9}.get(unit, 1)
        k = short(r["Kernel Name"])
        agg[k][0] += 1
        agg[k][1] += ns
        total += ns
    print("| kernel | launches | total ms | share |")
    print("|---|---:|---:|---:|")
    for k, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.3f | %.1f%% |" % (k, c, ns / 1e6, 100 * ns / total))
    print("| **total** | %d | %.3f | 100%% |" % (sum(c for c, _ in agg.values()), total / 1e6))


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(out.splitlines()))
    if not rd:
        print("no data")
        return
    header = rd[0]
    cols = {h: i for i, h in enumerate(header)}
    print("| launch | kernel | " + " | ".join(k for k in KEYS if k in cols) + " |")
    print("|---|---|" + "---:|" * len([k for k in KEYS if k in cols]))
    for r in rd[2:]:
        if len(r) < len(header):
            continue
        print("| %s | `%s` | " % (r[cols.get("ID", 0)], short(r[cols["Kernel Name"]])) +
              " | ".join(r[cols[k]] for k in KEYS if k in cols) + " |")
    print()
    print("units: " + ", ".join("%s=%s" % (k, rd[1][cols[k]]) for k in KEYS if k in cols))


def traffic(path):
    """DRAM bytes + time of every convolution GEMM launch of one embedding batch -> JSON on stdout."""
    import json
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.reader(lines)
    header = next(rd)
    per = defaultdict(dict)
    order = []
    for r in rd:
        row = dict(zip(header, r))
        i = row["ID"]
        if i not in per:
            order.append(i)
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6,
                 "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}.get(u, 1)
        per[i][row["Metric Name"]] = v * scale
        per[i]["kernel"] = short(row["Kernel Name"])
    launches_ = [{"kernel": per[i]["kernel"], "us": per[i].get("gpu__time_duration.sum", 0.0),
                  "dram_read": per[i].get("dram__bytes_read.sum", 0.0), "dram_write": per[i].get("dram__bytes_write.sum", 0.0)}
                 for i in order]
    out = {"launches": len(launches_), "batch_images": 512,
           "dram_bytes_per_batch": sum(l["dram_read"] + l["dram_write"] for l in launches_),
           "us_per_batch": sum(l["us"] for l in launches_), "per_launch": launches_}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](sys.argv[2])

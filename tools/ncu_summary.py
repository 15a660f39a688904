#!/usr/bin/env python
"""Summarise ncu output for profiles/: a launch list CSV (--metrics gpu__time_duration.sum) into
per-kernel totals/shares, and an .ncu-rep (--set full) into the handful of metrics the roofline
argument needs.  Usage:
    python tools/ncu_summary.py launches <launches.csv>
    python tools/ncu_summary.py rep <file.ncu-rep>
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_op_dmma.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]


def launches(path):
    rows = list(csv.reader(open(path)))
    h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H, data = rows[h], rows[h + 1:]
    ki, vi = H.index("Kernel Name"), H.index("Metric Value")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        a = agg.setdefault(r[ki].split("(")[0][-48:], [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    print("%-48s %6s %12s %8s %10s" % ("kernel", "n", "total_us", "share", "avg_us"))
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%-48s %6d %12.1f %7.1f%% %10.1f" % (k, n, t / 1e3, 100 * t / tot, t / n / 1e3))


def rep(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True,
                         text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    H, units = rows[0], rows[1]
    for r in rows[2:]:
        print("kernel:", r[H.index("Kernel Name")][:110])
        for k in KEYS:
            if k in H:
                print("  %-72s %s %s" % (k, r[H.index(k)], units[H.index(k)]))
        for i, name in enumerate(H):
            if ("pipe_fp64" in name or "dmma" in name) and name not in KEYS and "pct" in name:
                print("  %-72s %s %s" % (name, r[i], units[i]))
        print()


def traffic(path, workload, n_gpus, chains, out="profiles/r02_traffic.json"):
    """add / replace the entries of profiles/r02_traffic.json for the kernels of one --set full report:
    python tools/ncu_summary.py traffic <file.ncu-rep> <workload> <n_gpus> <chains>"""
    import json
    import os
    import re

    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True,
                         check=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    H, units = rows[0], rows[1]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    doc = json.load(open(out)) if os.path.exists(out) else {"entries": []}
    seen = {}
    for r in rows[2:]:
        name = re.sub(r"<.*", "", r[H.index("Kernel Name")].split("(")[0]).split("::")[-1]
        name = name.replace("void ", "").replace("_kernel", "").strip()
        vals = []
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = H.index(k)
            vals.append(float(r[i].replace(",", "")) * scale[units[i]])
        if any(v != v for v in vals):
            continue  # counters overflowed (nan)
        seen.setdefault(name, []).append((vals, r[H.index("gpu__time_duration.sum")]))
    for name, caps in seen.items():
        rd = sum(v[0][0] for v in caps) / len(caps)
        wr = sum(v[0][1] for v in caps) / len(caps)
        e = {"workload": workload, "n_gpus": int(n_gpus), "chains": int(chains), "kernel": name,
             "dram_bytes": rd + wr,
             "metric": "dram__bytes_read.sum %.6f GB + dram__bytes_write.sum %.6f MB (mean of %d launches)"
                       % (rd / 1e9, wr / 1e6, len(caps)),
             "source": "%s (ncu --set full --clock-control none)" % path}
        doc["entries"] = [x for x in doc["entries"]
                          if (x["workload"], x["n_gpus"], x["chains"], x["kernel"]) !=
                          (workload, int(n_gpus), int(chains), name)] + [e]
        print(e)
    json.dump(doc, open(out, "w"), indent=1)


if __name__ == "__main__":
    {"launches": launches, "rep": rep, "traffic": traffic}[sys.argv[1]](*sys.argv[2:])

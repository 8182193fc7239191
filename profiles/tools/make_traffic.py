#!/usr/bin/env python3
"""Writes profiles/traffic.json from `ncu --set full` captures: per workload and kernel the DRAM bytes per launch
(dram__bytes_read.sum + dram__bytes_write.sum) and the counters that say what bounds the kernel, tagged with the hash of
the CUDA sources they were taken from (bench.py drops them when the sources have changed).
usage: make_traffic.py <workload>:<kernel key>:<report.ncu-rep> ...      e.g.  c3:k_taa:gpurun_out/r2_k_taa.ncu-rep"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

KEEP = {"gpu__time_duration.sum": "ms", "smsp__thread_inst_executed_per_inst_executed.ratio": "active_lanes",
        "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_active": "l1_throughput_pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
        "l1tex__t_sector_hit_rate.pct": "l1_hit_pct", "lts__t_sector_hit_rate.pct": "l2_hit_pct"}


def read(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    H, units, vals = rows[0], rows[1], rows[2]
    g = lambda k: float(vals[H.index(k)].replace(",", ""))
    u = lambda k: units[H.index(k)]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    d = {"dram_bytes": g("dram__bytes_read.sum") * scale[u("dram__bytes_read.sum")] + g("dram__bytes_write.sum") * scale[u("dram__bytes_write.sum")]}
    for k, name in KEEP.items():
        v = g(k)
        if k == "gpu__time_duration.sum":
            v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u(k), 1.0)
        d[name] = v
    return d


def main():
    path = os.path.join(ROOT, "profiles", "traffic.json")
    doc = {"source_hash": bench.kernel_source_hash(), "note": "ncu --set full --clock-control none, one launch each, cold caches"}
    for spec in sys.argv[1:]:
        wl, key, rep = spec.split(":")
        doc.setdefault(wl, {})[key] = read(rep)
        doc[wl][key]["report"] = os.path.basename(rep)
    with open(path, "w") as f:
        json.dump(doc, f, indent=1)
    print(json.dumps(doc, indent=1))


if __name__ == "__main__":
    main()

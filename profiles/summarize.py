#!/usr/bin/env python3
"""Summarises an .ncu-rep (one kernel, `ncu --set full --import-source on`) into markdown:
headline counters, SASS opcode histogram (executed instructions vs stall samples) and the stall mix.
usage: summarize.py report.ncu-rep [title] > profiles/<name>.md"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__warps_eligible.avg.per_cycle_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "sm__inst_executed.sum", "smsp__inst_executed.sum"]


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep] + list(args), stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout


def main():
    rep = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else rep
    print("# %s\n" % title)
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "raw", "--csv"))))
    H = rows[0]
    print("Source: `%s` (ncu --set full --clock-control none). Columns = profiled launches.\n" % rep.split("/")[-1])
    print("| metric | unit | values |\n|---|---|---|")
    name_i = H.index("Kernel Name") if "Kernel Name" in H else None
    if name_i is not None:
        print("| kernel | | %s |" % " / ".join(sorted({r[name_i][:70] for r in rows[2:]})))
    for k in KEYS:
        if k in H:
            i = H.index(k)
            print("| %s | %s | %s |" % (k, rows[1][i], ", ".join(r[i] for r in rows[2:])))
    sass = list(csv.reader(io.StringIO(ncu(rep, "--page", "source", "--csv", "--print-source", "sass"))))
    if len(sass) < 3:
        return
    Hs = sass[1]
    si, ii, smp = Hs.index("Source"), Hs.index("Instructions Executed"), Hs.index("# Samples")
    stall_cols = [i for i, h in enumerate(Hs) if h.startswith("stall_") and "Not Issued" not in h]
    data = []
    for r in sass[2:]:
        if r and r[0] == "Kernel Name":
            break
        if len(r) >= len(Hs):
            data.append(r)
    tot_i = sum(int(r[ii]) for r in data) or 1
    tot_s = sum(int(r[smp]) for r in data) or 1
    op, ops = collections.Counter(), collections.Counter()
    for r in data:
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[si])
        o = m.group(2).split(".")[0] if m else "?"
        op[o] += int(r[ii])
        ops[o] += int(r[smp])
    print("\n## SASS opcode mix (first profiled launch): %d static instructions, %.3g executed warp-instructions\n" % (len(data), tot_i))
    print("| opcode | % of executed instructions | % of stall samples |\n|---|---|---|")
    for o, c in op.most_common(22):
        print("| %s | %.1f | %.1f |" % (o, 100 * c / tot_i, 100 * ops[o] / tot_s))
    st = collections.Counter()
    for r in data:
        for i in stall_cols:
            st[Hs[i]] += int(r[i] or 0)
    tt = sum(st.values()) or 1
    print("\n## Warp stall mix (all samples)\n")
    print(", ".join("%s %.1f%%" % (k, 100 * v / tt) for k, v in st.most_common(10)))


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""Executed warp-instructions of one kernel aggregated over source-line ranges.
usage: by_region.py report.ncu-rep file:name:first:last [...]"""
import collections, csv, io, subprocess, sys

def load(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    cur = None; hdr = None
    per = collections.Counter(); thr = collections.Counter(); smp = collections.Counter()
    for r in rows:
        if not r: continue
        if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
        if r[0] == "Line No": hdr = r; continue
        if hdr is None or len(r) < len(hdr) or not r[0].isdigit(): continue
        ie, te, ts = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
        try:
            per[(cur, int(r[0]))] += int(r[ie]); thr[(cur, int(r[0]))] += int(r[te]); smp[(cur, int(r[0]))] += int(r[ts] or 0)
        except ValueError:
            continue
    return per, thr, smp

def main():
    per, thr, smp = load(sys.argv[1])
    tot = sum(per.values()) or 1; tots = sum(smp.values()) or 1
    print("| region | % warp-instructions | % stall samples | avg active lanes |\n|---|---|---|---|")
    for spec in sys.argv[2:]:
        f, name, a, b = spec.split(":")
        keys = [k for k in per if k[0] == f and int(a) <= k[1] <= int(b)]
        n = sum(per[k] for k in keys); t = sum(thr[k] for k in keys); s = sum(smp[k] for k in keys)
        print("| %s (%s:%s-%s) | %.1f | %.1f | %.1f |" % (name, f, a, b, 100 * n / tot, 100 * s / tots, t / max(n, 1)))
main()

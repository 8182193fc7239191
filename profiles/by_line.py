#!/usr/bin/env python3
"""Per-source-line executed warp-instructions of one kernel from an .ncu-rep captured with --import-source on.
usage: by_line.py report.ncu-rep [top_n]"""
import collections, csv, io, subprocess, sys

def num(v):
    try: return int(v)
    except ValueError: return 0

def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    cur_file, hdr = None, None
    per = collections.Counter(); smp = collections.Counter(); thr = collections.Counter(); src = {}
    for r in rows:
        if not r: continue
        if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
        if r[0] == "Function Name": continue
        if r[0] == "Line No": hdr = r; continue
        if hdr is None or len(r) < len(hdr) or not r[0].isdigit(): continue
        ie, ts, te = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
        try: n = int(r[ie])
        except ValueError: continue
        k = (cur_file, int(r[0]))
        per[k] += n; smp[k] += num(r[ts]); thr[k] += num(r[te]); src[k] = r[1].strip()[:110]
    tot = sum(per.values()) or 1; tots = sum(smp.values()) or 1
    byfile = collections.Counter()
    for (f, l), n in per.items(): byfile[f] += n
    print("total warp-instructions %.4g" % tot)
    for f, n in byfile.most_common(): print("  %-18s %5.1f%%" % (f, 100 * n / tot))
    print("| file:line | % inst | % samples | avg lanes | source |\n|---|---|---|---|---|")
    for k, n in per.most_common(top):
        print("| %s:%d | %.2f | %.2f | %.1f | `%s` |" % (k[0], k[1], 100 * n / tot, 100 * smp[k] / tots, thr[k] / max(n, 1), src[k]))
main()

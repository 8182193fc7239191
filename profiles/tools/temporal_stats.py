#!/usr/bin/env python3
"""Hit rate of the per-ray temporal occluder hints on a benchmark scene (LUZRT_TEMPORAL_COUNT=1).
usage: LUZRT_TEMPORAL_COUNT=1 python profiles/tools/temporal_stats.py c3 [frames]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np

from luz_b200 import rt as R
from luz_b200 import workloads

bn = np.fromfile(os.path.join(ROOT, "tests", "golden", "blue_noise_256.rgba"), dtype=np.uint8).reshape(256, 256, 4)
cfg = sys.argv[1] if len(sys.argv) > 1 else "c3"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 6
rt = R.LuzRT(0)
wl = workloads.Workload(rt, cfg)
wl.upload(bn)
wl.step(first=True)
print("| frame | shadow rays | with a hint | settled by the hint | queued | queued and occluded |\n|---|---|---|---|---|---|")
for f in range(frames):
    wl.step()
    rt.sync()
    d = rt.read(R.STATS_DETAIL).astype(np.float64)
    n = max(d[24], 1)
    print("| %d | %d | %.3f | %.3f | %.3f | %.3f |" % (wl.frame, d[24], d[25] / n, d[26] / n, d[27] / n, d[28] / n))

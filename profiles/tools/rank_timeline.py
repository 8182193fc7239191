#!/usr/bin/env python3
"""One rank's share of a multi-GPU C3 frame on a single GPU (no gather): for `ncu --metrics gpu__time_duration.sum` launch
lists of what a rank of an 8-GPU frame executes.  usage: python profiles/tools/rank_timeline.py [world] [frames]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np

from luz_b200 import rt as R
from luz_b200 import workloads

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 5
bn = np.fromfile(os.path.join(ROOT, "tests", "golden", "blue_noise_256.rgba"), dtype=np.uint8).reshape(256, 256, 4)
rt = R.LuzRT(0, rank=0, world=world)
wl = workloads.Workload(rt, "c3")
wl.upload(bn)
app = wl.app
app.update_resources()
app.update_resources_gpu(0)
models, n_models = app.models()
rt.gbuffer_pass(models, n_models)
ts = []
for f in range(frames):
    rt.light_pass(f)
    rt.taa_pass(True)
    rt.swap_light_history()
    rt.sync()
    t = rt.read(R.TIMINGS)
    ts.append((t.light_ms, t.light_rays_ms, t.taa_ms))
print("world %d rank 0: light / rays / taa ms per frame:" % world, [tuple(round(x, 3) for x in t) for t in ts])

#!/usr/bin/env python3
"""Per ray class break-down of a benchmark frame (LUZRT_DEBUG_STATS variant): where the ray kernel's node visits go.
usage: python profiles/tools/ray_classes.py c3 [c4 ...]   (on the GPU box)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np

from luz_b200 import rt as R
from luz_b200 import workloads

bn = np.fromfile(os.path.join(ROOT, "tests", "golden", "blue_noise_256.rgba"), dtype=np.uint8).reshape(256, 256, 4)
for cfg in sys.argv[1:]:
    rt = R.LuzRT(0)
    wl = workloads.Workload(rt, cfg)
    wl.upload(bn)
    wl.step(first=True)
    wl.step()
    rt.set_debug(R.DEBUG_STATS)
    rt.light_pass(wl.app.frame_count)
    st, d = rt.read(R.STATS), rt.read(R.STATS_DETAIL).astype(np.float64)
    print("## %s: %d lit px, %d rays, nodes/ray %.2f tris/ray %.2f insts/ray %.2f" % (
        cfg, st.lit_pixels, st.rays, st.nodes_visited / st.rays, st.triangles_tested / st.rays, st.instances_entered / st.rays))
    print("| class | rays | occluded | TLAS nodes/ray | BLAS nodes/ray | tris/ray | instances/ray | root descents/ray | share of node visits |")
    print("|---|---|---|---|---|---|---|---|---|")
    for c, name in enumerate(("shadow, hinted", "shadow, unhinted", "AO, candidate list", "AO, root descent")):
        r = d[8 * c:8 * c + 8]
        if r[0]:
            print("| %s | %d | %.3f | %.2f | %.2f | %.2f | %.2f | %.3f | %.3f |" % (
                name, r[0], r[1] / r[0], r[2] / r[0], r[3] / r[0], r[4] / r[0], r[5] / r[0], r[6] / r[0],
                (r[2] + r[3]) / st.nodes_visited))
    if d[32]:
        print("AO pixels %d: empty list %.3f, overflowed %.3f, candidates/px %.2f, query+filter nodes/px %.2f (share of node "
              "visits %.3f)" % (d[32], d[33] / d[32], d[34] / d[32], d[35] / d[32], d[36] / d[32], d[36] / st.nodes_visited))
    if d[37]:
        print("hint rays %d, hit %.3f" % (d[37], d[38] / d[37]))
    rt.close()

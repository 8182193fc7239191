#!/usr/bin/env python3
"""Shape of the built 8-wide BVHs of a benchmark scene: slot fill, leaf sizes, depth (from luzrt_blas_dump / tlas_dump).
usage: python profiles/tools/bvh_shape.py c3   (on the GPU box)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np

from luz_b200 import rt as R
from luz_b200 import workloads

NODE = np.dtype([("cbi", "<u4"), ("prim_base", "<u4"), ("meta", "u1", 8), ("box", "<f4", 48)])


def shape(buf, n_prims, prim_bytes, name):
    n_nodes = (len(buf) - n_prims * prim_bytes) // 208
    nodes = np.frombuffer(buf[:n_nodes * 208], dtype=NODE)
    meta = nodes["meta"]
    internal = (meta >> 5 == 1) & ((meta & 31) >= 24)
    leaf = (meta != 0) & ~internal
    cnt = np.zeros_like(meta, dtype=np.int32)
    for u, c in ((1, 1), (3, 2), (7, 3)):
        cnt[leaf & ((meta >> 5) == u)] = c
    used = internal.sum(1) + leaf.sum(1)
    # depth by BFS
    child_base = nodes["cbi"] & 0xFFFFFF
    depth = np.zeros(n_nodes, np.int32)
    for i in range(n_nodes):
        k = int(internal[i].sum())
        depth[child_base[i]:child_base[i] + k] = depth[i] + 1
    lo = nodes["box"].reshape(-1, 6, 8)
    print("%s: %d nodes, %d prims, slots used/node mean %.2f hist %s; internal/node %.2f, leaf slots/node %.2f, prims/leaf slot %.2f; "
          "depth max %d, leaf-slot depth mean %.2f" % (
              name, n_nodes, n_prims, used.mean(), np.bincount(used, minlength=9).tolist(), internal.sum(1).mean(),
              leaf.sum(1).mean(), cnt[leaf].mean() if leaf.any() else 0, depth.max(),
              (np.repeat(depth, 8).reshape(-1, 8)[leaf] + 1).mean() if leaf.any() else 0))
    # surface area of node boxes relative to the root: sum over children (SAH cost proxy)
    ext = np.maximum(lo[:, 1::2, :] - lo[:, 0::2, :], 0)  # (n, 3, 8): hi - lo per axis
    sa = 2 * (ext[:, 0] * ext[:, 1] + ext[:, 1] * ext[:, 2] + ext[:, 0] * ext[:, 2])
    valid = internal | leaf
    root_ext = np.array([lo[0, 1, :][valid[0]].max() - lo[0, 0, :][valid[0]].min(), lo[0, 3, :][valid[0]].max() - lo[0, 2, :][valid[0]].min(),
                         lo[0, 5, :][valid[0]].max() - lo[0, 4, :][valid[0]].min()])
    root_sa = 2 * (root_ext[0] * root_ext[1] + root_ext[1] * root_ext[2] + root_ext[0] * root_ext[2])
    print("   SAH proxy: sum of internal child areas / root area = %.2f, sum of leaf slot areas x prims / root area = %.2f" % (
        sa[internal].sum() / root_sa, (sa * cnt)[leaf].sum() / root_sa))


bn = np.fromfile(os.path.join(ROOT, "tests", "golden", "blue_noise_256.rgba"), dtype=np.uint8).reshape(256, 256, 4)
cfg = sys.argv[1] if len(sys.argv) > 1 else "c3"
rt = R.LuzRT(0)
wl = workloads.Workload(rt, cfg)
wl.upload(bn)
wl.step(first=True)
meshes = wl.app.meshes()
for b in range(1, min(len(meshes), 3) + 1):
    shape(rt.blas_dump(b).tobytes(), len(meshes[b - 1][1]) // 3, 96, "BLAS %d" % b)
shape(rt.tlas_dump().tobytes(), len(wl.app.instances()), 4, "TLAS (%d instances)" % len(wl.app.instances()))

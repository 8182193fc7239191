#!/usr/bin/env python3
"""Writes tests/golden/glsl_ref_c1.npz and glsl_ref_taa.npz: outputs of the reference's own light.frag / taa.comp
(oracle/_ref/libglsl_ref.so, built from /root/reference by oracle/Makefile) on the seeded cases of
tests/glsl_pin_cases.py.  Run here, where the reference is mounted; the fixtures travel to machines without it."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import glsl_pin_cases as cases  # noqa: E402
import glsl_ref_api as G  # noqa: E402
import scene_util as S  # noqa: E402

bn = S.blue_noise()
n = 3000
case = cases.c1_case(n_pixels=n)
rad, sm, am = G.light_frag(case["sc"]["scene"], case["gb"], case["frame"], bn, case["world"], case["pixels"], exhaustive=True)
np.savez_compressed(os.path.join(HERE, "glsl_ref_c1.npz"), n_pixels=n, pixels=case["pixels"], radiance=rad, shadow_mask=sm, ao_mask=am)
case = cases.synthetic_case()
light, hist = cases.taa_images(case, bn)
w, h = case["w"], case["h"]
rows = np.array([0, 1, h // 3, h // 2, h - 2, h - 1])
px = np.concatenate([np.stack([np.arange(w), np.full(w, y)], 1) for y in rows]).astype(np.uint32)
out = {}
for reconstruct, key in ((True, "resolved_reconstruct"), (False, "resolved_plain")):
    out[key] = G.taa_comp(case["sc"]["scene"], light, hist, case["gb"].depth, reconstruct, px).reshape(len(rows), w, 4)
np.savez_compressed(os.path.join(HERE, "glsl_ref_taa.npz"), rows=rows, **out)
print("wrote glsl_ref_c1.npz (%d pixels), glsl_ref_taa.npz (%d rows)" % (case["pixels"].shape[0] if False else n + 3 * 1280, len(rows)))

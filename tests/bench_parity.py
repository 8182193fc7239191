"""Parity of a benchmark workload at its own scale: the frame the product kernels just produced (G-buffer, TLAS, scene
block as they are on the device) against the CPU oracle on a seeded sample of rows.  Used by
tests/test_gpu_bench_parity.py and, outside the timed region, by bench.py (the `parity` object of its JSON line).
Checker code: everything here runs the oracle; nothing in luz_b200/ imports it."""
import time

import numpy as np

import oracle_api as O
from luz_b200 import rt as R

MASK_AGREE = 0.999   # BASELINE.json north_star: >= 99.9 % of the rays
RADIANCE_TOL = 1e-3  # max-abs in linear HDR (or PSNR >= 50 dB)
TAA_TOL = 1e-3       # the resolve is a blend of radiance values: same tolerance
# Sanity bound on pixels whose masks agree, relative to max(|ref|, 1).  The relaxed-precision shading build is ~1e-6 off
# the oracle on ordinary pixels; on a specular highlight of a smooth material GGX's denominator NdotH^2 (a^2 - 1) + 1
# cancels down to a^2 (1e-4 at roughness 0.1), which amplifies the 1e-7 rounding of a differently ordered (contracted)
# evaluation to ~1e-3 -- the spread any two conforming GLSL compilers show on this expression.
REL_TOL = 2e-3


def popcount(a):
    return int(np.unpackbits(np.ascontiguousarray(a).view(np.uint8)).sum())


def psnr(a, b):
    mse = float(np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2))
    peak = float(max(np.abs(b).max(), 1e-12))
    return 99.0 if mse == 0 else float(10.0 * np.log10(peak * peak / mse))


def read_gbuffer(rt, w, h):
    gb = O.GBuffer(w, h)
    for sel, arr in ((R.GBUF_ALBEDO, gb.albedo), (R.GBUF_NORMAL, gb.normal), (R.GBUF_MATERIAL, gb.material),
                     (R.GBUF_EMISSION, gb.emission), (R.GBUF_DEPTH, gb.depth)):
        rt.read(sel, out=arr)
    return gb


def check_workload(wl, blue_noise, n_rows=64, seed=2024, candidate_rows=None, exhaustive_rows=4, exhaustive_cols=96,
                   taa=True, debug_flags=0):
    """wl: luz_b200.workloads.Workload whose last step() has been rendered.  Re-runs the light pass of the next frame id
    with the PRODUCT kernels (debug flags 0 unless given), compares visibility masks and radiance on `n_rows` seeded rows
    (drawn from candidate_rows, default all) with the BVH2 oracle, re-validates the BVH2 oracle against the
    hierarchy-free oracle on a sub-sample, and (taa) compares the resolve of four 4-row groups.  Returns a dict; the
    caller asserts / reports."""
    rt, app = wl.rt, wl.app
    w, h = wl.width, wl.height
    frame = app.frame_count
    sb, extra = app.scene_block(), app.extra_lights()
    t0 = time.perf_counter()
    meshes, instances = app.meshes(), app.instances()
    world = O.World(meshes, instances)
    gb = read_gbuffer(rt, w, h)
    rt.set_scene(sb, extra)
    rt.set_debug(debug_flags)
    rt.light_pass(frame)
    light = rt.read(R.IMG_LIGHT)
    gsm, gam = rt.read(R.SHADOW_MASK), rt.read(R.AO_MASK)
    sw, aw = gsm.shape[-1], gam.shape[-1]

    rng = np.random.default_rng(seed)
    cand = np.arange(h) if candidate_rows is None else np.asarray(candidate_rows)
    rows = np.sort(rng.choice(cand, size=min(n_rows, cand.size), replace=False)).astype(np.uint32)
    rc, ref, sm, am, st = O.light_pass(sb, gb, frame, blue_noise, world, extra_lights=extra, exhaustive=0, row_list=rows,
                                       shadow_words=sw, ao_words=aw)
    assert rc == 0
    bad = popcount(gsm[rows] ^ sm[rows]) + popcount(gam[rows] ^ am[rows])
    agree = 1.0 - bad / max(st.rays, 1)
    same = np.all(gsm[rows] == sm[rows], axis=-1) & np.all(gam[rows] == am[rows], axis=-1)
    got, want = light[rows], ref[rows]
    fin = np.isfinite(want).all(axis=-1)
    max_abs = float(np.abs(got[fin] - want[fin]).max()) if fin.any() else 0.0
    max_abs_same = float(np.abs(got[fin & same] - want[fin & same]).max()) if (fin & same).any() else 0.0
    max_rel_same = float((np.abs(got[fin & same] - want[fin & same]) / np.maximum(np.abs(want[fin & same]), 1.0)).max()) if (fin & same).any() else 0.0
    res = {"config": wl.config + ("-" + wl.variant if wl.variant else ""), "frame": int(frame), "rows": int(rows.size),
           "pixels": int(rows.size) * w, "rays": int(st.rays), "rays_differ": int(bad), "agree": agree,
           "max_abs": max_abs, "max_abs_mask_identical_pixels": max_abs_same, "max_rel_mask_identical_pixels": max_rel_same,
           "max_radiance": float(np.abs(want[fin]).max()) if fin.any() else 0.0, "psnr_db": psnr(got[fin], want[fin]),
           "nan_pixels_match": bool(np.array_equal(np.isnan(got), np.isnan(want))),
           "kernels": "product (set_debug(%d))" % debug_flags, "shadow_words": int(sw), "ao_words": int(aw),
           "oracle": "BVH2 traverser over the same instances; G-buffer, scene block and blue noise identical"}

    # the BVH2 oracle against the oracle without any hierarchy, on a sub-sample of those rows
    if exhaustive_rows:
        inst_tris = sum(len(meshes[m][1]) // 3 for (m, _, _) in instances)
        x0 = max(0, w // 2 - exhaustive_cols // 2)
        xr = (x0, min(w, x0 + exhaustive_cols))
        lit_in_window = (np.abs(gb.normal[rows, xr[0]:xr[1], :3]).sum(axis=-1) > 0).sum(axis=1)
        sub = np.sort(rows[np.argsort(-lit_in_window, kind="stable")[:exhaustive_rows]]).astype(np.uint32)
        # every triangle of every instance where that is affordable, else every instance box + every triangle inside
        rays_sub = exhaustive_rows * exhaustive_cols * max(1, st.rays // max(st.lit_pixels, 1))
        mode = 1 if inst_tris * rays_sub < 4e10 else 2
        rc, _, sm2, am2, st2 = O.light_pass(sb, gb, frame, blue_noise, world, extra_lights=extra, exhaustive=mode,
                                            row_list=sub, x_range=xr, shadow_words=sw, ao_words=aw)
        assert rc == 0
        diff = popcount(sm2[sub, xr[0]:xr[1]] ^ sm[sub, xr[0]:xr[1]]) + popcount(am2[sub, xr[0]:xr[1]] ^ am[sub, xr[0]:xr[1]])
        res["bvh2_check"] = {"mode": "every triangle of every instance" if mode == 1 else
                             "every instance box, every triangle inside (no hierarchy)",
                             "rays": int(st2.rays), "rays_differ": int(diff)}

    if taa:
        hist = rt.read(R.IMG_HISTORY)
        rt.taa_pass(True)
        resolved = rt.read(R.IMG_LIGHT)
        err = 0.0
        groups = [int(y) for y in np.sort(rng.choice(np.arange(1, max(h - 5, 2)), size=4, replace=False))]
        for y0 in groups:
            want_t = O.taa_pass(sb, light, hist, gb.depth, True, rows=(y0, y0 + 4))
            a, b = resolved[y0:y0 + 4], want_t[y0:y0 + 4]
            f = np.isfinite(b)
            err = max(err, float(np.abs(a[f] - b[f]).max()))
        res["taa_max_abs"] = err
        res["taa_rows"] = 4 * len(groups)
    res["seconds"] = time.perf_counter() - t0
    return res


def assert_parity(res):
    assert res["agree"] >= MASK_AGREE, res
    assert res["max_abs_mask_identical_pixels"] <= RADIANCE_TOL or res["max_rel_mask_identical_pixels"] <= REL_TOL, res
    assert res["max_abs"] <= RADIANCE_TOL or res["psnr_db"] >= 50.0, res
    if "bvh2_check" in res:
        assert res["bvh2_check"]["rays_differ"] == 0 and res["bvh2_check"]["rays"] > 0, res
    if "taa_max_abs" in res:
        assert res["taa_max_abs"] <= TAA_TOL, res

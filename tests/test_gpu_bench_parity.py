"""Parity at the benchmark's own scale: the configurations bench.py times -- C2, C3 (+ the unique-BLAS
variant), C4, C5 of SURVEY.md section 8(d) -- are loaded exactly as bench.py loads them (scenes.write_project -> .luz
loader -> GPUScene::AddAssets -> RenderFrame), stepped, and the frame the PRODUCT kernels produce (set_debug(0):
k_shadow_hints + the compacting ray kernels + k_light_shade + k_taa) is compared with the CPU oracle on a seeded sample
of 64 full-width rows: visibility masks >= 99.9 % of the rays, radiance max-abs <= 1e-3 (or PSNR >= 50 dB), TAA <= 1e-4.
The BVH2 traverser the oracle needs at this size is itself re-validated against the hierarchy-free oracle on a
sub-sample of the same rows.  Covers what the small parity scenes cannot: two AO mask words (64 spp, C5), 256 lights
through extra_lights at 4K (C4), 10 288 instances (> LUZ_MAX_MODELS), a TLAS after 30 refits (C2) and after per-frame
rebuilds (C5)."""
import numpy as np
import pytest

import bench_parity as BP
import scene_util as S
from luz_b200 import rt as R
from luz_b200 import workloads

pytestmark = pytest.mark.gpu


def run_config(rt_factory, tmp_path, config, variant=None, frames=2, n_rows=64):
    rt = rt_factory()
    bn = S.blue_noise()
    wl = workloads.Workload(rt, config, variant=variant, tmp=str(tmp_path))
    wl.upload(bn)
    rt.set_debug(0)
    wl.step(first=True)
    for _ in range(frames - 1):
        wl.step()
    rt.sync()
    res = BP.check_workload(wl, bn, n_rows=n_rows)
    print("%s: %d rays on %d rows, agree %.6f, max-abs %.3g (mask-identical pixels %.3g), PSNR %.1f dB, taa %.3g, "
          "bvh2 check %s, %.1f s" % (res["config"], res["rays"], res["rows"], res["agree"], res["max_abs"],
                                     res["max_abs_mask_identical_pixels"], res["psnr_db"], res.get("taa_max_abs", -1.0),
                                     res.get("bvh2_check"), res["seconds"]))
    BP.assert_parity(res)
    rt.close()
    return res, wl


def test_c2_after_30_refits(rt_factory, tmp_path):
    """1080p, 4 097 instances, 4 lights + 4 AO spp; every instance has turned for 30 frames, TLAS refit each frame."""
    res, wl = run_config(rt_factory, tmp_path, "c2", frames=31)
    assert res["rays"] > 300_000 and wl.animate == "refit"


def test_c3_instanced(rt_factory, tmp_path):
    """4K, 10 288 instances of 16 BLASes, 4 lights + 16 AO spp: the configuration the headline metric is quoted on."""
    res, wl = run_config(rt_factory, tmp_path, "c3", frames=2)
    assert len(wl.app.instances()) == 10288 and res["rays"] > 1_000_000


def test_c3_unique_blas(rt_factory, tmp_path):
    """The HBM-bound variant: 10 288 unique BLASes (~10 M unique triangles)."""
    res, wl = run_config(rt_factory, tmp_path, "c3", variant="unique", frames=2, n_rows=32)
    assert len(wl.app.meshes()) == 10288


def test_c4_256_lights(rt_factory, tmp_path):
    """4K, 256 lights (192 through extra_lights), 256 shadow rays per pixel = 8 shadow mask words."""
    res, wl = run_config(rt_factory, tmp_path, "c4", frames=2, n_rows=64)
    assert res["shadow_words"] == 8 and wl.app.light_count() == 256 and res["rays"] > 5_000_000


def test_c5_64_ao_spp_after_rebuild(rt_factory, tmp_path):
    """8K, 64 AO spp = two AO mask words, every instance translated per frame (TLAS rebuild per frame)."""
    res, wl = run_config(rt_factory, tmp_path, "c5", frames=3, n_rows=64)
    assert res["ao_words"] == 2 and wl.animate == "rebuild" and res["rays"] > 5_000_000

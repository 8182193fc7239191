"""GPU test of one whole RenderFrame through the C++ host mirror (libluzhost.so -> libluzrt.so) with the SURVEY 8(f)
rank 4 passes switched on: shadow-map shadows (scene.shadowType == 2), a screen-space volumetric light, TAA.  The
frame must equal the oracle's light.frag -> screenSpaceVolumetricLight.comp -> taa.comp chain on the same G-buffer
and the same (device-rendered, separately oracle-checked) shadow map."""
import gzip
import json
import os
import shutil

import numpy as np
import pytest

import oracle_api as O
import scene_util as S
from luz_b200 import host as H
from luz_b200 import rt as R

pytestmark = pytest.mark.gpu


def test_render_frame_shadow_map_and_volumetric(rt_factory, tmp_path):
    with open(os.path.join(S.GOLDEN, "default.luz")) as f:
        doc = json.load(f)
    for sc in doc["scenes"].values():
        sc["shadowType"] = 2
        for n in sc["nodes"]:
            if n.get("type") == 7:
                n["volumetricType"] = 1
                n["volumetricScreenSamples"] = 48
                n["shadowMapFar"] = 60.0
    (tmp_path / "p.luz").write_text(json.dumps(doc))
    with gzip.open(os.path.join(S.GOLDEN, "default.luzbin.gz"), "rb") as f:
        (tmp_path / "p.luzbin").write_bytes(f.read())
    w, h = 480, 270
    bn = S.blue_noise()
    rt = rt_factory()
    app = H.LuzHost(rt)
    app.load_project(str(tmp_path / "p.luz"), str(tmp_path / "p.luzbin"))
    app.set_extent(w, h)
    app.add_assets()
    app.scene_settings(ao_samples=2)
    rt.set_blue_noise(bn)
    rt.set_debug(0)
    frame = app.frame_count
    app.render_frame(H.FRAME_OPAQUE)
    resolved = rt.read(R.IMG_HISTORY)  # SwapLightHistory: the resolved frame is now the history
    t = rt.read(R.TIMINGS)
    assert t.shadow_map_ms > 0.0 and t.volumetric_ms > 0.0 and t.light_ms > 0.0 and t.taa_ms > 0.0

    sb = app.scene_block()
    assert sb.shadow_type == 2 and sb.num_lights == 1 and sb.lights[0].shadow_map != -1
    assert sb.lights[0].num_shadow_samples == 0 and sb.lights[0].volumetric_type == 1  # GPUScene.cpp:248, :252
    world = O.World(app.meshes(), app.instances())
    gb = O.GBuffer(w, h)
    for sel, arr in ((R.GBUF_ALBEDO, gb.albedo), (R.GBUF_NORMAL, gb.normal), (R.GBUF_MATERIAL, gb.material),
                     (R.GBUF_EMISSION, gb.emission), (R.GBUF_DEPTH, gb.depth)):
        rt.read(sel, out=arr)
    smap = rt.read_shadow_map(0, 1024, 6)  # scene->shadowResolution (AssetManager.hpp:287)
    ref_map = O.shadow_map_pass(sb.lights[0], world, 1024)
    close = np.abs(smap - ref_map) <= 1e-5 * np.maximum(np.abs(ref_map), 1e-3)
    assert float(close.mean()) >= 0.999 and float((ref_map < 1.0).mean()) > 0.01
    with O.BoundShadowMaps({0: smap}, 1):
        rc, light, _, _, st = O.light_pass(sb, gb, frame, bn, world, exhaustive=True)
    assert rc == 0
    lit = O.volumetric_screen_pass(sb, light, gb.depth, bn, frame)
    assert float(np.abs(lit - light).max()) > 1e-4
    ref = O.taa_pass(sb, lit, lit, gb.depth, True)  # first frame: history := current
    err = np.abs(resolved - ref).max(axis=-1)
    assert float((err <= 1e-3).mean()) >= 0.999, float((err <= 1e-3).mean())
    # the point light's map really shadows: the slab under the cube receives less light than its open part
    assert float(ref[..., :3].max()) > 10.0 * float(np.median(ref[..., :3][ref[..., :3] > 0]))

"""CPU tests: the C++ host mirror (libluzhost.so) against golden vectors produced by the REFERENCE'S
OWN compiled host code (oracle/ref_dump.cpp -> tests/golden/ref_host_1280x720.json)."""
import ctypes as C
import gzip
import json
import os
import shutil

import numpy as np
import pytest

import scene_util as S
from luz_b200 import host as H
from luz_b200 import wire

GOLDEN = S.GOLDEN


@pytest.fixture(scope="module")
def project(tmp_path_factory):
    d = tmp_path_factory.mktemp("proj")
    shutil.copy(os.path.join(GOLDEN, "default.luz"), d / "default.luz")
    with gzip.open(os.path.join(GOLDEN, "default.luzbin.gz"), "rb") as f:
        (d / "default.luzbin").write_bytes(f.read())
    return str(d / "default.luz"), str(d / "default.luzbin")


@pytest.fixture(scope="module")
def g():
    return S.golden_json()


def f32(a):
    return np.asarray(a, np.float32)


def test_wire_layout_matches_reference_headers(g):
    lay = g["layout"]
    assert C.sizeof(wire.LightBlock) == lay["LightBlock"]
    assert C.sizeof(wire.ModelBlock) == lay["ModelBlock"]
    assert C.sizeof(wire.SceneBlock) == lay["SceneBlock"]
    pairs = {
        "SceneBlock.ambientLightColor": wire.SceneBlock.ambient_light_color, "SceneBlock.proj": wire.SceneBlock.proj,
        "SceneBlock.view": wire.SceneBlock.view, "SceneBlock.viewProj": wire.SceneBlock.view_proj,
        "SceneBlock.prevViewProj": wire.SceneBlock.prev_view_proj, "SceneBlock.inverseProj": wire.SceneBlock.inverse_proj,
        "SceneBlock.inverseView": wire.SceneBlock.inverse_view, "SceneBlock.jitter": wire.SceneBlock.jitter,
        "SceneBlock.prevJitter": wire.SceneBlock.prev_jitter, "SceneBlock.camPos": wire.SceneBlock.cam_pos,
        "SceneBlock.numLights": wire.SceneBlock.num_lights, "SceneBlock.aoMin": wire.SceneBlock.ao_min,
        "SceneBlock.aoMax": wire.SceneBlock.ao_max, "SceneBlock.exposure": wire.SceneBlock.exposure,
        "SceneBlock.aoNumSamples": wire.SceneBlock.ao_num_samples,
        "SceneBlock.blueNoiseTexture": wire.SceneBlock.blue_noise_texture, "SceneBlock.tlasRid": wire.SceneBlock.tlas_rid,
        "SceneBlock.shadowType": wire.SceneBlock.shadow_type,
        "LightBlock.position": wire.LightBlock.position, "LightBlock.direction": wire.LightBlock.direction,
        "LightBlock.type": wire.LightBlock.type, "LightBlock.numShadowSamples": wire.LightBlock.num_shadow_samples,
        "LightBlock.radius": wire.LightBlock.radius, "LightBlock.viewProj": wire.LightBlock.view_proj,
        "LightBlock.zFar": wire.LightBlock.z_far, "ModelBlock.color": wire.ModelBlock.color,
        "ModelBlock.roughness": wire.ModelBlock.roughness, "ModelBlock.colorMap": wire.ModelBlock.color_map,
    }
    for k, fld in pairs.items():
        assert fld.offset == lay[k], k
    assert lay["LightConstants"] == 32 and lay["PostProcessingConstants"] == 52 and lay["MeshVertex"] == 48


def test_halton_bit_exact(g):
    for i in range(32):
        assert np.float32(H.halton(i, 2)) == np.float32(g["halton2"][i])
        assert np.float32(H.halton(i, 3)) == np.float32(g["halton3"][i])


def test_compose_transform_and_inverse_bit_exact(g):
    for t in g["transforms"]:
        m = H.compose_transform(t["pos"], t["rot"], t["scale"], t["parent"])
        assert np.array_equal(m, f32(t["mat"])), (m, t["mat"])
        inv = H.mat4_inverse(t["mat"])
        ref = f32(t["inverse"])
        assert np.array_equal(inv, ref) or np.allclose(inv, ref, rtol=2e-6, atol=1e-9)


def test_load_default_project_matches_reference_loader(project, g):
    app = H.LuzHost(None)
    app.load_project(*project)
    st = app.get_scene_settings()
    assert st["lightSamples"] == g["scene"]["lightSamples"] and st["aoSamples"] == g["scene"]["aoSamples"]
    assert st["shadowType"] == g["scene"]["shadowType"]
    assert st["taaEnabled"] == bool(g["scene"]["taaEnabled"]) and st["taaReconstruct"] == bool(g["scene"]["taaReconstruct"])
    ao_min, ao_max, exposure, ambient = g["scene"]["aoMin_aoMax_exposure_ambientLight"]
    assert (np.float32(st["aoMin"]), np.float32(st["aoMax"])) == (np.float32(ao_min), np.float32(ao_max))
    assert np.float32(st["exposure"]) == np.float32(exposure) and np.float32(st["ambientLight"]) == np.float32(ambient)
    assert app.mesh_node_count() == len(g["meshNodes"]) and app.light_count() == len(g["lights"])
    app.set_extent(1280, 720, create_images=False)
    app.add_assets()
    app.update_resources()
    insts = app.instances()
    meshes = app.meshes()
    for (mi, mat, ci), node in zip(insts, g["meshNodes"]):
        assert np.array_equal(mat, f32(node["world"]))
        v, idx = meshes[mi]
        assert v.shape[0] == node["vertexCount"] and idx.size == node["indexCount"]
        assert np.array_equal(v[0], f32(node["vertex0"]))
        assert np.array_equal(idx, np.asarray(node["indices"], np.uint32))
    models, n = app.models()
    assert n == len(g["meshNodes"])
    for i, node in enumerate(g["meshNodes"]):
        assert np.array_equal(f32(models[i].model_mat[:]), f32(node["world"]))
        assert np.array_equal(f32(models[i].color[:]), f32(node["color"]))
        assert np.float32(models[i].metallic) == np.float32(node["metallic_roughness"][0])
        assert np.float32(models[i].roughness) == np.float32(node["metallic_roughness"][1])
        assert (models[i].color_map >= 0) == (node["colorMapUuid"] != 0)


def test_scene_block_sequence_matches_reference_camera(project, g):
    """GPUScene::UpdateResources for 40 frames: jitter cycle, matrices, light block inputs."""
    app = H.LuzHost(None)
    app.load_project(*project)
    app.set_extent(1280, 720, create_images=False)
    app.scene_settings(light_samples=1, ao_samples=1)
    frames = g["camera"]["frames"]
    prev_vp = None
    for f, fr in enumerate(frames):
        app.update_resources()
        sb = app.scene_block()
        for name, fld in (("proj", sb.proj), ("view", sb.view), ("viewProj", sb.view_proj),
                          ("inverseProj", sb.inverse_proj), ("inverseView", sb.inverse_view)):
            got, ref = f32(fld[:]), f32(fr[name])
            assert np.array_equal(got, ref) or np.allclose(got, ref, rtol=3e-6, atol=1e-9), (f, name, got, ref)
        assert np.array_equal(f32(sb.jitter[:]), f32(fr["jitter"])), f
        assert np.array_equal(f32(sb.cam_pos[:]), f32(fr["camPos"]))
        if f == 0:  # defined first frame: no motion
            assert np.array_equal(f32(sb.prev_view_proj[:]), f32(sb.view_proj[:]))
            assert np.array_equal(f32(sb.prev_jitter[:]), f32(sb.jitter[:]))
        else:
            assert np.array_equal(f32(sb.prev_view_proj[:]), prev_vp)
            assert np.array_equal(f32(sb.prev_jitter[:]), f32(fr["prevJitter"]))
        prev_vp = f32(sb.view_proj[:])
        assert sb.num_lights == len(g["lights"])
        for i, l in enumerate(g["lights"]):
            lb = sb.lights[i]
            assert np.array_equal(f32(lb.position[:]), f32(l["position"]))
            assert np.array_equal(f32(lb.direction[:]), f32(l["direction"]))
            assert np.array_equal(f32([lb.inner_angle, lb.outer_angle]), f32(l["inner_outer_radians"]))
            assert np.array_equal(f32([lb.intensity, lb.radius, lb.z_far]), f32(l["intensity_radius_zfar"]))
            assert lb.type == l["type"] and lb.num_shadow_samples == 1 and lb.shadow_map == -1
        assert sb.ao_num_samples == 1 and sb.shadow_type == 1
    # jitter is a 16-cycle of Halton(2,3)
    assert frames[0]["jitter"] == frames[16]["jitter"] and frames[1]["jitter"] != frames[0]["jitter"]


def test_negative_ao_samples_and_shadow_type(project):
    app = H.LuzHost(None)
    app.load_project(*project)
    app.set_extent(640, 360, create_images=False)
    app.scene_settings(shadow_type=0)
    app.update_resources()
    sb = app.scene_block()
    assert sb.shadow_type == 0 and sb.lights[0].num_shadow_samples == 0  # GPUScene.cpp:248


def test_project_round_trip(project, tmp_path):
    """SaveProject -> LoadProject reproduces the same scene (the .luz writer, section 8f rank 3)."""
    app = H.LuzHost(None)
    app.load_project(*project)
    out_j, out_b = str(tmp_path / "rt.luz"), str(tmp_path / "rt.luzbin")
    app.save_project(out_j, out_b)
    a = json.load(open(project[0]))
    b = json.load(open(out_j))
    assert a["initialScene"] == b["initialScene"]
    assert set(a["scenes"].keys()) == set(b["scenes"].keys())

    def strip_blobs(x):
        if isinstance(x, dict):
            return {k: ("blob" if isinstance(v, dict) and set(v.keys()) == {"offset", "size"} else strip_blobs(v))
                    for k, v in x.items()}
        if isinstance(x, list):
            return [strip_blobs(v) for v in x]
        return x
    assert strip_blobs(a["scenes"]) == strip_blobs(b["scenes"])
    key = lambda o: o["uuid"]
    assert sorted(map(strip_blobs, a["assets"]), key=key) == sorted(map(strip_blobs, b["assets"]), key=key)
    app2 = H.LuzHost(None)
    app2.load_project(out_j, out_b)
    m1, m2 = app.meshes(), app2.meshes()
    assert len(m1) == len(m2)
    for (v1, i1), (v2, i2) in zip(m1, m2):
        assert np.array_equal(v1, v2) and np.array_equal(i1, i2)
    for t1, t2 in zip(app.textures(), app2.textures()):
        assert np.array_equal(t1, t2)


def test_loader_errors(tmp_path):
    app = H.LuzHost(None)
    with pytest.raises(H.HostError):
        app.load_project(str(tmp_path / "missing.luz"), str(tmp_path / "missing.luzbin"))
    (tmp_path / "bad.luz").write_text("{ not json")
    (tmp_path / "bad.luzbin").write_bytes(b"")
    with pytest.raises(H.HostError):
        app.load_project(str(tmp_path / "bad.luz"), str(tmp_path / "bad.luzbin"))
    (tmp_path / "oob.luz").write_text(json.dumps({"assets": [{"type": 2, "name": "m", "uuid": 5,
                                                             "vertices": {"offset": 0, "size": 4800},
                                                             "indices": {"offset": 0, "size": 12}}],
                                                  "scenes": {}, "initialScene": 0}))
    (tmp_path / "oob.luzbin").write_bytes(b"\0" * 16)
    with pytest.raises(H.HostError):
        app.load_project(str(tmp_path / "oob.luz"), str(tmp_path / "oob.luzbin"))

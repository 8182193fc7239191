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
            assert lb.type == l["type"] and lb.num_shadow_samples == 1 and lb.shadow_map == i  # 'has a map' (GPUScene.cpp:329)
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


def test_shadow_matrix_primitives_match_glm(project, g):
    """The glm / camera calls GPUScene.cpp:266-311 builds light.viewProj[] from, against the reference's own glm
    (oracle/ref_dump.cpp 'shadowGlm'): perspective with zNear = 0, lookAt with the cube-face ups, ortho with
    zNear > zFar, CameraNode::GetProj(near, far / range), products and the inverse."""
    sg = g["shadowGlm"]
    rad90 = float(np.float32(90.0) * np.float32(0.01745329251994329576923690768489))
    assert np.array_equal(H.perspective(rad90, 1.0, 0.0, 2000.0), f32(sg["persp90_far2000"]))
    assert np.array_equal(H.perspective(rad90, 1.0, 0.0, 37.5), f32(sg["persp90_far37"]))
    pos = f32(sg["pos"])
    for f, (axis, up) in enumerate(S.CUBE_FACES):
        v = H.look_at(pos, pos + f32(axis), up)
        assert np.array_equal(v, f32(sg["faces"][f]["lookAt"])), f
        assert np.array_equal(H.mat4_mul(sg["persp90_far2000"], v), f32(sg["faces"][f]["viewProj"])), f
    assert np.array_equal(H.ortho(-3.25, 5.5, -2.125, 7.75, 11.5, -9.25), f32(sg["ortho"]))
    app = H.LuzHost(None)
    app.load_project(*project)
    app.set_extent(1280, 720, create_images=False)
    near, far = g["camera"]["zoom_far_near_fov_w_h"][2], g["camera"]["zoom_far_near_fov_w_h"][1]
    cp = app.camera_proj(near, float(np.float32(far) / np.float32(3.0)))
    assert np.array_equal(cp, f32(sg["camProj_far_over_3"]))
    inv = H.mat4_inverse(H.mat4_mul(sg["camProj_far_over_3"], sg["camView"]))
    ref = f32(sg["inverse_camProjView"])
    assert np.array_equal(inv, ref) or np.allclose(inv, ref, rtol=3e-6, atol=1e-9)
    c, fr = f32([0.5, -1.25, 2.0]), f32([0.3, -1.0, 0.2])
    assert np.array_equal(H.look_at(c + fr, c, (0, 1, 0)), f32(sg["lookAt_front"]))


def test_light_view_proj_like_gpuscene_266_311(project, g):
    """LightBlock.viewProj[] of the host mirror: the default project's point light gets exactly the six cube-face
    matrices glm produces; a directional light gets the frustum-fitted orthographic matrix, checked against an
    independent float64 evaluation of GPUScene.cpp:278-310."""
    app = H.LuzHost(None)
    app.load_project(*project)
    app.set_extent(1280, 720, create_images=False)
    app.update_resources()
    sb = app.scene_block()
    lb = sb.lights[0]
    assert lb.type == wire.LIGHT_POINT and lb.shadow_map != -1 and lb.z_far == 2000.0
    assert np.array_equal(f32(lb.position[:]), f32(g["shadowGlm"]["pos"]))
    for f in range(6):
        assert np.array_equal(f32(lb.view_proj[f][:]), f32(g["shadowGlm"]["faces"][f]["viewProj"])), f
    # the six matrices map a direction to the (sc, tc) / |ma| of the Vulkan cube-face table (the property
    # luzrt_shadow_map_pass relies on for its texel-centre rays)
    rng = np.random.default_rng(3)
    pos = np.array(lb.position[:], np.float64)
    for f in range(6):
        vp = np.array(lb.view_proj[f][:], np.float64).reshape(4, 4).T
        for _ in range(50):
            sc, tc = rng.uniform(-1, 1, 2)
            d = [(1, -tc, -sc), (-1, -tc, sc), (sc, 1, tc), (sc, -1, -tc), (sc, -tc, 1), (-sc, -tc, -1)][f]
            c = vp @ np.array([*(pos + 5.0 * np.array(d)), 1.0])
            assert c[3] > 0 and abs(c[0] / c[3] - sc) < 1e-5 and abs(c[1] / c[3] - tc) < 1e-5
            assert abs(c[2] / c[3] - 1.0) < 1e-6  # zNear = 0: every fragment sits at depth 1, hence shadowMap.frag:15


def _expected_ortho_view_proj(cam_proj, view, front, rng_r):
    inv = np.linalg.inv(cam_proj @ view)
    corners = []
    for i in range(2):
        for j in range(2):
            for k in range(2):
                p = inv @ np.array([2.0 * i - 1, 2.0 * j - 1, 2.0 * k - 1, 1.0])
                corners.append(p / p[3])
    centre = np.mean([c[:3] for c in corners], axis=0)
    eye, up = centre + front, np.array([0.0, 1.0, 0.0])
    fwd = (centre - eye) / np.linalg.norm(centre - eye)
    s = np.cross(fwd, up)
    s /= np.linalg.norm(s)
    u = np.cross(s, fwd)
    lv = np.eye(4)
    lv[0, :3], lv[1, :3], lv[2, :3] = s, u, -fwd
    lv[:3, 3] = [-s @ eye, -u @ eye, fwd @ eye]
    mn = corners[0][:3].copy()  # seeded in world space (GPUScene.cpp:299-300)
    mx = mn.copy()
    for c in corners:
        q = (lv @ c)[:3]
        mn, mx = np.minimum(q, mn), np.maximum(q, mx)
    mn[2] = mn[2] * rng_r if mn[2] < 0 else mn[2] / rng_r
    mx[2] = mx[2] / rng_r if mx[2] < 0 else mx[2] * rng_r
    l, r_, b, t, n, f = mn[0], mx[0], mn[1], mx[1], mx[2], mn[2]
    o = np.eye(4)
    o[0, 0], o[1, 1], o[2, 2] = 2 / (r_ - l), 2 / (t - b), -1 / (f - n)
    o[0, 3], o[1, 3], o[2, 3] = -(r_ + l) / (r_ - l), -(t + b) / (t - b), -n / (f - n)
    return o @ lv


def test_directional_light_view_proj(project, g, tmp_path):
    import json as J
    with open(project[0]) as f:
        doc = J.load(f)
    for sc in doc["scenes"].values():
        for n in sc["nodes"]:
            if n.get("type") == 7:
                n["lightType"] = 2
                n["rotation"] = [25.0, 40.0, 10.0]
                n["shadowMapRange"] = 3.0
    path = tmp_path / "dir.luz"
    path.write_text(J.dumps(doc))
    app = H.LuzHost(None)
    app.load_project(str(path), project[1])
    app.set_extent(1280, 720, create_images=False)
    app.camera_use_jitter(False)
    app.update_resources()
    sb = app.scene_block()
    lb = sb.lights[0]
    assert lb.type == wire.LIGHT_DIRECTIONAL
    got = np.array(lb.view_proj[0][:], np.float64).reshape(4, 4).T
    near, far = g["camera"]["zoom_far_near_fov_w_h"][2], g["camera"]["zoom_far_near_fov_w_h"][1]
    cam_proj = np.array(app.camera_proj(near, float(np.float32(far) / np.float32(3.0))), np.float64).reshape(4, 4).T
    view = np.array(sb.view[:], np.float64).reshape(4, 4).T
    exp = _expected_ortho_view_proj(cam_proj, view, np.array(lb.direction[:], np.float64), 3.0)
    assert got[3].tolist() == [0.0, 0.0, 0.0, 1.0]  # orthographic: what luzrt_shadow_map_pass requires
    assert np.allclose(got, exp, rtol=2e-3, atol=2e-4), (got, exp)

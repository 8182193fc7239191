"""Test-side scene construction: the reference's default project from tests/golden (matrices come
from the REFERENCE'S OWN compiled host code, ref_host_*.json) and small seeded synthetic scenes.
Numpy fp32 camera helpers here are harness code for synthetic cases only."""
import gzip
import json
import os

import numpy as np

from luz_b200 import wire

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_json(w=1280, h=720):
    with open(os.path.join(GOLDEN, "ref_host_%dx%d.json" % (w, h))) as f:
        return json.load(f)


def blue_noise():
    return np.fromfile(os.path.join(GOLDEN, "blue_noise_256.rgba"), dtype=np.uint8).reshape(256, 256, 4)


def load_default_project():
    """Raw parse of default.luz/.luzbin (harness-side; the product loader is luz_b200/host)."""
    with open(os.path.join(GOLDEN, "default.luz")) as f:
        j = json.load(f)
    with gzip.open(os.path.join(GOLDEN, "default.luzbin.gz"), "rb") as f:
        blob = f.read()
    assets = {}
    for a in j["assets"]:
        if a["type"] == 2:
            v = np.frombuffer(blob, np.float32, a["vertices"]["size"] // 4, a["vertices"]["offset"]).reshape(-1, 12)
            i = np.frombuffer(blob, np.uint32, a["indices"]["size"] // 4, a["indices"]["offset"])
            assets[a["uuid"]] = ("mesh", v.copy(), i.copy())
        elif a["type"] == 1:
            t = np.frombuffer(blob, np.uint8, a["data"]["size"], a["data"]["offset"]).reshape(a["height"], a["width"], 4)
            assets[a["uuid"]] = ("texture", t.copy())
        elif a["type"] == 3:
            assets[a["uuid"]] = ("material", a)
    return j, assets


def set_mat(dst, m16):
    for i in range(16):
        dst[i] = float(m16[i])


def default_scene(frame=0, light_samples=1, ao_samples=1, w=1280, h=720):
    """Returns dict(scene=SceneBlock, meshes, instances, models, textures) for config C1 using the
    golden camera sequence at frame `frame` (frame 0: prevViewProj = viewProj, prevJitter = jitter)."""
    g = golden_json(w, h)
    j, assets = load_default_project()
    scene_json = j["scenes"][str(j["initialScene"])]
    sb = wire.SceneBlock()
    fr = g["camera"]["frames"][frame]
    prev = g["camera"]["frames"][frame - 1] if frame > 0 else fr
    set_mat(sb.proj, fr["proj"])
    set_mat(sb.view, fr["view"])
    set_mat(sb.view_proj, fr["viewProj"])
    set_mat(sb.prev_view_proj, prev["viewProj"])
    set_mat(sb.inverse_proj, fr["inverseProj"])
    set_mat(sb.inverse_view, fr["inverseView"])
    sb.jitter[0], sb.jitter[1] = fr["jitter"]
    pj = fr["prevJitter"] if frame > 0 else fr["jitter"]
    sb.prev_jitter[0], sb.prev_jitter[1] = pj
    for k in range(3):
        sb.cam_pos[k] = fr["camPos"][k]
        sb.ambient_light_color[k] = g["scene"]["ambientLightColor"][k]
    ao_min, ao_max, exposure, ambient = g["scene"]["aoMin_aoMax_exposure_ambientLight"]
    sb.ambient_light_intensity = ambient
    sb.ao_min, sb.ao_max, sb.exposure = ao_min, ao_max, exposure
    sb.ao_num_samples = ao_samples
    sb.shadow_type = g["scene"]["shadowType"]
    sb.num_lights = len(g["lights"])
    for i, l in enumerate(g["lights"]):
        lb = sb.lights[i]
        for k in range(3):
            lb.color[k] = l["color"][k]
            lb.position[k] = l["position"][k]
            lb.direction[k] = l["direction"][k]
        lb.intensity, lb.radius, lb.z_far = l["intensity_radius_zfar"]
        lb.inner_angle, lb.outer_angle = l["inner_outer_radians"]
        lb.type = l["type"]
        lb.num_shadow_samples = light_samples if sb.shadow_type == 1 else 0
        lb.shadow_map = -1
    # meshes in first-use order, instances in GPUScene::UpdateResources order
    mesh_ids, meshes, textures, tex_ids = {}, [], [], {}
    instances = []
    models = (wire.ModelBlock * len(g["meshNodes"]))()
    for n, node in enumerate(g["meshNodes"]):
        mu = node["meshUuid"]
        if mu not in mesh_ids:
            mesh_ids[mu] = len(meshes)
            meshes.append((assets[mu][1], assets[mu][2]))
        instances.append((mesh_ids[mu], np.array(node["world"], np.float32), n))
        mb = models[n]
        set_mat(mb.model_mat, node["world"])
        for k in range(4):
            mb.color[k] = node["color"][k]
        for k in range(3):
            mb.emission[k] = node["emission"][k]
        mb.metallic, mb.roughness = node["metallic_roughness"]
        mb.ao_map = mb.normal_map = mb.emission_map = mb.metallic_roughness_map = -1
        cu = node.get("colorMapUuid", 0)
        if cu:
            if cu not in tex_ids:
                tex_ids[cu] = len(textures)
                textures.append(assets[cu][1])
            mb.color_map = tex_ids[cu]
        else:
            mb.color_map = -1
    return dict(scene=sb, meshes=meshes, instances=instances, models=models, textures=textures, width=w, height=h)


# ---- numpy fp32 helpers for synthetic scenes (harness only) ------------------------------------------
def perspective_vk(fovy_deg, aspect, near, far):
    f = np.float32(1.0) / np.tan(np.float32(np.radians(fovy_deg)) / np.float32(2))
    m = np.zeros((4, 4), np.float32)  # m[col][row]
    m[0][0] = f / np.float32(aspect)
    m[1][1] = -f
    m[2][2] = np.float32(far) / np.float32(near - far)
    m[2][3] = -1.0
    m[3][2] = -np.float32(far * near) / np.float32(far - near)
    return m


def look_at(eye, center, up=(0, 1, 0)):
    eye, center, up = (np.asarray(v, np.float32) for v in (eye, center, up))
    f = center - eye
    f /= np.linalg.norm(f)
    s = np.cross(f, up)
    s /= np.linalg.norm(s)
    u = np.cross(s, f)
    m = np.eye(4, dtype=np.float32)  # m[col][row]
    m[0][0], m[1][0], m[2][0] = s
    m[0][1], m[1][1], m[2][1] = u
    m[0][2], m[1][2], m[2][2] = -f
    m[3][0], m[3][1], m[3][2] = -np.dot(s, eye), -np.dot(u, eye), np.dot(f, eye)
    return m


def perspective_zo(fovy_deg, aspect, near, far):
    """glm::perspective (RH, depth 0..1, GLM_FORCE_DEPTH_ZERO_TO_ONE) without Luz's y flip, stored m[col][row]."""
    m = perspective_vk(fovy_deg, aspect, near, far)
    m[1][1] = -m[1][1]
    return m


def ortho_zo(left, right, bottom, top, near, far):
    """glm::ortho (RH, depth 0..1), stored m[col][row]."""
    m = np.eye(4, dtype=np.float32)
    m[0][0] = 2.0 / (right - left)
    m[1][1] = 2.0 / (top - bottom)
    m[2][2] = -1.0 / (far - near)
    m[3][0] = -(right + left) / (right - left)
    m[3][1] = -(top + bottom) / (top - bottom)
    m[3][2] = -near / (far - near)
    return m


CUBE_FACES = [((1, 0, 0), (0, -1, 0)), ((-1, 0, 0), (0, -1, 0)), ((0, 1, 0), (0, 0, 1)), ((0, -1, 0), (0, 0, -1)),
              ((0, 0, 1), (0, -1, 0)), ((0, 0, -1), (0, -1, 0))]  # GPUScene.cpp:270-275


def set_light_shadow_matrices(lb, centre=(0, 0, 0), half_extent=12.0, z_far=60.0):
    """Fills light.viewProj[] / zFar like GPUScene.cpp:266-311 does: six perspective(90, 1, 0, zFar) * lookAt
    faces for a point light; for the others one orthographic matrix looking along the light's direction (the
    reference fits it to the camera frustum; a fixed box around `centre` exercises the same code path)."""
    pos = np.array([lb.position[k] for k in range(3)], np.float32)
    lb.z_far = z_far
    if lb.type == wire.LIGHT_POINT:
        proj = perspective_zo(90.0, 1.0, 0.0, z_far)
        for f, (axis, up) in enumerate(CUBE_FACES):
            set_mat(lb.view_proj[f], colmajor_mul(proj, look_at(pos, pos + np.array(axis, np.float32), up)).reshape(16))
    else:
        front = np.array([lb.direction[k] for k in range(3)], np.float32)
        c = np.array(centre, np.float32)
        view = look_at(c + front, c, (0, 1, 0))  # GPUScene.cpp:298 looks from centre + front towards the centre
        h = half_extent
        set_mat(lb.view_proj[0], colmajor_mul(ortho_zo(-h, h, -h, h, 3 * h, -3 * h), view).reshape(16))


def colmajor_mul(a, b):
    """a, b stored as m[col][row]; returns a*b in the same storage."""
    return (a.T @ b.T).T.astype(np.float32)


def colmajor_inv(a):
    return np.linalg.inv(a.T.astype(np.float64)).T.astype(np.float32)


def trs(pos, yaw_deg=0.0, scale=(1, 1, 1)):
    c, s = np.cos(np.radians(yaw_deg)), np.sin(np.radians(yaw_deg))
    r = np.array([[c, 0, s, 0], [0, 1, 0, 0], [-s, 0, c, 0], [0, 0, 0, 1]], np.float64)
    sc = np.diag([scale[0], scale[1], scale[2], 1.0])
    t = np.eye(4)
    t[:3, 3] = pos
    m = t @ r @ sc  # row-major math
    return m.T.astype(np.float32).reshape(16)  # column-major storage


def unit_cube():
    """24-vertex cube (+-1), 48-byte vertices: pos3 normal3 tangent4 uv2."""
    faces = [((1, 0, 0), (0, 1, 0), (0, 0, 1)), ((-1, 0, 0), (0, 0, 1), (0, 1, 0)), ((0, 1, 0), (0, 0, 1), (1, 0, 0)),
             ((0, -1, 0), (1, 0, 0), (0, 0, 1)), ((0, 0, 1), (1, 0, 0), (0, 1, 0)), ((0, 0, -1), (0, 1, 0), (1, 0, 0))]
    v, idx = [], []
    for n, a, b in faces:
        n, a, b = (np.array(t, np.float32) for t in (n, a, b))
        base = len(v)
        for (sa, sb_) in ((-1, -1), (1, -1), (1, 1), (-1, 1)):
            p = n + sa * a + sb_ * b
            v.append(np.concatenate([p, n, a, [1.0], [(sa + 1) / 2, (sb_ + 1) / 2]]).astype(np.float32))
        idx += [base, base + 1, base + 2, base, base + 2, base + 3]
    return np.array(v, np.float32), np.array(idx, np.uint32)


def synthetic_scene(w, h, grid=4, n_lights=3, light_samples=1, ao_samples=2, seed=11, eye=(9, 7, 11), shadow_type=1,
                    mixed_materials=True):
    """grid x grid cubes on a slab, point/spot/directional lights; returns the same dict as default_scene."""
    rng = np.random.default_rng(seed)
    cube = unit_cube()
    meshes = [cube]
    instances, mats = [], []
    half = (grid - 1) * 1.5
    instances.append((0, trs((0, -0.05, 0), 0, (half + 3, 0.05, half + 3)), 0))
    for gz in range(grid):
        for gx in range(grid):
            yaw = float(rng.uniform(0, 90))
            s = float(rng.uniform(0.3, 0.6))
            instances.append((0, trs((gx * 3.0 - half, s, gz * 3.0 - half), yaw, (s, s, s)), len(instances)))
    models = (wire.ModelBlock * len(instances))()
    for i, (_, m, _) in enumerate(instances):
        mb = models[i]
        set_mat(mb.model_mat, m)
        col = rng.uniform(0.3, 1.0, 3) if mixed_materials else (0.8, 0.8, 0.8)
        for k in range(3):
            mb.color[k] = float(col[k])
        mb.color[3] = 1.0
        mb.metallic = float(rng.uniform(0, 1)) if mixed_materials else 0.0
        mb.roughness = float(rng.uniform(0.05, 1)) if mixed_materials else 0.5
        if mixed_materials and i % 5 == 3:
            mb.emission[0], mb.emission[1], mb.emission[2] = 0.2, 0.1, 0.0
        mb.ao_map = mb.color_map = mb.normal_map = mb.emission_map = mb.metallic_roughness_map = -1
    sb = wire.SceneBlock()
    proj = perspective_vk(60.0, w / h, 0.01, 1000.0)
    jit = np.array([0.3 / w, -0.2 / h], np.float32)
    tj = np.eye(4, dtype=np.float32)
    tj[3][0], tj[3][1] = jit
    projj = colmajor_mul(tj, proj)
    view = look_at(eye, (0, 0.3, 0))
    vp = colmajor_mul(projj, view)
    set_mat(sb.proj, projj.reshape(16))
    set_mat(sb.view, view.reshape(16))
    set_mat(sb.view_proj, vp.reshape(16))
    # previous frame: slightly moved camera and different jitter, so TAA reprojects for real
    view_p = look_at((eye[0] + 0.05, eye[1], eye[2] - 0.03), (0, 0.3, 0))
    tjp = np.eye(4, dtype=np.float32)
    tjp[3][0], tjp[3][1] = -0.25 / w, 0.35 / h
    set_mat(sb.prev_view_proj, colmajor_mul(colmajor_mul(tjp, proj), view_p).reshape(16))
    set_mat(sb.inverse_proj, colmajor_inv(projj).reshape(16))
    set_mat(sb.inverse_view, colmajor_inv(view).reshape(16))
    sb.jitter[0], sb.jitter[1] = float(jit[0]), float(jit[1])
    sb.prev_jitter[0], sb.prev_jitter[1] = -0.25 / w, 0.35 / h
    for k in range(3):
        sb.cam_pos[k] = float(eye[k])
        sb.ambient_light_color[k] = 1.0
    sb.ambient_light_intensity = 0.05
    sb.ao_min, sb.ao_max, sb.exposure = 1e-4, 1.0, 2.0
    sb.ao_num_samples = ao_samples
    sb.shadow_type = shadow_type
    sb.num_lights = n_lights
    for i in range(n_lights):
        lb = sb.lights[i]
        kind = [wire.LIGHT_POINT, wire.LIGHT_SPOT, wire.LIGHT_DIRECTIONAL][i % 3]
        pos = np.array([rng.uniform(-half, half), rng.uniform(3, 6), rng.uniform(-half, half)], np.float32)
        d = np.array([rng.uniform(-0.4, 0.4), -1.0, rng.uniform(-0.4, 0.4)], np.float32)
        for k in range(3):
            lb.color[k] = float(rng.uniform(0.5, 1.0))
            lb.position[k] = float(pos[k])
            lb.direction[k] = float(d[k])
        lb.intensity = float(rng.uniform(5, 20)) if kind != wire.LIGHT_DIRECTIONAL else 1.5
        lb.inner_angle, lb.outer_angle = float(np.radians(60.0)), float(np.radians(50.0))
        lb.type = kind
        lb.num_shadow_samples = light_samples if shadow_type == 1 else 0
        lb.radius = float(rng.uniform(0.1, 0.6))
        lb.shadow_map = -1
    return dict(scene=sb, meshes=meshes, instances=instances, models=models, textures=[], width=w, height=h)


def make_rt_scene(rt, sc):
    """Uploads meshes/instances/textures of a scene dict into a LuzRT ctx; returns instance array."""
    blas = [rt.blas_create(v, i, stride=v.strides[0]) for (v, i) in sc["meshes"]]
    for t in sc["textures"]:
        rt.texture_create(t)
    inst = rt.make_instances([blas[m] for (m, _, _) in sc["instances"]], [mat for (_, mat, _) in sc["instances"]],
                             [ci for (_, _, ci) in sc["instances"]])
    rt.tlas_build(inst, len(sc["instances"]), 0)
    return blas, inst

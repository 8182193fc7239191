"""Synthetic benchmark scenes of BASELINE.json / SURVEY.md section 8(d), written as ordinary Luz
projects (.luz JSON + .luzbin blob) so that they enter the product exactly like a user's scene does:
through the .luz loader and GPUScene::UpdateResources of the host mirror.  Harness-side; seeded and
deterministic.  C1 is the reference's own assets/default.luz (tests/golden/)."""
import gzip
import json
import os
import shutil

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CONFIGS = {
    # name: (width, height, lightSamples, aoSamples, tlas mode per frame: None static / "refit" / "rebuild")
    "c1": dict(width=1280, height=720, light_samples=1, ao_samples=1, animate=None),
    "c2": dict(width=1920, height=1080, light_samples=1, ao_samples=4, animate="refit"),
    "c3": dict(width=3840, height=2160, light_samples=1, ao_samples=16, animate=None),
    "c4": dict(width=3840, height=2160, light_samples=1, ao_samples=0, animate=None),
    "c5": dict(width=7680, height=4320, light_samples=1, ao_samples=64, animate="rebuild"),
}


class _Uuid:
    def __init__(self, seed):
        self.rng = np.random.default_rng(seed)

    def __call__(self):
        return int(self.rng.integers(2 ** 61, 2 ** 62))


def _cube_mesh(tess=1, displace=0.0, phase=None):
    """Cube of half-size 1 with tess x tess quads per face; optional smooth radial displacement.
    Returns (vertices [n,12] float32: pos3 normal3 tangent4 uv2, indices uint32)."""
    faces = [((1, 0, 0), (0, 1, 0), (0, 0, 1)), ((-1, 0, 0), (0, 0, 1), (0, 1, 0)), ((0, 1, 0), (0, 0, 1), (1, 0, 0)),
             ((0, -1, 0), (1, 0, 0), (0, 0, 1)), ((0, 0, 1), (1, 0, 0), (0, 1, 0)), ((0, 0, -1), (0, 1, 0), (1, 0, 0))]
    n1 = tess + 1
    t = np.linspace(-1.0, 1.0, n1)
    verts, idx = [], []
    for n, a, b in faces:
        n, a, b = (np.asarray(v, np.float64) for v in (n, a, b))
        base = len(verts) * n1 * n1
        uu, vv = np.meshgrid(t, t, indexing="xy")
        p = n[None, None, :] + uu[..., None] * a + vv[..., None] * b
        if displace > 0.0:
            r = np.linalg.norm(p, axis=-1, keepdims=True)
            d = p / r
            k = phase
            f = (np.sin(k[0] * d[..., 0:1] + k[3]) * np.sin(k[1] * d[..., 1:2] + k[4]) * np.sin(k[2] * d[..., 2:3] + k[5]))
            p = p * (1.0 + displace * f)
        # per-vertex normals from the displaced grid (hard edges between faces, like a cube)
        du = np.gradient(p, axis=1)
        dv = np.gradient(p, axis=0)
        nn = np.cross(du, dv)
        nn /= np.maximum(np.linalg.norm(nn, axis=-1, keepdims=True), 1e-20)
        flip = (nn * n).sum(-1, keepdims=True) < 0
        nn = np.where(flip, -nn, nn)
        tg = du / np.maximum(np.linalg.norm(du, axis=-1, keepdims=True), 1e-20)
        uv = np.stack([(uu + 1) / 2, (vv + 1) / 2], -1)
        v = np.concatenate([p, nn, tg, np.ones_like(uu)[..., None], uv], -1).reshape(-1, 12)
        verts.append(v)
        ii = np.arange(n1 * n1).reshape(n1, n1)
        q = np.stack([ii[:-1, :-1], ii[:-1, 1:], ii[1:, 1:], ii[:-1, :-1], ii[1:, 1:], ii[1:, :-1]], -1).reshape(-1)
        # keep the winding consistent with the outward normal
        tri = q.reshape(-1, 3)
        if np.dot(np.cross(a, b), n) < 0:
            tri = tri[:, ::-1]
        idx.append(tri.reshape(-1) + base)
    return np.concatenate(verts).astype(np.float32), np.concatenate(idx).astype(np.uint32)


class ProjectWriter:
    """Builds the JSON/blob pair of a Luz project (AssetManager::SaveProject layout)."""

    def __init__(self, seed):
        self.uuid = _Uuid(seed)
        self.blob = bytearray()
        self.assets = []
        self.nodes = []
        self.camera_uuid = 0

    def _push(self, arr):
        off = len(self.blob)
        b = np.ascontiguousarray(arr).tobytes()
        self.blob += b
        return {"offset": off, "size": len(b)}

    def mesh(self, verts, idx, name="Mesh"):
        u = self.uuid()
        self.assets.append({"type": 2, "name": name, "uuid": u, "vertices": self._push(verts), "indices": self._push(idx)})
        return u

    def material(self, color=(1, 1, 1, 1), emission=(0, 0, 0), metallic=0.0, roughness=0.5):
        u = self.uuid()
        self.assets.append({"type": 3, "name": "Material", "uuid": u, "color": [float(c) for c in color],
                            "emission": [float(c) for c in emission], "metallic": float(metallic),
                            "roughness": float(roughness), "colorMap": 0, "aoMap": 0, "emissionMap": 0, "normalMap": 0,
                            "metallicRoughnessMap": 0})
        return u

    def mesh_node(self, mesh, material, pos=(0, 0, 0), rot=(0, 0, 0), scale=(1, 1, 1), name="Cube"):
        self.nodes.append({"type": 6, "name": name, "uuid": self.uuid(), "children": [],
                           "position": [float(x) for x in pos], "rotation": [float(x) for x in rot],
                           "scale": [float(x) for x in scale], "mesh": mesh, "material": material})

    def light(self, kind, pos, rot=(0, 0, 0), color=(1, 1, 1), intensity=10.0, radius=0.5, inner=60.0, outer=50.0):
        self.nodes.append({"type": 7, "name": "Light", "uuid": self.uuid(), "children": [],
                           "position": [float(x) for x in pos], "rotation": [float(x) for x in rot], "scale": [1.0, 1.0, 1.0],
                           "color": [float(c) for c in color], "intensity": float(intensity), "lightType": int(kind),
                           "innerAngle": float(inner), "outerAngle": float(outer), "radius": float(radius),
                           "shadowMapRange": 3.0, "shadowMapFar": 2000.0, "volumetricType": 0})

    def camera(self, center, rotation, zoom, fov=60.0, near=0.01, far=1000.0):
        u = self.uuid()
        rads = np.radians(np.asarray(rotation, np.float32) + np.array([90.0, 90.0, 0.0], np.float32)).astype(np.float32)
        d = np.array([np.cos(-rads[1]) * np.sin(rads[0]), np.cos(rads[0]), np.sin(-rads[1]) * np.sin(rads[0])], np.float32)
        eye = np.asarray(center, np.float32) - d * np.float32(zoom)
        self.nodes.append({"type": 8, "name": "Camera", "uuid": u, "children": [], "position": [0.0, 0.0, 0.0],
                           "rotation": [float(x) for x in rotation], "scale": [1.0, 1.0, 1.0], "cameraType": 0, "mode": 0,
                           "eye": [float(x) for x in eye], "center": [float(x) for x in center], "zoom": float(zoom),
                           "farDistance": float(far), "nearDistance": float(near), "horizontalFov": float(fov),
                           "orthoFarDistance": 10.0, "orthoNearDistance": -100.0})
        self.camera_uuid = u

    def write(self, path, bin_path, light_samples, ao_samples, ambient=0.03):
        su = self.uuid()
        scene = {"type": 4, "name": "Scene", "uuid": su, "nodes": self.nodes, "ambientLight": float(ambient),
                 "ambientLightColor": [1.0, 1.0, 1.0], "lightSamples": int(light_samples), "aoSamples": int(ao_samples),
                 "aoMin": 9.999999747378752e-05, "aoMax": 1.0, "exposure": 2.0, "shadowType": 1, "taaEnabled": True,
                 "taaReconstruct": True, "mainCamera": self.camera_uuid}
        j = {"assets": self.assets, "scenes": {str(su): scene}, "initialScene": su}
        with open(path, "w") as f:
            json.dump(j, f)
        with open(bin_path, "wb") as f:
            f.write(bytes(self.blob))


def _four_lights(w, extent, height, seed=11):
    """2 point, 1 spot (inner 60 / outer 50 as the reference's defaults), 1 directional (non-vertical)."""
    rng = np.random.default_rng(seed)
    e = extent
    for kind in (0, 0, 1, 2):
        pos = (rng.uniform(-0.4 * e, 0.4 * e), height * rng.uniform(0.8, 1.2), rng.uniform(-0.4 * e, 0.4 * e))
        col = rng.uniform(0.6, 1.0, 3)
        if kind == 2:
            w.light(2, (pos[0], 2.0 * height, pos[2]), rot=(25.0, 0.0, 18.0), color=col, intensity=1.2, radius=0.05)
        elif kind == 1:
            w.light(1, pos, rot=(10.0, 0.0, -8.0), color=col, intensity=0.6 * height * height, radius=0.3)
        else:
            w.light(0, pos, color=col, intensity=0.35 * height * height, radius=0.4)


def _grid_scene(w, n=64, pitch=3.0, seed=3):
    """C2/C4 geometry: n x n unit cubes on an XZ grid at y=1 over one slab instance."""
    rng = np.random.default_rng(seed)
    cube = w.mesh(*_cube_mesh(1), name="Cube")
    mats = [w.material(color=(*rng.uniform(0.3, 1.0, 3), 1.0), metallic=float(rng.uniform(0, 1)),
                       roughness=float(rng.uniform(0.1, 1.0))) for _ in range(8)]
    half = (n - 1) * pitch / 2
    w.mesh_node(cube, w.material(color=(0.8, 0.8, 0.8, 1), roughness=0.6), pos=(0, -0.05, 0),
                scale=(half + 3.0, 0.05, half + 3.0), name="Slab")
    for gz in range(n):
        for gx in range(n):
            w.mesh_node(cube, mats[(gx * 7 + gz * 3) % 8], pos=(gx * pitch - half, 1.0, gz * pitch - half),
                        rot=(0.0, float(rng.uniform(0, 90)), 0.0), scale=(0.8, 1.0, 0.8))
    return half


def _lattice_scene(w, n_inst=10288, lattice=22, pitch=4.0, unique=False, seed=42):
    """C3/C5 geometry: tessellated displaced cubes (972 triangles) on a jittered lattice."""
    rng = np.random.default_rng(seed)
    vrng = np.random.default_rng(1234)
    n_var = n_inst if unique else 16
    variants = []
    for k in range(n_var):
        phase = np.concatenate([vrng.uniform(2.0, 5.0, 3), vrng.uniform(0, 2 * np.pi, 3)])
        variants.append(w.mesh(*_cube_mesh(9, 0.12, phase), name="Blob%d" % k))
    mats = [w.material(color=(*rng.uniform(0.25, 1.0, 3), 1.0), metallic=float(rng.uniform(0, 1)),
                       roughness=float(rng.uniform(0.1, 1.0)),
                       emission=(0.3, 0.2, 0.1) if i % 7 == 6 else (0, 0, 0)) for i in range(12)]
    cells = rng.permutation(lattice ** 3)[:n_inst]
    cells.sort()
    half = (lattice - 1) * pitch / 2
    for i, c in enumerate(cells):
        cx, cy, cz = c % lattice, (c // lattice) % lattice, c // (lattice * lattice)
        pos = np.array([cx, cy, cz], np.float64) * pitch - half + rng.uniform(-0.8, 0.8, 3)
        s = float(rng.uniform(0.5, 1.5))
        w.mesh_node(variants[i % n_var], mats[i % 12], pos=pos, rot=rng.uniform(0, 360, 3), scale=(s, s, s), name="Blob")
    return half


def write_project(config, out_dir, variant=None):
    """Writes <out_dir>/<config>.luz/.luzbin and returns (luz_path, luzbin_path, settings dict)."""
    os.makedirs(out_dir, exist_ok=True)
    cfg = dict(CONFIGS[config])
    path, bin_path = os.path.join(out_dir, config + ".luz"), os.path.join(out_dir, config + ".luzbin")
    if config == "c1":
        shutil.copyfile(os.path.join(GOLDEN, "default.luz"), path)
        with gzip.open(os.path.join(GOLDEN, "default.luzbin.gz"), "rb") as f, open(bin_path, "wb") as o:
            o.write(f.read())
        return path, bin_path, cfg
    w = ProjectWriter(seed={"c2": 2, "c3": 3, "c4": 4, "c5": 5}[config])
    if config in ("c2", "c4"):
        half = _grid_scene(w)
        if config == "c2":
            _four_lights(w, 2 * half, 14.0)
        else:  # 256 lights on a 16 x 16 grid at y = 6: 192 point + 64 spot, intensity 5, radius 0.2 (seed 7)
            rng = np.random.default_rng(7)
            kinds = np.array([0] * 192 + [1] * 64)
            rng.shuffle(kinds)
            for i in range(256):
                gx, gz = i % 16, i // 16
                pos = ((gx - 7.5) / 7.5 * half, 6.0, (gz - 7.5) / 7.5 * half)
                w.light(int(kinds[i]), pos, rot=(float(rng.uniform(-15, 15)), 0.0, float(rng.uniform(-15, 15))),
                        color=rng.uniform(0.5, 1.0, 3), intensity=5.0, radius=0.2)
        w.camera(center=(0.0, 1.0, 0.0), rotation=(28.0, 35.0, 0.0), zoom=1.35 * half + 10.0)
    else:
        half = _lattice_scene(w, unique=(variant == "unique"))
        if config == "c3":
            _four_lights(w, 2 * half, 1.25 * half)
        else:
            w.light(0, (0.3 * half, 1.3 * half, -0.2 * half), intensity=0.5 * half * half, radius=0.5)
        w.camera(center=(0.0, 0.0, 0.0), rotation=(24.0, 38.0, 0.0), zoom=2.6 * half)
    w.write(path, bin_path, cfg["light_samples"], cfg["ao_samples"])
    return path, bin_path, cfg


def synthetic_blue_noise(size=1024, seed=1):
    """Stand-in for assets/blue_noise.png where the reference's asset is not available (the shader only
    needs two decorrelated 8-bit channels per texel; the identical texture goes to oracle and CUDA)."""
    return np.random.default_rng(seed).integers(0, 256, (size, size, 4), dtype=np.uint8)


def animate(config, frame, n_nodes, base_pos=None, base_rot=None):
    """Per-frame instance motion of the animated configs.  Returns (pos or None, rot or None) arrays for
    mesh nodes 1.. (C2: yaw += 0.5 deg per frame -> refit) or 0.. (C5: sinusoidal translation -> rebuild)."""
    if config == "c2":
        rot = base_rot.copy()
        rot[:, 1] += 0.5 * frame
        return None, rot
    if config == "c5":
        rng = np.random.default_rng(5)
        ph = rng.uniform(0, 2 * np.pi, (n_nodes, 3))
        pos = base_pos + 0.6 * np.sin(0.15 * frame + ph)
        return pos.astype(np.float32), None
    return None, None

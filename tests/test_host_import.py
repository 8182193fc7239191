"""Scene import: luz_b200/host/import.cpp against the REFERENCE'S OWN importers.

tests/golden/import/*.json.gz is what source/Resources/AssetIO.cpp (tiny_gltf / tiny_obj_loader / stb_image, compiled
in place as oracle/_ref/ref_import) produces for the scene files next to them: node tree, transforms, materials,
textures (size + hash of the RGBA8 bytes), meshes with every vertex float as its bit pattern.  The host mirror must
reproduce all of it bit for bit -- vertex order and de-duplication, OBJ v-flip, quad / polygon triangulation, per-face
material splitting, generated tangents, quaternion -> Euler degrees, matrix decomposition, PNG decoding."""
import ctypes as C
import gzip
import json
import os

import numpy as np
import pytest

from luz_b200 import host

HERE = os.path.dirname(os.path.abspath(__file__))
ASSETS = os.path.join(HERE, "golden", "import")
FILES = ["cube.glb", "point.obj", "directional.obj", "multi.gltf", "embedded.gltf", "shapes.obj"]


def asset_path(name, tmp_path):
    """Path of a scene file; the reference's own assets are stored gzip'ed and unpacked next to nothing else."""
    p = os.path.join(ASSETS, name)
    if os.path.exists(p):
        return p
    out = tmp_path / name
    with gzip.open(p + ".gz", "rb") as f:
        out.write_bytes(f.read())
    return str(out)


def import_dump(path, tmp_path):
    lib = host.load_library()
    err = C.create_string_buffer(512)
    out = str(tmp_path / "dump.json")
    rc = lib.luzhost_import_dump(path.encode(), out.encode(), err, 512)
    assert rc == 0, err.value.decode()
    with open(out) as f:
        return json.load(f)


def assert_same(got, ref, where=""):
    assert type(got) is type(ref), where
    if isinstance(ref, dict):
        assert sorted(got) == sorted(ref), where
        for k in ref:
            assert_same(got[k], ref[k], where + "/" + k)
    elif isinstance(ref, list):
        assert len(got) == len(ref), "%s: %d != %d" % (where, len(got), len(ref))
        if ref and isinstance(ref[0], int):
            g, r = np.array(got, dtype=np.uint64), np.array(ref, dtype=np.uint64)
            bad = np.nonzero(g != r)[0]
            assert len(bad) == 0, "%s: %d values differ, first at %d (%d != %d)" % (where, len(bad), bad[0], g[bad[0]], r[bad[0]])
        else:
            for i, (a, b) in enumerate(zip(got, ref)):
                assert_same(a, b, "%s[%d]" % (where, i))
    else:
        assert got == ref, "%s: %r != %r" % (where, got, ref)


@pytest.mark.parametrize("name", FILES)
def test_import_matches_reference_importer(name, tmp_path):
    with gzip.open(os.path.join(ASSETS, name + ".json.gz"), "rt") as f:
        ref = json.load(f)
    got = import_dump(asset_path(name, tmp_path), tmp_path)
    assert_same(got, ref)


def test_import_golden_covers_the_intended_branches():
    """The hand-written assets do reach what they were written for (guards against a silently trivial golden)."""
    def load(n):
        with gzip.open(os.path.join(ASSETS, n + ".json.gz"), "rt") as f:
            return json.load(f)
    multi = load("multi.gltf")
    names = [m["name"] for m in multi["meshes"]]
    assert "Quad_0" in names
    assert sorted(t["name"] for t in multi["textures"]) == ["", "CheckerFromView", "Deep", "Interlaced"]  # 16-bit + Adam7 PNGs too
    types = []

    def walk(n):
        types.append(n["type"])
        for c in n["children"]:
            walk(c)
    for n in multi["nodes"]:
        walk(n)
    assert 7 in types and 6 in types  # a LightNode and MeshNodes
    shapes = load("shapes.obj")
    assert len(shapes["materials"]) == 2 and len(shapes["textures"]) == 1
    assert sum(len(m["indices"]) for m in shapes["meshes"]) >= 3 * 16  # quads and polygons were triangulated
    cube = load("cube.glb")
    v = np.array(cube["meshes"][0]["vertices"], dtype=np.uint32).view(np.float32).reshape(-1, 12)
    assert np.allclose(np.linalg.norm(v[:, 6:9], axis=1), 1.0, atol=1e-6)  # generated tangents are unit vectors


def test_import_errors_are_reported_not_fatal(tmp_path):
    lib = host.load_library()
    err = C.create_string_buffer(512)
    out = str(tmp_path / "o.json").encode()
    bad = tmp_path / "bad.glb"
    bad.write_bytes(b"glTF\x02\x00\x00\x00\x10\x00\x00\x00garbage!")
    assert lib.luzhost_import_dump(str(bad).encode(), out, err, 512) != 0 and err.value
    trunc = tmp_path / "t.gltf"
    trunc.write_text('{"asset":{"version":"2.0"},"scenes":[{"nodes":[0]}],"nodes":[{"mesh":0}],"meshes":[{"primitives":'
                     '[{"attributes":{"POSITION":0},"indices":0}]}],"accessors":[{"bufferView":0,"componentType":5126,'
                     '"count":1000,"type":"VEC3"}],"bufferViews":[{"buffer":0,"byteLength":12}],'
                     '"buffers":[{"byteLength":12,"uri":"data:application/octet-stream;base64,AAAAAAAAAAAAAAAA"}]}')
    assert lib.luzhost_import_dump(str(trunc).encode(), out, err, 512) != 0
    assert b"past the end" in err.value
    obj = tmp_path / "z.obj"
    obj.write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 0 1 2\n")
    assert lib.luzhost_import_dump(str(obj).encode(), out, err, 512) != 0
    assert lib.luzhost_import_dump(str(tmp_path / "missing.obj").encode(), out, err, 512) != 0


def test_imported_scene_flattens_like_a_loaded_project(tmp_path):
    """An imported .glb goes through GPUScene::AddAssets / UpdateResources like a .luz project (CPU-only host): one
    instance of the 24-vertex cube with the material's checker texture bound, and it survives SaveProject ->
    LoadProject."""
    app = host.LuzHost(None)
    app.import_file(asset_path("cube.glb", tmp_path), as_scene=True)
    app.set_extent(320, 180, create_images=False)
    app.add_assets()
    app.update_resources()
    assert app.mesh_node_count() == 1
    (verts, idx), = app.meshes()
    assert verts.shape == (24, 12) and idx.shape == (36,)
    models, n = app.models()
    assert n == 1 and models[0].color_map >= 0 and abs(models[0].roughness - 0.4) < 1e-6
    tex, = app.textures()
    assert tex.shape == (1080, 1080, 4)
    app.save_project(str(tmp_path / "cube.luz"), str(tmp_path / "cube.luzbin"))
    app.close()
    again = host.LuzHost(None)
    again.load_project(str(tmp_path / "cube.luz"), str(tmp_path / "cube.luzbin"))
    (v2, i2), = again.meshes()
    assert np.array_equal(v2.view(np.uint32), verts.view(np.uint32)) and np.array_equal(i2, idx)
    again.close()

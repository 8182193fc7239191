#!/usr/bin/env python3
"""Regenerates tests/golden/import/: scene files for the importers (luz_b200/host/import.cpp) and what the REFERENCE'S
OWN importers make of them (oracle/_ref/ref_import = source/Resources/AssetIO.cpp with tiny_gltf / tiny_obj_loader /
stb_image, compiled in place).  Run in the build container only (needs /root/reference); the committed files are what
tests/test_host_import.py uses.

 - cube.glb.gz, point.obj.gz, directional.obj.gz : the reference's own data assets (assets/), gzip'ed
 - multi.gltf + multi.bin, embedded.gltf, shapes.obj + shapes.mtl + checker8.png / deep16.png / adam7.png : written by this script to reach
   the branches the reference's assets do not (interleaved views, u8 / u32 indices, supplied tangents, missing
   normals / uvs, node TRS / matrix / light extension, two scenes, data URIs, quads, polygons, negative indices,
   per-face materials, .mtl fields, textures incl. 16-bit and interlaced PNG)
 - *.json.gz : ref_import's dump of each (floats as bit patterns).  The reference accumulates tangents into
   `new glm::vec3[...]` without initialising it (AssetIO.cpp:314-315), so its output depends on stale heap contents;
   the goldens are produced with MALLOC_PERTURB_=255 and GLIBC_TUNABLES=glibc.malloc.tcache_count=0 (glibc then hands
   out zero-filled blocks, also for the small ones its thread cache would otherwise return untouched), i.e. the
   reference's own code under the initial state it assumes.  Without them its tangents differ from run to run.
"""
import base64, gzip, json, os, shutil, struct, subprocess, sys, tempfile, zlib

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "import")
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("LUZ_REFERENCE", "/root/reference")
REFERENCE_ASSETS = ("cube.glb", "point.obj", "directional.obj")


def png_rgb(w, h, pixel):
    raw = b"".join(b"\x00" + b"".join(bytes(pixel(x, y)) for x in range(w)) for y in range(h))
    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xFFFFFFFF)
    return b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(raw, 9)) + chunk(b"IEND", b"")


def png_raw(w, h, depth, ctype, rows, interlace=0):
    raw = b"".join(b"\x00" + r for r in rows)
    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xFFFFFFFF)
    return b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, interlace)) + \
        chunk(b"IDAT", zlib.compress(raw, 9)) + chunk(b"IEND", b"")


def png_rgba16(w, h):
    """16 bits per sample: stb's 8-bit API keeps the high byte."""
    rows = [b"".join(struct.pack(">4H", (x * 4099 + y * 257) & 0xFFFF, (x * 1543 + y * 7919) & 0xFFFF, (x * y * 911) & 0xFFFF,
                                 65535 - ((x + y) * 3001 & 0xFFFF)) for x in range(w)) for y in range(h)]
    return png_raw(w, h, 16, 6, rows)


def png_adam7_rgb(w, h):
    """Adam7-interlaced RGB8: seven reduced images, each with its own scanlines."""
    px = [[((x * 37 + y * 11) & 255, (x * 5 + y * 83) & 255, (x * y * 7 + 13) & 255) for x in range(w)] for y in range(h)]
    xs, ys, dxs, dys = [0, 4, 0, 2, 0, 1, 0], [0, 0, 4, 0, 2, 0, 1], [8, 8, 4, 4, 2, 2, 1], [8, 8, 8, 4, 4, 2, 2]
    rows = []
    for p in range(7):
        for y in range(ys[p], h, dys[p]):
            r = b"".join(bytes(px[y][x]) for x in range(xs[p], w, dxs[p]))
            if r:
                rows.append(r)
    return png_raw(w, h, 8, 2, rows, interlace=1)


def f32(*v):
    return struct.pack("<%df" % len(v), *v)


def write_gltf_assets():
    checker = png_rgb(8, 8, lambda x, y: (255, 200, 40) if (x // 2 + y // 2) % 2 else (20, 60, 220))
    with open(os.path.join(OUT, "checker8.png"), "wb") as f:
        f.write(checker)
    with open(os.path.join(OUT, "deep16.png"), "wb") as f:
        f.write(png_rgba16(5, 3))
    with open(os.path.join(OUT, "adam7.png"), "wb") as f:
        f.write(png_adam7_rgb(19, 13))
    # ---- multi.bin ----
    blob = bytearray()
    def add(b, align=4):
        while len(blob) % align:
            blob.append(0)
        off = len(blob)
        blob.extend(b)
        return off
    # primitive A0: interleaved position / normal / uv, stride 32, a quad in the XZ plane
    quad = [(-1, 0, -1, 0, 0), (1, 0, -1, 1, 0), (1, 0.25, 1, 1, 1), (-1, 0, 1, 0.125, 0.875)]
    inter = b"".join(f32(x, y, z) + f32(0, 1, 0) + f32(u, v) for (x, y, z, u, v) in quad)
    o_inter = add(inter)
    o_idx8 = add(bytes([0, 1, 2, 0, 2, 3]))
    # primitive A1: positions only, u32 indices
    tri = [(0, 0, 0), (2, 0, 0), (0, 3, 0), (0.5, 0.5, -1.5)]
    o_pos1 = add(b"".join(f32(*p) for p in tri))
    o_idx32 = add(struct.pack("<6I", 0, 1, 2, 0, 2, 3))
    # mesh B: position, normal, tangent, uv in separate views, u16 indices
    o_posb = add(b"".join(f32(*p) for p in [(0, 0, 0), (1, 0, 0), (0, 0, 1)]))
    o_nrmb = add(b"".join(f32(0, 1, 0) for _ in range(3)))
    o_tanb = add(b"".join(f32(1, 0, 0, -1) for _ in range(3)))
    o_uvb = add(b"".join(f32(*p) for p in [(0, 0), (1, 0), (0, 1)]))
    o_idx16 = add(struct.pack("<3H", 0, 2, 1))
    o_img = add(checker)
    with open(os.path.join(OUT, "multi.bin"), "wb") as f:
        f.write(blob)
    views = [
        {"buffer": 0, "byteOffset": o_inter, "byteLength": len(inter), "byteStride": 32},
        {"buffer": 0, "byteOffset": o_idx8, "byteLength": 6},
        {"buffer": 0, "byteOffset": o_pos1, "byteLength": 48},
        {"buffer": 0, "byteOffset": o_idx32, "byteLength": 24},
        {"buffer": 0, "byteOffset": o_posb, "byteLength": 36},
        {"buffer": 0, "byteOffset": o_nrmb, "byteLength": 36},
        {"buffer": 0, "byteOffset": o_tanb, "byteLength": 48},
        {"buffer": 0, "byteOffset": o_uvb, "byteLength": 24},
        {"buffer": 0, "byteOffset": o_idx16, "byteLength": 6},
        {"buffer": 0, "byteOffset": o_img, "byteLength": len(checker)},
    ]
    acc = [
        {"bufferView": 0, "byteOffset": 0, "componentType": 5126, "count": 4, "type": "VEC3", "min": [-1, 0, -1], "max": [1, 0.25, 1]},
        {"bufferView": 0, "byteOffset": 12, "componentType": 5126, "count": 4, "type": "VEC3"},
        {"bufferView": 0, "byteOffset": 24, "componentType": 5126, "count": 4, "type": "VEC2"},
        {"bufferView": 1, "componentType": 5121, "count": 6, "type": "SCALAR"},
        {"bufferView": 2, "componentType": 5126, "count": 4, "type": "VEC3", "min": [0, 0, -1.5], "max": [2, 3, 0]},
        {"bufferView": 3, "componentType": 5125, "count": 6, "type": "SCALAR"},
        {"bufferView": 4, "componentType": 5126, "count": 3, "type": "VEC3", "min": [0, 0, 0], "max": [1, 0, 1]},
        {"bufferView": 5, "componentType": 5126, "count": 3, "type": "VEC3"},
        {"bufferView": 6, "componentType": 5126, "count": 3, "type": "VEC4"},
        {"bufferView": 7, "componentType": 5126, "count": 3, "type": "VEC2"},
        {"bufferView": 8, "componentType": 5123, "count": 3, "type": "SCALAR"},
    ]
    s2 = 0.7071067811865476
    doc = {
        "asset": {"version": "2.0", "generator": "luz_b200 tests/golden/make_import_golden.py"},
        "scene": 1,
        "scenes": [{"name": "First", "nodes": [0, 5]}, {"name": "Second", "nodes": [5]}],
        "nodes": [
            {"name": "Root", "translation": [1.5, -2, 0.25], "rotation": [0.1, 0.2, 0.3, 0.9273618495495703], "scale": [2, 2, 0.5], "children": [1, 2, 3]},
            {"name": "QuadNode", "mesh": 0, "rotation": [0, s2, 0, s2]},
            {"name": "MatrixNode", "mesh": 1, "matrix": [0, 0, -2, 0, 0, 1.5, 0, 0, 3, 0, 0, 0, 4, 5, 6, 1]},
            {"name": "Lamp", "extensions": {"KHR_lights_punctual": {"light": 0}}, "translation": [0, 4, 0]},
            {"name": "Unused"},
            {"name": "Flat", "matrix": [1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 7, 8, 9, 1], "children": [4]},
        ],
        "meshes": [
            {"name": "Quad", "primitives": [
                {"attributes": {"POSITION": 0, "NORMAL": 1, "TEXCOORD_0": 2}, "indices": 3, "material": 0},
                {"attributes": {"POSITION": 4}, "indices": 5}]},
            {"primitives": [{"attributes": {"POSITION": 6, "NORMAL": 7, "TANGENT": 8, "TEXCOORD_0": 9}, "indices": 10, "material": 1}]},
        ],
        "materials": [
            {"name": "Painted", "pbrMetallicRoughness": {"baseColorFactor": [0.8, 0.1, 0.3, 0.5], "metallicFactor": 0.25,
                                                          "roughnessFactor": 0.65, "baseColorTexture": {"index": 0},
                                                          "metallicRoughnessTexture": {"index": 1}},
             "emissiveFactor": [0.1, 0.2, 0.3], "normalTexture": {"index": 1}, "occlusionTexture": {"index": 2},
             "emissiveTexture": {"index": 3}},
            {"pbrMetallicRoughness": {"metallicFactor": 1, "baseColorTexture": {"index": 2}}, "normalTexture": {"index": 3}},
        ],
        "textures": [{"name": "CheckerFromView", "source": 0}, {"source": 1}, {"name": "Deep", "source": 2}, {"name": "Interlaced", "source": 3}],
        "images": [{"bufferView": 9, "mimeType": "image/png"}, {"uri": "checker8.png"}, {"uri": "deep16.png"}, {"uri": "adam7.png"}],
        "extensions": {"KHR_lights_punctual": {"lights": [{"type": "point", "color": [1, 1, 1], "intensity": 5}]}},
        "extensionsUsed": ["KHR_lights_punctual"],
        "accessors": acc, "bufferViews": views, "buffers": [{"uri": "multi.bin", "byteLength": len(blob)}],
    }
    with open(os.path.join(OUT, "multi.gltf"), "w") as f:
        json.dump(doc, f, indent=1)
    # ---- embedded.gltf: everything in data URIs, a single triangle without material ----
    tri_blob = b"".join(f32(*p) for p in [(0, 0, 0), (1, 0, 0), (0, 1, 0)]) + b"".join(f32(0, 0, 1) for _ in range(3)) + \
        b"".join(f32(*p) for p in [(0, 0), (1, 0), (0, 1)]) + struct.pack("<3H", 0, 1, 2) + b"\x00\x00"
    emb = {
        "asset": {"version": "2.0"},
        "scenes": [{"nodes": [0]}],
        "nodes": [{"mesh": 0, "name": "Tri", "rotation": [0.5, 0.5, 0.5, 0.5]}],
        "meshes": [{"primitives": [{"attributes": {"POSITION": 0, "NORMAL": 1, "TEXCOORD_0": 2}, "indices": 3}]}],
        "accessors": [{"bufferView": 0, "componentType": 5126, "count": 3, "type": "VEC3", "min": [0, 0, 0], "max": [1, 1, 0]},
                      {"bufferView": 1, "componentType": 5126, "count": 3, "type": "VEC3"},
                      {"bufferView": 2, "componentType": 5126, "count": 3, "type": "VEC2"},
                      {"bufferView": 3, "componentType": 5123, "count": 3, "type": "SCALAR"}],
        "bufferViews": [{"buffer": 0, "byteOffset": 0, "byteLength": 36}, {"buffer": 0, "byteOffset": 36, "byteLength": 36},
                        {"buffer": 0, "byteOffset": 72, "byteLength": 24}, {"buffer": 0, "byteOffset": 96, "byteLength": 6}],
        "buffers": [{"byteLength": len(tri_blob), "uri": "data:application/octet-stream;base64," + base64.b64encode(tri_blob).decode()}],
    }
    with open(os.path.join(OUT, "embedded.gltf"), "w") as f:
        json.dump(emb, f)


def write_obj_assets():
    mtl = """# materials for shapes.obj
newmtl red
Kd 0.8 0.1 0.1
Ks 0.5 0.25 0.125
Ke 0 0.5 1
Pm 0.75
Pr 0.3
map_Kd checker8.png

newmtl plain
Kd 0.2 0.4 0.6
Ks 0 0 0
norm checker8.png
"""
    obj = """# written by tests/golden/make_import_golden.py
mtllib shapes.mtl
o Box Part
v -1 -1 -1
v 1 -1 -1
v 1 1 -1
v -1 1 -1
v -1 -1 1
v 1 -1 1
v 1.0e0 +1 1
v -1 1 1.000000
vt 0 0
vt 1 0
vt 1 1
vt 0 1
vn 0 0 -1
vn 0 0 1
vn 1 0 0
usemtl red
f 1/1/1 4/4/1 3/3/1 2/2/1
f 5/1/2 6/2/2 7/3/2 8/4/2
usemtl plain
f 2/1/3 3/2/3 7/3/3 6/4/3
f 1 2 6 5
usemtl red
f -8//1 -4//1 -1//1
o Fan
v 0 0 5
v 2 0 5
v 3 1.5 5
v 2.0 3 5
v .5 3.5 5
v -1 2 5
v 0.5 1 5
g cap extra words
f 9 10 11 12 13 14
f 9/1 10/2 15/3
g
usemtl missing
f 15 12 11 10 9 14 13
v 25e-1 -.5 5
f 16 10 9
"""
    with open(os.path.join(OUT, "shapes.mtl"), "w") as f:
        f.write(mtl)
    with open(os.path.join(OUT, "shapes.obj"), "w", newline="") as f:
        f.write(obj.replace("\n", "\r\n"))


def main():
    if not os.path.isdir(os.path.join(REF, "assets")):
        sys.exit("reference not mounted at %s" % REF)
    os.makedirs(OUT, exist_ok=True)
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "_ref/ref_import"])
    for name in REFERENCE_ASSETS:  # data assets of the reference, stored gzip'ed like tests/golden/default.luzbin.gz
        with open(os.path.join(REF, "assets", name), "rb") as f:
            data = f.read()
        with open(os.path.join(OUT, name + ".gz"), "wb") as f:
            f.write(gzip.compress(data, 9, mtime=0))
    write_gltf_assets()
    write_obj_assets()
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_import")
    env = dict(os.environ, MALLOC_PERTURB_="255", GLIBC_TUNABLES="glibc.malloc.tcache_count=0")
    with tempfile.TemporaryDirectory() as tmp:
        for name in REFERENCE_ASSETS + ("multi.gltf", "embedded.gltf", "shapes.obj"):
            out = os.path.join(tmp, "out.json")
            src = os.path.join(REF, "assets", name) if name in REFERENCE_ASSETS else os.path.join(OUT, name)
            subprocess.check_call([exe, src, out], cwd=tmp, env=env, stdout=subprocess.DEVNULL)
            with open(out, "rb") as f:
                text = f.read()
            json.loads(text)
            with open(os.path.join(OUT, name + ".json.gz"), "wb") as f:
                f.write(gzip.compress(text, 9, mtime=0))
            print(name, len(text), "bytes of JSON")


if __name__ == "__main__":
    main()

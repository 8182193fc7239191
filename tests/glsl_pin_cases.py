"""The inputs of the shader-side pinning cases, shared by tests/test_glsl_pin.py (which runs the reference's own GLSL
where /root/reference is mounted) and tests/golden/make_glsl_golden.py (which stores that GLSL's outputs as fixtures
for machines without the reference).  Everything is seeded."""
import numpy as np

import oracle_api as O
import scene_util as S


def c1_case(n_pixels=12000, frame_slot=3, frame=7, light_samples=2, ao_samples=4):
    """assets/default.luz at 1280x720 (config C1) with the file's own sample counts; a seeded pixel sample plus every
    pixel of three rows that cross the cube's silhouette."""
    w, h = 1280, 720
    sc = S.default_scene(frame=frame_slot, light_samples=light_samples, ao_samples=ao_samples)
    world = O.World(sc["meshes"], sc["instances"])
    gb = O.gbuffer_pass(sc["scene"], world, sc["models"], len(sc["instances"]), sc["textures"], w, h, exhaustive=False)
    rng = np.random.default_rng(1)
    px = np.stack([rng.integers(0, w, n_pixels), rng.integers(0, h, n_pixels)], 1)
    rows = np.concatenate([np.stack([np.arange(w), np.full(w, y)], 1) for y in (250, 360, 470)])
    return dict(sc=sc, world=world, gb=gb, frame=frame, pixels=np.concatenate([px, rows]).astype(np.uint32), w=w, h=h)


def synthetic_case(w=320, h=180):
    """Point + spot + directional lights over instanced cubes (BVH2 tracer on both sides), every pixel."""
    sc = S.synthetic_scene(w, h, grid=4, n_lights=3, light_samples=2, ao_samples=3)
    world = O.World(sc["meshes"], sc["instances"])
    gb = O.gbuffer_pass(sc["scene"], world, sc["models"], len(sc["instances"]), sc["textures"], w, h, exhaustive=False)
    xs, ys = np.meshgrid(np.arange(w), np.arange(h))
    return dict(sc=sc, world=world, gb=gb, frame=77, pixels=np.stack([xs.ravel(), ys.ravel()], 1).astype(np.uint32), w=w, h=h)


def adversarial_gbuffer(w, h):
    """Uniform-random G-buffer (seed 99): roughness 0, non-unit normals, zero normals, depth-1 pixels with N != 0."""
    rng = np.random.default_rng(99)
    gb = O.GBuffer(w, h)
    gb.albedo[:] = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    gb.material[:] = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    gb.material[::3, :, 0] = 0
    gb.emission[:] = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    gb.normal[..., :3] = rng.normal(0, 1, (h, w, 3)).astype(np.float32) * rng.uniform(0.2, 2.0, (h, w, 1)).astype(np.float32)
    gb.normal[::7, ::5] = 0
    gb.depth[:] = rng.uniform(0.9990, 1.0, (h, w)).astype(np.float32)
    gb.depth[::4, ::4] = 1.0
    return gb


def taa_images(case, bn):
    """Two consecutive light images of a case (the second is the history) for taa.comp."""
    sc, gb, world = case["sc"], case["gb"], case["world"]
    _, light, _, _, _ = O.light_pass(sc["scene"], gb, 7, bn, world, exhaustive=False)
    _, hist, _, _, _ = O.light_pass(sc["scene"], gb, 8, bn, world, exhaustive=False)
    return light, hist

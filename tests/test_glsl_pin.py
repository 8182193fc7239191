"""Pins the shader side of the oracle to the reference's OWN GLSL.

oracle/glsl_harness/preprocess.py rewrites /root/reference/source/Shaders/{light.frag, taa.comp, utils.glsl, LuzCommon.h}
lexically for the host (no statement is restated by hand), oracle/Makefile compiles the result against the reference's
glm into oracle/_ref/libglsl_ref.so, and these tests run it next to oracle/luz_oracle.cpp on identical inputs:

  * light.frag: 14 000+ pixels of config C1 (assets/default.luz, the file's 2 shadow + 4 AO samples, exhaustive tracer),
    every pixel of a point + spot + directional scene, and the adversarial G-buffer (roughness 0, non-unit and zero
    normals, depth 1 with N != 0): every ray-query bit identical, NaN pixels identical, radiance equal to a few ulp;
  * taa.comp: every invocation incl. the image borders (REPEAT wrap) and out-of-image invocations, both `reconstruct`
    values, genuine history.

Not bit-for-bit everywhere, and why: glm evaluates mat4 * vec4 as (m0 x + m1 y) + (m2 z + m3 w), normalize() as
v * inversesqrt(dot) and the oracle (like the CUDA kernels) as a left-to-right sum and v / sqrt(dot) -- GLSL leaves both
to the implementation -- so positions differ in the last bit and everything downstream by a few ulp (>= 70 % of the
pixels are bit-equal; the bound asserted is 2e-5 relative to max(|ref|, 1), 1e-3 on the roughness-0 pixels whose GGX
term is a 0/0-type expression).  The fixed-function units the shaders call (texture unit, ray query) are driver
territory: texture() is implemented in the harness from the sampler the reference creates, ray queries go to the
oracle's tracer.

Where /root/reference is absent (the GPU box) the same oracle outputs are checked against fixtures the harness produced
here: tests/golden/glsl_ref_*.npz (tests/golden/make_glsl_golden.py)."""
import os

import numpy as np
import pytest

import glsl_pin_cases as cases
import glsl_ref_api as G
import oracle_api as O
import scene_util as S

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
needs_ref = pytest.mark.skipif(not G.available(), reason="reference GLSL not available (no /root/reference, no oracle/_ref)")
TOL = 2e-5


def rel_err(got, ref):
    return np.abs(got - ref) / np.maximum(np.abs(ref), 1.0)


def oracle_at(case, bn, exhaustive, sw=1, aw=1, gb=None, frame=None):
    sc = case["sc"]
    px = case["pixels"]
    rows = np.unique(px[:, 1]).astype(np.uint32)
    rc, ref, sm, am, st = O.light_pass(sc["scene"], gb if gb is not None else case["gb"], case["frame"] if frame is None else frame,
                                       bn, case["world"], exhaustive=exhaustive, row_list=rows, shadow_words=sw, ao_words=aw)
    assert rc == 0
    return ref[px[:, 1], px[:, 0]], sm[px[:, 1], px[:, 0]], am[px[:, 1], px[:, 0]]


@needs_ref
def test_light_frag_c1_pixels_match_reference_glsl():
    bn = S.blue_noise()
    case = cases.c1_case()
    ref, sm, am = oracle_at(case, bn, exhaustive=True)
    got, gsm, gam = G.light_frag(case["sc"]["scene"], case["gb"], case["frame"], bn, case["world"], case["pixels"], exhaustive=True)
    assert case["pixels"].shape[0] >= 14000
    assert np.array_equal(gsm, sm) and np.array_equal(gam, am)
    assert (np.linalg.norm(case["gb"].normal[case["pixels"][:, 1], case["pixels"][:, 0], :3], axis=1) > 0).mean() > 0.3
    err = rel_err(got, ref)
    assert float(err.max()) <= TOL, float(err.max())
    assert float((err == 0).all(axis=1).mean()) >= 0.7


@needs_ref
def test_light_frag_three_light_types_and_adversarial_gbuffer_match_reference_glsl():
    bn = S.blue_noise()
    case = cases.synthetic_case()
    ref, sm, am = oracle_at(case, bn, exhaustive=False)
    got, gsm, gam = G.light_frag(case["sc"]["scene"], case["gb"], case["frame"], bn, case["world"], case["pixels"], exhaustive=False)
    assert np.array_equal(gsm, sm) and np.array_equal(gam, am) and int(np.unpackbits(sm.view(np.uint8)).sum()) > 1000
    assert float(rel_err(got, ref).max()) <= TOL
    # shadowType 0: every light fully shadowed (light.frag:166-168); ray tracing with 0 samples: unshadowed (:87-89)
    for st, ls in ((0, 1), (1, 0)):
        c2 = dict(case, sc=S.synthetic_scene(case["w"], case["h"], grid=4, n_lights=3, light_samples=ls, ao_samples=0, shadow_type=st))
        ref2, _, _ = oracle_at(c2, bn, exhaustive=False)
        got2, _, _ = G.light_frag(c2["sc"]["scene"], c2["gb"], c2["frame"], bn, c2["world"], c2["pixels"], exhaustive=False)
        assert float(rel_err(got2, ref2).max()) <= TOL
    gb = cases.adversarial_gbuffer(case["w"], case["h"])
    ref, sm, am = oracle_at(case, bn, exhaustive=False, gb=gb, frame=200)
    got, gsm, gam = G.light_frag(case["sc"]["scene"], gb, 200, bn, case["world"], case["pixels"], exhaustive=False)
    assert np.array_equal(gsm, sm) and np.array_equal(gam, am)
    assert np.array_equal(np.isnan(got), np.isnan(ref)) and np.array_equal(np.isinf(got), np.isinf(ref))
    fin = np.isfinite(ref).all(axis=1)
    err = rel_err(got[fin], ref[fin])
    assert float(np.quantile(err, 0.999)) <= TOL and float(err.max()) <= 1e-3, (float(err.max()), float(np.quantile(err, 0.999)))


@needs_ref
def test_taa_comp_matches_reference_glsl_including_borders():
    bn = S.blue_noise()
    case = cases.synthetic_case()
    w, h = case["w"], case["h"]
    light, hist = cases.taa_images(case, bn)
    # every pixel, plus invocations beyond the image (the dispatch is (W/32+1, H/32+1) groups of 32x32, taa.comp:274-276)
    px = np.concatenate([case["pixels"], np.array([[w, 0], [0, h], [w + 5, h + 7]], np.uint32)])
    for reconstruct in (True, False):
        ref = O.taa_pass(case["sc"]["scene"], light, hist, case["gb"].depth, reconstruct)
        got = G.taa_comp(case["sc"]["scene"], light, hist, case["gb"].depth, reconstruct, px[:-3]).reshape(h, w, 4)
        stray = G.taa_comp(case["sc"]["scene"], light, hist, case["gb"].depth, reconstruct, px[-3:], full_image=True)
        assert not stray.any()  # invocations beyond the image store nothing
        err = rel_err(got, ref)
        assert float(err.max()) <= TOL, float(err.max())
        border = np.concatenate([err[0].ravel(), err[-1].ravel(), err[:, 0].ravel(), err[:, -1].ravel()])
        assert float(border.max()) <= TOL
        assert float((err == 0).all(axis=2).mean()) >= 0.7
    # NaN in the neighbourhood: the reconstruction falls back to the centre tap and the result to the source (:287-289, :303-305)
    bad = light.copy()
    bad[h // 2, w // 2] = np.nan
    ref = O.taa_pass(case["sc"]["scene"], bad, hist, case["gb"].depth, True)
    got = G.taa_comp(case["sc"]["scene"], bad, hist, case["gb"].depth, True, case["pixels"]).reshape(h, w, 4)
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    fin = np.isfinite(ref)
    # min() / max() with a NaN operand are undefined in GLSL (glm: (b < a) ? b : a, the oracle: fminf): the AABB of the
    # eight pixels around the NaN may differ in which finite neighbour it keeps; everything else is as tight as above
    err = rel_err(got[fin], ref[fin])
    assert float(err.max()) <= 1e-3 and float((err > TOL).sum()) <= 9 * 4


def test_oracle_matches_committed_reference_glsl_fixtures():
    """Runs everywhere: the oracle against the outputs the reference's GLSL produced here (make_glsl_golden.py)."""
    bn = S.blue_noise()
    f = np.load(os.path.join(GOLDEN, "glsl_ref_c1.npz"))
    case = cases.c1_case(n_pixels=int(f["n_pixels"]))
    assert np.array_equal(case["pixels"], f["pixels"])
    ref, sm, am = oracle_at(case, bn, exhaustive=True)
    assert np.array_equal(sm, f["shadow_mask"]) and np.array_equal(am, f["ao_mask"])
    assert float(rel_err(ref, f["radiance"]).max()) <= TOL
    t = np.load(os.path.join(GOLDEN, "glsl_ref_taa.npz"))
    case = cases.synthetic_case()
    light, hist = cases.taa_images(case, bn)
    for reconstruct, key in ((True, "resolved_reconstruct"), (False, "resolved_plain")):
        ref = O.taa_pass(case["sc"]["scene"], light, hist, case["gb"].depth, reconstruct)
        rows = t["rows"]
        assert float(rel_err(ref[rows], t[key]).max()) <= TOL

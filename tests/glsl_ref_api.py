"""ctypes binding of oracle/_ref/libglsl_ref.so: the reference's own light.frag / taa.comp compiled for the host
(oracle/glsl_harness/).  Exists only where /root/reference is mounted; tests that need it skip otherwise and fall back
to the committed fixtures it produced (tests/golden/glsl_ref_*.npz)."""
import ctypes as C
import os
import subprocess

import numpy as np

import oracle_api as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_ref", "libglsl_ref.so")
REF_SHADERS = "/root/reference/source/Shaders"

_lib = None


def available():
    return os.path.exists(LIB) or os.path.isdir(REF_SHADERS)


def lib():
    global _lib
    if _lib is None:
        O.lib()
        if os.path.isdir(REF_SHADERS):  # (re)build when the reference is here: make decides whether anything is stale
            subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "_ref/libglsl_ref.so"])
        L = C.CDLL(LIB)
        vp, u32, i32 = C.c_void_p, C.c_uint32, C.c_int
        L.glsl_light_frag.restype = i32
        L.glsl_light_frag.argtypes = [vp, u32, u32, vp, u32, vp, u32, u32, vp, i32, vp, u32, vp, vp, u32, vp, u32]
        L.glsl_taa_comp.restype = i32
        L.glsl_taa_comp.argtypes = [vp, u32, u32, vp, vp, vp, i32, vp, u32, vp]
        _lib = L
    return _lib


def light_frag(scene, gb, frame, blue_noise, world, pixels_xy, exhaustive=True, shadow_words=1, ao_words=1):
    """light.frag main() at the listed (x, y) pixels: (radiance [n,4], shadow mask [n,sw], AO mask [n,aw])."""
    px = np.ascontiguousarray(pixels_xy, np.uint32).reshape(-1, 2)
    n = px.shape[0]
    out = np.zeros((n, 4), np.float32)
    sm = np.zeros((n, shadow_words), np.uint32)
    am = np.zeros((n, ao_words), np.uint32)
    bn = np.ascontiguousarray(blue_noise, np.uint8)
    g = gb.c()
    rc = lib().glsl_light_frag(O._p(scene), gb.w, gb.h, O._p(g), frame, O._p(bn), bn.shape[1], bn.shape[0], world.h,
                               int(exhaustive), O._p(px), n, O._p(out), O._p(sm), shadow_words, O._p(am), ao_words)
    assert rc == 0
    return out, sm, am


def taa_comp(scene, light_in, history, depth, reconstruct, pixels_xy, full_image=False):
    """taa.comp main() for the listed invocations; returns the values imageStore wrote at those pixels [n,4] (or the
    whole output image, zero where nothing was stored)."""
    h, w = depth.shape
    px = np.ascontiguousarray(pixels_xy, np.uint32).reshape(-1, 2)
    out = np.zeros((h, w, 4), np.float32)
    li = np.ascontiguousarray(light_in, np.float32)
    hi = np.ascontiguousarray(history, np.float32)
    d = np.ascontiguousarray(depth, np.float32)
    rc = lib().glsl_taa_comp(O._p(scene), w, h, O._p(li), O._p(hi), O._p(d), 1 if reconstruct else 0, O._p(px), px.shape[0],
                             O._p(out))
    assert rc == 0
    return out if full_image else out[px[:, 1], px[:, 0]]

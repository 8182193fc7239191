"""ctypes binding of oracle/libluz_oracle.so -- the CPU checker.  Imported only by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import os
import subprocess

import numpy as np

from luz_b200 import wire

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "libluz_oracle.so")


class OrcMesh(C.Structure):
    _fields_ = [("vertices", C.c_void_p), ("vertex_count", C.c_uint32), ("vertex_stride", C.c_uint32),
                ("indices", C.c_void_p), ("index_count", C.c_uint32)]


class OrcInstance(C.Structure):
    _fields_ = [("mesh", C.c_uint32), ("model_mat", C.c_float * 16), ("custom_index", C.c_uint32)]


class OrcTexture(C.Structure):
    _fields_ = [("rgba8", C.c_void_p), ("width", C.c_uint32), ("height", C.c_uint32)]


class OrcGbuffer(C.Structure):
    _fields_ = [("albedo", C.c_void_p), ("normal", C.c_void_p), ("material", C.c_void_p),
                ("emission", C.c_void_p), ("depth", C.c_void_p)]


class OrcShadowMap(C.Structure):
    _fields_ = [("data", C.c_void_p), ("res", C.c_uint32), ("layers", C.c_uint32)]


class OrcStats(C.Structure):
    _fields_ = [("lit_pixels", C.c_uint64), ("rays", C.c_uint64), ("rays_occluded", C.c_uint64)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "libluz_oracle.so"])
        L = C.CDLL(LIB)
        vp, u32, i32 = C.c_void_p, C.c_uint32, C.c_int
        L.orc_world_create.restype = vp
        L.orc_world_create.argtypes = [vp, u32, vp, u32]
        L.orc_world_destroy.argtypes = [vp]
        L.orc_set_threads.argtypes = [i32]
        L.orc_get_threads.restype = i32
        L.orc_trace_any.argtypes = [vp, u32, vp, vp, vp, vp, i32, vp]
        L.orc_trace_closest.argtypes = [vp, u32, vp, vp, vp, vp, i32, vp, vp, vp]
        L.orc_light_pass.restype = i32
        L.orc_light_pass.argtypes = [vp, vp, u32, u32, u32, vp, u32, vp, u32, u32, vp, i32, u32, u32, vp, vp, u32,
                                     vp, u32, vp]
        L.orc_light_pass_rows.restype = i32
        L.orc_light_pass_rows.argtypes = [vp, vp, u32, u32, u32, vp, u32, vp, u32, u32, vp, i32, vp, u32, u32, u32, vp, vp,
                                          u32, vp, u32, vp]
        L.orc_taa_pass.restype = i32
        L.orc_taa_pass.argtypes = [vp, u32, u32, vp, vp, vp, i32, u32, u32, vp]
        L.orc_volumetric_screen_pass.restype = i32
        L.orc_volumetric_screen_pass.argtypes = [vp, vp, u32, u32, u32, vp, vp, u32, u32, u32, u32, u32, vp]
        L.orc_shadow_map_pass.restype = i32
        L.orc_shadow_map_pass.argtypes = [vp, vp, u32, vp]
        L.orc_bind_shadow_maps.argtypes = [vp, u32]
        L.orc_shadow_factor.restype = C.c_float
        L.orc_shadow_factor.argtypes = [vp, vp, vp, vp]
        L.orc_volumetric_shadow_pass.restype = i32
        L.orc_volumetric_shadow_pass.argtypes = [vp, vp, u32, u32, u32, vp, vp, u32, u32, u32, u32, u32, vp]
        L.orc_compose_pass.restype = i32
        L.orc_compose_pass.argtypes = [u32, u32, vp, vp]
        L.orc_gbuffer_pass.restype = i32
        L.orc_gbuffer_pass.argtypes = [vp, vp, vp, u32, vp, u32, u32, u32, i32, u32, u32, vp]
        L.orc_blue_noise_sample.argtypes = [vp, u32, u32, u32, u32, i32, u32, vp]
        L.orc_depth_to_world.argtypes = [vp, C.c_float, C.c_float, C.c_float, vp]
        L.orc_mitchell.restype = C.c_float
        L.orc_mitchell.argtypes = [C.c_float]
        L.orc_tri_test.restype = i32
        L.orc_tri_test.argtypes = [vp, vp, vp, vp, vp, C.c_float, C.c_float, vp]
        _lib = L
    return _lib


def _p(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"]
        return a.ctypes.data_as(C.c_void_p)
    return C.cast(C.byref(a), C.c_void_p)


class World:
    """meshes: list of (vertices ndarray [n, stride/4] float32, indices uint32); instances: list of
    (mesh index, mat16 column-major, custom_index)."""

    def __init__(self, meshes, instances):
        self._keep = []
        ms = (OrcMesh * max(len(meshes), 1))()
        for i, (v, idx) in enumerate(meshes):
            v = np.ascontiguousarray(v, dtype=np.float32)
            idx = np.ascontiguousarray(idx, dtype=np.uint32)
            self._keep += [v, idx]
            ms[i] = OrcMesh(v.ctypes.data, v.shape[0], v.strides[0], idx.ctypes.data, idx.size)
        ins = (OrcInstance * max(len(instances), 1))()
        for i, (m, mat, ci) in enumerate(instances):
            ins[i].mesh = m
            ins[i].model_mat[:] = [float(x) for x in np.asarray(mat, dtype=np.float32).reshape(16)]
            ins[i].custom_index = ci
        self.h = lib().orc_world_create(_p(ms), len(meshes), _p(ins), len(instances))
        self.n_instances = len(instances)

    def __del__(self):
        try:
            if self.h:
                lib().orc_world_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def trace_any(self, o, d, tmin, tmax, exhaustive=True):
        o = np.ascontiguousarray(o, np.float32)
        d = np.ascontiguousarray(d, np.float32)
        n = o.shape[0]
        tmin = np.ascontiguousarray(np.broadcast_to(np.asarray(tmin, np.float32), (n,)))
        tmax = np.ascontiguousarray(np.broadcast_to(np.asarray(tmax, np.float32), (n,)))
        hit = np.zeros(n, np.uint8)
        lib().orc_trace_any(self.h, n, _p(o), _p(d), _p(tmin), _p(tmax), 1 if exhaustive else 0, _p(hit))
        return hit

    def trace_closest(self, o, d, tmin, tmax, exhaustive=True):
        o = np.ascontiguousarray(o, np.float32)
        d = np.ascontiguousarray(d, np.float32)
        n = o.shape[0]
        tmin = np.ascontiguousarray(np.broadcast_to(np.asarray(tmin, np.float32), (n,)))
        tmax = np.ascontiguousarray(np.broadcast_to(np.asarray(tmax, np.float32), (n,)))
        t = np.zeros(n, np.float32)
        inst = np.zeros(n, np.int32)
        prim = np.zeros(n, np.int32)
        lib().orc_trace_closest(self.h, n, _p(o), _p(d), _p(tmin), _p(tmax), 1 if exhaustive else 0, _p(t), _p(inst),
                                _p(prim))
        return t, inst, prim


class GBuffer:
    def __init__(self, w, h):
        self.w, self.h = w, h
        self.albedo = np.zeros((h, w, 4), np.uint8)
        self.normal = np.zeros((h, w, 4), np.float32)
        self.material = np.zeros((h, w, 4), np.uint8)
        self.emission = np.zeros((h, w, 4), np.uint8)
        self.depth = np.ones((h, w), np.float32)

    def c(self):
        return OrcGbuffer(self.albedo.ctypes.data, self.normal.ctypes.data, self.material.ctypes.data,
                          self.emission.ctypes.data, self.depth.ctypes.data)


def gbuffer_pass(scene, world, models, n_models, textures, w, h, exhaustive=False, rows=None):
    gb = GBuffer(w, h)
    y0, y1 = rows if rows else (0, h)
    tex = (OrcTexture * max(len(textures), 1))()
    keep = []
    for i, t in enumerate(textures):
        t = np.ascontiguousarray(t, np.uint8)
        keep.append(t)
        tex[i] = OrcTexture(t.ctypes.data, t.shape[1], t.shape[0])
    g = gb.c()
    rc = lib().orc_gbuffer_pass(_p(scene), world.h, _p(models), n_models, _p(tex), len(textures), w, h,
                                1 if exhaustive else 0, y0, y1, _p(g))
    assert rc == 0
    return gb


def light_pass(scene, gb, frame, blue_noise, world, extra_lights=None, exhaustive=True, rows=None,
               shadow_words=0, ao_words=0, row_list=None, x_range=None):
    """light.frag over rows [rows[0], rows[1]) (default: all), or over the rows in row_list restricted to the columns
    x_range.  exhaustive: False / 0 BVH2, True / 1 every triangle, 2 every instance box + every triangle inside.
    Returns full-frame arrays of which only the requested pixels are written."""
    w, h = gb.w, gb.h
    if row_list is None:
        y0, y1 = rows if rows else (0, h)
        row_list = np.arange(y0, y1, dtype=np.uint32)
    row_list = np.ascontiguousarray(row_list, np.uint32)
    x0, x1 = x_range if x_range else (0, w)
    out = np.zeros((h, w, 4), np.float32)
    sm = np.zeros((h, w, shadow_words), np.uint32) if shadow_words else None
    am = np.zeros((h, w, ao_words), np.uint32) if ao_words else None
    st = OrcStats()
    g = gb.c()
    bn = np.ascontiguousarray(blue_noise, np.uint8)
    n_extra = len(extra_lights) if extra_lights is not None else 0
    rc = lib().orc_light_pass_rows(_p(scene), _p(extra_lights) if n_extra else None, n_extra, w, h, _p(g), frame, _p(bn),
                                   bn.shape[1], bn.shape[0], world.h, int(exhaustive), _p(row_list), row_list.size, x0, x1,
                                   _p(out), _p(sm), shadow_words, _p(am), ao_words, _p(st))
    return rc, out, sm, am, st


def taa_pass(scene, light_in, history, depth, reconstruct=True, rows=None):
    h, w = depth.shape
    y0, y1 = rows if rows else (0, h)
    out = np.zeros((h, w, 4), np.float32)
    light_in = np.ascontiguousarray(light_in, np.float32)
    history = np.ascontiguousarray(history, np.float32)
    depth = np.ascontiguousarray(depth, np.float32)
    rc = lib().orc_taa_pass(_p(scene), w, h, _p(light_in), _p(history), _p(depth), 1 if reconstruct else 0, y0, y1,
                            _p(out))
    assert rc == 0
    return out


def volumetric_screen_pass(scene, light, depth, blue_noise, frame, extra_lights=None, rows=None):
    """screenSpaceVolumetricLight.comp over rows: returns a copy of `light` with the shafts added."""
    h, w = depth.shape
    y0, y1 = rows if rows else (0, h)
    out = np.array(light, np.float32, copy=True, order="C")
    depth = np.ascontiguousarray(depth, np.float32)
    bn = np.ascontiguousarray(blue_noise, np.uint8)
    n_extra = len(extra_lights) if extra_lights is not None else 0
    rc = lib().orc_volumetric_screen_pass(_p(scene), _p(extra_lights) if n_extra else None, n_extra, w, h, _p(depth),
                                          _p(bn), bn.shape[1], bn.shape[0], frame, y0, y1, _p(out))
    assert rc == 0
    return out


def shadow_map_pass(light, world, res):
    """DeferredRenderer::ShadowMapPass for one light: [layers, res, res] float32."""
    layers = 6 if light.type == wire.LIGHT_POINT else 1
    out = np.zeros((layers, res, res), np.float32)
    assert lib().orc_shadow_map_pass(_p(light), world.h, res, _p(out)) == 0
    return out


def shadow_factor(light, shadow_map, frag_pos, shadow_origin=None):
    m = np.ascontiguousarray(shadow_map, np.float32)
    rec = OrcShadowMap(m.ctypes.data, m.shape[1], m.shape[0])
    fp = np.asarray(frag_pos, np.float32)
    so = np.asarray(shadow_origin if shadow_origin is not None else frag_pos, np.float32)
    return float(lib().orc_shadow_factor(_p(light), _p(rec), _p(fp), _p(so)))


class BoundShadowMaps:
    """with BoundShadowMaps({light index: map array}): orc_light_pass / volumetric_shadow_pass sample these."""

    def __init__(self, maps, n_lights):
        self.arr = (OrcShadowMap * max(n_lights, 1))()
        self.keep = []
        for i, m in maps.items():
            m = np.ascontiguousarray(m, np.float32)
            self.keep.append(m)
            self.arr[i] = OrcShadowMap(m.ctypes.data, m.shape[1], m.shape[0])
        self.n = n_lights

    def __enter__(self):
        lib().orc_bind_shadow_maps(_p(self.arr), self.n)
        return self

    def __exit__(self, *a):
        lib().orc_bind_shadow_maps(None, 0)


def volumetric_shadow_pass(scene, light, depth, blue_noise, frame, extra_lights=None, rows=None):
    """shadowMapVolumetricLight.comp over rows (maps from BoundShadowMaps): returns a copy of `light` + shafts."""
    h, w = depth.shape
    y0, y1 = rows if rows else (0, h)
    out = np.array(light, np.float32, copy=True, order="C")
    depth = np.ascontiguousarray(depth, np.float32)
    bn = np.ascontiguousarray(blue_noise, np.uint8)
    n_extra = len(extra_lights) if extra_lights is not None else 0
    rc = lib().orc_volumetric_shadow_pass(_p(scene), _p(extra_lights) if n_extra else None, n_extra, w, h, _p(depth),
                                          _p(bn), bn.shape[1], bn.shape[0], frame, y0, y1, _p(out))
    assert rc == 0
    return out


def compose_pass(light_in):
    h, w, _ = light_in.shape
    out = np.zeros((h, w, 4), np.uint8)
    light_in = np.ascontiguousarray(light_in, np.float32)
    assert lib().orc_compose_pass(w, h, _p(light_in), _p(out)) == 0
    return out

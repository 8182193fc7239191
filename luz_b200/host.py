"""ctypes binding of libluzhost.so -- the C++ mirror of Luz's host side (scene format, GPUScene,
DeferredRenderer calls).  Harness only; a C++ Luz host includes luz_b200/host/*.hpp directly."""
import ctypes as C
import os

import numpy as np

from . import rt as _rt
from . import wire

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libluzhost.so")

FRAME_OPAQUE, FRAME_COMPOSE, FRAME_TLAS_REFIT, FRAME_NO_UPDATE = 1, 2, 4, 8

_lib = None


def load_library():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(LIB_PATH + " is not built (python -m luz_b200.build)")
    _rt.load_library()  # libluzhost.so links against libluzrt.so
    L = C.CDLL(LIB_PATH)
    vp, u32, i32, fp = C.c_void_p, C.c_uint32, C.c_int, C.POINTER(C.c_float)
    sig = {
        "luzhost_create": (vp, [vp]),
        "luzhost_destroy": (None, [vp]),
        "luzhost_last_error": (C.c_char_p, [vp]),
        "luzhost_load_project": (i32, [vp, C.c_char_p, C.c_char_p]),
        "luzhost_save_project": (i32, [vp, C.c_char_p, C.c_char_p]),
        "luzhost_import": (i32, [vp, C.c_char_p, i32]),
        "luzhost_import_dump": (i32, [C.c_char_p, C.c_char_p, C.c_char_p, u32]),
        "luzhost_set_extent": (i32, [vp, u32, u32, i32]),
        "luzhost_add_assets": (i32, [vp]),
        "luzhost_update_resources": (i32, [vp]),
        "luzhost_update_resources_gpu": (i32, [vp, i32]),
        "luzhost_render_frame": (i32, [vp, u32]),
        "luzhost_frame_count": (i32, [vp]),
        "luzhost_set_frame_count": (None, [vp, i32]),
        "luzhost_scene_block": (vp, [vp]),
        "luzhost_models": (vp, [vp, C.POINTER(u32)]),
        "luzhost_extra_lights": (vp, [vp, C.POINTER(u32)]),
        "luzhost_instances": (vp, [vp, C.POINTER(u32)]),
        "luzhost_instance_mesh": (i32, [vp, u32]),
        "luzhost_mesh_count": (u32, [vp]),
        "luzhost_mesh": (i32, [vp, u32, C.POINTER(vp), C.POINTER(u32), C.POINTER(vp), C.POINTER(u32), C.POINTER(C.c_uint64)]),
        "luzhost_texture_count": (u32, [vp]),
        "luzhost_texture": (i32, [vp, u32, C.POINTER(vp), C.POINTER(u32), C.POINTER(u32)]),
        "luzhost_mesh_node_count": (u32, [vp]),
        "luzhost_light_count": (u32, [vp]),
        "luzhost_mesh_node_set_transform": (i32, [vp, u32, fp, fp, fp]),
        "luzhost_mesh_node_get_transform": (i32, [vp, u32, fp, fp, fp]),
        "luzhost_mesh_nodes_set_transforms": (i32, [vp, u32, u32, vp, vp, vp]),
        "luzhost_scene_settings": (i32, [vp, i32, i32, i32, i32, i32]),
        "luzhost_scene_get_settings": (i32, [vp, C.POINTER(i32), fp]),
        "luzhost_camera_set_orbit": (i32, [vp, fp, fp, C.c_float]),
        "luzhost_camera_use_jitter": (i32, [vp, i32]),
        "luzhost_halton": (C.c_float, [u32, u32]),
        "luzhost_compose_transform": (None, [fp, fp, fp, fp, fp]),
        "luzhost_mat4_inverse": (None, [fp, fp]),
        "luzhost_perspective": (None, [C.c_float, C.c_float, C.c_float, C.c_float, fp]),
        "luzhost_ortho": (None, [C.c_float] * 6 + [fp]),
        "luzhost_look_at": (None, [fp, fp, fp, fp]),
        "luzhost_mat4_mul": (None, [fp, fp, fp]),
        "luzhost_camera_proj": (i32, [vp, C.c_float, C.c_float, fp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def _f(a):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a.ctypes.data_as(C.POINTER(C.c_float)), a


class HostError(RuntimeError):
    pass


class LuzHost:
    """The host application state: AssetManager + scene + camera + GPUScene + DeferredRenderer.
    `rt` is a luz_b200.rt.LuzRT or None (CPU-only use: loading, transforms, UpdateResources)."""

    def __init__(self, rt=None):
        self.lib = load_library()
        self.rt = rt
        self.h = self.lib.luzhost_create(rt.h if rt is not None else None)

    def close(self):
        if self.h:
            self.lib.luzhost_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise HostError("luzhost error %d: %s" % (rc, self.lib.luzhost_last_error(self.h).decode()))

    def load_project(self, path, bin_path):
        self._ck(self.lib.luzhost_load_project(self.h, path.encode(), bin_path.encode()))

    def import_file(self, path, as_scene=True):
        """AssetIO::Import of a .glb / .gltf / .obj scene or a .png texture (luz_b200/host/import.cpp)."""
        self._ck(self.lib.luzhost_import(self.h, path.encode(), 1 if as_scene else 0))

    def save_project(self, path, bin_path):
        self._ck(self.lib.luzhost_save_project(self.h, path.encode(), bin_path.encode()))

    def set_extent(self, w, h, create_images=True):
        self._ck(self.lib.luzhost_set_extent(self.h, w, h, 1 if create_images else 0))
        if create_images and self.rt is not None:
            self.rt.width, self.rt.height = w, h

    def add_assets(self):
        self._ck(self.lib.luzhost_add_assets(self.h))

    def update_resources(self):
        self._ck(self.lib.luzhost_update_resources(self.h))

    def update_resources_gpu(self, tlas_mode=0):
        self._ck(self.lib.luzhost_update_resources_gpu(self.h, tlas_mode))

    def render_frame(self, flags=0):
        self._ck(self.lib.luzhost_render_frame(self.h, flags))

    @property
    def frame_count(self):
        return self.lib.luzhost_frame_count(self.h)

    @frame_count.setter
    def frame_count(self, v):
        self.lib.luzhost_set_frame_count(self.h, v)

    def scene_block(self):
        """A copy of the SceneBlock UpdateResources produced."""
        p = self.lib.luzhost_scene_block(self.h)
        sb = wire.SceneBlock()
        C.memmove(C.byref(sb), p, C.sizeof(sb))
        return sb

    def models(self):
        n = C.c_uint32()
        p = self.lib.luzhost_models(self.h, C.byref(n))
        arr = (wire.ModelBlock * max(n.value, 1))()
        if n.value:
            C.memmove(arr, p, n.value * C.sizeof(wire.ModelBlock))
        return arr, n.value

    def extra_lights(self):
        n = C.c_uint32()
        p = self.lib.luzhost_extra_lights(self.h, C.byref(n))
        if not n.value:
            return None
        arr = (wire.LightBlock * n.value)()
        C.memmove(arr, p, n.value * C.sizeof(wire.LightBlock))
        return arr

    def instances(self):
        """(mesh index, mat16, custom_index) per instance, in TLAS input order."""
        n = C.c_uint32()
        p = self.lib.luzhost_instances(self.h, C.byref(n))
        out = []
        if n.value:
            a = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), (n.value * C.sizeof(wire.Instance),))
            rec = a.view(np.dtype([("blas", "<u4"), ("m", "<f4", 16), ("ci", "<u4")]))
            for i in range(n.value):
                out.append((self.lib.luzhost_instance_mesh(self.h, i), rec["m"][i].copy(), int(rec["ci"][i])))
        return out

    def meshes(self):
        out = []
        for i in range(self.lib.luzhost_mesh_count(self.h)):
            v, nv, ix, ni, uu = C.c_void_p(), C.c_uint32(), C.c_void_p(), C.c_uint32(), C.c_uint64()
            self._ck(self.lib.luzhost_mesh(self.h, i, C.byref(v), C.byref(nv), C.byref(ix), C.byref(ni), C.byref(uu)))
            verts = np.ctypeslib.as_array(C.cast(v, C.POINTER(C.c_float)), (nv.value, 12)).copy() if nv.value else np.zeros((0, 12), np.float32)
            idx = np.ctypeslib.as_array(C.cast(ix, C.POINTER(C.c_uint32)), (ni.value,)).copy() if ni.value else np.zeros(0, np.uint32)
            out.append((verts, idx))
        return out

    def textures(self):
        out = []
        for i in range(self.lib.luzhost_texture_count(self.h)):
            d, w, h = C.c_void_p(), C.c_uint32(), C.c_uint32()
            self._ck(self.lib.luzhost_texture(self.h, i, C.byref(d), C.byref(w), C.byref(h)))
            out.append(np.ctypeslib.as_array(C.cast(d, C.POINTER(C.c_uint8)), (h.value, w.value, 4)).copy())
        return out

    def mesh_node_count(self):
        return self.lib.luzhost_mesh_node_count(self.h)

    def light_count(self):
        return self.lib.luzhost_light_count(self.h)

    def set_mesh_node_transform(self, i, pos=None, rot=None, scale=None):
        keep = [_f(pos), _f(rot), _f(scale)]
        self._ck(self.lib.luzhost_mesh_node_set_transform(self.h, i, *[k[0] if k else None for k in keep]))

    def get_mesh_node_transform(self, i):
        p, r, s = (np.zeros(3, np.float32) for _ in range(3))
        self._ck(self.lib.luzhost_mesh_node_get_transform(self.h, i, _f(p)[0], _f(r)[0], _f(s)[0]))
        return p, r, s

    def set_mesh_node_transforms(self, first, pos=None, rot=None, scale=None):
        arrs = [np.ascontiguousarray(a, np.float32) if a is not None else None for a in (pos, rot, scale)]
        n = next(a.shape[0] for a in arrs if a is not None)
        self._ck(self.lib.luzhost_mesh_nodes_set_transforms(
            self.h, first, n, *[a.ctypes.data_as(C.c_void_p) if a is not None else None for a in arrs]))

    def scene_settings(self, light_samples=-1, ao_samples=-1, shadow_type=-1, taa_enabled=-1, taa_reconstruct=-1):
        self._ck(self.lib.luzhost_scene_settings(self.h, light_samples, ao_samples, shadow_type, int(taa_enabled),
                                                 int(taa_reconstruct)))

    def get_scene_settings(self):
        i5 = (C.c_int * 5)()
        f4 = (C.c_float * 4)()
        self._ck(self.lib.luzhost_scene_get_settings(self.h, i5, f4))
        return dict(lightSamples=i5[0], aoSamples=i5[1], shadowType=i5[2], taaEnabled=bool(i5[3]),
                    taaReconstruct=bool(i5[4]), aoMin=f4[0], aoMax=f4[1], exposure=f4[2], ambientLight=f4[3])

    def camera_set_orbit(self, center=None, rotation=None, zoom=-1.0):
        keep = [_f(center), _f(rotation)]
        self._ck(self.lib.luzhost_camera_set_orbit(self.h, *[k[0] if k else None for k in keep], zoom))

    def camera_proj(self, near, far):
        out, p = _out16()
        self._ck(self.lib.luzhost_camera_proj(self.h, near, far, p))
        return out

    def camera_use_jitter(self, on):
        self._ck(self.lib.luzhost_camera_use_jitter(self.h, 1 if on else 0))


def halton(i, b):
    return float(load_library().luzhost_halton(i, b))


def compose_transform(pos, rot, scale, parent=None):
    out = np.zeros(16, np.float32)
    p, r, s = _f(pos), _f(rot), _f(scale)
    par = _f(parent) if parent is not None else None
    load_library().luzhost_compose_transform(p[0], r[0], s[0], par[0] if par else None, _f(out)[0])
    return out


def _out16():
    out = np.zeros(16, np.float32)
    return out, out.ctypes.data_as(C.POINTER(C.c_float))


def perspective(fovy, aspect, near, far):
    out, p = _out16()
    load_library().luzhost_perspective(fovy, aspect, near, far, p)
    return out


def ortho(left, right, bottom, top, near, far):
    out, p = _out16()
    load_library().luzhost_ortho(left, right, bottom, top, near, far, p)
    return out


def look_at(eye, center, up):
    out, p = _out16()
    e, c, u = _f(eye), _f(center), _f(up)
    load_library().luzhost_look_at(e[0], c[0], u[0], p)
    return out


def mat4_mul(a, b):
    out, p = _out16()
    aa, bb = _f(a), _f(b)
    load_library().luzhost_mat4_mul(aa[0], bb[0], p)
    return out


def mat4_inverse(m):
    out = np.zeros(16, np.float32)
    load_library().luzhost_mat4_inverse(_f(m)[0], out.ctypes.data_as(C.POINTER(C.c_float)))
    return out

"""ctypes binding of libluzrt.so (include/luzrt.h).  This is the harness the tests and bench.py use
to call the C ABI exactly as a C++ Luz host would; it adds nothing but argument marshalling.
There is no fallback: if the library is missing or no sm_100 GPU is present, it raises."""
import ctypes as C
import os

import numpy as np

from . import wire

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libluzrt.so")

# luzrt_read selectors
IMG_LIGHT, IMG_HISTORY, SHADOW_MASK, AO_MASK, STATS = 0, 1, 2, 3, 4
GBUF_ALBEDO, GBUF_NORMAL, GBUF_MATERIAL, GBUF_EMISSION, GBUF_DEPTH, IMG_COMPOSE, TIMINGS = 5, 6, 7, 8, 9, 10, 11
STATS_DETAIL = 12
DEBUG_MASKS, DEBUG_STATS, DEBUG_NO_HINTS, DEBUG_EXACT_MATH, DEBUG_NO_TEMPORAL = 1, 2, 4, 8, 16

EXPORTS = [
    "luzrt_create", "luzrt_destroy", "luzrt_last_error", "luzrt_version", "luzrt_comm_unique_id",
    "luzrt_comm_init", "luzrt_resize", "luzrt_set_blue_noise", "luzrt_texture_create", "luzrt_blas_create",
    "luzrt_blas_destroy", "luzrt_blas_dump", "luzrt_tlas_dump", "luzrt_tlas_build", "luzrt_set_scene",
    "luzrt_set_gbuffer", "luzrt_gbuffer_pass", "luzrt_set_debug", "luzrt_light_pass", "luzrt_taa_pass",
    "luzrt_gather", "luzrt_compose_pass", "luzrt_swap_light_history", "luzrt_read", "luzrt_device_ptr",
    "luzrt_sync", "luzrt_stream", "luzrt_launch_count", "luzrt_read_rows", "luzrt_owned_bands", "luzrt_read_owned",
    "luzrt_prefetch_gbuffer", "luzrt_flip_gbuffer", "luzrt_read_owned_async", "luzrt_read_wait",
    "luzrt_probe_read_bandwidth", "luzrt_volumetric_pass", "luzrt_shadow_map_pass",
    "luzrt_read_shadow_map", "luzrt_create_multi", "luzrt_gather_multi", "luzrt_comm_check_bvh_multi", "luzrt_bvh_hash",
    "luzrt_comm_check_bvh",
]


class LuzError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("luzrt error %d: %s" % (code, msg))
        self.code = code


_lib = None


def load_library():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(LIB_PATH + " is not built (python -m luz_b200.build); there is no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    vp, u32, i32, u64 = C.c_void_p, C.c_uint32, C.c_int, C.c_uint64
    sig = {
        "luzrt_create": (i32, [i32, i32, i32, C.POINTER(vp)]),
        "luzrt_destroy": (None, [vp]),
        "luzrt_last_error": (C.c_char_p, [vp]),
        "luzrt_version": (C.c_char_p, []),
        "luzrt_comm_unique_id": (i32, [vp]),
        "luzrt_comm_init": (i32, [vp, vp]),
        "luzrt_resize": (i32, [vp, u32, u32]),
        "luzrt_set_blue_noise": (i32, [vp, vp, u32, u32]),
        "luzrt_texture_create": (i32, [vp, vp, u32, u32, C.POINTER(C.c_int32)]),
        "luzrt_blas_create": (i32, [vp, vp, u32, u32, vp, u32, C.POINTER(u32)]),
        "luzrt_blas_destroy": (i32, [vp, u32]),
        "luzrt_blas_dump": (i32, [vp, u32, vp, C.POINTER(C.c_size_t)]),
        "luzrt_tlas_dump": (i32, [vp, vp, C.POINTER(C.c_size_t)]),
        "luzrt_tlas_build": (i32, [vp, vp, u32, i32]),
        "luzrt_set_scene": (i32, [vp, vp, vp, u32]),
        "luzrt_set_gbuffer": (i32, [vp, vp, vp, vp, vp, vp, i32]),
        "luzrt_gbuffer_pass": (i32, [vp, vp, u32]),
        "luzrt_set_debug": (i32, [vp, u32]),
        "luzrt_light_pass": (i32, [vp, u32]),
        "luzrt_taa_pass": (i32, [vp, i32]),
        "luzrt_volumetric_pass": (i32, [vp, u32]),
        "luzrt_shadow_map_pass": (i32, [vp, u32]),
        "luzrt_read_shadow_map": (i32, [vp, u32, vp, C.c_size_t]),
        "luzrt_gather": (i32, [vp]),
        "luzrt_compose_pass": (i32, [vp, C.c_float]),
        "luzrt_swap_light_history": (i32, [vp]),
        "luzrt_read": (i32, [vp, i32, vp, C.c_size_t]),
        "luzrt_device_ptr": (i32, [vp, i32, C.POINTER(vp), C.POINTER(C.c_size_t)]),
        "luzrt_sync": (i32, [vp]),
        "luzrt_stream": (i32, [vp, C.POINTER(u64)]),
        "luzrt_launch_count": (u64, [vp]),
        "luzrt_read_rows": (i32, [vp, i32, u32, u32, vp, C.c_size_t]),
        "luzrt_owned_bands": (i32, [vp, C.POINTER(u32), C.POINTER(u32), C.POINTER(u32), C.POINTER(u32)]),
        "luzrt_read_owned": (i32, [vp, i32, vp, C.c_size_t]),
        "luzrt_prefetch_gbuffer": (i32, [vp, vp, vp, vp, vp, vp]),
        "luzrt_flip_gbuffer": (i32, [vp]),
        "luzrt_read_owned_async": (i32, [vp, i32, vp, C.c_size_t]),
        "luzrt_read_wait": (i32, [vp]),
        "luzrt_probe_read_bandwidth": (i32, [vp, C.c_size_t, i32, C.POINTER(C.c_double)]),
        "luzrt_create_multi": (i32, [C.POINTER(i32), i32, C.POINTER(vp)]),
        "luzrt_gather_multi": (i32, [C.POINTER(vp), i32]),
        "luzrt_comm_check_bvh_multi": (i32, [C.POINTER(vp), i32]),
        "luzrt_bvh_hash": (i32, [vp, C.POINTER(u64)]),
        "luzrt_comm_check_bvh": (i32, [vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def create_multi(device_ids):
    """One process, several GPUs: [LuzRT] with rank i on device_ids[i], communicators initialised (luzrt_create_multi)."""
    lib = load_library()
    n = len(device_ids)
    ids = (C.c_int * n)(*device_ids)
    hs = (C.c_void_p * n)()
    rc = lib.luzrt_create_multi(ids, n, hs)
    if rc != 0:
        raise LuzError(rc, "luzrt_create_multi failed")
    return [LuzRT(device=device_ids[i], rank=i, world=n, _handle=C.c_void_p(hs[i])) for i in range(n)]


def _group(ctxs):
    arr = (C.c_void_p * len(ctxs))(*[c.h for c in ctxs])
    return arr, len(ctxs)


def gather_multi(ctxs):
    arr, n = _group(ctxs)
    rc = load_library().luzrt_gather_multi(arr, n)
    if rc != 0:
        raise LuzError(rc, ctxs[0].lib.luzrt_last_error(ctxs[0].h).decode())


def comm_check_bvh_multi(ctxs):
    arr, n = _group(ctxs)
    rc = load_library().luzrt_comm_check_bvh_multi(arr, n)
    if rc != 0:
        raise LuzError(rc, "; ".join(c.lib.luzrt_last_error(c.h).decode() for c in ctxs))


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"]
        return a.ctypes.data_as(C.c_void_p)
    if isinstance(a, int):
        return C.c_void_p(a)
    return C.cast(C.byref(a), C.c_void_p)


class LuzRT:
    """One context == one GPU.  Method names follow the C ABI one to one."""

    def __init__(self, device=0, rank=0, world=1, _handle=None):
        self.lib = load_library()
        if _handle is not None:  # a ctx made by luzrt_create_multi
            h = _handle
        else:
            h = C.c_void_p()
            rc = self.lib.luzrt_create(device, rank, world, C.byref(h))
            if rc != 0:
                raise LuzError(rc, "luzrt_create failed (no usable sm_100 GPU?)")
        self.h = h
        self.width = self.height = 0
        self.rank, self.world = rank, world

    def close(self):
        if self.h:
            self.lib.luzrt_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise LuzError(rc, self.lib.luzrt_last_error(self.h).decode())

    def resize(self, w, h):
        self._ck(self.lib.luzrt_resize(self.h, w, h))
        self.width, self.height = w, h

    def set_blue_noise(self, rgba8):
        rgba8 = np.ascontiguousarray(rgba8, dtype=np.uint8)
        self._ck(self.lib.luzrt_set_blue_noise(self.h, _ptr(rgba8), rgba8.shape[1], rgba8.shape[0]))

    def texture_create(self, rgba8):
        rgba8 = np.ascontiguousarray(rgba8, dtype=np.uint8)
        rid = C.c_int32(-1)
        self._ck(self.lib.luzrt_texture_create(self.h, _ptr(rgba8), rgba8.shape[1], rgba8.shape[0], C.byref(rid)))
        return rid.value

    def blas_create(self, vertices, indices, stride=None):
        vertices = np.ascontiguousarray(vertices)
        indices = np.ascontiguousarray(indices, dtype=np.uint32)
        if stride is None:
            stride = vertices.strides[0] if vertices.ndim > 1 else 48
        nverts = vertices.nbytes // stride
        out = C.c_uint32(0)
        self._ck(self.lib.luzrt_blas_create(self.h, _ptr(vertices), nverts, stride, _ptr(indices), indices.size,
                                            C.byref(out)))
        return out.value

    def blas_destroy(self, blas):
        self._ck(self.lib.luzrt_blas_destroy(self.h, blas))

    def blas_dump(self, blas):
        n = C.c_size_t(0)
        self._ck(self.lib.luzrt_blas_dump(self.h, blas, None, C.byref(n)))
        buf = np.empty(n.value, dtype=np.uint8)
        self._ck(self.lib.luzrt_blas_dump(self.h, blas, _ptr(buf), C.byref(n)))
        return buf

    def tlas_dump(self):
        n = C.c_size_t(0)
        self._ck(self.lib.luzrt_tlas_dump(self.h, None, C.byref(n)))
        buf = np.empty(n.value, dtype=np.uint8)
        self._ck(self.lib.luzrt_tlas_dump(self.h, _ptr(buf), C.byref(n)))
        return buf

    @staticmethod
    def make_instances(blas_ids, mats, custom_indices=None):
        """mats: (n,16) float32 column-major."""
        n = len(blas_ids)
        arr = (wire.Instance * max(n, 1))()
        a = np.frombuffer(arr, dtype=np.dtype([("blas", "<u4"), ("m", "<f4", 16), ("ci", "<u4")]), count=max(n, 1))
        if n:
            a["blas"][:n] = np.asarray(blas_ids, dtype=np.uint32)
            a["m"][:n] = np.asarray(mats, dtype=np.float32).reshape(n, 16)
            a["ci"][:n] = np.arange(n, dtype=np.uint32) if custom_indices is None else np.asarray(custom_indices)
        return arr

    def tlas_build(self, instances, count, mode=0):
        self._ck(self.lib.luzrt_tlas_build(self.h, _ptr(instances) if count else None, count, mode))

    def set_scene(self, scene_block, extra_lights=None):
        n_extra = len(extra_lights) if extra_lights is not None else 0
        self._ck(self.lib.luzrt_set_scene(self.h, _ptr(scene_block), _ptr(extra_lights) if n_extra else None, n_extra))

    def set_gbuffer(self, albedo=None, normal=None, material=None, emission=None, depth=None, device=False):
        self._ck(self.lib.luzrt_set_gbuffer(self.h, _ptr(albedo), _ptr(normal), _ptr(material), _ptr(emission),
                                            _ptr(depth), 1 if device else 0))

    def gbuffer_pass(self, models, n_models):
        self._ck(self.lib.luzrt_gbuffer_pass(self.h, _ptr(models) if n_models else None, n_models))

    def set_debug(self, flags):
        self._ck(self.lib.luzrt_set_debug(self.h, flags))

    def light_pass(self, frame):
        self._ck(self.lib.luzrt_light_pass(self.h, frame))

    def shadow_map_pass(self, resolution=1024):
        self._ck(self.lib.luzrt_shadow_map_pass(self.h, resolution))

    def read_shadow_map(self, light, resolution, layers):
        out = np.zeros((layers, resolution, resolution), np.float32)
        self._ck(self.lib.luzrt_read_shadow_map(self.h, light, out.ctypes.data_as(C.c_void_p), out.nbytes))
        return out

    def volumetric_pass(self, frame):
        self._ck(self.lib.luzrt_volumetric_pass(self.h, frame))

    def taa_pass(self, reconstruct=True):
        self._ck(self.lib.luzrt_taa_pass(self.h, 1 if reconstruct else 0))

    def gather(self):
        self._ck(self.lib.luzrt_gather(self.h))

    def compose_pass(self, exposure=2.0):
        self._ck(self.lib.luzrt_compose_pass(self.h, exposure))

    def swap_light_history(self):
        self._ck(self.lib.luzrt_swap_light_history(self.h))

    def sync(self):
        self._ck(self.lib.luzrt_sync(self.h))

    def comm_unique_id(self):
        buf = np.zeros(128, dtype=np.uint8)
        rc = self.lib.luzrt_comm_unique_id(_ptr(buf))
        if rc != 0:
            raise LuzError(rc, "luzrt_comm_unique_id failed")
        return buf

    def comm_init(self, id128):
        id128 = np.ascontiguousarray(id128, dtype=np.uint8)
        self._ck(self.lib.luzrt_comm_init(self.h, _ptr(id128)))

    def bvh_hash(self):
        h = C.c_uint64(0)
        self._ck(self.lib.luzrt_bvh_hash(self.h, C.byref(h)))
        return h.value

    def comm_check_bvh(self):
        """Collective: raises if any rank's acceleration structures differ from this rank's."""
        self._ck(self.lib.luzrt_comm_check_bvh(self.h))

    def stream(self):
        s = C.c_uint64(0)
        self._ck(self.lib.luzrt_stream(self.h, C.byref(s)))
        return s.value

    def launch_count(self):
        return int(self.lib.luzrt_launch_count(self.h))

    def owned_bands(self):
        """(first_row, band_rows, pitch, n_bands): this ctx owns rows [first + k*pitch, + band_rows), k < n_bands."""
        v = [C.c_uint32() for _ in range(4)]
        self._ck(self.lib.luzrt_owned_bands(self.h, *[C.byref(x) for x in v]))
        return tuple(x.value for x in v)

    def prefetch_gbuffer(self, albedo, normal, material, emission, depth):
        self._ck(self.lib.luzrt_prefetch_gbuffer(self.h, _ptr(albedo), _ptr(normal), _ptr(material), _ptr(emission), _ptr(depth)))

    def flip_gbuffer(self):
        self._ck(self.lib.luzrt_flip_gbuffer(self.h))

    def read_owned_async(self, which, out):
        self._ck(self.lib.luzrt_read_owned_async(self.h, which, _ptr(out), out.nbytes))

    def read_wait(self):
        self._ck(self.lib.luzrt_read_wait(self.h))

    def read_owned(self, which, out):
        self._ck(self.lib.luzrt_read_owned(self.h, which, _ptr(out), out.nbytes))
        return out

    def read_rows(self, which, y0, y1, out):
        self._ck(self.lib.luzrt_read_rows(self.h, which, y0, y1, _ptr(out), out.nbytes))
        return out

    def probe_read_bandwidth(self, nbytes, iters):
        g = C.c_double(0.0)
        self._ck(self.lib.luzrt_probe_read_bandwidth(self.h, nbytes, iters, C.byref(g)))
        return g.value

    def device_ptr(self, which):
        p, n = C.c_void_p(), C.c_size_t()
        self._ck(self.lib.luzrt_device_ptr(self.h, which, C.byref(p), C.byref(n)))
        return p.value, n.value

    def read(self, which, out=None):
        w, h = self.width, self.height
        if which in (IMG_LIGHT, IMG_HISTORY, GBUF_NORMAL):
            buf = np.empty((h, w, 4), dtype=np.float32) if out is None else out
        elif which in (GBUF_ALBEDO, GBUF_MATERIAL, GBUF_EMISSION, IMG_COMPOSE):
            buf = np.empty((h, w, 4), dtype=np.uint8) if out is None else out
        elif which == GBUF_DEPTH:
            buf = np.empty((h, w), dtype=np.float32) if out is None else out
        elif which in (SHADOW_MASK, AO_MASK):
            _, n = self.device_ptr(which)
            buf = np.empty((h, w, n // (4 * w * h)), dtype=np.uint32)
        elif which == STATS:
            s = wire.Stats()
            self._ck(self.lib.luzrt_read(self.h, which, _ptr(s), C.sizeof(s)))
            return s
        elif which == STATS_DETAIL:
            d = np.zeros(40, np.uint64)
            self._ck(self.lib.luzrt_read(self.h, which, _ptr(d), d.nbytes))
            return d
        elif which == TIMINGS:
            t = wire.Timings()
            self._ck(self.lib.luzrt_read(self.h, which, _ptr(t), C.sizeof(t)))
            return t
        else:
            raise ValueError(which)
        self._ck(self.lib.luzrt_read(self.h, which, _ptr(buf), buf.nbytes))
        return buf

"""ctypes mirrors of include/luz_wire.h (== source/Shaders/LuzCommon.h of the reference) and of
the small structs in include/luzrt.h.  Harness-side only: the product code is the C++/CUDA in
csrc/ and host/."""
import ctypes as C

LUZ_MAX_LIGHTS = 64
LUZ_MAX_MODELS = 8192
LIGHT_POINT, LIGHT_SPOT, LIGHT_DIRECTIONAL = 0, 1, 2
SHADOW_DISABLED, SHADOW_RAYTRACING, SHADOW_MAP = 0, 1, 2

F = C.c_float
I = C.c_int32
U = C.c_uint32


class LightBlock(C.Structure):
    _fields_ = [
        ("color", F * 3), ("intensity", F),
        ("position", F * 3), ("inner_angle", F),
        ("direction", F * 3), ("outer_angle", F),
        ("type", I), ("num_shadow_samples", I), ("radius", F), ("shadow_map", I),
        ("view_proj", (F * 16) * 6),
        ("z_far", F), ("volumetric_type", I), ("volumetric_weight", F), ("volumetric_absorption", F),
        ("volumetric_density", F), ("volumetric_samples", I), ("pad", I * 2),
    ]


class ModelBlock(C.Structure):
    _fields_ = [
        ("model_mat", F * 16), ("color", F * 4), ("emission", F * 3), ("metallic", F),
        ("roughness", F), ("ao_map", I), ("color_map", I), ("normal_map", I),
        ("emission_map", I), ("metallic_roughness_map", I), ("vertex_buffer", I), ("index_buffer", I),
    ]


class SceneBlock(C.Structure):
    _fields_ = [
        ("lights", LightBlock * LUZ_MAX_LIGHTS),
        ("ambient_light_color", F * 3), ("ambient_light_intensity", F),
        ("proj", F * 16), ("view", F * 16), ("view_proj", F * 16), ("prev_view_proj", F * 16),
        ("inverse_proj", F * 16), ("inverse_view", F * 16),
        ("jitter", F * 2), ("prev_jitter", F * 2),
        ("cam_pos", F * 3), ("num_lights", I),
        ("ao_min", F), ("ao_max", F), ("exposure", F), ("ao_num_samples", I),
        ("white_texture", I), ("black_texture", I), ("blue_noise_texture", I), ("tlas_rid", I),
        ("shadow_type", I), ("pad", I * 3),
    ]


class Instance(C.Structure):  # luzrt_instance
    _fields_ = [("blas", U), ("model_mat", F * 16), ("custom_index", U)]


class Stats(C.Structure):  # luzrt_stats
    _fields_ = [("lit_pixels", C.c_uint64), ("rays", C.c_uint64), ("nodes_visited", C.c_uint64),
                ("triangles_tested", C.c_uint64), ("instances_entered", C.c_uint64),
                ("rays_occluded", C.c_uint64)]


class Timings(C.Structure):  # luzrt_timings
    _fields_ = [("tlas_ms", F), ("gbuffer_ms", F), ("light_ms", F), ("taa_ms", F), ("gather_ms", F),
                ("compose_ms", F), ("volumetric_ms", F), ("shadow_map_ms", F), ("light_rays_ms", F),
                ("temporal_settled", F), ("temporal_on", F)]


assert C.sizeof(LightBlock) == 480
assert C.sizeof(ModelBlock) == 128
assert C.sizeof(SceneBlock) == 31200
assert C.sizeof(Instance) == 72

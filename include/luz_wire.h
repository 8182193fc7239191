/*
 * luz_wire.h -- byte-exact wire structs shared by the Luz host and the lighting path.
 *
 * These restate, in plain C (no glm), the layouts that the reference declares once for
 * both C++ and GLSL in source/Shaders/LuzCommon.h (LightBlock :37-61, ModelBlock :76-92,
 * SceneBlock :94-125, LightConstants :180-189, PostProcessingConstants :206-221) and the
 * mesh vertex of source/Resources/AssetManager.hpp:83-91.  Sizes/offsets are the ones a
 * g++ 13 build of the reference headers reports (SURVEY.md probe table, section 8 a8) and
 * are re-checked against the compiled reference by oracle/ref_dump.cpp.
 *
 * A Luz host can memcpy its own SceneBlock / ModelBlock / MeshVertex straight into these.
 */
#ifndef LUZ_WIRE_H
#define LUZ_WIRE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LUZW_MAX_LIGHTS 64   /* LuzCommon.h:20 */
#define LUZW_MAX_MODELS 8192 /* LuzCommon.h:21 */

#define LUZW_LIGHT_POINT 0       /* LuzCommon.h:24 */
#define LUZW_LIGHT_SPOT 1        /* LuzCommon.h:25 */
#define LUZW_LIGHT_DIRECTIONAL 2 /* LuzCommon.h:26 */

#define LUZW_SHADOW_DISABLED 0   /* AssetManager.hpp:41 */
#define LUZW_SHADOW_RAYTRACING 1 /* LuzCommon.h:28 */
#define LUZW_SHADOW_MAP 2        /* LuzCommon.h:29 */

#define LUZW_VOLUMETRIC_DISABLED 0     /* AssetManager.hpp:194 */
#define LUZW_VOLUMETRIC_SCREEN_SPACE 1 /* LuzCommon.h:31 */
#define LUZW_VOLUMETRIC_SHADOW_MAP 2   /* LuzCommon.h:32 */

typedef struct luzw_light_block { /* 480 B */
    float color[3];
    float intensity;
    float position[3];
    float inner_angle; /* radians (GPUScene.cpp:244) */
    float direction[3];
    float outer_angle; /* radians */
    int32_t type;
    int32_t num_shadow_samples;
    float radius;
    int32_t shadow_map;
    float view_proj[6][16]; /* column-major mat4 x6 (shadow maps; unused on this path) */
    float z_far;
    int32_t volumetric_type;
    float volumetric_weight;
    float volumetric_absorption;
    float volumetric_density;
    int32_t volumetric_samples;
    int32_t pad[2];
} luzw_light_block;

typedef struct luzw_model_block { /* 128 B */
    float model_mat[16];          /* column-major */
    float color[4];
    float emission[3];
    float metallic;
    float roughness;
    int32_t ao_map;
    int32_t color_map;
    int32_t normal_map;
    int32_t emission_map;
    int32_t metallic_roughness_map;
    int32_t vertex_buffer;
    int32_t index_buffer;
} luzw_model_block;

typedef struct luzw_scene_block { /* 31200 B */
    luzw_light_block lights[LUZW_MAX_LIGHTS];
    float ambient_light_color[3];
    float ambient_light_intensity;
    float proj[16];
    float view[16];
    float view_proj[16];
    float prev_view_proj[16];
    float inverse_proj[16];
    float inverse_view[16];
    float jitter[2];
    float prev_jitter[2];
    float cam_pos[3];
    int32_t num_lights;
    float ao_min;
    float ao_max;
    float exposure;
    int32_t ao_num_samples;
    int32_t white_texture;
    int32_t black_texture;
    int32_t blue_noise_texture;
    int32_t tlas_rid;
    int32_t shadow_type;
    int32_t pad[3];
} luzw_scene_block;

typedef struct luzw_light_constants { /* 32 B push constants, LuzCommon.h:180-189 */
    int32_t scene_buffer_index;
    int32_t model_buffer_index;
    int32_t frame;
    int32_t albedo_rid;
    int32_t normal_rid;
    int32_t material_rid;
    int32_t emission_rid;
    int32_t depth_rid;
} luzw_light_constants;

typedef struct luzw_post_constants { /* 52 B, LuzCommon.h:206-221 */
    int32_t light_input_rid;
    int32_t light_output_rid;
    int32_t light_history_rid;
    int32_t depth_rid;
    float size[2];
    int32_t scene_buffer_index;
    int32_t reconstruct;
    float delta_time;
    int32_t histogram_rid;
    int32_t histogram_average_rid;
    float histogram_min_log;
    float histogram_one_over_log;
} luzw_post_constants;

typedef struct luzw_mesh_vertex { /* 48 B, AssetManager.hpp:83-91 */
    float position[3];
    float normal[3];
    float tangent[4];
    float tex_coord[2];
} luzw_mesh_vertex;

#ifdef __cplusplus
}
#define LUZW_CHECK(c, m) static_assert(c, m)
#else
#define LUZW_CHECK(c, m) _Static_assert(c, m)
#endif

LUZW_CHECK(sizeof(luzw_light_block) == 480, "LightBlock is 480 B");
LUZW_CHECK(sizeof(luzw_model_block) == 128, "ModelBlock is 128 B");
LUZW_CHECK(sizeof(luzw_scene_block) == 31200, "SceneBlock is 31200 B");
LUZW_CHECK(sizeof(luzw_light_constants) == 32, "LightConstants is 32 B");
LUZW_CHECK(sizeof(luzw_post_constants) == 52, "PostProcessingConstants is 52 B");
LUZW_CHECK(sizeof(luzw_mesh_vertex) == 48, "MeshVertex is 48 B");
LUZW_CHECK(offsetof(luzw_scene_block, ambient_light_color) == 30720, "SceneBlock.ambientLightColor");
LUZW_CHECK(offsetof(luzw_scene_block, proj) == 30736, "SceneBlock.proj");
LUZW_CHECK(offsetof(luzw_scene_block, inverse_proj) == 30992, "SceneBlock.inverseProj");
LUZW_CHECK(offsetof(luzw_scene_block, inverse_view) == 31056, "SceneBlock.inverseView");
LUZW_CHECK(offsetof(luzw_scene_block, jitter) == 31120, "SceneBlock.jitter");
LUZW_CHECK(offsetof(luzw_scene_block, cam_pos) == 31136, "SceneBlock.camPos");
LUZW_CHECK(offsetof(luzw_scene_block, num_lights) == 31148, "SceneBlock.numLights");
LUZW_CHECK(offsetof(luzw_scene_block, ao_num_samples) == 31164, "SceneBlock.aoNumSamples");
LUZW_CHECK(offsetof(luzw_scene_block, shadow_type) == 31184, "SceneBlock.shadowType");
LUZW_CHECK(offsetof(luzw_post_constants, size) == 16, "PostProcessingConstants.size");

#endif /* LUZ_WIRE_H */

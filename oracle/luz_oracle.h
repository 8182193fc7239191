/*
 * luz_oracle.h -- C interface of the CPU oracle (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * The oracle restates, in scalar fp32 C++ built with -ffp-contract=off, the arithmetic of the
 * reference's lighting path: source/Shaders/light.frag, taa.comp, utils.glsl, present.frag
 * and (as an input producer) opaque.vert/frag, plus the any-hit ray-query semantics the
 * Vulkan driver supplies.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  libluzrt.so never does.
 *
 * PARITY PINNING: the reference ships no tests, golden images or known-answer vectors for this path and its Vulkan
 * pipeline cannot run in this environment (no Vulkan loader / glslang / lavapipe).  The restatement is pinned to the
 * reference's own code instead:
 *   - shader side: oracle/glsl_harness/ rewrites the reference's light.frag / taa.comp / utils.glsl / LuzCommon.h
 *     lexically for the host and compiles them against the reference's glm (oracle/_ref/libglsl_ref.so);
 *     tests/test_glsl_pin.py runs that next to this oracle on identical inputs (every ray-query bit identical, radiance
 *     and resolve equal to a few ulp; >= 70 % of the pixels bit-equal), and tests/golden/glsl_ref_*.npz keep its outputs
 *     for machines without /root/reference;
 *   - host side: the reference's own compiled loader / camera / struct layouts (oracle/ref_dump.cpp -> tests/golden);
 *   - the fixed-function units the shaders call (texture unit, VK_KHR_ray_query, rasteriser) are driver territory with no
 *     code in the reference tree: those are restated from the Vulkan rules the reference selects (sampler, cull mode,
 *     ray flags) and checked by known answers (tests/test_oracle_kat.py).
 */
#ifndef LUZ_ORACLE_H
#define LUZ_ORACLE_H

#include <stdint.h>
#include "../include/luz_wire.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_mesh {
    const void* vertices;   /* vertex_stride bytes each, position = first 12 bytes */
    uint32_t vertex_count;
    uint32_t vertex_stride;
    const uint32_t* indices;
    uint32_t index_count;
} orc_mesh;

typedef struct orc_instance {
    uint32_t mesh;          /* index into the mesh array */
    float model_mat[16];    /* column-major glm::mat4 */
    uint32_t custom_index;  /* index into ModelBlock[] */
} orc_instance;

typedef struct orc_texture {
    const uint8_t* rgba8;
    uint32_t width, height;
} orc_texture;

typedef struct orc_gbuffer {
    uint8_t* albedo;   /* RGBA8   */
    float* normal;     /* RGBA32F */
    uint8_t* material; /* RGBA8   */
    uint8_t* emission; /* RGBA8   */
    float* depth;      /* F32     */
} orc_gbuffer;

typedef struct orc_stats {
    uint64_t lit_pixels;
    uint64_t rays;
    uint64_t rays_occluded;
} orc_stats;

typedef struct orc_world orc_world;

/* Builds per-mesh BVH2s and a BVH2 over instances (used only when exhaustive == 0). */
orc_world* orc_world_create(const orc_mesh* meshes, uint32_t n_meshes, const orc_instance* instances,
                            uint32_t n_instances);
void orc_world_destroy(orc_world* w);
void orc_set_threads(int n); /* 0 = all cores */
int orc_get_threads(void);

/* Any-hit over n rays: hit[i] = 1 if some triangle is hit with tmin < t < tmax.
 * exhaustive != 0: every triangle of every instance is tested (truth); else BVH2 traversal. */
void orc_trace_any(const orc_world* w, uint32_t n, const float* origins3, const float* dirs3, const float* tmin,
                   const float* tmax, int exhaustive, uint8_t* hit);
/* Closest hit: t[i] (inf if none), inst[i], prim[i]. */
void orc_trace_closest(const orc_world* w, uint32_t n, const float* origins3, const float* dirs3, const float* tmin,
                       const float* tmax, int exhaustive, float* t, int32_t* inst, int32_t* prim);

/* light.frag main() over rows [y0, y1).  out_rgba32f is the full W*H*4 image (only the rows
 * are written).  Masks may be NULL.  Returns 0, or -1 if shadowType == 2 and a light's map is not bound
 * (orc_bind_shadow_maps). */
int orc_light_pass(const luzw_scene_block* scene, const luzw_light_block* extra_lights, uint32_t n_extra,
                   uint32_t width, uint32_t height, const orc_gbuffer* gb, uint32_t frame,
                   const uint8_t* blue_noise_rgba8, uint32_t bn_w, uint32_t bn_h, const orc_world* world,
                   int exhaustive, uint32_t y0, uint32_t y1, float* out_rgba32f, uint32_t* shadow_mask,
                   uint32_t shadow_words, uint32_t* ao_mask, uint32_t ao_words, orc_stats* stats);
/* The same over columns [x0, x1) of the n_rows rows listed in row_list (any order, each < height).
 * exhaustive: 0 BVH2; 1 every triangle of every instance; 2 every instance whose world box the ray segment meets,
 * every triangle inside (no hierarchy; what validates the BVH2 on 10 M-triangle scenes). */
int orc_light_pass_rows(const luzw_scene_block* scene, const luzw_light_block* extra_lights, uint32_t n_extra,
                        uint32_t width, uint32_t height, const orc_gbuffer* gb, uint32_t frame,
                        const uint8_t* blue_noise_rgba8, uint32_t bn_w, uint32_t bn_h, const orc_world* world,
                        int exhaustive, const uint32_t* row_list, uint32_t n_rows, uint32_t x0, uint32_t x1,
                        float* out_rgba32f, uint32_t* shadow_mask, uint32_t shadow_words, uint32_t* ao_mask, uint32_t ao_words,
                        orc_stats* stats);

/* taa.comp main() over rows [y0, y1). */
int orc_taa_pass(const luzw_scene_block* scene, uint32_t width, uint32_t height, const float* light_in,
                 const float* history, const float* depth, int reconstruct, uint32_t y0, uint32_t y1,
                 float* out_rgba32f);

/* screenSpaceVolumetricLight.comp main() over rows [y0, y1): adds into light_inout (full frame RGBA32F). */
int orc_volumetric_screen_pass(const luzw_scene_block* scene, const luzw_light_block* extra_lights, uint32_t n_extra,
                               uint32_t width, uint32_t height, const float* depth, const uint8_t* blue_noise_rgba8,
                               uint32_t bn_w, uint32_t bn_h, uint32_t frame, uint32_t y0, uint32_t y1,
                               float* light_inout);

/* One light's shadow map, as orc_shadow_map_pass writes it: layers * res * res floats (6 layers for a point
 * light, else 1). */
typedef struct orc_shadow_map {
    const float* data;
    uint32_t res;
    uint32_t layers;
} orc_shadow_map;
/* DeferredRenderer::ShadowMapPass for one light (exhaustive, per texel).  Returns -1 for a singular viewProj[0]. */
int orc_shadow_map_pass(const luzw_light_block* light, const orc_world* world, uint32_t res, float* out);
/* The SHADOW_TYPE_MAP branch of EvaluateShadow (light.frag:147-165) for one point: 1 = in shadow. */
float orc_shadow_factor(const luzw_light_block* light, const orc_shadow_map* map, const float* frag_pos,
                        const float* shadow_origin);
/* The maps orc_light_pass (shadowType 2) and orc_volumetric_shadow_pass sample, indexed by light; the array
 * must stay alive while they run.  NULL / 0 unbinds. */
void orc_bind_shadow_maps(const orc_shadow_map* maps, uint32_t n);
/* shadowMapVolumetricLight.comp main() over rows [y0, y1): adds into light_inout. */
int orc_volumetric_shadow_pass(const luzw_scene_block* scene, const luzw_light_block* extra_lights, uint32_t n_extra,
                               uint32_t width, uint32_t height, const float* depth, const uint8_t* blue_noise_rgba8,
                               uint32_t bn_w, uint32_t bn_h, uint32_t frame, uint32_t y0, uint32_t y1,
                               float* light_inout);

/* present.frag imageType 0 -> BGRA8. */
int orc_compose_pass(uint32_t width, uint32_t height, const float* light_in, uint8_t* out_bgra8);

/* opaque.vert/frag as primary visibility (input producer), rows [y0, y1) of full-frame buffers. */
int orc_gbuffer_pass(const luzw_scene_block* scene, const orc_world* world, const luzw_model_block* models,
                     uint32_t n_models, const orc_texture* textures, uint32_t n_textures, uint32_t width,
                     uint32_t height, int exhaustive, uint32_t y0, uint32_t y1, orc_gbuffer* out);

/* Small exported helpers for known-answer tests. */
void orc_blue_noise_sample(const uint8_t* bn, uint32_t bn_w, uint32_t bn_h, uint32_t px, uint32_t py, int i,
                           uint32_t frame, float out2[2]);
void orc_depth_to_world(const luzw_scene_block* scene, float u, float v, float depth, float out3[3]);
float orc_mitchell(float x);
int orc_tri_test(const float* v0, const float* v1, const float* v2, const float* org, const float* dir, float tmin,
                 float tmax, float* t_out);

#ifdef __cplusplus
}
#endif
#endif

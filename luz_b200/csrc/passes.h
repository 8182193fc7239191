// passes.h -- launch interfaces of the frame kernels (light, TAA, compose, G-buffer).
#pragma once

#include <cuda_runtime.h>

#include "common.cuh"
#include "shadow_map.cuh"

namespace luz {

struct LightArgs {
    FrameConst fc;
    const uchar4* albedo;
    const float4* normal;
    const uchar4* material;
    const uchar4* emission;
    const float* depth;
    const uchar4* blue_noise;
    const LightRec* lights;
    float4* out;
    TraceScene scene;
    BandSet rows;       // rows to shade: own bands plus the halo rows TAA's 3x3 taps read
    uint32_t* shadow_mask;
    uint32_t shadow_words;
    uint32_t* ao_mask;
    uint32_t ao_words;
    DeviceStats* stats;
    unsigned long long* lit_counters; // 64 counters, 128 B apart
    uint32_t count_row_begin, count_row_end; // rows (relative to a band's first row) whose lit pixels are counted
    uint32_t cand_offset;                    // byte offset of the AO candidate lists in dynamic shared memory
    uint32_t* hints;                         // occluder hints [tile][light] of the shadow rays (nullptr: none), light_pass.cu
    uint32_t hint_sx, hint_sy;               // a hint tile is (1 << hint_sx) x (1 << hint_sy) pixels: 16 x 8 or 8 x 4
    uint32_t* tile_counter;                  // work counter of the persistent ray kernel (reset per launch)
    uint2* ray_hints;                        // per-ray temporal occluder hints [shadow bit][pixel][slots] = (instance, triangle): two
                                             // slots per ray while shadow_words <= 2, one beyond; nullptr: off
    uint32_t n_instances;                    // instances of the current TLAS (bounds of the hints)
    const uint32_t* inst_order;              // leaf position -> index in the host's instance array (tlas.prim_order)
    const uint32_t* inst_leaf;               // index in the host's instance array -> leaf position
    unsigned long long* temporal_counters;   // 64 x {shadow rays fired, settled by their temporal hints}, 128 B apart
    uint32_t count_temporal;                 // LUZRT_TEMPORAL_COUNT=1: count hint tests / hits / queued rays in stats->detail[24..28]
    const ShadowMapRec* shadow_maps;         // per light, scene order (read only when fc.shadow_type == LUZW_SHADOW_MAP)
    const float* pow22;                      // 256 floats: (c / 255)^2.2 (launch_pow22_table)
    uint32_t exact_math;                     // 1: the bit-faithful shading kernel (LUZRT_DEBUG_EXACT_MATH)
};

struct TaaArgs {
    FrameConst fc;
    const float4* light_in;
    const float4* history;
    const float* depth;
    float4* out;
    BandSet rows; // own bands
    int reconstruct;
};

struct GbufferArgs {
    FrameConst fc;
    TraceScene scene;
    const InstanceMeta* inst_meta;
    const BlasAttr* blas_attr;
    const luzw_model_block* models;
    uint32_t n_models;
    const uchar4* const* tex_data; // per RID: device pointer (or null)
    const uint2* tex_size;
    uint32_t n_textures;
    uchar4* albedo;
    float4* normal;
    uchar4* material;
    uchar4* emission;
    float* depth;
    BandSet rows;
    float cull_sign; // -s_view of the camera: back faces are culled (api.cu luzrt_gbuffer_pass)
};

// one light with a volumetric type (LightBlock fields the volumetric shaders read), scene order
struct __align__(16) VolLight {
    float4 color_intensity;
    float4 position_type;        // w: light type (int bits)
    float4 direction_absorption; // w: volumetricAbsorption
    int samples;
    int volumetric_type;
    int light_index;
    int pad;
};
static_assert(sizeof(VolLight) == 64, "VolLight");

struct VolumetricArgs {
    FrameConst fc; // frame_mod, bn_w, bn_h filled like the light pass
    const float* depth;
    const uchar4* blue_noise;
    const VolLight* lights;
    int n_lights;
    float4* light; // lightA, read-modify-write
    BandSet rows;  // the rows the light pass shaded (own bands + halo rows)
    const ShadowMapRec* shadow_maps; // per light, scene order (shadow-map volumetrics)
};

// one shadow map (all its layers) of one light
struct ShadowMapArgs {
    TraceScene scene;
    float* out; // layers * res * res
    uint32_t res, layers;
    int is_cube;
    float eye[3]; // light.position (cube)
    float z_far;
    float inv_view_proj[16]; // inverse of light.viewProj[0], column-major (2-D)
    float cull_sign[6];      // s_view per layer (shadow_map.cu)
};

cudaError_t launch_volumetric_screen(cudaStream_t stream, const VolumetricArgs& args);
cudaError_t launch_volumetric_shadow_map(cudaStream_t stream, const VolumetricArgs& args);
cudaError_t launch_shadow_map(cudaStream_t stream, const ShadowMapArgs& args);
// ray kernel + shading kernel (light_pass.cu); the per-pixel visibility masks are cleared, written and read by it
// aux / ev_fork / ev_join: a second stream and two events of the caller; the hint pass + shadow-ray launch run there while
// the AO-ray launch runs on `stream`, joined before the shading kernel (all nullptr: one stream)
cudaError_t launch_light_pass(cudaStream_t stream, const LightArgs& args, bool stats, cudaEvent_t rays_done, uint64_t* launches,
                              cudaStream_t aux, cudaEvent_t ev_fork, cudaEvent_t ev_join);
// exact: the bit-faithful build (LUZRT_DEBUG_EXACT_MATH) instead of the relaxed-precision one (relaxed.cu)
cudaError_t launch_taa_pass(cudaStream_t stream, const TaaArgs& args, bool exact);
cudaError_t launch_taa_relaxed(cudaStream_t stream, const TaaArgs& args);
cudaError_t launch_light_shade_relaxed(cudaStream_t stream, const LightArgs& args);
// fills the 256-entry table (c / 255)^2.2 the shading kernels look albedo up in
cudaError_t launch_pow22_table(cudaStream_t stream, float* table256);
cudaError_t launch_compose_pass(cudaStream_t stream, const FrameConst& fc, const float4* light_in, uchar4* out_bgra,
                                const BandSet& rows);
// copies rows between the banded storage order of the light images and the natural row order
cudaError_t launch_unpermute_rows(cudaStream_t stream, const FrameConst& fc, const float4* banded, float4* natural,
                                  uint32_t y0, uint32_t y1);
cudaError_t launch_gbuffer_pass(cudaStream_t stream, const GbufferArgs& args);
// bandwidth probe: every SM streams `bytes` of `buf` `iters` times with 128-bit loads
cudaError_t launch_probe_read(cudaStream_t stream, const void* buf, size_t bytes, int iters, float* sink);

} // namespace luz

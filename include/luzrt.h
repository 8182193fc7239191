/*
 * luzrt.h -- C ABI of libluzrt.so, the B200 (sm_100a) implementation of Luz's
 * ray-traced deferred lighting path.
 *
 * Luz has no plugin/FFI layer: the seam this library sits behind is the set of C++ calls
 * that main.cpp RenderFrame (source/Core/main.cpp:223-311) and GPUScene
 * (source/Graphics/GPUScene.cpp:108-138, :348-366) make into vkw:: and DeferredRenderer::.
 * Each entry point below names the reference call it replaces.  INTEGRATION.md shows the
 * few lines a Luz maintainer adds to GPUScene.cpp / DeferredRenderer.cpp to route through it.
 *
 * Conventions
 *  - plain pointers and sizes only; no C++/torch types cross this boundary.
 *  - every call returns LUZRT_OK (0) or a negative luzrt_status; nothing aborts or throws.
 *    (The reference aborts via ASSERT/LOG_CRITICAL, Base.hpp:51-59; a library cannot.)
 *  - luzrt_last_error(ctx) gives the message for the last failing call on that ctx.
 *  - the caller owns every input buffer.  Pageable host buffers are staged before the call returns
 *    and may be freed at once; page-locked (cudaHostAlloc) buffers are read by DMA and must stay
 *    valid until the next luzrt_sync / luzrt_read on the ctx (the usual CUDA rule).
 *  - calls are asynchronous and ordered on the ctx's CUDA stream, except luzrt_read,
 *    luzrt_sync, luzrt_blas_create (blocking like GPUScene::AddMesh's WaitQueue,
 *    GPUScene.cpp:132-137) and luzrt_create/destroy.
 *  - one host thread per ctx (the reference is single threaded).  The only process-global state is the
 *    lazily dlopen'ed NCCL library handle shared by every ctx (loaded on the first luzrt_comm_* call).
 *  - there is no CPU fallback: if no sm_100-class CUDA device is usable luzrt_create fails.
 *
 * Images are row-major, row 0 = top (Vulkan viewport, VulkanWrapper.cpp:1216-1222).
 */
#ifndef LUZRT_H
#define LUZRT_H

#include <stddef.h>
#include <stdint.h>

#include "luz_wire.h"

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define LUZRT_API
#else
#define LUZRT_API __attribute__((visibility("default")))
#endif

typedef struct luzrt_ctx luzrt_ctx;
typedef uint32_t luzrt_blas; /* opaque handle, 0 = invalid; owned by the ctx */

typedef enum luzrt_status {
    LUZRT_OK = 0,
    LUZRT_E_INVALID = -1, /* bad argument                                    */
    LUZRT_E_CUDA = -2,    /* a CUDA runtime call failed (message has detail) */
    LUZRT_E_NOMEM = -3,   /* device or pinned allocation failed              */
    LUZRT_E_STATE = -4,   /* call order violated (e.g. light pass before resize) */
    LUZRT_E_COMM = -5,    /* NCCL failure                                    */
    LUZRT_E_NODEVICE = -6 /* no usable sm_100 device                         */
} luzrt_status;

/* One TLAS instance; mirrors vkw::BLASInstance {blas, modelMat, customIndex}
 * (source/Graphics/VulkanWrapper.h:196-200).  model_mat is a glm::mat4 as stored
 * (column-major); rows 0..2 are used, like VulkanWrapper.cpp:1122-1126. */
typedef struct luzrt_instance {
    luzrt_blas blas;
    float model_mat[16];
    uint32_t custom_index;
} luzrt_instance;

/* luzrt_read selectors */
enum {
    LUZRT_IMG_LIGHT = 0,    /* current lightA, RGBA32F  W*H*16 B (what ComposePass reads)     */
    LUZRT_IMG_HISTORY = 1,  /* lightHistory, RGBA32F                                          */
    LUZRT_SHADOW_MASK = 2,  /* u32[H][W][shadow_words]; bit (l*S+s) set = ray occluded        */
    LUZRT_AO_MASK = 3,      /* u32[H][W][ao_words];     bit i set       = ray occluded        */
    LUZRT_STATS = 4,        /* luzrt_stats of the last light pass (needs LUZRT_DEBUG_STATS)   */
    LUZRT_GBUF_ALBEDO = 5,  /* RGBA8                                                          */
    LUZRT_GBUF_NORMAL = 6,  /* RGBA32F                                                        */
    LUZRT_GBUF_MATERIAL = 7,/* RGBA8                                                          */
    LUZRT_GBUF_EMISSION = 8,/* RGBA8                                                          */
    LUZRT_GBUF_DEPTH = 9,   /* F32                                                            */
    LUZRT_IMG_COMPOSE = 10, /* BGRA8 from luzrt_compose_pass                                  */
    LUZRT_TIMINGS = 11,     /* luzrt_timings of the last frame (CUDA events, ms)              */
    LUZRT_STATS_DETAIL = 12 /* uint64[40]: per ray class break-down of the last light pass run with LUZRT_DEBUG_STATS
                               (layout: csrc/common.cuh DeviceStats::detail; measurement aid, profiles/)   */
};

/* luzrt_set_debug flags */
enum {
    LUZRT_DEBUG_MASKS = 1, /* kept for ABI stability: the light pass always writes the per-ray visibility
                              bitmasks (they carry the rays' results to its shading kernel)            */
    LUZRT_DEBUG_STATS = 2, /* light pass counts nodes / triangles / instances per ray */
    LUZRT_DEBUG_NO_HINTS = 4, /* shadow rays descend from the TLAS root without trying the tile's occluder hint first */
    LUZRT_DEBUG_NO_TEMPORAL = 16, /* shadow rays do not try the occluder they found last frame first (per-ray temporal hints
                                    off: every frame costs what the first frame costs; same bits) */
    LUZRT_DEBUG_EXACT_MATH = 8 /* luzrt_light_pass / luzrt_taa_pass run the bit-faithful builds of the shading and resolve
                                  kernels (every operation of light.frag / taa.comp in the shader's order, IEEE division
                                  and square root, no FMA contraction) instead of the relaxed-precision ones a host gets
                                  by default (MUFU reciprocals, contraction; ~1e-6 relative, inside the 1e-3 / 50 dB
                                  tolerance).  Ray visibility is identical in both.                                  */
};

typedef struct luzrt_stats {
    uint64_t lit_pixels;        /* pixels with length(N) != 0                         */
    uint64_t rays;              /* shadow + AO any-hit rays actually traced           */
    uint64_t nodes_visited;     /* 8-wide nodes fetched, TLAS + BLAS (this build: 208-B fp32 nodes; the
                                   roofline's algorithmic figure is SURVEY 8(d)'s 80 B)              */
    uint64_t triangles_tested;  /* triangles fetched (this build: 96-B Pluecker records; SURVEY 8(d): 48 B) */
    uint64_t instances_entered; /* 64-B instance records fetched                      */
    uint64_t rays_occluded;
} luzrt_stats;

typedef struct luzrt_timings {
    float tlas_ms;    /* last luzrt_tlas_build     ("GPUScene::BuildTLAS", GPUScene.cpp:354) */
    float gbuffer_ms; /* last luzrt_gbuffer_pass   ("OpaquePass", main.cpp:242)              */
    float light_ms;   /* last luzrt_light_pass     ("LightPass", main.cpp:266)               */
    float taa_ms;     /* last luzrt_taa_pass       ("TAAPass", main.cpp:281)                 */
    float gather_ms;  /* last luzrt_gather                                                    */
    float compose_ms; /* last luzrt_compose_pass   ("ComposePass", main.cpp:293)             */
    float volumetric_ms; /* last luzrt_volumetric_pass ("VolumetricLightPass", main.cpp:274)    */
    float shadow_map_ms; /* last luzrt_shadow_map_pass ("ShadowMaps", main.cpp:260)              */
    float light_rays_ms; /* the ray kernel's share of light_ms (mask clears + k_light_rays)      */
    float temporal_settled; /* fraction of the shadow rays of the last measured frame that their temporal occluder hints
                               settled (-1: never measured); temporal_on: 1 if the last light pass used the hints */
    float temporal_on;
} luzrt_timings;

/* ---- lifetime ------------------------------------------------------------------------- */

/* Replaces vkw::Init + DeferredRenderer::CreateResources for this path.  One ctx drives one
 * GPU.  For an image-partitioned multi-GPU frame create one ctx per GPU with its rank in [0, world)
 * (one process per GPU, or all of them in one process through luzrt_create_multi).  The frame is cut
 * into bands of rows that are dealt ROUND-ROBIN (band b belongs to rank b % world, see luzrt_owned_bands):
 * a rank does NOT own one contiguous strip.  Hosts place luzrt_read_owned / luzrt_device_ptr data with
 * luzrt_owned_bands, never by assuming [r*H/world, (r+1)*H/world).  world = 1 for a single GPU. */
LUZRT_API int luzrt_create(int device_id, int rank, int world, luzrt_ctx** out);
LUZRT_API void luzrt_destroy(luzrt_ctx* ctx);
LUZRT_API const char* luzrt_last_error(luzrt_ctx* ctx);
LUZRT_API const char* luzrt_version(void);

/* Multi-GPU plumbing: rank 0 calls luzrt_comm_unique_id, the host broadcasts the 128 bytes
 * by any means (bench.py uses torch.distributed), every rank calls luzrt_comm_init. */
LUZRT_API int luzrt_comm_unique_id(void* out128);
LUZRT_API int luzrt_comm_init(luzrt_ctx* ctx, const void* id128);

/* One process, several GPUs (Luz is one process and one thread, main.cpp:356-366): creates n_devices contexts, ctx i on
 * device_ids[i] with rank i of n_devices, and their communicators in one go (ncclCommInitAll; nothing to broadcast).
 * out receives n_devices pointers; each is driven like any ctx (the same calls in the same order on every one, each
 * followed through before the next is fine: calls are asynchronous) and destroyed with luzrt_destroy.  The two
 * collectives of contexts that share a thread go through the _multi entry points (one NCCL group): */
LUZRT_API int luzrt_create_multi(const int* device_ids, int n_devices, luzrt_ctx** out);
LUZRT_API int luzrt_gather_multi(luzrt_ctx** ctxs, int n);           /* == luzrt_gather on every ctx          */
LUZRT_API int luzrt_comm_check_bvh_multi(luzrt_ctx** ctxs, int n);   /* == luzrt_comm_check_bvh on every ctx  */

/* Content hash of everything rays traverse on this ctx (TLAS nodes, leaf order, instance transforms and world boxes,
 * nodes and triangles of every live BLAS).  Equal scenes give equal hashes on every run and every GPU. */
LUZRT_API int luzrt_bvh_hash(luzrt_ctx* ctx, uint64_t* out);
/* Collective over the communicator: every rank hashes its replica of the acceleration structures, the hashes are
 * all-gathered and compared; LUZRT_E_STATE names the first rank that differs.  ("BVH build deterministic across GPUs":
 * each rank builds its own replica instead of receiving a broadcast, so this is the check that they agree.) */
LUZRT_API int luzrt_comm_check_bvh(luzrt_ctx* ctx);

/* ---- resources ------------------------------------------------------------------------ */

/* == DeferredRenderer::CreateImages(w, h) (DeferredRenderer.cpp:175-248): (re)allocates the
 * G-buffer (RGBA8 albedo, RGBA32F normal, RGBA8 material, RGBA8 emission, D32F depth),
 * lightA / lightB / lightHistory (RGBA32F) and the BGRA8 compose target; history becomes
 * invalid (the first TAA pass after a resize uses the current light buffer as history). */
LUZRT_API int luzrt_resize(luzrt_ctx* ctx, uint32_t width, uint32_t height);

/* == the blue-noise upload in GPUScene::Create (GPUScene.cpp:57-74); RGBA8, any size. */
LUZRT_API int luzrt_set_blue_noise(luzrt_ctx* ctx, const uint8_t* rgba8, uint32_t width, uint32_t height);

/* == GPUScene::AddTexture (GPUScene.cpp:140-156); returns the bindless RID that
 * ModelBlock.colorMap etc. refer to.  Only used by luzrt_gbuffer_pass. */
LUZRT_API int luzrt_texture_create(luzrt_ctx* ctx, const uint8_t* rgba8, uint32_t width, uint32_t height,
                                   int32_t* out_rid);

/* == vkw::CreateBLAS + vkw::CmdBuildBLAS as driven by GPUScene::AddMesh
 * (VulkanWrapper.cpp:743-836, :1089-1104; GPUScene.cpp:108-138).  positions are the first
 * 12 bytes of each vertex_stride-byte vertex; index_count/3 opaque two-sided triangles.
 * When vertex_stride == 48 the normal/tangent/uv attributes are kept for luzrt_gbuffer_pass.
 * Deterministic: the same input gives a bitwise identical BVH on every run and every GPU. */
LUZRT_API int luzrt_blas_create(luzrt_ctx* ctx, const void* vertices, uint32_t vertex_count,
                                uint32_t vertex_stride, const uint32_t* indices, uint32_t index_count,
                                luzrt_blas* out);
LUZRT_API int luzrt_blas_destroy(luzrt_ctx* ctx, luzrt_blas blas);

/* Copies the built BLAS (nodes then triangles) to dst for determinism checks; *bytes is
 * in: capacity, out: size needed/written. */
LUZRT_API int luzrt_blas_dump(luzrt_ctx* ctx, luzrt_blas blas, void* dst, size_t* bytes);
LUZRT_API int luzrt_tlas_dump(luzrt_ctx* ctx, void* dst, size_t* bytes);

/* == vkw::CmdBuildTLAS (VulkanWrapper.cpp:1106-1139) as called every frame by
 * GPUScene::UpdateResourcesGPU (GPUScene.cpp:354-365).  mode 0 = full rebuild (what the
 * reference always does), mode 1 = refit (same instance count/BLAS set as the last rebuild,
 * only transforms changed; falls back to a rebuild when that does not hold). */
LUZRT_API int luzrt_tlas_build(luzrt_ctx* ctx, const luzrt_instance* instances, uint32_t count, int mode);

/* == the SceneBlock staging copy in GPUScene::UpdateResourcesGPU (GPUScene.cpp:353).
 * scene_block is the 31 200-byte SceneBlock.  Lights beyond LUZ_MAX_LIGHTS (benchmark
 * config 4) are passed as n_extra further LightBlocks; the pass then iterates
 * scene.numLights + n_extra lights. */
LUZRT_API int luzrt_set_scene(luzrt_ctx* ctx, const luzw_scene_block* scene_block,
                              const luzw_light_block* extra_lights, uint32_t n_extra);

/* The G-buffer the reference's opaque pass leaves in its attachments
 * (DeferredRenderer.cpp:176-238): full W*H images; NULL keeps the current contents.
 * src_is_device != 0: the pointers are device memory on this ctx's GPU.  The pointers always address
 * full-frame images; a ctx of a multi-GPU frame copies only the rows it shades. */
LUZRT_API int luzrt_set_gbuffer(luzrt_ctx* ctx, const void* albedo_rgba8, const void* normal_rgba32f,
                                const void* material_rgba8, const void* emission_rgba8,
                                const void* depth_f32, int src_is_device);

/* Pipelined host path (the reference keeps 3 frames in flight, VulkanWrapper.cpp:178): the five planes of the
 * NEXT frame are copied from page-locked host memory into a second G-buffer set on a copy stream while the
 * current frame is still being shaded; luzrt_flip_gbuffer then makes that set current (the ctx stream waits
 * for the copy).  The host buffers must stay valid until the flip. */
LUZRT_API int luzrt_prefetch_gbuffer(luzrt_ctx* ctx, const void* albedo_rgba8, const void* normal_rgba32f,
                                     const void* material_rgba8, const void* emission_rgba8, const void* depth_f32);
LUZRT_API int luzrt_flip_gbuffer(luzrt_ctx* ctx);

/* produces the same five attachments on the device by primary
 * visibility through the TLAS (closest hit), restating opaque.vert:21-31 / opaque.frag:21-59;
 * models[i] is addressed by luzrt_instance.custom_index (GPUScene.cpp:188-191). */
LUZRT_API int luzrt_gbuffer_pass(luzrt_ctx* ctx, const luzw_model_block* models, uint32_t n_models);

/* ---- the frame ------------------------------------------------------------------------ */

LUZRT_API int luzrt_set_debug(luzrt_ctx* ctx, uint32_t flags);

/* == DeferredRenderer::LightPass(LightConstants{.frameID = frame}) running light.frag
 * (DeferredRenderer.cpp:324-345, light.frag:171-235): writes lightA for this rank's rows
 * (plus the rows TAA's 3x3 taps need). */
LUZRT_API int luzrt_light_pass(luzrt_ctx* ctx, uint32_t frame);

/* == the loop over DeferredRenderer::ShadowMapPass(light, scene, gpuScene)
 * of RenderFrame (main.cpp:260-264, DeferredRenderer.cpp:268-291; shadowMap.vert/.geom/.frag): renders
 * the D32F shadow map of every light of the current scene block that is sampled on this path -- lights
 * with shadowMap != -1 when scene.shadowType == 2, and lights with volumetricType == 2 -- at
 * resolution^2 texels (scene->shadowResolution, default 1024, AssetManager.hpp:287): six cube-face
 * layers of |light.position - fragPos| / zFar for point lights, one layer of gl_FragCoord.z under the
 * orthographic light.viewProj[0] for spot / directional lights.  There is no rasteriser here: one
 * closest-hit ray per texel centre through the TLAS, keeping only the triangles the pipeline's
 * front-face culling keeps (see csrc/shadow_map.cu).  Call after luzrt_set_scene + luzrt_tlas_build
 * and before luzrt_light_pass / luzrt_volumetric_pass of the frame; both fail with LUZRT_E_STATE if
 * the maps are older than the scene block or the TLAS.  In light.shadowMap only "-1 / not -1" matters
 * (the reference stores a bindless texture index there, GPUScene.cpp:329). */
LUZRT_API int luzrt_shadow_map_pass(luzrt_ctx* ctx, uint32_t resolution);
/* Blocking read-back of light `light_index`'s map: layers * resolution^2 floats, layer-major, row 0 = top. */
LUZRT_API int luzrt_read_shadow_map(luzrt_ctx* ctx, uint32_t light_index, void* dst, size_t bytes);

/* == the block `if (gpuScene.AnyVolumetricLight())` of RenderFrame
 * (main.cpp:274-279): DeferredRenderer::ScreenSpaceVolumetricLightPass(gpuScene, frame)
 * (DeferredRenderer.cpp:294-307, screenSpaceVolumetricLight.comp:22-62) for the lights with
 * volumetricType == 1, then DeferredRenderer::ShadowMapVolumetricLightPass (DeferredRenderer.cpp:309-322,
 * shadowMapVolumetricLight.comp:42-74) for the lights with volumetricType == 2; both add into lightA,
 * after luzrt_light_pass and before luzrt_taa_pass.  A no-op when the scene block has no such light
 * (== AnyVolumetricLight() false).  Both passes read depth outside the rows a rank shades: with
 * world > 1, luzrt_set_gbuffer / luzrt_prefetch_gbuffer / luzrt_gbuffer_pass fill the whole depth
 * plane on every rank while the scene block set by luzrt_set_scene has such a light. */
LUZRT_API int luzrt_volumetric_pass(luzrt_ctx* ctx, uint32_t frame);

/* == DeferredRenderer::TAAPass with scene->taaEnabled (DeferredRenderer.cpp:425-445,
 * taa.comp:271-308) including its swap(lightA, lightB).  Not calling it == taaEnabled false. */
LUZRT_API int luzrt_taa_pass(luzrt_ctx* ctx, int reconstruct);

/* World > 1 only: NCCL all-gather of every rank's resolved rows of lightA so that each GPU
 * holds the whole frame (it becomes the next frame's TAA history).  No-op when world == 1. */
LUZRT_API int luzrt_gather(luzrt_ctx* ctx);

/* == DeferredRenderer::ComposePass for imageType 0
 * (present.frag:27-35, :85-95): exposure, ACES, gamma 1/2.2 -> BGRA8. */
LUZRT_API int luzrt_compose_pass(luzrt_ctx* ctx, float exposure);

/* == DeferredRenderer::SwapLightHistory (DeferredRenderer.cpp:469-471), end of frame. */
LUZRT_API int luzrt_swap_light_history(luzrt_ctx* ctx);

/* Blocking read-back of one of the LUZRT_* selectors into host memory. */
LUZRT_API int luzrt_read(luzrt_ctx* ctx, int which, void* dst, size_t bytes);
/* Same for image selectors, restricted to rows [y0, y1): dst receives (y1-y0) packed rows. */
LUZRT_API int luzrt_read_rows(luzrt_ctx* ctx, int which, uint32_t y0, uint32_t y1, void* dst, size_t bytes);
/* The rows this ctx owns.  A frame shared by `world` GPUs is cut into bands of *band_rows rows that are dealt
 * round-robin (band b belongs to rank b % world, which balances sky against geometry): this ctx owns rows
 * [*first_row + k * *pitch, + *band_rows) for k = 0 .. *n_bands-1.  world == 1: one band, the whole image. */
LUZRT_API int luzrt_owned_bands(luzrt_ctx* ctx, uint32_t* first_row, uint32_t* band_rows, uint32_t* pitch,
                                uint32_t* n_bands);
/* Blocking read-back of the rows this ctx owns, packed in band order (height / world rows). */
LUZRT_API int luzrt_read_owned(luzrt_ctx* ctx, int which, void* dst, size_t bytes);
/* Non-blocking variant for LUZRT_IMG_LIGHT / LUZRT_IMG_HISTORY: the copy into page-locked dst runs on a copy
 * stream after the work enqueued so far; luzrt_read_wait blocks until it has landed.  One read in flight. */
LUZRT_API int luzrt_read_owned_async(luzrt_ctx* ctx, int which, void* dst, size_t bytes);
LUZRT_API int luzrt_read_wait(luzrt_ctx* ctx);
/* Measurement aid ("L2 peak must be measured by the builder"): streams a
 * `bytes`-sized device buffer `iters` times with every SM and returns the achieved read GB/s.
 * A buffer well below the 126 MB L2 measures L2 bandwidth, one far above it measures HBM. */
LUZRT_API int luzrt_probe_read_bandwidth(luzrt_ctx* ctx, size_t bytes, int iters, double* out_gbs);
/* Device pointer of an image selector (for hosts that keep working on the GPU).  With world > 1 the two
 * light images are stored band-permuted (each rank's rows contiguous, rank-major): row y lives at
 * (b % world) * (H / world) + (b / world) * band_rows + y % band_rows with b = y / band_rows. */
LUZRT_API int luzrt_device_ptr(luzrt_ctx* ctx, int which, void** out_ptr, size_t* out_bytes);
LUZRT_API int luzrt_sync(luzrt_ctx* ctx);
/* The ctx's cudaStream_t as an integer, so a host can time with events on the same stream. */
LUZRT_API int luzrt_stream(luzrt_ctx* ctx, uint64_t* out_stream);
/* Count of kernels this ctx has launched so far (bench.py's gpu_launches). */
LUZRT_API uint64_t luzrt_launch_count(luzrt_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* LUZRT_H */

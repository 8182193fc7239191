// api.cu -- the C ABI of libluzrt.so (include/luzrt.h): context, resources and the frame calls.
// Every entry point cites the reference call it replaces in luzrt.h; nothing here computes on the
// CPU beyond marshalling (there is no CPU fallback by design).
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "bvh.h"
#include "passes.h"
#include "traverse.cuh"

using namespace luz;

namespace {

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool load() {
        if (lib) return true;
        // RTLD_NOLOAD first: share the copy a host process (e.g. torch) already mapped
        lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
        if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW);
        if (!lib) lib = dlopen("libnccl.so", RTLD_NOW);
        if (!lib) return false;
        GetUniqueId = (decltype(GetUniqueId))dlsym(lib, "ncclGetUniqueId");
        CommInitRank = (decltype(CommInitRank))dlsym(lib, "ncclCommInitRank");
        CommInitAll = (decltype(CommInitAll))dlsym(lib, "ncclCommInitAll");
        GroupStart = (decltype(GroupStart))dlsym(lib, "ncclGroupStart");
        GroupEnd = (decltype(GroupEnd))dlsym(lib, "ncclGroupEnd");
        AllGather = (decltype(AllGather))dlsym(lib, "ncclAllGather");
        CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
        GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
        return GetUniqueId && CommInitRank && CommInitAll && GroupStart && GroupEnd && AllGather && CommDestroy && GetErrorString;
    }
};
NcclApi g_nccl;

struct Blas {
    bool alive = false;
    WideBvh bvh;
    WideTri* tris = nullptr;
    uint8_t* verts = nullptr;
    uint32_t* indices = nullptr;
    uint32_t stride = 0, n_verts = 0, n_tris = 0;
    BoxF bounds{};
    uint32_t levels = 0;
};

struct Texture {
    uchar4* data = nullptr;
    uint2 size{0, 0};
};

enum { EV_TLAS, EV_GBUF, EV_LIGHT, EV_TAA, EV_GATHER, EV_COMPOSE, EV_VOLUMETRIC, EV_SHADOWMAP, EV_LIGHT_RAYS, EV_COUNT };

} // namespace

struct luzrt_ctx {
    int device = 0, rank = 0, world = 1;
    cudaStream_t stream = nullptr;
    std::string err;
    uint64_t launches = 0;
    uint32_t debug = 0;

    uint32_t w = 0, h = 0;
    uint32_t band_rows = 0, n_bands = 0, rows_per_rank = 0; // image partition
    float4* unperm = nullptr;                               // staging for reads of banded images (world > 1)
    size_t unperm_cap = 0;
    uchar4 *albedo = nullptr, *material = nullptr, *emission = nullptr, *compose = nullptr;
    float4 *normal = nullptr, *lightA = nullptr, *lightB = nullptr, *lightHist = nullptr;
    float* depth = nullptr;
    // back G-buffer set + streams of the pipelined host path (luzrt_prefetch_gbuffer / luzrt_read_owned_async)
    uchar4 *albedo_b = nullptr, *material_b = nullptr, *emission_b = nullptr;
    float4* normal_b = nullptr;
    float* depth_b = nullptr;
    cudaStream_t upload_stream = nullptr, download_stream = nullptr;
    cudaEvent_t ev_flip = nullptr, ev_uploaded = nullptr, ev_result = nullptr, ev_downloaded = nullptr;
    bool upload_pending = false, download_pending = false;
    const void* download_src = nullptr; // light image the asynchronous read-back in flight is reading
    // Page-locked staging ring for the small per-frame uploads (light records ...): a cudaMemcpyAsync from pageable
    // memory first synchronises the stream, which would stall the host behind the previous frame every frame.
    struct Staging {
        void* host = nullptr;
        size_t cap = 0;
        cudaEvent_t done = nullptr;
    } staging[4];
    uint32_t staging_next = 0;
    bool history_valid = false;

    uchar4* blue_noise = nullptr;
    uint32_t bn_w = 0, bn_h = 0;

    bool have_scene = false;
    FrameConst fc{};
    LightRec* d_lights = nullptr;
    size_t lights_cap = 0;
    uint32_t rays_per_lit_pixel = 0, shadow_bits = 0;
    VolLight* d_vol_lights = nullptr; // lights with a volumetric type, scene order
    size_t vol_lights_cap = 0;
    int n_vol_lights = 0;
    bool need_full_depth = false; // screen-space volumetrics read depth anywhere in the frame
    bool has_shadow_map_volumetric = false;
    // shadow maps: one per light that needs it, re-rendered by luzrt_shadow_map_pass
    std::vector<luzw_light_block> host_lights; // the lights of the current scene block
    std::vector<float*> shadow_data;           // per light (grow-only device buffers)
    std::vector<size_t> shadow_cap;            // floats
    std::vector<ShadowMapRec> host_shadow_recs;
    ShadowMapRec* d_shadow_recs = nullptr;
    size_t shadow_recs_cap = 0;
    bool shadow_maps_current = false; // rendered for the current scene block and TLAS

    std::vector<Blas> blas;
    BuildScratch scratch;
    BoxF* d_boxes = nullptr; // triangle / instance boxes (grow-only)
    size_t boxes_cap = 0;
    BlasAttr* d_blas_attr = nullptr;
    size_t blas_attr_cap = 0;
    bool blas_attr_dirty = true;

    WideBvh tlas;
    bool have_tlas = false;
    InstanceIn* d_inst_in = nullptr;
    uint32_t* d_inst_inv = nullptr; // leaf position of input instance i (inverse of tlas.prim_order)
    InstanceRec *d_recs_in = nullptr, *d_recs = nullptr;
    InstanceMeta *d_meta_in = nullptr, *d_meta = nullptr;
    float4* d_inst_boxes = nullptr; // 2 per instance, TLAS leaf order
    size_t inst_cap = 0;
    uint32_t n_inst = 0;
    std::vector<luzrt_blas> last_blas;
    std::vector<InstanceIn> host_inst;

    std::vector<Texture> textures;
    const uchar4** d_tex_data = nullptr;
    uint2* d_tex_size = nullptr;
    size_t tex_cap = 0;
    bool tex_dirty = true;
    luzw_model_block* d_models = nullptr;
    size_t models_cap = 0;

    uint32_t *d_shadow_mask = nullptr, *d_ao_mask = nullptr;
    uint32_t* d_hints = nullptr; // occluder hints of the shadow rays, [tile][light] (light_pass.cu)
    size_t hints_cap = 0;
    size_t shadow_mask_words = 0, ao_mask_words = 0; // per pixel
    size_t shadow_mask_cap = 0, ao_mask_cap = 0;
    DeviceStats* d_stats = nullptr;
    unsigned long long* d_lit = nullptr;
    float* d_pow22 = nullptr; // (c / 255)^2.2, c = 0..255 (light.frag:172), for the shading kernels
    uint2* d_ray_hints = nullptr; // per-ray temporal occluder hints of the shadow rays, [shadow bit][pixel][slots] (light_pass.cu)
    size_t ray_hints_cap = 0;     // in uint2 elements
    size_t ray_hints_refused = 0; // an allocation of this many elements failed: the pass runs without hints
    // the hints pay when last frame's occluders still occlude (static or slowly changing views) and cost a few per cent
    // when they do not (every instance moving every frame): the kernel counts how many rays they settle, the counters
    // come back asynchronously, and the pass uses the hints while the settled fraction stays above a threshold; while it
    // is off, two consecutive frames out of 32 run with hints (the first refreshes them, the second measures)
    unsigned long long* d_temporal_cnt = nullptr; // 64 x 16 u64
    unsigned long long* h_temporal_cnt = nullptr; // pinned copy
    cudaEvent_t ev_temporal = nullptr;
    bool temporal_pending = false; // a copy of the counters is in flight
    bool temporal_use = true;
    uint32_t temporal_frame = 0;
    float temporal_rate = -1.0f;   // last measured settled fraction
    bool temporal_last_on = false; // the last light pass ran the temporal-hint kernel
    uint32_t temporal_consecutive = 0;  // consecutive light passes that ran it
    bool temporal_pending_warm = false; // the counters in flight belong to a frame whose hints the frame before had written
    unsigned long long* d_hash = nullptr; // [0] content hash of this ctx's BVHs, [1 + r] the hash of rank r (luzrt_comm_check_bvh)

    cudaEvent_t ev[EV_COUNT][2]{};
    bool ev_valid[EV_COUNT]{};

    ncclComm_t comm = nullptr;
    // the frame gather runs on its own stream so that it overlaps the next frame's light pass; whoever touches
    // the buffer being gathered waits for ev_gathered first (wait_gather)
    cudaStream_t comm_stream = nullptr;
    cudaStream_t aux_stream = nullptr; // light pass: hint pass + shadow-ray launch (the AO-ray launch runs on `stream`)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaEvent_t ev_resolved = nullptr, ev_gathered = nullptr;
    const void* gather_buf = nullptr; // image with a gather in flight (nullptr: none)
    uint32_t min_node_lanes = 33; // > 32: one node visit per pass of the traversal loop (fastest on C2-C4; LUZRT_MIN_NODE_LANES)
};

namespace {

int fail(luzrt_ctx* c, int code, const char* fmt, ...) {
    if (c) {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof(buf), fmt, ap);
        va_end(ap);
        c->err = buf;
    }
    return code;
}

#define CU(c, expr)                                                                                         \
    do {                                                                                                    \
        cudaError_t _e = (expr);                                                                            \
        if (_e != cudaSuccess)                                                                              \
            return fail(c, _e == cudaErrorMemoryAllocation ? LUZRT_E_NOMEM : LUZRT_E_CUDA, "%s: %s", #expr, \
                        cudaGetErrorString(_e));                                                            \
    } while (0)

#define REQUIRE(c, cond, msg) \
    do {                      \
        if (!(cond)) return fail(c, LUZRT_E_INVALID, "%s", msg); \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

template <class T>
int grow(luzrt_ctx* c, T*& p, size_t& cap, size_t need) {
    if (cap >= need) return LUZRT_OK;
    if (p) {
        CU(c, cudaStreamSynchronize(c->stream));
        CU(c, cudaFree(p));
        p = nullptr;
        cap = 0;
    }
    const size_t n = need + need / 2;
    CU(c, cudaMalloc(&p, n * sizeof(T)));
    cap = n;
    return LUZRT_OK;
}

// stream-ordered upload of a small host array through the page-locked staging ring
int stage_upload(luzrt_ctx* c, void* dst, const void* src, size_t bytes) {
    if (bytes == 0) return LUZRT_OK;
    luzrt_ctx::Staging& s = c->staging[c->staging_next++ % 4];
    if (!s.done)
        CU(c, cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
    else
        CU(c, cudaEventSynchronize(s.done)); // four uploads ago: long finished
    if (s.cap < bytes) {
        if (s.host) cudaFreeHost(s.host);
        s.host = nullptr;
        s.cap = 0;
        if (cudaHostAlloc(&s.host, bytes + bytes / 2 + 256, cudaHostAllocDefault) != cudaSuccess)
            return fail(c, LUZRT_E_NOMEM, "cudaHostAlloc(%zu) failed", bytes);
        s.cap = bytes + bytes / 2 + 256;
    }
    memcpy(s.host, src, bytes);
    CU(c, cudaMemcpyAsync(dst, s.host, bytes, cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaEventRecord(s.done, c->stream));
    return LUZRT_OK;
}

// makes the ctx stream wait for the in-flight gather if it targets one of the given buffers (nullptr = any)
cudaError_t wait_gather(luzrt_ctx* c, const void* a = nullptr, const void* b = nullptr, const void* d = nullptr) {
    if (!c->gather_buf) return cudaSuccess;
    if (a && c->gather_buf != a && c->gather_buf != b && c->gather_buf != d) return cudaSuccess;
    c->gather_buf = nullptr;
    return cudaStreamWaitEvent(c->stream, c->ev_gathered, 0);
}

// Makes the ctx stream wait for the asynchronous read-back in flight if it reads the image about to be WRITTEN.
// (After SwapLightHistory the image being read back is lightHistory, which the next frame only reads: its light
// pass and TAA then overlap the download instead of queueing behind it.)
cudaError_t wait_download(luzrt_ctx* c, const void* written) {
    if (!c->download_pending || c->download_src != written) return cudaSuccess;
    return cudaStreamWaitEvent(c->stream, c->ev_downloaded, 0);
}

void ev_begin(luzrt_ctx* c, int which) { cudaEventRecord(c->ev[which][0], c->stream); }
void ev_end(luzrt_ctx* c, int which) {
    cudaEventRecord(c->ev[which][1], c->stream);
    c->ev_valid[which] = true;
}

void free_images(luzrt_ctx* c) {
    void* ptrs[] = {c->albedo, c->material, c->emission, c->compose, c->normal, c->lightA, c->lightB, c->lightHist,
                    c->depth, c->albedo_b, c->material_b, c->emission_b, c->normal_b, c->depth_b};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    c->albedo_b = c->material_b = c->emission_b = nullptr;
    c->normal_b = nullptr;
    c->depth_b = nullptr;
    c->upload_pending = c->download_pending = false;
    c->albedo = c->material = c->emission = c->compose = nullptr;
    c->normal = c->lightA = c->lightB = c->lightHist = nullptr;
    c->depth = nullptr;
}

void free_blas(Blas& b) {
    free_wide_bvh(b.bvh);
    if (b.tris) cudaFree(b.tris);
    if (b.verts) cudaFree(b.verts);
    if (b.indices) cudaFree(b.indices);
    b = Blas();
}

// rows this rank resolves / shades (own bands; own bands + one halo row each side for TAA's 3x3 taps)
BandSet own_bands(const luzrt_ctx* c) {
    if (c->world == 1) return BandSet{0, c->h, c->h, 1};
    return BandSet{(int)((uint32_t)c->rank * c->band_rows), (uint32_t)c->world * c->band_rows, c->band_rows, c->n_bands};
}
BandSet shade_bands(const luzrt_ctx* c) {
    if (c->world == 1) return BandSet{0, c->h, c->h, 1};
    return BandSet{(int)((uint32_t)c->rank * c->band_rows) - 1, (uint32_t)c->world * c->band_rows, c->band_rows + 2,
                   c->n_bands};
}
// smallest band height >= 48 rows that divides a rank's share (more bands = better balance, more halo rows)
uint32_t choose_band_rows(uint32_t rows_per_rank) {
    static const uint32_t min_rows = [] { // LUZRT_BAND_ROWS_MIN: tuning knob (luz_b200/strips.py reads it too)
        const char* e = getenv("LUZRT_BAND_ROWS_MIN");
        const int v = e ? atoi(e) : 48;
        return (uint32_t)(v >= 2 ? v : 2);
    }();
    uint32_t best = rows_per_rank;
    for (uint32_t k = 1; k <= rows_per_rank; k++)
        if (rows_per_rank % k == 0 && rows_per_rank / k >= min_rows) best = rows_per_rank / k;
    return best;
}
bool is_light_image(int which) { return which == LUZRT_IMG_LIGHT || which == LUZRT_IMG_HISTORY; }

void* image_ptr(luzrt_ctx* c, int which, size_t* bytes) {
    const size_t px = (size_t)c->w * c->h;
    switch (which) {
        case LUZRT_IMG_LIGHT: *bytes = px * 16; return c->lightA;
        case LUZRT_IMG_HISTORY: *bytes = px * 16; return c->lightHist;
        case LUZRT_SHADOW_MASK: *bytes = px * c->shadow_mask_words * 4; return c->d_shadow_mask;
        case LUZRT_AO_MASK: *bytes = px * c->ao_mask_words * 4; return c->d_ao_mask;
        case LUZRT_GBUF_ALBEDO: *bytes = px * 4; return c->albedo;
        case LUZRT_GBUF_NORMAL: *bytes = px * 16; return c->normal;
        case LUZRT_GBUF_MATERIAL: *bytes = px * 4; return c->material;
        case LUZRT_GBUF_EMISSION: *bytes = px * 4; return c->emission;
        case LUZRT_GBUF_DEPTH: *bytes = px * 4; return c->depth;
        case LUZRT_IMG_COMPOSE: *bytes = px * 4; return c->compose;
        default: *bytes = 0; return nullptr;
    }
}

} // namespace

extern "C" {

const char* luzrt_version(void) { return "luzrt 0.1 (sm_100a)"; }

int luzrt_create(int device_id, int rank, int world, luzrt_ctx** out) {
    if (!out) return LUZRT_E_INVALID;
    *out = nullptr;
    if (world < 1 || rank < 0 || rank >= world) return LUZRT_E_INVALID;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device_id < 0 || device_id >= n) return LUZRT_E_NODEVICE;
    cudaDeviceProp prop{};
    if (cudaGetDeviceProperties(&prop, device_id) != cudaSuccess) return LUZRT_E_NODEVICE;
    if (prop.major != 10) return LUZRT_E_NODEVICE; // the kernels are sm_100a SASS only
    if (cudaSetDevice(device_id) != cudaSuccess) return LUZRT_E_NODEVICE;
    luzrt_ctx* c = new luzrt_ctx();
    c->device = device_id;
    c->rank = rank;
    c->world = world;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete c;
        return LUZRT_E_CUDA;
    }
    for (int i = 0; i < EV_COUNT; i++)
        for (int k = 0; k < 2; k++) cudaEventCreate(&c->ev[i][k]);
    if (cudaMalloc(&c->d_stats, sizeof(DeviceStats)) != cudaSuccess ||
        cudaMalloc(&c->d_lit, 64 * 16 * sizeof(unsigned long long)) != cudaSuccess ||
        cudaMalloc(&c->d_pow22, 256 * sizeof(float)) != cudaSuccess) {
        luzrt_destroy(c);
        return LUZRT_E_NOMEM;
    }
    if (launch_pow22_table(c->stream, c->d_pow22) != cudaSuccess) {
        luzrt_destroy(c);
        return LUZRT_E_CUDA;
    }
    cudaMemsetAsync(c->d_stats, 0, sizeof(DeviceStats), c->stream);
    cudaMemsetAsync(c->d_lit, 0, 64 * 16 * sizeof(unsigned long long), c->stream);
    if (const char* e = getenv("LUZRT_MIN_NODE_LANES")) c->min_node_lanes = (uint32_t)atoi(e);
    c->blas.emplace_back(); // handle 0 is invalid
    *out = c;
    return LUZRT_OK;
}

void luzrt_destroy(luzrt_ctx* c) {
    if (!c) return;
    DeviceGuard g(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->comm_stream) cudaStreamSynchronize(c->comm_stream);
    if (c->aux_stream) {
        cudaStreamSynchronize(c->aux_stream);
        cudaStreamDestroy(c->aux_stream);
        cudaEventDestroy(c->ev_fork);
        cudaEventDestroy(c->ev_join);
    }
    if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
    for (cudaStream_t st : {c->upload_stream, c->download_stream})
        if (st) {
            cudaStreamSynchronize(st);
            cudaStreamDestroy(st);
        }
    for (auto& st : c->staging) {
        if (st.done) cudaEventDestroy(st.done);
        if (st.host) cudaFreeHost(st.host);
    }
    for (cudaEvent_t e : {c->ev_flip, c->ev_uploaded, c->ev_result, c->ev_downloaded})
        if (e) cudaEventDestroy(e);
    if (c->ev_resolved) cudaEventDestroy(c->ev_resolved);
    if (c->ev_gathered) cudaEventDestroy(c->ev_gathered);
    if (c->comm_stream) cudaStreamDestroy(c->comm_stream);
    free_images(c);
    for (auto& b : c->blas) free_blas(b);
    free_wide_bvh(c->tlas);
    for (auto& t : c->textures)
        if (t.data) cudaFree(t.data);
    if (c->unperm) cudaFree(c->unperm);
    void* ptrs[] = {c->blue_noise, c->d_lights, c->d_boxes,   c->d_blas_attr, c->d_inst_in, c->d_recs_in,
                    c->d_recs,     c->d_meta_in, c->d_meta,   c->d_tex_data,  c->d_tex_size, c->d_models, c->d_inst_boxes,
                    c->d_shadow_mask, c->d_ao_mask, c->d_stats, c->d_lit, c->d_vol_lights, c->d_shadow_recs, c->d_hints, c->d_pow22, c->d_hash, c->d_ray_hints, c->d_inst_inv, c->d_temporal_cnt};
    if (c->h_temporal_cnt) cudaFreeHost(c->h_temporal_cnt);
    if (c->ev_temporal) cudaEventDestroy(c->ev_temporal);
    for (float* p : c->shadow_data)
        if (p) cudaFree(p);
    for (void* p : ptrs)
        if (p) cudaFree(p);
    for (int i = 0; i < EV_COUNT; i++)
        for (int k = 0; k < 2; k++)
            if (c->ev[i][k]) cudaEventDestroy(c->ev[i][k]);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

const char* luzrt_last_error(luzrt_ctx* c) { return c ? c->err.c_str() : "null context"; }

int luzrt_comm_unique_id(void* out128) {
    if (!out128) return LUZRT_E_INVALID;
    if (!g_nccl.load()) return LUZRT_E_COMM;
    ncclUniqueId id;
    if (g_nccl.GetUniqueId(&id) != ncclSuccess) return LUZRT_E_COMM;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    memcpy(out128, &id, 128);
    return LUZRT_OK;
}

int luzrt_comm_init(luzrt_ctx* c, const void* id128) {
    if (!c) return LUZRT_E_INVALID;
    REQUIRE(c, id128, "id128 is null");
    if (c->world == 1) return LUZRT_OK;
    if (!g_nccl.load()) return fail(c, LUZRT_E_COMM, "libnccl.so.2 could not be loaded: %s", dlerror());
    DeviceGuard g(c->device);
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    ncclResult_t r = g_nccl.CommInitRank(&c->comm, c->world, id, c->rank);
    if (r != ncclSuccess) return fail(c, LUZRT_E_COMM, "ncclCommInitRank: %s", g_nccl.GetErrorString(r));
    CU(c, cudaStreamCreateWithFlags(&c->comm_stream, cudaStreamNonBlocking));
    CU(c, cudaEventCreateWithFlags(&c->ev_resolved, cudaEventDisableTiming));
    CU(c, cudaEventCreateWithFlags(&c->ev_gathered, cudaEventDisableTiming));
    return LUZRT_OK;
}

// One process, n GPUs (Luz is one process and one thread, main.cpp:356-366): one ctx per device with
// rank i of n, their communicators created together by ncclCommInitAll -- no unique id to broadcast.  Collectives of
// contexts that live in one thread must be issued as one NCCL group: luzrt_gather_multi / luzrt_comm_check_bvh_multi.
int luzrt_create_multi(const int* device_ids, int n_devices, luzrt_ctx** out) {
    if (!device_ids || !out || n_devices < 1 || n_devices > 64) return LUZRT_E_INVALID;
    for (int i = 0; i < n_devices; i++) out[i] = nullptr;
    auto destroy_all = [&] {
        for (int i = 0; i < n_devices; i++) {
            if (out[i]) luzrt_destroy(out[i]);
            out[i] = nullptr;
        }
    };
    for (int i = 0; i < n_devices; i++) {
        const int rc = luzrt_create(device_ids[i], i, n_devices, &out[i]);
        if (rc != LUZRT_OK) {
            destroy_all();
            return rc;
        }
    }
    if (n_devices == 1) return LUZRT_OK;
    if (!g_nccl.load()) {
        destroy_all();
        return LUZRT_E_COMM;
    }
    std::vector<ncclComm_t> comms((size_t)n_devices);
    if (g_nccl.CommInitAll(comms.data(), n_devices, device_ids) != ncclSuccess) {
        destroy_all();
        return LUZRT_E_COMM;
    }
    for (int i = 0; i < n_devices; i++) {
        luzrt_ctx* c = out[i];
        c->comm = comms[(size_t)i];
        DeviceGuard g(c->device);
        if (cudaStreamCreateWithFlags(&c->comm_stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&c->ev_resolved, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&c->ev_gathered, cudaEventDisableTiming) != cudaSuccess) {
            destroy_all();
            return LUZRT_E_CUDA;
        }
    }
    return LUZRT_OK;
}

// Content hash of everything the rays traverse on this ctx: TLAS nodes, leaf order, instance transforms (the device
// pointers in the records are skipped) and world boxes, and the nodes + triangles of every live BLAS.
static int bvh_hash_enqueue(luzrt_ctx* c) {
    REQUIRE(c, c->have_tlas, "no TLAS built");
    if (!c->d_hash) CU(c, cudaMalloc(&c->d_hash, sizeof(unsigned long long) * 65));
    unsigned long long* acc = c->d_hash;
    CU(c, cudaMemsetAsync(acc, 0, sizeof(unsigned long long), c->stream));
    CU(c, launch_hash_words(c->stream, c->tlas.nodes, (size_t)c->tlas.n_nodes * sizeof(WideNode), 0x11ull << 32, 4, 4, acc));
    CU(c, launch_hash_words(c->stream, c->tlas.prim_order, (size_t)c->n_inst * 4, 0x22ull << 32, 4, 4, acc));
    CU(c, launch_hash_words(c->stream, c->d_recs, (size_t)c->n_inst * sizeof(InstanceRec), 0x33ull << 32, sizeof(InstanceRec), 48, acc));
    CU(c, launch_hash_words(c->stream, c->d_inst_boxes, (size_t)c->n_inst * 32, 0x44ull << 32, 4, 4, acc));
    for (size_t i = 0; i < c->blas.size(); i++) {
        const Blas& b = c->blas[i];
        if (!b.alive) continue;
        CU(c, launch_hash_words(c->stream, b.bvh.nodes, (size_t)b.bvh.n_nodes * sizeof(WideNode), (0x1000ull + i) << 32, 4, 4, acc));
        CU(c, launch_hash_words(c->stream, b.tris, (size_t)b.n_tris * sizeof(WideTri), (0x800000ull + i) << 32, 4, 4, acc));
    }
    return LUZRT_OK;
}

int luzrt_bvh_hash(luzrt_ctx* c, uint64_t* out) {
    if (!c) return LUZRT_E_INVALID;
    REQUIRE(c, out, "out is null");
    DeviceGuard g(c->device);
    int rc = bvh_hash_enqueue(c);
    if (rc != LUZRT_OK) return rc;
    unsigned long long h = 0;
    CU(c, cudaMemcpyAsync(&h, c->d_hash, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    *out = h;
    return LUZRT_OK;
}

// "BVH build must be deterministic across GPUs": every rank hashes its replica, one all-gather of the 8-byte hashes, and
// every rank compares.  Collective: all ranks call it (contexts of one process through luzrt_comm_check_bvh_multi).
static int check_bvh_enqueue(luzrt_ctx* c) {
    if (!c->comm) return fail(c, LUZRT_E_STATE, "luzrt_comm_init has not been called");
    int rc = bvh_hash_enqueue(c);
    if (rc != LUZRT_OK) return rc;
    const ncclResult_t r = g_nccl.AllGather(c->d_hash, c->d_hash + 1, 1, ncclUint64, c->comm, c->stream);
    if (r != ncclSuccess) return fail(c, LUZRT_E_COMM, "ncclAllGather: %s", g_nccl.GetErrorString(r));
    return LUZRT_OK;
}
static int check_bvh_compare(luzrt_ctx* c) {
    unsigned long long h[65];
    REQUIRE(c, c->world <= 64, "world > 64");
    CU(c, cudaMemcpyAsync(h, c->d_hash, sizeof(unsigned long long) * (size_t)(1 + c->world), cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    for (int r = 0; r < c->world; r++)
        if (h[1 + r] != h[0])
            return fail(c, LUZRT_E_STATE, "acceleration structures differ: rank %d has hash %016llx, rank %d has %016llx", c->rank,
                        h[0], r, h[1 + r]);
    return LUZRT_OK;
}
int luzrt_comm_check_bvh(luzrt_ctx* c) {
    if (!c) return LUZRT_E_INVALID;
    if (c->world == 1) return LUZRT_OK;
    DeviceGuard g(c->device);
    const int rc = check_bvh_enqueue(c);
    return rc != LUZRT_OK ? rc : check_bvh_compare(c);
}
int luzrt_comm_check_bvh_multi(luzrt_ctx** ctxs, int n) {
    if (!ctxs || n < 1) return LUZRT_E_INVALID;
    if (n == 1) return luzrt_comm_check_bvh(ctxs[0]);
    if (!g_nccl.load()) return LUZRT_E_COMM;
    g_nccl.GroupStart();
    int rc = LUZRT_OK;
    for (int i = 0; i < n && rc == LUZRT_OK; i++) {
        DeviceGuard g(ctxs[i]->device);
        rc = check_bvh_enqueue(ctxs[i]);
    }
    const ncclResult_t r = g_nccl.GroupEnd();
    if (rc == LUZRT_OK && r != ncclSuccess) rc = fail(ctxs[0], LUZRT_E_COMM, "ncclGroupEnd: %s", g_nccl.GetErrorString(r));
    for (int i = 0; i < n && rc == LUZRT_OK; i++) {
        DeviceGuard g(ctxs[i]->device);
        rc = check_bvh_compare(ctxs[i]);
    }
    return rc;
}

int luzrt_gather_multi(luzrt_ctx** ctxs, int n) {
    if (!ctxs || n < 1) return LUZRT_E_INVALID;
    if (n == 1) return luzrt_gather(ctxs[0]);
    if (!g_nccl.load()) return LUZRT_E_COMM;
    g_nccl.GroupStart();
    int rc = LUZRT_OK;
    for (int i = 0; i < n && rc == LUZRT_OK; i++) rc = luzrt_gather(ctxs[i]);
    const ncclResult_t r = g_nccl.GroupEnd();
    if (rc == LUZRT_OK && r != ncclSuccess) rc = fail(ctxs[0], LUZRT_E_COMM, "ncclGroupEnd: %s", g_nccl.GetErrorString(r));
    return rc;
}

int luzrt_resize(luzrt_ctx* c, uint32_t width, uint32_t height) {
    if (!c) return LUZRT_E_INVALID;
    REQUIRE(c, width > 0 && height > 0 && width <= 65536 && height <= 65536, "bad image size");
    REQUIRE(c, height % (uint32_t)c->world == 0, "height must be divisible by the number of ranks");
    REQUIRE(c, c->world == 1 || height / (uint32_t)c->world >= 2, "a rank needs at least two rows");
    REQUIRE(c, (c->world & (c->world - 1)) == 0, "the number of ranks must be a power of two");
    DeviceGuard g(c->device);
    CU(c, cudaStreamSynchronize(c->stream));
    if (c->comm_stream) CU(c, cudaStreamSynchronize(c->comm_stream));
    if (c->upload_stream) CU(c, cudaStreamSynchronize(c->upload_stream));
    if (c->download_stream) CU(c, cudaStreamSynchronize(c->download_stream));
    c->gather_buf = nullptr;
    free_images(c);
    const size_t px = (size_t)width * height;
    CU(c, cudaMalloc(&c->albedo, px * 4));
    CU(c, cudaMalloc(&c->material, px * 4));
    CU(c, cudaMalloc(&c->emission, px * 4));
    CU(c, cudaMalloc(&c->compose, px * 4));
    CU(c, cudaMalloc(&c->normal, px * 16));
    CU(c, cudaMalloc(&c->lightA, px * 16));
    CU(c, cudaMalloc(&c->lightB, px * 16));
    CU(c, cudaMalloc(&c->lightHist, px * 16));
    CU(c, cudaMalloc(&c->depth, px * 4));
    // cleared like the attachments: colour 0, depth 1 (VulkanWrapper.cpp:1194-1196, :1212)
    CU(c, cudaMemsetAsync(c->albedo, 0, px * 4, c->stream));
    CU(c, cudaMemsetAsync(c->material, 0, px * 4, c->stream));
    CU(c, cudaMemsetAsync(c->emission, 0, px * 4, c->stream));
    CU(c, cudaMemsetAsync(c->compose, 0, px * 4, c->stream));
    CU(c, cudaMemsetAsync(c->normal, 0, px * 16, c->stream));
    CU(c, cudaMemsetAsync(c->lightA, 0, px * 16, c->stream));
    CU(c, cudaMemsetAsync(c->lightB, 0, px * 16, c->stream));
    CU(c, cudaMemsetAsync(c->lightHist, 0, px * 16, c->stream));
    {
        std::vector<float> ones(px, 1.0f);
        CU(c, cudaMemcpyAsync(c->depth, ones.data(), px * 4, cudaMemcpyHostToDevice, c->stream));
        CU(c, cudaStreamSynchronize(c->stream));
    }
    c->w = width;
    c->h = height;
    c->rows_per_rank = height / (uint32_t)c->world;
    c->band_rows = c->world == 1 ? height : choose_band_rows(c->rows_per_rank);
    c->n_bands = c->rows_per_rank / c->band_rows;
    c->history_valid = false;
    c->fc.width = width;
    c->fc.height = height;
    c->fc.band_rows = c->band_rows;
    c->fc.band_magic = (uint32_t)((0x100000000ull + c->band_rows - 1) / c->band_rows);
    c->fc.world_shift = 0;
    while ((1 << c->fc.world_shift) < c->world) c->fc.world_shift++;
    c->fc.rows_per_rank = c->rows_per_rank;
    return LUZRT_OK;
}

int luzrt_set_blue_noise(luzrt_ctx* c, const uint8_t* rgba8, uint32_t width, uint32_t height) {
    if (!c) return LUZRT_E_INVALID;
    REQUIRE(c, rgba8 && width > 0 && height > 0, "bad blue-noise texture");
    DeviceGuard g(c->device);
    CU(c, cudaStreamSynchronize(c->stream));
    if (c->blue_noise) cudaFree(c->blue_noise);
    c->blue_noise = nullptr;
    CU(c, cudaMalloc(&c->blue_noise, (size_t)width * height * 4));
    CU(c, cudaMemcpyAsync(c->blue_noise, rgba8, (size_t)width * height * 4, cudaMemcpyHostToDevice, c->stream));
    c->bn_w = width;
    c->bn_h = height;
    return LUZRT_OK;
}

int luzrt_texture_create(luzrt_ctx* c, const uint8_t* rgba8, uint32_t width, uint32_t height, int32_t* out_rid) {
    if (!c) return LUZRT_E_INVALID;
    REQUIRE(c, rgba8 && width > 0 && height > 0 && out_rid, "bad texture");
    DeviceGuard g(c->device);
    Texture t;
    CU(c, cudaMalloc(&t.data, (size_t)width * height * 4));
    CU(c, cudaMemcpyAsync(t.data, rgba8, (size_t)width * height * 4, cudaMemcpyHostToDevice, c->stream));
    t.size = make_uint2(width, height);
    c->textures.push_back(t);
    c->tex_dirty = true;
    *out_rid = (int32_t)c->textures.size() - 1;
    return LUZRT_OK;
}

int luzrt_blas_create(luzrt_ctx* c, const void* vertices, uint32_t vertex_count, uint32_t vertex_stride,
                      const uint32_t* indices, uint32_t index_count, luzrt_blas* out) {
    if (!c) return LUZRT_E_INVALID;
    REQUIRE(c, out, "out is null");
    *out = 0;
    REQUIRE(c, vertex_stride >= 12 && vertex_stride % 4 == 0, "vertex_stride must be a multiple of 4 and >= 12");
    const uint32_t n_tris = index_count / 3; // primitiveCount = indexCount / 3 (VulkanWrapper.cpp:761)
    REQUIRE(c, (vertices && indices) || n_tris == 0, "null vertex/index data");
    for (uint32_t i = 0; i < 3 * n_tris; i++)
        if (indices[i] >= vertex_count) return fail(c, LUZRT_E_INVALID, "index %u out of range at %u", indices[i], i);
    DeviceGuard g(c->device);
    Blas b;
    b.alive = true;
    b.stride = vertex_stride;
    b.n_verts = vertex_count;
    b.n_tris = n_tris;
    int rc = LUZRT_OK;
    auto bail = [&](int code) {
        free_blas(b);
        return code;
    };
#define CUB_(expr)                                                                                   \
    do {                                                                                             \
        cudaError_t _e = (expr);                                                                     \
        if (_e != cudaSuccess)                                                                       \
            return bail(fail(c, _e == cudaErrorMemoryAllocation ? LUZRT_E_NOMEM : LUZRT_E_CUDA,     \
                             "%s: %s", #expr, cudaGetErrorString(_e)));                              \
    } while (0)
    CUB_(cudaMalloc(&b.verts, std::max<size_t>((size_t)vertex_count * vertex_stride, 16)));
    CUB_(cudaMalloc(&b.indices, std::max<size_t>((size_t)n_tris * 12, 16)));
    if (vertex_count)
        CUB_(cudaMemcpyAsync(b.verts, vertices, (size_t)vertex_count * vertex_stride, cudaMemcpyHostToDevice, c->stream));
    if (n_tris) CUB_(cudaMemcpyAsync(b.indices, indices, (size_t)n_tris * 12, cudaMemcpyHostToDevice, c->stream));
    if ((rc = grow(c, c->d_boxes, c->boxes_cap, std::max<size_t>(n_tris, 1))) != LUZRT_OK) return bail(rc);
    CUB_(launch_triangle_boxes(c->stream, b.verts, b.stride, b.indices, n_tris, c->d_boxes));
    c->launches += n_tris ? 1 : 0;
    CUB_(build_wide_bvh(c->stream, c->scratch, c->d_boxes, n_tris, 3, b.bvh, &c->launches));
    CUB_(cudaMalloc(&b.tris, std::max<size_t>((size_t)n_tris, 1) * sizeof(WideTri)));
    CUB_(launch_gather_triangles(c->stream, b.verts, b.stride, b.indices, b.bvh.prim_order, n_tris, b.tris));
    c->launches += n_tris ? 1 : 0;
    CUB_(cudaMemcpyAsync(&b.bounds, b.bvh.node_bounds, sizeof(BoxF), cudaMemcpyDeviceToHost, c->stream));
    CUB_(cudaStreamSynchronize(c->stream));
    // shrink the node array to its final size (the builder allocates for the worst case)
    if (b.bvh.node_capacity > (size_t)b.bvh.n_nodes + 16) {
        WideNode* nn = nullptr;
        BoxF* nb = nullptr;
        CUB_(cudaMalloc(&nn, (size_t)b.bvh.n_nodes * sizeof(WideNode)));
        CUB_(cudaMalloc(&nb, (size_t)b.bvh.n_nodes * sizeof(BoxF)));
        CUB_(cudaMemcpyAsync(nn, b.bvh.nodes, (size_t)b.bvh.n_nodes * sizeof(WideNode), cudaMemcpyDeviceToDevice, c->stream));
        CUB_(cudaMemcpyAsync(nb, b.bvh.node_bounds, (size_t)b.bvh.n_nodes * sizeof(BoxF), cudaMemcpyDeviceToDevice, c->stream));
        CUB_(cudaStreamSynchronize(c->stream));
        cudaFree(b.bvh.nodes);
        cudaFree(b.bvh.node_bounds);
        b.bvh.nodes = nn;
        b.bvh.node_bounds = nb;
        b.bvh.node_capacity = b.bvh.n_nodes;
    }
#undef CUB_
    b.levels = (uint32_t)b.bvh.levels.size();
    // reuse a dead slot if there is one
    uint32_t slot = 0;
    for (uint32_t i = 1; i < c->blas.size(); i++)
        if (!c->blas[i].alive) {
            slot = i;
            break;
        }
    if (!slot) {
        c->blas.emplace_back();
        slot = (uint32_t)c->blas.size() - 1;
    }
    c->blas[slot] = b;
    c->blas_attr_dirty = true;
    *out = slot;
    return LUZRT_OK;
}

int luzrt_blas_destroy(luzrt_ctx* c, luzrt_blas h) {
    if (!c) return LUZRT_E_INVALID;
    REQUIRE(c, h > 0 && h < c->blas.size() && c->blas[h].alive, "invalid BLAS handle");
    DeviceGuard g(c->device);
    CU(c, cudaStreamSynchronize(c->stream));
    free_blas(c->blas[h]);
    c->blas_attr_dirty = true;
    // the current TLAS's instance records hold raw device pointers into the BLAS that was just freed: passes must not
    // traverse it, and a later "refit" against a BLAS that reuses the slot must be a rebuild
    for (luzrt_blas used : c->last_blas)
        if (used == h) {
            c->have_tlas = false;
            c->shadow_maps_current = false;
            c->last_blas.clear();
            break;
        }
    return LUZRT_OK;
}

static int dump_bvh(luzrt_ctx* c, const WideBvh& bvh, const void* prims, size_t prim_bytes, void* dst, size_t* bytes) {
    REQUIRE(c, bytes, "bytes is null");
    const size_t need = (size_t)bvh.n_nodes * sizeof(WideNode) + prim_bytes;
    if (!dst || *bytes < need) {
        *bytes = need;
        return dst ? fail(c, LUZRT_E_INVALID, "buffer too small") : LUZRT_OK;
    }
    CU(c, cudaStreamSynchronize(c->stream));
    CU(c, cudaMemcpy(dst, bvh.nodes, (size_t)bvh.n_nodes * sizeof(WideNode), cudaMemcpyDeviceToHost));
    if (prim_bytes)
        CU(c, cudaMemcpy((char*)dst + (size_t)bvh.n_nodes * sizeof(WideNode), prims, prim_bytes, cudaMemcpyDeviceToHost));
    *bytes = need;
    return LUZRT_OK;
}

int luzrt_blas_dump(luzrt_ctx* c, luzrt_blas h, void* dst, size_t* bytes) {
    if (!c) return LUZRT_E_INVALID;
    REQUIRE(c, h > 0 && h < c->blas.size() && c->blas[h].alive, "invalid BLAS handle");
    DeviceGuard g(c->device);
    const Blas& b = c->blas[h];
    return dump_bvh(c, b.bvh, b.tris, (size_t)b.n_tris * sizeof(WideTri), dst, bytes);
}

int luzrt_tlas_dump(luzrt_ctx* c, void* dst, size_t* bytes) {
    if (!c) return LUZRT_E_INVALID;
    REQUIRE(c, c->have_tlas, "no TLAS built");
    DeviceGuard g(c->device);
    // instance records hold device pointers (not reproducible across runs): dump nodes + leaf order
    return dump_bvh(c, c->tlas, c->tlas.prim_order, (size_t)c->n_inst * sizeof(uint32_t), dst, bytes);
}

int luzrt_tlas_build(luzrt_ctx* c, const luzrt_instance* instances, uint32_t count, int mode) {
    if (!c) return LUZRT_E_INVALID;
    REQUIRE(c, instances || count == 0, "instances is null");
    REQUIRE(c, mode == 0 || mode == 1, "mode must be 0 (rebuild) or 1 (refit)");
    DeviceGuard g(c->device);
    uint32_t max_blas_levels = 1;
    c->host_inst.resize(count);
    bool same_set = c->have_tlas && c->last_blas.size() == count;
    for (uint32_t i = 0; i < count; i++) {
        const luzrt_blas h = instances[i].blas;
        if (!(h > 0 && h < c->blas.size() && c->blas[h].alive))
            return fail(c, LUZRT_E_INVALID, "instance %u has an invalid BLAS handle", i);
        const Blas& b = c->blas[h];
        InstanceIn& in = c->host_inst[i];
        memcpy(in.m, instances[i].model_mat, sizeof(in.m));
        in.nodes = b.bvh.nodes;
        in.tris = b.tris;
        in.blas_bounds = b.bounds;
        in.custom_index = instances[i].custom_index;
        in.blas_slot = h;
        in.n_tris = b.n_tris;
        in.pad = 0;
        max_blas_levels = std::max(max_blas_levels, b.levels);
        if (same_set && c->last_blas[i] != h) same_set = false;
    }
    const bool refit = mode == 1 && same_set && count > 0;
    const size_t cap_need = std::max<uint32_t>(count, 1);
    if (c->inst_cap < cap_need) {
        CU(c, cudaStreamSynchronize(c->stream));
        void* olds[] = {c->d_inst_in, c->d_recs_in, c->d_recs, c->d_meta_in, c->d_meta, c->d_inst_boxes, c->d_inst_inv};
        for (void* p : olds)
            if (p) cudaFree(p);
        c->d_inst_in = nullptr;
        c->d_recs_in = c->d_recs = nullptr;
        c->d_meta_in = c->d_meta = nullptr;
        c->d_inst_boxes = nullptr;
        c->d_inst_inv = nullptr;
        c->inst_cap = 0;
        const size_t n = cap_need + cap_need / 2;
        CU(c, cudaMalloc(&c->d_inst_in, n * sizeof(InstanceIn)));
        CU(c, cudaMalloc(&c->d_recs_in, n * sizeof(InstanceRec)));
        CU(c, cudaMalloc(&c->d_recs, n * sizeof(InstanceRec)));
        CU(c, cudaMalloc(&c->d_meta_in, n * sizeof(InstanceMeta)));
        CU(c, cudaMalloc(&c->d_meta, n * sizeof(InstanceMeta)));
        CU(c, cudaMalloc(&c->d_inst_boxes, n * 2 * sizeof(float4)));
        CU(c, cudaMalloc(&c->d_inst_inv, n * sizeof(uint32_t)));
        c->inst_cap = n;
    }
    int rc;
    if ((rc = grow(c, c->d_boxes, c->boxes_cap, cap_need)) != LUZRT_OK) return rc;
    ev_begin(c, EV_TLAS);
    if (count)
        CU(c, cudaMemcpyAsync(c->d_inst_in, c->host_inst.data(), (size_t)count * sizeof(InstanceIn),
                              cudaMemcpyHostToDevice, c->stream));
    CU(c, launch_instance_prepare(c->stream, c->d_inst_in, count, c->d_boxes, c->d_recs_in, c->d_meta_in));
    c->launches += count ? 1 : 0;
    if (refit) {
        CU(c, refit_wide_bvh(c->stream, c->d_boxes, c->tlas, &c->launches));
    } else {
        CU(c, build_wide_bvh(c->stream, c->scratch, c->d_boxes, count, 1, c->tlas, &c->launches));
    }
    CU(c, launch_instance_gather(c->stream, c->d_inst_in, c->d_recs_in, c->d_meta_in, c->d_boxes, c->tlas.prim_order, count, c->d_recs,
                                 c->d_meta, c->d_inst_boxes, c->d_inst_inv));
    c->launches += count ? 1 : 0;
    ev_end(c, EV_TLAS);
    c->n_inst = count;
    c->have_tlas = true;
    c->shadow_maps_current = false; // the maps were rendered from the previous TLAS
    c->last_blas.resize(count);
    for (uint32_t i = 0; i < count; i++) c->last_blas[i] = instances[i].blas;
    if (c->tlas.levels.size() + max_blas_levels + 3 > LUZ_STACK_SIZE) {
        c->have_tlas = false; // no pass may traverse it: the stack would overflow
        c->last_blas.clear();
        return fail(c, LUZRT_E_INVALID, "BVH too deep for the traversal stack (%zu + %u levels)", c->tlas.levels.size(),
                    max_blas_levels);
    }
    return LUZRT_OK;
}

int luzrt_set_scene(luzrt_ctx* c, const luzw_scene_block* s, const luzw_light_block* extra, uint32_t n_extra) {
    if (!c) return LUZRT_E_INVALID;
    REQUIRE(c, s, "scene_block is null");
    REQUIRE(c, s->num_lights >= 0 && s->num_lights <= LUZW_MAX_LIGHTS, "numLights out of range");
    REQUIRE(c, extra || n_extra == 0, "extra_lights is null");
    REQUIRE(c, n_extra == 0 || s->num_lights == LUZW_MAX_LIGHTS, "extra lights require numLights == LUZ_MAX_LIGHTS");
    REQUIRE(c, s->ao_num_samples >= 0, "aoNumSamples must be >= 0");
    DeviceGuard g(c->device);
    const int n = s->num_lights + (int)n_extra;
    std::vector<LightRec> recs((size_t)std::max(n, 1));
    std::vector<VolLight> vols;
    c->host_lights.resize((size_t)n);
    c->shadow_maps_current = false; // the reference re-renders every map every frame (main.cpp:260-264)
    uint32_t shadow_bits = 0;
    for (int i = 0; i < n; i++) {
        const luzw_light_block& l = i < LUZW_MAX_LIGHTS ? s->lights[i] : extra[i - LUZW_MAX_LIGHTS];
        c->host_lights[i] = l;
        LightRec& r = recs[i];
        r.color_intensity = make_float4(l.color[0], l.color[1], l.color[2], l.intensity);
        r.position_inner = make_float4(l.position[0], l.position[1], l.position[2], l.inner_angle);
        r.direction_outer = make_float4(l.direction[0], l.direction[1], l.direction[2], l.outer_angle);
        r.type = l.type;
        r.num_shadow_samples = l.num_shadow_samples;
        r.radius = l.radius;
        r.shadow_map = l.shadow_map;
        if (s->shadow_type == LUZW_SHADOW_RAYTRACING && l.num_shadow_samples > 0) shadow_bits += (uint32_t)l.num_shadow_samples;
        if (l.volumetric_type == LUZW_VOLUMETRIC_SCREEN_SPACE || l.volumetric_type == LUZW_VOLUMETRIC_SHADOW_MAP) {
            VolLight v{};
            v.color_intensity = r.color_intensity;
            v.position_type = make_float4(l.position[0], l.position[1], l.position[2], 0.0f);
            memcpy(&v.position_type.w, &l.type, 4);
            v.direction_absorption = make_float4(l.direction[0], l.direction[1], l.direction[2], l.volumetric_absorption);
            v.samples = l.volumetric_samples;
            v.volumetric_type = l.volumetric_type;
            v.light_index = i;
            vols.push_back(v);
        }
    }
    int rc;
    c->n_vol_lights = (int)vols.size();
    c->need_full_depth = false;
    c->has_shadow_map_volumetric = false;
    for (const VolLight& v : vols) {
        c->need_full_depth = true; // both passes tap depth outside the rows a rank shades
        if (v.volumetric_type == LUZW_VOLUMETRIC_SHADOW_MAP) c->has_shadow_map_volumetric = true;
    }
    if (!vols.empty()) {
        if ((rc = grow(c, c->d_vol_lights, c->vol_lights_cap, vols.size())) != LUZRT_OK) return rc;
        if ((rc = stage_upload(c, c->d_vol_lights, vols.data(), vols.size() * sizeof(VolLight))) != LUZRT_OK) return rc;
    }
    if ((rc = grow(c, c->d_lights, c->lights_cap, recs.size())) != LUZRT_OK) return rc;
    if ((rc = stage_upload(c, c->d_lights, recs.data(), recs.size() * sizeof(LightRec))) != LUZRT_OK) return rc;
    FrameConst& fc = c->fc;
    memcpy(fc.inverse_proj, s->inverse_proj, 64);
    memcpy(fc.inverse_view, s->inverse_view, 64);
    memcpy(fc.view_proj, s->view_proj, 64);
    memcpy(fc.prev_view_proj, s->prev_view_proj, 64);
    memcpy(fc.jitter, s->jitter, 8);
    memcpy(fc.prev_jitter, s->prev_jitter, 8);
    memcpy(fc.cam_pos, s->cam_pos, 12);
    for (int k = 0; k < 3; k++) fc.ambient[k] = s->ambient_light_color[k] * s->ambient_light_intensity;
    fc.ao_min = s->ao_min;
    fc.ao_max = s->ao_max;
    fc.ao_num_samples = s->ao_num_samples;
    fc.num_lights = n;
    fc.shadow_type = s->shadow_type;
    c->shadow_bits = shadow_bits;
    c->rays_per_lit_pixel = shadow_bits + (uint32_t)s->ao_num_samples;
    c->have_scene = true;
    return LUZRT_OK;
}

namespace {
struct GbufPlanes {
    void* p[5]; // albedo, normal, material, emission, depth
};
// Copies the rows this ctx shades (everything on one GPU, else each own band + one wrapped row on each side) of
// the given full-frame planes; null sources are skipped.
int copy_gbuffer(luzrt_ctx* c, const GbufPlanes& dst, const void* const src[5], cudaMemcpyKind kind, cudaStream_t stream) {
    std::vector<std::pair<uint32_t, uint32_t>> seg;
    const BandSet sb = shade_bands(c);
    for (uint32_t k = 0; k < sb.n_bands; k++) {
        const int lo = sb.first + (int)(k * sb.pitch), hi = lo + (int)sb.rows;
        if (lo < 0) seg.emplace_back(c->h - 1, c->h);
        if (hi > (int)c->h) seg.emplace_back(0, 1);
        seg.emplace_back((uint32_t)std::max(lo, 0), (uint32_t)std::min(hi, (int)c->h));
    }
    static const size_t bpp[5] = {4, 16, 4, 4, 4};
    for (int pl = 0; pl < 5; pl++) {
        if (!src[pl]) continue;
        const size_t row = (size_t)c->w * bpp[pl];
        if (pl == 4 && c->need_full_depth && c->world > 1) { // the volumetric march reads depth anywhere in the frame
            CU(c, cudaMemcpyAsync(dst.p[pl], src[pl], (size_t)c->h * row, kind, stream));
            continue;
        }
        // bands that lie fully inside the image repeat at a fixed pitch: one strided copy covers them all
        size_t first_regular = seg.size(), n_regular = 0;
        for (size_t s2 = 0; s2 < seg.size(); s2++) {
            const bool regular = seg[s2].second - seg[s2].first == sb.rows;
            if (regular && n_regular == 0) first_regular = s2;
            if (regular && s2 == first_regular + n_regular &&
                (n_regular == 0 || seg[s2].first == seg[first_regular].first + n_regular * sb.pitch))
                n_regular++;
        }
        for (size_t s2 = 0; s2 < seg.size(); s2++) {
            if (n_regular > 1 && s2 >= first_regular && s2 < first_regular + n_regular) {
                if (s2 == first_regular) {
                    const size_t off = (size_t)seg[s2].first * row;
                    CU(c, cudaMemcpy2DAsync((char*)dst.p[pl] + off, (size_t)sb.pitch * row, (const char*)src[pl] + off,
                                            (size_t)sb.pitch * row, (size_t)sb.rows * row, n_regular, kind, stream));
                }
                continue;
            }
            const size_t off = (size_t)seg[s2].first * row, n = (size_t)(seg[s2].second - seg[s2].first) * row;
            CU(c, cudaMemcpyAsync((char*)dst.p[pl] + off, (const char*)src[pl] + off, n, kind, stream));
        }
    }
    return LUZRT_OK;
}
} // namespace

int luzrt_set_gbuffer(luzrt_ctx* c, const void* albedo, const void* normal, const void* material,
                      const void* emission, const void* depth, int src_is_device) {
    if (!c) return LUZRT_E_INVALID;
    if (!c->w) return fail(c, LUZRT_E_STATE, "luzrt_resize has not been called");
    DeviceGuard g(c->device);
    const void* src[5] = {albedo, normal, material, emission, depth};
    return copy_gbuffer(c, GbufPlanes{{c->albedo, c->normal, c->material, c->emission, c->depth}}, src,
                        src_is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, c->stream);
}

int luzrt_prefetch_gbuffer(luzrt_ctx* c, const void* albedo, const void* normal, const void* material,
                           const void* emission, const void* depth) {
    if (!c) return LUZRT_E_INVALID;
    if (!c->w) return fail(c, LUZRT_E_STATE, "luzrt_resize has not been called");
    REQUIRE(c, albedo && normal && material && emission && depth, "all five planes are required");
    DeviceGuard g(c->device);
    if (!c->upload_stream) {
        CU(c, cudaStreamCreateWithFlags(&c->upload_stream, cudaStreamNonBlocking));
        CU(c, cudaEventCreateWithFlags(&c->ev_flip, cudaEventDisableTiming));
        CU(c, cudaEventCreateWithFlags(&c->ev_uploaded, cudaEventDisableTiming));
        CU(c, cudaEventRecord(c->ev_flip, c->stream));
    }
    if (!c->albedo_b) {
        const size_t px = (size_t)c->w * c->h;
        CU(c, cudaMalloc(&c->albedo_b, px * 4));
        CU(c, cudaMalloc(&c->material_b, px * 4));
        CU(c, cudaMalloc(&c->emission_b, px * 4));
        CU(c, cudaMalloc(&c->normal_b, px * 16));
        CU(c, cudaMalloc(&c->depth_b, px * 4));
    }
    // the back set was the front set until the last flip: its readers were all enqueued before ev_flip
    CU(c, cudaStreamWaitEvent(c->upload_stream, c->ev_flip, 0));
    const void* src[5] = {albedo, normal, material, emission, depth};
    int rc = copy_gbuffer(c, GbufPlanes{{c->albedo_b, c->normal_b, c->material_b, c->emission_b, c->depth_b}}, src,
                          cudaMemcpyHostToDevice, c->upload_stream);
    if (rc != LUZRT_OK) return rc;
    CU(c, cudaEventRecord(c->ev_uploaded, c->upload_stream));
    c->upload_pending = true;
    return LUZRT_OK;
}

int luzrt_flip_gbuffer(luzrt_ctx* c) {
    if (!c) return LUZRT_E_INVALID;
    if (!c->upload_pending) return fail(c, LUZRT_E_STATE, "luzrt_prefetch_gbuffer has not been called");
    DeviceGuard g(c->device);
    CU(c, cudaEventRecord(c->ev_flip, c->stream)); // everything enqueued so far read the old front set
    CU(c, cudaStreamWaitEvent(c->stream, c->ev_uploaded, 0));
    std::swap(c->albedo, c->albedo_b);
    std::swap(c->normal, c->normal_b);
    std::swap(c->material, c->material_b);
    std::swap(c->emission, c->emission_b);
    std::swap(c->depth, c->depth_b);
    c->upload_pending = false;
    return LUZRT_OK;
}

int luzrt_read_owned_async(luzrt_ctx* c, int which, void* dst, size_t bytes) {
    if (!c) return LUZRT_E_INVALID;
    REQUIRE(c, dst, "dst is null");
    REQUIRE(c, is_light_image(which), "only the light images can be read asynchronously");
    if (!c->w) return fail(c, LUZRT_E_STATE, "luzrt_resize has not been called");
    DeviceGuard g(c->device);
    if (!c->download_stream) {
        CU(c, cudaStreamCreateWithFlags(&c->download_stream, cudaStreamNonBlocking));
        CU(c, cudaEventCreateWithFlags(&c->ev_result, cudaEventDisableTiming));
        CU(c, cudaEventCreateWithFlags(&c->ev_downloaded, cudaEventDisableTiming));
    }
    size_t have = 0;
    void* src = image_ptr(c, which, &have);
    const size_t row = have / c->h, need = row * c->rows_per_rank;
    REQUIRE(c, bytes >= need, "destination buffer too small");
    CU(c, cudaEventRecord(c->ev_result, c->stream));
    CU(c, cudaStreamWaitEvent(c->download_stream, c->ev_result, 0));
    // a rank's own rows are contiguous in the banded storage order and are not written by a gather in flight
    CU(c, cudaMemcpyAsync(dst, (const char*)src + row * c->rows_per_rank * (size_t)c->rank, need, cudaMemcpyDeviceToHost,
                          c->download_stream));
    CU(c, cudaEventRecord(c->ev_downloaded, c->download_stream));
    c->download_pending = true;
    c->download_src = src;
    return LUZRT_OK;
}

int luzrt_read_wait(luzrt_ctx* c) {
    if (!c) return LUZRT_E_INVALID;
    if (!c->download_pending) return LUZRT_OK;
    DeviceGuard g(c->device);
    CU(c, cudaEventSynchronize(c->ev_downloaded));
    c->download_pending = false;
    return LUZRT_OK;
}

namespace {
double det3_rows(const float* m, int r0, int r1, int r2);
}

int luzrt_gbuffer_pass(luzrt_ctx* c, const luzw_model_block* models, uint32_t n_models) {
    if (!c) return LUZRT_E_INVALID;
    if (!c->w) return fail(c, LUZRT_E_STATE, "luzrt_resize has not been called");
    if (!c->have_scene) return fail(c, LUZRT_E_STATE, "luzrt_set_scene has not been called");
    if (!c->have_tlas) return fail(c, LUZRT_E_STATE, "luzrt_tlas_build has not been called");
    REQUIRE(c, models || n_models == 0, "models is null");
    DeviceGuard g(c->device);
    int rc;
    if ((rc = grow(c, c->d_models, c->models_cap, std::max<size_t>(n_models, 1))) != LUZRT_OK) return rc;
    if (n_models)
        CU(c, cudaMemcpyAsync(c->d_models, models, (size_t)n_models * sizeof(luzw_model_block), cudaMemcpyHostToDevice,
                              c->stream));
    if (c->blas_attr_dirty) {
        std::vector<BlasAttr> attr(c->blas.size());
        for (size_t i = 0; i < c->blas.size(); i++) {
            const Blas& b = c->blas[i];
            attr[i] = BlasAttr{b.verts, b.indices, b.stride, (b.alive && b.stride == 48) ? 1u : 0u};
        }
        if ((rc = grow(c, c->d_blas_attr, c->blas_attr_cap, attr.size())) != LUZRT_OK) return rc;
        CU(c, cudaMemcpyAsync(c->d_blas_attr, attr.data(), attr.size() * sizeof(BlasAttr), cudaMemcpyHostToDevice, c->stream));
        c->blas_attr_dirty = false;
    }
    if (c->tex_dirty) {
        const size_t n = std::max<size_t>(c->textures.size(), 1);
        std::vector<const uchar4*> ptrs(n, nullptr);
        std::vector<uint2> sizes(n, make_uint2(0, 0));
        for (size_t i = 0; i < c->textures.size(); i++) {
            ptrs[i] = c->textures[i].data;
            sizes[i] = c->textures[i].size;
        }
        if (c->tex_cap < n) {
            CU(c, cudaStreamSynchronize(c->stream));
            if (c->d_tex_data) cudaFree(c->d_tex_data);
            if (c->d_tex_size) cudaFree(c->d_tex_size);
            c->d_tex_data = nullptr;
            c->d_tex_size = nullptr;
            CU(c, cudaMalloc(&c->d_tex_data, n * 2 * sizeof(void*)));
            CU(c, cudaMalloc(&c->d_tex_size, n * 2 * sizeof(uint2)));
            c->tex_cap = n * 2;
        }
        CU(c, cudaMemcpyAsync(c->d_tex_data, ptrs.data(), n * sizeof(void*), cudaMemcpyHostToDevice, c->stream));
        CU(c, cudaMemcpyAsync(c->d_tex_size, sizes.data(), n * sizeof(uint2), cudaMemcpyHostToDevice, c->stream));
        c->tex_dirty = false;
    }
    GbufferArgs a{};
    a.fc = c->fc;
    a.scene = TraceScene{c->tlas.nodes, c->d_recs, c->d_inst_boxes, c->min_node_lanes};
    a.inst_meta = c->d_meta;
    a.blas_attr = c->d_blas_attr;
    a.models = c->d_models;
    a.n_models = n_models;
    a.tex_data = c->d_tex_data;
    a.tex_size = c->d_tex_size;
    a.n_textures = (uint32_t)c->textures.size();
    a.albedo = c->albedo;
    a.normal = c->normal;
    a.material = c->material;
    a.emission = c->emission;
    a.depth = c->depth;
    { // the Opaque Pipeline culls back faces (front = counter-clockwise in the framebuffer; DeferredRenderer.cpp
      // "Opaque Pipeline" leaves cullFront false, VulkanWrapper.cpp:941-946): with the sign s_view of the view's linear
      // part (rows x, y, w of a perspective viewProj; x, y, z of an orthographic one) a triangle is front-facing iff
      // s_view * sign(det M_instance) * dot(d, (v1 - v0) x (v2 - v0)) < 0 along the primary ray d (shadow_map.cu has
      // the derivation); trace_ray<FACE_CULL> keeps the triangles with cull_sign * sign(det M) * dot(d, n) > 0
        const float* m = c->fc.view_proj;
        const bool ortho = m[3] == 0.0f && m[7] == 0.0f && m[11] == 0.0f;
        const double d = det3_rows(m, 0, 1, ortho ? 2 : 3);
        if (d == 0.0 || d != d) return fail(c, LUZRT_E_INVALID, "scene.viewProj is singular");
        a.cull_sign = d < 0.0 ? 1.0f : -1.0f;
    }
    // screen-space volumetrics read depth anywhere in the frame: every rank then renders all rows
    a.rows = (c->need_full_depth && c->world > 1) ? BandSet{0, c->h, c->h, 1} : shade_bands(c);
    ev_begin(c, EV_GBUF);
    CU(c, launch_gbuffer_pass(c->stream, a));
    c->launches++;
    ev_end(c, EV_GBUF);
    return LUZRT_OK;
}

int luzrt_set_debug(luzrt_ctx* c, uint32_t flags) {
    if (!c) return LUZRT_E_INVALID;
    c->debug = flags;
    return LUZRT_OK;
}

int luzrt_light_pass(luzrt_ctx* c, uint32_t frame) {
    if (!c) return LUZRT_E_INVALID;
    if (!c->w) return fail(c, LUZRT_E_STATE, "luzrt_resize has not been called");
    if (!c->have_scene) return fail(c, LUZRT_E_STATE, "luzrt_set_scene has not been called");
    if (!c->have_tlas) return fail(c, LUZRT_E_STATE, "luzrt_tlas_build has not been called");
    if (!c->blue_noise) return fail(c, LUZRT_E_STATE, "luzrt_set_blue_noise has not been called");
    DeviceGuard g(c->device);
    const bool stats = (c->debug & LUZRT_DEBUG_STATS) != 0;
    if (c->fc.shadow_type == LUZW_SHADOW_MAP && !c->shadow_maps_current)
        return fail(c, LUZRT_E_STATE, "shadowType 2: luzrt_shadow_map_pass has not been called for this scene block");
    const size_t px = (size_t)c->w * c->h;
    { // the visibility masks carry the rays' results from the ray kernel to the shading kernel
        const size_t sw = std::max<size_t>((c->shadow_bits + 31) / 32, 1);
        const size_t aw = std::max<size_t>(((size_t)c->fc.ao_num_samples + 31) / 32, 1);
        int rc;
        if ((rc = grow(c, c->d_shadow_mask, c->shadow_mask_cap, px * sw)) != LUZRT_OK) return rc;
        if ((rc = grow(c, c->d_ao_mask, c->ao_mask_cap, px * aw)) != LUZRT_OK) return rc;
        c->shadow_mask_words = sw;
        c->ao_mask_words = aw;
    }
    LightArgs a{};
    a.fc = c->fc;
    a.fc.frame_mod = (int)((int32_t)frame % 128); // ctx.frame % 128 on a GLSL int (frameCount < 32768, main.cpp:321)
    a.fc.bn_w = c->bn_w;
    a.fc.bn_h = c->bn_h;
    a.albedo = c->albedo;
    a.normal = c->normal;
    a.material = c->material;
    a.emission = c->emission;
    a.depth = c->depth;
    a.blue_noise = c->blue_noise;
    a.lights = c->d_lights;
    a.out = c->lightA;
    a.scene = TraceScene{c->tlas.nodes, c->d_recs, c->d_inst_boxes, c->min_node_lanes};
    a.rows = shade_bands(c); // own bands plus the row above and below each (wrapping) that TAA's 3x3 taps read
    a.shadow_mask = c->d_shadow_mask;
    a.shadow_words = (uint32_t)c->shadow_mask_words;
    a.ao_mask = c->d_ao_mask;
    a.ao_words = (uint32_t)c->ao_mask_words;
    a.stats = c->d_stats;
    a.lit_counters = c->d_lit;
    a.tile_counter = reinterpret_cast<uint32_t*>(c->d_lit + 8); // a spare word of the counter block (only every 16th u64 counts)
    a.shadow_maps = c->d_shadow_recs;
    a.pow22 = c->d_pow22;
    a.exact_math = (c->debug & LUZRT_DEBUG_EXACT_MATH) ? 1u : 0u;
    a.hints = nullptr;
    static const bool hints_env = [] { // LUZRT_SHADOW_HINTS=0: no occluder hints (tuning / A-B runs)
        const char* e = getenv("LUZRT_SHADOW_HINTS");
        return !(e && e[0] == '0');
    }();
    if (hints_env && !(c->debug & LUZRT_DEBUG_NO_HINTS) && c->fc.shadow_type == LUZW_SHADOW_RAYTRACING && c->fc.num_lights > 0) {
        // hint tiles of 16 x 8 pixels; one hint per warp tile (8 x 4) was measured and lost: four times the hint rays for
        // a fallback rate that barely moves (C3 rays 4.15 -> 4.26 ms, C4 71.7 -> 75.0; LUZRT_HINT_TILE=8x4 for A/B runs)
        static const bool fine = [] {
            const char* e = getenv("LUZRT_HINT_TILE");
            return e && e[0] == '8';
        }();
        a.hint_sx = fine ? 3u : 4u;
        a.hint_sy = fine ? 2u : 3u;
        const size_t tiles = (size_t)((c->w + (1u << a.hint_sx) - 1u) >> a.hint_sx) * ((a.rows.rows + (1u << a.hint_sy) - 1u) >> a.hint_sy) * a.rows.n_bands;
        int rc;
        if ((rc = grow(c, c->d_hints, c->hints_cap, tiles * (size_t)c->fc.num_lights)) != LUZRT_OK) return rc;
        a.hints = c->d_hints;
    }
    a.ray_hints = nullptr;
    a.n_instances = c->n_inst;
    a.inst_order = c->tlas.prim_order;
    a.inst_leaf = c->d_inst_inv;
    {   // per-ray temporal occluder hints: 16 bytes per shadow ray of the frame, kept from frame to frame.  They never need
        // invalidating (an entry is bounds-checked and then only decides which triangle is tried first), so a resize or a
        // different light count just starts from a cleared array.  LUZRT_TEMPORAL_HINTS=0 / LUZRT_DEBUG_NO_TEMPORAL: off;
        // LUZRT_TEMPORAL_HINTS_MB: memory budget (default 24 GB of the 180).
        static const long budget_mb = [] {
            const char* e = getenv("LUZRT_TEMPORAL_HINTS");
            if (e && e[0] == '0') return 0L;
            const char* m = getenv("LUZRT_TEMPORAL_HINTS_MB");
            return m ? atol(m) : 24576L;
        }();
        const size_t need = (size_t)c->shadow_bits * px * (c->shadow_mask_words <= 2 ? 2u : 1u); // two slots per ray up to 64 rays per pixel
        if (budget_mb > 0 && !(c->debug & LUZRT_DEBUG_NO_TEMPORAL) && c->shadow_bits > 0 && c->shadow_mask_words <= 8 &&
            c->fc.shadow_type == LUZW_SHADOW_RAYTRACING && need * sizeof(uint2) <= (size_t)budget_mb << 20) {
            if (c->ray_hints_cap != need) { // (re)allocate and clear: all entries out of bounds
                CU(c, cudaStreamSynchronize(c->stream));
                if (c->d_ray_hints) cudaFree(c->d_ray_hints);
                c->d_ray_hints = nullptr;
                c->ray_hints_cap = 0;
                if (c->ray_hints_refused != need && cudaMalloc(&c->d_ray_hints, need * sizeof(uint2)) == cudaSuccess) {
                    c->ray_hints_cap = need;
                    CU(c, cudaMemsetAsync(c->d_ray_hints, 0xFF, need * sizeof(uint2), c->stream));
                } else { // no room for the hints: an optimisation the frame can do without (do not ask again for this size)
                    cudaGetLastError();
                    c->d_ray_hints = nullptr;
                    c->ray_hints_refused = need;
                }
            }
            if (c->d_ray_hints && !c->d_temporal_cnt) {
                CU(c, cudaMalloc(&c->d_temporal_cnt, 64 * 16 * sizeof(unsigned long long)));
                CU(c, cudaMallocHost(&c->h_temporal_cnt, 64 * 16 * sizeof(unsigned long long)));
                CU(c, cudaEventCreateWithFlags(&c->ev_temporal, cudaEventDisableTiming));
            }
            if (c->temporal_pending && cudaEventQuery(c->ev_temporal) == cudaSuccess) { // counters of an earlier frame have landed
                unsigned long long fired = 0, settled = 0;
                for (int i = 0; i < 64; i++) fired += c->h_temporal_cnt[16 * i], settled += c->h_temporal_cnt[16 * i + 1];
                c->temporal_pending = false;
                if (fired && c->temporal_pending_warm) { // (a frame whose hints were not written by the frame before says nothing)
                    c->temporal_rate = (float)((double)settled / (double)fired);
                    // hysteresis: on above 30 % settled, off below 20 % (C3 0.60, C2 0.24-0.39, C5 with every instance moving 0.17)
                    if (c->temporal_rate >= 0.30f) c->temporal_use = true;
                    else if (c->temporal_rate < 0.20f) c->temporal_use = false;
                }
            }
            static const bool adaptive_env = [] { // LUZRT_TEMPORAL_HINTS=force: always on (A/B runs)
                const char* e = getenv("LUZRT_TEMPORAL_HINTS");
                return !(e && e[0] == 'f');
            }();
            c->temporal_frame++;
            const bool probe = (c->temporal_frame & 31u) < 2u; // two frames out of 32
            if (c->d_ray_hints && (c->temporal_use || probe || !adaptive_env)) {
                a.ray_hints = c->d_ray_hints;
                a.temporal_counters = c->d_temporal_cnt;
            }
            static const bool count_env = [] { // measurement aid (profiles/tools/temporal_stats.py)
                const char* e = getenv("LUZRT_TEMPORAL_COUNT");
                return e && e[0] == '1';
            }();
            a.count_temporal = count_env ? 1u : 0u;
            if (count_env) CU(c, cudaMemsetAsync(c->d_stats, 0, sizeof(DeviceStats), c->stream));
        }
    }
    a.count_row_begin = c->world == 1 ? 0u : 1u; // the halo rows are recomputation, not frame rays
    a.count_row_end = c->world == 1 ? c->h : 1u + c->band_rows;
    CU(c, wait_gather(c, a.out));
    CU(c, wait_download(c, a.out));
    ev_begin(c, EV_LIGHT);
    CU(c, cudaMemsetAsync(c->d_lit, 0, 64 * 16 * sizeof(unsigned long long), c->stream));
    if (stats) CU(c, cudaMemsetAsync(c->d_stats, 0, sizeof(DeviceStats), c->stream));
    CU(c, cudaEventRecord(c->ev[EV_LIGHT_RAYS][0], c->stream));
    if (!c->aux_stream) { // second stream of the light pass (hint pass + shadow rays next to the AO rays)
        CU(c, cudaStreamCreateWithFlags(&c->aux_stream, cudaStreamNonBlocking));
        CU(c, cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
        CU(c, cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    }
    const bool temporal_kernel = a.ray_hints && !stats && a.shadow_words <= 8;
    // Once the per-ray hints are warm (written by the frame before) the tile hints only serve the rays that were
    // unoccluded last time -- and those mostly are again: the hint pass costs more than it saves (C3 rays 2.81 -> 2.76 ms,
    // C4 30.9 -> 28.8, a rank of eight 0.51 -> 0.47).  The first frame of a run of temporal frames keeps them.
    if (temporal_kernel && c->temporal_consecutive >= 1u) a.hints = nullptr;
    c->temporal_last_on = temporal_kernel;
    c->temporal_consecutive = temporal_kernel ? c->temporal_consecutive + 1u : 0u;
    if (temporal_kernel) CU(c, cudaMemsetAsync(c->d_temporal_cnt, 0, 64 * 16 * sizeof(unsigned long long), c->stream));
    CU(c, launch_light_pass(c->stream, a, stats, c->ev[EV_LIGHT_RAYS][1], &c->launches, c->aux_stream, c->ev_fork, c->ev_join));
    if (temporal_kernel && !c->temporal_pending) { // bring this frame's counters back without stalling anybody
        CU(c, cudaMemcpyAsync(c->h_temporal_cnt, c->d_temporal_cnt, 64 * 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
        CU(c, cudaEventRecord(c->ev_temporal, c->stream));
        c->temporal_pending = true;
        c->temporal_pending_warm = c->temporal_consecutive >= 2u;
    }
    c->ev_valid[EV_LIGHT_RAYS] = true;
    ev_end(c, EV_LIGHT);
    return LUZRT_OK;
}

namespace {
// 4x4 inverse (column-major) in double by Gauss-Jordan with partial pivoting; false if singular
bool invert4(const float* m, float* out) {
    double a[4][8];
    for (int r = 0; r < 4; r++)
        for (int col = 0; col < 4; col++) {
            a[r][col] = m[col * 4 + r];
            a[r][4 + col] = r == col ? 1.0 : 0.0;
        }
    for (int i = 0; i < 4; i++) {
        int piv = i;
        for (int r = i + 1; r < 4; r++)
            if (fabs(a[r][i]) > fabs(a[piv][i])) piv = r;
        if (a[piv][i] == 0.0 || a[piv][i] != a[piv][i]) return false;
        if (piv != i)
            for (int k = 0; k < 8; k++) std::swap(a[i][k], a[piv][k]);
        const double d = a[i][i];
        for (int k = 0; k < 8; k++) a[i][k] /= d;
        for (int r = 0; r < 4; r++)
            if (r != i) {
                const double f = a[r][i];
                for (int k = 0; k < 8; k++) a[r][k] -= f * a[i][k];
            }
    }
    for (int r = 0; r < 4; r++)
        for (int col = 0; col < 4; col++) out[col * 4 + r] = (float)a[r][4 + col];
    return true;
}
// determinant of rows (r0, r1, r2) x columns 0..2 of a column-major mat4
double det3_rows(const float* m, int r0, int r1, int r2) {
    const int rows[3] = {r0, r1, r2};
    double a[3][3];
    for (int i = 0; i < 3; i++)
        for (int col = 0; col < 3; col++) a[i][col] = m[col * 4 + rows[i]];
    return a[0][0] * (a[1][1] * a[2][2] - a[1][2] * a[2][1]) - a[0][1] * (a[1][0] * a[2][2] - a[1][2] * a[2][0]) +
           a[0][2] * (a[1][0] * a[2][1] - a[1][1] * a[2][0]);
}
} // namespace

int luzrt_shadow_map_pass(luzrt_ctx* c, uint32_t resolution) {
    if (!c) return LUZRT_E_INVALID;
    if (!c->have_scene) return fail(c, LUZRT_E_STATE, "luzrt_set_scene has not been called");
    if (!c->have_tlas) return fail(c, LUZRT_E_STATE, "luzrt_tlas_build has not been called");
    REQUIRE(c, resolution >= 1 && resolution <= 16384, "shadow map resolution must be in [1, 16384]");
    DeviceGuard g(c->device);
    const size_t n = c->host_lights.size();
    c->shadow_data.resize(std::max(c->shadow_data.size(), n), nullptr);
    c->shadow_cap.resize(std::max(c->shadow_cap.size(), n), 0);
    c->host_shadow_recs.assign(std::max<size_t>(n, 1), ShadowMapRec{});
    ev_begin(c, EV_SHADOWMAP);
    for (size_t i = 0; i < n; i++) {
        const luzw_light_block& l = c->host_lights[i];
        // the reference renders a map for every light every frame (main.cpp:260-264); only these are ever sampled
        const bool needed = (c->fc.shadow_type == LUZW_SHADOW_MAP && l.shadow_map != -1) ||
                            l.volumetric_type == LUZW_VOLUMETRIC_SHADOW_MAP;
        if (!needed) continue;
        ShadowMapArgs a{};
        a.scene = TraceScene{c->tlas.nodes, c->d_recs, c->d_inst_boxes, c->min_node_lanes};
        a.res = resolution;
        a.is_cube = l.type == LUZW_LIGHT_POINT ? 1 : 0;
        a.layers = a.is_cube ? 6u : 1u;
        a.z_far = l.z_far;
        memcpy(a.eye, l.position, 12);
        if (a.is_cube) {
            for (int f = 0; f < 6; f++) {
                const double d = det3_rows(l.view_proj[f], 0, 1, 3);
                if (d == 0.0 || d != d) return fail(c, LUZRT_E_INVALID, "light %zu: viewProj[%d] is singular", i, f);
                a.cull_sign[f] = d < 0.0 ? -1.0f : 1.0f;
            }
        } else {
            const float* m = l.view_proj[0];
            if (m[3] != 0.0f || m[7] != 0.0f || m[11] != 0.0f || m[15] != 1.0f)
                return fail(c, LUZRT_E_INVALID, "light %zu: viewProj[0] of a spot/directional light must be orthographic "
                                                "(GPUScene.cpp:278-310)", i);
            const double d = det3_rows(m, 0, 1, 2);
            if (d == 0.0 || d != d || !invert4(m, a.inv_view_proj))
                return fail(c, LUZRT_E_INVALID, "light %zu: viewProj[0] is singular", i);
            a.cull_sign[0] = d < 0.0 ? -1.0f : 1.0f;
        }
        const size_t need = (size_t)a.layers * resolution * resolution;
        int rc;
        if ((rc = grow(c, c->shadow_data[i], c->shadow_cap[i], need)) != LUZRT_OK) return rc;
        a.out = c->shadow_data[i];
        CU(c, launch_shadow_map(c->stream, a));
        c->launches++;
        ShadowMapRec& r = c->host_shadow_recs[i];
        memcpy(r.view_proj, l.view_proj[0], 64);
        r.data = c->shadow_data[i];
        r.res = resolution;
        r.layers = a.layers;
        r.z_far = l.z_far;
    }
    int rc;
    if ((rc = grow(c, c->d_shadow_recs, c->shadow_recs_cap, c->host_shadow_recs.size())) != LUZRT_OK) return rc;
    if ((rc = stage_upload(c, c->d_shadow_recs, c->host_shadow_recs.data(),
                           c->host_shadow_recs.size() * sizeof(ShadowMapRec))) != LUZRT_OK)
        return rc;
    ev_end(c, EV_SHADOWMAP);
    c->shadow_maps_current = true;
    return LUZRT_OK;
}

int luzrt_read_shadow_map(luzrt_ctx* c, uint32_t light, void* dst, size_t bytes) {
    if (!c) return LUZRT_E_INVALID;
    REQUIRE(c, dst, "dst is null");
    if (!c->shadow_maps_current) return fail(c, LUZRT_E_STATE, "luzrt_shadow_map_pass has not been called for this scene block");
    REQUIRE(c, light < c->host_shadow_recs.size() && c->host_shadow_recs[light].data, "this light has no shadow map");
    const ShadowMapRec& r = c->host_shadow_recs[light];
    const size_t need = (size_t)r.layers * r.res * r.res * sizeof(float);
    REQUIRE(c, bytes >= need, "buffer too small for the shadow map");
    DeviceGuard g(c->device);
    CU(c, cudaMemcpyAsync(dst, r.data, need, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return LUZRT_OK;
}

int luzrt_volumetric_pass(luzrt_ctx* c, uint32_t frame) {
    if (!c) return LUZRT_E_INVALID;
    if (!c->w) return fail(c, LUZRT_E_STATE, "luzrt_resize has not been called");
    if (!c->have_scene) return fail(c, LUZRT_E_STATE, "luzrt_set_scene has not been called");
    if (c->n_vol_lights == 0) return LUZRT_OK; // AnyVolumetricLight() == false (main.cpp:275)
    if (c->has_shadow_map_volumetric && !c->shadow_maps_current)
        return fail(c, LUZRT_E_STATE, "volumetricType 2: luzrt_shadow_map_pass has not been called for this scene block");
    if (!c->blue_noise) return fail(c, LUZRT_E_STATE, "luzrt_set_blue_noise has not been called");
    DeviceGuard g(c->device);
    VolumetricArgs a{};
    a.fc = c->fc;
    a.fc.frame_mod = (int)((int32_t)frame % 128);
    a.fc.bn_w = c->bn_w;
    a.fc.bn_h = c->bn_h;
    a.depth = c->depth;
    a.blue_noise = c->blue_noise;
    a.lights = c->d_vol_lights;
    a.n_lights = c->n_vol_lights;
    a.light = c->lightA;
    a.rows = shade_bands(c);
    CU(c, wait_gather(c, a.light));
    CU(c, wait_download(c, a.light));
    ev_begin(c, EV_VOLUMETRIC);
    a.shadow_maps = c->d_shadow_recs;
    bool any_screen = false;
    for (const luzw_light_block& l : c->host_lights) any_screen |= l.volumetric_type == LUZW_VOLUMETRIC_SCREEN_SPACE;
    if (any_screen) { // DeferredRenderer::ScreenSpaceVolumetricLightPass
        CU(c, launch_volumetric_screen(c->stream, a));
        c->launches++;
    }
    if (c->has_shadow_map_volumetric) { // DeferredRenderer::ShadowMapVolumetricLightPass
        CU(c, launch_volumetric_shadow_map(c->stream, a));
        c->launches++;
    }
    ev_end(c, EV_VOLUMETRIC);
    return LUZRT_OK;
}

int luzrt_taa_pass(luzrt_ctx* c, int reconstruct) {
    if (!c) return LUZRT_E_INVALID;
    if (!c->w) return fail(c, LUZRT_E_STATE, "luzrt_resize has not been called");
    if (!c->have_scene) return fail(c, LUZRT_E_STATE, "luzrt_set_scene has not been called");
    DeviceGuard g(c->device);
    TaaArgs a{};
    a.fc = c->fc;
    a.light_in = c->lightA;
    // the reference never initialises lightHistory (DeferredRenderer.cpp:225-231); this path defines
    // the first frame after a resize as history == current light buffer
    a.history = c->history_valid ? c->lightHist : c->lightA;
    a.depth = c->depth;
    a.out = c->lightB;
    a.rows = own_bands(c);
    a.reconstruct = reconstruct ? 1 : 0;
    CU(c, wait_gather(c, a.light_in, a.history, a.out));
    CU(c, wait_download(c, a.out));
    ev_begin(c, EV_TAA);
    CU(c, launch_taa_pass(c->stream, a, (c->debug & LUZRT_DEBUG_EXACT_MATH) != 0));
    c->launches++;
    ev_end(c, EV_TAA);
    std::swap(c->lightA, c->lightB); // DeferredRenderer.cpp:443
    return LUZRT_OK;
}

int luzrt_gather(luzrt_ctx* c) {
    if (!c) return LUZRT_E_INVALID;
    if (c->world == 1) return LUZRT_OK;
    if (!c->comm) return fail(c, LUZRT_E_STATE, "luzrt_comm_init has not been called");
    if (!c->w) return fail(c, LUZRT_E_STATE, "luzrt_resize has not been called");
    DeviceGuard g(c->device);
    const size_t strip_floats = (size_t)c->rows_per_rank * c->w * 4; // a rank's bands are contiguous in storage order
    CU(c, wait_gather(c)); // one gather in flight at a time
    CU(c, cudaEventRecord(c->ev_resolved, c->stream));
    CU(c, cudaStreamWaitEvent(c->comm_stream, c->ev_resolved, 0));
    cudaEventRecord(c->ev[EV_GATHER][0], c->comm_stream);
    // in place: this rank's rows already sit at their offset inside the full-frame buffer
    ncclResult_t r = g_nccl.AllGather((const float*)c->lightA + (size_t)c->rank * strip_floats, c->lightA, strip_floats,
                                      ncclFloat, c->comm, c->comm_stream);
    cudaEventRecord(c->ev[EV_GATHER][1], c->comm_stream);
    c->ev_valid[EV_GATHER] = true;
    if (r != ncclSuccess) return fail(c, LUZRT_E_COMM, "ncclAllGather: %s", g_nccl.GetErrorString(r));
    CU(c, cudaEventRecord(c->ev_gathered, c->comm_stream));
    c->gather_buf = c->lightA;
    return LUZRT_OK;
}

int luzrt_compose_pass(luzrt_ctx* c, float exposure) {
    if (!c) return LUZRT_E_INVALID;
    (void)exposure; // present.frag's active path (TonemapACES, :87) ignores ctx.exposure
    if (!c->w) return fail(c, LUZRT_E_STATE, "luzrt_resize has not been called");
    DeviceGuard g(c->device);
    ev_begin(c, EV_COMPOSE);
    // reads only this rank's own rows of lightA, which a gather in flight does not write
    CU(c, launch_compose_pass(c->stream, c->fc, c->lightA, c->compose, own_bands(c)));
    c->launches++;
    ev_end(c, EV_COMPOSE);
    return LUZRT_OK;
}

int luzrt_swap_light_history(luzrt_ctx* c) {
    if (!c) return LUZRT_E_INVALID;
    std::swap(c->lightA, c->lightHist); // DeferredRenderer.cpp:469-471
    c->history_valid = true;
    return LUZRT_OK;
}

int luzrt_read(luzrt_ctx* c, int which, void* dst, size_t bytes) {
    if (!c) return LUZRT_E_INVALID;
    REQUIRE(c, dst, "dst is null");
    DeviceGuard g(c->device);
    CU(c, wait_gather(c));
    CU(c, cudaStreamSynchronize(c->stream));
    if (which == LUZRT_STATS_DETAIL) {
        REQUIRE(c, bytes >= sizeof(unsigned long long) * 40, "buffer too small for the 40 detail counters");
        CU(c, cudaMemcpy(dst, reinterpret_cast<const char*>(c->d_stats) + offsetof(DeviceStats, detail),
                         sizeof(unsigned long long) * 40, cudaMemcpyDeviceToHost));
        return LUZRT_OK;
    }
    if (which == LUZRT_STATS) {
        REQUIRE(c, bytes >= sizeof(luzrt_stats), "buffer too small for luzrt_stats");
        DeviceStats ds{};
        unsigned long long lit[64 * 16];
        CU(c, cudaMemcpy(&ds, c->d_stats, sizeof(ds), cudaMemcpyDeviceToHost));
        CU(c, cudaMemcpy(lit, c->d_lit, sizeof(lit), cudaMemcpyDeviceToHost));
        luzrt_stats s{};
        for (int i = 0; i < 64; i++) s.lit_pixels += lit[16 * i];
        if (c->debug & LUZRT_DEBUG_STATS) {
            s.rays = ds.rays;
            s.nodes_visited = ds.nodes;
            s.triangles_tested = ds.tris;
            s.instances_entered = ds.insts;
            s.rays_occluded = ds.occluded;
        } else { // every lit pixel traces the same number of rays
            s.rays = s.lit_pixels * c->rays_per_lit_pixel;
        }
        memcpy(dst, &s, sizeof(s));
        return LUZRT_OK;
    }
    if (which == LUZRT_TIMINGS) {
        REQUIRE(c, bytes >= sizeof(luzrt_timings), "buffer too small for luzrt_timings");
        float ms[EV_COUNT] = {0};
        for (int i = 0; i < EV_COUNT; i++)
            if (c->ev_valid[i]) cudaEventElapsedTime(&ms[i], c->ev[i][0], c->ev[i][1]);
        luzrt_timings t{ms[EV_TLAS], ms[EV_GBUF], ms[EV_LIGHT], ms[EV_TAA], ms[EV_GATHER], ms[EV_COMPOSE], ms[EV_VOLUMETRIC], ms[EV_SHADOWMAP], ms[EV_LIGHT_RAYS],
                        c->temporal_rate, c->temporal_last_on ? 1.0f : 0.0f};
        memcpy(dst, &t, sizeof(t));
        return LUZRT_OK;
    }
    size_t have = 0;
    void* src = image_ptr(c, which, &have);
    if (!src) return fail(c, LUZRT_E_INVALID, "selector %d has no data", which);
    REQUIRE(c, bytes >= have, "destination buffer too small");
    if (c->world > 1 && is_light_image(which)) return luzrt_read_rows(c, which, 0, c->h, dst, bytes);
    CU(c, cudaMemcpy(dst, src, have, cudaMemcpyDeviceToHost));
    return LUZRT_OK;
}

int luzrt_read_rows(luzrt_ctx* c, int which, uint32_t y0, uint32_t y1, void* dst, size_t bytes) {
    if (!c) return LUZRT_E_INVALID;
    REQUIRE(c, dst, "dst is null");
    REQUIRE(c, y0 <= y1 && y1 <= c->h, "row range out of bounds");
    DeviceGuard g(c->device);
    CU(c, wait_gather(c));
    size_t have = 0;
    void* src = image_ptr(c, which, &have);
    if (!src || !c->h) return fail(c, LUZRT_E_INVALID, "selector %d has no data", which);
    const size_t row = have / c->h, need = row * (y1 - y0);
    REQUIRE(c, bytes >= need, "destination buffer too small");
    if (c->world > 1 && is_light_image(which) && y1 > y0) {
        // the light images of a partitioned frame are stored band-permuted: bring the rows to natural order first
        int rc;
        if ((rc = grow(c, c->unperm, c->unperm_cap, (size_t)c->w * (y1 - y0))) != LUZRT_OK) return rc;
        CU(c, launch_unpermute_rows(c->stream, c->fc, (const float4*)src, c->unperm, y0, y1));
        c->launches++;
        CU(c, cudaMemcpyAsync(dst, c->unperm, need, cudaMemcpyDeviceToHost, c->stream));
    } else {
        CU(c, cudaMemcpyAsync(dst, (const char*)src + row * y0, need, cudaMemcpyDeviceToHost, c->stream));
    }
    CU(c, cudaStreamSynchronize(c->stream));
    return LUZRT_OK;
}

int luzrt_owned_bands(luzrt_ctx* c, uint32_t* first_row, uint32_t* band_rows, uint32_t* pitch, uint32_t* n_bands) {
    if (!c || !first_row || !band_rows || !pitch || !n_bands) return LUZRT_E_INVALID;
    if (!c->w) return fail(c, LUZRT_E_STATE, "luzrt_resize has not been called");
    const BandSet b = own_bands(c);
    *first_row = (uint32_t)b.first;
    *band_rows = b.rows;
    *pitch = b.pitch;
    *n_bands = b.n_bands;
    return LUZRT_OK;
}

int luzrt_read_owned(luzrt_ctx* c, int which, void* dst, size_t bytes) {
    if (!c) return LUZRT_E_INVALID;
    REQUIRE(c, dst, "dst is null");
    if (!c->w) return fail(c, LUZRT_E_STATE, "luzrt_resize has not been called");
    DeviceGuard g(c->device);
    size_t have = 0;
    void* src = image_ptr(c, which, &have);
    if (!src) return fail(c, LUZRT_E_INVALID, "selector %d has no data", which);
    const size_t row = have / c->h, need = row * c->rows_per_rank;
    REQUIRE(c, bytes >= need, "destination buffer too small");
    if (is_light_image(which) || c->world == 1) {
        // banded storage keeps a rank's rows contiguous: one copy
        CU(c, cudaMemcpyAsync(dst, (const char*)src + row * c->rows_per_rank * (size_t)c->rank, need, cudaMemcpyDeviceToHost,
                              c->stream));
    } else {
        const BandSet b = own_bands(c);
        CU(c, cudaMemcpy2DAsync(dst, row * b.rows, (const char*)src + row * (size_t)b.first, row * b.pitch, row * b.rows,
                                b.n_bands, cudaMemcpyDeviceToHost, c->stream));
    }
    CU(c, cudaStreamSynchronize(c->stream));
    return LUZRT_OK;
}

int luzrt_probe_read_bandwidth(luzrt_ctx* c, size_t bytes, int iters, double* out_gbs) {
    if (!c) return LUZRT_E_INVALID;
    REQUIRE(c, out_gbs && bytes >= (1u << 20) && iters > 0, "bad probe arguments");
    DeviceGuard g(c->device);
    void* buf = nullptr;
    float* sink = nullptr;
    CU(c, cudaMalloc(&buf, bytes));
    CU(c, cudaMalloc(&sink, 4096));
    CU(c, cudaMemsetAsync(buf, 1, bytes, c->stream));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaError_t err = launch_probe_read(c->stream, buf, bytes, 2, sink); // warm the cache
    cudaEventRecord(e0, c->stream);
    if (err == cudaSuccess) err = launch_probe_read(c->stream, buf, bytes, iters, sink);
    cudaEventRecord(e1, c->stream);
    c->launches += 2;
    cudaStreamSynchronize(c->stream);
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(buf);
    cudaFree(sink);
    if (err != cudaSuccess) return fail(c, LUZRT_E_CUDA, "probe kernel: %s", cudaGetErrorString(err));
    *out_gbs = ms > 0.0f ? (double)bytes * iters / (ms * 1e6) : 0.0;
    return LUZRT_OK;
}

int luzrt_device_ptr(luzrt_ctx* c, int which, void** out_ptr, size_t* out_bytes) {
    if (!c) return LUZRT_E_INVALID;
    REQUIRE(c, out_ptr && out_bytes, "null output");
    DeviceGuard g(c->device);
    CU(c, wait_gather(c)); // work the host enqueues on the ctx stream after this call sees the gathered frame
    *out_ptr = image_ptr(c, which, out_bytes);
    return *out_ptr ? LUZRT_OK : fail(c, LUZRT_E_INVALID, "selector %d has no data", which);
}

int luzrt_sync(luzrt_ctx* c) {
    if (!c) return LUZRT_E_INVALID;
    DeviceGuard g(c->device);
    CU(c, wait_gather(c));
    CU(c, cudaStreamSynchronize(c->stream));
    return LUZRT_OK;
}

int luzrt_stream(luzrt_ctx* c, uint64_t* out_stream) {
    if (!c || !out_stream) return LUZRT_E_INVALID;
    *out_stream = (uint64_t)(uintptr_t)c->stream;
    return LUZRT_OK;
}

uint64_t luzrt_launch_count(luzrt_ctx* c) { return c ? c->launches : 0; }

} // extern "C"

// common.cuh -- shared types for libluzrt (sm_100a).  See DESIGN.md for the memory layout.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/luzrt.h"

namespace luz {

// ---- 8-wide BVH node, 208 bytes: a 16-byte header and six planes x eight children in fp32 ----------
// The acceleration structure of every instanced configuration is a few MB and lives in L1/L2, and the
// traversal is bound by instruction issue on the ALU pipe, not by memory (profiles/r1_light_pass_v4.md:
// ALU pipe 62 % of peak, FMA pipe 24 %, DRAM 0.6 %).  Child boxes are therefore kept as plain floats:
// a slab distance is one FFMA on the plane (no byte extraction, no per-node grid set-up), and the near /
// far plane of an axis is picked by the ADDRESS the ray loads from (lo* and hi* of an axis are 32 bytes
// apart), not by selects.  The 8-bit quantised layout this replaces (Ylitie et al. 2017, 80 B) cost
// 181 instructions per node visit, two thirds of them on the ALU pipe; this one costs ~115.
//   meta[i] == 0                      : empty slot (its box is lo = +inf, hi = -inf: never hit)
//   meta[i] == (1<<5) | (24 + i)      : internal child; its index is child_base + popc(imask & ((1<<i)-1))
//   meta[i] == (unary(count)<<5) | off: leaf of `count` (1..3) primitives prim_base + off .. +count-1
struct __align__(16) WideNode {
    uint32_t child_base_imask; // bits 23..0: index of the first internal child; bits 31..24: imask
    uint32_t prim_base;
    uint8_t meta[8];
    float lox[8], hix[8], loy[8], hiy[8], loz[8], hiz[8];
};
static_assert(sizeof(WideNode) == 208, "WideNode must be 208 bytes");
constexpr uint32_t kMaxWideNodes = 1u << 24; // child_base is 24 bits

// One triangle, 96 bytes, prepared for the edge test in Pluecker coordinates (traverse.cuh tri_test): per edge the
// moment q x p and the direction p - q of its two vertices, so that the signed volume of the ray against the edge is
// six multiply-adds on the ray's (d, o x d) with no per-triangle translation; the plane (N, k = N . p0) gives t.
// Reversing an edge negates its moment and its direction bit by bit, which is what makes shared edges watertight.
struct __align__(16) WideTri {
    float4 mu; // xyz = p2 x p1 (edge opposite p0),  w = original triangle index bits
    float4 eu; // xyz = p1 - p2,                     w = k = N . p0
    float4 mv; // xyz = p0 x p2 (edge opposite p1),  w = N.x     N = (p1 - p0) x (p2 - p0)
    float4 ev; // xyz = p2 - p0,                     w = N.y
    float4 mw; // xyz = p1 x p0 (edge opposite p2),  w = N.z
    float4 ew; // xyz = p0 - p1,                     w = 0
};
static_assert(sizeof(WideTri) == 96, "WideTri must be 96 bytes");

// One TLAS leaf record, 64 bytes: world->object 3x4 (rows) + the BLAS it points to.
struct __align__(16) InstanceRec {
    float4 r0, r1, r2; // rows of the inverse affine transform
    const WideNode* nodes;
    const WideTri* tris;
};
static_assert(sizeof(InstanceRec) == 64, "InstanceRec must be 64 bytes");

struct InstanceMeta { // parallel to InstanceRec (G-buffer pass only)
    uint32_t custom_index;
    uint32_t blas_slot;
};

struct BlasAttr { // per-BLAS vertex attributes kept for the G-buffer pass
    const uint8_t* vertices;
    const uint32_t* indices;
    uint32_t stride;
    uint32_t has_attr;
};

struct TraceScene {
    const WideNode* tlas_nodes;
    const InstanceRec* instances;
    const float4* inst_boxes; // world boxes of the instances, TLAS leaf order: [2i] = lo.xyz, [2i+1] = hi.xyz
    uint32_t min_node_lanes; // the node loop of trace_ray yields once fewer lanes than this remain in it
};

// compact light record staged in shared memory (first 64 bytes of a LightBlock)
struct __align__(16) LightRec {
    float4 color_intensity;
    float4 position_inner;
    float4 direction_outer;
    int type;
    int num_shadow_samples;
    float radius;
    int shadow_map;
};
static_assert(sizeof(LightRec) == 64, "LightRec");

struct FrameConst { // what the kernels need from SceneBlock, passed by value (kernel param space)
    float inverse_proj[16];
    float inverse_view[16];
    float view_proj[16];
    float prev_view_proj[16];
    float jitter[2], prev_jitter[2];
    float cam_pos[3];
    float ambient[3]; // ambientLightColor * ambientLightIntensity
    float ao_min, ao_max;
    int ao_num_samples;
    int num_lights;
    int shadow_type;
    int frame_mod; // frame % 128
    uint32_t width, height;
    uint32_t bn_w, bn_h;
    // Image partition over the GPUs of a box: the frame is cut into bands of band_rows
    // rows, band b belongs to rank b % world.  The RGBA32F light images are stored with every rank's rows
    // contiguous (rank-major, then band, then row) so that one in-place all-gather assembles them; the
    // G-buffer keeps the natural row order.  world == 1: band_rows == height and storage_row is the identity.
    uint32_t band_rows;     // rows per band
    uint32_t band_magic;    // ceil(2^32 / band_rows): y / band_rows == umulhi(y, band_magic) for y < 65536
    uint32_t world_shift;   // log2(world)
    uint32_t rows_per_rank; // height / world
};

// Rows [band_first + z * band_pitch, + rows) for z = 0 .. n_bands-1 (wrapping at the image border like the
// REPEAT sampler): the row sets a rank shades (own bands + one halo row each side) or resolves (own bands).
struct BandSet {
    int first;          // first row of band 0 (may be -1: the halo row above row 0 wraps to height-1)
    uint32_t pitch;     // rows between the starts of successive bands of this rank
    uint32_t rows;      // rows per band in this set
    uint32_t n_bands;
};

struct DeviceStats {
    unsigned long long lit_pixels, rays, nodes, tris, insts, occluded;
    // LUZRT_STATS_DETAIL (statistics variant only): per ray class c in {0 hinted shadow, 1 unhinted shadow, 2 AO with a
    // candidate list, 3 AO descending from the root} detail[8 c + {0 rays, 1 occluded, 2 TLAS nodes, 3 BLAS nodes,
    // 4 triangles, 5 instances entered, 6 root descents, 7 unused}]; detail[32..] = {AO pixels, AO pixels with an empty
    // list, AO pixels whose list overflowed, candidates kept, nodes of the per-pixel box queries + filters, hint rays,
    // hint rays that hit}
    unsigned long long detail[40];
};

// ---- small vector helpers ------------------------------------------------------------------
__host__ __device__ inline float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
__host__ __device__ inline float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
__host__ __device__ inline float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
__host__ __device__ inline float3 operator-(float3 a) { return f3(-a.x, -a.y, -a.z); }
__host__ __device__ inline float3 operator*(float3 a, float3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
__host__ __device__ inline float3 operator*(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
__host__ __device__ inline float3 operator*(float s, float3 a) { return f3(a.x * s, a.y * s, a.z * s); }
__host__ __device__ inline float3 operator/(float3 a, float s) { return f3(a.x / s, a.y / s, a.z / s); }
__host__ __device__ inline float3 operator/(float3 a, float3 b) { return f3(a.x / b.x, a.y / b.y, a.z / b.z); }
__host__ __device__ inline float dot3(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__host__ __device__ inline float3 cross3(float3 a, float3 b) {
    return f3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}
__device__ inline float length3(float3 a) { return sqrtf(dot3(a, a)); }
__device__ inline float3 normalize3(float3 a) { return a / sqrtf(dot3(a, a)); }
__device__ inline float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

__host__ __device__ inline float4 f4(float x, float y, float z, float w) { return make_float4(x, y, z, w); }
__host__ __device__ inline float4 operator+(float4 a, float4 b) { return f4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__host__ __device__ inline float4 operator-(float4 a, float4 b) { return f4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
__host__ __device__ inline float4 operator*(float4 a, float s) { return f4(a.x * s, a.y * s, a.z * s, a.w * s); }
__host__ __device__ inline float4 operator/(float4 a, float s) { return f4(a.x / s, a.y / s, a.z / s, a.w / s); }
__device__ inline float4 min4(float4 a, float4 b) { return f4(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z), fminf(a.w, b.w)); }
__device__ inline float4 max4(float4 a, float4 b) { return f4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w)); }
__device__ inline bool any_nan4(float4 a) { return isnan(a.x) || isnan(a.y) || isnan(a.z) || isnan(a.w); }

// column-major mat4 * vec4, summed left to right over columns
__device__ inline float4 mat_mul(const float* m, float4 v) {
    float4 r;
    r.x = m[0] * v.x + m[4] * v.y + m[8] * v.z + m[12] * v.w;
    r.y = m[1] * v.x + m[5] * v.y + m[9] * v.z + m[13] * v.w;
    r.z = m[2] * v.x + m[6] * v.y + m[10] * v.z + m[14] * v.w;
    r.w = m[3] * v.x + m[7] * v.y + m[11] * v.z + m[15] * v.w;
    return r;
}

__device__ __forceinline__ uint32_t storage_row(const FrameConst& fc, uint32_t y) {
    const uint32_t b = __umulhi(y, fc.band_magic);
    return (b & ((1u << fc.world_shift) - 1u)) * fc.rows_per_rank + (b >> fc.world_shift) * fc.band_rows +
           (y - b * fc.band_rows);
}
// index of the occluder-hint tile of pixel (x, row r of band z) in a hint grid of (1 << sx) x (1 << sy)-pixel tiles
__device__ __forceinline__ size_t hint_tile_index(const FrameConst& fc, const BandSet& bs, uint32_t sx, uint32_t sy, uint32_t z,
                                                  uint32_t x, uint32_t r) {
    const uint32_t hx = (fc.width + (1u << sx) - 1u) >> sx, hy = (bs.rows + (1u << sy) - 1u) >> sy;
    return (size_t)(z * hy + (r >> sy)) * hx + (x >> sx);
}
__device__ __forceinline__ uint32_t band_row(const FrameConst& fc, const BandSet& bs, uint32_t z, uint32_t r) {
    int y = bs.first + (int)(z * bs.pitch + r);
    if (y < 0) y += (int)fc.height;
    if (y >= (int)fc.height) y -= (int)fc.height;
    return (uint32_t)y;
}

// utils.glsl:1-7
__device__ inline float3 depth_to_world(const FrameConst& fc, float u, float v, float depth) {
    float4 clip = f4(u * 2.0f - 1.0f, v * 2.0f - 1.0f, depth, 1.0f);
    float4 view = mat_mul(fc.inverse_proj, clip);
    view = view / view.w;
    float4 world = mat_mul(fc.inverse_view, view);
    return f3(world.x, world.y, world.z);
}

} // namespace luz

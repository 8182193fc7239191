// light_pass.cu -- the deferred lighting pass.  Ray kernels fire every shadow and AO ray of light.frag and leave one bit
// per ray (1 = occluded) in the per-pixel visibility masks (all ray state in registers, no ray buffers in HBM);
// k_light_shade evaluates light.frag:171-235 with the occluded fractions read back from those bits.
//
//   k_shadow_hints           one ray per 16x8 tile and light: an occluding instance the tile's shadow rays try first
//   k_light_rays_persistent  the ray kernels a host gets: persistent warps pull 8x4-pixel tiles from an atomic counter;
//                            one specialised launch per kind of ray (PART 0 shadow rays, PART 1 AO rays), the AO launch
//                            overlapping the hint pass and filling the tail of the shadow launch
//   k_light_rays             every ray of a pixel in one CTA: the LUZRT_DEBUG_STATS variant (counts nodes / triangles /
//                            instances) and the reference the persistent kernels are tested against bit for bit
//   k_light_shade            (shade_kernel.cuh) Cook-Torrance over the lights, shadow factors and AO from the masks
// All ray kernels produce identical bits (tests/test_gpu_parity.py run_light).
//
// Why rays and shading are separate kernels: a single fused kernel (this file up to commit "Host path: read-back in flight ...") keeps the whole
// BRDF state of the pixel alive inside the traversal loop (albedo, F0, V, Lo, roughness ... ~30 registers) next to
// ~45 registers of traversal state, which at the occupancy that hides the traversal's latencies best (80 registers,
// 6 CTAs of 128 threads per SM) is spilled to local memory inside the hottest loop.  The ray kernel only carries what
// ray generation needs (fragPos, N, blue noise); the shading kernel has no traversal in it and runs at full lane
// utilisation.  The price is a second read of the G-buffer and 4 * (shadow_words + ao_words) bytes per pixel of
// mask traffic, ~0.1 ms of HBM time at 4K.  Measured on B200 against the fused kernel (bit-identical results):
// C3 7.31 -> 7.10 ms, C4 134.5 -> 125.5 ms, C2 1.00 -> 0.94 ms.
//
// Restates source/Shaders/light.frag:171-235 (main), :86-109 (TraceShadowRay), :111-135 (TraceAORays), :137-169
// (EvaluateShadow), :57-75 (samplers), :17-49 (BRDF); launched where DeferredRenderer::LightPass
// (DeferredRenderer.cpp:324-345) draws its quad.
#include <cstdlib>

#include "passes.h"
#include "shade_kernel.cuh"
#include "traverse.cuh"

namespace luz {

namespace {

constexpr float kPI = 3.14159265359f;              // LuzCommon.h:11
constexpr float kGoldenRatio = 2.118033988749895f; // LuzCommon.h:12 (sic)
constexpr int kLightChunk = 256;
#ifndef LUZ_MIN_CAND_SAMPLES
#define LUZ_MIN_CAND_SAMPLES 6
#endif
constexpr int kMinCandSamples = LUZ_MIN_CAND_SAMPLES; // below this the one TLAS walk per pixel does not pay for itself
constexpr int kMaxCand = 8;        // instances an AO candidate list holds before falling back to the root descent

__device__ __forceinline__ float fractf(float x) { return x - floorf(x); }

// light.frag:71-75, multiply and add rounded separately (bit-identical to the oracle's plain fp32)
__device__ __forceinline__ float2 blue_noise_sample(float bn_r, float bn_g, int i, int frame_mod) {
    const float k = (float)(128 * i + frame_mod);
    const float off = __fmul_rn(kGoldenRatio, k);
    return make_float2(fractf(__fadd_rn(bn_r, off)), fractf(__fadd_rn(bn_g, off)));
}

// ---- ray generation arithmetic ---------------------------------------------------------------------------------
// GLSL leaves the precision of normalize / sqrt / sin / cos to the driver (inversesqrt and sqrt 2 ulp, sin / cos 2^-11
// absolute in the Vulkan precision table), and every desktop driver evaluates them on the special
// function unit.  The ray kernel does the same: MUFU.RSQ / MUFU.SQRT / MUFU.SIN / MUFU.COS instead of the IEEE
// division, square root and the 40-instruction sincosf, and explicit FMAs for the frame combination (ray generation was
// 16 % of the kernel's issue slots).  Directions move by <= 1e-6 relative against the oracle's correctly rounded
// evaluation, which only flips rays that graze an edge (the parity tests bound it: >= 99.9 % of the rays agree).  The
// shading kernel keeps the correctly rounded forms: its output is compared value by value.
#ifndef LUZ_FAST_RAYGEN
#define LUZ_FAST_RAYGEN 1
#endif
#ifndef LUZ_AO_HEMISPHERE
#define LUZ_AO_HEMISPHERE 1 // hemisphere reach boxes for the per-pixel AO candidate lists (traverse.cuh)
#endif
#if LUZ_FAST_RAYGEN
__device__ __forceinline__ float rg_sqrt(float x) {
    float r;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float3 rg_normalize(float3 a) {
    const float s = rsqrtf(fmaf(a.x, a.x, fmaf(a.y, a.y, a.z * a.z))); // 0 -> inf -> NaN components, like a / 0
    return f3(a.x * s, a.y * s, a.z * s);
}
__device__ __forceinline__ float rg_length(float3 a) { return rg_sqrt(fmaf(a.x, a.x, fmaf(a.y, a.y, a.z * a.z))); }
__device__ __forceinline__ void rg_sincos(float x, float* sn, float* cs) { __sincosf(x, sn, cs); }
// a * x + b * y + c * z
__device__ __forceinline__ float3 rg_combine(float3 a, float x, float3 b, float y, float3 c, float z) {
    return f3(fmaf(a.x, x, fmaf(b.x, y, c.x * z)), fmaf(a.y, x, fmaf(b.y, y, c.y * z)), fmaf(a.z, x, fmaf(b.z, y, c.z * z)));
}
#else
__device__ __forceinline__ float rg_sqrt(float x) { return sqrtf(x); }
__device__ __forceinline__ float3 rg_normalize(float3 a) { return normalize3(a); }
__device__ __forceinline__ float rg_length(float3 a) { return length3(a); }
__device__ __forceinline__ void rg_sincos(float x, float* sn, float* cs) { sincosf(x, sn, cs); }
__device__ __forceinline__ float3 rg_combine(float3 a, float x, float3 b, float y, float3 c, float z) {
    return a * x + b * y + c * z;
}
#endif

// Appends one visibility bit to a pixel's mask; words are stored when they fill up (and by flush()).
struct BitWriter {
    uint32_t* words;
    uint32_t bit = 0, cur = 0;
    __device__ __forceinline__ void push(bool set) {
        if (set) cur |= 1u << (bit & 31u);
        bit++;
        if ((bit & 31u) == 0u) {
            if (cur) words[(bit >> 5) - 1u] = cur; // the masks are cleared before the launch
            cur = 0;
        }
    }
    __device__ __forceinline__ void flush() {
        if (cur) words[bit >> 5] = cur;
        cur = 0;
    }
};

// ---- kernel 1: rays ------------------------------------------------------------------------------------------
// One warp = 8x4 pixel tile; lights staged in shared memory; one loop over "ray sources" (the lights, then one
// pseudo source for AO) so that the kernel holds a single inlined copy of the traversal.
// EvaluateShadow (light.frag:137-146) + the set-up of TraceShadowRay (:86-91): origin O and the centre vector C of the
// shadow rays of one light at one pixel (C = unnormalised vector to the light; its length is the rays' tMax).
__device__ __forceinline__ void shadow_ray_frame(const LightRec& L4, const float3 fragPos, const float3 N, const float camDist,
                                                 float3& O, float3& C) {
    const float3 lpos = f3(L4.position_inner.x, L4.position_inner.y, L4.position_inner.z);
    const float3 ldir = f3(L4.direction_outer.x, L4.direction_outer.y, L4.direction_outer.z);
    const float3 Lvec = lpos - fragPos;
    const float dist = rg_length(Lvec);
    float3 L = rg_normalize(Lvec);
    if (L4.type == LUZW_LIGHT_DIRECTIONAL) L = rg_normalize(-ldir);
    O = fragPos + N * fmaxf(camDist * 0.01f, 0.05f);
    C = (L4.type == LUZW_LIGHT_DIRECTIONAL) ? ldir * dot3(ldir, L) * dist : L * dist;
}

// ---- occluder hints ----------------------------------------------------------------------------------------------
// The shadow rays of a 16x8 pixel tile towards one light form a beam far thinner than an instance, and in scenes
// where lights sit among geometry most tiles are entirely in shadow (C3: 95 % of the lit tiles for the three local
// lights, C4: 52 %).  An occluded ray still pays the whole TLAS descent plus the instances it enters in vain before it
// meets an occluder (~12 node visits on C3).  So one ray per tile and light -- from the tile's centre pixel along the
// centre vector -- is traced first (k_shadow_hints, < 2 % of the frame's shadow rays) and the instance it hits is the
// tile's hint: every shadow ray of the tile tries that instance first, all lanes of a warp entering the same instance
// together, and only if it misses descends from the root (trace_ray's then_root).  Any-hit visibility does not depend
// on the order in which occluders are tried, so the bits are identical with and without hints
// (test_shadow_hints_do_not_change_visibility); nothing is carried from frame to frame.
template <bool STATS>
__global__ void __launch_bounds__(128) k_shadow_hints(const LightArgs a, uint32_t* __restrict__ hints, const uint32_t tiles_x,
                                                      const uint32_t tiles_y) {
    const FrameConst& fc = a.fc;
    const uint32_t n_tiles = tiles_x * tiles_y * a.rows.n_bands;
    const uint32_t g = blockIdx.x * 128u + threadIdx.x;
    const bool valid = g < n_tiles * (uint32_t)fc.num_lights;
    LocalStats st = {0, 0, 0};
    if (valid) {
        const uint32_t light = g / n_tiles, tile = g % n_tiles; // a warp = 32 neighbouring tiles, one light
        const uint32_t bx = tile % tiles_x, by = (tile / tiles_x) % tiles_y, band = tile / (tiles_x * tiles_y);
        const uint32_t x = min((bx << a.hint_sx) + (1u << a.hint_sx) / 2u, fc.width - 1u);
        const uint32_t r = min((by << a.hint_sy) + (1u << a.hint_sy) / 2u, a.rows.rows - 1u);
        const uint32_t y = band_row(fc, a.rows, band, r);
        const size_t pix = (size_t)y * fc.width + x;
        const float4 n4 = __ldg(a.normal + pix);
        const float3 N = f3(n4.x, n4.y, n4.z);
        uint32_t hint = kNoInstance;
        const LightRec L4 = a.lights[light];
        if (length3(N) != 0.0f && L4.num_shadow_samples > 0) {
            const float u = ((float)x + 0.5f) / (float)fc.width, v = ((float)y + 0.5f) / (float)fc.height;
            const float3 fragPos = depth_to_world(fc, u, v, __ldg(a.depth + pix));
            const float camDist = length3(fragPos - f3(fc.cam_pos[0], fc.cam_pos[1], fc.cam_pos[2]));
            float3 O, C;
            shadow_ray_frame(L4, fragPos, N, camDist, O, C);
            uint2 stack[LUZ_STACK_SIZE];
            HitInfo h;
            if (trace_ray<false, STATS>(a.scene, O, rg_normalize(C), 0.001f, rg_length(C), &h, &st, stack)) hint = h.inst;
            if (STATS) {
                atomicAdd(&a.stats->detail[37], 1ull);
                if (hint != kNoInstance) atomicAdd(&a.stats->detail[38], 1ull);
            }
        }
        hints[(size_t)tile * (uint32_t)fc.num_lights + light] = hint;
    }
    if (STATS) { // the hint rays are not frame rays, but what they fetch is part of the pass
        unsigned long long vals[3] = {st.nodes, st.tris, st.insts};
#pragma unroll
        for (int k = 0; k < 3; k++)
            for (int off = 16; off; off >>= 1) vals[k] += __shfl_xor_sync(0xFFFFFFFFu, vals[k], off);
        if ((threadIdx.x & 31) == 0) {
            atomicAdd(&a.stats->nodes, vals[0]);
            atomicAdd(&a.stats->tris, vals[1]);
            atomicAdd(&a.stats->insts, vals[2]);
        }
    }
}

template <bool STATS>
__device__ __forceinline__ void light_rays_body(const LightArgs& a, const uint32_t band) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    LightRec* s_lights = reinterpret_cast<LightRec*>(smem_raw);
    uint32_t* s_cand = reinterpret_cast<uint32_t*>(smem_raw + a.cand_offset) + threadIdx.x; // [k][thread]

    const FrameConst& fc = a.fc;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t x = blockIdx.x * 16 + (warp & 1) * 8 + (lane & 7);
    const uint32_t r = blockIdx.y * 8 + (warp >> 1) * 4 + (lane >> 3);
    const bool in_image = x < fc.width && r < a.rows.rows;
    const uint32_t y = in_image ? band_row(fc, a.rows, band, r) : 0u;
    const size_t pix = (size_t)y * fc.width + x;

    float3 N = f3(0.0f, 0.0f, 0.0f);
    float depth = 1.0f;
    uchar4 bn8 = make_uchar4(0, 0, 0, 0);
    if (in_image) {
        const float4 n4 = __ldg(a.normal + pix);
        N = f3(n4.x, n4.y, n4.z);
        depth = __ldg(a.depth + pix);
        bn8 = __ldg(a.blue_noise + (size_t)(y % fc.bn_h) * fc.bn_w + (x % fc.bn_w));
    }
    const bool lit = in_image && (length3(N) != 0.0f); // light.frag:178
    const float u = ((float)x + 0.5f) / (float)fc.width, v = ((float)y + 0.5f) / (float)fc.height;
    const float3 fragPos = depth_to_world(fc, u, v, depth);
    const float3 camPos = f3(fc.cam_pos[0], fc.cam_pos[1], fc.cam_pos[2]);
    const float camDist = length3(fragPos - camPos);
    const float bn_r = (float)bn8.x / 255.0f, bn_g = (float)bn8.y / 255.0f;

    uint2 stack[LUZ_STACK_SIZE];
    LocalStats st = {0, 0, 0};
    uint32_t det[40]; // STATS: per-class detail, see DeviceStats
    if (STATS)
        for (int k = 0; k < 40; k++) det[k] = 0;
    uint32_t n_rays = 0, n_occl = 0;
    const uint32_t counted = (r >= a.count_row_begin && r < a.count_row_end) ? 1u : 0u; // halo rows are recomputation
    BitWriter bits; // the shadow bits of all lights in light order, then (restarted) the AO bits
    bits.words = a.shadow_mask + pix * a.shadow_words;

    const int n_sources = fc.num_lights + 1;
    for (int base = 0; base < n_sources; base += kLightChunk) {
        const int chunk = min(kLightChunk, n_sources - base);
        const int chunk_lights = min(chunk, fc.num_lights - base);
        __syncthreads();
        for (int k = threadIdx.x; k < chunk_lights * 4; k += blockDim.x)
            reinterpret_cast<float4*>(s_lights)[k] = __ldg(reinterpret_cast<const float4*>(a.lights + base) + k);
        __syncthreads();
        if (!lit) continue;
        for (int li = 0; li < chunk; li++) {
            const bool is_ao = base + li == fc.num_lights;
            float3 O, T, B, C; // ray origin and sampling frame
            float radius = 0.0f, tMinRay, tMaxRay;
            int n_samples;
            int n_cand = -1; // < 0: rays descend from the TLAS root
            bool hinted = false;
            if (is_ao) { // TraceAORays (light.frag:111-135)
                bits.flush();
                bits.words = a.ao_mask + pix * a.ao_words;
                bits.bit = 0;
                O = fragPos + N * (camDist * 0.01f);
                T = fabsf(N.z) > 0.5f ? f3(0.0f, -N.z, N.y) : f3(-N.y, N.x, 0.0f);
                B = cross3(N, T);
                C = N;
                tMinRay = fc.ao_min;
                tMaxRay = fc.ao_max;
                n_samples = fc.ao_num_samples;
                if (n_samples >= kMinCandSamples) {
                    // every AO ray of this pixel stays inside O +- aoMax * |dir| per axis (|dir_k| <= |(T_k, B_k, N_k)|):
                    // one TLAS walk with that box replaces the TLAS levels of all aoNumSamples rays
                    const float m = fabsf(tMaxRay) * 1.001f;
                    if (LUZ_AO_HEMISPHERE && tMinRay >= 0.0f && tMaxRay >= 0.0f) {
                        // the rays only leave towards the hemisphere of C = N: hemisphere reach box (traverse.cuh),
                        // then the same argument in the object space of every candidate against its BLAS root
                        float3 lo, hi;
                        hemisphere_box(O, T, B, C, m, f3(2e-6f * fabsf(O.x) + 1e-6f, 2e-6f * fabsf(O.y) + 1e-6f, 2e-6f * fabsf(O.z) + 1e-6f), lo, hi);
                        const uint32_t nodes0 = st.nodes;
                        n_cand = collect_instances<STATS>(a.scene, lo, hi, s_cand, 128, kMaxCand, stack, &st);
                        if (n_cand > 0) n_cand = filter_candidates<STATS>(a.scene, O, T, B, C, m, s_cand, 128, n_cand, &st);
                        if (STATS && counted) {
                            det[32]++;
                            det[33] += n_cand == 0;
                            det[34] += n_cand < 0;
                            det[35] += (uint32_t)max(n_cand, 0);
                            det[36] += st.nodes - nodes0;
                        }
                    } else {
                        const float3 ext = f3(m * sqrtf(T.x * T.x + B.x * B.x + C.x * C.x) + 1e-6f,
                                              m * sqrtf(T.y * T.y + B.y * B.y + C.y * C.y) + 1e-6f,
                                              m * sqrtf(T.z * T.z + B.z * B.z + C.z * C.z) + 1e-6f);
                        n_cand = collect_instances<STATS>(a.scene, O - ext, O + ext, s_cand, 128, kMaxCand, stack, &st);
                    }
                }
            } else {
                const LightRec L4 = s_lights[li];
                n_samples = fc.shadow_type == LUZW_SHADOW_RAYTRACING ? L4.num_shadow_samples : 0;
                if (n_samples <= 0) continue; // no rays, no bits (light.frag:87-89)
                radius = L4.radius;
                shadow_ray_frame(L4, fragPos, N, camDist, O, C);
                if (a.hints) { // the tile's occluder hint for this light is tried first (see k_shadow_hints)
                    const uint32_t hint = __ldg(a.hints + hint_tile_index(fc, a.rows, a.hint_sx, a.hint_sy, band, x, r) * (uint32_t)fc.num_lights + (uint32_t)(base + li));
                    if (hint != kNoInstance) {
                        s_cand[0] = hint;
                        n_cand = 1;
                        hinted = true;
                    }
                }
                T = rg_normalize(cross3(C, f3(0.0f, 1.0f, 0.0f)));
                B = rg_normalize(cross3(T, C));
                tMinRay = 0.001f;
                tMaxRay = rg_length(C);
            }
            if (n_cand == 0) { // no instance within reach of any AO ray of this pixel: every one of them misses
                n_rays += counted * (uint32_t)max(n_samples, 0);
                continue;
            }
            for (int i = 0; i < n_samples; i++) {
                const float2 rng = blue_noise_sample(bn_r, bn_g, i, fc.frame_mod);
                float sn, cs;
                float3 dir;
                if (is_ao) { // HemisphereSample (light.frag:63-69)
                    const float rr = rg_sqrt(rng.x);
                    rg_sincos(6.283f * rng.y, &sn, &cs);
                    dir = rg_combine(T, rr * cs, B, rr * sn, C, rg_sqrt(fmaxf(0.0f, 1.0f - rng.x)));
                } else { // DiskSample (light.frag:57-61)
                    const float pointRadius = radius * rg_sqrt(rng.x);
                    rg_sincos(rng.y * 2.0f * kPI, &sn, &cs);
                    dir = rg_normalize(rg_combine(T, pointRadius * cs, B, pointRadius * sn, C, 1.0f));
                }
                n_rays += counted;
                LocalStats rst = {0, 0, 0};
                const bool hit = trace_ray<false, STATS, false, false>(a.scene, O, dir, tMinRay, tMaxRay, nullptr, STATS ? &rst : &st, stack, s_cand, 128, n_cand, hinted);
                if (STATS) {
                    st.nodes += rst.nodes, st.tris += rst.tris, st.insts += rst.insts;
                    if (counted) {
                        uint32_t* d8 = det + 8 * (is_ao ? (n_cand > 0 ? 2 : 3) : (hinted ? 0 : 1));
                        d8[0]++;
                        d8[1] += hit;
                        d8[2] += rst.tlas_nodes;
                        d8[3] += rst.nodes - rst.tlas_nodes;
                        d8[4] += rst.tris;
                        d8[5] += rst.insts;
                        d8[6] += rst.root_descents;
                    }
                }
                n_occl += hit ? counted : 0u;
                bits.push(hit);
            }
        }
    }
    if (lit) bits.flush();

    if (STATS) {
        unsigned long long vals[5] = {n_rays, st.nodes, st.tris, st.insts, n_occl};
#pragma unroll
        for (int k = 0; k < 5; k++) {
            unsigned long long vsum = vals[k];
            for (int off = 16; off; off >>= 1) vsum += __shfl_xor_sync(0xFFFFFFFFu, vsum, off);
            vals[k] = vsum;
        }
        if (lane == 0) {
            atomicAdd(&a.stats->rays, vals[0]);
            atomicAdd(&a.stats->nodes, vals[1]);
            atomicAdd(&a.stats->tris, vals[2]);
            atomicAdd(&a.stats->insts, vals[3]);
            atomicAdd(&a.stats->occluded, vals[4]);
        }
        for (int k = 0; k < 37; k++) {
            unsigned long long vsum = det[k];
            for (int off = 16; off; off >>= 1) vsum += __shfl_xor_sync(0xFFFFFFFFu, vsum, off);
            if (lane == 0 && vsum) atomicAdd(&a.stats->detail[k], vsum);
        }
    }
}

template <bool STATS, int MIN_BLOCKS>
__global__ void __launch_bounds__(128, MIN_BLOCKS) k_light_rays(const LightArgs a) {
    light_rays_body<STATS>(a, blockIdx.z);
}

// ---- the ray kernels a host gets: persistent warps ---------------------------------------------------------------------
// The same per-lane ray loop as above inside a tile loop: a warp pulls 8x4-pixel tiles from an atomic counter until none
// is left (a launch has no tail of half-empty waves, whatever share of the frame a rank owns) and clears the mask words of
// its tile itself (no memset pass over the frame).  Lights are read through the read-only cache at a warp-uniform address
// (no staging, no barriers).  What was tried on top of this and lost is recorded in profiles/r2_ray_queue_experiment.md:
// ballot-compacted ray queues, shadow rays as warp packets with one shared stack, both kinds of ray in one body.
// PART 0: the shadow rays, 1: the AO rays (two launches, each body specialised at compile time); -1: both in one launch.
template <bool ONE_VISIT, int MIN_BLOCKS, int PART>
__global__ void __launch_bounds__(128, MIN_BLOCKS) k_light_rays_persistent(const LightArgs a, uint32_t* __restrict__ tile_counter,
                                                                           const uint32_t tiles_x, const uint32_t tiles_y,
                                                                           const uint32_t n_tiles) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const FrameConst& fc = a.fc;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t* s_cand = reinterpret_cast<uint32_t*>(smem_raw) + warp * (kMaxCand * 32) + lane; // candidate lists [k][lane]
    const float3 camPos = f3(fc.cam_pos[0], fc.cam_pos[1], fc.cam_pos[2]);
    const bool shadows = fc.shadow_type == LUZW_SHADOW_RAYTRACING && fc.num_lights > 0;
    uint2 stack[LUZ_STACK_SIZE];

    while (true) {
        uint32_t t = 0;
        if (lane == 0) t = atomicAdd(tile_counter, 1u);
        t = __shfl_sync(0xFFFFFFFFu, t, 0);
        if (t >= n_tiles) break;
        const uint32_t bx = t % tiles_x, by = (t / tiles_x) % tiles_y, band = t / (tiles_x * tiles_y);
        const uint32_t x = bx * 8u + (lane & 7), r = by * 4u + (lane >> 3);
        const bool in_image = x < fc.width && r < a.rows.rows;
        const uint32_t y = in_image ? band_row(fc, a.rows, band, r) : 0u;
        const size_t pix = (size_t)y * fc.width + x;
        float3 N = f3(0.0f, 0.0f, 0.0f);
        float depth = 1.0f;
        uchar4 bn8 = make_uchar4(0, 0, 0, 0);
        if (in_image) {
            const float4 n4 = __ldg(a.normal + pix);
            N = f3(n4.x, n4.y, n4.z);
            depth = __ldg(a.depth + pix);
            bn8 = __ldg(a.blue_noise + (size_t)(y % fc.bn_h) * fc.bn_w + (x % fc.bn_w));
            if (PART != 1)
                for (uint32_t w = 0; w < a.shadow_words; w++) a.shadow_mask[pix * a.shadow_words + w] = 0u;
            if (PART != 0)
                for (uint32_t w = 0; w < a.ao_words; w++) a.ao_mask[pix * a.ao_words + w] = 0u;
        }
        const bool lit = in_image && (length3(N) != 0.0f); // light.frag:178
        if (lit) {
            const float u = ((float)x + 0.5f) / (float)fc.width, v = ((float)y + 0.5f) / (float)fc.height;
            const float3 fragPos = depth_to_world(fc, u, v, depth);
            const float camDist = length3(fragPos - camPos);
            const float bn_r = (float)bn8.x / 255.0f, bn_g = (float)bn8.y / 255.0f;
            const size_t hint_base = hint_tile_index(fc, a.rows, a.hint_sx, a.hint_sy, band, bx * 8u, by * 4u) * (uint32_t)fc.num_lights;
            BitWriter bits;
            bits.words = a.shadow_mask + pix * a.shadow_words;
            // one loop over the ray sources (the lights, then AO) so that the kernel holds one inlined copy of the traversal
            const int li_first = (PART == 1 || !shadows) ? fc.num_lights : 0;
            const int li_last = PART == 0 ? fc.num_lights - 1 : fc.num_lights;
            for (int li = li_first; li <= li_last; li++) {
                const bool is_ao = PART == 1 || (PART < 0 && li == fc.num_lights);
                float3 O, T, B, C;
                float radius = 0.0f, tMinRay, tMaxRay;
                int n_samples, n_cand = -1;
                bool hinted = false;
                if (is_ao) { // TraceAORays (light.frag:111-135)
                    bits.flush();
                    n_samples = fc.ao_num_samples;
                    if (n_samples <= 0) break;
                    bits.words = a.ao_mask + pix * a.ao_words;
                    bits.bit = 0;
                    O = fragPos + N * (camDist * 0.01f);
                    T = fabsf(N.z) > 0.5f ? f3(0.0f, -N.z, N.y) : f3(-N.y, N.x, 0.0f);
                    B = cross3(N, T);
                    C = N;
                    tMinRay = fc.ao_min;
                    tMaxRay = fc.ao_max;
                    if (n_samples >= kMinCandSamples) {
                        const float m = fabsf(tMaxRay) * 1.001f;
                        if (LUZ_AO_HEMISPHERE && tMinRay >= 0.0f && tMaxRay >= 0.0f) {
                            float3 lo, hi;
                            hemisphere_box(O, T, B, C, m, f3(2e-6f * fabsf(O.x) + 1e-6f, 2e-6f * fabsf(O.y) + 1e-6f, 2e-6f * fabsf(O.z) + 1e-6f), lo, hi);
                            n_cand = collect_instances<false>(a.scene, lo, hi, s_cand, 32, kMaxCand, stack, nullptr);
                            if (n_cand > 0) n_cand = filter_candidates<false>(a.scene, O, T, B, C, m, s_cand, 32, n_cand, nullptr);
                        } else {
                            const float3 ext = f3(m * sqrtf(T.x * T.x + B.x * B.x + C.x * C.x) + 1e-6f,
                                                  m * sqrtf(T.y * T.y + B.y * B.y + C.y * C.y) + 1e-6f,
                                                  m * sqrtf(T.z * T.z + B.z * B.z + C.z * C.z) + 1e-6f);
                            n_cand = collect_instances<false>(a.scene, O - ext, O + ext, s_cand, 32, kMaxCand, stack, nullptr);
                        }
                    }
                    if (n_cand == 0) break; // no instance within reach of any AO ray of this pixel: every one of them misses
                } else { // EvaluateShadow + TraceShadowRay (light.frag:137-146, :86-109)
                    const float4* lp = reinterpret_cast<const float4*>(a.lights + li); // same address in every lane
                    LightRec L4;
                    reinterpret_cast<float4*>(&L4)[0] = __ldg(lp + 0);
                    reinterpret_cast<float4*>(&L4)[1] = __ldg(lp + 1);
                    reinterpret_cast<float4*>(&L4)[2] = __ldg(lp + 2);
                    reinterpret_cast<float4*>(&L4)[3] = __ldg(lp + 3);
                    n_samples = L4.num_shadow_samples;
                    if (n_samples <= 0) continue; // no rays, no bits (light.frag:87-89)
                    radius = L4.radius;
                    shadow_ray_frame(L4, fragPos, N, camDist, O, C);
                    if (a.hints) { // the tile's occluder hint for this light is tried first (see k_shadow_hints)
                        const uint32_t hint = __ldg(a.hints + hint_base + (uint32_t)li);
                        if (hint != kNoInstance) {
                            s_cand[0] = hint;
                            n_cand = 1;
                            hinted = true;
                        }
                    }
                    T = rg_normalize(cross3(C, f3(0.0f, 1.0f, 0.0f)));
                    B = rg_normalize(cross3(T, C));
                    tMinRay = 0.001f;
                    tMaxRay = rg_length(C);
                }
                for (int i = 0; i < n_samples; i++) {
                    const float2 rng = blue_noise_sample(bn_r, bn_g, i, fc.frame_mod);
                    float sn, cs;
                    float3 dir;
                    if (is_ao) { // HemisphereSample (light.frag:63-69)
                        const float rr = rg_sqrt(rng.x);
                        rg_sincos(6.283f * rng.y, &sn, &cs);
                        dir = rg_combine(T, rr * cs, B, rr * sn, C, rg_sqrt(fmaxf(0.0f, 1.0f - rng.x)));
                    } else { // DiskSample (light.frag:57-61)
                        const float pointRadius = radius * rg_sqrt(rng.x);
                        rg_sincos(rng.y * 2.0f * kPI, &sn, &cs);
                        dir = rg_normalize(rg_combine(T, pointRadius * cs, B, pointRadius * sn, C, 1.0f));
                    }
                    bits.push(trace_ray<false, false, false, ONE_VISIT>(a.scene, O, dir, tMinRay, tMaxRay, nullptr, nullptr, stack, s_cand, 32,
                                                                        n_cand, hinted));
                }
            }
            bits.flush();
        }
        __syncwarp();
    }
}

// ---- AO rays, candidate-major ----------------------------------------------------------------------------------------------
// The AO rays of a pixel leave one origin and visit the same short candidate list (usually one instance).  Looping
// candidates outside and samples inside loads the instance record and transforms the origin once per pixel and
// candidate instead of once per ray, and the inner traversal is trace_blas_any -- one BLAS, no TLAS / candidate state
// machine, fewer live registers, fewer phases for the lanes of a warp to disagree on.  The per-ray arithmetic (world
// direction, world-box pre-test, object-space direction, node and triangle tests) is what trace_ray's candidate mode
// does, so the bits are the same.  Pixels whose list overflowed (n_cand < 0) descend from the TLAS root with trace_ray.
template <int MIN_BLOCKS>
__global__ void __launch_bounds__(128, MIN_BLOCKS) k_ao_rays_persistent(const LightArgs a, uint32_t* __restrict__ tile_counter,
                                                                        const uint32_t tiles_x, const uint32_t tiles_y,
                                                                        const uint32_t n_tiles) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const FrameConst& fc = a.fc;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t* s_cand = reinterpret_cast<uint32_t*>(smem_raw) + warp * (kMaxCand * 32) + lane; // candidate lists [k][lane]
    const float3 camPos = f3(fc.cam_pos[0], fc.cam_pos[1], fc.cam_pos[2]);
    const int n_samples = fc.ao_num_samples; // <= 64 (two mask words in registers); more samples take the generic kernel
    uint2 stack[LUZ_STACK_SIZE];

    while (true) {
        uint32_t t = 0;
        if (lane == 0) t = atomicAdd(tile_counter, 1u);
        t = __shfl_sync(0xFFFFFFFFu, t, 0);
        if (t >= n_tiles) break;
        const uint32_t bx = t % tiles_x, by = (t / tiles_x) % tiles_y, band = t / (tiles_x * tiles_y);
        const uint32_t x = bx * 8u + (lane & 7), r = by * 4u + (lane >> 3);
        const bool in_image = x < fc.width && r < a.rows.rows;
        const uint32_t y = in_image ? band_row(fc, a.rows, band, r) : 0u;
        const size_t pix = (size_t)y * fc.width + x;
        float3 N = f3(0.0f, 0.0f, 0.0f);
        float depth = 1.0f;
        uchar4 bn8 = make_uchar4(0, 0, 0, 0);
        if (in_image) {
            const float4 n4 = __ldg(a.normal + pix);
            N = f3(n4.x, n4.y, n4.z);
            depth = __ldg(a.depth + pix);
            bn8 = __ldg(a.blue_noise + (size_t)(y % fc.bn_h) * fc.bn_w + (x % fc.bn_w));
        }
        const bool lit = in_image && (length3(N) != 0.0f); // light.frag:178
        uint32_t occl0 = 0u, occl1 = 0u; // bit i: AO ray i occluded
        if (lit) { // TraceAORays (light.frag:111-135)
            const float u = ((float)x + 0.5f) / (float)fc.width, v = ((float)y + 0.5f) / (float)fc.height;
            const float3 fragPos = depth_to_world(fc, u, v, depth);
            const float camDist = length3(fragPos - camPos);
            const float bn_r = (float)bn8.x / 255.0f, bn_g = (float)bn8.y / 255.0f;
            const float3 O = fragPos + N * (camDist * 0.01f);
            const float3 T = fabsf(N.z) > 0.5f ? f3(0.0f, -N.z, N.y) : f3(-N.y, N.x, 0.0f);
            const float3 B = cross3(N, T);
            const float tMinRay = fc.ao_min, tMaxRay = fc.ao_max;
            int n_cand = -1; // < 0: rays descend from the TLAS root
            if (n_samples >= kMinCandSamples) {
                const float m = fabsf(tMaxRay) * 1.001f;
                if (LUZ_AO_HEMISPHERE && tMinRay >= 0.0f && tMaxRay >= 0.0f) {
                    float3 lo, hi;
                    hemisphere_box(O, T, B, N, m, f3(2e-6f * fabsf(O.x) + 1e-6f, 2e-6f * fabsf(O.y) + 1e-6f, 2e-6f * fabsf(O.z) + 1e-6f), lo, hi);
                    n_cand = collect_instances<false>(a.scene, lo, hi, s_cand, 32, kMaxCand, stack, nullptr);
                    if (n_cand > 0) n_cand = filter_candidates<false>(a.scene, O, T, B, N, m, s_cand, 32, n_cand, nullptr);
                } else {
                    const float3 ext = f3(m * sqrtf(T.x * T.x + B.x * B.x + N.x * N.x) + 1e-6f,
                                          m * sqrtf(T.y * T.y + B.y * B.y + N.y * N.y) + 1e-6f,
                                          m * sqrtf(T.z * T.z + B.z * B.z + N.z * N.z) + 1e-6f);
                    n_cand = collect_instances<false>(a.scene, O - ext, O + ext, s_cand, 32, kMaxCand, stack, nullptr);
                }
            }
            const bool o_ok = O.x == O.x && O.y == O.y && O.z == O.z && tMinRay == tMinRay && tMaxRay == tMaxRay;
            auto ao_dir = [&](const int i) -> float3 { // HemisphereSample (light.frag:63-69)
                const float2 rng = blue_noise_sample(bn_r, bn_g, i, fc.frame_mod);
                const float rr = rg_sqrt(rng.x);
                float sn, cs;
                rg_sincos(6.283f * rng.y, &sn, &cs);
                return rg_combine(T, rr * cs, B, rr * sn, N, rg_sqrt(fmaxf(0.0f, 1.0f - rng.x)));
            };
            if (n_cand < 0) {
                for (int i = 0; i < n_samples; i++)
                    if (trace_ray<false, false, false, true>(a.scene, O, ao_dir(i), tMinRay, tMaxRay, nullptr, nullptr, stack)) {
                        if (i < 32) occl0 |= 1u << i; else occl1 |= 1u << (i - 32);
                    }
            } else if (o_ok) {
                for (int c = 0; c < n_cand; c++) {
                    const uint32_t id = s_cand[c * 32];
                    const InstanceRec* rec = a.scene.instances + id;
                    const float4 r0 = __ldg(&rec->r0), r1 = __ldg(&rec->r1), r2 = __ldg(&rec->r2);
                    const uint4 ptrs = __ldg(reinterpret_cast<const uint4*>(&rec->nodes));
                    const WideNode* nodes = reinterpret_cast<const WideNode*>(((unsigned long long)ptrs.y << 32) | ptrs.x);
                    const WideTri* tris = reinterpret_cast<const WideTri*>(((unsigned long long)ptrs.w << 32) | ptrs.z);
                    const float4 blo = __ldg(a.scene.inst_boxes + 2 * id), bhi = __ldg(a.scene.inst_boxes + 2 * id + 1);
                    const float3 o = xform_point(r0, r1, r2, O);
                    const bool oo_ok = o.x == o.x && o.y == o.y && o.z == o.z;
                    for (int i = 0; i < n_samples; i++) {
                        if ((i < 32 ? occl0 >> i : occl1 >> (i - 32)) & 1u) continue; // an earlier candidate occludes it
                        const float3 wd = ao_dir(i);
                        if (!(wd.x == wd.x && wd.y == wd.y && wd.z == wd.z) || (wd.x == 0.0f && wd.y == 0.0f && wd.z == 0.0f)) continue;
                        // the ray segment against the instance's world box (trace_ray's candidate pre-test)
                        const float3 widir = f3(safe_rcp(wd.x), safe_rcp(wd.y), safe_rcp(wd.z));
                        const float tx0 = (blo.x - O.x) * widir.x, tx1 = (bhi.x - O.x) * widir.x;
                        const float ty0 = (blo.y - O.y) * widir.y, ty1 = (bhi.y - O.y) * widir.y;
                        const float tz0 = (blo.z - O.z) * widir.z, tz1 = (bhi.z - O.z) * widir.z;
                        const float tn = fmaxf(fmaxf(fminf(tx0, tx1), fminf(ty0, ty1)), fmaxf(fminf(tz0, tz1), tMinRay));
                        const float tf = fminf(fminf(fmaxf(tx0, tx1), fmaxf(ty0, ty1)), fminf(fmaxf(tz0, tz1), tMaxRay));
                        if (!(tn - tf <= 2e-6f * fmaxf(fabsf(tn), fabsf(tf)))) continue;
                        const float3 d = xform_dir(r0, r1, r2, wd);
                        if (!oo_ok || !(d.x == d.x && d.y == d.y && d.z == d.z) || (d.x == 0.0f && d.y == 0.0f && d.z == 0.0f)) continue;
                        if (trace_blas_any<true>(nodes, tris, o, d, tMinRay, tMaxRay, stack)) {
                            if (i < 32) occl0 |= 1u << i; else occl1 |= 1u << (i - 32);
                        }
                    }
                }
            }
        }
        if (in_image) { // every pixel of the tile gets its words (zero for background and unoccluded pixels): no clear pass
            a.ao_mask[pix * a.ao_words] = occl0;
            if (a.ao_words > 1) a.ao_mask[pix * a.ao_words + 1] = occl1;
        }
        __syncwarp();
    }
}

// ---- AO rays, (pixel, candidate) pairs compacted over the warp ---------------------------------------------------------------
// 71 % of C3's pixels have an empty candidate list; in the kernel above their lanes idle through the candidate and sample
// loops of the 9 that do not (ncu: 8.8 active lanes in the node test, 2.6 in the triangle test).  Here the per-pixel part
// (G-buffer load, box query, filter) stays one pixel per lane, but what it produces -- one work item per (pixel,
// candidate instance), each worth ao_num_samples rays against ONE BLAS -- goes into a per-warp ring in shared memory
// (ballot + prefix popcount) and is executed 32 items at a time, one per lane, whichever tile they came from.  An item
// carries what its rays need (origin, normal, blue-noise texel, pixel, instance: 36 bytes) and is worth ~16 rays, so the
// queue traffic that sank the all-rays experiment (profiles/r2_ray_queue_experiment.md) is amortised 16 times.  The
// occlusion bits of a pixel's items are OR-ed into its mask word with atomicOr (the owner lane stored the zero word
// before, ordered by __syncwarp); an item no ray of which is occluded -- most -- touches nothing.  Bits: OR over the
// candidates of the per-candidate any-hit, as above (the "already occluded by an earlier candidate" skip is lost between
// items in flight together; it only saved work).
constexpr int kAoQueueCap = 64; // < 32 queued before a push round, <= 32 pushed by it
struct __align__(16) AoPairQueue { // one per warp
    float4 qa[kAoQueueCap];        // O.xyz, pixel index
    float4 qb[kAoQueueCap];        // N.xyz, instance
    uint32_t qc[kAoQueueCap];      // blue-noise texel: r | g << 8
    uint32_t cand[kMaxCand * 32];  // candidate lists [k][lane]
    float4 ra[kAoQueueCap];        // a ray: object-space origin, instance
    float4 rb[kAoQueueCap];        // object-space direction, pixel | sample << 26
};

template <int MIN_BLOCKS>
__global__ void __launch_bounds__(128, MIN_BLOCKS) k_ao_rays_compact(const LightArgs a, uint32_t* __restrict__ tile_counter,
                                                                     const uint32_t tiles_x, const uint32_t tiles_y,
                                                                     const uint32_t n_tiles) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const FrameConst& fc = a.fc;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    AoPairQueue& ws = reinterpret_cast<AoPairQueue*>(smem_raw)[warp];
    uint32_t* s_cand = ws.cand + lane;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const float3 camPos = f3(fc.cam_pos[0], fc.cam_pos[1], fc.cam_pos[2]);
    const int n_samples = fc.ao_num_samples; // <= 64 (two mask words)
    const float tMinRay = fc.ao_min, tMaxRay = fc.ao_max;
    uint2 stack[LUZ_STACK_SIZE];
    uint32_t q_head = 0, q_count = 0, r_head = 0, r_count = 0; // warp-uniform

    auto sample_dir = [&](const float3& T, const float3& B, const float3& N, const float bn_r, const float bn_g, const int i) -> float3 {
        const float2 rng = blue_noise_sample(bn_r, bn_g, i, fc.frame_mod); // HemisphereSample (light.frag:63-69)
        const float rr = rg_sqrt(rng.x);
        float sn, cs;
        rg_sincos(6.283f * rng.y, &sn, &cs);
        return rg_combine(T, rr * cs, B, rr * sn, N, rg_sqrt(fmaxf(0.0f, 1.0f - rng.x)));
    };
    // rays that passed the world-box pre-test: traced 32 at a time, whichever item they came from
    auto trace_rays = [&](const uint32_t n) {
        if ((uint32_t)lane < n) {
            const uint32_t e = (r_head + (uint32_t)lane) & (kAoQueueCap - 1);
            const float4 ra = ws.ra[e], rb = ws.rb[e];
            const uint32_t id = __float_as_uint(ra.w), meta = __float_as_uint(rb.w);
            const uint4 ptrs = __ldg(reinterpret_cast<const uint4*>(&a.scene.instances[id].nodes));
            const WideNode* nodes = reinterpret_cast<const WideNode*>(((unsigned long long)ptrs.y << 32) | ptrs.x);
            const WideTri* tris = reinterpret_cast<const WideTri*>(((unsigned long long)ptrs.w << 32) | ptrs.z);
            if (trace_blas_any<true>(nodes, tris, f3(ra.x, ra.y, ra.z), f3(rb.x, rb.y, rb.z), tMinRay, tMaxRay, stack)) {
                const uint32_t bit = meta >> 26;
                atomicOr(a.ao_mask + (size_t)(meta & 0x3FFFFFFu) * a.ao_words + (bit >> 5), 1u << (bit & 31u));
            }
        }
        __syncwarp();
        r_head = (r_head + n) & (kAoQueueCap - 1);
        r_count -= n;
    };
    auto drain = [&](const uint32_t n) { // the n (<= 32) oldest items, one per lane: generate their rays
        const bool have = (uint32_t)lane < n;
        const uint32_t e = (q_head + (uint32_t)(have ? lane : 0)) & (kAoQueueCap - 1);
        const float4 qa = ws.qa[e], qb = ws.qb[e];
        const uint32_t bn = ws.qc[e], id = __float_as_uint(qb.w);
        const uint32_t pix = __float_as_uint(qa.w);
        const float3 O = f3(qa.x, qa.y, qa.z), N = f3(qb.x, qb.y, qb.z);
        const float3 T = fabsf(N.z) > 0.5f ? f3(0.0f, -N.z, N.y) : f3(-N.y, N.x, 0.0f);
        const float3 B = cross3(N, T);
        const float bn_r = (float)(bn & 255u) / 255.0f, bn_g = (float)(bn >> 8) / 255.0f;
        const InstanceRec* rec = a.scene.instances + id;
        const float4 r0 = __ldg(&rec->r0), r1 = __ldg(&rec->r1), r2 = __ldg(&rec->r2);
        const float4 blo = __ldg(a.scene.inst_boxes + 2 * id), bhi = __ldg(a.scene.inst_boxes + 2 * id + 1);
        const float3 o = xform_point(r0, r1, r2, O);
        const bool oo_ok = have && o.x == o.x && o.y == o.y && o.z == o.z;
        for (int i = 0; i < n_samples; i++) {
            const float3 wd = sample_dir(T, B, N, bn_r, bn_g, i);
            bool fire = oo_ok && (wd.x == wd.x && wd.y == wd.y && wd.z == wd.z) && !(wd.x == 0.0f && wd.y == 0.0f && wd.z == 0.0f);
            // the ray segment against the instance's world box (trace_ray's candidate pre-test)
            const float3 widir = f3(safe_rcp(wd.x), safe_rcp(wd.y), safe_rcp(wd.z));
            const float tx0 = (blo.x - O.x) * widir.x, tx1 = (bhi.x - O.x) * widir.x;
            const float ty0 = (blo.y - O.y) * widir.y, ty1 = (bhi.y - O.y) * widir.y;
            const float tz0 = (blo.z - O.z) * widir.z, tz1 = (bhi.z - O.z) * widir.z;
            const float tn = fmaxf(fmaxf(fminf(tx0, tx1), fminf(ty0, ty1)), fmaxf(fminf(tz0, tz1), tMinRay));
            const float tf = fminf(fminf(fmaxf(tx0, tx1), fmaxf(ty0, ty1)), fminf(fmaxf(tz0, tz1), tMaxRay));
            fire = fire && (tn - tf <= 2e-6f * fmaxf(fabsf(tn), fabsf(tf)));
            const float3 d = xform_dir(r0, r1, r2, wd);
            fire = fire && (d.x == d.x && d.y == d.y && d.z == d.z) && !(d.x == 0.0f && d.y == 0.0f && d.z == 0.0f);
            const uint32_t bl = __ballot_sync(0xFFFFFFFFu, fire);
            if (fire) {
                const uint32_t k = (r_head + r_count + __popc(bl & lt_mask)) & (kAoQueueCap - 1);
                ws.ra[k] = make_float4(o.x, o.y, o.z, __uint_as_float(id));
                ws.rb[k] = make_float4(d.x, d.y, d.z, __uint_as_float(pix | ((uint32_t)i << 26)));
            }
            r_count += __popc(bl);
            __syncwarp();
            if (r_count >= 32) trace_rays(32);
        }
        q_head = (q_head + n) & (kAoQueueCap - 1);
        q_count -= n;
    };

    while (true) {
        uint32_t t = 0;
        if (lane == 0) t = atomicAdd(tile_counter, 1u);
        t = __shfl_sync(0xFFFFFFFFu, t, 0);
        if (t >= n_tiles) break;
        const uint32_t bx = t % tiles_x, by = (t / tiles_x) % tiles_y, band = t / (tiles_x * tiles_y);
        const uint32_t x = bx * 8u + (lane & 7), r = by * 4u + (lane >> 3);
        const bool in_image = x < fc.width && r < a.rows.rows;
        const uint32_t y = in_image ? band_row(fc, a.rows, band, r) : 0u;
        const size_t pix = (size_t)y * fc.width + x;
        float3 N = f3(0.0f, 0.0f, 0.0f);
        float depth = 1.0f;
        uchar4 bn8 = make_uchar4(0, 0, 0, 0);
        if (in_image) {
            const float4 n4 = __ldg(a.normal + pix);
            N = f3(n4.x, n4.y, n4.z);
            depth = __ldg(a.depth + pix);
            bn8 = __ldg(a.blue_noise + (size_t)(y % fc.bn_h) * fc.bn_w + (x % fc.bn_w));
        }
        const bool lit = in_image && (length3(N) != 0.0f); // light.frag:178
        uint32_t occl0 = 0u, occl1 = 0u; // bits found here (pixels whose list overflowed); the items add theirs later
        int n_cand = 0;
        float3 O = f3(0.0f, 0.0f, 0.0f);
        if (lit) { // TraceAORays (light.frag:111-135)
            const float u = ((float)x + 0.5f) / (float)fc.width, v = ((float)y + 0.5f) / (float)fc.height;
            const float3 fragPos = depth_to_world(fc, u, v, depth);
            const float camDist = length3(fragPos - camPos);
            O = fragPos + N * (camDist * 0.01f);
            const float3 T = fabsf(N.z) > 0.5f ? f3(0.0f, -N.z, N.y) : f3(-N.y, N.x, 0.0f);
            const float3 B = cross3(N, T);
            n_cand = -1; // < 0: rays descend from the TLAS root
            if (n_samples >= kMinCandSamples) {
                const float m = fabsf(tMaxRay) * 1.001f;
                if (LUZ_AO_HEMISPHERE && tMinRay >= 0.0f && tMaxRay >= 0.0f) {
                    float3 lo, hi;
                    hemisphere_box(O, T, B, N, m, f3(2e-6f * fabsf(O.x) + 1e-6f, 2e-6f * fabsf(O.y) + 1e-6f, 2e-6f * fabsf(O.z) + 1e-6f), lo, hi);
                    n_cand = collect_instances<false>(a.scene, lo, hi, s_cand, 32, kMaxCand, stack, nullptr);
                    if (n_cand > 0) n_cand = filter_candidates<false>(a.scene, O, T, B, N, m, s_cand, 32, n_cand, nullptr);
                } else {
                    const float3 ext = f3(m * sqrtf(T.x * T.x + B.x * B.x + N.x * N.x) + 1e-6f,
                                          m * sqrtf(T.y * T.y + B.y * B.y + N.y * N.y) + 1e-6f,
                                          m * sqrtf(T.z * T.z + B.z * B.z + N.z * N.z) + 1e-6f);
                    n_cand = collect_instances<false>(a.scene, O - ext, O + ext, s_cand, 32, kMaxCand, stack, nullptr);
                }
            }
            const bool o_ok = O.x == O.x && O.y == O.y && O.z == O.z && tMinRay == tMinRay && tMaxRay == tMaxRay;
            if (n_cand < 0) {
                const float bn_r = (float)bn8.x / 255.0f, bn_g = (float)bn8.y / 255.0f;
                for (int i = 0; i < n_samples; i++)
                    if (trace_ray<false, false, false, true>(a.scene, O, sample_dir(T, B, N, bn_r, bn_g, i), tMinRay, tMaxRay, nullptr, nullptr, stack)) {
                        if (i < 32) occl0 |= 1u << i; else occl1 |= 1u << (i - 32);
                    }
                n_cand = 0;
            } else if (!o_ok) {
                n_cand = 0;
            }
        }
        if (in_image) { // every pixel of the tile gets its words: no clear pass
            a.ao_mask[pix * a.ao_words] = occl0;
            if (a.ao_words > 1) a.ao_mask[pix * a.ao_words + 1] = occl1;
        }
        __syncwarp(); // the words are stored before any item of this tile ORs into them
        const int max_c = __reduce_max_sync(0xFFFFFFFFu, n_cand);
        for (int c = 0; c < max_c; c++) {
            const bool push = c < n_cand;
            const uint32_t b = __ballot_sync(0xFFFFFFFFu, push);
            if (push) {
                const uint32_t e = (q_head + q_count + __popc(b & lt_mask)) & (kAoQueueCap - 1);
                ws.qa[e] = make_float4(O.x, O.y, O.z, __uint_as_float((uint32_t)pix));
                ws.qb[e] = make_float4(N.x, N.y, N.z, __uint_as_float(s_cand[c * 32]));
                ws.qc[e] = (uint32_t)bn8.x | ((uint32_t)bn8.y << 8);
            }
            q_count += __popc(b);
            __syncwarp();
            if (q_count >= 32) drain(32);
        }
    }
    if (q_count) drain(q_count);
    if (r_count) trace_rays(r_count);
}

// ---- shadow rays with per-ray temporal occluder hints ---------------------------------------------------------------------
// 87 % of C3's shadow rays are occluded, and the ray a pixel fires towards a light this frame is -- up to the TAA jitter
// and one step of the blue-noise sequence -- the ray it fired last frame.  The triangle that occluded it then very
// probably occludes it now.  So every shadow ray remembers its last two occluders: ray_hints[bit][pixel] = 2 x (instance,
// triangle) is written by whoever finds a hit, and the next frame's ray tests THOSE triangles first -- an instance
// transform and a triangle test each, ~150 instructions against ~1900 for a hinted traversal, executed by all lit lanes of
// the tile together.  Rays it does not settle go into a per-warp queue in shared memory (ballot + prefix popcount, as in the
// compaction experiment of profiles/r2_ray_queue_experiment.md -- which lost because EVERY ray went through the queue;
// here the queue only sees the few survivors) and are traced 32 at a time: tile hint first, then the TLAS root.
// Exactness: a hint only changes the ORDER in which occluders are tried.  Whatever the hint array holds -- last frame's
// occluders, the occluders of another scene, garbage -- an entry is either out of bounds (skipped) or names a real
// triangle of a real instance, and a hit on it is a hit; a miss falls through to the full traversal.  The bits are those
// of the kernels above on the first frame, on the hundredth, and after the scene has changed under the hints
// (tests/test_gpu_parity.py::test_temporal_hints_do_not_change_visibility).  Used while a pixel has at most 256 shadow
// rays (the masks of a tile are assembled in shared memory) and the hint array fits its budget; else the kernel above.
constexpr int kShadowQueueCap = 64; // < 32 queued before a round, <= 32 pushed by it; a drain pops 32 (nothing is re-queued)
// Two shapes: up to 64 shadow rays per pixel keep the last TWO occluders of every ray (16 bytes per ray); up to 256 (C4's
// 256 lights) keep one (8 bytes per ray: 17 GB at 4K -- HBM is there to be used -- read once per frame at ~3 ms).
template <uint32_t WORDS>
struct __align__(16) ShadowQueue { // one per warp
    float4 q0[kShadowQueueCap];   // O.xyz, tmax
    float4 q1[kShadowQueueCap];   // dir.xyz, meta: owner lane | mask bit << 8
    uint32_t aux[kShadowQueueCap]; // first candidate instance of the traversal (kNoInstance: none)
    uint32_t mask[32 * WORDS];     // the shadow mask words of the tile's 32 pixels
};

template <bool ONE_VISIT, int MIN_BLOCKS, int SLOTS, uint32_t kSmemMaskWords>
__global__ void __launch_bounds__(128, MIN_BLOCKS) k_shadow_rays_temporal(const LightArgs a, uint32_t* __restrict__ tile_counter,
                                                                          const uint32_t tiles_x, const uint32_t tiles_y,
                                                                          const uint32_t n_tiles, uint2* __restrict__ ray_hints /* [bit][pixel][SLOTS] */,
                                                                          const uint32_t n_inst) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const FrameConst& fc = a.fc;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    ShadowQueue<kSmemMaskWords>& ws = reinterpret_cast<ShadowQueue<kSmemMaskWords>*>(smem_raw)[warp];
    const uint32_t lt_mask = (1u << lane) - 1u;
    const float3 camPos = f3(fc.cam_pos[0], fc.cam_pos[1], fc.cam_pos[2]);
    const size_t px_total = (size_t)fc.width * fc.height;
    uint2 stack[LUZ_STACK_SIZE];

    while (true) {
        uint32_t t = 0;
        if (lane == 0) t = atomicAdd(tile_counter, 1u);
        t = __shfl_sync(0xFFFFFFFFu, t, 0);
        if (t >= n_tiles) break;
        const uint32_t bx = t % tiles_x, by = (t / tiles_x) % tiles_y, band = t / (tiles_x * tiles_y);
        const uint32_t x = bx * 8u + (lane & 7), r = by * 4u + (lane >> 3);
        const bool in_image = x < fc.width && r < a.rows.rows;
        const uint32_t y = in_image ? band_row(fc, a.rows, band, r) : 0u;
        const size_t pix = (size_t)y * fc.width + x;
        float3 N = f3(0.0f, 0.0f, 0.0f);
        float depth = 1.0f;
        uchar4 bn8 = make_uchar4(0, 0, 0, 0);
        if (in_image) {
            const float4 n4 = __ldg(a.normal + pix);
            N = f3(n4.x, n4.y, n4.z);
            depth = __ldg(a.depth + pix);
            bn8 = __ldg(a.blue_noise + (size_t)(y % fc.bn_h) * fc.bn_w + (x % fc.bn_w));
        }
        const bool lit = in_image && (length3(N) != 0.0f); // light.frag:178
        if (!__any_sync(0xFFFFFFFFu, lit)) { // nothing to trace: the tile's masks are zero
            if (in_image)
                for (uint32_t w = 0; w < a.shadow_words; w++) a.shadow_mask[pix * a.shadow_words + w] = 0u;
            continue;
        }
#pragma unroll
        for (uint32_t w = 0; w < kSmemMaskWords; w++) ws.mask[lane * kSmemMaskWords + w] = 0u;
        const float u = ((float)x + 0.5f) / (float)fc.width, v = ((float)y + 0.5f) / (float)fc.height;
        const float3 fragPos = depth_to_world(fc, u, v, depth);
        const float camDist = length3(fragPos - camPos);
        const float bn_r = (float)bn8.x / 255.0f, bn_g = (float)bn8.y / 255.0f;
        const size_t hint_base = hint_tile_index(fc, a.rows, a.hint_sx, a.hint_sy, band, bx * 8u, by * 4u) * (uint32_t)fc.num_lights;
        __syncwarp();

        int q_count = 0, light = 0, sample = 0;
        uint32_t bit0 = 0; // first mask bit of the current light
        bool producing = true;
        uint32_t n_fired = 0, n_settled = 0; // this tile's shadow rays / those the temporal hints settled (feeds the host's on / off decision)
        while (true) {
            // ---- produce: one (light, sample) round for every lit lane; the temporal hint settles most of them at once ----
            while (q_count < 32 && producing) {
                if (light >= fc.num_lights) {
                    producing = false;
                    break;
                }
                const float4* lp = reinterpret_cast<const float4*>(a.lights + light); // same address in every lane
                LightRec L4;
                reinterpret_cast<float4*>(&L4)[0] = __ldg(lp + 0);
                reinterpret_cast<float4*>(&L4)[1] = __ldg(lp + 1);
                reinterpret_cast<float4*>(&L4)[2] = __ldg(lp + 2);
                reinterpret_cast<float4*>(&L4)[3] = __ldg(lp + 3);
                const int n_samples = L4.num_shadow_samples;
                if (sample >= n_samples) { // no (more) rays for this light (light.frag:87-89)
                    bit0 += (uint32_t)max(n_samples, 0);
                    light++;
                    sample = 0;
                    continue;
                }
                const uint32_t bit = bit0 + (uint32_t)sample;
                float3 O = f3(0.0f, 0.0f, 0.0f), dir = O;
                float tMaxRay = 0.0f;
                bool pending = false; // this lane's ray is not settled yet
                bool fired = false;   // this lane has a (valid) ray in this round
                uint32_t hint_own = kNoInstance;
                if (lit) { // EvaluateShadow + TraceShadowRay (light.frag:137-146, :86-100)
                    float3 C;
                    shadow_ray_frame(L4, fragPos, N, camDist, O, C);
                    const float3 T = rg_normalize(cross3(C, f3(0.0f, 1.0f, 0.0f)));
                    const float3 B = rg_normalize(cross3(T, C));
                    tMaxRay = rg_length(C);
                    const float2 rng = blue_noise_sample(bn_r, bn_g, sample, fc.frame_mod);
                    const float pointRadius = L4.radius * rg_sqrt(rng.x); // DiskSample (light.frag:57-61)
                    float sn, cs;
                    rg_sincos(rng.y * 2.0f * kPI, &sn, &cs);
                    dir = rg_normalize(rg_combine(T, pointRadius * cs, B, pointRadius * sn, C, 1.0f));
                    // rays with NaNs (the vertical-light tangent of light.frag:90) and null directions miss, as in trace_ray
                    pending = (O.x == O.x && O.y == O.y && O.z == O.z && dir.x == dir.x && dir.y == dir.y && dir.z == dir.z &&
                               tMaxRay == tMaxRay) && !(dir.x == 0.0f && dir.y == 0.0f && dir.z == 0.0f);
                    fired = pending;
                    uint32_t own_inst = kNoInstance; // the instance that occluded this ray last time: first candidate if it has to be traced
                    if (pending) { // the last two occluders of this very ray, most recent first
                        uint4 th = make_uint4(kNoInstance, 0u, kNoInstance, 0u);
                        if (SLOTS == 2) {
                            th = __ldg(reinterpret_cast<const uint4*>(ray_hints) + (size_t)bit * px_total + pix);
                        } else {
                            const uint2 t1 = __ldg(ray_hints + (size_t)bit * px_total + pix);
                            th.x = t1.x, th.y = t1.y;
                        }
                        if (a.count_temporal) atomicAdd(&a.stats->detail[24], 1ull), atomicAdd(&a.stats->detail[25], th.x < n_inst ? 1ull : 0ull);
                        // hints name instances by their index in the host's array, which survives TLAS rebuilds and refits
                        own_inst = th.x < n_inst ? __ldg(a.inst_leaf + th.x) : kNoInstance;
#pragma unroll
                        for (int k = 0; k < SLOTS; k++) {
                            const uint32_t hin = k ? th.z : th.x, ht = k ? th.w : th.y;
                            if (!pending || hin >= n_inst) continue;
                            const uint32_t hi = k ? __ldg(a.inst_leaf + hin) : own_inst;
                            const float4 bhi = __ldg(a.scene.inst_boxes + 2 * (size_t)hi + 1);
                            if (ht >= __float_as_uint(bhi.w)) continue; // not a triangle of that instance's BLAS
                            const InstanceRec* rec = a.scene.instances + hi;
                            const float4 r0 = __ldg(&rec->r0), r1 = __ldg(&rec->r1), r2 = __ldg(&rec->r2);
                            const uint4 ptrs = __ldg(reinterpret_cast<const uint4*>(&rec->nodes));
                            const WideTri* tris = reinterpret_cast<const WideTri*>(((unsigned long long)ptrs.w << 32) | ptrs.z);
                            const float3 o = xform_point(r0, r1, r2, O), d = xform_dir(r0, r1, r2, dir);
                            const TriData q = load_tri(tris + ht);
                            float tt, bu, bv;
                            if (tri_test<false>(q, o, d, cross3_rn(o, d), 0.001f, tMaxRay, tt, bu, bv)) {
                                ws.mask[lane * kSmemMaskWords + (bit >> 5)] |= 1u << (bit & 31u);
                                pending = false;
                                if (a.count_temporal) atomicAdd(&a.stats->detail[26], 1ull);
                            }
                        }
                    }
                    hint_own = own_inst;
                }
                const uint32_t m = __ballot_sync(0xFFFFFFFFu, pending);
                n_fired += __popc(__ballot_sync(0xFFFFFFFFu, fired));
                n_settled += __popc(__ballot_sync(0xFFFFFFFFu, fired && !pending));
                if (m) {
                    const uint32_t hint = a.hints ? __ldg(a.hints + hint_base + (uint32_t)light) : kNoInstance;
                    if (pending) { // first candidate of the traversal: the ray's own last occluder instance, else the tile's hint
                        const int pos = q_count + __popc(m & lt_mask);
                        ws.q0[pos] = make_float4(O.x, O.y, O.z, tMaxRay);
                        ws.q1[pos] = make_float4(dir.x, dir.y, dir.z, __uint_as_float((uint32_t)lane | (bit << 8)));
                        ws.aux[pos] = hint_own != kNoInstance ? hint_own : hint;
                    }
                    q_count += __popc(m);
                }
                sample++;
                __syncwarp();
            }
            if (q_count == 0) break; // nothing queued, nothing left to produce

            // ---- drain one batch of unsettled rays: tile hint first, then the TLAS root; remember the occluder found ----
            const int n = min(q_count, 32);
            q_count -= n;
            if (lane < n) {
                const float4 i0 = ws.q0[q_count + lane], i1 = ws.q1[q_count + lane];
                uint32_t aux = ws.aux[q_count + lane];
                const uint32_t meta = __float_as_uint(i1.w), owner = meta & 31u, bit = meta >> 8;
                HitInfo h;
                h.inst = kNoInstance;
                h.slot = 0u;
                const bool hit = trace_ray<false, false, false, ONE_VISIT>(a.scene, f3(i0.x, i0.y, i0.z), f3(i1.x, i1.y, i1.z), 0.001f, i0.w, &h,
                                                                           nullptr, stack, &aux, 1, aux != kNoInstance ? 1 : -1,
                                                                           aux != kNoInstance);
                if (hit) atomicOr(&ws.mask[owner * kSmemMaskWords + (bit >> 5)], 1u << (bit & 31u));
                if (a.count_temporal) atomicAdd(&a.stats->detail[27], 1ull), atomicAdd(&a.stats->detail[28], hit ? 1ull : 0ull);
                const uint32_t ox = bx * 8u + (owner & 7u), orow = by * 4u + (owner >> 3);
                const size_t opix = (size_t)band_row(fc, a.rows, band, orow) * fc.width + ox;
                if (SLOTS == 2) {
                    uint4* slot = reinterpret_cast<uint4*>(ray_hints) + (size_t)bit * px_total + opix;
                    if (hit) { // the new occluder in front, the previous one behind it
                        const uint4 old = *slot;
                        *slot = make_uint4(__ldg(a.inst_order + h.inst), h.slot, old.x, old.y);
                    } else {
                        *slot = make_uint4(kNoInstance, 0u, kNoInstance, 0u); // unoccluded: nothing to try next frame
                    }
                } else {
                    ray_hints[(size_t)bit * px_total + opix] = hit ? make_uint2(__ldg(a.inst_order + h.inst), h.slot) : make_uint2(kNoInstance, 0u);
                }
            }
            __syncwarp();
        }
        if (in_image)
            for (uint32_t w = 0; w < a.shadow_words; w++) a.shadow_mask[pix * a.shadow_words + w] = ws.mask[lane * kSmemMaskWords + w];
        if (lane == 0 && n_fired) { // 64 counter pairs, 128 bytes apart
            unsigned long long* cnt = a.temporal_counters + 16u * (t & 63u);
            atomicAdd(cnt, (unsigned long long)n_fired);
            atomicAdd(cnt + 1, (unsigned long long)n_settled);
        }
        __syncwarp();
    }
}

__global__ void k_pow22_table(float* __restrict__ t) { t[threadIdx.x] = powf((float)threadIdx.x / 255.0f, 2.2f); } // light.frag:172

} // namespace

cudaError_t launch_pow22_table(cudaStream_t stream, float* table256) {
    k_pow22_table<<<1, 256, 0, stream>>>(table256);
    return cudaGetLastError();
}

// The mask buffers (shadow_words / ao_words words per pixel) are part of the pass, not a debug option.
cudaError_t launch_light_pass(cudaStream_t stream, const LightArgs& args, bool stats, cudaEvent_t rays_done, uint64_t* launches,
                              cudaStream_t aux, cudaEvent_t ev_fork, cudaEvent_t ev_join) {
    if (args.rows.rows == 0 || args.rows.n_bands == 0 || args.fc.width == 0) return cudaSuccess;
    const size_t px = (size_t)args.fc.width * args.fc.height;
    cudaError_t e;
    LightArgs a2 = args;
    a2.cand_offset = (uint32_t)(sizeof(LightRec) * (size_t)max(1, min(args.fc.num_lights, kLightChunk)));
    const size_t smem = a2.cand_offset + sizeof(uint32_t) * kMaxCand * 128;
    static const int kernel_env = [] { // LUZRT_RAY_KERNEL=plain: the per-pixel kernel without statistics (A/B runs)
        const char* e2 = getenv("LUZRT_RAY_KERNEL");
        return (e2 && e2[0] == 'p') ? 1 : 0;
    }();
    const bool any_shadow = args.fc.shadow_type == LUZW_SHADOW_RAYTRACING && args.fc.num_lights > 0;
    const bool any_ao = args.fc.ao_num_samples > 0;
    const bool persistent = !stats && kernel_env == 0 && (any_shadow || any_ao);
    const dim3 grid((args.fc.width + 15) / 16, (args.rows.rows + 7) / 8, args.rows.n_bands); // 16x8 tiles: hints, plain kernel
    if (!(any_shadow && args.hints)) a2.hints = nullptr;
    auto launch_hints = [&](cudaStream_t st) -> cudaError_t { // one hint ray per hint tile and light
        const uint32_t hx = (args.fc.width + (1u << args.hint_sx) - 1u) >> args.hint_sx;
        const uint32_t hy = (args.rows.rows + (1u << args.hint_sy) - 1u) >> args.hint_sy;
        const uint32_t n = hx * hy * args.rows.n_bands * (uint32_t)args.fc.num_lights;
        if (stats)
            k_shadow_hints<true><<<(n + 127) / 128, 128, 0, st>>>(a2, a2.hints, hx, hy);
        else
            k_shadow_hints<false><<<(n + 127) / 128, 128, 0, st>>>(a2, a2.hints, hx, hy);
        ++*launches;
        return cudaGetLastError();
    };
    if (persistent) {
        static const int pminb = [] { // resident CTAs per SM the persistent kernels are compiled for (LUZRT_PERSIST_MINB: tuning runs)
            const char* e2 = getenv("LUZRT_PERSIST_MINB");
            return e2 ? atoi(e2) : 6;
        }();
        static const bool overlap_env = [] { // LUZRT_RAY_OVERLAP=0: hints, shadow rays, AO rays strictly one after the other
            const char* e2 = getenv("LUZRT_RAY_OVERLAP");
            return !(e2 && e2[0] == '0');
        }();
        static int sms = 0;
        if (!sms) {
            int dev = 0;
            if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
            if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
        }
        const uint32_t tx = (args.fc.width + 7) / 8, ty = (args.rows.rows + 3) / 4;
        const uint32_t n_tiles = tx * ty * args.rows.n_bands;
        if ((e = cudaMemsetAsync(args.tile_counter, 0, 2 * sizeof(uint32_t), stream)) != cudaSuccess) return e;
        // a kind of ray the frame does not have leaves its mask untouched: clear it here (the shading kernel does not read
        // it, the debug read-back does)
        if (!any_shadow && (e = cudaMemsetAsync(args.shadow_mask, 0, px * args.shadow_words * 4, stream)) != cudaSuccess) return e;
        if (!any_ao && (e = cudaMemsetAsync(args.ao_mask, 0, px * args.ao_words * 4, stream)) != cudaSuccess) return e;
        const size_t smem_p = 4 * sizeof(uint32_t) * kMaxCand * 32;
        auto launch = [&](auto kern, uint32_t* counter, cudaStream_t st) -> cudaError_t {
            int per_sm = 0;
            cudaError_t e3 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 128, smem_p);
            if (e3 != cudaSuccess) return e3;
            kern<<<min((uint32_t)(sms * max(per_sm, 1)), (n_tiles + 3) / 4), 128, smem_p, st>>>(a2, counter, tx, ty, n_tiles);
            ++*launches;
            return cudaGetLastError();
        };
        // The AO launch needs no hints: it goes first on the pass's stream while the hint pass (a few latency-bound warps)
        // and then the shadow launch run on a second stream -- the hint pass hides behind AO work, and the CTAs of the
        // shadow launch move in as the persistent AO CTAs run out of tiles (and the other way round at its end).
        const bool fork = overlap_env && any_shadow && any_ao && aux && ev_fork && ev_join;
        cudaStream_t shadow_stream = fork ? aux : stream;
        if (fork) {
            if ((e = cudaEventRecord(ev_fork, stream)) != cudaSuccess) return e;
            if ((e = cudaStreamWaitEvent(aux, ev_fork, 0)) != cudaSuccess) return e;
        }
        if (any_shadow) {
            if (a2.hints && (e = launch_hints(shadow_stream)) != cudaSuccess) return e;
            if (args.ray_hints && args.shadow_words <= 8) { // per-ray temporal occluder hints
                auto launch_t = [&](auto kern, size_t smem_t) -> cudaError_t {
                    int per_sm = 0;
                    cudaError_t e3 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 128, smem_t);
                    if (e3 != cudaSuccess) return e3;
                    kern<<<min((uint32_t)(sms * max(per_sm, 1)), (n_tiles + 3) / 4), 128, smem_t, shadow_stream>>>(
                        a2, args.tile_counter, tx, ty, n_tiles, args.ray_hints, args.n_instances);
                    ++*launches;
                    return cudaGetLastError();
                };
                static const int tminb = [] { // resident CTAs per SM (LUZRT_TEMPORAL_MINB: tuning runs)
                    const char* e2 = getenv("LUZRT_TEMPORAL_MINB");
                    return e2 ? atoi(e2) : 6;
                }();
                if (args.shadow_words <= 2)
                    e = tminb == 5 ? launch_t(k_shadow_rays_temporal<true, 5, 2, 2>, 4 * sizeof(ShadowQueue<2>))
                                   : tminb == 7 ? launch_t(k_shadow_rays_temporal<true, 7, 2, 2>, 4 * sizeof(ShadowQueue<2>))
                                                : launch_t(k_shadow_rays_temporal<true, 6, 2, 2>, 4 * sizeof(ShadowQueue<2>));
                else
                    e = launch_t(k_shadow_rays_temporal<true, 6, 1, 8>, 4 * sizeof(ShadowQueue<8>));
            } else {
                e = pminb == 5 ? launch(k_light_rays_persistent<true, 5, 0>, args.tile_counter, shadow_stream)
                               : pminb == 7 ? launch(k_light_rays_persistent<true, 7, 0>, args.tile_counter, shadow_stream)
                                            : launch(k_light_rays_persistent<true, 6, 0>, args.tile_counter, shadow_stream);
            }
            if (e != cudaSuccess) return e;
        }
        if (any_ao) {
            static const bool ao_generic_env = [] { // LUZRT_AO_KERNEL=generic: the per-ray trace_ray loop for AO too (A/B runs)
                const char* e2 = getenv("LUZRT_AO_KERNEL");
                return e2 && e2[0] == 'g';
            }();
            static const int ao_minb = [] { // resident CTAs per SM the AO kernel is compiled for (LUZRT_AO_MINB: tuning runs)
                const char* e2 = getenv("LUZRT_AO_MINB");
                return e2 ? atoi(e2) : 0;
            }();
            static const bool ao_compact_env = [] { // LUZRT_AO_KERNEL=candidate: the kernel without item compaction (A/B runs)
                const char* e2 = getenv("LUZRT_AO_KERNEL");
                return !(e2 && e2[0] == 'c');
            }();
            if (!ao_generic_env && ao_compact_env && args.ao_words <= 2 && px <= 0x4000000ull) { // pixel | sample << 26 in a word
                auto launch_c = [&](auto kern) -> cudaError_t {
                    const size_t smem_c = 4 * sizeof(AoPairQueue);
                    int per_sm = 0;
                    cudaError_t e3 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 128, smem_c);
                    if (e3 != cudaSuccess) return e3;
                    kern<<<min((uint32_t)(sms * max(per_sm, 1)), (n_tiles + 3) / 4), 128, smem_c, stream>>>(a2, args.tile_counter + 1, tx, ty, n_tiles);
                    ++*launches;
                    return cudaGetLastError();
                };
                e = ao_minb == 4 ? launch_c(k_ao_rays_compact<4>) : ao_minb == 5 ? launch_c(k_ao_rays_compact<5>) : launch_c(k_ao_rays_compact<6>);
            } else if (!ao_generic_env && args.ao_words <= 2)
                e = ao_minb == 4 ? launch(k_ao_rays_persistent<4>, args.tile_counter + 1, stream)
                                 : ao_minb == 6 ? launch(k_ao_rays_persistent<6>, args.tile_counter + 1, stream)
                                                : launch(k_ao_rays_persistent<5>, args.tile_counter + 1, stream);
            else
                e = launch(k_light_rays_persistent<true, 6, 1>, args.tile_counter + 1, stream);
            if (e != cudaSuccess) return e;
        }
        if (fork) {
            if ((e = cudaEventRecord(ev_join, aux)) != cudaSuccess) return e;
            if ((e = cudaStreamWaitEvent(stream, ev_join, 0)) != cudaSuccess) return e;
        }
    } else {
        // the per-pixel kernel: statistics variant, LUZRT_RAY_KERNEL=plain, and frames without rays (it then only exists
        // to leave cleared masks behind)
        if ((e = cudaMemsetAsync(args.shadow_mask, 0, px * args.shadow_words * 4, stream)) != cudaSuccess) return e;
        if ((e = cudaMemsetAsync(args.ao_mask, 0, px * args.ao_words * 4, stream)) != cudaSuccess) return e;
        if (a2.hints && (e = launch_hints(stream)) != cudaSuccess) return e;
        if (stats)
            k_light_rays<true, 4><<<grid, 128, smem, stream>>>(a2);
        else
            k_light_rays<false, 6><<<grid, 128, smem, stream>>>(a2);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        ++*launches;
    }
    if (rays_done && (e = cudaEventRecord(rays_done, stream)) != cudaSuccess) return e;
    ++*launches;
    if (!args.exact_math) return launch_light_shade_relaxed(stream, a2);
    const dim3 sgrid((args.fc.width + 31) / 32, (args.rows.rows + 3) / 4, args.rows.n_bands);
    if (args.fc.shadow_type == LUZW_SHADOW_MAP)
        k_light_shade<true, false><<<sgrid, 128, 0, stream>>>(a2);
    else
        k_light_shade<false, false><<<sgrid, 128, 0, stream>>>(a2);
    return cudaGetLastError();
}

} // namespace luz

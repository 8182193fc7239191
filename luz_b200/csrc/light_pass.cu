// light_pass.cu -- the fused deferred lighting kernel: G-buffer decode, Cook-Torrance shading,
// shadow-ray and AO-ray generation, any-hit traversal and accumulation in one launch.  Rays never
// touch HBM: each warp keeps a private ray queue in shared memory.
//
// Restates source/Shaders/light.frag:171-235 (main), :86-109 (TraceShadowRay), :111-135
// (TraceAORays), :137-169 (EvaluateShadow), :57-75 (samplers) and :17-49 (BRDF) of the reference;
// launched where DeferredRenderer::LightPass (DeferredRenderer.cpp:324-345) draws its quad.
//
// Structure (one warp = one 8x4 pixel tile, warps are independent, no CTA barriers):
//   per chunk of at most RMAX rays per pixel
//     GEN    pixel-per-lane, converged: every lit pixel writes its rays of the chunk (direction + id) to
//            the warp's queue; origins and tMax go to small per-pixel tables
//     TRACE  ray-per-lane, persistent: a lane that finishes its ray takes the next one from the queue
//            (the head is a warp-uniform register, so the fetch needs no atomics), which keeps the lanes
//            of the expensive node test full although rays differ widely in length
//     SHADE  pixel-per-lane, converged: hit counts -> shadow / AO factors -> BRDF accumulation, in the
//            reference's light order
#include <cstdlib>

#include "passes.h"
#include "traverse.cuh"

namespace luz {

namespace {

constexpr float kPI = 3.14159265359f;              // LuzCommon.h:11
constexpr float kGoldenRatio = 2.118033988749895f; // LuzCommon.h:12 (sic)
constexpr int kMinCandSamples = 6; // below this the one TLAS walk per pixel does not pay for itself (C2: 4 spp)
constexpr int kMaxCand = 8;        // instances an AO candidate list holds before falling back to the root descent
constexpr int kWarps = 4;
constexpr unsigned kFull = 0xFFFFFFFFu;

__device__ __forceinline__ float fractf(float x) { return x - floorf(x); }

// light.frag:71-75.  The multiply and the add are rounded separately (no FMA) so that the sample
// sequence is bit-identical to the oracle's plain fp32 evaluation.
__device__ __forceinline__ float2 blue_noise_sample(float bn_r, float bn_g, int i, int frame_mod) {
    const float k = (float)(128 * i + frame_mod);
    const float off = __fmul_rn(kGoldenRatio, k);
    return make_float2(fractf(__fadd_rn(bn_r, off)), fractf(__fadd_rn(bn_g, off)));
}

// light.frag:17-26
__device__ __forceinline__ float distribution_ggx(float3 N, float3 H, float roughness) {
    const float a = roughness * roughness;
    const float a2 = a * a;
    const float NdotH = fmaxf(dot3(N, H), 0.0f);
    const float NdotH2 = NdotH * NdotH;
    float denom = (NdotH2 * (a2 - 1.0f) + 1.0f);
    denom = kPI * denom * denom;
    return a2 / denom;
}
// light.frag:28-36
__device__ __forceinline__ float geometry_schlick_ggx(float NdotV, float roughness) {
    const float r = roughness + 1.0f;
    const float k = (r * r) / 8.0f;
    return NdotV / (NdotV * (1.0f - k) + k);
}

// What light.frag:192-210 and EvaluateShadow (:137-146) derive from one light for one fragment.
struct LightEval {
    float3 L;          // unit direction to the light (BRDF)
    float3 C;          // centre of the shadow-ray cone, un-normalised (TraceShadowRay's `L`)
    float attenuation; // incl. the spot cone
    float4 color_intensity;
    float radius;
};
__device__ __forceinline__ LightEval eval_light(const LightRec& L4, const float3 fragPos) {
    LightEval e;
    const float3 lpos = f3(L4.position_inner.x, L4.position_inner.y, L4.position_inner.z);
    const float3 ldir = f3(L4.direction_outer.x, L4.direction_outer.y, L4.direction_outer.z);
    const float3 Lvec = lpos - fragPos;
    const float dist = length3(Lvec);
    e.L = Lvec / dist; // normalize(L_)
    e.attenuation = 1.0f;
    if (L4.type == LUZW_LIGHT_DIRECTIONAL) {
        e.L = normalize3(-ldir);
    } else if (L4.type == LUZW_LIGHT_SPOT) {
        e.attenuation = 1.0f / (dist * dist);
        const float theta = dot3(e.L, normalize3(-ldir));
        const float epsilon = L4.position_inner.w - L4.direction_outer.w;
        e.attenuation *= clampf((theta - L4.direction_outer.w) / epsilon, 0.0f, 1.0f);
    } else if (L4.type == LUZW_LIGHT_POINT) {
        e.attenuation = 1.0f / (dist * dist);
    }
    e.C = (L4.type == LUZW_LIGHT_DIRECTIONAL) ? ldir * dot3(ldir, e.L) * dist : e.L * dist; // :140-146
    e.color_intensity = L4.color_intensity;
    e.radius = L4.radius;
    return e;
}

__device__ __forceinline__ LightRec load_light(const LightRec* lights, int li) {
    const float4* p = reinterpret_cast<const float4*>(lights + li);
    LightRec r;
    r.color_intensity = __ldg(p + 0);
    r.position_inner = __ldg(p + 1);
    r.direction_outer = __ldg(p + 2);
    const float4 t = __ldg(p + 3);
    r.type = __float_as_int(t.x);
    r.num_shadow_samples = __float_as_int(t.y);
    r.radius = t.z;
    r.shadow_map = __float_as_int(t.w);
    return r;
}

// ---- warp-private shared memory ---------------------------------------------------------------------
template <int RMAX>
struct WarpSmem {
    float4 ray[RMAX * 32];       // [q * nlit + rank]: direction xyz, w = id bits
    float4 org[2][32];           // per pixel: shadow-ray origin, AO-ray origin (light.frag:138-139, :229-230)
    float tmax[RMAX][32];        // per (source slot of the chunk, pixel)
    uint32_t hits[RMAX][32];     // occluded rays per (source slot, pixel)
    uint32_t cand[kMaxCand][32]; // AO candidate instances per pixel
    int ncand[32];
};

// id bits of a queued ray
__device__ __forceinline__ uint32_t make_ray_id(int lane, int slot, bool is_ao, uint32_t bit) {
    return (uint32_t)lane | ((uint32_t)slot << 5) | ((is_ao ? 1u : 0u) << 10) | (bit << 11);
}

struct MaskOut { // debug visibility masks (LUZRT_DEBUG_MASKS)
    uint32_t* shadow_mask;
    uint32_t* ao_mask;
    uint32_t shadow_words, ao_words;
    uint32_t x0, r0; // tile origin (row relative to row_start)
    int row_start;
    uint32_t width, height;
};

// TRACE phase: any-hit traversal of the n_rays queued rays of this warp, one ray per lane with refill.
//
// "If-if" form of the two-level wide-BVH traversal of traverse.cuh (same node test, triangle test and
// candidate-list semantics as trace_ray): every iteration each lane first does whatever cheap bookkeeping
// it needs to arrive at a node (test the leaf primitives it holds, leave an instance, pop the stack, take
// the next candidate, take a new ray, enter an instance), and then ALL lanes that hold a node execute the
// expensive 8-child box test together, once.  Rays here are short (a few nodes each), so keeping that test
// full matters more than the extra pass over the bookkeeping code.
template <int RMAX, bool MASKS, bool STATS>
__device__ __forceinline__ void trace_queue(const TraceScene& sc, WarpSmem<RMAX>& q, const int n_rays, const float ao_min,
                                            uint2* stack, LocalStats* st, uint32_t& n_occluded, const MaskOut& mo,
                                            const int node_repeat) {
    const int lane = threadIdx.x & 31;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint32_t one_bits = sc.one_bits;
    int head = 0; // warp-uniform
    bool have = false;

    int idx = 0;
    uint32_t id = 0;
    float tmin = 0.0f, tmax = 0.0f, inv_dd = 0.0f;
    float3 o = f3(0, 0, 0), d = f3(0, 0, 0), idir = f3(0, 0, 0);
    const WideNode* nodes = sc.tlas_nodes;
    const WideTri* tris = nullptr;
    int sp = 0, inst_sp = -1, ci = 0, n_cand = -1;
    uint2 ngroup = make_uint2(0u, 0u), tgroup = make_uint2(0u, 0u);

    while (true) {
        uint32_t pending = kNoInstance; // instance to enter in this iteration
        bool want_cand = false, finished = false, hit = false;

        // ---- (a) leaf primitives held from the last node visit (or popped from the stack) ----
        while (tgroup.y != 0u) {
            const int j = __ffs(tgroup.y) - 1;
            tgroup.y &= tgroup.y - 1u;
            const uint32_t prim = tgroup.x + (uint32_t)j;
            if (inst_sp < 0) {
                // TLAS leaf: postpone the rest of this node, enter the instance below
                if (tgroup.y != 0u) stack[sp++] = tgroup;
                if (ngroup.y > 0x00FFFFFFu) stack[sp++] = ngroup;
                pending = prim;
                tgroup = make_uint2(0u, 0u);
                ngroup = make_uint2(0u, 0u);
            } else {
                const float4* tp = reinterpret_cast<const float4*>(tris + prim);
                const float4 p0 = __ldg(tp + 0), p1 = __ldg(tp + 1), p2 = __ldg(tp + 2);
                if (STATS) st->tris++;
                float t, bu, bv;
                if (tri_test(p0, p1, p2, o, d, inv_dd, tmin, tmax, t, bu, bv)) {
                    hit = true;
                    finished = true;
                    tgroup.y = 0u;
                }
            }
        }

        // ---- (b) out of nodes at this level: leave the instance / pop / ask for the next candidate ----
        if (have && !finished && pending == kNoInstance && ngroup.y <= 0x00FFFFFFu) {
            const bool from_root = n_cand < 0;
            if (inst_sp >= 0 && sp == inst_sp) {
                inst_sp = -1; // BLAS exhausted
                if (from_root) { // back to world space
                    const float4 wr = q.ray[idx];
                    const float4 wg = q.org[(id >> 10) & 1u][id & 31u];
                    o = f3(wg.x, wg.y, wg.z);
                    d = f3(wr.x, wr.y, wr.z);
                    idir = f3(safe_rcp(d.x), safe_rcp(d.y), safe_rcp(d.z));
                    inv_dd = fast_rcp(dot3_fma(d, d));
                    nodes = sc.tlas_nodes;
                }
            }
            if (sp == 0) {
                if (from_root) finished = true;
                else want_cand = true;
            } else {
                const uint2 e = stack[--sp];
                if (e.y > 0x00FFFFFFu) ngroup = e;
                else tgroup = e; // a postponed primitive group: tested at the top of the next iteration
            }
        }

        // ---- (c) retire finished rays ----
        if (finished) {
            have = false;
            if (hit) {
                n_occluded++;
                const int pl = id & 31u, slot = (id >> 5) & 31u;
                atomicAdd(&q.hits[slot][pl], 1u);
                if (MASKS) {
                    const uint32_t x = mo.x0 + (uint32_t)(pl & 7), r = mo.r0 + (uint32_t)(pl >> 3);
                    int yy = (mo.row_start + (int)r) % (int)mo.height;
                    if (yy < 0) yy += (int)mo.height;
                    const size_t pix = (size_t)yy * mo.width + x;
                    const uint32_t b = id >> 11;
                    if ((id >> 10) & 1u)
                        atomicOr(mo.ao_mask + pix * mo.ao_words + (b >> 5), 1u << (b & 31u));
                    else
                        atomicOr(mo.shadow_mask + pix * mo.shadow_words + (b >> 5), 1u << (b & 31u));
                }
            }
        }

        // ---- (d) refill idle lanes (converged: the queue head lives in a register, no atomics) ----
        const unsigned need = __ballot_sync(kFull, !have);
        if (need != 0u && head < n_rays) {
            if (!have) {
                idx = head + __popc(need & lt_mask);
                if (idx < n_rays) {
                    const float4 r = q.ray[idx];
                    id = __float_as_uint(r.w);
                    const int pl = id & 31u, slot = (id >> 5) & 31u;
                    const bool is_ao = (id >> 10) & 1u;
                    const float4 og = q.org[is_ao ? 1 : 0][pl];
                    o = f3(og.x, og.y, og.z);
                    d = f3(r.x, r.y, r.z);
                    tmin = is_ao ? ao_min : 0.001f;
                    tmax = q.tmax[slot][pl];
                    n_cand = is_ao ? q.ncand[pl] : -1;
                    // rays with NaNs (e.g. the vertical-light tangent of light.frag:90) and null directions miss
                    have = (o.x == o.x && o.y == o.y && o.z == o.z && d.x == d.x && d.y == d.y && d.z == d.z &&
                            tmin == tmin && tmax == tmax) &&
                           !(d.x == 0.0f && d.y == 0.0f && d.z == 0.0f);
                    sp = 0;
                    inst_sp = -1;
                    ci = 0;
                    tgroup = make_uint2(0u, 0u);
                    nodes = sc.tlas_nodes;
                    if (n_cand < 0) { // descend from the TLAS root
                        ngroup = make_uint2(0u, 0x80000000u);
                        idir = f3(safe_rcp(d.x), safe_rcp(d.y), safe_rcp(d.z));
                        inv_dd = fast_rcp(dot3_fma(d, d));
                    } else {
                        ngroup = make_uint2(0u, 0u);
                        want_cand = have;
                    }
                }
            }
            head += __popc(need);
        }
        if (!__any_sync(kFull, have)) break;

        // ---- (e) next candidate instance whose world box the ray segment meets ----
        if (want_cand) {
            const float4 wr = q.ray[idx];
            const float4 wg = q.org[1][id & 31u];
            const float3 widir = f3(safe_rcp(wr.x), safe_rcp(wr.y), safe_rcp(wr.z));
            const uint32_t* cand = &q.cand[0][id & 31u];
            while (ci < n_cand) {
                const uint32_t inst = cand[ci * 32];
                ci++;
                const float4 blo = __ldg(sc.inst_boxes + 2 * inst), bhi = __ldg(sc.inst_boxes + 2 * inst + 1);
                // (b - o) first: exact to half an ulp even when the origin sits on the box
                const float tx0 = (blo.x - wg.x) * widir.x, tx1 = (bhi.x - wg.x) * widir.x;
                const float ty0 = (blo.y - wg.y) * widir.y, ty1 = (bhi.y - wg.y) * widir.y;
                const float tz0 = (blo.z - wg.z) * widir.z, tz1 = (bhi.z - wg.z) * widir.z;
                const float tn = fmaxf(fmaxf(fminf(tx0, tx1), fminf(ty0, ty1)), fmaxf(fminf(tz0, tz1), tmin));
                const float tf = fminf(fminf(fmaxf(tx0, tx1), fmaxf(ty0, ty1)), fminf(fmaxf(tz0, tz1), tmax));
                // slab distances carry a few ulps of relative error: keep everything within 2e-6 relative
                if (tn - tf <= 2e-6f * fmaxf(fabsf(tn), fabsf(tf))) {
                    pending = inst;
                    break;
                }
            }
            if (pending == kNoInstance) have = false; // no candidate left: the ray is unoccluded
        }

        // ---- (f) enter an instance: the ray goes to object space, t is preserved ----
        if (pending != kNoInstance) {
            const InstanceRec* rec = sc.instances + pending;
            const float4 r0 = __ldg(&rec->r0), r1 = __ldg(&rec->r1), r2 = __ldg(&rec->r2);
            const uint4 ptrs = __ldg(reinterpret_cast<const uint4*>(&rec->nodes));
            if (STATS) st->insts++;
            const float4 wr = q.ray[idx];
            const float4 wg = q.org[(id >> 10) & 1u][id & 31u];
            o = xform_point(r0, r1, r2, f3(wg.x, wg.y, wg.z));
            d = xform_dir(r0, r1, r2, f3(wr.x, wr.y, wr.z));
            idir = f3(safe_rcp(d.x), safe_rcp(d.y), safe_rcp(d.z));
            inv_dd = fast_rcp(dot3_fma(d, d));
            nodes = reinterpret_cast<const WideNode*>(((unsigned long long)ptrs.y << 32) | ptrs.x);
            tris = reinterpret_cast<const WideTri*>(((unsigned long long)ptrs.w << 32) | ptrs.z);
            inst_sp = sp;
            // a transformed ray with NaN/inf components (singular instance matrix) misses the instance
            const bool ok = (d.x == d.x && d.y == d.y && d.z == d.z && o.x == o.x && o.y == o.y && o.z == o.z) &&
                            !(d.x == 0.0f && d.y == 0.0f && d.z == 0.0f);
            ngroup = ok ? make_uint2(0u, 0x80000000u) : make_uint2(0u, 0u);
        }

        // ---- (g) node visits: one for every lane that holds a node, then more for as long as at least
        //      `node_repeat` lanes can go on without testing primitives (amortises the bookkeeping above) ----
        for (int pass = 0;; pass++) {
            const bool work = have && ngroup.y > 0x00FFFFFFu && tgroup.y == 0u;
            const int n_work = __popc(__ballot_sync(kFull, work));
            if (n_work == 0 || (pass > 0 && n_work < node_repeat)) break;
            if (work) {
                const uint32_t hits = ngroup.y;
                const uint32_t imask = hits & 0xFFu;
                const int bit = 31 - __clz(hits);
                ngroup.y &= ~(1u << bit);
                if (ngroup.y > 0x00FFFFFFu) stack[sp++] = ngroup;
                const int slot = bit - 24;
                const uint32_t rel = __popc(imask & ~(0xFFFFFFFFu << slot));
                const uint4* np = reinterpret_cast<const uint4*>(nodes + (ngroup.x + rel));
                const uint4 n0 = __ldg(np + 0), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3),
                            n4 = __ldg(np + 4);
                if (STATS) st->nodes++;
                const uint32_t slots = intersect_node(n0, n2, n3, n4, o, idir, tmin, tmax, one_bits);
                const uint32_t node_imask = n0.w >> 24;
                ngroup = make_uint2(n1.x, ((slots & node_imask) << 24) | node_imask);
                tgroup = make_uint2(n1.y, leaf_bits(slots & ~node_imask, n1.z, n1.w));
            }
        }
    }
}

template <bool MASKS, bool STATS, int MIN_BLOCKS, int RMAX>
__global__ void __launch_bounds__(32 * kWarps, MIN_BLOCKS) k_light_pass(const LightArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const FrameConst& fc = a.fc;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpSmem<RMAX>& q = reinterpret_cast<WarpSmem<RMAX>*>(smem_raw)[warp];

    const uint32_t x0 = blockIdx.x * 16 + (warp & 1) * 8, r0 = blockIdx.y * 8 + (warp >> 1) * 4;
    const uint32_t x = x0 + (lane & 7);
    const uint32_t r = r0 + (lane >> 3);
    const bool in_image = x < fc.width && r < a.row_count;
    uint32_t y = 0;
    if (in_image) {
        int yy = (a.row_start + (int)r) % (int)fc.height;
        if (yy < 0) yy += (int)fc.height;
        y = (uint32_t)yy;
    }
    const size_t pix = (size_t)y * fc.width + x;

    // ---- G-buffer fetch (light.frag:172-176; texel loads) ----
    float3 N = f3(0.0f, 0.0f, 0.0f);
    uchar4 a8 = make_uchar4(0, 0, 0, 0), m8 = a8, e8 = a8;
    float depth = 1.0f;
    uchar4 bn8 = a8;
    if (in_image) {
        const float4 n4 = __ldg(a.normal + pix);
        N = f3(n4.x, n4.y, n4.z);
        a8 = __ldg(a.albedo + pix);
        m8 = __ldg(a.material + pix);
        e8 = __ldg(a.emission + pix);
        depth = __ldg(a.depth + pix);
        // gl_FragCoord = (x+.5, y+.5); ivec2(mod(fragCoord, size)) == (x % w, y % h)
        bn8 = __ldg(a.blue_noise + (size_t)(y % fc.bn_h) * fc.bn_w + (x % fc.bn_w));
    }
    const float3 ambientLight = f3(fc.ambient[0], fc.ambient[1], fc.ambient[2]);
    const bool lit = in_image && (length3(N) != 0.0f); // :178
    if (in_image && !lit) a.out[pix] = make_float4(ambientLight.x, ambientLight.y, ambientLight.z, 1.0f);

    const unsigned lit_mask = __ballot_sync(kFull, lit);
    // lit pixels (the ray count of the frame follows from it); halo rows of a multi-GPU strip are not counted
    const unsigned int lit_warp = __popc(__ballot_sync(kFull, lit && r >= a.count_row_begin && r < a.count_row_end));
    if (lane == 0 && lit_warp)
        atomicAdd(a.lit_counters + 16 * ((blockIdx.x * 4u + blockIdx.y * 29u + warp) & 63u), (unsigned long long)lit_warp);
    if (MASKS && in_image) {
        uint32_t* smask = a.shadow_mask + pix * a.shadow_words;
        uint32_t* amask = a.ao_mask + pix * a.ao_words;
        for (uint32_t k = 0; k < a.shadow_words; k++) smask[k] = 0;
        for (uint32_t k = 0; k < a.ao_words; k++) amask[k] = 0;
        __threadfence_block();
    }
    if (lit_mask == 0u) return; // background tile (warps are independent: no CTA barrier below)
    const int nlit = __popc(lit_mask);
    const int rank = __popc(lit_mask & ((1u << lane) - 1u));

    const float3 albedo = f3(powf((float)a8.x / 255.0f, 2.2f), powf((float)a8.y / 255.0f, 2.2f),
                             powf((float)a8.z / 255.0f, 2.2f));
    const float roughness = (float)m8.x / 255.0f, metallic = (float)m8.y / 255.0f, occlusion = (float)m8.z / 255.0f;
    const float u = ((float)x + 0.5f) / (float)fc.width, v = ((float)y + 0.5f) / (float)fc.height;
    const float3 fragPos = depth_to_world(fc, u, v, depth);
    const float3 camPos = f3(fc.cam_pos[0], fc.cam_pos[1], fc.cam_pos[2]);
    const float3 V = normalize3(camPos - fragPos);
    const float3 F0 = f3(0.04f, 0.04f, 0.04f) * (1.0f - metallic) + albedo * metallic;
    const float camDist = length3(fragPos - camPos);
    const float bn_r = (float)bn8.x / 255.0f, bn_g = (float)bn8.y / 255.0f;
    const float NdotV = fmaxf(dot3(N, V), 0.0f);
    const float ggxV = geometry_schlick_ggx(NdotV, roughness);

    uint2 stack[LUZ_STACK_SIZE];
    LocalStats st = {0, 0, 0};
    uint32_t n_rays = 0, n_occl = 0;

    // ray origins (light.frag:138-139 shadow bias with its 0.05 floor, :229-230 AO bias without) and the AO frame
    const float3 Oshadow = fragPos + N * fmaxf(camDist * 0.01f, 0.05f);
    const float3 Oao = fragPos + N * (camDist * 0.01f);
    const float3 Tao = fabsf(N.z) > 0.5f ? f3(0.0f, -N.z, N.y) : f3(-N.y, N.x, 0.0f); // :116-117
    const float3 Bao = cross3(N, Tao);
    q.org[0][lane] = make_float4(Oshadow.x, Oshadow.y, Oshadow.z, 0.0f);
    q.org[1][lane] = make_float4(Oao.x, Oao.y, Oao.z, 0.0f);

    // One TLAS walk per pixel instead of one per AO ray: see collect_instances (traverse.cuh)
    int my_cand = -1;
    if (lit && fc.ao_num_samples >= kMinCandSamples) {
        // every AO ray of this pixel stays inside O +- aoMax * |dir| per axis, and
        // |dir_k| = |T_k h.x + B_k h.y + N_k h.z| <= sqrt(T_k^2 + B_k^2 + N_k^2) * |h| with |h| = 1 (+ rounding)
        const float m = fabsf(fc.ao_max) * 1.001f;
        const float3 ext = f3(m * sqrtf(Tao.x * Tao.x + Bao.x * Bao.x + N.x * N.x) + 1e-6f,
                              m * sqrtf(Tao.y * Tao.y + Bao.y * Bao.y + N.y * N.y) + 1e-6f,
                              m * sqrtf(Tao.z * Tao.z + Bao.z * Bao.z + N.z * N.z) + 1e-6f);
        my_cand = collect_instances<STATS>(a.scene, Oao - ext, Oao + ext, &q.cand[0][lane], 32, kMaxCand, stack, &st);
    }
    q.ncand[lane] = my_cand;

    MaskOut mo;
    mo.shadow_mask = a.shadow_mask;
    mo.ao_mask = a.ao_mask;
    mo.shadow_words = a.shadow_words;
    mo.ao_words = a.ao_words;
    mo.x0 = x0;
    mo.r0 = r0;
    mo.row_start = a.row_start;
    mo.width = fc.width;
    mo.height = fc.height;

    // ---- chunks of the pixel's ray list: lights in order (shadow rays, light.frag:192-227), then AO (:229-231) ----
    const int n_sources = fc.num_lights + 1;
    const int total = (int)a.rays_per_pixel;
    const int n_chunks = max(1, (total + RMAX - 1) / RMAX);
    const int Rc = (total + n_chunks - 1) / n_chunks;
    const bool rt_shadows = fc.shadow_type == LUZW_SHADOW_RAYTRACING;

    float3 Lo = f3(0.0f, 0.0f, 0.0f);
    float rayTracedAo = 1.0f;
    float carry = 0.0f;       // hits of a source whose samples straddle a chunk boundary
    int cur_l = 0, cur_i = 0; // warp-uniform cursor into the ray list
    uint32_t cur_bit = 0;     // shadow-mask bit of sample 0 of light cur_l

    for (int chunk = 0; chunk < n_chunks; chunk++) {
        // ---------------- GEN ----------------
        int qn = 0, slot = 0;
        {
            int l = cur_l, i = cur_i;
            uint32_t bit0 = cur_bit;
            while (l < n_sources) {
                const bool is_ao = l == fc.num_lights;
                LightRec L4;
                int ns;
                if (is_ao) {
                    ns = fc.ao_num_samples;
                } else {
                    L4 = load_light(a.lights, l);
                    ns = rt_shadows ? max(L4.num_shadow_samples, 0) : 0;
                }
                const int take = min(ns - i, Rc - qn);
                if (ns > 0 && take == 0) break; // the queue of this chunk is full
                if (take > 0 && lit) {
                    float3 C, T, B;
                    float radius = 0.0f;
                    if (is_ao) {
                        C = N;
                        T = Tao;
                        B = Bao;
                        q.tmax[slot][lane] = fc.ao_max;
                    } else {
                        const LightEval e = eval_light(L4, fragPos);
                        C = e.C;
                        T = normalize3(cross3(C, f3(0.0f, 1.0f, 0.0f))); // TraceShadowRay :90-91
                        B = normalize3(cross3(T, C));
                        radius = e.radius;
                        q.tmax[slot][lane] = length3(C); // :100
                    }
                    q.hits[slot][lane] = 0u;
                    for (int k = 0; k < take; k++) {
                        const int s = i + k;
                        const float2 rng = blue_noise_sample(bn_r, bn_g, s, fc.frame_mod);
                        float sn, cs;
                        float3 dir;
                        if (is_ao) { // HemisphereSample (light.frag:63-69)
                            const float rr = sqrtf(rng.x);
                            sincosf(6.283f * rng.y, &sn, &cs);
                            dir = T * (rr * cs) + B * (rr * sn) + C * sqrtf(fmaxf(0.0f, 1.0f - rng.x));
                        } else { // DiskSample (light.frag:57-61)
                            const float pointRadius = radius * sqrtf(rng.x);
                            sincosf(rng.y * 2.0f * kPI, &sn, &cs);
                            dir = normalize3(C + (pointRadius * cs) * T + (pointRadius * sn) * B);
                        }
                        const uint32_t bit = is_ao ? (uint32_t)s : bit0 + (uint32_t)s;
                        q.ray[(qn + k) * nlit + rank] =
                            make_float4(dir.x, dir.y, dir.z, __uint_as_float(make_ray_id(lane, slot, is_ao, bit)));
                    }
                    n_rays += (uint32_t)take;
                }
                if (take > 0) {
                    qn += take;
                    slot++;
                }
                i += take;
                if (i >= ns) {
                    l++;
                    i = 0;
                    bit0 += (uint32_t)ns;
                }
            }
        }
        __syncwarp();

        // ---------------- TRACE ----------------
        trace_queue<RMAX, MASKS, STATS>(a.scene, q, qn * nlit, fc.ao_min, stack, &st, n_occl, mo, (int)a.node_repeat);
        __syncwarp();

        // ---------------- SHADE ----------------
        {
            int l = cur_l, i = cur_i, qs = 0, s_slot = 0;
            uint32_t bit0 = cur_bit;
            while (l < n_sources) {
                const bool is_ao = l == fc.num_lights;
                LightRec L4;
                int ns;
                if (is_ao) {
                    ns = fc.ao_num_samples;
                } else {
                    L4 = load_light(a.lights, l);
                    ns = rt_shadows ? max(L4.num_shadow_samples, 0) : 0;
                }
                const int take = min(ns - i, Rc - qs);
                if (ns > 0 && take == 0) break;
                if (i == 0) carry = 0.0f;
                if (take > 0 && lit) carry += (float)q.hits[s_slot][lane];
                if (i + take >= ns && lit) { // all samples of this source are in: finalise it
                    if (is_ao) {
                        if (ns != 0) rayTracedAo = ((float)ns - carry) / (float)ns; // ao / aoNumSamples (:133-134)
                    } else {
                        const LightEval e = eval_light(L4, fragPos);
                        // shadow factor: RT with samples -> occluded fraction; RT with 0 samples -> 0 (:87-89);
                        // any other shadowType -> 1 (:166-168)
                        float shadowFactor = rt_shadows ? 0.0f : 1.0f;
                        if (ns > 0) shadowFactor = carry / (float)ns;
                        const float3 lcol = f3(e.color_intensity.x, e.color_intensity.y, e.color_intensity.z);
                        const float3 radiance = lcol * e.color_intensity.w * e.attenuation * (1.0f - shadowFactor);
                        const float3 L = e.L;
                        const float3 H = normalize3(V + L);
                        const float NDF = distribution_ggx(N, H, roughness);
                        const float NdotL = fmaxf(dot3(N, L), 0.0f);
                        const float G = geometry_schlick_ggx(NdotL, roughness) * ggxV; // GeometrySmith :38-45
                        const float fp = powf(clampf(1.0f - clampf(dot3(H, V), 0.0f, 1.0f), 0.0f, 1.0f), 5.0f);
                        const float3 F = F0 + (f3(1.0f, 1.0f, 1.0f) - F0) * fp; // FresnelSchlick :47-49
                        const float3 num = NDF * G * F;
                        const float denom = 4.0f * NdotV * NdotL + 0.0001f;
                        const float3 spec = num / denom;
                        float3 kD = f3(1.0f, 1.0f, 1.0f) - F;
                        kD = kD * (1.0f - metallic);
                        Lo = Lo + (kD * albedo / kPI + spec) * radiance * NdotL;
                    }
                }
                if (take > 0) {
                    qs += take;
                    s_slot++;
                }
                i += take;
                if (i >= ns) {
                    l++;
                    i = 0;
                    bit0 += (uint32_t)ns;
                }
            }
            cur_l = l;
            cur_i = i;
            cur_bit = bit0;
        }
        __syncwarp();
    }

    if (lit) {
        const float3 emission = f3((float)e8.x / 255.0f, (float)e8.y / 255.0f, (float)e8.z / 255.0f);
        const float3 ambient = ambientLight * albedo * occlusion * rayTracedAo;
        const float3 color = ambient + Lo + emission;
        a.out[pix] = make_float4(color.x, color.y, color.z, 1.0f);
    }

    if (STATS) {
        unsigned long long vals[5] = {n_rays, st.nodes, st.tris, st.insts, n_occl};
#pragma unroll
        for (int k = 0; k < 5; k++) {
            unsigned long long vsum = vals[k];
            for (int off = 16; off; off >>= 1) vsum += __shfl_xor_sync(kFull, vsum, off);
            vals[k] = vsum;
        }
        if (lane == 0) {
            atomicAdd(&a.stats->rays, vals[0]);
            atomicAdd(&a.stats->nodes, vals[1]);
            atomicAdd(&a.stats->tris, vals[2]);
            atomicAdd(&a.stats->insts, vals[3]);
            atomicAdd(&a.stats->occluded, vals[4]);
        }
    }
}

template <bool MASKS, bool STATS, int MINB, int RMAX>
cudaError_t launch_variant(cudaStream_t stream, const LightArgs& args, dim3 grid) {
    const size_t smem = sizeof(WarpSmem<RMAX>) * kWarps;
    auto kern = k_light_pass<MASKS, STATS, MINB, RMAX>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, 32 * kWarps, smem, stream>>>(args);
    return cudaGetLastError();
}

cudaError_t launch_sel(cudaStream_t stream, const LightArgs& args, dim3 grid, bool masks, bool stats, int rmax) {
    if (masks && stats) return launch_variant<true, true, 3, 10>(stream, args, grid);
    if (masks) return launch_variant<true, false, 3, 10>(stream, args, grid);
    if (stats) return launch_variant<false, true, 3, 10>(stream, args, grid);
    if (rmax == 16) return launch_variant<false, false, 3, 16>(stream, args, grid);
    if (rmax == 20) return launch_variant<false, false, 3, 20>(stream, args, grid);
    if (rmax == 8) return launch_variant<false, false, 4, 8>(stream, args, grid);
    return launch_variant<false, false, 4, 10>(stream, args, grid);
}

} // namespace

cudaError_t launch_light_pass(cudaStream_t stream, const LightArgs& args, bool masks, bool stats) {
    if (args.row_count == 0 || args.fc.width == 0) return cudaSuccess;
    const dim3 grid((args.fc.width + 15) / 16, (args.row_count + 7) / 8);
    // LUZRT_LIGHT_RMAX selects among the compiled queue depths (rays per pixel per chunk) for tuning runs
    static const int rmax = [] {
        const char* e = getenv("LUZRT_LIGHT_RMAX");
        return e ? atoi(e) : 10;
    }();
    static const int node_repeat = [] {
        const char* e = getenv("LUZRT_NODE_REPEAT");
        return e ? atoi(e) : 20;
    }();
    LightArgs a2 = args;
    a2.node_repeat = (uint32_t)node_repeat;
    return launch_sel(stream, a2, grid, masks, stats, rmax);
}

} // namespace luz

// traverse.cuh -- two-level stack traversal of the 8-wide compressed BVH (device code).
//
// Replaces what VK_KHR_ray_query does for light.frag:99-106 / :125-132 (the driver's traversal is
// not in the reference tree): opaque two-sided triangles, cull mask 0xFF, any-hit terminates on
// the first committed hit with tmin < t < tmax, t parametric along the un-normalised direction
// and preserved across the instance transform (VulkanWrapper.cpp:1119-1126).
#pragma once

#include "common.cuh"

namespace luz {

#define LUZ_STACK_SIZE 40

struct HitInfo {
    float t;
    uint32_t inst; // index into the TLAS-ordered instance array
    uint32_t prim; // original triangle index inside the BLAS
    float bu, bv;  // barycentric weights of v0, v1
};

struct LocalStats {
    uint32_t nodes, tris, insts;
};

__device__ __forceinline__ float safe_rcp(float d) {
    const float ooeps = 8.271806125530277e-25f; // 2^-80
    return 1.0f / (fabsf(d) > ooeps ? d : copysignf(ooeps, d));
}

// Watertight two-sided ray/triangle test by signed volumes (scalar triple products) evaluated
// with explicitly rounded operations: the edge function of a shared edge is bitwise
// antisymmetric between the two triangles that share it, so no ray slips between them.
__device__ __forceinline__ bool tri_test(const float4 p0, const float4 p1, const float4 p2, const float3 o,
                                         const float3 d, const float inv_dd, const float tmin, const float tmax,
                                         float& t_out, float& bu, float& bv) {
    const float3 A = f3(p0.x - o.x, p0.y - o.y, p0.z - o.z);
    const float3 B = f3(p1.x - o.x, p1.y - o.y, p1.z - o.z);
    const float3 C = f3(p2.x - o.x, p2.y - o.y, p2.z - o.z);
#define LUZ_TRIPLE(P, Q)                                                                                      \
    __fadd_rn(__fadd_rn(__fmul_rn(d.x, __fsub_rn(__fmul_rn(P.y, Q.z), __fmul_rn(P.z, Q.y))),                  \
                        __fmul_rn(d.y, __fsub_rn(__fmul_rn(P.z, Q.x), __fmul_rn(P.x, Q.z)))),                 \
              __fmul_rn(d.z, __fsub_rn(__fmul_rn(P.x, Q.y), __fmul_rn(P.y, Q.x))))
    const float U = LUZ_TRIPLE(C, B);
    const float V = LUZ_TRIPLE(A, C);
    const float W = LUZ_TRIPLE(B, A);
#undef LUZ_TRIPLE
    const float mn = fminf(U, fminf(V, W)), mx = fmaxf(U, fmaxf(V, W));
    if (mn < 0.0f && mx > 0.0f) return false;
    const float det = U + V + W;
    if (det == 0.0f) return false;
    // hit point relative to the origin = (U*A + V*B + W*C) / det; t = (P . d) / (d . d)
    const float3 P = f3(fmaf(U, A.x, fmaf(V, B.x, W * C.x)), fmaf(U, A.y, fmaf(V, B.y, W * C.y)),
                        fmaf(U, A.z, fmaf(V, B.z, W * C.z)));
    const float t = (fmaf(P.x, d.x, fmaf(P.y, d.y, P.z * d.z)) * inv_dd) / det;
    if (!(t > tmin && t < tmax)) return false;
    t_out = t;
    bu = U / det;
    bv = V / det;
    return true;
}

// Intersects the ray with the eight quantised child boxes; returns the hit mask in the
// "bits 31..24 = internal children by slot, bits 23..0 = leaf primitives" form.
__device__ __forceinline__ uint32_t intersect_node(const uint4 n0, const uint4 n1, const uint4 n2, const uint4 n3,
                                                   const uint4 n4, const float3 o, const float3 idir,
                                                   const float tmin, const float tmax) {
    const float adjx = __uint_as_float((n0.w & 0xFFu) << 23) * idir.x;
    const float adjy = __uint_as_float(((n0.w >> 8) & 0xFFu) << 23) * idir.y;
    const float adjz = __uint_as_float(((n0.w >> 16) & 0xFFu) << 23) * idir.z;
    const float ox = (__uint_as_float(n0.x) - o.x) * idir.x;
    const float oy = (__uint_as_float(n0.y) - o.y) * idir.y;
    const float oz = (__uint_as_float(n0.z) - o.z) * idir.z;
    uint32_t hitmask = 0;
#pragma unroll
    for (int half = 0; half < 2; half++) {
        const uint32_t meta4 = half ? n1.w : n1.z;
        const uint32_t lox = half ? n2.y : n2.x, loy = half ? n2.w : n2.z, loz = half ? n3.y : n3.x;
        const uint32_t hix = half ? n3.w : n3.z, hiy = half ? n4.y : n4.x, hiz = half ? n4.w : n4.z;
        const uint32_t nx = idir.x < 0.0f ? hix : lox, fx = idir.x < 0.0f ? lox : hix;
        const uint32_t ny = idir.y < 0.0f ? hiy : loy, fy = idir.y < 0.0f ? loy : hiy;
        const uint32_t nz = idir.z < 0.0f ? hiz : loz, fz = idir.z < 0.0f ? loz : hiz;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const float tnx = fmaf((float)((nx >> (8 * j)) & 0xFFu), adjx, ox);
            const float tny = fmaf((float)((ny >> (8 * j)) & 0xFFu), adjy, oy);
            const float tnz = fmaf((float)((nz >> (8 * j)) & 0xFFu), adjz, oz);
            const float tfx = fmaf((float)((fx >> (8 * j)) & 0xFFu), adjx, ox);
            const float tfy = fmaf((float)((fy >> (8 * j)) & 0xFFu), adjy, oy);
            const float tfz = fmaf((float)((fz >> (8 * j)) & 0xFFu), adjz, oz);
            const float cmin = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, tmin));
            const float cmax = fminf(fminf(tfx, tfy), fminf(tfz, tmax));
            if (cmin <= cmax * 1.0000005f) {
                const uint32_t meta = (meta4 >> (8 * j)) & 0xFFu;
                hitmask |= (meta >> 5) << (meta & 31u);
            }
        }
    }
    return hitmask;
}

// CLOSEST == false: any-hit, returns true at the first committed intersection.
// CLOSEST == true : returns true if something was hit; *hit describes the nearest one.
template <bool CLOSEST, bool STATS>
__device__ __forceinline__ bool trace_ray(const TraceScene& sc, const float3 wo, const float3 wd, const float tmin,
                                          float tmax, HitInfo* hit, LocalStats* st) {
    // rays with NaNs (e.g. the vertical-light tangent of light.frag:90) and null directions miss
    if (!(wo.x == wo.x && wo.y == wo.y && wo.z == wo.z && wd.x == wd.x && wd.y == wd.y && wd.z == wd.z &&
          tmin == tmin && tmax == tmax))
        return false;
    if (wd.x == 0.0f && wd.y == 0.0f && wd.z == 0.0f) return false;

    uint2 stack[LUZ_STACK_SIZE];
    int sp = 0;
    int inst_sp = -1; // stack height at which the current instance was entered, -1 = in the TLAS
    uint32_t cur_inst = 0;

    float3 o = wo, d = wd;
    float3 idir = f3(safe_rcp(d.x), safe_rcp(d.y), safe_rcp(d.z));
    float inv_dd = 1.0f / dot3(d, d);
    const WideNode* nodes = sc.tlas_nodes;
    const WideTri* tris = nullptr;
    bool found = false;

    uint2 ngroup = make_uint2(0u, 0x80000000u); // root: slot 7 of a virtual parent with imask 0
    uint2 tgroup = make_uint2(0u, 0u);

    while (true) {
        if (ngroup.y > 0x00FFFFFFu) {
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
            const uint32_t hm = intersect_node(n0, n1, n2, n3, n4, o, idir, tmin, tmax);
            ngroup = make_uint2(n1.x, (hm & 0xFF000000u) | (n0.w >> 24));
            tgroup = make_uint2(n1.y, hm & 0x00FFFFFFu);
        } else {
            tgroup = ngroup;
            ngroup = make_uint2(0u, 0u);
        }

        while (tgroup.y != 0u) {
            const int j = __ffs(tgroup.y) - 1;
            tgroup.y &= tgroup.y - 1u;
            const uint32_t prim = tgroup.x + (uint32_t)j;
            if (inst_sp < 0) {
                // TLAS leaf: enter the instance
                if (tgroup.y != 0u) stack[sp++] = tgroup;
                if (ngroup.y > 0x00FFFFFFu) stack[sp++] = ngroup;
                const InstanceRec* rec = sc.instances + prim;
                const float4 r0 = __ldg(&rec->r0), r1 = __ldg(&rec->r1), r2 = __ldg(&rec->r2);
                const uint4 ptrs = __ldg(reinterpret_cast<const uint4*>(&rec->nodes));
                if (STATS) st->insts++;
                o = f3(r0.x * wo.x + r0.y * wo.y + r0.z * wo.z + r0.w, r1.x * wo.x + r1.y * wo.y + r1.z * wo.z + r1.w,
                       r2.x * wo.x + r2.y * wo.y + r2.z * wo.z + r2.w);
                d = f3(r0.x * wd.x + r0.y * wd.y + r0.z * wd.z, r1.x * wd.x + r1.y * wd.y + r1.z * wd.z,
                       r2.x * wd.x + r2.y * wd.y + r2.z * wd.z);
                idir = f3(safe_rcp(d.x), safe_rcp(d.y), safe_rcp(d.z));
                inv_dd = 1.0f / dot3(d, d);
                nodes = reinterpret_cast<const WideNode*>(((unsigned long long)ptrs.y << 32) | ptrs.x);
                tris = reinterpret_cast<const WideTri*>(((unsigned long long)ptrs.w << 32) | ptrs.z);
                cur_inst = prim;
                inst_sp = sp;
                tgroup = make_uint2(0u, 0u);
                // a transformed ray with NaN/inf components (singular instance matrix) misses the instance
                const bool ok = (d.x == d.x && d.y == d.y && d.z == d.z && o.x == o.x && o.y == o.y && o.z == o.z) &&
                                !(d.x == 0.0f && d.y == 0.0f && d.z == 0.0f);
                ngroup = ok ? make_uint2(0u, 0x80000000u) : make_uint2(0u, 0u);
                break;
            } else {
                const float4* tp = reinterpret_cast<const float4*>(tris + prim);
                const float4 p0 = __ldg(tp + 0), p1 = __ldg(tp + 1), p2 = __ldg(tp + 2);
                if (STATS) st->tris++;
                float t, bu, bv;
                if (tri_test(p0, p1, p2, o, d, inv_dd, tmin, tmax, t, bu, bv)) {
                    if (!CLOSEST) return true;
                    tmax = t;
                    found = true;
                    hit->t = t;
                    hit->inst = cur_inst;
                    hit->prim = __float_as_uint(p0.w);
                    hit->bu = bu;
                    hit->bv = bv;
                }
            }
        }

        if (ngroup.y <= 0x00FFFFFFu) {
            if (inst_sp >= 0 && sp == inst_sp) {
                // BLAS exhausted: back to world space
                inst_sp = -1;
                o = wo;
                d = wd;
                idir = f3(safe_rcp(d.x), safe_rcp(d.y), safe_rcp(d.z));
                inv_dd = 1.0f / dot3(d, d);
                nodes = sc.tlas_nodes;
            }
            if (sp == 0) break;
            ngroup = stack[--sp];
        }
    }
    return found;
}

} // namespace luz

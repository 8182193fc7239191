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

// Ray vs the eight quantised child boxes of a wide node.
//
// Dequantisation without integer->float conversions (I2F runs on the quarter-rate XU pipe and was
// the top stall in the first ncu profile, profiles/r1_light_pass_v0.md): one PRMT drops the byte q
// into the mantissa of 1.0f, giving v = 1 + q*2^-15 exactly, and one FFMA evaluates
//     t = q*adj + o  ==  v*A + (o - A),   A = adj * 2^15.
// Rounding (o - A) costs at most |adj|*2^-9 (1/512 of a grid cell); both planes are pushed outwards
// by |adj|*2^-8 and the far plane is scaled by (1 + 2^-21) so the test stays conservative.
__device__ __forceinline__ float q_as_float(uint32_t packed, uint32_t one_bits, uint32_t selector) {
    return __uint_as_float(__byte_perm(packed, one_bits, selector));
}

// Returns the 8-bit mask of child slots whose box the ray overlaps in [tmin, tmax].
__device__ __forceinline__ uint32_t intersect_node(const uint4 n0, const uint4 n2, const uint4 n3, const uint4 n4,
                                                   const float3 o, const float3 idir, const float tmin,
                                                   const float tmax, const uint32_t one_bits) {
    const float kFar = 1.0000005f;
    // per-axis constants
    const float adjx = __uint_as_float((n0.w & 0xFFu) << 23) * idir.x;
    const float adjy = __uint_as_float(((n0.w >> 8) & 0xFFu) << 23) * idir.y;
    const float adjz = __uint_as_float(((n0.w >> 16) & 0xFFu) << 23) * idir.z;
    const float Ax = adjx * 32768.0f, Ay = adjy * 32768.0f, Az = adjz * 32768.0f;
    const float Bx = (__uint_as_float(n0.x) - o.x) * idir.x - Ax;
    const float By = (__uint_as_float(n0.y) - o.y) * idir.y - Ay;
    const float Bz = (__uint_as_float(n0.z) - o.z) * idir.z - Az;
    const float px = fabsf(adjx) * 0.00390625f, py = fabsf(adjy) * 0.00390625f, pz = fabsf(adjz) * 0.00390625f;
    const float Bnx = Bx - px, Bny = By - py, Bnz = Bz - pz;
    const float Afx = Ax * kFar, Afy = Ay * kFar, Afz = Az * kFar;
    const float Bfx = fmaf(Bx, kFar, px), Bfy = fmaf(By, kFar, py), Bfz = fmaf(Bz, kFar, pz);
    const bool negx = idir.x < 0.0f, negy = idir.y < 0.0f, negz = idir.z < 0.0f;
    uint32_t slots = 0;
#pragma unroll
    for (int half = 0; half < 2; half++) {
        const uint32_t lox = half ? n2.y : n2.x, loy = half ? n2.w : n2.z, loz = half ? n3.y : n3.x;
        const uint32_t hix = half ? n3.w : n3.z, hiy = half ? n4.y : n4.x, hiz = half ? n4.w : n4.z;
        const uint32_t nx = negx ? hix : lox, fx = negx ? lox : hix;
        const uint32_t ny = negy ? hiy : loy, fy = negy ? loy : hiy;
        const uint32_t nz = negz ? hiz : loz, fz = negz ? loz : hiz;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t sel = 0x7604u | ((uint32_t)j << 4); // bytes: [3F][80][q_j][00]
            const float tnx = fmaf(q_as_float(nx, one_bits, sel), Ax, Bnx);
            const float tny = fmaf(q_as_float(ny, one_bits, sel), Ay, Bny);
            const float tnz = fmaf(q_as_float(nz, one_bits, sel), Az, Bnz);
            const float tfx = fmaf(q_as_float(fx, one_bits, sel), Afx, Bfx);
            const float tfy = fmaf(q_as_float(fy, one_bits, sel), Afy, Bfy);
            const float tfz = fmaf(q_as_float(fz, one_bits, sel), Afz, Bfz);
            const float cmin = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, tmin));
            const float cmax = fminf(fminf(tfx, tfy), fminf(tfz, tmax));
            if (cmin <= cmax) slots |= 1u << (4 * half + j);
        }
    }
    return slots;
}

// Expands the hit leaf slots of a node into the primitive bit field (bits 23..0).  Empty slots carry
// an inverted box and never hit; internal slots are masked out by the caller.
__device__ __forceinline__ uint32_t leaf_bits(uint32_t leaf_slots, const uint32_t meta_lo, const uint32_t meta_hi) {
    uint32_t bits = 0;
    while (leaf_slots) {
        const int s = __ffs(leaf_slots) - 1;
        leaf_slots &= leaf_slots - 1u;
        const uint32_t m4 = s < 4 ? meta_lo : meta_hi;
        const uint32_t meta = (m4 >> (8 * (s & 3))) & 0xFFu;
        bits |= (meta >> 5) << (meta & 31u);
    }
    return bits;
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

    // 0x3F800000 comes in as a kernel parameter: PRMT then takes it straight from the constant bank and keeps
    // its selector as an immediate (with a literal, ptxas puts the selector in a register and re-creates it
    // with a MOV in front of every PRMT: profiles/r1_light_pass_v1.md)
    const uint32_t one_bits = sc.one_bits;
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
        // node phase: every lane keeps descending until it holds primitives to test (or runs out of nodes), so
        // that the expensive leaf work below is entered by as many lanes of the warp together as possible
        while (ngroup.y > 0x00FFFFFFu && tgroup.y == 0u) {
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
            const uint2 e = stack[--sp];
            if (e.y > 0x00FFFFFFu) {
                ngroup = e;
            } else { // a postponed primitive group
                tgroup = e;
                ngroup = make_uint2(0u, 0u);
            }
        }
    }
    return found;
}

} // namespace luz

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

// Reciprocal for slab tests: one MUFU.RCP (relative error <= 2^-23) instead of the ~10-instruction IEEE
// division.  The box test absorbs it: far planes are scaled by kFar = 1 + 2^-20 (see intersect_node).
__device__ __forceinline__ float fast_rcp(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float safe_rcp(float d) {
    const float ooeps = 8.271806125530277e-25f; // 2^-80
    return fast_rcp(fabsf(d) > ooeps ? d : copysignf(ooeps, d));
}

// Affine 3x4 transforms of a point / a direction with explicit FMAs (this header is also compiled into
// translation units built with -fmad=false, where a*b+c would otherwise cost two instructions).
__device__ __forceinline__ float3 xform_point(const float4 r0, const float4 r1, const float4 r2, const float3 p) {
    return f3(fmaf(r0.x, p.x, fmaf(r0.y, p.y, fmaf(r0.z, p.z, r0.w))), fmaf(r1.x, p.x, fmaf(r1.y, p.y, fmaf(r1.z, p.z, r1.w))),
              fmaf(r2.x, p.x, fmaf(r2.y, p.y, fmaf(r2.z, p.z, r2.w))));
}
__device__ __forceinline__ float3 xform_dir(const float4 r0, const float4 r1, const float4 r2, const float3 v) {
    return f3(fmaf(r0.x, v.x, fmaf(r0.y, v.y, r0.z * v.z)), fmaf(r1.x, v.x, fmaf(r1.y, v.y, r1.z * v.z)),
              fmaf(r2.x, v.x, fmaf(r2.y, v.y, r2.z * v.z)));
}
__device__ __forceinline__ float dot3_fma(const float3 a, const float3 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)); }

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
    const float kFar = 1.000001f; // 1 + 2^-20: covers the rounding of the FFMAs and of the approximate reciprocal
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

// ---- shared-origin candidate lists ------------------------------------------------------------------
// All AO rays of a pixel leave one origin and are at most aoMax * |dir| long (light.frag:116-126), so every
// instance any of them can reach overlaps one small world-space box around that origin.  collect_instances
// walks the TLAS ONCE per pixel with that box and records the overlapping leaves; trace_ray then visits only
// those instances for each of the pixel's AO rays and skips the TLAS levels entirely.  The set is a superset
// of the instances the ordinary descent would enter (a ray segment that meets a leaf box lies inside the
// query box), and the per-instance work is the same code on the same transformed ray, so visibility is
// identical to a root descent.
//
// Returns the number of instances written to out[k * stride], or -1 if there are more than max_out (the
// caller then falls back to the root descent).  Child boxes are dequantised exactly as the builder checked
// them (origin + q * 2^e, one rounding), so the test is conservative.
template <bool STATS>
__device__ __forceinline__ int collect_instances(const TraceScene& sc, const float3 lo, const float3 hi, uint32_t* out,
                                                 const int stride, const int max_out, uint2* stack, LocalStats* st) {
    if (!(lo.x <= hi.x && lo.y <= hi.y && lo.z <= hi.z)) return -1; // NaN bounds
    int sp = 0, n = 0;
    uint2 ngroup = make_uint2(0u, 0x80000000u);
    while (true) {
        while (ngroup.y > 0x00FFFFFFu) {
            const uint32_t hits = ngroup.y;
            const uint32_t imask = hits & 0xFFu;
            const int bit = 31 - __clz(hits);
            ngroup.y &= ~(1u << bit);
            if (ngroup.y > 0x00FFFFFFu) stack[sp++] = ngroup;
            const int slot = bit - 24;
            const uint32_t rel = __popc(imask & ~(0xFFFFFFFFu << slot));
            const uint4* np = reinterpret_cast<const uint4*>(sc.tlas_nodes + (ngroup.x + rel));
            const uint4 n0 = __ldg(np + 0), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3),
                        n4 = __ldg(np + 4);
            if (STATS) st->nodes++;
            // the query box in the node's 8-bit grid, widened by a cell on each side (absorbs every rounding of
            // the conversion and of the builder's dequantisation check), then pure byte compares per child
            const uint32_t ex = n0.w & 0xFFu, ey = (n0.w >> 8) & 0xFFu, ez = (n0.w >> 16) & 0xFFu;
            if (max(ex, max(ey, ez)) >= 254u) return -1; // 2^-(e-127) not representable: let the root descent handle it
            const float ix = __uint_as_float((254u - ex) << 23), iy = __uint_as_float((254u - ey) << 23),
                        iz = __uint_as_float((254u - ez) << 23);
            const float px = __uint_as_float(n0.x), py = __uint_as_float(n0.y), pz = __uint_as_float(n0.z);
            const int glx = (int)fminf(fmaxf(floorf((lo.x - px) * ix) - 1.0f, 0.0f), 255.0f);
            const int gly = (int)fminf(fmaxf(floorf((lo.y - py) * iy) - 1.0f, 0.0f), 255.0f);
            const int glz = (int)fminf(fmaxf(floorf((lo.z - pz) * iz) - 1.0f, 0.0f), 255.0f);
            const int ghx = (int)fminf(fmaxf(floorf((hi.x - px) * ix) + 2.0f, 0.0f), 255.0f);
            const int ghy = (int)fminf(fmaxf(floorf((hi.y - py) * iy) + 2.0f, 0.0f), 255.0f);
            const int ghz = (int)fminf(fmaxf(floorf((hi.z - pz) * iz) + 2.0f, 0.0f), 255.0f);
            uint32_t slots = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const uint32_t sel = 0x4440u + (uint32_t)(k & 3); // byte k&3 zero-extended
                const int meta = (int)__byte_perm(k < 4 ? n1.z : n1.w, 0u, sel);
                const int qlx = (int)__byte_perm(k < 4 ? n2.x : n2.y, 0u, sel), qly = (int)__byte_perm(k < 4 ? n2.z : n2.w, 0u, sel);
                const int qlz = (int)__byte_perm(k < 4 ? n3.x : n3.y, 0u, sel), qhx = (int)__byte_perm(k < 4 ? n3.z : n3.w, 0u, sel);
                const int qhy = (int)__byte_perm(k < 4 ? n4.x : n4.y, 0u, sel), qhz = (int)__byte_perm(k < 4 ? n4.z : n4.w, 0u, sel);
                const bool ov = meta != 0 && qlx <= ghx && qhx >= glx && qly <= ghy && qhy >= gly && qlz <= ghz && qhz >= glz;
                if (ov) slots |= 1u << k;
            }
            const uint32_t node_imask = n0.w >> 24;
            uint32_t prims = leaf_bits(slots & ~node_imask, n1.z, n1.w);
            while (prims) {
                const int j = __ffs(prims) - 1;
                prims &= prims - 1u;
                if (n >= max_out) return -1;
                out[n * stride] = n1.y + (uint32_t)j;
                n++;
            }
            ngroup = make_uint2(n1.x, ((slots & node_imask) << 24) | node_imask);
        }
        if (sp == 0) break;
        ngroup = stack[--sp];
    }
    return n;
}

// CLOSEST == false: any-hit, returns true at the first committed intersection.
// CLOSEST == true : returns true if something was hit; *hit describes the nearest one.
// n_cand < 0: descend from the TLAS root.  n_cand >= 0: visit only the instances cand[k * cand_stride]
// (TLAS leaf order indices from collect_instances).  `stack` is LUZ_STACK_SIZE entries of caller storage.
constexpr uint32_t kNoInstance = 0xFFFFFFFFu;

template <bool CLOSEST, bool STATS>
__device__ __forceinline__ bool trace_ray(const TraceScene& sc, const float3 wo, const float3 wd, const float tmin,
                                          float tmax, HitInfo* hit, LocalStats* st, uint2* stack,
                                          const uint32_t* cand = nullptr, const int cand_stride = 0,
                                          const int n_cand = -1) {
    // rays with NaNs (e.g. the vertical-light tangent of light.frag:90) and null directions miss
    if (!(wo.x == wo.x && wo.y == wo.y && wo.z == wo.z && wd.x == wd.x && wd.y == wd.y && wd.z == wd.z &&
          tmin == tmin && tmax == tmax))
        return false;
    if (wd.x == 0.0f && wd.y == 0.0f && wd.z == 0.0f) return false;

    // 0x3F800000 comes in as a kernel parameter: PRMT then takes it straight from the constant bank and keeps
    // its selector as an immediate (with a literal, ptxas puts the selector in a register and re-creates it
    // with a MOV in front of every PRMT: profiles/r1_light_pass_v1.md)
    const uint32_t one_bits = sc.one_bits;
    int sp = 0;
    int inst_sp = -1; // stack height at which the current instance was entered, -1 = in the TLAS
    uint32_t cur_inst = 0;
    const bool from_root = n_cand < 0;

    float3 o = wo, d = wd;
    float3 idir;
    float inv_dd;
    const WideNode* nodes = sc.tlas_nodes;
    const WideTri* tris = nullptr;
    bool found = false;

    uint2 ngroup = make_uint2(0u, 0u);
    uint2 tgroup = make_uint2(0u, 0u);
    uint32_t pending = kNoInstance; // instance to enter at the top of the loop
    int ci = 0;
    if (from_root) {
        ngroup = make_uint2(0u, 0x80000000u); // root: slot 7 of a virtual parent with imask 0
        idir = f3(safe_rcp(d.x), safe_rcp(d.y), safe_rcp(d.z));
        inv_dd = fast_rcp(dot3_fma(d, d));
    } else {
        // candidate mode keeps the world-space reciprocal direction for the box pre-test below
        idir = f3(safe_rcp(d.x), safe_rcp(d.y), safe_rcp(d.z));
        inv_dd = 0.0f;
    }
    const float3 widir = idir;

    while (true) {
        if (pending != kNoInstance) {
            // enter an instance: the ray goes to object space, t is preserved (direction not re-normalised)
            const InstanceRec* rec = sc.instances + pending;
            const float4 r0 = __ldg(&rec->r0), r1 = __ldg(&rec->r1), r2 = __ldg(&rec->r2);
            const uint4 ptrs = __ldg(reinterpret_cast<const uint4*>(&rec->nodes));
            if (STATS) st->insts++;
            o = xform_point(r0, r1, r2, wo);
            d = xform_dir(r0, r1, r2, wd);
            idir = f3(safe_rcp(d.x), safe_rcp(d.y), safe_rcp(d.z));
            inv_dd = fast_rcp(dot3_fma(d, d));
            nodes = reinterpret_cast<const WideNode*>(((unsigned long long)ptrs.y << 32) | ptrs.x);
            tris = reinterpret_cast<const WideTri*>(((unsigned long long)ptrs.w << 32) | ptrs.z);
            cur_inst = pending;
            pending = kNoInstance;
            inst_sp = sp;
            tgroup = make_uint2(0u, 0u);
            // a transformed ray with NaN/inf components (singular instance matrix) misses the instance
            const bool ok = (d.x == d.x && d.y == d.y && d.z == d.z && o.x == o.x && o.y == o.y && o.z == o.z) &&
                            !(d.x == 0.0f && d.y == 0.0f && d.z == 0.0f);
            ngroup = ok ? make_uint2(0u, 0x80000000u) : make_uint2(0u, 0u);
        }

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
            // when only a few lanes are still descending, yield so that the lanes waiting at the end of this loop
            // (they hold primitives or need a pop) get served and everybody re-enters the node test together
            if (__popc(__activemask()) < (int)sc.min_node_lanes) break;
        }

        while (tgroup.y != 0u) {
            const int j = __ffs(tgroup.y) - 1;
            tgroup.y &= tgroup.y - 1u;
            const uint32_t prim = tgroup.x + (uint32_t)j;
            if (inst_sp < 0) {
                // TLAS leaf: postpone the rest of this node and enter the instance
                if (tgroup.y != 0u) stack[sp++] = tgroup;
                if (ngroup.y > 0x00FFFFFFu) stack[sp++] = ngroup;
                pending = prim;
                tgroup = make_uint2(0u, 0u);
                ngroup = make_uint2(0u, 0u);
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
        if (pending != kNoInstance) continue;

        if (ngroup.y <= 0x00FFFFFFu) {
            if (inst_sp >= 0 && sp == inst_sp) {
                // BLAS exhausted
                inst_sp = -1;
                if (from_root) { // back to world space
                    o = wo;
                    d = wd;
                    idir = f3(safe_rcp(d.x), safe_rcp(d.y), safe_rcp(d.z));
                    inv_dd = fast_rcp(dot3_fma(d, d));
                    nodes = sc.tlas_nodes;
                }
            }
            if (sp == 0) {
                if (from_root) break;
                // next candidate whose world box the ray segment meets (the same slab test, with the same
                // outward padding, that the TLAS descent applies to the quantised leaf box)
                while (ci < n_cand) {
                    const uint32_t id = cand[ci * cand_stride];
                    ci++;
                    const float4 blo = __ldg(sc.inst_boxes + 2 * id), bhi = __ldg(sc.inst_boxes + 2 * id + 1);
                    // (b - o) first: exact to half an ulp even when the origin sits on the box
                    const float tx0 = (blo.x - wo.x) * widir.x, tx1 = (bhi.x - wo.x) * widir.x;
                    const float ty0 = (blo.y - wo.y) * widir.y, ty1 = (bhi.y - wo.y) * widir.y;
                    const float tz0 = (blo.z - wo.z) * widir.z, tz1 = (bhi.z - wo.z) * widir.z;
                    const float tn = fmaxf(fmaxf(fminf(tx0, tx1), fminf(ty0, ty1)), fmaxf(fminf(tz0, tz1), tmin));
                    const float tf = fminf(fminf(fmaxf(tx0, tx1), fmaxf(ty0, ty1)), fminf(fmaxf(tz0, tz1), tmax));
                    // slab distances carry a few ulps of relative error: keep everything within 2e-6 relative
                    if (tn - tf <= 2e-6f * fmaxf(fabsf(tn), fabsf(tf))) {
                        pending = id;
                        break;
                    }
                }
                if (pending != kNoInstance) continue;
                break;
            }
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

// traverse.cuh -- two-level stack traversal of the 8-wide BVH (device code).
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
    uint32_t slot; // any-hit calls: index of the triangle that was hit in its BLAS's leaf-order triangle array
    float bu, bv;  // barycentric weights of v0, v1
    // input of trace_ray<.., FACE_CULL = true>: +1 / -1 = the sign s_view of the rasteriser view being emulated; a
    // triangle is kept iff s_view * sign(det instance) * dot(d, (v1 - v0) x (v2 - v0)) > 0 (shadow_map.cu)
    float cull_sign;
};

struct LocalStats {
    uint32_t nodes, tris, insts;
    uint32_t tlas_nodes, root_descents; // of `nodes`: visited in the TLAS; descents from the TLAS root started
};

// Reciprocal for slab tests: one MUFU.RCP (relative error <= 2^-23) instead of the ~10-instruction IEEE
// division.  The box test absorbs it: far planes are scaled by kFar = 1 + 2^-20 (see intersect_node).
__device__ __forceinline__ float fast_rcp(float x) {
#ifdef __CUDA_ARCH__
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else // host build of this header (tests/cpu_traverse: the traversal logic checked without a GPU)
    return 1.0f / x;
#endif
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
__device__ __forceinline__ float3 cross3_rn(const float3 a, const float3 b) { // explicitly rounded, never contracted
    return f3(__fsub_rn(__fmul_rn(a.y, b.z), __fmul_rn(a.z, b.y)), __fsub_rn(__fmul_rn(a.z, b.x), __fmul_rn(a.x, b.z)),
              __fsub_rn(__fmul_rn(a.x, b.y), __fmul_rn(a.y, b.x)));
}

// Watertight two-sided ray/triangle test by signed volumes in Pluecker coordinates.  With A = p0 - o etc. the volume
// d . (C x B) of the ray against the edge (p1, p2) expands to d . (p2 x p1) + (p1 - p2) . (o x d): the builder stores
// the edge's moment and direction (WideTri), the ray carries d and its moment m = o x d, and the
// volume is six multiply-adds.  Reversing the edge negates moment and direction bit by bit and every multiply-add
// below is odd in them, so the two triangles sharing an edge see exactly opposite values: no ray slips between them.
// (The previous form translated the vertices by -o and cost 51 instructions up to the sign test; this one 18.)
struct TriData {
    float4 mu, eu, mv, ev, mw, ew;
};
__device__ __forceinline__ TriData load_tri(const WideTri* tri) {
    const float4* tp = reinterpret_cast<const float4*>(tri);
    TriData q;
    q.mu = __ldg(tp + 0), q.eu = __ldg(tp + 1), q.mv = __ldg(tp + 2);
    q.ev = __ldg(tp + 3), q.mw = __ldg(tp + 4), q.ew = __ldg(tp + 5);
    return q;
}
__device__ __forceinline__ float edge_volume(const float3 d, const float3 m, const float4 M, const float4 E) {
    return fmaf(d.x, M.x, fmaf(d.y, M.y, fmaf(d.z, M.z, fmaf(m.x, E.x, fmaf(m.y, E.y, __fmul_rn(m.z, E.z))))));
}
// NEED_T == false (any-hit callers): only "tmin < t < tmax" is decided, by cross-multiplication with the sign of the
// denominator instead of the IEEE division (a dozen instructions executed by the few lanes that sit in a leaf).
template <bool NEED_T = true>
__device__ __forceinline__ bool tri_test(const TriData& q, const float3 o, const float3 d, const float3 m, const float tmin,
                                         const float tmax, float& t_out, float& bu, float& bv) {
    const float U = edge_volume(d, m, q.mu, q.eu);
    const float V = edge_volume(d, m, q.mv, q.ev);
    const float W = edge_volume(d, m, q.mw, q.ew);
    const float mn = fminf(U, fminf(V, W)), mx = fmaxf(U, fmaxf(V, W));
    if (mn < 0.0f && mx > 0.0f) return false;
    const float det = U + V + W;
    if (det == 0.0f) return false;
    // the plane: t = N . (p0 - o) / (N . d)
    const float3 N = f3(q.mv.w, q.ev.w, q.mw.w);
    const float num = q.eu.w - dot3_fma(N, o), den = dot3_fma(N, d);
    if (!NEED_T) { // t = num / den lies in (tmin, tmax); den == 0 or NaN operands fail both forms
        const float lo = __fmul_rn(tmin, den), hi = __fmul_rn(tmax, den);
        return den > 0.0f ? (num > lo && num < hi) : (num < lo && num > hi);
    }
    const float t = num / den;
    if (!(t > tmin && t < tmax)) return false;
    t_out = t;
    bu = U / det;
    bv = V / det;
    return true;
}

// ---- ray vs the eight child boxes of a wide node -------------------------------------------------------
// Per ray space (world, or the object space of the instance being traversed) the ray keeps
//     idir = 1 / d (MUFU, relative error 2^-23),  p = o * idir,  npn = -(p + |p| 2^-22),  npf = -(p - |p| 2^-22)
// and the slab distance of a plane b is half an instruction: t = fma(b, idir, np*) is issued as the packed
// FFMA2 of sm_100 (fma.rn.f32x2: two IEEE fp32 FMAs per issue slot, the ray constants broadcast from one
// register, the plane pair exactly as LDG.128 delivered it), because this kernel is bound by instruction issue,
// not by the FMA pipe.  Error budget: the product o * idir is rounded once (<= |p| 2^-24, covered four times over
// by the +-|p| 2^-22 built into npn / npf, which push near planes down and far planes up) and the FMA rounds once
// more (<= |t| 2^-24, which together with the reciprocal's 2^-23 is covered by scaling the far distance by
// 1 + 2^-20, one packed FMUL2 per child pair).  The test is therefore conservative for rays with tmin >= 0 (what
// VK_KHR_ray_query requires of its callers): a ray that meets a box in exact arithmetic is never rejected.
//
// lo and hi planes of an axis are 32 bytes apart in the node, so the near / far selection by the sign of the
// direction is an address offset fixed per ray space (no selects in the per-child code).  Empty slots hold
// lo = +inf, hi = -inf: near = +inf or NaN, and the comparison below is false.
struct RaySpace {
    float3 idir, npn, npf;
    uint32_t off; // byte 0/1/2: 0 or 32 = offset of the NEAR plane array of x / y / z inside its axis pair
};

__device__ __forceinline__ RaySpace make_ray_space(const float3 o, const float3 d) {
    RaySpace rs;
    rs.idir = f3(safe_rcp(d.x), safe_rcp(d.y), safe_rcp(d.z));
    const float px = __fmul_rn(o.x, rs.idir.x), py = __fmul_rn(o.y, rs.idir.y), pz = __fmul_rn(o.z, rs.idir.z);
    const float k = 2.384185791015625e-07f; // 2^-22
    rs.npn = f3(-fmaf(fabsf(px), k, px), -fmaf(fabsf(py), k, py), -fmaf(fabsf(pz), k, pz));
    rs.npf = f3(-fmaf(fabsf(px), -k, px), -fmaf(fabsf(py), -k, py), -fmaf(fabsf(pz), -k, pz));
    rs.off = ((__float_as_uint(rs.idir.x) >> 31) << 5) | ((__float_as_uint(rs.idir.y) >> 31) << 13) |
             ((__float_as_uint(rs.idir.z) >> 31) << 21);
    return rs;
}

// (a0, a1) * b + c and (a0, a1) * b as one packed instruction each; bitwise equal to two fmaf() / __fmul_rn().
__device__ __forceinline__ float2 fma2_bcast(const float a0, const float a1, const float b, const float c) {
#ifdef __CUDA_ARCH__
    unsigned long long A, B, C, D;
    float2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(A) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %1};" : "=l"(B) : "f"(b));
    asm("mov.b64 %0, {%1, %1};" : "=l"(C) : "f"(c));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(D) : "l"(A), "l"(B), "l"(C));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(D));
    return r;
#else
    return make_float2(fmaf(a0, b, c), fmaf(a1, b, c));
#endif
}
__device__ __forceinline__ float2 mul2_bcast(const float a0, const float a1, const float b) {
#ifdef __CUDA_ARCH__
    unsigned long long A, B, D;
    float2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(A) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %1};" : "=l"(B) : "f"(b));
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(D) : "l"(A), "l"(B));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(D));
    return r;
#else
    return make_float2(a0 * b, a1 * b);
#endif
}

// Returns the 8-bit mask of child slots whose box the ray overlaps in [tmin, tmax].
__device__ __forceinline__ uint32_t intersect_node(const WideNode* __restrict__ node, const RaySpace& rs, const float tmin,
                                                   const float tmax) {
    const char* base = reinterpret_cast<const char*>(node) + 16;
    const uint32_t ox = rs.off & 0xFFu, oy = (rs.off >> 8) & 0xFFu, oz = (rs.off >> 16) & 0xFFu;
    const float4* nxp = reinterpret_cast<const float4*>(base + ox);
    const float4* fxp = reinterpret_cast<const float4*>(base + (ox ^ 32u));
    const float4* nyp = reinterpret_cast<const float4*>(base + 64 + oy);
    const float4* fyp = reinterpret_cast<const float4*>(base + 64 + (oy ^ 32u));
    const float4* nzp = reinterpret_cast<const float4*>(base + 128 + oz);
    const float4* fzp = reinterpret_cast<const float4*>(base + 128 + (oz ^ 32u));
    const float kFar = 1.00000095367431640625f; // 1 + 2^-20
    uint32_t slots = 0;
#pragma unroll
    for (int half = 0; half < 2; half++) {
        const float4 nx = __ldg(nxp + half), ny = __ldg(nyp + half), nz = __ldg(nzp + half);
        const float4 fx = __ldg(fxp + half), fy = __ldg(fyp + half), fz = __ldg(fzp + half);
        const float anx[4] = {nx.x, nx.y, nx.z, nx.w}, any[4] = {ny.x, ny.y, ny.z, ny.w}, anz[4] = {nz.x, nz.y, nz.z, nz.w};
        const float afx[4] = {fx.x, fx.y, fx.z, fx.w}, afy[4] = {fy.x, fy.y, fy.z, fy.w}, afz[4] = {fz.x, fz.y, fz.z, fz.w};
#pragma unroll
        for (int j = 0; j < 4; j += 2) {
            const float2 tnx = fma2_bcast(anx[j], anx[j + 1], rs.idir.x, rs.npn.x);
            const float2 tny = fma2_bcast(any[j], any[j + 1], rs.idir.y, rs.npn.y);
            const float2 tnz = fma2_bcast(anz[j], anz[j + 1], rs.idir.z, rs.npn.z);
            const float2 tfx = fma2_bcast(afx[j], afx[j + 1], rs.idir.x, rs.npf.x);
            const float2 tfy = fma2_bcast(afy[j], afy[j + 1], rs.idir.y, rs.npf.y);
            const float2 tfz = fma2_bcast(afz[j], afz[j + 1], rs.idir.z, rs.npf.z);
            const float cmin0 = fmaxf(fmaxf(tnx.x, tny.x), fmaxf(tnz.x, tmin));
            const float cmin1 = fmaxf(fmaxf(tnx.y, tny.y), fmaxf(tnz.y, tmin));
            const float cmax0 = fminf(fminf(tfx.x, tfy.x), fminf(tfz.x, tmax));
            const float cmax1 = fminf(fminf(tfx.y, tfy.y), fminf(tfz.y, tmax));
            const float2 cfar = mul2_bcast(cmax0, cmax1, kFar);
            if (cmin0 <= cfar.x) slots |= 1u << (4 * half + j);
            if (cmin1 <= cfar.y) slots |= 1u << (4 * half + j + 1);
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
// caller then falls back to the root descent).  Child boxes are the exact fp32 boxes the builder stored.
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
            const WideNode* node = sc.tlas_nodes + (ngroup.x + rel);
            const uint4 hdr = __ldg(reinterpret_cast<const uint4*>(node));
            if (STATS) st->nodes++;
            // exact child boxes against the query box (empty slots hold an inverted infinite box)
            const float4* bp = reinterpret_cast<const float4*>(reinterpret_cast<const char*>(node) + 16);
            uint32_t slots = 0;
#pragma unroll
            for (int half = 0; half < 2; half++) {
                const float4 lx = __ldg(bp + half), hx = __ldg(bp + 2 + half), ly = __ldg(bp + 4 + half);
                const float4 hy = __ldg(bp + 6 + half), lz = __ldg(bp + 8 + half), hz = __ldg(bp + 10 + half);
                const float alx[4] = {lx.x, lx.y, lx.z, lx.w}, ahx[4] = {hx.x, hx.y, hx.z, hx.w};
                const float aly[4] = {ly.x, ly.y, ly.z, ly.w}, ahy[4] = {hy.x, hy.y, hy.z, hy.w};
                const float alz[4] = {lz.x, lz.y, lz.z, lz.w}, ahz[4] = {hz.x, hz.y, hz.z, hz.w};
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const bool ov = alx[j] <= hi.x && ahx[j] >= lo.x && aly[j] <= hi.y && ahy[j] >= lo.y &&
                                    alz[j] <= hi.z && ahz[j] >= lo.z;
                    if (ov) slots |= 1u << (4 * half + j);
                }
            }
            const uint32_t node_imask = hdr.x >> 24;
            uint32_t prims = leaf_bits(slots & ~node_imask, hdr.z, hdr.w);
            while (prims) {
                const int j = __ffs(prims) - 1;
                prims &= prims - 1u;
                if (n >= max_out) return -1;
                out[n * stride] = hdr.y + (uint32_t)j;
                n++;
            }
            ngroup = make_uint2(hdr.x & 0x00FFFFFFu, ((slots & node_imask) << 24) | node_imask);
        }
        if (sp == 0) break;
        ngroup = stack[--sp];
    }
    return n;
}

// ---- hemisphere reach ---------------------------------------------------------------------------------
// The AO directions of a pixel are T h.x + B h.y + C h.z with h on the upper unit hemisphere (h.z >= 0,
// light.frag:63-69) and 0 <= aoMin < t < aoMax, so along an axis whose frame coefficients are a = (T_k, B_k, C_k)
// the segment end points stay inside O_k + reach * [dmin, dmax]:  dmax = |a| if a.z >= 0 (the maximiser a / |a| lies
// on the hemisphere), else the rim value |a.xy|;  dmin likewise with the signs flipped.  For a pixel on a flat face
// this box starts AT the (biased) origin instead of reaching one aoMax below the surface, which is what keeps the
// pixel's own instance out of the candidate list.
__device__ __forceinline__ void hemisphere_axis(const float ax, const float ay, const float az, float& dmin, float& dmax) {
    const float r2 = fmaf(ax, ax, ay * ay);
    const float len = sqrtf(fmaf(az, az, r2)), rim = sqrtf(r2);
    dmax = az >= 0.0f ? len : rim;
    dmin = az <= 0.0f ? -len : -rim;
}
__device__ __forceinline__ void hemisphere_box(const float3 O, const float3 T, const float3 B, const float3 C,
                                               const float reach, const float3 slack, float3& lo, float3& hi) {
    float mn, mx;
    hemisphere_axis(T.x, B.x, C.x, mn, mx);
    lo.x = fmaf(reach, mn, O.x) - slack.x, hi.x = fmaf(reach, mx, O.x) + slack.x;
    hemisphere_axis(T.y, B.y, C.y, mn, mx);
    lo.y = fmaf(reach, mn, O.y) - slack.y, hi.y = fmaf(reach, mx, O.y) + slack.y;
    hemisphere_axis(T.z, B.z, C.z, mn, mx);
    lo.z = fmaf(reach, mn, O.z) - slack.z, hi.z = fmaf(reach, mx, O.z) + slack.z;
}

// Drops from a candidate list (collect_instances) every instance whose BLAS the pixel's AO rays cannot reach: the
// hemisphere frame goes to the instance's object space (the map is linear, so the hemisphere reach argument holds
// there unchanged), and the reach box is tested against the exact fp32 boxes of the BLAS root's eight children.
// Every triangle lies inside one of them, so a dropped instance holds no triangle any of the pixel's rays can hit:
// visibility is unchanged, and each of the pixel's rays is spared the instance entry and the root visit (C3: 0.9
// entries per AO ray before, almost all of them into the instance the pixel itself lies on).  The slack covers the
// rounding of the transform (the rays are transformed one by one, with their own rounding, in trace_ray).
// Returns the new count; NaN boxes (singular instance matrices) keep the candidate.
template <bool STATS>
__device__ __forceinline__ int filter_candidates(const TraceScene& sc, const float3 O, const float3 T, const float3 B,
                                                 const float3 C, const float reach, uint32_t* cand, const int stride,
                                                 const int n_cand, LocalStats* st) {
    int kept = 0;
    for (int k = 0; k < n_cand; k++) {
        const uint32_t id = cand[k * stride];
        const InstanceRec* rec = sc.instances + id;
        const float4 r0 = __ldg(&rec->r0), r1 = __ldg(&rec->r1), r2 = __ldg(&rec->r2);
        const uint4 ptrs = __ldg(reinterpret_cast<const uint4*>(&rec->nodes));
        const float3 Oo = xform_point(r0, r1, r2, O);
        const float3 To = xform_dir(r0, r1, r2, T), Bo = xform_dir(r0, r1, r2, B), Co = xform_dir(r0, r1, r2, C);
        // |terms| of the point transform: its rounding error is <= 4 * 2^-24 of their sum
        const float3 mag = f3(fmaf(fabsf(r0.x), fabsf(O.x), fmaf(fabsf(r0.y), fabsf(O.y), fmaf(fabsf(r0.z), fabsf(O.z), fabsf(r0.w)))),
                              fmaf(fabsf(r1.x), fabsf(O.x), fmaf(fabsf(r1.y), fabsf(O.y), fmaf(fabsf(r1.z), fabsf(O.z), fabsf(r1.w)))),
                              fmaf(fabsf(r2.x), fabsf(O.x), fmaf(fabsf(r2.y), fabsf(O.y), fmaf(fabsf(r2.z), fabsf(O.z), fabsf(r2.w)))));
        float3 lo, hi;
        hemisphere_box(Oo, To, Bo, Co, reach, f3(2e-6f * mag.x + 1e-30f, 2e-6f * mag.y + 1e-30f, 2e-6f * mag.z + 1e-30f), lo, hi);
        bool keep = !(lo.x <= hi.x && lo.y <= hi.y && lo.z <= hi.z); // NaN: leave it to trace_ray
        if (!keep) {
            const WideNode* root = reinterpret_cast<const WideNode*>(((unsigned long long)ptrs.y << 32) | ptrs.x);
            const float4* bp = reinterpret_cast<const float4*>(reinterpret_cast<const char*>(root) + 16);
            if (STATS) st->nodes++;
#pragma unroll
            for (int half = 0; half < 2; half++) {
                const float4 lx = __ldg(bp + half), hx = __ldg(bp + 2 + half), ly = __ldg(bp + 4 + half);
                const float4 hy = __ldg(bp + 6 + half), lz = __ldg(bp + 8 + half), hz = __ldg(bp + 10 + half);
                const float alx[4] = {lx.x, lx.y, lx.z, lx.w}, ahx[4] = {hx.x, hx.y, hx.z, hx.w};
                const float aly[4] = {ly.x, ly.y, ly.z, ly.w}, ahy[4] = {hy.x, hy.y, hy.z, hy.w};
                const float alz[4] = {lz.x, lz.y, lz.z, lz.w}, ahz[4] = {hz.x, hz.y, hz.z, hz.w};
#pragma unroll
                for (int j = 0; j < 4; j++)
                    keep |= alx[j] <= hi.x && ahx[j] >= lo.x && aly[j] <= hi.y && ahy[j] >= lo.y && alz[j] <= hi.z && ahz[j] >= lo.z;
            }
        }
        if (keep) {
            cand[kept * stride] = id;
            kept++;
        }
    }
    return kept;
}

// CLOSEST == false: any-hit, returns true at the first committed intersection.
// CLOSEST == true : returns true if something was hit; *hit describes the nearest one.
// n_cand < 0: descend from the TLAS root.  n_cand >= 0: visit only the instances cand[k * cand_stride]
// (TLAS leaf order indices from collect_instances); with then_root the candidates are only tried FIRST (occluder
// hints of shadow rays: an any-hit result does not depend on the order) and the root descent follows if none of
// them is hit.  Any-hit calls may pass `hit` to learn the instance that was hit (hit->inst).
// `stack` is LUZ_STACK_SIZE entries of caller storage.
constexpr uint32_t kNoInstance = 0xFFFFFFFFu;

// ONE_VISIT: the node phase is a plain `if` (one node visit per pass of the loop) instead of the loop that yields
// through sc.min_node_lanes.  Both forms visit one node per pass with the default min_node_lanes = 33, but the
// compiler places the warp's re-join points differently, and which form is faster depends on the kernel around it
// (profiles/r1_ab_onepass.md): the ray kernels pick per instantiation.
template <bool CLOSEST, bool STATS, bool FACE_CULL = false, bool ONE_VISIT = false>
__device__ __forceinline__ bool trace_ray(const TraceScene& sc, const float3 wo, const float3 wd, const float tmin,
                                          float tmax, HitInfo* hit, LocalStats* st, uint2* stack,
                                          const uint32_t* cand = nullptr, const int cand_stride = 0,
                                          const int n_cand = -1, const bool then_root = false) {
    // rays with NaNs (e.g. the vertical-light tangent of light.frag:90) and null directions miss
    if (!(wo.x == wo.x && wo.y == wo.y && wo.z == wo.z && wd.x == wd.x && wd.y == wd.y && wd.z == wd.z &&
          tmin == tmin && tmax == tmax))
        return false;
    if (wd.x == 0.0f && wd.y == 0.0f && wd.z == 0.0f) return false;

    int sp = 0;
    int inst_sp = -1; // stack height at which the current instance was entered, -1 = in the TLAS
    uint32_t cur_inst = 0;
    bool from_root = n_cand < 0;

    float3 o = wo, d = wd;
    RaySpace rs;
    const WideNode* nodes = sc.tlas_nodes;
    const WideTri* tris = nullptr;
    bool found = false;

    uint2 ngroup = make_uint2(0u, 0u);
    uint2 tgroup = make_uint2(0u, 0u);
    uint32_t pending = kNoInstance; // instance to enter at the top of the loop
    int ci = 0;
    float face_sign = 0.0f; // FACE_CULL: cull_sign * sign(det of the current instance's matrix)
    float3 widir; // candidate mode keeps the world-space reciprocal direction for the box pre-test below
    if (from_root) {
        if (STATS) st->root_descents++;
        rs = make_ray_space(o, d);
        widir = rs.idir;
        ngroup = make_uint2(0u, 0x80000000u); // root: slot 7 of a virtual parent with imask 0
    } else { // no world-space node is ever tested: the ray space is set up per instance
        widir = f3(safe_rcp(d.x), safe_rcp(d.y), safe_rcp(d.z));
        rs.idir = widir;
        rs.npn = rs.npf = f3(0.0f, 0.0f, 0.0f);
        rs.off = 0u;
    }

    while (true) {
        if (pending != kNoInstance) {
            // enter an instance: the ray goes to object space, t is preserved (direction not re-normalised)
            const InstanceRec* rec = sc.instances + pending;
            const float4 r0 = __ldg(&rec->r0), r1 = __ldg(&rec->r1), r2 = __ldg(&rec->r2);
            const uint4 ptrs = __ldg(reinterpret_cast<const uint4*>(&rec->nodes));
            if (STATS) st->insts++;
            o = xform_point(r0, r1, r2, wo);
            d = xform_dir(r0, r1, r2, wd);
            rs = make_ray_space(o, d);
            if (FACE_CULL) {
                // det of the world->object rows has the sign of det of the instance matrix
                const float det = r0.x * (r1.y * r2.z - r1.z * r2.y) - r0.y * (r1.x * r2.z - r1.z * r2.x) +
                                  r0.z * (r1.x * r2.y - r1.y * r2.x);
                face_sign = det < 0.0f ? -hit->cull_sign : hit->cull_sign;
            }
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
            const WideNode* node = nodes + (ngroup.x + rel);
            const uint4 hdr = __ldg(reinterpret_cast<const uint4*>(node));
            if (STATS) {
                st->nodes++;
                if (inst_sp < 0) st->tlas_nodes++;
            }
            const uint32_t slots = intersect_node(node, rs, tmin, tmax);
            const uint32_t node_imask = hdr.x >> 24;
            ngroup = make_uint2(hdr.x & 0x00FFFFFFu, ((slots & node_imask) << 24) | node_imask);
            tgroup = make_uint2(hdr.y, leaf_bits(slots & ~node_imask, hdr.z, hdr.w));
            // when only a few lanes are still descending, yield so that the lanes waiting at the end of this loop
            // (they hold primitives or need a pop) get served and everybody re-enters the node test together
            if (ONE_VISIT) break;
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
                const TriData q = load_tri(tris + prim);
                if (STATS) st->tris++;
                float t, bu, bv;
                // the ray's moment is recomputed here rather than kept alive across the node loop (registers)
                if (tri_test<CLOSEST>(q, o, d, cross3_rn(o, d), tmin, tmax, t, bu, bv)) {
                    if (FACE_CULL) { // front faces (in the framebuffer of the emulated view) are not rasterised
                        const float3 n = f3(q.mv.w, q.ev.w, q.mw.w); // (p1 - p0) x (p2 - p0)
                        if (!(face_sign * dot3(d, n) > 0.0f)) continue;
                    }
                    if (!CLOSEST) {
                        if (hit) {
                            hit->inst = cur_inst;
                            hit->slot = prim;
                        }
                        return true;
                    }
                    tmax = t;
                    found = true;
                    hit->t = t;
                    hit->inst = cur_inst;
                    hit->prim = __float_as_uint(q.mu.w);
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
                    rs = make_ray_space(o, d);
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
                if (then_root && !from_root) { // none of the hinted instances was hit: the ordinary descent
                    if (STATS) st->root_descents++;
                    from_root = true;
                    o = wo;
                    d = wd;
                    rs = make_ray_space(o, d);
                    nodes = sc.tlas_nodes;
                    ngroup = make_uint2(0u, 0x80000000u);
                    continue;
                }
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

// Any-hit against ONE BLAS, for callers that already hold the object-space ray (the AO kernel: all rays of a pixel share
// their origin and their short candidate list, so the instance record is loaded and the origin transformed once per
// pixel and candidate, and the loop below carries no TLAS / candidate state).  Same node test, leaf expansion and
// triangle test as trace_ray, same arithmetic on the same operands.  The caller has checked the ray for NaNs and a
// null direction.  `stack` is LUZ_STACK_SIZE entries of caller storage.
template <bool ONE_VISIT = true>
__device__ __forceinline__ bool trace_blas_any(const WideNode* __restrict__ nodes, const WideTri* __restrict__ tris, const float3 o,
                                               const float3 d, const float tmin, const float tmax, uint2* stack) {
    const RaySpace rs = make_ray_space(o, d);
    int sp = 0;
    uint2 ngroup = make_uint2(0u, 0x80000000u); // root: slot 7 of a virtual parent with imask 0
    uint2 tgroup = make_uint2(0u, 0u);
    while (true) {
        if (ngroup.y > 0x00FFFFFFu && tgroup.y == 0u) {
            const uint32_t hits = ngroup.y;
            const uint32_t imask = hits & 0xFFu;
            const int bit = 31 - __clz(hits);
            ngroup.y &= ~(1u << bit);
            if (ngroup.y > 0x00FFFFFFu) stack[sp++] = ngroup;
            const int slot = bit - 24;
            const uint32_t rel = __popc(imask & ~(0xFFFFFFFFu << slot));
            const WideNode* node = nodes + (ngroup.x + rel);
            const uint4 hdr = __ldg(reinterpret_cast<const uint4*>(node));
            const uint32_t slots = intersect_node(node, rs, tmin, tmax);
            const uint32_t node_imask = hdr.x >> 24;
            ngroup = make_uint2(hdr.x & 0x00FFFFFFu, ((slots & node_imask) << 24) | node_imask);
            tgroup = make_uint2(hdr.y, leaf_bits(slots & ~node_imask, hdr.z, hdr.w));
        }
        while (tgroup.y != 0u) {
            const int j = __ffs(tgroup.y) - 1;
            tgroup.y &= tgroup.y - 1u;
            const TriData q = load_tri(tris + tgroup.x + (uint32_t)j);
            float t, bu, bv;
            if (tri_test<false>(q, o, d, cross3_rn(o, d), tmin, tmax, t, bu, bv)) return true;
        }
        if (ngroup.y <= 0x00FFFFFFu) {
            if (sp == 0) return false;
            ngroup = stack[--sp];
        }
    }
}

} // namespace luz

// harness.cpp -- luz_b200/csrc/traverse.cuh compiled for the HOST (g++, no nvcc, no GPU) and run against an
// exhaustive double-precision ray/triangle test: the traversal LOGIC of the CUDA kernels (stack handling, two-level
// descent, candidate lists, hemisphere filter, occluder hints = then_root, closest hit, Pluecker triangle test) can be
// checked, and edited, without a B200.  TEST INFRASTRUCTURE (tests/test_cpu_traverse.py builds and runs it); the
// product never runs this code on the CPU.  The device intrinsics the header uses are shimmed below; its three inline
// PTX sequences have host branches (#ifdef __CUDA_ARCH__).  The wide BVHs are built here by a simple median splitter
// that emits the same node format as csrc/bvh_build.cu (common.cuh WideNode / WideTri / InstanceRec).
//
// usage: harness <seed> [instances=40] [rays=6000] [extent=6]   -> one JSON line with agreement counts
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <random>
#include <vector>

// ---- shims for device intrinsics (host semantics are exact for all of them under -ffp-contract=off) ----
template <class T>
static inline T __ldg(const T* p) { return *p; }
static inline int __ffs(unsigned x) { return __builtin_ffs((int)x); }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __clz(unsigned x) { return x ? __builtin_clz(x) : 32; }
static inline unsigned __activemask() { return 1u; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
using std::isnan;

#include "../../luz_b200/csrc/traverse.cuh"

using namespace luz;

struct Box {
    float lo[3], hi[3];
    void init() { for (int k = 0; k < 3; k++) lo[k] = std::numeric_limits<float>::infinity(), hi[k] = -lo[k]; }
    void grow(const float* p) { for (int k = 0; k < 3; k++) lo[k] = std::min(lo[k], p[k]), hi[k] = std::max(hi[k], p[k]); }
    void grow(const Box& b) { grow(b.lo); grow(b.hi); }
};

// ---- median-split builder emitting WideNode (208 B) trees ----------------------------------------------------
struct WideTree {
    std::vector<WideNode> nodes;
    std::vector<uint32_t> order; // leaf order -> input primitive
    int levels = 0;
};

static void split_groups(std::vector<uint32_t>& prims, const std::vector<Box>& boxes, size_t max_leaf,
                         std::vector<std::vector<uint32_t>>& out) {
    std::vector<std::vector<uint32_t>> work{prims};
    while (work.size() < 8) { // split the largest group that is still too big for a leaf
        size_t best = work.size();
        for (size_t i = 0; i < work.size(); i++)
            if (work[i].size() > max_leaf && (best == work.size() || work[i].size() > work[best].size())) best = i;
        if (best == work.size()) break;
        std::vector<uint32_t> g = work[best];
        Box cb; cb.init();
        for (uint32_t p : g) { float c[3]; for (int k = 0; k < 3; k++) c[k] = 0.5f * (boxes[p].lo[k] + boxes[p].hi[k]); cb.grow(c); }
        int axis = 0;
        for (int k = 1; k < 3; k++) if (cb.hi[k] - cb.lo[k] > cb.hi[axis] - cb.lo[axis]) axis = k;
        std::sort(g.begin(), g.end(), [&](uint32_t a, uint32_t b) {
            const float ca = boxes[a].lo[axis] + boxes[a].hi[axis], cbb = boxes[b].lo[axis] + boxes[b].hi[axis];
            return ca < cbb || (ca == cbb && a < b);
        });
        const size_t half = g.size() / 2;
        work[best] = std::vector<uint32_t>(g.begin(), g.begin() + (ptrdiff_t)half);
        work.push_back(std::vector<uint32_t>(g.begin() + (ptrdiff_t)half, g.end()));
    }
    out = work;
}

static WideTree build_tree(const std::vector<Box>& boxes, size_t max_leaf) {
    WideTree t;
    struct Item { std::vector<uint32_t> prims; int level; };
    std::vector<Item> queue;
    std::vector<uint32_t> all(boxes.size());
    for (size_t i = 0; i < all.size(); i++) all[i] = (uint32_t)i;
    queue.push_back({all, 1});
    for (size_t q = 0; q < queue.size(); q++) {
        Item item = queue[q];
        t.levels = std::max(t.levels, item.level);
        std::vector<std::vector<uint32_t>> groups;
        split_groups(item.prims, boxes, max_leaf, groups);
        WideNode n;
        memset(&n, 0, sizeof n);
        for (int i = 0; i < 8; i++) {
            n.lox[i] = n.loy[i] = n.loz[i] = std::numeric_limits<float>::infinity();
            n.hix[i] = n.hiy[i] = n.hiz[i] = -std::numeric_limits<float>::infinity();
        }
        uint32_t imask = 0, off = 0;
        const uint32_t child_base = (uint32_t)queue.size();
        n.prim_base = (uint32_t)t.order.size();
        for (size_t i = 0; i < groups.size() && i < 8; i++) {
            if (groups[i].empty()) continue;
            Box b; b.init();
            for (uint32_t p : groups[i]) b.grow(boxes[p]);
            n.lox[i] = b.lo[0], n.loy[i] = b.lo[1], n.loz[i] = b.lo[2];
            n.hix[i] = b.hi[0], n.hiy[i] = b.hi[1], n.hiz[i] = b.hi[2];
            if (groups[i].size() > max_leaf) { // internal child
                imask |= 1u << i;
                n.meta[i] = (uint8_t)((1u << 5) | (24u + (uint32_t)i));
                queue.push_back({groups[i], item.level + 1});
            } else {
                const uint32_t cnt = (uint32_t)groups[i].size();
                n.meta[i] = (uint8_t)((((1u << cnt) - 1u) << 5) | off);
                for (uint32_t p : groups[i]) t.order.push_back(p);
                off += cnt;
            }
        }
        n.child_base_imask = (child_base & 0x00FFFFFFu) | (imask << 24);
        t.nodes.push_back(n);
    }
    return t;
}

// ---- scene ------------------------------------------------------------------------------------------------------
struct Mesh {
    std::vector<float> pos; // 9 floats per triangle
    WideTree tree;
    std::vector<WideTri> tris;
    Box bounds;
};
static float3 crs(float3 q, float3 p) { return f3(q.y * p.z - q.z * p.y, q.z * p.x - q.x * p.z, q.x * p.y - q.y * p.x); }
static void prepare(Mesh& m) { // what bvh_build.cu k_gather_triangles stores
    const size_t nt = m.pos.size() / 9;
    std::vector<Box> boxes(nt);
    m.bounds.init();
    for (size_t t = 0; t < nt; t++) {
        boxes[t].init();
        for (int v = 0; v < 3; v++) boxes[t].grow(&m.pos[t * 9 + (size_t)v * 3]);
        m.bounds.grow(boxes[t]);
    }
    m.tree = build_tree(boxes, 3);
    for (uint32_t src : m.tree.order) {
        const float* p = &m.pos[(size_t)src * 9];
        const float3 a = f3(p[0], p[1], p[2]), b = f3(p[3], p[4], p[5]), c = f3(p[6], p[7], p[8]);
        const float3 mu = crs(c, b), mv = crs(a, c), mw = crs(b, a);
        const float3 eu = b - c, ev = c - a, ew = a - b, n = crs(b - a, c - a);
        const float k = (n.x * a.x + n.y * a.y) + n.z * a.z;
        WideTri w;
        w.mu = make_float4(mu.x, mu.y, mu.z, __uint_as_float(src));
        w.eu = make_float4(eu.x, eu.y, eu.z, k);
        w.mv = make_float4(mv.x, mv.y, mv.z, n.x);
        w.ev = make_float4(ev.x, ev.y, ev.z, n.y);
        w.mw = make_float4(mw.x, mw.y, mw.z, n.z);
        w.ew = make_float4(ew.x, ew.y, ew.z, 0.0f);
        m.tris.push_back(w);
    }
}

struct Instance {
    int mesh;
    double m[12], inv[12]; // rows of the 3x4 object->world and world->object maps
};
static void invert(const double* m, double* inv) {
    const double a = m[0], b = m[1], c = m[2], d = m[4], e = m[5], f = m[6], g = m[8], h = m[9], i = m[10];
    const double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
    const double r[9] = {(e * i - f * h) / det, (c * h - b * i) / det, (b * f - c * e) / det, (f * g - d * i) / det, (a * i - c * g) / det,
                         (c * d - a * f) / det, (d * h - e * g) / det, (b * g - a * h) / det, (a * e - b * d) / det};
    for (int row = 0; row < 3; row++) {
        for (int col = 0; col < 3; col++) inv[row * 4 + col] = r[row * 3 + col];
        inv[row * 4 + 3] = -(r[row * 3] * m[3] + r[row * 3 + 1] * m[7] + r[row * 3 + 2] * m[11]);
    }
}

// exhaustive reference: every triangle of every instance in world space, double precision, two-sided
static bool exhaustive(const std::vector<Mesh>& meshes, const std::vector<Instance>& inst, const double* o, const double* d,
                       double tmin, double tmax, double* t_closest, double* margin) {
    bool any = false;
    double best = tmax, closest_edge = 1e30;
    for (const Instance& in : inst) {
        const Mesh& me = meshes[(size_t)in.mesh];
        for (size_t t = 0; t < me.pos.size() / 9; t++) {
            double P[3][3];
            for (int v = 0; v < 3; v++)
                for (int r = 0; r < 3; r++)
                    P[v][r] = in.m[r * 4] * me.pos[t * 9 + (size_t)v * 3] + in.m[r * 4 + 1] * me.pos[t * 9 + (size_t)v * 3 + 1] +
                              in.m[r * 4 + 2] * me.pos[t * 9 + (size_t)v * 3 + 2] + in.m[r * 4 + 3];
            double e1[3], e2[3], pv[3], tv[3], qv[3];
            for (int k = 0; k < 3; k++) e1[k] = P[1][k] - P[0][k], e2[k] = P[2][k] - P[0][k], tv[k] = o[k] - P[0][k];
            pv[0] = d[1] * e2[2] - d[2] * e2[1], pv[1] = d[2] * e2[0] - d[0] * e2[2], pv[2] = d[0] * e2[1] - d[1] * e2[0];
            const double det = e1[0] * pv[0] + e1[1] * pv[1] + e1[2] * pv[2];
            if (det == 0.0) continue;
            const double u = (tv[0] * pv[0] + tv[1] * pv[1] + tv[2] * pv[2]) / det;
            qv[0] = tv[1] * e1[2] - tv[2] * e1[1], qv[1] = tv[2] * e1[0] - tv[0] * e1[2], qv[2] = tv[0] * e1[1] - tv[1] * e1[0];
            const double v = (d[0] * qv[0] + d[1] * qv[1] + d[2] * qv[2]) / det;
            const double tt = (e2[0] * qv[0] + e2[1] * qv[1] + e2[2] * qv[2]) / det;
            // distance (in barycentric units / t units) from the decision boundaries: rays closer than 1e-4 are "grazing"
            const double edge = std::min(std::min(std::fabs(u), std::fabs(v)), std::fabs(1.0 - u - v));
            const double tedge = std::min(std::fabs(tt - tmin), std::fabs(tt - tmax)) / std::max(1.0, std::fabs(tmax));
            if (u >= -1e-4 && v >= -1e-4 && u + v <= 1.0 + 1e-4 && tt > tmin - 1e-4 && tt < tmax + 1e-4) closest_edge = std::min(closest_edge, std::min(edge, tedge));
            if (u >= 0 && v >= 0 && u + v <= 1 && tt > tmin && tt < tmax) {
                any = true;
                best = std::min(best, tt);
            }
        }
    }
    *t_closest = best;
    *margin = closest_edge;
    return any;
}

int main(int argc, char** argv) {
    const unsigned seed = argc > 1 ? (unsigned)atoi(argv[1]) : 1u;
    const int n_inst = argc > 2 ? atoi(argv[2]) : 40, n_rays = argc > 3 ? atoi(argv[3]) : 6000;
    const double extent = argc > 4 ? atof(argv[4]) : 6.0;
    std::mt19937 rng(seed);
    auto uni = [&](double a, double b) { return std::uniform_real_distribution<double>(a, b)(rng); };
    // meshes: a unit cube, a triangle soup, a displaced grid
    std::vector<Mesh> meshes(3);
    {
        static const float c[8][3] = {{-1, -1, -1}, {1, -1, -1}, {1, 1, -1}, {-1, 1, -1}, {-1, -1, 1}, {1, -1, 1}, {1, 1, 1}, {-1, 1, 1}};
        static const int q[6][4] = {{0, 3, 2, 1}, {4, 5, 6, 7}, {0, 1, 5, 4}, {2, 3, 7, 6}, {1, 2, 6, 5}, {0, 4, 7, 3}};
        for (auto& f : q)
            for (int t = 0; t < 2; t++) {
                const int idx[3] = {f[0], f[t + 1], f[t + 2]};
                for (int v : idx) for (int k = 0; k < 3; k++) meshes[0].pos.push_back(c[v][k]);
            }
    }
    for (int t = 0; t < 160; t++) {
        const double cx = uni(-1, 1), cy = uni(-1, 1), cz = uni(-1, 1);
        for (int v = 0; v < 3; v++) { meshes[1].pos.push_back((float)(cx + uni(-.3, .3))); meshes[1].pos.push_back((float)(cy + uni(-.3, .3))); meshes[1].pos.push_back((float)(cz + uni(-.3, .3))); }
    }
    {
        const int n = 12;
        std::vector<float> hgt((size_t)(n + 1) * (n + 1));
        for (auto& hh : hgt) hh = (float)uni(-0.15, 0.15);
        auto P = [&](int i, int j, float* out) { out[0] = -1.0f + 2.0f * i / n; out[1] = hgt[(size_t)j * (n + 1) + i]; out[2] = -1.0f + 2.0f * j / n; };
        for (int j = 0; j < n; j++)
            for (int i = 0; i < n; i++) {
                float a[3], b[3], c2[3], d2[3];
                P(i, j, a); P(i + 1, j, b); P(i + 1, j + 1, c2); P(i, j + 1, d2);
                const float* tri[6] = {a, b, c2, a, c2, d2};
                for (auto* v : tri) for (int k = 0; k < 3; k++) meshes[2].pos.push_back(v[k]);
            }
    }
    for (auto& m : meshes) prepare(m);
    // instances: random rotation about a random axis, anisotropic scale (one mirrored), translation in a 12^3 box
    std::vector<Instance> inst;
    for (int i = 0; i < n_inst; i++) {
        Instance in;
        in.mesh = i % 3;
        double ax[3] = {uni(-1, 1), uni(-1, 1), uni(-1, 1)};
        const double al = std::sqrt(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]) + 1e-9;
        for (double& a : ax) a /= al;
        const double ang = uni(0, 6.28), cs = std::cos(ang), sn = std::sin(ang);
        const double sc[3] = {uni(0.4, 1.6) * (i == 7 ? -1 : 1), uni(0.4, 1.6), uni(0.4, 1.6)};
        const double R[9] = {cs + ax[0] * ax[0] * (1 - cs), ax[0] * ax[1] * (1 - cs) - ax[2] * sn, ax[0] * ax[2] * (1 - cs) + ax[1] * sn,
                             ax[1] * ax[0] * (1 - cs) + ax[2] * sn, cs + ax[1] * ax[1] * (1 - cs), ax[1] * ax[2] * (1 - cs) - ax[0] * sn,
                             ax[2] * ax[0] * (1 - cs) - ax[1] * sn, ax[2] * ax[1] * (1 - cs) + ax[0] * sn, cs + ax[2] * ax[2] * (1 - cs)};
        for (int r = 0; r < 3; r++) {
            for (int c2 = 0; c2 < 3; c2++) in.m[r * 4 + c2] = R[r * 3 + c2] * sc[c2];
            in.m[r * 4 + 3] = uni(-extent, extent);
        }
        invert(in.m, in.inv);
        inst.push_back(in);
    }
    // TLAS over the world boxes of the instances (8 corners of the BLAS bounds)
    std::vector<Box> wboxes(inst.size());
    for (size_t i = 0; i < inst.size(); i++) {
        const Box& b = meshes[(size_t)inst[i].mesh].bounds;
        wboxes[i].init();
        for (int c2 = 0; c2 < 8; c2++) {
            const double p[3] = {(c2 & 1) ? b.hi[0] : b.lo[0], (c2 & 2) ? b.hi[1] : b.lo[1], (c2 & 4) ? b.hi[2] : b.lo[2]};
            float w[3];
            for (int r = 0; r < 3; r++) {
                const double x = inst[i].m[r * 4] * p[0] + inst[i].m[r * 4 + 1] * p[1] + inst[i].m[r * 4 + 2] * p[2] + inst[i].m[r * 4 + 3];
                w[r] = (float)x;
            }
            float lo[3], hi[3];
            for (int r = 0; r < 3; r++) lo[r] = std::nextafter(w[r], -INFINITY), hi[r] = std::nextafter(w[r], INFINITY);
            wboxes[i].grow(lo); wboxes[i].grow(hi);
        }
    }
    WideTree tlas = build_tree(wboxes, 1);
    std::vector<InstanceRec> recs(inst.size());
    std::vector<float4> iboxes(inst.size() * 2);
    std::vector<Instance> ordered(inst.size());
    for (size_t k = 0; k < tlas.order.size(); k++) {
        const Instance& in = inst[tlas.order[k]];
        ordered[k] = in;
        InstanceRec& r = recs[k];
        r.r0 = make_float4((float)in.inv[0], (float)in.inv[1], (float)in.inv[2], (float)in.inv[3]);
        r.r1 = make_float4((float)in.inv[4], (float)in.inv[5], (float)in.inv[6], (float)in.inv[7]);
        r.r2 = make_float4((float)in.inv[8], (float)in.inv[9], (float)in.inv[10], (float)in.inv[11]);
        r.nodes = meshes[(size_t)in.mesh].tree.nodes.data();
        r.tris = meshes[(size_t)in.mesh].tris.data();
        const Box& wb = wboxes[tlas.order[k]];
        iboxes[2 * k] = make_float4(wb.lo[0], wb.lo[1], wb.lo[2], 0);
        iboxes[2 * k + 1] = make_float4(wb.hi[0], wb.hi[1], wb.hi[2], 0);
    }
    TraceScene sc{tlas.nodes.data(), recs.data(), iboxes.data(), 33u};

    long hits = 0, rays = 0, agree = 0, clear_rays = 0, clear_agree = 0, hint_same = 0, cand_rays = 0, cand_same = 0, closest_ok = 0, closest_n = 0;
    uint2 stack[LUZ_STACK_SIZE];
    LocalStats st = {0, 0, 0};
    for (int r = 0; r < n_rays; r++) {
        const double e1 = extent + 1.0;
        double o[3] = {uni(-e1, e1), uni(-e1, e1), uni(-e1, e1)}, tgt[3] = {uni(-e1, e1), uni(-e1, e1), uni(-e1, e1)}, d[3];
        const bool shortray = r % 3 == 0;
        if (shortray) { // AO-like: start on an instance's surface region, short reach
            const Instance& in = ordered[(size_t)(rng() % ordered.size())];
            const double p[3] = {uni(-1, 1), uni(-1, 1), uni(-1, 1)};
            for (int k = 0; k < 3; k++) o[k] = in.m[k * 4] * p[0] + in.m[k * 4 + 1] * p[1] + in.m[k * 4 + 2] * p[2] + in.m[k * 4 + 3];
            for (int k = 0; k < 3; k++) tgt[k] = o[k] + uni(-1, 1);
        }
        for (int k = 0; k < 3; k++) d[k] = tgt[k] - o[k];
        const float3 fo = f3((float)o[0], (float)o[1], (float)o[2]), fd = f3((float)d[0], (float)d[1], (float)d[2]);
        const double od[3] = {fo.x, fo.y, fo.z}, dd[3] = {fd.x, fd.y, fd.z}; // the reference sees the rounded ray
        const float tmin = 1e-3f, tmax = shortray ? 1.0f : 0.95f;
        double tbest, margin;
        const bool ref = exhaustive(meshes, ordered, od, dd, tmin, tmax, &tbest, &margin);
        const bool got = trace_ray<false, true>(sc, fo, fd, tmin, tmax, nullptr, &st, stack);
        rays++;
        hits += ref;
        agree += got == ref;
        if (margin > 1e-4) { clear_rays++; clear_agree += got == ref; }
        // occluder hint: any instance tried first, then the root descent -> the same answer
        uint32_t hint = (uint32_t)(rng() % ordered.size());
        HitInfo hi;
        const bool hinted = trace_ray<false, false>(sc, fo, fd, tmin, tmax, &hi, &st, stack, &hint, 1, 1, true);
        hint_same += hinted == got;
        // closest hit
        HitInfo ch;
        const bool cgot = trace_ray<true, false>(sc, fo, fd, tmin, tmax, &ch, &st, stack);
        if (margin > 1e-4) {
            closest_n++;
            closest_ok += (cgot == ref) && (!ref || std::fabs(ch.t - tbest) <= 1e-4 * std::max(1.0, std::fabs(tbest)));
        }
        if (shortray) { // candidate list from the reach box of this single ray, filtered like the AO path does per pixel
            const float m = tmax * 1.001f;
            const float3 ext = f3(m * std::fabs(fd.x) + 1e-6f, m * std::fabs(fd.y) + 1e-6f, m * std::fabs(fd.z) + 1e-6f);
            uint32_t cand[16];
            uint2 cstack[LUZ_STACK_SIZE];
            int n = collect_instances<false>(sc, fo - ext, fo + ext, cand, 1, 16, cstack, &st);
            if (n >= 0) {
                // hemisphere frame whose axis is the ray itself: T = B = 0, C = d  ->  reach = the segment's box
                n = filter_candidates<false>(sc, fo, f3(0, 0, 0), f3(0, 0, 0), fd, m, cand, 1, n, &st);
                const bool cres = n == 0 ? false : trace_ray<false, false>(sc, fo, fd, tmin, tmax, nullptr, &st, stack, cand, 1, n);
                cand_rays++;
                cand_same += cres == got;
            }
        }
    }
    printf("{\"rays\":%ld,\"hits\":%ld,\"agree\":%ld,\"clear_rays\":%ld,\"clear_agree\":%ld,\"hint_same\":%ld,\"cand_rays\":%ld,\"cand_same\":%ld,"
           "\"closest_n\":%ld,\"closest_ok\":%ld,\"tlas_nodes\":%zu,\"tlas_levels\":%d,\"nodes_visited\":%u,\"tris_tested\":%u}\n",
           rays, hits, agree, clear_rays, clear_agree, hint_same, cand_rays, cand_same, closest_n, closest_ok, tlas.nodes.size(), tlas.levels,
           st.nodes, st.tris);
    return 0;
}

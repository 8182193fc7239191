// luz_oracle.cpp -- CPU oracle for the Luz lighting path.  TEST INFRASTRUCTURE ONLY (see
// luz_oracle.h).  Build: g++ -O2 -fopenmp -ffp-contract=off -fno-fast-math (oracle/Makefile).
//
// Every function names the reference lines (relative to /root/reference) it restates.  All
// arithmetic is IEEE fp32 with no contraction; GLSL built-ins map to: pow->powf, sqrt->sqrtf,
// sin/cos->sinf/cosf, normalize(v)->v/sqrtf(dot(v,v)), length->sqrtf(dot), fract(x)->x-floorf(x),
// mod(x,y)->x-y*floorf(x/y), mix(a,b,t)->a*(1-t)+b*t, min/max->fminf/fmaxf, mat*vec summed
// left to right over columns.
#include "luz_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

struct V3 {
    float x, y, z;
};
struct V4 {
    float x, y, z, w;
};
inline V3 v3(float x, float y, float z) { return V3{x, y, z}; }
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator*(float s, V3 a) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator/(V3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline V3 operator/(V3 a, V3 b) { return {a.x / b.x, a.y / b.y, a.z / b.z}; }
inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }
inline float length(V3 a) { return sqrtf(dot(a, a)); }
inline V3 normalize(V3 a) { return a / sqrtf(dot(a, a)); }
inline float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
inline float fractf(float x) { return x - floorf(x); }
inline float modf_glsl(float x, float y) { return x - y * floorf(x / y); }

inline V4 operator+(V4 a, V4 b) { return {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
inline V4 operator-(V4 a, V4 b) { return {a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w}; }
inline V4 operator*(V4 a, float s) { return {a.x * s, a.y * s, a.z * s, a.w * s}; }
inline V4 operator/(V4 a, float s) { return {a.x / s, a.y / s, a.z / s, a.w / s}; }
inline V4 min4(V4 a, V4 b) { return {fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z), fminf(a.w, b.w)}; }
inline V4 max4(V4 a, V4 b) { return {fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w)}; }
inline bool any_nan(V4 a) { return std::isnan(a.x) || std::isnan(a.y) || std::isnan(a.z) || std::isnan(a.w); }

// column-major mat4 (glm) times vec4
inline V4 mul(const float* m, V4 v) {
    V4 r;
    r.x = m[0] * v.x + m[4] * v.y + m[8] * v.z + m[12] * v.w;
    r.y = m[1] * v.x + m[5] * v.y + m[9] * v.z + m[13] * v.w;
    r.z = m[2] * v.x + m[6] * v.y + m[10] * v.z + m[14] * v.w;
    r.w = m[3] * v.x + m[7] * v.y + m[11] * v.z + m[15] * v.w;
    return r;
}

const float kPI = 3.14159265359f;            // LuzCommon.h:11
const float kGoldenRatio = 2.118033988749895f; // LuzCommon.h:12 (sic)

// utils.glsl:1-7
inline V3 depth_to_world(const luzw_scene_block* s, float u, float v, float depth) {
    V4 clip = {u * 2.0f - 1.0f, v * 2.0f - 1.0f, depth, 1.0f};
    V4 view = mul(s->inverse_proj, clip);
    view = view / view.w;
    V4 world = mul(s->inverse_view, view);
    return {world.x, world.y, world.z};
}

// ------------------------------------------------------------------------------------------
// Geometry: instanced triangle meshes, watertight ray/triangle test, exhaustive + BVH2 any-hit.
// Semantics of the ray query (light.frag:99-106, :125-132; VulkanWrapper.cpp:780, :1119-1120):
// opaque, two-sided, mask 0xFF, terminate on first hit, committed hit iff tmin < t < tmax with
// t parametric along the (possibly non-unit) direction.  The ray is taken into object space
// with the inverse of rows 0..2 of the instance matrix (VulkanWrapper.cpp:1122-1126), t is
// preserved.  Rays with a NaN origin/direction miss.
// ------------------------------------------------------------------------------------------

struct RayPre {
    V3 o, d;
    int kx, ky, kz;
    float sx, sy, sz;
};

inline RayPre make_ray(V3 o, V3 d) {
    RayPre r;
    r.o = o;
    r.d = d;
    float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
    int kz = (ax > ay) ? ((ax > az) ? 0 : 2) : ((ay > az) ? 1 : 2);
    int kx = kz + 1;
    if (kx == 3) kx = 0;
    int ky = kx + 1;
    if (ky == 3) ky = 0;
    const float dv[3] = {d.x, d.y, d.z};
    if (dv[kz] < 0.0f) std::swap(kx, ky);
    r.kx = kx;
    r.ky = ky;
    r.kz = kz;
    r.sx = dv[kx] / dv[kz];
    r.sy = dv[ky] / dv[kz];
    r.sz = 1.0f / dv[kz];
    return r;
}

// Watertight test after Woop, Benthin, Wald, "Watertight Ray/Triangle Intersection" (JCGT 2013).
inline bool tri_hit(const RayPre& r, const float* p0, const float* p1, const float* p2, float tmin, float tmax,
                    float* t_out, float* bu = nullptr, float* bv = nullptr) {
    const float o[3] = {r.o.x, r.o.y, r.o.z};
    const float A[3] = {p0[0] - o[0], p0[1] - o[1], p0[2] - o[2]};
    const float B[3] = {p1[0] - o[0], p1[1] - o[1], p1[2] - o[2]};
    const float C[3] = {p2[0] - o[0], p2[1] - o[1], p2[2] - o[2]};
    const float Ax = A[r.kx] - r.sx * A[r.kz], Ay = A[r.ky] - r.sy * A[r.kz];
    const float Bx = B[r.kx] - r.sx * B[r.kz], By = B[r.ky] - r.sy * B[r.kz];
    const float Cx = C[r.kx] - r.sx * C[r.kz], Cy = C[r.ky] - r.sy * C[r.kz];
    float U = Cx * By - Cy * Bx;
    float V = Ax * Cy - Ay * Cx;
    float W = Bx * Ay - By * Ax;
    if (U == 0.0f || V == 0.0f || W == 0.0f) {
        double CxBy = (double)Cx * (double)By, CyBx = (double)Cy * (double)Bx;
        U = (float)(CxBy - CyBx);
        double AxCy = (double)Ax * (double)Cy, AyCx = (double)Ay * (double)Cx;
        V = (float)(AxCy - AyCx);
        double BxAy = (double)Bx * (double)Ay, ByAx = (double)By * (double)Ax;
        W = (float)(BxAy - ByAx);
    }
    if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;
    const float det = U + V + W;
    if (det == 0.0f) return false;
    const float Az = r.sz * A[r.kz], Bz = r.sz * B[r.kz], Cz = r.sz * C[r.kz];
    const float T = U * Az + V * Bz + W * Cz;
    const float t = T / det;
    if (!(t > tmin && t < tmax)) return false;
    *t_out = t;
    if (bu) {
        *bu = U / det; // weight of p0
        *bv = V / det; // weight of p1
    }
    return true;
}

struct Aabb {
    float lo[3], hi[3];
    void reset() {
        for (int k = 0; k < 3; k++) {
            lo[k] = std::numeric_limits<float>::infinity();
            hi[k] = -std::numeric_limits<float>::infinity();
        }
    }
    void grow(const float* p) {
        for (int k = 0; k < 3; k++) {
            lo[k] = fminf(lo[k], p[k]);
            hi[k] = fmaxf(hi[k], p[k]);
        }
    }
    void grow(const Aabb& b) {
        for (int k = 0; k < 3; k++) {
            lo[k] = fminf(lo[k], b.lo[k]);
            hi[k] = fmaxf(hi[k], b.hi[k]);
        }
    }
};

struct Bvh2Node {
    Aabb box;
    uint32_t left;  // internal: index of left child (right = left+1); leaf: first primitive
    uint32_t count; // 0 = internal
};

struct Bvh2 {
    std::vector<Bvh2Node> nodes;
    std::vector<uint32_t> prims;

    void build(const std::vector<Aabb>& boxes, uint32_t leaf_size) {
        const uint32_t n = (uint32_t)boxes.size();
        prims.resize(n);
        for (uint32_t i = 0; i < n; i++) prims[i] = i;
        nodes.clear();
        nodes.reserve(2 * n + 1);
        nodes.push_back(Bvh2Node{});
        if (n == 0) {
            nodes[0].box.reset();
            nodes[0].left = 0;
            nodes[0].count = 0;
            return;
        }
        std::vector<float> cent(3 * (size_t)n);
        for (uint32_t i = 0; i < n; i++)
            for (int k = 0; k < 3; k++) cent[3 * (size_t)i + k] = 0.5f * (boxes[i].lo[k] + boxes[i].hi[k]);
        struct Item {
            uint32_t node, first, count;
        };
        std::vector<Item> todo;
        todo.push_back({0, 0, n});
        while (!todo.empty()) {
            Item it = todo.back();
            todo.pop_back();
            Aabb b, cb;
            b.reset();
            cb.reset();
            for (uint32_t i = it.first; i < it.first + it.count; i++) {
                b.grow(boxes[prims[i]]);
                cb.grow(&cent[3 * (size_t)prims[i]]);
            }
            nodes[it.node].box = b;
            if (it.count <= leaf_size) {
                nodes[it.node].left = it.first;
                nodes[it.node].count = it.count;
                continue;
            }
            int axis = 0;
            float ext = cb.hi[0] - cb.lo[0];
            for (int k = 1; k < 3; k++)
                if (cb.hi[k] - cb.lo[k] > ext) {
                    ext = cb.hi[k] - cb.lo[k];
                    axis = k;
                }
            uint32_t mid = it.count / 2;
            std::nth_element(prims.begin() + it.first, prims.begin() + it.first + mid,
                             prims.begin() + it.first + it.count, [&](uint32_t a, uint32_t c) {
                                 float ca = cent[3 * (size_t)a + axis], cc = cent[3 * (size_t)c + axis];
                                 return ca < cc || (ca == cc && a < c);
                             });
            uint32_t l = (uint32_t)nodes.size();
            nodes.push_back(Bvh2Node{});
            nodes.push_back(Bvh2Node{});
            nodes[it.node].left = l;
            nodes[it.node].count = 0;
            todo.push_back({l, it.first, mid});
            todo.push_back({l + 1, it.first + mid, it.count - mid});
        }
    }
};

inline bool box_hit(const Aabb& b, V3 o, V3 id, float tmin, float tmax) {
    float t0 = (b.lo[0] - o.x) * id.x, t1 = (b.hi[0] - o.x) * id.x;
    float lo = fminf(t0, t1), hi = fmaxf(t0, t1);
    t0 = (b.lo[1] - o.y) * id.y;
    t1 = (b.hi[1] - o.y) * id.y;
    lo = fmaxf(lo, fminf(t0, t1));
    hi = fminf(hi, fmaxf(t0, t1));
    t0 = (b.lo[2] - o.z) * id.z;
    t1 = (b.hi[2] - o.z) * id.z;
    lo = fmaxf(lo, fminf(t0, t1));
    hi = fminf(hi, fmaxf(t0, t1));
    lo = fmaxf(lo, tmin);
    hi = fminf(hi, tmax) * 1.0000004f;
    // NaN (0*inf) must not reject: use negated comparison
    return !(lo > hi);
}

struct MeshData {
    std::vector<float> pos;      // 3 per vertex
    std::vector<float> attr;     // normal3, tangent4, uv2 per vertex (9 floats) when stride == 48
    std::vector<uint32_t> idx;
    Bvh2 bvh;
    Aabb bounds;
    bool has_attr = false;
};

struct InstData {
    uint32_t mesh;
    float m[16];
    float inv[12]; // rows of the 3x4 inverse: inv[r*4 + c]
    uint32_t custom_index;
    Aabb world_box;
};

} // namespace

struct orc_world {
    std::vector<MeshData> meshes;
    std::vector<InstData> inst;
    Bvh2 tlas;
};

namespace {

// inverse of the affine 3x4 formed by rows 0..2 of the column-major mat4 (cofactor form)
void affine_inverse(const float* m, float* inv) {
    const float a00 = m[0], a01 = m[4], a02 = m[8], t0 = m[12];
    const float a10 = m[1], a11 = m[5], a12 = m[9], t1 = m[13];
    const float a20 = m[2], a21 = m[6], a22 = m[10], t2 = m[14];
    const float c00 = a11 * a22 - a12 * a21;
    const float c01 = a12 * a20 - a10 * a22;
    const float c02 = a10 * a21 - a11 * a20;
    const float det = a00 * c00 + a01 * c01 + a02 * c02;
    const float id = 1.0f / det;
    const float i00 = c00 * id, i01 = (a02 * a21 - a01 * a22) * id, i02 = (a01 * a12 - a02 * a11) * id;
    const float i10 = c01 * id, i11 = (a00 * a22 - a02 * a20) * id, i12 = (a02 * a10 - a00 * a12) * id;
    const float i20 = c02 * id, i21 = (a01 * a20 - a00 * a21) * id, i22 = (a00 * a11 - a01 * a10) * id;
    inv[0] = i00;
    inv[1] = i01;
    inv[2] = i02;
    inv[3] = -(i00 * t0 + i01 * t1 + i02 * t2);
    inv[4] = i10;
    inv[5] = i11;
    inv[6] = i12;
    inv[7] = -(i10 * t0 + i11 * t1 + i12 * t2);
    inv[8] = i20;
    inv[9] = i21;
    inv[10] = i22;
    inv[11] = -(i20 * t0 + i21 * t1 + i22 * t2);
}

inline V3 xform_point(const float* inv, V3 p) {
    return {inv[0] * p.x + inv[1] * p.y + inv[2] * p.z + inv[3], inv[4] * p.x + inv[5] * p.y + inv[6] * p.z + inv[7],
            inv[8] * p.x + inv[9] * p.y + inv[10] * p.z + inv[11]};
}
inline V3 xform_dir(const float* inv, V3 d) {
    return {inv[0] * d.x + inv[1] * d.y + inv[2] * d.z, inv[4] * d.x + inv[5] * d.y + inv[6] * d.z,
            inv[8] * d.x + inv[9] * d.y + inv[10] * d.z};
}

inline bool ray_is_nan(V3 o, V3 d) {
    return std::isnan(o.x) || std::isnan(o.y) || std::isnan(o.z) || std::isnan(d.x) || std::isnan(d.y) ||
           std::isnan(d.z);
}

struct Hit {
    float t;
    int32_t inst, prim;
    float bu, bv;
};

// Back-face culling of a rasterising pipeline (the Opaque Pipeline: VK_CULL_MODE_BACK_BIT, front = counter-clockwise in
// framebuffer space, DeferredRenderer.cpp "Opaque Pipeline" / VulkanWrapper.cpp:941-946), decided the way the
// rasteriser decides it: from the clip coordinates c_i = viewProj * model * p_i of the triangle's vertices.  The signed
// framebuffer area has the sign of -det[c_0; c_1; c_2] over (x, y, w); the triangle is front-facing iff the determinant
// is negative (see the shadow-map pass below, which culls the other side).
struct FaceCull {
    const float* view_proj; // column-major mat4, or nullptr: two-sided
    const float* model;     // the instance's model matrix
};
inline bool front_facing(const FaceCull& fc, const float* p0, const float* p1, const float* p2) {
    double c[3][3];
    const float* ps[3] = {p0, p1, p2};
    for (int k = 0; k < 3; k++) {
        const V4 wv = mul(fc.model, V4{ps[k][0], ps[k][1], ps[k][2], 1.0f});
        const V4 cv = mul(fc.view_proj, wv);
        c[k][0] = cv.x, c[k][1] = cv.y, c[k][2] = cv.w;
    }
    const double D = c[0][0] * (c[1][1] * c[2][2] - c[1][2] * c[2][1]) - c[0][1] * (c[1][0] * c[2][2] - c[1][2] * c[2][0]) +
                     c[0][2] * (c[1][0] * c[2][1] - c[1][1] * c[2][0]);
    return D < 0.0;
}

// mesh-level search in object space; closest == false returns at the first accepted hit
bool mesh_trace(const MeshData& md, V3 o, V3 d, float tmin, float& tmax, bool closest, int exhaustive, Hit& hit,
                const FaceCull* cull = nullptr) {
    if (d.x == 0.0f && d.y == 0.0f && d.z == 0.0f) return false;
    const RayPre r = make_ray(o, d);
    bool found = false;
    const uint32_t ntri = (uint32_t)md.idx.size() / 3;
    auto test = [&](uint32_t p) -> bool {
        const float* p0 = &md.pos[3 * (size_t)md.idx[3 * p + 0]];
        const float* p1 = &md.pos[3 * (size_t)md.idx[3 * p + 1]];
        const float* p2 = &md.pos[3 * (size_t)md.idx[3 * p + 2]];
        float t, bu, bv;
        if (tri_hit(r, p0, p1, p2, tmin, tmax, &t, &bu, &bv)) {
            if (cull && !front_facing(*cull, p0, p1, p2)) return false; // culled triangles produce no fragment
            tmax = t;
            hit.t = t;
            hit.prim = (int32_t)p;
            hit.bu = bu;
            hit.bv = bv;
            found = true;
            return true;
        }
        return false;
    };
    if (exhaustive) {
        for (uint32_t p = 0; p < ntri; p++)
            if (test(p) && !closest) return true;
        return found;
    }
    const V3 id = {1.0f / d.x, 1.0f / d.y, 1.0f / d.z};
    uint32_t stack[64];
    int sp = 0;
    stack[sp++] = 0;
    while (sp) {
        const Bvh2Node& n = md.bvh.nodes[stack[--sp]];
        if (!box_hit(n.box, o, id, tmin, tmax)) continue;
        if (n.count) {
            for (uint32_t i = 0; i < n.count; i++)
                if (test(md.bvh.prims[n.left + i]) && !closest) return true;
        } else {
            stack[sp++] = n.left;
            stack[sp++] = n.left + 1;
        }
    }
    return found;
}

// exhaustive: 0 = BVH2 traversal; 1 = every triangle of every instance (truth, O(rays * triangles)); 2 = every instance
// whose world box the ray segment meets (no hierarchy), every triangle inside it -- the check of the BVH2 topology for
// scenes of 10 M instanced triangles, where mode 1 is out of reach
bool world_trace(const orc_world* w, V3 o, V3 d, float tmin, float tmax, bool closest, int exhaustive, Hit& hit,
                 const float* cull_view_proj = nullptr) {
    hit.t = std::numeric_limits<float>::infinity();
    hit.inst = -1;
    hit.prim = -1;
    hit.bu = hit.bv = 0.0f;
    if (ray_is_nan(o, d) || std::isnan(tmin) || std::isnan(tmax)) return false;
    bool found = false;
    auto visit = [&](uint32_t ii) -> bool {
        const InstData& in = w->inst[ii];
        const V3 oo = xform_point(in.inv, o);
        const V3 od = xform_dir(in.inv, d);
        if (ray_is_nan(oo, od)) return false;
        Hit h = hit;
        const FaceCull fcull{cull_view_proj, in.m};
        if (mesh_trace(w->meshes[in.mesh], oo, od, tmin, tmax, closest, exhaustive, h, cull_view_proj ? &fcull : nullptr)) {
            hit = h;
            hit.inst = (int32_t)ii;
            found = true;
            return true;
        }
        return false;
    };
    const V3 id = {1.0f / d.x, 1.0f / d.y, 1.0f / d.z};
    if (exhaustive) {
        for (uint32_t ii = 0; ii < w->inst.size(); ii++) {
            if (exhaustive == 2 && !box_hit(w->inst[ii].world_box, o, id, tmin, tmax)) continue;
            if (visit(ii) && !closest) return true;
        }
        return found;
    }
    if (w->inst.empty()) return false;
    uint32_t stack[64];
    int sp = 0;
    stack[sp++] = 0;
    while (sp) {
        const Bvh2Node& n = w->tlas.nodes[stack[--sp]];
        if (!box_hit(n.box, o, id, tmin, tmax)) continue;
        if (n.count) {
            for (uint32_t i = 0; i < n.count; i++)
                if (visit(w->tlas.prims[n.left + i]) && !closest) return true;
        } else {
            stack[sp++] = n.left;
            stack[sp++] = n.left + 1;
        }
    }
    return found;
}

inline bool occluded(const orc_world* w, V3 o, V3 d, float tmin, float tmax, int exhaustive) {
    Hit h;
    return world_trace(w, o, d, tmin, tmax, false, exhaustive, h);
}

// ------------------------------------------------------------------------------------------
// Shadow maps: the two texture() look-ups of light.frag:147-165 and
// shadowMapVolumetricLight.comp:22-40.  D32F maps of res^2 texels: six cube-face layers
// (+X -X +Y -Y +Z -Z) holding |light.position - fragPos| / zFar for point lights, one layer of
// gl_FragCoord.z under the orthographic light.viewProj[0] otherwise; cleared to 1.0.  LINEAR /
// REPEAT sampler (VulkanWrapper.cpp:2429-2461): bilinear taps as nested lerps a + w * (b - a);
// the 2-D tap wraps; the cube tap picks the face by the Vulkan rules (major axis, z over y over
// x on ties, sc / tc table) and clamps its 2x2 footprint to the face (no blending with the
// adjacent face at cube edges: documented deviation).
// ------------------------------------------------------------------------------------------
const orc_shadow_map* g_shadow_maps = nullptr;
uint32_t g_n_shadow_maps = 0;

inline int wrap_mod(int i, int n) {
    int r = i % n;
    return r < 0 ? r + n : r;
}
inline float lerp1(float a, float b, float w) { return a + w * (b - a); }

float shadow_tap_2d(const float* d, int res, float u, float v) {
    const float x = u * (float)res - 0.5f, y = v * (float)res - 0.5f;
    const float fx0 = floorf(x), fy0 = floorf(y);
    const float fx = x - fx0, fy = y - fy0;
    if (std::isnan(fx0) || std::isnan(fy0)) return std::numeric_limits<float>::quiet_NaN();
    const int ix = (int)fminf(fmaxf(fx0, -1.0e9f), 1.0e9f), iy = (int)fminf(fmaxf(fy0, -1.0e9f), 1.0e9f);
    const int x0 = wrap_mod(ix, res), y0 = wrap_mod(iy, res), x1 = wrap_mod(ix + 1, res), y1 = wrap_mod(iy + 1, res);
    const float top = lerp1(d[(size_t)y0 * res + x0], d[(size_t)y0 * res + x1], fx);
    const float bot = lerp1(d[(size_t)y1 * res + x0], d[(size_t)y1 * res + x1], fx);
    return lerp1(top, bot, fy);
}

float shadow_tap_cube(const float* d, int res, V3 r) {
    const float ax = fabsf(r.x), ay = fabsf(r.y), az = fabsf(r.z);
    int face;
    float sc, tc, ma;
    if (az >= ax && az >= ay) {
        face = r.z < 0.0f ? 5 : 4;
        sc = r.z < 0.0f ? -r.x : r.x;
        tc = -r.y;
        ma = az;
    } else if (ay >= ax) {
        face = r.y < 0.0f ? 3 : 2;
        sc = r.x;
        tc = r.y < 0.0f ? -r.z : r.z;
        ma = ay;
    } else {
        face = r.x < 0.0f ? 1 : 0;
        sc = r.x < 0.0f ? r.z : -r.z;
        tc = -r.y;
        ma = ax;
    }
    const float u = 0.5f * (sc / ma) + 0.5f, v = 0.5f * (tc / ma) + 0.5f;
    const float x = u * (float)res - 0.5f, y = v * (float)res - 0.5f;
    const float fx0 = floorf(x), fy0 = floorf(y);
    const float fx = x - fx0, fy = y - fy0;
    if (std::isnan(fx0) || std::isnan(fy0)) return std::numeric_limits<float>::quiet_NaN();
    auto cl = [&](float f) { return std::min(std::max((int)fminf(fmaxf(f, -1.0e9f), 1.0e9f), 0), res - 1); };
    const int x0 = cl(fx0), y0 = cl(fy0), x1 = cl(fx0 + 1.0f), y1 = cl(fy0 + 1.0f);
    const float* f = d + (size_t)face * res * res;
    const float top = lerp1(f[(size_t)y0 * res + x0], f[(size_t)y0 * res + x1], fx);
    const float bot = lerp1(f[(size_t)y1 * res + x0], f[(size_t)y1 * res + x1], fx);
    return lerp1(top, bot, fy);
}

// the SHADOW_TYPE_MAP branch of EvaluateShadow (light.frag:147-165); 1 = in shadow
float shadow_map_factor(const luzw_light_block& light, const orc_shadow_map& m, V3 fragPos, V3 shadowOrigin) {
    const V3 lpos = v3(light.position[0], light.position[1], light.position[2]);
    if (light.type == LUZW_LIGHT_POINT) {
        const V3 lightToFrag = fragPos - lpos;
        const float shadowDepth = shadow_tap_cube(m.data, (int)m.res, lightToFrag);
        return (length(lightToFrag) - 0.05f >= shadowDepth * light.z_far) ? 1.0f : 0.0f;
    }
    const V4 fragInLight = mul(light.view_proj[0], V4{shadowOrigin.x, shadowOrigin.y, shadowOrigin.z, 1.0f});
    const float shadowDepth = shadow_tap_2d(m.data, (int)m.res, fragInLight.x * 0.5f + 0.5f, fragInLight.y * 0.5f + 0.5f);
    return (fragInLight.z >= shadowDepth) ? 1.0f : 0.0f;
}

// ------------------------------------------------------------------------------------------
// light.frag
// ------------------------------------------------------------------------------------------

// light.frag:17-26
float distribution_ggx(V3 N, V3 H, float roughness) {
    float a = roughness * roughness;
    float a2 = a * a;
    float NdotH = fmaxf(dot(N, H), 0.0f);
    float NdotH2 = NdotH * NdotH;
    float nom = a2;
    float denom = (NdotH2 * (a2 - 1.0f) + 1.0f);
    denom = kPI * denom * denom;
    return nom / denom;
}
// light.frag:28-36
float geometry_schlick_ggx(float NdotV, float roughness) {
    float r = roughness + 1.0f;
    float k = (r * r) / 8.0f;
    float nom = NdotV;
    float denom = NdotV * (1.0f - k) + k;
    return nom / denom;
}
// light.frag:38-45
float geometry_smith(V3 N, V3 V, V3 L, float roughness) {
    float NdotV = fmaxf(dot(N, V), 0.0f);
    float NdotL = fmaxf(dot(N, L), 0.0f);
    float ggx2 = geometry_schlick_ggx(NdotV, roughness);
    float ggx1 = geometry_schlick_ggx(NdotL, roughness);
    return ggx1 * ggx2;
}
// light.frag:47-49
V3 fresnel_schlick(float cosTheta, V3 F0) {
    float p = powf(clampf(1.0f - cosTheta, 0.0f, 1.0f), 5.0f);
    return F0 + (v3(1.0f, 1.0f, 1.0f) - F0) * p;
}

struct PixelCtx {
    const luzw_scene_block* scene;
    const orc_world* world;
    int exhaustive;
    float bn_r, bn_g; // blue-noise texel .rg for this pixel, already /255
    int frame_mod;    // frame % 128
    uint64_t rays, occl;
};

// light.frag:71-75: fract(texel + GOLDEN_RATIO*(128*i + frame%128)), .rg only
inline void blue_noise(const PixelCtx& c, int i, float& r0, float& r1) {
    const float k = (float)(128 * i + c.frame_mod);
    const float off = kGoldenRatio * k;
    r0 = fractf(c.bn_r + off);
    r1 = fractf(c.bn_g + off);
}

// light.frag:86-109.  mask_bits (may be null) receives bit (bit0 + i) per occluded sample.
float trace_shadow_ray(PixelCtx& c, V3 O, V3 L, float numSamples, float radius, uint32_t* mask, uint32_t bit0) {
    if (numSamples == 0.0f) return 0.0f;
    const V3 lightTangent = normalize(cross(L, v3(0.0f, 1.0f, 0.0f)));
    const V3 lightBitangent = normalize(cross(lightTangent, L));
    float numShadows = 0.0f;
    for (int i = 0; (float)i < numSamples; i++) {
        float r0, r1;
        blue_noise(c, i, r0, r1);
        // DiskSample light.frag:57-61
        const float pointRadius = radius * sqrtf(r0);
        const float pointAngle = r1 * 2.0f * kPI;
        const float dx = pointRadius * cosf(pointAngle), dy = pointRadius * sinf(pointAngle);
        const float tMax = length(L);
        const V3 direction = normalize(L + dx * lightTangent + dy * lightBitangent);
        c.rays++;
        if (occluded(c.world, O, direction, 0.001f, tMax, c.exhaustive)) {
            numShadows += 1.0f;
            c.occl++;
            if (mask) {
                uint32_t b = bit0 + (uint32_t)i;
                mask[b >> 5] |= 1u << (b & 31);
            }
        }
    }
    return numShadows / numSamples;
}

// light.frag:111-135
float trace_ao_rays(PixelCtx& c, V3 fragPos, V3 normal, uint32_t* mask) {
    const luzw_scene_block* s = c.scene;
    if (s->ao_num_samples == 0) return 1.0f;
    float ao = 0.0f;
    const V3 tangent = fabsf(normal.z) > 0.5f ? v3(0.0f, -normal.z, normal.y) : v3(-normal.y, normal.x, 0.0f);
    const V3 bitangent = cross(normal, tangent);
    const float tMin = s->ao_min, tMax = s->ao_max;
    for (int i = 0; i < s->ao_num_samples; i++) {
        float r0, r1;
        blue_noise(c, i, r0, r1);
        // HemisphereSample light.frag:63-69
        const float r = sqrtf(r0);
        const float theta = 6.283f * r1;
        const float hx = r * cosf(theta), hy = r * sinf(theta);
        const float hz = sqrtf(fmaxf(0.0f, 1.0f - r0));
        const V3 direction = tangent * hx + bitangent * hy + normal * hz;
        c.rays++;
        if (!occluded(c.world, fragPos, direction, tMin, tMax, c.exhaustive)) {
            ao += 1.0f;
        } else {
            c.occl++;
            if (mask) mask[i >> 5] |= 1u << (i & 31);
        }
    }
    return ao / (float)s->ao_num_samples;
}

inline const luzw_light_block& light_at(const luzw_scene_block* s, const luzw_light_block* extra, int i) {
    return i < LUZW_MAX_LIGHTS ? s->lights[i] : extra[i - LUZW_MAX_LIGHTS];
}

} // namespace

extern "C" {

static int g_threads = 0;
void orc_set_threads(int n) { g_threads = n; }
int orc_get_threads(void) {
#ifdef _OPENMP
    return g_threads > 0 ? g_threads : omp_get_max_threads();
#else
    return 1;
#endif
}
#ifdef _OPENMP
#define ORC_NT num_threads(orc_get_threads())
#else
#define ORC_NT
#endif

orc_world* orc_world_create(const orc_mesh* meshes, uint32_t n_meshes, const orc_instance* instances,
                            uint32_t n_instances) {
    orc_world* w = new orc_world();
    w->meshes.resize(n_meshes);
    for (uint32_t m = 0; m < n_meshes; m++) {
        MeshData& md = w->meshes[m];
        const orc_mesh& src = meshes[m];
        md.pos.resize(3 * (size_t)src.vertex_count);
        md.has_attr = src.vertex_stride == 48;
        if (md.has_attr) md.attr.resize(9 * (size_t)src.vertex_count);
        for (uint32_t v = 0; v < src.vertex_count; v++) {
            const float* p = (const float*)((const char*)src.vertices + (size_t)v * src.vertex_stride);
            for (int k = 0; k < 3; k++) md.pos[3 * (size_t)v + k] = p[k];
            if (md.has_attr)
                for (int k = 0; k < 9; k++) md.attr[9 * (size_t)v + k] = p[3 + k];
        }
        const uint32_t ntri = src.index_count / 3; // VulkanWrapper.cpp:761
        md.idx.assign(src.indices, src.indices + 3 * (size_t)ntri);
        std::vector<Aabb> boxes(ntri);
        md.bounds.reset();
        for (uint32_t t = 0; t < ntri; t++) {
            boxes[t].reset();
            for (int k = 0; k < 3; k++) boxes[t].grow(&md.pos[3 * (size_t)md.idx[3 * t + k]]);
            md.bounds.grow(boxes[t]);
        }
        md.bvh.build(boxes, 4);
    }
    w->inst.resize(n_instances);
    std::vector<Aabb> ib(n_instances);
    for (uint32_t i = 0; i < n_instances; i++) {
        InstData& in = w->inst[i];
        in.mesh = instances[i].mesh;
        memcpy(in.m, instances[i].model_mat, sizeof(in.m));
        in.custom_index = instances[i].custom_index;
        affine_inverse(in.m, in.inv);
        const Aabb& b = w->meshes[in.mesh].bounds;
        in.world_box.reset();
        for (int c = 0; c < 8; c++) {
            V4 p = {(c & 1) ? b.hi[0] : b.lo[0], (c & 2) ? b.hi[1] : b.lo[1], (c & 4) ? b.hi[2] : b.lo[2], 1.0f};
            V4 q = mul(in.m, p);
            const float qq[3] = {q.x, q.y, q.z};
            in.world_box.grow(qq);
        }
        // pad slightly: the object-space test is exact, the world box is only a culling aid
        for (int k = 0; k < 3; k++) {
            float e = 1e-5f * fmaxf(fabsf(in.world_box.lo[k]), fabsf(in.world_box.hi[k])) + 1e-7f;
            in.world_box.lo[k] -= e;
            in.world_box.hi[k] += e;
        }
        ib[i] = in.world_box;
    }
    w->tlas.build(ib, 1);
    return w;
}

void orc_world_destroy(orc_world* w) { delete w; }

void orc_trace_any(const orc_world* w, uint32_t n, const float* o, const float* d, const float* tmin,
                   const float* tmax, int exhaustive, uint8_t* hit) {
#pragma omp parallel for schedule(dynamic, 256) ORC_NT
    for (int64_t i = 0; i < (int64_t)n; i++)
        hit[i] = occluded(w, v3(o[3 * i], o[3 * i + 1], o[3 * i + 2]), v3(d[3 * i], d[3 * i + 1], d[3 * i + 2]),
                          tmin[i], tmax[i], exhaustive)
                     ? 1
                     : 0;
}

void orc_trace_closest(const orc_world* w, uint32_t n, const float* o, const float* d, const float* tmin,
                       const float* tmax, int exhaustive, float* t, int32_t* inst, int32_t* prim) {
#pragma omp parallel for schedule(dynamic, 256) ORC_NT
    for (int64_t i = 0; i < (int64_t)n; i++) {
        Hit h;
        world_trace(w, v3(o[3 * i], o[3 * i + 1], o[3 * i + 2]), v3(d[3 * i], d[3 * i + 1], d[3 * i + 2]), tmin[i],
                    tmax[i], true, exhaustive, h);
        t[i] = h.t;
        inst[i] = h.inst;
        prim[i] = h.prim;
    }
}

int orc_tri_test(const float* v0, const float* v1, const float* v2, const float* org, const float* dir, float tmin,
                 float tmax, float* t_out) {
    RayPre r = make_ray(v3(org[0], org[1], org[2]), v3(dir[0], dir[1], dir[2]));
    float t = 0.0f;
    int h = tri_hit(r, v0, v1, v2, tmin, tmax, &t) ? 1 : 0;
    if (t_out) *t_out = t;
    return h;
}

void orc_blue_noise_sample(const uint8_t* bn, uint32_t bn_w, uint32_t bn_h, uint32_t px, uint32_t py, int i,
                           uint32_t frame, float out2[2]) {
    PixelCtx c{};
    const int bx = (int)modf_glsl((float)px + 0.5f, (float)bn_w);
    const int by = (int)modf_glsl((float)py + 0.5f, (float)bn_h);
    const uint8_t* t = bn + 4 * ((size_t)by * bn_w + bx);
    c.bn_r = (float)t[0] / 255.0f;
    c.bn_g = (float)t[1] / 255.0f;
    c.frame_mod = (int)(frame % 128u);
    blue_noise(c, i, out2[0], out2[1]);
}

void orc_depth_to_world(const luzw_scene_block* scene, float u, float v, float depth, float out3[3]) {
    V3 p = depth_to_world(scene, u, v, depth);
    out3[0] = p.x;
    out3[1] = p.y;
    out3[2] = p.z;
}

// light.frag:171-235
int orc_light_pass(const luzw_scene_block* scene, const luzw_light_block* extra_lights, uint32_t n_extra,
                   uint32_t width, uint32_t height, const orc_gbuffer* gb, uint32_t frame,
                   const uint8_t* blue_noise_rgba8, uint32_t bn_w, uint32_t bn_h, const orc_world* world,
                   int exhaustive, uint32_t y0, uint32_t y1, float* out, uint32_t* shadow_mask,
                   uint32_t shadow_words, uint32_t* ao_mask, uint32_t ao_words, orc_stats* stats) {
    if (y1 < y0) return -2;
    std::vector<uint32_t> rows(y1 - y0);
    for (uint32_t y = y0; y < y1; y++) rows[y - y0] = y;
    return orc_light_pass_rows(scene, extra_lights, n_extra, width, height, gb, frame, blue_noise_rgba8, bn_w, bn_h, world,
                               exhaustive, rows.data(), (uint32_t)rows.size(), 0u, width, out, shadow_mask, shadow_words,
                               ao_mask, ao_words, stats);
}

// The same over an arbitrary list of rows (benchmark-scale parity checks sample rows of a 4K / 8K frame).
int orc_light_pass_rows(const luzw_scene_block* scene, const luzw_light_block* extra_lights, uint32_t n_extra,
                        uint32_t width, uint32_t height, const orc_gbuffer* gb, uint32_t frame,
                        const uint8_t* blue_noise_rgba8, uint32_t bn_w, uint32_t bn_h, const orc_world* world,
                        int exhaustive, const uint32_t* row_list, uint32_t n_rows, uint32_t x0, uint32_t x1, float* out,
                        uint32_t* shadow_mask, uint32_t shadow_words, uint32_t* ao_mask, uint32_t ao_words,
                        orc_stats* stats) {
    if (x1 > width || x0 > x1) return -2;
    for (uint32_t k = 0; k < n_rows; k++)
        if (row_list[k] >= height) return -2;
    const int numLights = scene->num_lights + (int)n_extra;
    if (scene->shadow_type == LUZW_SHADOW_MAP) { // the maps come from orc_bind_shadow_maps
        for (int i = 0; i < numLights; i++)
            if (light_at(scene, extra_lights, i).shadow_map != -1 &&
                ((uint32_t)i >= g_n_shadow_maps || !g_shadow_maps[i].data))
                return -1;
    }
    uint64_t tot_rays = 0, tot_occl = 0, tot_lit = 0;
    int frame_signed = (int)frame;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : tot_rays, tot_occl, tot_lit) ORC_NT
    for (int64_t yy = 0; yy < (int64_t)n_rows; yy++) {
        const uint32_t y = row_list[yy];
        for (uint32_t x = x0; x < x1; x++) {
            const size_t pix = (size_t)y * width + x;
            float* o = out + 4 * pix;
            uint32_t* smask = shadow_mask ? shadow_mask + pix * shadow_words : nullptr;
            uint32_t* amask = ao_mask ? ao_mask + pix * ao_words : nullptr;
            if (smask) memset(smask, 0, sizeof(uint32_t) * shadow_words);
            if (amask) memset(amask, 0, sizeof(uint32_t) * ao_words);
            // G-buffer fetches (light.frag:172-176): texel loads, see SURVEY section 9 item 13
            const uint8_t* a8 = gb->albedo + 4 * pix;
            const float* n4 = gb->normal + 4 * pix;
            const uint8_t* m8 = gb->material + 4 * pix;
            const uint8_t* e8 = gb->emission + 4 * pix;
            const float depth = gb->depth[pix];
            const V3 albedo = {powf((float)a8[0] / 255.0f, 2.2f), powf((float)a8[1] / 255.0f, 2.2f),
                               powf((float)a8[2] / 255.0f, 2.2f)};
            const V3 N = {n4[0], n4[1], n4[2]};
            const V3 ambientLight = v3(scene->ambient_light_color[0], scene->ambient_light_color[1],
                                       scene->ambient_light_color[2]) *
                                    scene->ambient_light_intensity;
            if (length(N) == 0.0f) { // :178-181
                o[0] = ambientLight.x;
                o[1] = ambientLight.y;
                o[2] = ambientLight.z;
                o[3] = 1.0f;
                continue;
            }
            tot_lit++;
            const float roughness = (float)m8[0] / 255.0f;
            const float metallic = (float)m8[1] / 255.0f;
            const float occlusion = (float)m8[2] / 255.0f;
            const V3 emission = {(float)e8[0] / 255.0f, (float)e8[1] / 255.0f, (float)e8[2] / 255.0f};
            const float u = ((float)x + 0.5f) / (float)width, v = ((float)y + 0.5f) / (float)height;
            const V3 fragPos = depth_to_world(scene, u, v, depth);
            const V3 camPos = v3(scene->cam_pos[0], scene->cam_pos[1], scene->cam_pos[2]);
            const V3 V = normalize(camPos - fragPos);
            V3 F0 = v3(0.04f, 0.04f, 0.04f);
            F0 = F0 * (1.0f - metallic) + albedo * metallic; // mix
            V3 Lo = v3(0.0f, 0.0f, 0.0f);

            PixelCtx c{};
            c.scene = scene;
            c.world = world;
            c.exhaustive = exhaustive;
            c.frame_mod = frame_signed % 128;
            {
                const int bx = (int)modf_glsl((float)x + 0.5f, (float)bn_w);
                const int by = (int)modf_glsl((float)y + 0.5f, (float)bn_h);
                const uint8_t* t = blue_noise_rgba8 + 4 * ((size_t)by * bn_w + bx);
                c.bn_r = (float)t[0] / 255.0f;
                c.bn_g = (float)t[1] / 255.0f;
            }

            uint32_t shadow_bit = 0;
            for (int i = 0; i < numLights; i++) {
                const luzw_light_block& light = light_at(scene, extra_lights, i);
                const V3 lpos = v3(light.position[0], light.position[1], light.position[2]);
                const V3 ldir = v3(light.direction[0], light.direction[1], light.direction[2]);
                const V3 L_ = lpos - fragPos;
                V3 L = normalize(L_);
                float attenuation = 1.0f;
                if (light.type == LUZW_LIGHT_DIRECTIONAL) {
                    L = normalize(-ldir);
                } else if (light.type == LUZW_LIGHT_SPOT) {
                    const float dist = length(lpos - fragPos);
                    attenuation = 1.0f / (dist * dist);
                    const float theta = dot(L, normalize(-ldir));
                    const float epsilon = light.inner_angle - light.outer_angle;
                    attenuation *= clampf((theta - light.outer_angle) / epsilon, 0.0f, 1.0f);
                } else if (light.type == LUZW_LIGHT_POINT) {
                    const float dist = length(lpos - fragPos);
                    attenuation = 1.0f / (dist * dist);
                }
                // EvaluateShadow light.frag:137-169
                float shadowFactor;
                {
                    const float shadowBias = fmaxf(length(fragPos - camPos) * 0.01f, 0.05f);
                    const V3 shadowOrigin = fragPos + N * shadowBias;
                    const float dist = length(lpos - fragPos);
                    if (scene->shadow_type == LUZW_SHADOW_RAYTRACING) {
                        const V3 Lr = (light.type == LUZW_LIGHT_DIRECTIONAL) ? ldir * dot(ldir, L) * dist : L * dist;
                        shadowFactor = trace_shadow_ray(c, shadowOrigin, Lr, (float)light.num_shadow_samples,
                                                        light.radius, smask, shadow_bit);
                        shadow_bit += (uint32_t)(light.num_shadow_samples > 0 ? light.num_shadow_samples : 0);
                    } else if (scene->shadow_type == LUZW_SHADOW_MAP && light.shadow_map != -1) {
                        shadowFactor = shadow_map_factor(light, g_shadow_maps[i], fragPos, shadowOrigin); // :147-165
                    } else {
                        shadowFactor = 1.0f; // :166-168
                    }
                }
                const V3 lcol = v3(light.color[0], light.color[1], light.color[2]);
                const V3 radiance = lcol * light.intensity * attenuation * (1.0f - shadowFactor);

                const V3 H = normalize(V + L);
                const float NDF = distribution_ggx(N, H, roughness);
                const float G = geometry_smith(N, V, L, roughness);
                const V3 F = fresnel_schlick(clampf(dot(H, V), 0.0f, 1.0f), F0);
                const V3 num = NDF * G * F;
                const float denom = 4.0f * fmaxf(dot(N, V), 0.0f) * fmaxf(dot(N, L), 0.0f) + 0.0001f;
                const V3 spec = num / denom;
                const V3 kS = F;
                V3 kD = v3(1.0f, 1.0f, 1.0f) - kS;
                kD = kD * (1.0f - metallic);
                const float NdotL = fmaxf(dot(N, L), 0.0f);
                Lo = Lo + (kD * albedo / kPI + spec) * radiance * NdotL;
            }
            const float aoBias = length(fragPos - camPos) * 0.01f;
            const V3 aoOrigin = fragPos + N * aoBias;
            const float rayTracedAo = trace_ao_rays(c, aoOrigin, N, amask);
            const V3 ambient = ambientLight * albedo * occlusion * rayTracedAo;
            const V3 color = ambient + Lo + emission;
            o[0] = color.x;
            o[1] = color.y;
            o[2] = color.z;
            o[3] = 1.0f;
            tot_rays += c.rays;
            tot_occl += c.occl;
        }
    }
    if (stats) {
        stats->lit_pixels = tot_lit;
        stats->rays = tot_rays;
        stats->rays_occluded = tot_occl;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------
// taa.comp.  All texture() taps go through the one LINEAR/REPEAT sampler
// (VulkanWrapper.cpp:2429-2461): taps at texel centres are texel loads with wrap; the history
// tap is an fp32 bilinear fetch with wrap.
// ------------------------------------------------------------------------------------------
namespace {
struct Img4 {
    const float* p;
    int w, h;
};
inline int wrapi(int i, int n) {
    int r = i % n;
    return r < 0 ? r + n : r;
}
inline V4 texel4(const Img4& im, int x, int y) {
    const float* q = im.p + 4 * ((size_t)wrapi(y, im.h) * im.w + wrapi(x, im.w));
    return {q[0], q[1], q[2], q[3]};
}
// nearest texel of a uv that sits on a texel centre
inline V4 tap4(const Img4& im, float u, float v) {
    return texel4(im, (int)floorf(u * (float)im.w), (int)floorf(v * (float)im.h));
}
inline float tap1(const float* d, int w, int h, float u, float v) {
    return d[(size_t)wrapi((int)floorf(v * (float)h), h) * w + wrapi((int)floorf(u * (float)w), w)];
}
inline V4 bilinear4(const Img4& im, float u, float v) {
    const float x = u * (float)im.w - 0.5f, y = v * (float)im.h - 0.5f;
    const float fx0 = floorf(x), fy0 = floorf(y);
    const float fx = x - fx0, fy = y - fy0;
    const int x0 = (int)fx0, y0 = (int)fy0;
    const V4 t00 = texel4(im, x0, y0), t10 = texel4(im, x0 + 1, y0);
    const V4 t01 = texel4(im, x0, y0 + 1), t11 = texel4(im, x0 + 1, y0 + 1);
    const V4 top = t00 * (1.0f - fx) + t10 * fx;
    const V4 bot = t01 * (1.0f - fx) + t11 * fx;
    return top * (1.0f - fy) + bot * fy;
}
// utils.glsl:9-15
inline float mitchell(float x) {
    const float B = 1.0f / 3.0f, C = 1.0f / 3.0f;
    const float x2 = x * x, x3 = x2 * x;
    return (6.0f - 2.0f * B) * x3 - (6.0f - 2.0f * B - 3.0f * C) * x2 + 1.0f;
}
inline float luminance(V3 c) { return dot(c, v3(0.2127f, 0.7152f, 0.0722f)); } // utils.glsl:95-97
} // namespace

float orc_mitchell(float x) { return mitchell(x); }

int orc_taa_pass(const luzw_scene_block* scene, uint32_t width, uint32_t height, const float* light_in,
                 const float* history, const float* depth, int reconstruct, uint32_t y0, uint32_t y1, float* out) {
    const Img4 light{light_in, (int)width, (int)height};
    const Img4 hist{history, (int)width, (int)height};
    const float sw = (float)width, sh = (float)height;
    const float w_corner = mitchell(sqrtf(2.0f)), w_edge = mitchell(1.0f), w_centre = mitchell(0.0f);
#pragma omp parallel for schedule(dynamic, 4) ORC_NT
    for (int64_t yy = (int64_t)y0; yy < (int64_t)y1; yy++) {
        for (uint32_t x = 0; x < width; x++) {
            const uint32_t y = (uint32_t)yy;
            float* o = out + 4 * ((size_t)y * width + x);
            const float su = ((float)x + 0.5f) / sw, sv = ((float)y + 0.5f) / sh; // get_uv :15-17
            // find_closest_3x3 :88-119
            const float ddx = fabsf(1.0f / sw), ddy = fabsf(1.0f / sh);
            float dminx = -1.0f, dminy = -1.0f, dminz = tap1(depth, width, height, su - ddx, sv - ddy);
            for (int j = -1; j <= 1; j++)
                for (int i = -1; i <= 1; i++) {
                    if (i == -1 && j == -1) continue;
                    const float uu = (i < 0) ? su - ddx : (i > 0 ? su + ddx : su);
                    const float vv = (j < 0) ? sv - ddy : (j > 0 ? sv + ddy : sv);
                    const float z = tap1(depth, width, height, uu, vv);
                    if (dminz > z) {
                        dminx = (float)i;
                        dminy = (float)j;
                        dminz = z;
                    }
                }
            const float cu = su + ddx * dminx, cv = sv + ddy * dminy;
            // get_motion_vector :19-27
            float mvx, mvy;
            {
                const float d = tap1(depth, width, height, cu, cv);
                const V3 wp = depth_to_world(scene, cu, cv, d);
                V4 prevNDC = mul(scene->prev_view_proj, V4{wp.x, wp.y, wp.z, 1.0f});
                V4 curNDC = mul(scene->view_proj, V4{wp.x, wp.y, wp.z, 1.0f});
                prevNDC.x /= prevNDC.w;
                prevNDC.y /= prevNDC.w;
                curNDC.x /= curNDC.w;
                curNDC.y /= curNDC.w;
                mvx = ((curNDC.x - scene->jitter[0]) - (prevNDC.x - scene->prev_jitter[0])) * 0.5f;
                mvy = ((curNDC.y - scene->jitter[1]) - (prevNDC.y - scene->prev_jitter[1])) * 0.5f;
            }
            const float hu = su - mvx, hv = sv - mvy;
            V4 historySample = bilinear4(hist, hu, hv);
            // get_neighbor_3x3 :29-86
            const float du = 1.0f / sw, dv = 1.0f / sh;
            const V4 ctl = tap4(light, su - du, sv - dv), ctc = tap4(light, su, sv - dv),
                     ctr = tap4(light, su + du, sv - dv);
            const V4 cml = tap4(light, su - du, sv), cmc = tap4(light, su, sv), cmr = tap4(light, su + du, sv);
            const V4 cbl = tap4(light, su - du, sv + dv), cbc = tap4(light, su, sv + dv),
                     cbr = tap4(light, su + du, sv + dv);
            V4 cmin = min4(ctl, min4(ctc, min4(ctr, min4(cml, min4(cmc, min4(cmr, min4(cbl, min4(cbc, cbr))))))));
            V4 cmax = max4(ctl, max4(ctc, max4(ctr, max4(cml, max4(cmc, max4(cmr, max4(cbl, max4(cbc, cbr))))))));
            V4 cavg = (ctl + ctc + ctr + cml + cmc + cmr + cbl + cbc + cbr) / 9.0f;
            const V4 cmin5 = min4(ctc, min4(cml, min4(cmc, min4(cmr, cbc))));
            const V4 cmax5 = max4(ctc, max4(cml, max4(cmc, max4(cmr, cbc))));
            const V4 cavg5 = (ctc + cml + cmc + cmr + cbc) / 5.0f;
            cmin = (cmin + cmin5) * 0.5f;
            cmax = (cmax + cmax5) * 0.5f;
            cavg = (cavg + cavg5) * 0.5f;
            V4 sourceSample = {0, 0, 0, 0};
            if (reconstruct == 1) {
                float weightSum = 0.0f;
                const V4* taps[9] = {&ctl, &ctc, &ctr, &cml, &cmc, &cmr, &cbl, &cbc, &cbr};
                const float wts[9] = {w_corner, w_edge, w_corner, w_edge, w_centre, w_edge, w_corner, w_edge, w_corner};
                for (int k = 0; k < 9; k++) {
                    sourceSample = sourceSample + (*taps[k]) * wts[k];
                    weightSum += wts[k];
                }
                sourceSample = sourceSample / weightSum;
            }
            if (reconstruct == 0 || any_nan(sourceSample)) sourceSample = cmc; // :287-289
            if (hu > 1.0f || hv > 1.0f || hu < 0.0f || hv < 0.0f) {          // :291-294
                o[0] = sourceSample.x;
                o[1] = sourceSample.y;
                o[2] = sourceSample.z;
                o[3] = sourceSample.w;
                continue;
            }
            // clip_aabb :121-143 with p = clamp(cavg, cmin, cmax), q = history
            {
                const V4 p = {clampf(cavg.x, cmin.x, cmax.x), clampf(cavg.y, cmin.y, cmax.y),
                              clampf(cavg.z, cmin.z, cmax.z), clampf(cavg.w, cmin.w, cmax.w)};
                V4 r = historySample - p;
                const V3 rmax = {cmax.x - p.x, cmax.y - p.y, cmax.z - p.z};
                const V3 rmin = {cmin.x - p.x, cmin.y - p.y, cmin.z - p.z};
                const float eps = 0.00000001f;
                if (r.x > rmax.x + eps) r = r * (rmax.x / r.x);
                if (r.y > rmax.y + eps) r = r * (rmax.y / r.y);
                if (r.z > rmax.z + eps) r = r * (rmax.z / r.z);
                if (r.x < rmin.x - eps) r = r * (rmin.x / r.x);
                if (r.y < rmin.y - eps) r = r * (rmin.y / r.y);
                if (r.z < rmin.z - eps) r = r * (rmin.z / r.z);
                historySample = p + r;
            }
            float sourceWeight = 0.05f;
            float historyWeight = 1.0f - sourceWeight;
            { // anti_flicker :145-152
                const V3 s3 = {sourceSample.x, sourceSample.y, sourceSample.z};
                const V3 h3 = {historySample.x, historySample.y, historySample.z};
                const V3 cs = s3 * (1.0f / (fmaxf(fmaxf(s3.x, s3.y), s3.z) + 1.0f));
                const V3 ch = h3 * (1.0f / (fmaxf(fmaxf(h3.x, h3.y), h3.z) + 1.0f));
                sourceWeight *= 1.0f / (1.0f + luminance(cs));
                historyWeight *= 1.0f / (1.0f + luminance(ch));
            }
            V4 result = (sourceSample * sourceWeight + historySample * historyWeight) /
                        fmaxf(sourceWeight + historyWeight, 0.0000001f);
            if (any_nan(result)) result = sourceSample;
            o[0] = result.x;
            o[1] = result.y;
            o[2] = result.z;
            o[3] = result.w;
        }
    }
    return 0;
}

// ------------------------------------------------------------------------------------------
// screenSpaceVolumetricLight.comp:22-62, dispatched between the light
// pass and TAA (main.cpp:274-279, DeferredRenderer.cpp:294-307) when any light has a volumetric
// type.  Per pixel and per light with volumetricType == VOLUMETRIC_TYPE_SCREEN_SPACE the shader
// marches `volumetricSamples` steps from the pixel towards the light's screen position and adds
// light for every step that lands on background (depth == 1).  pixelUV = pixelPos / imageSize
// (texel CORNER, :27); the depth fetches are texture() taps through the LINEAR / REPEAT sampler
// at arbitrary uv, i.e. genuine bilinear fetches.  They are restated as nested lerps
// a + w * (b - a), which return a constant neighbourhood exactly like the fixed-point weights
// of a texture unit do, so that `== 1.0` (:50) means "all four texels are background".
// ------------------------------------------------------------------------------------------
namespace {
inline float bilinear1(const float* d, int w, int h, float u, float v) {
    const float x = u * (float)w - 0.5f, y = v * (float)h - 0.5f;
    const float fx0 = floorf(x), fy0 = floorf(y);
    const float fx = x - fx0, fy = y - fy0;
    const int x0 = wrapi((int)fx0, w), y0 = wrapi((int)fy0, h);
    const int x1 = wrapi((int)fx0 + 1, w), y1 = wrapi((int)fy0 + 1, h);
    const float t00 = d[(size_t)y0 * w + x0], t10 = d[(size_t)y0 * w + x1];
    const float t01 = d[(size_t)y1 * w + x0], t11 = d[(size_t)y1 * w + x1];
    const float top = t00 + fx * (t10 - t00);
    const float bot = t01 + fx * (t11 - t01);
    return top + fy * (bot - top);
}
} // namespace

int orc_volumetric_screen_pass(const luzw_scene_block* scene, const luzw_light_block* extra_lights, uint32_t n_extra,
                               uint32_t width, uint32_t height, const float* depth, const uint8_t* bn, uint32_t bn_w,
                               uint32_t bn_h, uint32_t frame, uint32_t y0, uint32_t y1, float* light_inout) {
    const int n_lights = scene->num_lights + (int)n_extra;
    const int frame_mod = (int)((int32_t)frame % 128);
    const int W = (int)width, H = (int)height;
#pragma omp parallel for schedule(dynamic, 4) num_threads(orc_get_threads())
    for (int y = (int)y0; y < (int)y1; y++) {
        for (int x = 0; x < W; x++) {
            const float pu = (float)x / (float)W, pv = (float)y / (float)H; // :27
            const uint8_t* t = bn + 4 * ((size_t)(y % (int)bn_h) * bn_w + (size_t)(x % (int)bn_w));
            const float bn_r = (float)t[0] / 255.0f; // :16-20, .r only
            V3 radiance = v3(0.0f, 0.0f, 0.0f);
            for (int li = 0; li < n_lights; li++) {
                const luzw_light_block& light = light_at(scene, extra_lights, li);
                if (light.volumetric_type != 1) continue; // VOLUMETRIC_TYPE_SCREEN_SPACE, :32
                const V3 lpos = v3(light.position[0], light.position[1], light.position[2]);
                V4 lp = mul(scene->view_proj, V4{lpos.x, lpos.y, lpos.z, 1.0f});
                if (light.type == 2) // LUZ_LIGHT_TYPE_DIRECTIONAL, :37-39
                    lp = mul(scene->view_proj, V4{-light.direction[0] * 10000.0f, -light.direction[1] * 10000.0f,
                                                  -light.direction[2] * 10000.0f, 1.0f});
                const float lu = (lp.x / lp.w) * 0.5f + 0.5f, lv = (lp.y / lp.w) * 0.5f + 0.5f; // :41
                const int samples = light.volumetric_samples;
                const float absorption = light.volumetric_absorption / 1000.0f;
                const float inv_n = 1.0f / (float)samples;
                const float du = (pu - lu) * inv_n, dv = (pv - lv) * inv_n; // :45
                const float j0 = (fractf(bn_r + kGoldenRatio * (float)(128 * 0 + frame_mod)) * 2.0f - 1.0f) * 0.003f;
                float su = pu + j0, sv = pv + j0; // :46
                for (int i = 0; i < samples; i++) {
                    const float ji = (fractf(bn_r + kGoldenRatio * (float)(128 * (i + 1) + frame_mod)) * 2.0f - 1.0f) * 0.003f;
                    su -= du + ji; // :48
                    sv -= dv + ji;
                    const bool inside = su >= 0.0f && su <= 1.0f && sv >= 0.0f && sv <= 1.0f;
                    if (!inside) continue;
                    const float sd = bilinear1(depth, W, H, su, sv);
                    if (sd != 1.0f) continue; // :50
                    V3 sr = v3(light.color[0], light.color[1], light.color[2]) * light.intensity * absorption;
                    if (light.type == 0) { // LUZ_LIGHT_TYPE_POINT, :52-55
                        const V3 wp = depth_to_world(scene, su, sv, sd);
                        sr = sr * (5.0f / length(wp - lpos));
                    }
                    radiance = radiance + sr;
                }
            }
            float* px = light_inout + 4 * ((size_t)y * W + x); // :59-61
            px[0] += radiance.x;
            px[1] += radiance.y;
            px[2] += radiance.z;
            px[3] += 0.0f;
        }
    }
    return 0;
}


// ------------------------------------------------------------------------------------------
// DeferredRenderer::ShadowMapPass (DeferredRenderer.cpp:268-291) with shadowMap.vert:15-17,
// shadowMap.geom:17-38, shadowMap.frag:12-18, restated per texel: every triangle goes through
// the pipeline's own steps -- world position = modelMat * pos (vert), clip = light.viewProj[f] *
// world (geom), FRONT-face culling with front = counter-clockwise in framebuffer space
// (`.cullFront = true`, DeferredRenderer.cpp:100; VulkanWrapper.cpp:941-946): the framebuffer
// area of the Vulkan spec has the sign of -det[c0; c1; c2] over (x, y, w), so a triangle is
// front-facing, and dropped, iff that determinant is negative -- and the fragment of a texel is
// found with a ray through the texel centre (exhaustive over all triangles), depth test LESS
// against the clear value 1.0, no depth clamp.  Point lights: the texel-centre direction of layer
// f is the Vulkan cube-face table inverted (light.viewProj[f] of GPUScene.cpp:268-276 maps to
// exactly that table) and the value is |light.position - fragPos| / zFar; others: the ray runs
// from NDC z = 0 to z = 1 of light.viewProj[0] and the value is its parameter (= gl_FragCoord.z).
// ------------------------------------------------------------------------------------------
namespace {
bool invert4d(const float* m, double inv[16]) {
    double a[4][8];
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) {
            a[r][c] = m[c * 4 + r];
            a[r][4 + c] = r == c ? 1.0 : 0.0;
        }
    for (int i = 0; i < 4; i++) {
        int piv = i;
        for (int r = i + 1; r < 4; r++)
            if (fabs(a[r][i]) > fabs(a[piv][i])) piv = r;
        if (a[piv][i] == 0.0 || std::isnan(a[piv][i])) return false;
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
        for (int c = 0; c < 4; c++) inv[c * 4 + r] = a[r][4 + c];
    return true;
}
} // namespace

int orc_shadow_map_pass(const luzw_light_block* light, const orc_world* world, uint32_t res, float* out) {
    const bool cube = light->type == LUZW_LIGHT_POINT;
    const int layers = cube ? 6 : 1;
    double inv[16];
    float invf[16];
    if (!cube) {
        if (!invert4d(light->view_proj[0], inv)) return -1;
        for (int k = 0; k < 16; k++) invf[k] = (float)inv[k];
    }
    // world-space triangles once: modelMat * vec4(inPosition, 1) (shadowMap.vert:16)
    struct WTri {
        float p[3][3];
    };
    std::vector<WTri> tris;
    for (const InstData& in : world->inst) {
        const MeshData& md = world->meshes[in.mesh];
        for (size_t t = 0; t + 2 < md.idx.size(); t += 3) {
            WTri w;
            for (int k = 0; k < 3; k++) {
                const float* p = &md.pos[3 * (size_t)md.idx[t + k]];
                const V4 wp = mul(in.m, V4{p[0], p[1], p[2], 1.0f});
                w.p[k][0] = wp.x;
                w.p[k][1] = wp.y;
                w.p[k][2] = wp.z;
            }
            tris.push_back(w);
        }
    }
    const V3 eye = v3(light->position[0], light->position[1], light->position[2]);
    for (int layer = 0; layer < layers; layer++) {
        const float* vp = light->view_proj[cube ? layer : 0];
        // front-face culling per triangle, from its clip coordinates (shadowMap.geom:25, :33)
        std::vector<const WTri*> kept;
        for (const WTri& w : tris) {
            double c[3][3];
            for (int k = 0; k < 3; k++) {
                const V4 cp = mul(vp, V4{w.p[k][0], w.p[k][1], w.p[k][2], 1.0f});
                c[k][0] = cp.x;
                c[k][1] = cp.y;
                c[k][2] = cp.w;
            }
            const double D = c[0][0] * (c[1][1] * c[2][2] - c[1][2] * c[2][1]) -
                             c[0][1] * (c[1][0] * c[2][2] - c[1][2] * c[2][0]) +
                             c[0][2] * (c[1][0] * c[2][1] - c[1][1] * c[2][0]);
            if (D > 0.0) kept.push_back(&w); // D < 0: front-facing, culled; D == 0: no area
        }
#pragma omp parallel for schedule(dynamic, 4) num_threads(orc_get_threads())
        for (int y = 0; y < (int)res; y++) {
            for (int x = 0; x < (int)res; x++) {
                const float sc = ((float)x + 0.5f) / (float)res * 2.0f - 1.0f;
                const float tc = ((float)y + 0.5f) / (float)res * 2.0f - 1.0f;
                V3 o, d;
                float tmax;
                if (cube) {
                    o = eye;
                    switch (layer) {
                        case 0: d = v3(1.0f, -tc, -sc); break;
                        case 1: d = v3(-1.0f, -tc, sc); break;
                        case 2: d = v3(sc, 1.0f, tc); break;
                        case 3: d = v3(sc, -1.0f, -tc); break;
                        case 4: d = v3(sc, -tc, 1.0f); break;
                        default: d = v3(-sc, -tc, -1.0f); break;
                    }
                    tmax = 3.0e38f;
                } else {
                    const V4 o4 = mul(invf, V4{sc, tc, 0.0f, 1.0f});
                    o = v3(o4.x, o4.y, o4.z);
                    d = v3(invf[8], invf[9], invf[10]);
                    tmax = 1.0f;
                }
                float best = tmax;
                bool found = false;
                if (!ray_is_nan(o, d) && !(d.x == 0.0f && d.y == 0.0f && d.z == 0.0f)) {
                    const RayPre r = make_ray(o, d);
                    for (const WTri* w : kept) {
                        float t;
                        if (tri_hit(r, w->p[0], w->p[1], w->p[2], 0.0f, best, &t)) {
                            best = t;
                            found = true;
                        }
                    }
                }
                float depth = 1.0f;
                if (found) {
                    const float z = cube ? (best * length(d)) / light->z_far : best;
                    if (z < 1.0f) depth = z;
                }
                out[((size_t)layer * res + y) * res + x] = depth;
            }
        }
    }
    return 0;
}

float orc_shadow_factor(const luzw_light_block* light, const orc_shadow_map* map, const float* frag_pos,
                        const float* shadow_origin) {
    return shadow_map_factor(*light, *map, v3(frag_pos[0], frag_pos[1], frag_pos[2]),
                             v3(shadow_origin[0], shadow_origin[1], shadow_origin[2]));
}

void orc_bind_shadow_maps(const orc_shadow_map* maps, uint32_t n) {
    g_shadow_maps = maps;
    g_n_shadow_maps = n;
}

// shadowMapVolumetricLight.comp:42-74 (weight, samples, decay, density are the shader's locals :57-60)
int orc_volumetric_shadow_pass(const luzw_scene_block* scene, const luzw_light_block* extra_lights, uint32_t n_extra,
                               uint32_t width, uint32_t height, const float* depth, const uint8_t* bn, uint32_t bn_w,
                               uint32_t bn_h, uint32_t frame, uint32_t y0, uint32_t y1, float* light_inout) {
    const int n_lights = scene->num_lights + (int)n_extra;
    const int frame_mod = (int)((int32_t)frame % 128);
    const int W = (int)width, H = (int)height;
    for (int li = 0; li < n_lights; li++)
        if (light_at(scene, extra_lights, li).volumetric_type == 2 &&
            ((uint32_t)li >= g_n_shadow_maps || !g_shadow_maps[li].data))
            return -1;
    const V3 camPos = v3(scene->cam_pos[0], scene->cam_pos[1], scene->cam_pos[2]);
#pragma omp parallel for schedule(dynamic, 4) num_threads(orc_get_threads())
    for (int y = (int)y0; y < (int)y1; y++) {
        for (int x = 0; x < W; x++) {
            const float pu = (float)x / (float)W, pv = (float)y / (float)H; // :48
            const float pixelDepth = bilinear1(depth, W, H, pu, pv);          // :49
            const V3 worldPos = depth_to_world(scene, pu, pv, pixelDepth);
            const uint8_t* t = bn + 4 * ((size_t)(y % (int)bn_h) * bn_w + (size_t)(x % (int)bn_w));
            const float bn_r = (float)t[0] / 255.0f;
            const float noise0 = fractf(bn_r + kGoldenRatio * (float)(128 * 0 + frame_mod));
            V3 radiance = v3(0.0f, 0.0f, 0.0f);
            for (int li = 0; li < n_lights; li++) {
                const luzw_light_block& light = light_at(scene, extra_lights, li);
                if (light.volumetric_type != 2) continue; // VOLUMETRIC_TYPE_SHADOW_MAP, :54
                const float weight = 0.000005f, decay = 1.0f, density = 1.094f;
                const int samples = 128;
                const V3 deltaPos = (camPos - worldPos) * density * (1.0f / (float)samples);
                const float off = noise0 * length(deltaPos);
                V3 samplePos = v3(camPos.x + off, camPos.y + off, camPos.z + off); // :62
                const V3 lcol = v3(light.color[0], light.color[1], light.color[2]);
                for (int i = 0; i < samples; i++) {
                    samplePos = samplePos - deltaPos;
                    const float sh = shadow_map_factor(light, g_shadow_maps[li], samplePos, samplePos); // :22-40
                    radiance = radiance + (1.0f - sh) * lcol * light.intensity * weight * decay;     // :65
                }
            }
            float* px = light_inout + 4 * ((size_t)y * W + x);
            px[0] += radiance.x;
            px[1] += radiance.y;
            px[2] += radiance.z;
        }
    }
    return 0;
}


// present.frag:27-35, :85-95 (imageType 0, debug overlay alpha 0), BGRA8_unorm target
int orc_compose_pass(uint32_t width, uint32_t height, const float* light_in, uint8_t* out_bgra8) {
    const float a = 2.51f, b = 0.03f, c = 2.43f, d = 0.59f, e = 0.14f;
#pragma omp parallel for ORC_NT
    for (int64_t i = 0; i < (int64_t)width * height; i++) {
        float rgb[3];
        for (int k = 0; k < 3; k++) {
            const float x = light_in[4 * i + k];
            const float m = (x * (a * x + b)) / (x * (c * x + d) + e);
            rgb[k] = powf(m, 1.0f / 2.2f);
        }
        auto q = [](float v) -> uint8_t {
            if (!(v > 0.0f)) return 0; // NaN and negatives -> 0
            if (v > 1.0f) v = 1.0f;
            return (uint8_t)(int)floorf(v * 255.0f + 0.5f);
        };
        out_bgra8[4 * i + 0] = q(rgb[2]);
        out_bgra8[4 * i + 1] = q(rgb[1]);
        out_bgra8[4 * i + 2] = q(rgb[0]);
        out_bgra8[4 * i + 3] = 255;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------
// G-buffer producer: opaque.vert:21-31 / opaque.frag:21-59 evaluated at the closest hit of the
// primary ray through each pixel centre (the reference rasterises; coverage differs only at
// triangle edges).  Attachment formats DeferredRenderer.cpp:176-238, clears
// VulkanWrapper.cpp:1194-1196, :1212 (colour 0, depth 1).
// ------------------------------------------------------------------------------------------
namespace {
inline uint8_t unorm8(float v) {
    if (!(v > 0.0f)) return 0;
    if (v > 1.0f) v = 1.0f;
    return (uint8_t)(int)floorf(v * 255.0f + 0.5f);
}
// bilinear REPEAT fetch of an RGBA8 texture, LOD 0
V4 tex_rgba8(const orc_texture& t, float u, float v) {
    const float x = u * (float)t.width - 0.5f, y = v * (float)t.height - 0.5f;
    const float fx0 = floorf(x), fy0 = floorf(y);
    const float fx = x - fx0, fy = y - fy0;
    auto at = [&](int xi, int yi) -> V4 {
        const uint8_t* p = t.rgba8 + 4 * ((size_t)wrapi(yi, (int)t.height) * t.width + wrapi(xi, (int)t.width));
        return {(float)p[0] / 255.0f, (float)p[1] / 255.0f, (float)p[2] / 255.0f, (float)p[3] / 255.0f};
    };
    const int x0 = (int)fx0, y0 = (int)fy0;
    const V4 top = at(x0, y0) * (1.0f - fx) + at(x0 + 1, y0) * fx;
    const V4 bot = at(x0, y0 + 1) * (1.0f - fx) + at(x0 + 1, y0 + 1) * fx;
    return top * (1.0f - fy) + bot * fy;
}
// transpose(inverse(mat3(M))) * n  == cofactor(M)/det * n
inline V3 normal_xform(const float* inv, V3 n) {
    // inv holds rows of A^-1; (A^-1)^T * n = sum_r inv[r][c] * n[r]
    return {inv[0] * n.x + inv[4] * n.y + inv[8] * n.z, inv[1] * n.x + inv[5] * n.y + inv[9] * n.z,
            inv[2] * n.x + inv[6] * n.y + inv[10] * n.z};
}
} // namespace

int orc_gbuffer_pass(const luzw_scene_block* scene, const orc_world* world, const luzw_model_block* models,
                     uint32_t n_models, const orc_texture* textures, uint32_t n_textures, uint32_t width,
                     uint32_t height, int exhaustive, uint32_t y0, uint32_t y1, orc_gbuffer* out) {
    int err = 0;
#pragma omp parallel for schedule(dynamic, 2) ORC_NT
    for (int64_t yy = (int64_t)y0; yy < (int64_t)y1; yy++) {
        for (uint32_t x = 0; x < width; x++) {
            const size_t pix = (size_t)yy * width + x;
            const float u = ((float)x + 0.5f) / (float)width, v = ((float)yy + 0.5f) / (float)height;
            const V3 pn = depth_to_world(scene, u, v, 0.0f);
            const V3 pf = depth_to_world(scene, u, v, 1.0f);
            const V3 d = pf - pn;
            float tmin = 0.0f;
            bool have = false;
            Hit h{};
            V4 albedo{}, emission{};
            float roughness = 0, metallic = 0, occl = 1;
            V3 N{};
            float depth = 1.0f;
            for (int iter = 0; iter < 16 && !have; iter++) {
                // the Opaque Pipeline culls back faces: only front-facing triangles produce fragments
                if (!world_trace(world, pn, d, tmin, 1.0f, true, exhaustive, h, scene->view_proj)) break;
                const InstData& in = world->inst[h.inst];
                const MeshData& md = world->meshes[in.mesh];
                if (in.custom_index >= n_models) {
                    err = -1;
                    break;
                }
                const luzw_model_block& mb = models[in.custom_index];
                const uint32_t i0 = md.idx[3 * h.prim], i1 = md.idx[3 * h.prim + 1], i2 = md.idx[3 * h.prim + 2];
                const float b0 = h.bu, b1 = h.bv, b2 = 1.0f - h.bu - h.bv;
                V3 n0 = v3(0, 0, 0), n1 = n0, n2 = n0;
                V4 tg0{}, tg1{}, tg2{};
                float uv[2] = {0, 0};
                if (md.has_attr) {
                    const float* a0 = &md.attr[9 * (size_t)i0];
                    const float* a1 = &md.attr[9 * (size_t)i1];
                    const float* a2 = &md.attr[9 * (size_t)i2];
                    n0 = v3(a0[0], a0[1], a0[2]);
                    n1 = v3(a1[0], a1[1], a1[2]);
                    n2 = v3(a2[0], a2[1], a2[2]);
                    tg0 = {a0[3], a0[4], a0[5], a0[6]};
                    tg1 = {a1[3], a1[4], a1[5], a1[6]};
                    tg2 = {a2[3], a2[4], a2[5], a2[6]};
                    uv[0] = a0[7] * b0 + a1[7] * b1 + a2[7] * b2;
                    uv[1] = a0[8] * b0 + a1[8] * b1 + a2[8] * b2;
                }
                albedo = {mb.color[0], mb.color[1], mb.color[2], mb.color[3]};
                if (mb.color_map >= 0 && (uint32_t)mb.color_map < n_textures) {
                    const V4 t = tex_rgba8(textures[mb.color_map], uv[0], uv[1]);
                    albedo = {albedo.x * t.x, albedo.y * t.y, albedo.z * t.z, albedo.w * t.w};
                }
                if (albedo.w < 0.5f) { // discard: continue behind this surface
                    tmin = h.t;
                    continue;
                }
                roughness = mb.roughness;
                metallic = mb.metallic;
                occl = 1.0f;
                emission = {mb.emission[0], mb.emission[1], mb.emission[2], 1.0f};
                V3 normalSample = v3(1, 1, 1);
                if (mb.metallic_roughness_map >= 0 && (uint32_t)mb.metallic_roughness_map < n_textures) {
                    const V4 t = tex_rgba8(textures[mb.metallic_roughness_map], uv[0], uv[1]);
                    roughness *= t.y;
                    metallic *= t.z;
                }
                if (mb.ao_map >= 0 && (uint32_t)mb.ao_map < n_textures)
                    occl = tex_rgba8(textures[mb.ao_map], uv[0], uv[1]).x;
                if (mb.normal_map >= 0 && (uint32_t)mb.normal_map < n_textures) {
                    const V4 t = tex_rgba8(textures[mb.normal_map], uv[0], uv[1]);
                    normalSample = v3(t.x, t.y, t.z);
                }
                if (mb.emission_map >= 0 && (uint32_t)mb.emission_map < n_textures) {
                    const V4 t = tex_rgba8(textures[mb.emission_map], uv[0], uv[1]);
                    emission = {emission.x * t.x, emission.y * t.y, emission.z * t.z, emission.w * t.w};
                }
                // opaque.vert:25-30 per vertex, then interpolate
                auto vert = [&](V3 n, V4 tg, V3& fn, V3& ft, V3& fb) {
                    ft = normalize(normal_xform(in.inv, v3(tg.x, tg.y, tg.z)));
                    fn = normalize(normal_xform(in.inv, n));
                    ft = normalize(ft - dot(ft, fn) * fn);
                    fb = cross(fn, ft) * tg.w;
                };
                V3 fn0, ft0, fb0, fn1, ft1, fb1, fn2, ft2, fb2;
                vert(n0, tg0, fn0, ft0, fb0);
                vert(n1, tg1, fn1, ft1, fb1);
                vert(n2, tg2, fn2, ft2, fb2);
                const V3 fragNormal = fn0 * b0 + fn1 * b1 + fn2 * b2;
                const V3 fragTangent = ft0 * b0 + ft1 * b1 + ft2 * b2;
                const V3 fragBitan = fb0 * b0 + fb1 * b1 + fb2 * b2;
                const bool tangent_zero = fragTangent.x == 0.0f && fragTangent.y == 0.0f && fragTangent.z == 0.0f;
                const bool ns_one = normalSample.x == 1.0f && normalSample.y == 1.0f && normalSample.z == 1.0f;
                if (tangent_zero || ns_one) {
                    N = normalize(fragNormal);
                } else {
                    const V3 ts = normalize(normalSample * 2.0f - v3(1, 1, 1));
                    N = normalize(fragTangent * ts.x + fragBitan * ts.y + fragNormal * ts.z);
                }
                const V3 wp = pn + d * h.t;
                const V4 clip = mul(scene->view_proj, V4{wp.x, wp.y, wp.z, 1.0f});
                depth = clip.z / clip.w;
                have = true;
            }
            uint8_t* a8 = out->albedo + 4 * pix;
            float* n4 = out->normal + 4 * pix;
            uint8_t* m8 = out->material + 4 * pix;
            uint8_t* e8 = out->emission + 4 * pix;
            if (!have) {
                memset(a8, 0, 4);
                memset(m8, 0, 4);
                memset(e8, 0, 4);
                n4[0] = n4[1] = n4[2] = n4[3] = 0.0f;
                out->depth[pix] = 1.0f;
                continue;
            }
            a8[0] = unorm8(albedo.x);
            a8[1] = unorm8(albedo.y);
            a8[2] = unorm8(albedo.z);
            a8[3] = unorm8(albedo.w);
            n4[0] = N.x;
            n4[1] = N.y;
            n4[2] = N.z;
            n4[3] = 1.0f;
            m8[0] = unorm8(roughness);
            m8[1] = unorm8(metallic);
            m8[2] = unorm8(occl);
            m8[3] = 255;
            e8[0] = unorm8(emission.x);
            e8[1] = unorm8(emission.y);
            e8[2] = unorm8(emission.z);
            e8[3] = unorm8(emission.w);
            out->depth[pix] = depth;
        }
    }
    return err;
}

} // extern "C"

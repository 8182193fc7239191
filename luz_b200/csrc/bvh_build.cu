// bvh_build.cu -- deterministic GPU builder for the 8-wide compressed BVH.
//
// Replaces the driver-side vkCmdBuildAccelerationStructuresKHR that the reference reaches through
// vkw::CmdBuildBLAS / vkw::CmdBuildTLAS (source/Graphics/VulkanWrapper.cpp:1089-1104, :1106-1139).
// Pipeline: centroid bounds -> 30-bit Morton codes -> stable LSD radix sort (own kernels) ->
// Karras LBVH (index tie-break for duplicate codes) -> bottom-up boxes + SAH cost tables ->
// level-synchronous optimal collapse to 8-wide nodes (Ylitie et al. 2017 dynamic programme) with
// prefix-sum allocation (no allocation atomics, so the node/primitive layout is identical on every
// run and every GPU).  Small trees (TLAS rebuilds) are collapsed by one CTA without host round trips.
#include "bvh.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>

namespace luz {

namespace {

constexpr uint32_t kNone = 0xFFFFFFFFu;

#define LUZ_CK(expr)                         \
    do {                                     \
        cudaError_t _e = (expr);             \
        if (_e != cudaSuccess) return _e;    \
    } while (0)

inline uint32_t div_up(uint32_t a, uint32_t b) { return (a + b - 1) / b; }

// ---- ordered-int encoding of floats for atomicMin/Max ------------------------------------------
__device__ __forceinline__ int f2ord(float f) {
    int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7FFFFFFF;
}
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7FFFFFFF); }

__global__ void k_init_bounds(int* b) {
    if (threadIdx.x < 3) b[threadIdx.x] = 0x7F800000;                       // +inf
    else if (threadIdx.x < 6) b[threadIdx.x] = (int)0xFF800000 ^ 0x7FFFFFFF; // -inf, ordered
}

__global__ void k_centroid_bounds(const BoxF* __restrict__ boxes, uint32_t n, int* __restrict__ b) {
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const BoxF bx = boxes[i];
        const float c[3] = {0.5f * (bx.lox + bx.hix), 0.5f * (bx.loy + bx.hiy), 0.5f * (bx.loz + bx.hiz)};
#pragma unroll
        for (int k = 0; k < 3; k++) {
            lo[k] = fminf(lo[k], c[k]);
            hi[k] = fmaxf(hi[k], c[k]);
        }
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
        for (int off = 16; off; off >>= 1) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(0xFFFFFFFFu, lo[k], off));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xFFFFFFFFu, hi[k], off));
        }
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            if (lo[k] <= hi[k]) {
                atomicMin(&b[k], f2ord(lo[k]));
                atomicMax(&b[3 + k], f2ord(hi[k]));
            }
        }
    }
}

__device__ __forceinline__ uint32_t expand_bits10(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}

__global__ void k_morton(const BoxF* __restrict__ boxes, uint32_t n, const int* __restrict__ b,
                         uint32_t* __restrict__ codes, uint32_t* __restrict__ idx) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float lo[3] = {ord2f(b[0]), ord2f(b[1]), ord2f(b[2])};
    const float hi[3] = {ord2f(b[3]), ord2f(b[4]), ord2f(b[5])};
    const BoxF bx = boxes[i];
    const float c[3] = {0.5f * (bx.lox + bx.hix), 0.5f * (bx.loy + bx.hiy), 0.5f * (bx.loz + bx.hiz)};
    uint32_t q[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float ext = hi[k] - lo[k];
        float u = ext > 0.0f ? (c[k] - lo[k]) / ext : 0.0f;
        u = fminf(fmaxf(u * 1024.0f, 0.0f), 1023.0f);
        q[k] = (uint32_t)u;
    }
    codes[i] = (expand_bits10(q[0]) << 2) | (expand_bits10(q[1]) << 1) | expand_bits10(q[2]);
    idx[i] = i;
}

// ---- stable LSD radix sort, 8 bits per pass -----------------------------------------------------
constexpr int kSortThreads = 256;
constexpr int kSortItems = 8;
constexpr int kSortTile = kSortThreads * kSortItems;

__global__ void k_sort_hist(const uint32_t* __restrict__ keys, uint32_t n, int shift, uint32_t nblocks,
                            uint32_t* __restrict__ hist) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t base = blockIdx.x * kSortTile;
    for (int r = 0; r < kSortItems; r++) {
        const uint32_t i = base + r * kSortThreads + threadIdx.x;
        if (i < n) atomicAdd(&h[(keys[i] >> shift) & 0xFFu], 1u);
    }
    __syncthreads();
    hist[threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];
}

// exclusive scan of n u64 values by one block (values are small; two u32 lanes packed per u64)
__global__ void k_scan_u64(const uint64_t* __restrict__ in, uint64_t* __restrict__ out, uint32_t n,
                           uint64_t* __restrict__ total) {
    __shared__ uint64_t warp_sums[32];
    __shared__ uint64_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t base = 0; base < n; base += blockDim.x) {
        const uint32_t i = base + threadIdx.x;
        const uint64_t v = i < n ? in[i] : 0ull;
        uint64_t x = v;
        for (int off = 1; off < 32; off <<= 1) {
            const uint64_t y = __shfl_up_sync(0xFFFFFFFFu, x, off);
            if (lane >= off) x += y;
        }
        if (lane == 31) warp_sums[warp] = x;
        __syncthreads();
        if (warp == 0) {
            uint64_t w = lane < (int)(blockDim.x >> 5) ? warp_sums[lane] : 0ull;
            for (int off = 1; off < 32; off <<= 1) {
                const uint64_t y = __shfl_up_sync(0xFFFFFFFFu, w, off);
                if (lane >= off) w += y;
            }
            warp_sums[lane] = w; // inclusive over warps
        }
        __syncthreads();
        const uint64_t warp_off = warp ? warp_sums[warp - 1] : 0ull;
        const uint64_t c = carry;
        if (i < n) out[i] = c + warp_off + x - v;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = c + warp_off + x;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total) *total = carry;
}

__global__ void k_u32_to_u64(const uint32_t* __restrict__ in, uint64_t* __restrict__ out, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i];
}

__global__ void k_sort_scatter(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                               uint32_t n, int shift, uint32_t nblocks, const uint64_t* __restrict__ scanned,
                               uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out) {
    __shared__ uint32_t running[256];
    __shared__ uint32_t warp_cnt[kSortThreads / 32][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    running[threadIdx.x] = (uint32_t)scanned[threadIdx.x * nblocks + blockIdx.x];
    for (int w = 0; w < kSortThreads / 32; w++) warp_cnt[w][threadIdx.x] = 0;
    __syncthreads();
    const uint32_t base = blockIdx.x * kSortTile;
    for (int r = 0; r < kSortItems; r++) {
        const uint32_t i = base + r * kSortThreads + threadIdx.x;
        const bool valid = i < n;
        uint32_t key = 0, val = 0;
        if (valid) {
            key = keys_in[i];
            val = vals_in[i];
        }
        const uint32_t digit = valid ? ((key >> shift) & 0xFFu) : 256u;
        const uint32_t peers = __match_any_sync(0xFFFFFFFFu, digit);
        const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
        if (valid && rank == 0) warp_cnt[warp][digit] = __popc(peers);
        __syncthreads();
        if (valid) {
            uint32_t pre = 0;
            for (int w = 0; w < warp; w++) pre += warp_cnt[w][digit];
            const uint32_t pos = running[digit] + pre + rank;
            keys_out[pos] = key;
            vals_out[pos] = val;
        }
        __syncthreads();
        {
            uint32_t s = 0;
            for (int w = 0; w < kSortThreads / 32; w++) {
                s += warp_cnt[w][threadIdx.x];
                warp_cnt[w][threadIdx.x] = 0;
            }
            running[threadIdx.x] += s;
        }
        __syncthreads();
    }
}

// ---- Karras 2012 LBVH ----------------------------------------------------------------------------
__device__ __forceinline__ int lcp(const uint32_t* __restrict__ codes, int n, int i, long long j) {
    if (j < 0 || j >= n) return -1;
    const uint32_t a = codes[i], b = codes[j];
    if (a == b) return 32 + __clz((uint32_t)i ^ (uint32_t)j);
    return __clz(a ^ b);
}

// node refs: [0, n-1) internal nodes, [n-1, 2n-1) leaves (sorted position + n-1)
__global__ void k_karras(const uint32_t* __restrict__ codes, int n, uint32_t* __restrict__ left,
                         uint32_t* __restrict__ right, uint32_t* __restrict__ first, uint32_t* __restrict__ last,
                         uint32_t* __restrict__ parent) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int d = (lcp(codes, n, i, i + 1) - lcp(codes, n, i, i - 1)) >= 0 ? 1 : -1;
    const int dmin = lcp(codes, n, i, i - d);
    long long lmax = 2;
    while (lcp(codes, n, i, i + lmax * d) > dmin) lmax *= 2;
    long long l = 0;
    for (long long t = lmax / 2; t >= 1; t /= 2)
        if (lcp(codes, n, i, i + (l + t) * d) > dmin) l += t;
    const long long j = i + l * d;
    const int dnode = lcp(codes, n, i, j);
    long long s = 0;
    long long t = l;
    do {
        t = (t + 1) / 2;
        if (lcp(codes, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    const long long gamma = i + s * d + (d < 0 ? -1 : 0);
    const long long lo = i < j ? i : j, hi = i < j ? j : i;
    const uint32_t l_ref = (lo == gamma) ? (uint32_t)(n - 1 + gamma) : (uint32_t)gamma;
    const uint32_t r_ref = (hi == gamma + 1) ? (uint32_t)(n - 1 + gamma + 1) : (uint32_t)(gamma + 1);
    left[i] = l_ref;
    right[i] = r_ref;
    first[i] = (uint32_t)lo;
    last[i] = (uint32_t)hi;
    parent[l_ref] = (uint32_t)i;
    parent[r_ref] = (uint32_t)i;
    if (i == 0) parent[0] = kNone;
}

__device__ __forceinline__ BoxF box_union(const BoxF a, const BoxF b) {
    BoxF r;
    r.lox = fminf(a.lox, b.lox);
    r.loy = fminf(a.loy, b.loy);
    r.loz = fminf(a.loz, b.loz);
    r.hix = fmaxf(a.hix, b.hix);
    r.hiy = fmaxf(a.hiy, b.hiy);
    r.hiz = fmaxf(a.hiz, b.hiz);
    return r;
}
__device__ __forceinline__ BoxF load_box_cg(const BoxF* p) {
    const float* f = reinterpret_cast<const float*>(p);
    BoxF r;
    r.lox = __ldcg(f + 0);
    r.loy = __ldcg(f + 1);
    r.loz = __ldcg(f + 2);
    r.hix = __ldcg(f + 3);
    r.hiy = __ldcg(f + 4);
    r.hiz = __ldcg(f + 5);
    return r;
}

// ---- optimal collapse: cost tables (Ylitie, Karras, Laine 2017, section 3) ----------------------------
// For a binary node n and a budget i of 1..7 slots of a wide node, C(n, i) is the cheapest SAH cost of
// representing n's subtree by at most i slots (each a leaf slot of <= max_leaf primitives or an
// internal child node, which in turn distributes 8 slots):
//   C_leaf(n) = A(n) * P(n) * c_prim            if P(n) <= max_leaf, else infinity
//   D(n, j)   = min_{0<k<j} C(left, k) + C(right, j-k)
//   C(n, 1)   = min(C_leaf(n), D(n, 8) + A(n) * c_node)
//   C(n, i)   = min(D(n, i), C(n, i-1))
// cost[8n + 0] holds C_leaf(n), cost[8n + i] holds C(n, i).  A single primitive costs A * c_prim whatever
// the budget.  Every value is a min over plain fp32 sums evaluated in one fixed order, so the plan step
// below re-derives the arg-mins bit-exactly and the tree is identical on every run and every GPU.
struct CostParams {
    float c_node, c_prim;
    uint32_t max_leaf;
};

__device__ __forceinline__ float box_area(const BoxF b) {
    const float dx = b.hix - b.lox, dy = b.hiy - b.loy, dz = b.hiz - b.loz;
    const float a = dx * dy + dy * dz + dz * dx;
    return a >= 0.0f ? a : 0.0f; // NaN boxes cost nothing
}

__device__ __forceinline__ void distribute_costs(const float* cl, const float* cr, float* dist /*[9], 2..8 used*/) {
#pragma unroll
    for (int j = 2; j <= 8; j++) {
        float d = INFINITY;
#pragma unroll
        for (int k = 1; k <= 7; k++)
            if (k < j && j - k <= 7) d = fminf(d, cl[k] + cr[j - k]);
        dist[j] = d;
    }
}

__device__ __forceinline__ void node_costs(const float* cl, const float* cr, const float area, const uint32_t count,
                                           const CostParams p, float* c /*[8]*/) {
    float dist[9];
    distribute_costs(cl, cr, dist);
    c[0] = count <= p.max_leaf ? area * (float)count * p.c_prim : INFINITY;
    c[1] = fminf(c[0], dist[8] + area * p.c_node);
#pragma unroll
    for (int i = 2; i <= 7; i++) c[i] = fminf(dist[i], c[i - 1]);
}

template <bool CG>
__device__ __forceinline__ void load_costs(const float* __restrict__ cost, const BoxF* __restrict__ box, const int n,
                                           const uint32_t ref, const float c_prim, float* c /*[8]*/) {
    if (ref >= (uint32_t)(n - 1)) {
        const float a = box_area(CG ? load_box_cg(&box[ref]) : box[ref]) * c_prim;
#pragma unroll
        for (int i = 0; i < 8; i++) c[i] = a;
    } else {
        const float4* p4 = reinterpret_cast<const float4*>(cost) + 2 * (size_t)ref;
        const float4 a = CG ? __ldcg(p4) : p4[0], b = CG ? __ldcg(p4 + 1) : p4[1];
        c[0] = a.x, c[1] = a.y, c[2] = a.z, c[3] = a.w, c[4] = b.x, c[5] = b.y, c[6] = b.z, c[7] = b.w;
    }
}

// bottom-up: boxes and cost tables of the binary tree (the last of the two children to arrive continues)
__global__ void k_fit(const BoxF* __restrict__ prim_boxes, const uint32_t* __restrict__ sorted_idx, int n,
                      const uint32_t* __restrict__ left, const uint32_t* __restrict__ right,
                      const uint32_t* __restrict__ first, const uint32_t* __restrict__ last,
                      const uint32_t* __restrict__ parent, uint32_t* __restrict__ flags, BoxF* __restrict__ box,
                      float* __restrict__ cost, const CostParams cp) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    box[n - 1 + k] = prim_boxes[sorted_idx[k]];
    if (n == 1) return;
    __threadfence();
    uint32_t cur = parent[n - 1 + k];
    while (cur != kNone) {
        const uint32_t old = atomicAdd(&flags[cur], 1u);
        if (old == 0) return; // the sibling subtree is not finished: its thread will continue
        __threadfence();
        const uint32_t l = left[cur], r = right[cur];
        const BoxF a = load_box_cg(&box[l]), b = load_box_cg(&box[r]);
        const BoxF u = box_union(a, b);
        box[cur] = u;
        float cl[8], cr[8], c[8];
        load_costs<true>(cost, box, n, l, cp.c_prim, cl);
        load_costs<true>(cost, box, n, r, cp.c_prim, cr);
        node_costs(cl, cr, box_area(u), last[cur] - first[cur] + 1u, cp, c);
        float4* dst = reinterpret_cast<float4*>(cost) + 2 * (size_t)cur;
        dst[0] = make_float4(c[0], c[1], c[2], c[3]);
        dst[1] = make_float4(c[4], c[5], c[6], c[7]);
        __threadfence();
        cur = parent[cur];
    }
}

// ---- collapse to 8-wide ----------------------------------------------------------------------------
struct Tree2 {
    const uint32_t* left;
    const uint32_t* right;
    const uint32_t* first;
    const uint32_t* last;
    const BoxF* box;
    int n;
};
__device__ __forceinline__ uint32_t ref_count(const Tree2& t, uint32_t ref) {
    return ref >= (uint32_t)(t.n - 1) ? 1u : t.last[ref] - t.first[ref] + 1u;
}
__device__ __forceinline__ uint32_t ref_first(const Tree2& t, uint32_t ref) {
    return ref >= (uint32_t)(t.n - 1) ? ref - (uint32_t)(t.n - 1) : t.first[ref];
}
constexpr uint32_t kSlotInternal = 0x80000000u; // flag on a planned slot: the subtree becomes a child node

// Chooses the (at most 8) slots of the wide node rooted at binary node r by following the arg-mins of the
// cost tables, left subtree first (so the slots are in Morton order).  s[k] = binary ref | kSlotInternal.
// Returns the slot count; *counts = (#leaf primitives << 32) | #internal children.
__device__ __forceinline__ int plan_slots(const Tree2& t, const float* __restrict__ cost, const uint32_t r,
                                          const CostParams p, uint32_t* s, uint64_t* counts) {
    int ns = 0;
    uint32_t n_int = 0, n_prim = 0;
    if (r >= (uint32_t)(t.n - 1)) { // a tree of one primitive: the root holds it in its only slot
        s[0] = r;
        *counts = 1ull << 32;
        return 1;
    }
    uint32_t st_ref[8];
    int st_budget[8];
    int sp = 0;
    st_ref[sp] = r;
    st_budget[sp++] = 8;
    while (sp) {
        const uint32_t ref = st_ref[--sp];
        int i = st_budget[sp];
        if (ref >= (uint32_t)(t.n - 1)) { // a single primitive
            s[ns++] = ref;
            n_prim++;
            continue;
        }
        if (i < 8) {
            float c[8];
            load_costs<false>(cost, t.box, t.n, ref, p.c_prim, c);
            while (i > 1 && c[i] == c[i - 1]) i--;
            if (i == 1) {
                const uint32_t cnt = ref_count(t, ref);
                if (cnt <= p.max_leaf && c[0] <= c[1]) {
                    s[ns++] = ref;
                    n_prim += cnt;
                } else {
                    s[ns++] = ref | kSlotInternal;
                    n_int++;
                }
                continue;
            }
        }
        // split the budget i between the two children: first k reaching the minimum wins
        const uint32_t l = t.left[ref], rr = t.right[ref];
        float cl[8], cr[8];
        load_costs<false>(cost, t.box, t.n, l, p.c_prim, cl);
        load_costs<false>(cost, t.box, t.n, rr, p.c_prim, cr);
        float best = INFINITY;
        int bk = 1;
        for (int k = 1; k <= 7; k++)
            if (k < i && i - k <= 7) {
                const float v = cl[k] + cr[i - k];
                if (v < best) {
                    best = v;
                    bk = k;
                }
            }
        if (i - bk > 7) bk = i - 7; // all-infinite tables: any feasible split
        st_ref[sp] = rr;
        st_budget[sp++] = i - bk;
        st_ref[sp] = l;
        st_budget[sp++] = bk;
    }
    *counts = ((uint64_t)n_prim << 32) | n_int;
    return ns;
}

// plan: per level item, choose up to 8 slots; counts[i] = (#leaf primitives << 32) | #internal children
__global__ void k_collapse_plan(Tree2 t, const float* __restrict__ cost, const uint32_t* __restrict__ items,
                                uint32_t n_items, CostParams p, uint32_t* __restrict__ slots,
                                uint64_t* __restrict__ counts) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_items) return;
    uint32_t s[8];
    uint64_t cnt;
    const int ns = plan_slots(t, cost, items[i], p, s, &cnt);
    for (int k = 0; k < 8; k++) slots[i * 8 + k] = k < ns ? s[k] : kNone;
    counts[i] = cnt;
}

__device__ __forceinline__ void write_node(WideNode* dst, const uint32_t child_base, const uint32_t prim_base,
                                           const uint8_t* meta, const uint8_t imask, const BoxF* cb, const bool* used) {
    WideNode w;
    w.child_base_imask = (child_base & 0x00FFFFFFu) | ((uint32_t)imask << 24);
    w.prim_base = prim_base;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        w.meta[k] = meta[k];
        const bool u = used[k];
        w.lox[k] = u ? cb[k].lox : INFINITY;
        w.loy[k] = u ? cb[k].loy : INFINITY;
        w.loz[k] = u ? cb[k].loz : INFINITY;
        w.hix[k] = u ? cb[k].hix : -INFINITY;
        w.hiy[k] = u ? cb[k].hiy : -INFINITY;
        w.hiz[k] = u ? cb[k].hiz : -INFINITY;
    }
    const uint4* src = reinterpret_cast<const uint4*>(&w);
    uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
    for (int k = 0; k < (int)(sizeof(WideNode) / 16); k++) d4[k] = src[k];
}

// Writes the wide node planned as s[0..8) (item i of its level) and queues its internal children.
__device__ __forceinline__ void emit_node(const Tree2& t, const uint32_t r, const uint32_t* s, const uint32_t int_off,
                                          const uint32_t prim_off, const uint32_t node_index,
                                          const uint32_t next_level_start, const uint32_t prim_cursor,
                                          const uint32_t* __restrict__ sorted_idx, WideNode* __restrict__ nodes,
                                          BoxF* __restrict__ node_bounds, uint32_t* __restrict__ prim_order,
                                          uint32_t* __restrict__ next_items) {
    const uint32_t child_base = next_level_start + int_off;
    const uint32_t prim_base = prim_cursor + prim_off;
    uint8_t meta[8];
    bool used[8];
    BoxF cb[8];
    uint8_t imask = 0;
    uint32_t n_int = 0, poff = 0;
    for (int k = 0; k < 8; k++) {
        meta[k] = 0;
        used[k] = false;
        if (s[k] == kNone) continue;
        const uint32_t ref = s[k] & ~kSlotInternal;
        used[k] = true;
        cb[k] = t.box[ref];
        if (!(s[k] & kSlotInternal)) {
            const uint32_t cnt = ref_count(t, ref);
            const uint32_t f = ref_first(t, ref);
            for (uint32_t q = 0; q < cnt; q++) prim_order[prim_base + poff + q] = sorted_idx[f + q];
            const uint32_t unary = (1u << cnt) - 1u; // 1 -> 001, 2 -> 011, 3 -> 111
            meta[k] = (uint8_t)((unary << 5) | poff);
            poff += cnt;
        } else {
            meta[k] = (uint8_t)((1u << 5) | (24u + (uint32_t)k));
            imask |= (uint8_t)(1u << k);
            next_items[int_off + n_int] = ref;
            n_int++;
        }
    }
    const BoxF nb = t.box[r];
    node_bounds[node_index] = nb;
    write_node(nodes + node_index, child_base, prim_base, meta, imask, cb, used);
}

__global__ void k_collapse_emit(Tree2 t, const uint32_t* __restrict__ items, uint32_t n_items,
                                const uint32_t* __restrict__ slots, const uint64_t* __restrict__ scanned,
                                uint32_t level_start, uint32_t next_level_start, uint32_t prim_cursor,
                                const uint32_t* __restrict__ sorted_idx, WideNode* __restrict__ nodes,
                                BoxF* __restrict__ node_bounds, uint32_t* __restrict__ prim_order,
                                uint32_t* __restrict__ next_items) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_items) return;
    uint32_t s[8];
    for (int k = 0; k < 8; k++) s[k] = slots[i * 8 + k];
    const uint64_t sc = scanned[i];
    emit_node(t, items[i], s, (uint32_t)(sc & 0xFFFFFFFFull), (uint32_t)(sc >> 32), level_start + i, next_level_start,
              prim_cursor, sorted_idx, nodes, node_bounds, prim_order, next_items);
}

// The whole collapse of a small tree (TLAS rebuilds, small meshes) by one CTA: the level loop, the prefix
// sums that allocate child nodes and leaf primitives, and the level table stay on the device, so a rebuild
// costs no host round trip per level.  level_table: [0] = #levels, [1] = #nodes, [2] = error flag,
// [4 + 2l] = first node of level l, [5 + 2l] = node count.
constexpr int kSmallThreads = 1024;
constexpr uint32_t kMaxLevels = 60;
__global__ void __launch_bounds__(kSmallThreads) k_collapse_small(Tree2 t, const float* __restrict__ cost, CostParams p,
                                                               const uint32_t* __restrict__ sorted_idx,
                                                               WideNode* __restrict__ nodes, BoxF* __restrict__ node_bounds,
                                                               uint32_t* __restrict__ prim_order, uint32_t* items,
                                                               uint32_t* next_items, uint32_t node_capacity,
                                                               uint32_t* __restrict__ level_table) {
    __shared__ uint64_t warp_sums[32];
    __shared__ uint64_t carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) items[0] = 0; // the root of the binary tree (internal node 0, or leaf 0 when n == 1)
    uint32_t level_start = 0, level_count = 1, prim_cursor = 0, level = 0, err = 0;
    while (level_count) {
        if (threadIdx.x == 0) carry = 0;
        __syncthreads();
        const uint32_t next_start = level_start + level_count;
        for (uint32_t base = 0; base < level_count; base += kSmallThreads) {
            const uint32_t i = base + threadIdx.x;
            const bool valid = i < level_count;
            uint32_t s[8];
            uint64_t v = 0;
            uint32_t r = 0;
            if (valid) {
                r = items[i];
                const int ns = plan_slots(t, cost, r, p, s, &v);
                for (int k = ns; k < 8; k++) s[k] = kNone;
            }
            uint64_t x = v;
            for (int off = 1; off < 32; off <<= 1) {
                const uint64_t y = __shfl_up_sync(0xFFFFFFFFu, x, off);
                if (lane >= off) x += y;
            }
            if (lane == 31) warp_sums[warp] = x;
            __syncthreads();
            if (warp == 0) {
                uint64_t w = warp_sums[lane];
                for (int off = 1; off < 32; off <<= 1) {
                    const uint64_t y = __shfl_up_sync(0xFFFFFFFFu, w, off);
                    if (lane >= off) w += y;
                }
                warp_sums[lane] = w; // inclusive over warps
            }
            __syncthreads();
            const uint64_t c = carry;
            const uint64_t excl = c + (warp ? warp_sums[warp - 1] : 0ull) + x - v;
            if (valid)
                emit_node(t, r, s, (uint32_t)(excl & 0xFFFFFFFFull), (uint32_t)(excl >> 32), level_start + i, next_start,
                          prim_cursor, sorted_idx, nodes, node_bounds, prim_order, next_items);
            __syncthreads();
            if (threadIdx.x == kSmallThreads - 1) carry = excl + v;
            __syncthreads();
        }
        const uint64_t tot = carry;
        if (threadIdx.x == 0 && level < kMaxLevels) {
            level_table[4 + 2 * level] = level_start;
            level_table[5 + 2 * level] = level_count;
        }
        level++;
        level_start = next_start;
        level_count = (uint32_t)(tot & 0xFFFFFFFFull);
        prim_cursor += (uint32_t)(tot >> 32);
        uint32_t* tmp = items;
        items = next_items;
        next_items = tmp;
        if ((size_t)level_start + level_count > node_capacity || level >= kMaxLevels ||
            (size_t)level_start + level_count > kMaxWideNodes) {
            err = 1;
            break;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        level_table[0] = level;
        level_table[1] = level_start;
        level_table[2] = err;
    }
}

__global__ void k_empty_root(WideNode* nodes, BoxF* node_bounds) {
    uint8_t meta[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    bool used[8] = {false, false, false, false, false, false, false, false};
    BoxF cb[8];
    BoxF nb = {0, 0, 0, 0, 0, 0};
    node_bounds[0] = nb;
    write_node(nodes, 0, 0, meta, 0, cb, used);
}

// refit one level: every node rebuilds its child boxes from leaf primitive boxes / child node bounds
__global__ void k_refit_level(WideNode* __restrict__ nodes, BoxF* __restrict__ node_bounds, uint32_t level_start,
                              uint32_t n_nodes, const BoxF* __restrict__ prim_boxes,
                              const uint32_t* __restrict__ prim_order) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    WideNode* node = nodes + level_start + i;
    const uint4 hdr = *reinterpret_cast<const uint4*>(node);
    const uint32_t child_base = hdr.x & 0x00FFFFFFu, imask = hdr.x >> 24, prim_base = hdr.y;
    uint8_t meta[8];
    bool used[8];
    BoxF cb[8];
    BoxF nb = {INFINITY, INFINITY, INFINITY, -INFINITY, -INFINITY, -INFINITY};
    bool any = false;
    for (int k = 0; k < 8; k++) {
        meta[k] = (uint8_t)(((k < 4 ? hdr.z : hdr.w) >> (8 * (k & 3))) & 0xFFu);
        used[k] = meta[k] != 0;
        if (!used[k]) continue;
        BoxF b;
        if (imask & (1u << k)) {
            const uint32_t rel = __popc(imask & ((1u << k) - 1u));
            b = node_bounds[child_base + rel];
        } else {
            const uint32_t off = meta[k] & 31u;
            const uint32_t cnt = __popc((uint32_t)(meta[k] >> 5));
            b = prim_boxes[prim_order[prim_base + off]];
            for (uint32_t q = 1; q < cnt; q++) b = box_union(b, prim_boxes[prim_order[prim_base + off + q]]);
        }
        cb[k] = b;
        nb = any ? box_union(nb, b) : b;
        any = true;
    }
    if (!any) nb = BoxF{0, 0, 0, 0, 0, 0};
    node_bounds[level_start + i] = nb;
    write_node(node, child_base, prim_base, meta, (uint8_t)imask, cb, used);
}

// ---- mesh / instance helpers -------------------------------------------------------------------------
__global__ void k_triangle_boxes(const uint8_t* __restrict__ verts, uint32_t stride, const uint32_t* __restrict__ idx,
                                 uint32_t n_tris, BoxF* __restrict__ boxes) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tris) return;
    BoxF b = {INFINITY, INFINITY, INFINITY, -INFINITY, -INFINITY, -INFINITY};
    for (int k = 0; k < 3; k++) {
        const float* p = reinterpret_cast<const float*>(verts + (size_t)idx[3 * t + k] * stride);
        b.lox = fminf(b.lox, p[0]);
        b.loy = fminf(b.loy, p[1]);
        b.loz = fminf(b.loz, p[2]);
        b.hix = fmaxf(b.hix, p[0]);
        b.hiy = fmaxf(b.hiy, p[1]);
        b.hiz = fmaxf(b.hiz, p[2]);
    }
    boxes[t] = b;
}

__global__ void k_gather_triangles(const uint8_t* __restrict__ verts, uint32_t stride,
                                   const uint32_t* __restrict__ idx, const uint32_t* __restrict__ prim_order,
                                   uint32_t n_tris, WideTri* __restrict__ tris) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_tris) return;
    const uint32_t t = prim_order[k];
    const float* p0 = reinterpret_cast<const float*>(verts + (size_t)idx[3 * t + 0] * stride);
    const float* p1 = reinterpret_cast<const float*>(verts + (size_t)idx[3 * t + 1] * stride);
    const float* p2 = reinterpret_cast<const float*>(verts + (size_t)idx[3 * t + 2] * stride);
    // explicitly rounded (no contraction): cross(q, p) must be the exact negation of cross(p, q)
    const float3 a = f3(p0[0], p0[1], p0[2]), b = f3(p1[0], p1[1], p1[2]), c = f3(p2[0], p2[1], p2[2]);
    auto crs = [](const float3 q, const float3 p) {
        return f3(__fsub_rn(__fmul_rn(q.y, p.z), __fmul_rn(q.z, p.y)), __fsub_rn(__fmul_rn(q.z, p.x), __fmul_rn(q.x, p.z)),
                  __fsub_rn(__fmul_rn(q.x, p.y), __fmul_rn(q.y, p.x)));
    };
    auto sub = [](const float3 q, const float3 p) { return f3(__fsub_rn(q.x, p.x), __fsub_rn(q.y, p.y), __fsub_rn(q.z, p.z)); };
    const float3 mu = crs(c, b), mv = crs(a, c), mw = crs(b, a);
    const float3 eu = sub(b, c), ev = sub(c, a), ew = sub(a, b);
    const float3 n = crs(sub(b, a), sub(c, a));
    const float kk = __fadd_rn(__fadd_rn(__fmul_rn(n.x, a.x), __fmul_rn(n.y, a.y)), __fmul_rn(n.z, a.z));
    WideTri w;
    w.mu = make_float4(mu.x, mu.y, mu.z, __uint_as_float(t));
    w.eu = make_float4(eu.x, eu.y, eu.z, kk);
    w.mv = make_float4(mv.x, mv.y, mv.z, n.x);
    w.ev = make_float4(ev.x, ev.y, ev.z, n.y);
    w.mw = make_float4(mw.x, mw.y, mw.z, n.z);
    w.ew = make_float4(ew.x, ew.y, ew.z, 0.0f);
    tris[k] = w;
}

// inverse of rows 0..2 of a column-major mat4, in explicitly rounded fp32 (no contraction) so that
// it is bitwise reproducible and matches a plain CPU evaluation of the same cofactor formula
__global__ void k_instance_prepare(const InstanceIn* __restrict__ in, uint32_t n, BoxF* __restrict__ boxes,
                                   InstanceRec* __restrict__ recs, InstanceMeta* __restrict__ meta) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const InstanceIn I = in[i];
    const float* m = I.m;
    const float a00 = m[0], a01 = m[4], a02 = m[8], t0 = m[12];
    const float a10 = m[1], a11 = m[5], a12 = m[9], t1 = m[13];
    const float a20 = m[2], a21 = m[6], a22 = m[10], t2 = m[14];
#define MS(a, b, c, d) __fsub_rn(__fmul_rn(a, b), __fmul_rn(c, d))
    const float c00 = MS(a11, a22, a12, a21);
    const float c01 = MS(a12, a20, a10, a22);
    const float c02 = MS(a10, a21, a11, a20);
    const float det = __fadd_rn(__fadd_rn(__fmul_rn(a00, c00), __fmul_rn(a01, c01)), __fmul_rn(a02, c02));
    const float id = __fdiv_rn(1.0f, det);
    const float i00 = __fmul_rn(c00, id), i01 = __fmul_rn(MS(a02, a21, a01, a22), id),
                i02 = __fmul_rn(MS(a01, a12, a02, a11), id);
    const float i10 = __fmul_rn(c01, id), i11 = __fmul_rn(MS(a00, a22, a02, a20), id),
                i12 = __fmul_rn(MS(a02, a10, a00, a12), id);
    const float i20 = __fmul_rn(c02, id), i21 = __fmul_rn(MS(a01, a20, a00, a21), id),
                i22 = __fmul_rn(MS(a00, a11, a01, a10), id);
#undef MS
#define D3(x, y, z) (-__fadd_rn(__fadd_rn(__fmul_rn(x, t0), __fmul_rn(y, t1)), __fmul_rn(z, t2)))
    InstanceRec r;
    r.r0 = make_float4(i00, i01, i02, D3(i00, i01, i02));
    r.r1 = make_float4(i10, i11, i12, D3(i10, i11, i12));
    r.r2 = make_float4(i20, i21, i22, D3(i20, i21, i22));
#undef D3
    r.nodes = I.nodes;
    r.tris = I.tris;
    recs[i] = r;
    meta[i] = InstanceMeta{I.custom_index, I.blas_slot};
    // world box: the eight corners of the BLAS root box through the forward transform, padded by a
    // few ulps (culling aid only: the object-space triangle test is the exact part)
    BoxF b = {INFINITY, INFINITY, INFINITY, -INFINITY, -INFINITY, -INFINITY};
    const BoxF s = I.blas_bounds;
    for (int c = 0; c < 8; c++) {
        const float x = (c & 1) ? s.hix : s.lox, y = (c & 2) ? s.hiy : s.loy, z = (c & 4) ? s.hiz : s.loz;
        const float wx = a00 * x + a01 * y + a02 * z + t0;
        const float wy = a10 * x + a11 * y + a12 * z + t1;
        const float wz = a20 * x + a21 * y + a22 * z + t2;
        b.lox = fminf(b.lox, wx);
        b.loy = fminf(b.loy, wy);
        b.loz = fminf(b.loz, wz);
        b.hix = fmaxf(b.hix, wx);
        b.hiy = fmaxf(b.hiy, wy);
        b.hiz = fmaxf(b.hiz, wz);
    }
    const float ex = 4e-6f * fmaxf(fabsf(b.lox), fabsf(b.hix)) + 1e-7f;
    const float ey = 4e-6f * fmaxf(fabsf(b.loy), fabsf(b.hiy)) + 1e-7f;
    const float ez = 4e-6f * fmaxf(fabsf(b.loz), fabsf(b.hiz)) + 1e-7f;
    b.lox -= ex;
    b.hix += ex;
    b.loy -= ey;
    b.hiy += ey;
    b.loz -= ez;
    b.hiz += ez;
    if (!(b.lox <= b.hix && b.loy <= b.hiy && b.loz <= b.hiz)) b = BoxF{0, 0, 0, 0, 0, 0}; // NaN transform
    boxes[i] = b;
}

__global__ void k_instance_gather(const InstanceIn* __restrict__ in, const InstanceRec* __restrict__ rin, const InstanceMeta* __restrict__ min_,
                                  const BoxF* __restrict__ bin, const uint32_t* __restrict__ order, uint32_t n,
                                  InstanceRec* __restrict__ rout, InstanceMeta* __restrict__ mout,
                                  float4* __restrict__ bout, uint32_t* __restrict__ inv) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint32_t src = order[k];
    inv[src] = k; // leaf position of input instance src
    rout[k] = rin[src];
    mout[k] = min_[src];
    const BoxF b = bin[src];
    bout[2 * k] = make_float4(b.lox, b.loy, b.loz, 0.0f);
    bout[2 * k + 1] = make_float4(b.hix, b.hiy, b.hiz, __uint_as_float(in[src].n_tris)); // w: triangle count of the BLAS
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Carver {
    uint8_t* base;
    size_t off = 0;
    template <class T>
    T* take(size_t count) {
        off = align_up(off, 256);
        T* p = reinterpret_cast<T*>(base + off);
        off += count * sizeof(T);
        return p;
    }
};

// trees up to this many primitives are collapsed by the single-CTA kernel
constexpr uint32_t kSmallBuild = 1u << 17;

// SAH constants of the collapse, in units of one wide-node visit.  A leaf slot is tested triangle by
// triangle by few lanes of a warp (profiles/r1_light_pass_v4.md: 3.6 active lanes in the triangle test vs
// 20.6 in the node test), which is why a primitive is priced above the textbook 0.3.  LUZRT_BVH_CPRIM /
// LUZRT_BVH_CPRIM_TLAS override them for tuning runs.
CostParams cost_params(uint32_t max_leaf) {
    static const float c_prim_tri = [] {
        const char* e = getenv("LUZRT_BVH_CPRIM");
        return e ? (float)atof(e) : 0.6f;
    }();
    static const float c_prim_inst = [] {
        const char* e = getenv("LUZRT_BVH_CPRIM_TLAS");
        return e ? (float)atof(e) : 2.0f;
    }();
    return CostParams{1.0f, max_leaf > 1 ? c_prim_tri : c_prim_inst, max_leaf};
}

cudaError_t ensure_capacity(WideBvh& out, size_t nodes, size_t prims) {
    if (out.node_capacity < nodes) {
        if (out.nodes) cudaFree(out.nodes);
        if (out.node_bounds) cudaFree(out.node_bounds);
        out.nodes = nullptr;
        out.node_bounds = nullptr;
        out.node_capacity = 0;
        LUZ_CK(cudaMalloc(&out.nodes, nodes * sizeof(WideNode)));
        LUZ_CK(cudaMalloc(&out.node_bounds, nodes * sizeof(BoxF)));
        out.node_capacity = nodes;
    }
    if (out.prim_capacity < prims) {
        if (out.prim_order) cudaFree(out.prim_order);
        out.prim_order = nullptr;
        out.prim_capacity = 0;
        LUZ_CK(cudaMalloc(&out.prim_order, std::max<size_t>(prims, 1) * sizeof(uint32_t)));
        out.prim_capacity = prims;
    }
    return cudaSuccess;
}

} // namespace

BuildScratch::~BuildScratch() {
    if (mem) cudaFree(mem);
    if (host_pair) cudaFreeHost(host_pair);
    if (host_levels) cudaFreeHost(host_levels);
}

void free_wide_bvh(WideBvh& b) {
    if (b.nodes) cudaFree(b.nodes);
    if (b.node_bounds) cudaFree(b.node_bounds);
    if (b.prim_order) cudaFree(b.prim_order);
    b = WideBvh();
}

cudaError_t build_wide_bvh(cudaStream_t stream, BuildScratch& scratch, const BoxF* d_boxes, uint32_t n,
                           uint32_t max_leaf, WideBvh& out, uint64_t* launches) {
    uint64_t nl = 0;
    out.levels.clear();
    out.n_prims = n;
    if (!scratch.host_pair) LUZ_CK(cudaMallocHost(&scratch.host_pair, 2 * sizeof(uint64_t)));
    if (!scratch.host_levels) LUZ_CK(cudaMallocHost(&scratch.host_levels, sizeof(uint32_t) * (4 + 2 * kMaxLevels)));
    if (n == 0) {
        LUZ_CK(ensure_capacity(out, 1, 1));
        k_empty_root<<<1, 1, 0, stream>>>(out.nodes, out.node_bounds);
        nl++;
        out.n_nodes = 1;
        out.levels.push_back(make_uint2(0, 1));
        if (launches) *launches += nl;
        return cudaGetLastError();
    }
    // worst case: every wide node has two slots -> fewer than n wide nodes (+1 for tiny trees)
    LUZ_CK(ensure_capacity(out, (size_t)n + 1, n));

    const uint32_t sort_blocks = div_up(n, kSortTile);
    size_t need = 0;
    {
        Carver c{nullptr};
        c.take<int>(8);
        c.take<uint32_t>(n);
        c.take<uint32_t>(n);
        c.take<uint32_t>(n);
        c.take<uint32_t>(n);
        c.take<uint32_t>((size_t)256 * sort_blocks);
        c.take<uint64_t>(std::max<size_t>((size_t)256 * sort_blocks, n));
        c.take<uint64_t>(std::max<size_t>((size_t)256 * sort_blocks, n));
        c.take<uint64_t>(2);
        for (int k = 0; k < 4; k++) c.take<uint32_t>(n);
        c.take<uint32_t>(2 * (size_t)n);
        c.take<uint32_t>(n);
        c.take<BoxF>(2 * (size_t)n);
        c.take<uint32_t>(n);
        c.take<uint32_t>(n);
        c.take<uint32_t>(8 * (size_t)n);
        c.take<float>(8 * (size_t)n);
        c.take<uint32_t>(4 + 2 * kMaxLevels);
        need = c.off + 256;
    }
    if (scratch.bytes < need) {
        if (scratch.mem) cudaFree(scratch.mem);
        scratch.mem = nullptr;
        scratch.bytes = 0;
        const size_t grow = need + need / 2;
        LUZ_CK(cudaMalloc(&scratch.mem, grow));
        scratch.bytes = grow;
    }
    Carver c{reinterpret_cast<uint8_t*>(scratch.mem)};
    int* d_bounds = c.take<int>(8);
    uint32_t* codes_a = c.take<uint32_t>(n);
    uint32_t* codes_b = c.take<uint32_t>(n);
    uint32_t* idx_a = c.take<uint32_t>(n);
    uint32_t* idx_b = c.take<uint32_t>(n);
    uint32_t* hist = c.take<uint32_t>((size_t)256 * sort_blocks);
    uint64_t* tmp64_a = c.take<uint64_t>(std::max<size_t>((size_t)256 * sort_blocks, n));
    uint64_t* tmp64_b = c.take<uint64_t>(std::max<size_t>((size_t)256 * sort_blocks, n));
    uint64_t* d_total = c.take<uint64_t>(2);
    uint32_t* left = c.take<uint32_t>(n);
    uint32_t* right = c.take<uint32_t>(n);
    uint32_t* first = c.take<uint32_t>(n);
    uint32_t* last = c.take<uint32_t>(n);
    uint32_t* parent = c.take<uint32_t>(2 * (size_t)n);
    uint32_t* flags = c.take<uint32_t>(n);
    BoxF* box = c.take<BoxF>(2 * (size_t)n);
    uint32_t* items_a = c.take<uint32_t>(n);
    uint32_t* items_b = c.take<uint32_t>(n);
    uint32_t* slots = c.take<uint32_t>(8 * (size_t)n);
    float* cost = c.take<float>(8 * (size_t)n);
    uint32_t* level_table = c.take<uint32_t>(4 + 2 * kMaxLevels);
    const CostParams cp = cost_params(max_leaf);

    const int T = 256;
    k_init_bounds<<<1, 32, 0, stream>>>(d_bounds);
    k_centroid_bounds<<<std::min<uint32_t>(div_up(n, T), 1184u), T, 0, stream>>>(d_boxes, n, d_bounds);
    k_morton<<<div_up(n, T), T, 0, stream>>>(d_boxes, n, d_bounds, codes_a, idx_a);
    nl += 3;

    uint32_t *kin = codes_a, *kout = codes_b, *vin = idx_a, *vout = idx_b;
    if (n > 1) {
        for (int pass = 0; pass < 4; pass++) {
            const int shift = 8 * pass;
            k_sort_hist<<<sort_blocks, kSortThreads, 0, stream>>>(kin, n, shift, sort_blocks, hist);
            const uint32_t hn = 256u * sort_blocks;
            k_u32_to_u64<<<div_up(hn, T), T, 0, stream>>>(hist, tmp64_a, hn);
            k_scan_u64<<<1, 1024, 0, stream>>>(tmp64_a, tmp64_b, hn, nullptr);
            k_sort_scatter<<<sort_blocks, kSortThreads, 0, stream>>>(kin, vin, n, shift, sort_blocks, tmp64_b, kout,
                                                                     vout);
            nl += 4;
            std::swap(kin, kout);
            std::swap(vin, vout);
        }
        k_karras<<<div_up(n - 1, T), T, 0, stream>>>(kin, (int)n, left, right, first, last, parent);
        LUZ_CK(cudaMemsetAsync(flags, 0, sizeof(uint32_t) * n, stream));
        nl += 1;
    }
    k_fit<<<div_up(n, T), T, 0, stream>>>(d_boxes, vin, (int)n, left, right, first, last, parent, flags, box, cost, cp);
    nl += 1;

    Tree2 tree{left, right, first, last, box, (int)n};
    if (n <= kSmallBuild) {
        // one CTA runs every level; a single read-back of the level table at the end (the host needs the
        // node count and the per-level ranges for refits and for the stack-depth check)
        k_collapse_small<<<1, kSmallThreads, 0, stream>>>(tree, cost, cp, vin, out.nodes, out.node_bounds, out.prim_order,
                                                          items_a, items_b, (uint32_t)std::min<size_t>(out.node_capacity, 0xFFFFFFFFu),
                                                          level_table);
        nl += 1;
        LUZ_CK(cudaMemcpyAsync(scratch.host_levels, level_table, sizeof(uint32_t) * (4 + 2 * kMaxLevels),
                               cudaMemcpyDeviceToHost, stream));
        LUZ_CK(cudaStreamSynchronize(stream));
        const uint32_t* lt = scratch.host_levels;
        if (lt[2]) return cudaErrorMemoryAllocation;
        for (uint32_t l = 0; l < lt[0]; l++) out.levels.push_back(make_uint2(lt[4 + 2 * l], lt[5 + 2 * l]));
        out.n_nodes = lt[1];
        if (launches) *launches += nl;
        return cudaGetLastError();
    }
    // level 0 = the root of the binary tree (internal node 0, or leaf 0 when n == 1)
    const uint32_t root_ref = 0;
    LUZ_CK(cudaMemcpyAsync(items_a, &root_ref, sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
    uint32_t *items = items_a, *next_items = items_b;
    uint32_t level_start = 0, level_count = 1, prim_cursor = 0;
    while (level_count) {
        k_collapse_plan<<<div_up(level_count, 128), 128, 0, stream>>>(tree, cost, items, level_count, cp, slots, tmp64_a);
        k_scan_u64<<<1, 1024, 0, stream>>>(tmp64_a, tmp64_b, level_count, d_total);
        const uint32_t next_start = level_start + level_count;
        k_collapse_emit<<<div_up(level_count, 128), 128, 0, stream>>>(tree, items, level_count, slots, tmp64_b, level_start,
                                                                      next_start, prim_cursor, vin, out.nodes,
                                                                      out.node_bounds, out.prim_order, next_items);
        nl += 3;
        LUZ_CK(cudaMemcpyAsync(scratch.host_pair, d_total, sizeof(uint64_t), cudaMemcpyDeviceToHost, stream));
        LUZ_CK(cudaStreamSynchronize(stream));
        const uint64_t tot = scratch.host_pair[0];
        out.levels.push_back(make_uint2(level_start, level_count));
        level_start = next_start;
        level_count = (uint32_t)(tot & 0xFFFFFFFFull);
        prim_cursor += (uint32_t)(tot >> 32);
        std::swap(items, next_items);
        if ((size_t)level_start + level_count > out.node_capacity) return cudaErrorMemoryAllocation;
        if ((size_t)level_start + level_count > kMaxWideNodes) return cudaErrorInvalidValue; // 24-bit child index
    }
    out.n_nodes = level_start;
    if (launches) *launches += nl;
    return cudaGetLastError();
}

cudaError_t refit_wide_bvh(cudaStream_t stream, const BoxF* d_boxes, WideBvh& bvh, uint64_t* launches) {
    for (int l = (int)bvh.levels.size() - 1; l >= 0; l--) {
        const uint2 lv = bvh.levels[l];
        k_refit_level<<<div_up(lv.y, 128), 128, 0, stream>>>(bvh.nodes, bvh.node_bounds, lv.x, lv.y, d_boxes,
                                                             bvh.prim_order);
        if (launches) (*launches)++;
    }
    return cudaGetLastError();
}

// ---- content hash (cross-GPU determinism check) ---------------------------------------------------------------------
// acc += sum_i odd(splitmix64(seed + i)) * (word_i + 1) mod 2^64: position dependent, and -- being a sum of integers --
// independent of the order in which threads add their parts, so equal data gives an equal hash on every run and GPU.
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__global__ void k_hash_words(const uint32_t* __restrict__ w, size_t n, uint64_t seed, size_t stride_words, size_t take_words,
                             unsigned long long* __restrict__ acc) {
    // hashes the first take_words of every stride_words-long record (records with device pointers in their tail)
    unsigned long long h = 0;
    const size_t step = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
        if (i % stride_words >= take_words) continue;
        h += (splitmix64(seed + i) | 1ull) * ((unsigned long long)w[i] + 1ull);
    }
    for (int off = 16; off; off >>= 1) h += __shfl_xor_sync(0xFFFFFFFFu, h, off);
    if ((threadIdx.x & 31) == 0 && h) atomicAdd(acc, h);
}

cudaError_t launch_hash_words(cudaStream_t stream, const void* data, size_t bytes, uint64_t seed, size_t stride_bytes,
                              size_t take_bytes, unsigned long long* d_acc) {
    const size_t n = bytes / 4;
    if (n == 0) return cudaSuccess;
    const uint32_t blocks = (uint32_t)std::min<size_t>((n + 255) / 256, 148 * 8);
    k_hash_words<<<blocks, 256, 0, stream>>>(reinterpret_cast<const uint32_t*>(data), n, seed, stride_bytes / 4, take_bytes / 4, d_acc);
    return cudaGetLastError();
}

cudaError_t launch_triangle_boxes(cudaStream_t stream, const uint8_t* v, uint32_t stride, const uint32_t* idx,
                                  uint32_t n_tris, BoxF* boxes) {
    if (n_tris) k_triangle_boxes<<<div_up(n_tris, 256), 256, 0, stream>>>(v, stride, idx, n_tris, boxes);
    return cudaGetLastError();
}
cudaError_t launch_gather_triangles(cudaStream_t stream, const uint8_t* v, uint32_t stride, const uint32_t* idx,
                                    const uint32_t* order, uint32_t n_tris, WideTri* tris) {
    if (n_tris) k_gather_triangles<<<div_up(n_tris, 256), 256, 0, stream>>>(v, stride, idx, order, n_tris, tris);
    return cudaGetLastError();
}
cudaError_t launch_instance_prepare(cudaStream_t stream, const InstanceIn* in, uint32_t n, BoxF* boxes,
                                    InstanceRec* recs, InstanceMeta* meta) {
    if (n) k_instance_prepare<<<div_up(n, 128), 128, 0, stream>>>(in, n, boxes, recs, meta);
    return cudaGetLastError();
}
cudaError_t launch_instance_gather(cudaStream_t stream, const InstanceIn* in, const InstanceRec* rin, const InstanceMeta* min_,
                                   const BoxF* bin, const uint32_t* order, uint32_t n, InstanceRec* rout,
                                   InstanceMeta* mout, float4* bout, uint32_t* inv) {
    if (n) k_instance_gather<<<div_up(n, 128), 128, 0, stream>>>(in, rin, min_, bin, order, n, rout, mout, bout, inv);
    return cudaGetLastError();
}

} // namespace luz

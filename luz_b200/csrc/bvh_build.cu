// bvh_build.cu -- deterministic GPU builder for the 8-wide compressed BVH.
//
// Replaces the driver-side vkCmdBuildAccelerationStructuresKHR that the reference reaches through
// vkw::CmdBuildBLAS / vkw::CmdBuildTLAS (source/Graphics/VulkanWrapper.cpp:1089-1104, :1106-1139).
// Pipeline: centroid bounds -> 30-bit Morton codes -> stable LSD radix sort (own kernels) ->
// Karras LBVH (index tie-break for duplicate codes) -> bottom-up boxes -> level-synchronous
// greedy collapse to 8-wide nodes with prefix-sum allocation (no allocation atomics, so the
// node/primitive layout is identical on every run and every GPU) -> 80-byte quantised nodes.
#include "bvh.h"

#include <algorithm>
#include <cstdio>

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

__global__ void k_fit(const BoxF* __restrict__ prim_boxes, const uint32_t* __restrict__ sorted_idx, int n,
                      const uint32_t* __restrict__ left, const uint32_t* __restrict__ right,
                      const uint32_t* __restrict__ parent, uint32_t* __restrict__ flags, BoxF* __restrict__ box) {
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
        const BoxF a = load_box_cg(&box[left[cur]]), b = load_box_cg(&box[right[cur]]);
        box[cur] = box_union(a, b);
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
__device__ __forceinline__ float box_area(const BoxF b) {
    const float dx = b.hix - b.lox, dy = b.hiy - b.loy, dz = b.hiz - b.loz;
    const float a = dx * dy + dy * dz + dz * dx;
    return a >= 0.0f ? a : 0.0f; // NaN boxes sort last among expandable slots
}
// expansion priority of a slot: -1 for leaf slots (never expanded), else the surface area
__device__ __forceinline__ float slot_area(const Tree2& t, uint32_t ref, uint32_t max_leaf) {
    return ref_count(t, ref) <= max_leaf ? -1.0f : box_area(t.box[ref]);
}

// plan: per level item, choose up to 8 slots; counts[i] = (#leaf primitives << 32) | #internal children
__global__ void k_collapse_plan(Tree2 t, const uint32_t* __restrict__ items, uint32_t n_items, uint32_t max_leaf,
                                uint32_t* __restrict__ slots, uint64_t* __restrict__ counts) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_items) return;
    const uint32_t r = items[i];
    uint32_t s[8];
    float area[8];
    int ns;
    if (ref_count(t, r) <= max_leaf) { // only the root of a tiny tree
        s[0] = r;
        area[0] = -1.0f;
        ns = 1;
    } else {
        s[0] = t.left[r];
        s[1] = t.right[r];
        ns = 2;
        for (int k = 0; k < 2; k++) area[k] = slot_area(t, s[k], max_leaf);
        while (ns < 8) {
            int best = -1;
            float ba = -1.0f;
            for (int k = 0; k < ns; k++) // largest surface first; leaves carry area -1; first wins ties
                if (area[k] > ba) {
                    ba = area[k];
                    best = k;
                }
            if (best < 0) break;
            const uint32_t e = s[best];
            const uint32_t a = t.left[e], b = t.right[e];
            s[best] = a;
            area[best] = slot_area(t, a, max_leaf);
            s[ns] = b;
            area[ns] = slot_area(t, b, max_leaf);
            ns++;
        }
    }
    uint32_t n_int = 0, n_prim = 0;
    for (int k = 0; k < 8; k++) {
        if (k < ns) {
            slots[i * 8 + k] = s[k];
            if (area[k] >= 0.0f) n_int++;
            else n_prim += ref_count(t, s[k]);
        } else {
            slots[i * 8 + k] = kNone;
        }
    }
    counts[i] = ((uint64_t)n_prim << 32) | n_int;
}

// exponent byte e such that 255 * 2^(e-127) >= ext
__device__ __forceinline__ uint32_t grid_exponent(float ext) {
    const float s = ext / 255.0f;
    const uint32_t bits = __float_as_uint(s);
    uint32_t e = (bits >> 23) & 0xFFu;
    if (bits & 0x7FFFFFu) e++;
    e = max(e, 24u); // keep 2^(e-127) a normal number with headroom
    e = min(e, 254u);
    // guard against the rounding of ext/255
    if (255.0f * __uint_as_float(e << 23) < ext) e = min(e + 1u, 254u);
    return e;
}

__device__ __forceinline__ void quantize_box(const BoxF node, const uint32_t ex, const uint32_t ey, const uint32_t ez,
                                             const BoxF c, uint8_t* q /* lox,loy,loz,hix,hiy,hiz */) {
    const float sx = __uint_as_float(ex << 23), sy = __uint_as_float(ey << 23), sz = __uint_as_float(ez << 23);
    const float lo[3] = {(c.lox - node.lox) / sx, (c.loy - node.loy) / sy, (c.loz - node.loz) / sz};
    const float hi[3] = {(c.hix - node.lox) / sx, (c.hiy - node.loy) / sy, (c.hiz - node.loz) / sz};
    const float sc[3] = {sx, sy, sz};
    const float org[3] = {node.lox, node.loy, node.loz};
    const float clo[3] = {c.lox, c.loy, c.loz}, chi[3] = {c.hix, c.hiy, c.hiz};
#pragma unroll
    for (int k = 0; k < 3; k++) {
        float ql = fminf(fmaxf(floorf(lo[k]), 0.0f), 255.0f);
        float qh = fminf(fmaxf(ceilf(hi[k]), 0.0f), 255.0f);
        // conservative after rounding of the dequantised corner
        if (ql > 0.0f && org[k] + ql * sc[k] > clo[k]) ql -= 1.0f;
        if (qh < 255.0f && org[k] + qh * sc[k] < chi[k]) qh += 1.0f;
        q[k] = (uint8_t)ql;
        q[3 + k] = (uint8_t)qh;
    }
}

__device__ __forceinline__ void write_node(WideNode* dst, const BoxF nb, const uint32_t child_base,
                                           const uint32_t prim_base, const uint8_t* meta, const uint8_t imask,
                                           const BoxF* cb, const bool* used) {
    WideNode w;
    const uint32_t ex = grid_exponent(nb.hix - nb.lox), ey = grid_exponent(nb.hiy - nb.loy),
                   ez = grid_exponent(nb.hiz - nb.loz);
    w.px = nb.lox;
    w.py = nb.loy;
    w.pz = nb.loz;
    w.ex = (uint8_t)ex;
    w.ey = (uint8_t)ey;
    w.ez = (uint8_t)ez;
    w.imask = imask;
    w.child_base = child_base;
    w.prim_base = prim_base;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        w.meta[k] = meta[k];
        if (used[k]) {
            uint8_t q[6];
            quantize_box(nb, ex, ey, ez, cb[k], q);
            w.qlox[k] = q[0];
            w.qloy[k] = q[1];
            w.qloz[k] = q[2];
            w.qhix[k] = q[3];
            w.qhiy[k] = q[4];
            w.qhiz[k] = q[5];
        } else { // inverted box: never hit
            w.qlox[k] = w.qloy[k] = w.qloz[k] = 255;
            w.qhix[k] = w.qhiy[k] = w.qhiz[k] = 0;
        }
    }
    const uint4* src = reinterpret_cast<const uint4*>(&w);
    uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
    for (int k = 0; k < 5; k++) d4[k] = src[k];
}

__global__ void k_collapse_emit(Tree2 t, const uint32_t* __restrict__ items, uint32_t n_items, uint32_t max_leaf,
                                const uint32_t* __restrict__ slots, const uint64_t* __restrict__ scanned,
                                uint32_t level_start, uint32_t next_level_start, uint32_t prim_cursor,
                                const uint32_t* __restrict__ sorted_idx, WideNode* __restrict__ nodes,
                                BoxF* __restrict__ node_bounds, uint32_t* __restrict__ prim_order,
                                uint32_t* __restrict__ next_items) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_items) return;
    const uint32_t r = items[i];
    const uint64_t sc = scanned[i];
    const uint32_t int_off = (uint32_t)(sc & 0xFFFFFFFFull), prim_off = (uint32_t)(sc >> 32);
    const uint32_t child_base = next_level_start + int_off;
    const uint32_t prim_base = prim_cursor + prim_off;
    uint8_t meta[8];
    bool used[8];
    BoxF cb[8];
    uint8_t imask = 0;
    uint32_t n_int = 0, poff = 0;
    for (int k = 0; k < 8; k++) {
        const uint32_t s = slots[i * 8 + k];
        meta[k] = 0;
        used[k] = false;
        if (s == kNone) continue;
        used[k] = true;
        cb[k] = t.box[s];
        const uint32_t cnt = ref_count(t, s);
        if (cnt <= max_leaf) {
            const uint32_t f = ref_first(t, s);
            for (uint32_t q = 0; q < cnt; q++) prim_order[prim_base + poff + q] = sorted_idx[f + q];
            const uint32_t unary = (1u << cnt) - 1u; // 1 -> 001, 2 -> 011, 3 -> 111
            meta[k] = (uint8_t)((unary << 5) | poff);
            poff += cnt;
        } else {
            meta[k] = (uint8_t)((1u << 5) | (24u + (uint32_t)k));
            imask |= (uint8_t)(1u << k);
            next_items[int_off + n_int] = s;
            n_int++;
        }
    }
    const BoxF nb = t.box[r];
    node_bounds[level_start + i] = nb;
    write_node(nodes + level_start + i, nb, child_base, prim_base, meta, imask, cb, used);
}

__global__ void k_empty_root(WideNode* nodes, BoxF* node_bounds) {
    uint8_t meta[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    bool used[8] = {false, false, false, false, false, false, false, false};
    BoxF cb[8];
    BoxF nb = {0, 0, 0, 0, 0, 0};
    node_bounds[0] = nb;
    write_node(nodes, nb, 0, 0, meta, 0, cb, used);
}

// refit one level: every node rebuilds its child boxes from leaf primitive boxes / child node bounds
__global__ void k_refit_level(WideNode* __restrict__ nodes, BoxF* __restrict__ node_bounds, uint32_t level_start,
                              uint32_t n_nodes, const BoxF* __restrict__ prim_boxes,
                              const uint32_t* __restrict__ prim_order) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    WideNode* node = nodes + level_start + i;
    const WideNode w = *node;
    uint8_t meta[8];
    bool used[8];
    BoxF cb[8];
    BoxF nb = {INFINITY, INFINITY, INFINITY, -INFINITY, -INFINITY, -INFINITY};
    bool any = false;
    for (int k = 0; k < 8; k++) {
        meta[k] = w.meta[k];
        used[k] = meta[k] != 0;
        if (!used[k]) continue;
        BoxF b;
        if (w.imask & (1u << k)) {
            const uint32_t rel = __popc((uint32_t)w.imask & ((1u << k) - 1u));
            b = node_bounds[w.child_base + rel];
        } else {
            const uint32_t off = meta[k] & 31u;
            const uint32_t cnt = __popc((uint32_t)(meta[k] >> 5));
            b = prim_boxes[prim_order[w.prim_base + off]];
            for (uint32_t q = 1; q < cnt; q++) b = box_union(b, prim_boxes[prim_order[w.prim_base + off + q]]);
        }
        cb[k] = b;
        nb = any ? box_union(nb, b) : b;
        any = true;
    }
    if (!any) nb = BoxF{0, 0, 0, 0, 0, 0};
    node_bounds[level_start + i] = nb;
    write_node(node, nb, w.child_base, w.prim_base, meta, w.imask, cb, used);
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
    WideTri w;
    w.v0 = make_float4(p0[0], p0[1], p0[2], __uint_as_float(t));
    w.v1 = make_float4(p1[0], p1[1], p1[2], 0.0f);
    w.v2 = make_float4(p2[0], p2[1], p2[2], 0.0f);
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

__global__ void k_instance_gather(const InstanceRec* __restrict__ rin, const InstanceMeta* __restrict__ min_,
                                  const BoxF* __restrict__ bin, const uint32_t* __restrict__ order, uint32_t n,
                                  InstanceRec* __restrict__ rout, InstanceMeta* __restrict__ mout,
                                  float4* __restrict__ bout) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint32_t src = order[k];
    rout[k] = rin[src];
    mout[k] = min_[src];
    const BoxF b = bin[src];
    bout[2 * k] = make_float4(b.lox, b.loy, b.loz, 0.0f);
    bout[2 * k + 1] = make_float4(b.hix, b.hiy, b.hiz, 0.0f);
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
    k_fit<<<div_up(n, T), T, 0, stream>>>(d_boxes, vin, (int)n, left, right, parent, flags, box);
    nl += 1;

    Tree2 tree{left, right, first, last, box, (int)n};
    // level 0 = the root of the binary tree (internal node 0, or leaf 0 when n == 1)
    const uint32_t root_ref = 0;
    LUZ_CK(cudaMemcpyAsync(items_a, &root_ref, sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
    uint32_t *items = items_a, *next_items = items_b;
    uint32_t level_start = 0, level_count = 1, prim_cursor = 0;
    while (level_count) {
        k_collapse_plan<<<div_up(level_count, 128), 128, 0, stream>>>(tree, items, level_count, max_leaf, slots,
                                                                      tmp64_a);
        k_scan_u64<<<1, 1024, 0, stream>>>(tmp64_a, tmp64_b, level_count, d_total);
        const uint32_t next_start = level_start + level_count;
        k_collapse_emit<<<div_up(level_count, 128), 128, 0, stream>>>(tree, items, level_count, max_leaf, slots,
                                                                      tmp64_b, level_start, next_start, prim_cursor,
                                                                      vin, out.nodes, out.node_bounds, out.prim_order,
                                                                      next_items);
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
cudaError_t launch_instance_gather(cudaStream_t stream, const InstanceRec* rin, const InstanceMeta* min_,
                                   const BoxF* bin, const uint32_t* order, uint32_t n, InstanceRec* rout,
                                   InstanceMeta* mout, float4* bout) {
    if (n) k_instance_gather<<<div_up(n, 128), 128, 0, stream>>>(rin, min_, bin, order, n, rout, mout, bout);
    return cudaGetLastError();
}

} // namespace luz

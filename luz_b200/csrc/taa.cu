// taa.cu -- temporal anti-aliasing resolve and the compose (tonemap) epilogue.
//
// Restates source/Shaders/taa.comp:271-308 (main) with its helpers :15-27 (uv, motion vector),
// :29-86 (3x3 neighbourhood + Mitchell reconstruction), :88-119 (closest depth), :121-143
// (clip_aabb), :145-152 (anti_flicker) and utils.glsl:9-15, :91-97; dispatched where
// DeferredRenderer::TAAPass does (DeferredRenderer.cpp:425-445).  Every texture() of the shader
// goes through one LINEAR / REPEAT sampler (VulkanWrapper.cpp:2429-2461): taps that land on texel
// centres are texel loads with wrap-around, the history tap is an fp32 bilinear fetch with wrap.
// Compose restates present.frag:27-35, :85-95 (imageType 0).
#include <cstdlib>

#include "passes.h"

namespace luz {

namespace {

__device__ __forceinline__ int wrapi(int i, int n) {
    int r = i % n;
    return r < 0 ? r + n : r;
}

// REPEAT addressing for coordinates known to lie in [-n, 2n): no integer division
__device__ __forceinline__ int wrap1(int i, int n) { return i < 0 ? i + n : (i >= n ? i - n : i); }

struct Img { // an RGBA32F light image in banded storage order (common.cuh: storage_row)
    const float4* p;
    const FrameConst* fc;
    int w, h;
    __device__ __forceinline__ float4 texel(int x, int y) const { // x, y within one period of the image
        return __ldg(p + (size_t)storage_row(*fc, (uint32_t)wrap1(y, h)) * w + wrap1(x, w));
    }
};

__device__ __forceinline__ float4 bilinear(const Img& im, float u, float v) {
    const float x = u * (float)im.w - 0.5f, y = v * (float)im.h - 0.5f;
    const float fx0 = floorf(x), fy0 = floorf(y);
    const float fx = x - fx0, fy = y - fy0;
    const int x0 = (int)fx0, y0 = (int)fy0;
    const float4 t00 = im.texel(x0, y0), t10 = im.texel(x0 + 1, y0);
    const float4 t01 = im.texel(x0, y0 + 1), t11 = im.texel(x0 + 1, y0 + 1);
    const float4 top = t00 * (1.0f - fx) + t10 * fx;
    const float4 bot = t01 * (1.0f - fx) + t11 * fx;
    return top * (1.0f - fy) + bot * fy;
}

// utils.glsl:9-15 (the reference's non-standard cubic, evaluated literally)
__device__ __forceinline__ float mitchell(float x) {
    const float B = 1.0f / 3.0f, C = 1.0f / 3.0f;
    const float x2 = x * x, x3 = x2 * x;
    return (6.0f - 2.0f * B) * x3 - (6.0f - 2.0f * B - 3.0f * C) * x2 + 1.0f;
}
__device__ __forceinline__ float luminance(float3 c) { return dot3(c, f3(0.2127f, 0.7152f, 0.0722f)); }

// x / c for a constant c with rc = RN(1 / c): Markstein's correction gives the correctly rounded quotient in
// three instructions (the IEEE division sequence is ~9), so results stay bit-identical to `x / c`.
__device__ __forceinline__ float div_const(float x, float c, float rc) {
    const float q = __fmul_rn(x, rc);
    const float r = __fmaf_rn(-c, q, x);
    const float q2 = __fmaf_rn(r, rc, q);
    return (fabsf(q) <= 3.0e38f) ? q2 : q; // inf / NaN pass through as the division would give them
}
__device__ __forceinline__ float4 div_const4(float4 v, float c, float rc) {
    return f4(div_const(v.x, c, rc), div_const(v.y, c, rc), div_const(v.z, c, rc), div_const(v.w, c, rc));
}
__device__ __forceinline__ float4 min3_4(float4 a, float4 b, float4 c) { return min4(a, min4(b, c)); }
__device__ __forceinline__ float4 max3_4(float4 a, float4 b, float4 c) { return max4(a, max4(b, c)); }


struct TapRow { // one row of the 3x3 neighbourhood
    float4 l, c, r;   // lightInput taps
    float4 mn, mx;    // min / max over the three taps (exact in any order)
    float dl, dc, dr; // depth taps
};

// Loads the taps of image row y (any integer: wraps like the REPEAT sampler) for column x and its neighbours.
__device__ __forceinline__ TapRow load_row(const TaaArgs& a, const int xl, const int x, const int xr, const int y) {
    const FrameConst& fc = a.fc;
    const int H = (int)fc.height, W = (int)fc.width;
    const uint32_t yw = (uint32_t)wrap1(y, H);
    const float4* lp = a.light_in + (size_t)storage_row(fc, yw) * W;
    const float* dp = a.depth + (size_t)yw * W;
    TapRow t;
    t.l = __ldg(lp + xl);
    t.c = __ldg(lp + x);
    t.r = __ldg(lp + xr);
    t.dl = __ldg(dp + xl);
    t.dc = __ldg(dp + x);
    t.dr = __ldg(dp + xr);
    t.mn = min3_4(t.l, t.c, t.r);
    t.mx = max3_4(t.l, t.c, t.r);
    return t;
}

// One thread resolves kTaaRows consecutive rows of one column with a sliding three-row window (kTaaRows > 1 saves
// tap loads; measured on B200 it does not pay: the pass executes ~900 non-contractable fp32 operations per
// pixel, i.e. ~17 flop per algorithmic byte, and is bound by fp32 issue, not by HBM -- DESIGN.md section 4).
// Everything whose rounding depends on the order of operations (sums, Mitchell filter, reprojection, clip, blend)
// is evaluated in the reference's order; min/max are order independent and shared per row.
template <int kTaaRows, int kMinBlocks>
__global__ void __launch_bounds__(256, kMinBlocks) k_taa(const TaaArgs a) {
    const FrameConst& fc = a.fc;
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int ry0 = (blockIdx.y * 8 + (threadIdx.x >> 5)) * kTaaRows;
    if (x >= (int)fc.width || ry0 >= (int)a.rows.rows) return;
    const int W = (int)fc.width, H = (int)fc.height;
    const float sw = (float)W, sh = (float)H;
    const Img hist{a.history, &fc, W, H};
    const int xl = wrap1(x - 1, W), xr = wrap1(x + 1, W);
    const int ybase = a.rows.first + (int)(blockIdx.z * a.rows.pitch); // first row of this band
    const int n_rows = min(kTaaRows, (int)a.rows.rows - ry0);

    const float ddx = fabsf(1.0f / sw), ddy = fabsf(1.0f / sh);
    const float wc = mitchell(sqrtf(2.0f)), we = mitchell(1.0f), w0 = mitchell(0.0f);
    float weightSum = 0.0f; // accumulated in the shader's tap order (taa.comp:56-82)
    weightSum += wc; weightSum += we; weightSum += wc; weightSum += we; weightSum += w0;
    weightSum += we; weightSum += wc; weightSum += we; weightSum += wc;
    const float rws = 1.0f / weightSum, r9 = 1.0f / 9.0f, r5 = 1.0f / 5.0f;

    TapRow top = load_row(a, xl, x, xr, ybase + ry0 - 1);
    TapRow mid = load_row(a, xl, x, xr, ybase + ry0);
    for (int k = 0; k < n_rows; k++) {
        const int y = wrap1(ybase + ry0 + k, H);
        const TapRow bot = load_row(a, xl, x, xr, ybase + ry0 + k + 1);

        const float su = ((float)x + 0.5f) / sw, sv = ((float)y + 0.5f) / sh; // get_uv
        // find_closest_3x3: first strict minimum in row-major order
        int bi = -1, bj = -1;
        float dminz = top.dl;
#define LUZ_CLOSEST(z, i, j) \
    if (dminz > (z)) {       \
        bi = (i);            \
        bj = (j);            \
        dminz = (z);         \
    }
        LUZ_CLOSEST(top.dc, 0, -1)
        LUZ_CLOSEST(top.dr, 1, -1)
        LUZ_CLOSEST(mid.dl, -1, 0)
        LUZ_CLOSEST(mid.dc, 0, 0)
        LUZ_CLOSEST(mid.dr, 1, 0)
        LUZ_CLOSEST(bot.dl, -1, 1)
        LUZ_CLOSEST(bot.dc, 0, 1)
        LUZ_CLOSEST(bot.dr, 1, 1)
#undef LUZ_CLOSEST
        const float cu = su + ddx * (float)bi, cv = sv + ddy * (float)bj;
        // get_motion_vector(closest.xy): depth re-fetched at that uv == the minimum itself
        float mvx, mvy;
        {
            const float3 wp = depth_to_world(fc, cu, cv, dminz);
            float4 prevNDC = mat_mul(fc.prev_view_proj, f4(wp.x, wp.y, wp.z, 1.0f));
            float4 curNDC = mat_mul(fc.view_proj, f4(wp.x, wp.y, wp.z, 1.0f));
            prevNDC.x /= prevNDC.w;
            prevNDC.y /= prevNDC.w;
            curNDC.x /= curNDC.w;
            curNDC.y /= curNDC.w;
            mvx = ((curNDC.x - fc.jitter[0]) - (prevNDC.x - fc.prev_jitter[0])) * 0.5f;
            mvy = ((curNDC.y - fc.jitter[1]) - (prevNDC.y - fc.prev_jitter[1])) * 0.5f;
        }
        const float hu = su - mvx, hv = sv - mvy;

        // get_neighbor_3x3
        const float4 &ctl = top.l, &ctc = top.c, &ctr = top.r;
        const float4 &cml = mid.l, &cmc = mid.c, &cmr = mid.r;
        const float4 &cbl = bot.l, &cbc = bot.c, &cbr = bot.r;
        float4 cmin = min3_4(top.mn, mid.mn, bot.mn);
        float4 cmax = max3_4(top.mx, mid.mx, bot.mx);
        float4 cavg = div_const4(ctl + ctc + ctr + cml + cmc + cmr + cbl + cbc + cbr, 9.0f, r9);
        const float4 cmin5 = min3_4(ctc, mid.mn, cbc);
        const float4 cmax5 = max3_4(ctc, mid.mx, cbc);
        const float4 cavg5 = div_const4(ctc + cml + cmc + cmr + cbc, 5.0f, r5);
        cmin = (cmin + cmin5) * 0.5f;
        cmax = (cmax + cmax5) * 0.5f;
        cavg = (cavg + cavg5) * 0.5f;

        float4 sourceSample = f4(0.0f, 0.0f, 0.0f, 0.0f);
        if (a.reconstruct == 1) {
            sourceSample = sourceSample + ctl * wc;
            sourceSample = sourceSample + ctc * we;
            sourceSample = sourceSample + ctr * wc;
            sourceSample = sourceSample + cml * we;
            sourceSample = sourceSample + cmc * w0;
            sourceSample = sourceSample + cmr * we;
            sourceSample = sourceSample + cbl * wc;
            sourceSample = sourceSample + cbc * we;
            sourceSample = sourceSample + cbr * wc;
            sourceSample = div_const4(sourceSample, weightSum, rws);
        }
        if (a.reconstruct == 0 || any_nan4(sourceSample)) sourceSample = cmc;

        float4 result = sourceSample;
        if (!(hu > 1.0f || hv > 1.0f || hu < 0.0f || hv < 0.0f)) {
            // texture(lightHistory, historyUv): the shader fetches it before the bounds test (taa.comp:281) but only
            // uses it past this point; here 0 <= uv <= 1, so the four taps are within one period of the image
            float4 historySample = bilinear(hist, hu, hv);
            { // clip_aabb(cmin.rgb, cmax.rgb, clamp(cavg, cmin, cmax), history)
                const float4 p = f4(clampf(cavg.x, cmin.x, cmax.x), clampf(cavg.y, cmin.y, cmax.y),
                                    clampf(cavg.z, cmin.z, cmax.z), clampf(cavg.w, cmin.w, cmax.w));
                float4 r = historySample - p;
                const float3 rmax = f3(cmax.x - p.x, cmax.y - p.y, cmax.z - p.z);
                const float3 rmin = f3(cmin.x - p.x, cmin.y - p.y, cmin.z - p.z);
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
            { // anti_flicker
                const float3 s3 = f3(sourceSample.x, sourceSample.y, sourceSample.z);
                const float3 h3 = f3(historySample.x, historySample.y, historySample.z);
                const float3 cs = s3 * (1.0f / (fmaxf(fmaxf(s3.x, s3.y), s3.z) + 1.0f));
                const float3 ch = h3 * (1.0f / (fmaxf(fmaxf(h3.x, h3.y), h3.z) + 1.0f));
                sourceWeight *= 1.0f / (1.0f + luminance(cs));
                historyWeight *= 1.0f / (1.0f + luminance(ch));
            }
            const float wsum = fmaxf(sourceWeight + historyWeight, 0.0000001f);
            result = div_const4(sourceSample * sourceWeight + historySample * historyWeight, wsum, 1.0f / wsum);
            if (any_nan4(result)) result = sourceSample;
        }
        a.out[(size_t)storage_row(fc, (uint32_t)y) * W + x] = result;
        top = mid;
        mid = bot;
    }
}

__device__ __forceinline__ unsigned char unorm8(float v) {
    if (!(v > 0.0f)) return 0;
    if (v > 1.0f) v = 1.0f;
    return (unsigned char)(int)floorf(v * 255.0f + 0.5f);
}

__global__ void __launch_bounds__(256) k_compose(const FrameConst fc, const float4* __restrict__ light_in,
                                                 uchar4* __restrict__ out, const BandSet rows) {
    const uint32_t x = blockIdx.x * 32 + (threadIdx.x & 31), r = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= fc.width || r >= rows.rows) return;
    const uint32_t y = band_row(fc, rows, blockIdx.z, r);
    const float4 c = __ldg(light_in + (size_t)storage_row(fc, y) * fc.width + x);
    const float a = 2.51f, b = 0.03f, cc = 2.43f, d = 0.59f, e = 0.14f;
    const float in[3] = {c.x, c.y, c.z};
    float o[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float x = in[k];
        const float m = (x * (a * x + b)) / (x * (cc * x + d) + e);
        o[k] = powf(m, 1.0f / 2.2f);
    }
    out[(size_t)y * fc.width + x] = make_uchar4(unorm8(o[2]), unorm8(o[1]), unorm8(o[0]), 255); // BGRA8, natural rows
}

__global__ void __launch_bounds__(256) k_unpermute(const FrameConst fc, const float4* __restrict__ banded,
                                                   float4* __restrict__ natural, uint32_t y0, uint32_t y1) {
    const uint32_t x = blockIdx.x * 32 + (threadIdx.x & 31), y = y0 + blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= fc.width || y >= y1) return;
    natural[(size_t)(y - y0) * fc.width + x] = __ldg(banded + (size_t)storage_row(fc, y) * fc.width + x);
}

__global__ void __launch_bounds__(512) k_probe_read(const uint4* __restrict__ buf, size_t n16, int iters,
                                                    float* __restrict__ sink) {
    uint4 acc = make_uint4(0, 0, 0, 0);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int it = 0; it < iters; it++) {
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
            const uint4 v = __ldcg(buf + i); // L2 (not L1) is what is being measured
            acc.x ^= v.x;
            acc.y ^= v.y;
            acc.z ^= v.z;
            acc.w ^= v.w;
        }
    }
    if ((acc.x ^ acc.y ^ acc.z ^ acc.w) == 0x12345678u) sink[threadIdx.x] = 1.0f; // keep the loads alive
}

} // namespace

cudaError_t launch_probe_read(cudaStream_t stream, const void* buf, size_t bytes, int iters, float* sink) {
    k_probe_read<<<148 * 4, 512, 0, stream>>>(reinterpret_cast<const uint4*>(buf), bytes / 16, iters, sink);
    return cudaGetLastError();
}

cudaError_t launch_taa_pass(cudaStream_t stream, const TaaArgs& args) {
    if (args.rows.rows == 0 || args.rows.n_bands == 0 || args.fc.width == 0) return cudaSuccess;
    static const int variant = [] {
        const char* e = getenv("LUZRT_TAA_VARIANT");
        return e ? atoi(e) : 0;
    }();
    auto grid = [&](int rows_per_thread) {
        return dim3((args.fc.width + 31) / 32, (args.rows.rows + 8 * rows_per_thread - 1) / (8 * rows_per_thread),
                    args.rows.n_bands);
    };
    if (variant == 1) k_taa<8, 2><<<grid(8), 256, 0, stream>>>(args);
    else if (variant == 2) k_taa<4, 3><<<grid(4), 256, 0, stream>>>(args);
    else if (variant == 3) k_taa<2, 4><<<grid(2), 256, 0, stream>>>(args);
    else if (variant == 4) k_taa<4, 2><<<grid(4), 256, 0, stream>>>(args);
    else if (variant == 5) k_taa<4, 4><<<grid(4), 256, 0, stream>>>(args);
    else k_taa<1, 4><<<grid(1), 256, 0, stream>>>(args); // fastest measured on B200 (0.287 ms at 4K): the pass is fp32-issue bound
    return cudaGetLastError();
}

cudaError_t launch_compose_pass(cudaStream_t stream, const FrameConst& fc, const float4* light_in, uchar4* out_bgra,
                                const BandSet& rows) {
    if (rows.rows == 0 || rows.n_bands == 0 || fc.width == 0) return cudaSuccess;
    const dim3 grid((fc.width + 31) / 32, (rows.rows + 7) / 8, rows.n_bands);
    k_compose<<<grid, 256, 0, stream>>>(fc, light_in, out_bgra, rows);
    return cudaGetLastError();
}

cudaError_t launch_unpermute_rows(cudaStream_t stream, const FrameConst& fc, const float4* banded, float4* natural,
                                  uint32_t y0, uint32_t y1) {
    if (y1 <= y0 || fc.width == 0) return cudaSuccess;
    const dim3 grid((fc.width + 31) / 32, (y1 - y0 + 7) / 8);
    k_unpermute<<<grid, 256, 0, stream>>>(fc, banded, natural, y0, y1);
    return cudaGetLastError();
}

} // namespace luz

// taa_kernel.cuh -- the TAA resolve kernel, instantiated twice: FAST == false in taa.cu (compiled with -fmad=false: every
// operation of taa.comp in the reference's order, IEEE divisions, bit-faithful to the oracle) and FAST == true in
// relaxed.cu (compiled with FMA contraction: divisions by MUFU.RCP, constant divisions by the rounded reciprocal), the
// kernel a host gets unless it sets LUZRT_DEBUG_EXACT_MATH.  The relaxed build stays inside the contract's tolerance
// (radiance max-abs 1e-3 / PSNR 50 dB; measured ~1e-6 relative) and is ~2x faster: the bit-faithful pass is bound by
// fp32 issue (~900 non-contractable operations per pixel), not by HBM.
#pragma once

#include "passes.h"

namespace luz {
namespace {

// utils.glsl:1-7 with the perspective division of the build in use
template <bool FAST>
__device__ __forceinline__ float3 depth_to_world_t(const FrameConst& fc, float u, float v, float depth) {
    if (!FAST) return depth_to_world(fc, u, v, depth);
    const float4 clip = f4(u * 2.0f - 1.0f, v * 2.0f - 1.0f, depth, 1.0f);
    float4 view = mat_mul(fc.inverse_proj, clip);
    view = view * __fdividef(1.0f, view.w);
    const float4 world = mat_mul(fc.inverse_view, view);
    return f3(world.x, world.y, world.z);
}

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
        const uint32_t yw = (uint32_t)wrap1(y, h);
        return __ldg(p + (size_t)(fc->world_shift ? storage_row(*fc, yw) : yw) * w + wrap1(x, w)); // one GPU: natural rows
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
template <bool FAST>
__device__ __forceinline__ float4 div_const4(float4 v, float c, float rc) {
    if (FAST) return v * rc; // relaxed build: the rounded reciprocal (<= 1 ulp off the quotient)
    return f4(div_const(v.x, c, rc), div_const(v.y, c, rc), div_const(v.z, c, rc), div_const(v.w, c, rc));
}
// a / b: the IEEE division of the bit-faithful build, MUFU.RCP + multiply (2 ulp) in the relaxed one
template <bool FAST>
__device__ __forceinline__ float fdiv(float a, float b) {
    return FAST ? __fdividef(a, b) : a / b;
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
    const float4* lp = a.light_in + (size_t)(fc.world_shift ? storage_row(fc, yw) : yw) * W;
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
template <bool FAST, int kTaaRows, int kMinBlocks>
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

    const float ddx = fabsf(fdiv<FAST>(1.0f, sw)), ddy = fabsf(fdiv<FAST>(1.0f, sh));
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

        const float su = fdiv<FAST>((float)x + 0.5f, sw), sv = fdiv<FAST>((float)y + 0.5f, sh); // get_uv
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
            const float3 wp = depth_to_world_t<FAST>(fc, cu, cv, dminz);
            float4 prevNDC = mat_mul(fc.prev_view_proj, f4(wp.x, wp.y, wp.z, 1.0f));
            float4 curNDC = mat_mul(fc.view_proj, f4(wp.x, wp.y, wp.z, 1.0f));
            prevNDC.x = fdiv<FAST>(prevNDC.x, prevNDC.w);
            prevNDC.y = fdiv<FAST>(prevNDC.y, prevNDC.w);
            curNDC.x = fdiv<FAST>(curNDC.x, curNDC.w);
            curNDC.y = fdiv<FAST>(curNDC.y, curNDC.w);
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
        float4 cavg = div_const4<FAST>(ctl + ctc + ctr + cml + cmc + cmr + cbl + cbc + cbr, 9.0f, r9);
        const float4 cmin5 = min3_4(ctc, mid.mn, cbc);
        const float4 cmax5 = max3_4(ctc, mid.mx, cbc);
        const float4 cavg5 = div_const4<FAST>(ctc + cml + cmc + cmr + cbc, 5.0f, r5);
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
            sourceSample = div_const4<FAST>(sourceSample, weightSum, rws);
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
                if (r.x > rmax.x + eps) r = r * fdiv<FAST>(rmax.x, r.x);
                if (r.y > rmax.y + eps) r = r * fdiv<FAST>(rmax.y, r.y);
                if (r.z > rmax.z + eps) r = r * fdiv<FAST>(rmax.z, r.z);
                if (r.x < rmin.x - eps) r = r * fdiv<FAST>(rmin.x, r.x);
                if (r.y < rmin.y - eps) r = r * fdiv<FAST>(rmin.y, r.y);
                if (r.z < rmin.z - eps) r = r * fdiv<FAST>(rmin.z, r.z);
                historySample = p + r;
            }
            float sourceWeight = 0.05f;
            float historyWeight = 1.0f - sourceWeight;
            { // anti_flicker
                const float3 s3 = f3(sourceSample.x, sourceSample.y, sourceSample.z);
                const float3 h3 = f3(historySample.x, historySample.y, historySample.z);
                const float3 cs = s3 * fdiv<FAST>(1.0f, fmaxf(fmaxf(s3.x, s3.y), s3.z) + 1.0f);
                const float3 ch = h3 * fdiv<FAST>(1.0f, fmaxf(fmaxf(h3.x, h3.y), h3.z) + 1.0f);
                sourceWeight *= fdiv<FAST>(1.0f, 1.0f + luminance(cs));
                historyWeight *= fdiv<FAST>(1.0f, 1.0f + luminance(ch));
            }
            const float wsum = fmaxf(sourceWeight + historyWeight, 0.0000001f);
            result = div_const4<FAST>(sourceSample * sourceWeight + historySample * historyWeight, wsum, fdiv<FAST>(1.0f, wsum));
            if (any_nan4(result)) result = sourceSample;
        }
        a.out[(size_t)(fc.world_shift ? storage_row(fc, (uint32_t)y) : (uint32_t)y) * W + x] = result;
        top = mid;
        mid = bot;
    }
}

} // namespace
} // namespace luz

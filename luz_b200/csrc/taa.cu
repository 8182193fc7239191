// taa.cu -- temporal anti-aliasing resolve and the compose (tonemap) epilogue.
//
// Restates source/Shaders/taa.comp:271-308 (main) with its helpers :15-27 (uv, motion vector),
// :29-86 (3x3 neighbourhood + Mitchell reconstruction), :88-119 (closest depth), :121-143
// (clip_aabb), :145-152 (anti_flicker) and utils.glsl:9-15, :91-97; dispatched where
// DeferredRenderer::TAAPass does (DeferredRenderer.cpp:425-445).  Every texture() of the shader
// goes through one LINEAR / REPEAT sampler (VulkanWrapper.cpp:2429-2461): taps that land on texel
// centres are texel loads with wrap-around, the history tap is an fp32 bilinear fetch with wrap.
// Compose restates present.frag:27-35, :85-95 (imageType 0).
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

__global__ void __launch_bounds__(256) k_taa(const TaaArgs a) {
    const FrameConst& fc = a.fc;
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int ry = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= (int)fc.width || ry >= (int)a.rows.rows) return;
    const int y = (int)band_row(fc, a.rows, blockIdx.z, (uint32_t)ry);
    const int W = (int)fc.width, H = (int)fc.height;
    const float sw = (float)W, sh = (float)H;
    const Img light{a.light_in, &fc, W, H}, hist{a.history, &fc, W, H};

    const float su = ((float)x + 0.5f) / sw, sv = ((float)y + 0.5f) / sh; // get_uv
    // find_closest_3x3: first strict minimum in row-major order
    const float ddx = fabsf(1.0f / sw), ddy = fabsf(1.0f / sh);
    int bi = -1, bj = -1;
    float dminz = __ldg(a.depth + (size_t)wrap1(y - 1, H) * W + wrap1(x - 1, W));
#pragma unroll
    for (int j = -1; j <= 1; j++)
#pragma unroll
        for (int i = -1; i <= 1; i++) {
            if (i == -1 && j == -1) continue;
            const float z = __ldg(a.depth + (size_t)wrap1(y + j, H) * W + wrap1(x + i, W));
            if (dminz > z) {
                bi = i;
                bj = j;
                dminz = z;
            }
        }
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
    const float4 ctl = light.texel(x - 1, y - 1), ctc = light.texel(x, y - 1), ctr = light.texel(x + 1, y - 1);
    const float4 cml = light.texel(x - 1, y), cmc = light.texel(x, y), cmr = light.texel(x + 1, y);
    const float4 cbl = light.texel(x - 1, y + 1), cbc = light.texel(x, y + 1), cbr = light.texel(x + 1, y + 1);
    float4 cmin = min4(ctl, min4(ctc, min4(ctr, min4(cml, min4(cmc, min4(cmr, min4(cbl, min4(cbc, cbr))))))));
    float4 cmax = max4(ctl, max4(ctc, max4(ctr, max4(cml, max4(cmc, max4(cmr, max4(cbl, max4(cbc, cbr))))))));
    float4 cavg = (ctl + ctc + ctr + cml + cmc + cmr + cbl + cbc + cbr) / 9.0f;
    const float4 cmin5 = min4(ctc, min4(cml, min4(cmc, min4(cmr, cbc))));
    const float4 cmax5 = max4(ctc, max4(cml, max4(cmc, max4(cmr, cbc))));
    const float4 cavg5 = (ctc + cml + cmc + cmr + cbc) / 5.0f;
    cmin = (cmin + cmin5) * 0.5f;
    cmax = (cmax + cmax5) * 0.5f;
    cavg = (cavg + cavg5) * 0.5f;

    float4 sourceSample = f4(0.0f, 0.0f, 0.0f, 0.0f);
    if (a.reconstruct == 1) {
        const float wc = mitchell(sqrtf(2.0f)), we = mitchell(1.0f), w0 = mitchell(0.0f);
        float weightSum = 0.0f;
        sourceSample = sourceSample + ctl * wc;
        weightSum += wc;
        sourceSample = sourceSample + ctc * we;
        weightSum += we;
        sourceSample = sourceSample + ctr * wc;
        weightSum += wc;
        sourceSample = sourceSample + cml * we;
        weightSum += we;
        sourceSample = sourceSample + cmc * w0;
        weightSum += w0;
        sourceSample = sourceSample + cmr * we;
        weightSum += we;
        sourceSample = sourceSample + cbl * wc;
        weightSum += wc;
        sourceSample = sourceSample + cbc * we;
        weightSum += we;
        sourceSample = sourceSample + cbr * wc;
        weightSum += wc;
        sourceSample = sourceSample / weightSum;
    }
    if (a.reconstruct == 0 || any_nan4(sourceSample)) sourceSample = cmc;

    float4* dst = a.out + (size_t)storage_row(fc, (uint32_t)y) * W + x;
    if (hu > 1.0f || hv > 1.0f || hu < 0.0f || hv < 0.0f) {
        *dst = sourceSample;
        return;
    }
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
    float4 result = (sourceSample * sourceWeight + historySample * historyWeight) /
                    fmaxf(sourceWeight + historyWeight, 0.0000001f);
    if (any_nan4(result)) result = sourceSample;
    *dst = result;
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
    const dim3 grid((args.fc.width + 31) / 32, (args.rows.rows + 7) / 8, args.rows.n_bands);
    k_taa<<<grid, 256, 0, stream>>>(args);
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

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
#include "taa_kernel.cuh"

namespace luz {

namespace {

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

cudaError_t launch_taa_pass(cudaStream_t stream, const TaaArgs& args, bool exact) {
    if (args.rows.rows == 0 || args.rows.n_bands == 0 || args.fc.width == 0) return cudaSuccess;
    if (!exact) return launch_taa_relaxed(stream, args);
    // the bit-faithful build (one row per thread: it is bound by fp32 issue, the sliding window does not pay)
    const dim3 grid((args.fc.width + 31) / 32, (args.rows.rows + 7) / 8, args.rows.n_bands);
    k_taa<false, 1, 4><<<grid, 256, 0, stream>>>(args);
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

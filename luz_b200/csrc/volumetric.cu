// volumetric.cu -- the volumetric light passes that sit between the light pass and TAA (SURVEY section 8f
// rank 4; main.cpp:274-279).
//
// k_volumetric_screen restates source/Shaders/screenSpaceVolumetricLight.comp:22-62, dispatched by
// DeferredRenderer::ScreenSpaceVolumetricLightPass (DeferredRenderer.cpp:294-307): per pixel and per light
// with volumetricType == VOLUMETRIC_TYPE_SCREEN_SPACE, march volumetricSamples steps in screen space from the
// pixel towards the light and add light.color * intensity * absorption for every step whose (bilinear) depth
// is the clear value 1.0.  The pass reads depth anywhere in the frame, so every rank of a partitioned frame
// holds the full depth plane when the scene has such lights (api.cu: need_full_depth).
//
// One thread per pixel, a warp covers 32 x 1 pixels: the march of neighbouring pixels runs along nearly the
// same screen line, so the four depth taps of a step coalesce into a few sectors that stay in L1.
#include "passes.h"

namespace luz {

namespace {

constexpr float kGoldenRatio = 2.118033988749895f; // LuzCommon.h:12 (sic)

__device__ __forceinline__ float fractf(float x) { return x - floorf(x); }

__device__ __forceinline__ int wrap_coord(int i, int n) {
    // one conditional add / subtract is enough: |i| < 2n for uv in [0, 1]
    if (i < 0) i += n;
    if (i >= n) i -= n;
    return i;
}

// texture(depth, uv) through the global LINEAR / REPEAT sampler (VulkanWrapper.cpp:2429-2461), restated as
// nested lerps a + w * (b - a), which return a constant neighbourhood exactly -- like the fixed-point weights of
// a texture unit do -- so that `== 1.0` (:50) means "all four texels are background"
__device__ __forceinline__ float depth_bilinear(const float* __restrict__ d, int w, int h, float u, float v) {
    const float x = u * (float)w - 0.5f, y = v * (float)h - 0.5f;
    const float fx0 = floorf(x), fy0 = floorf(y);
    const float fx = x - fx0, fy = y - fy0;
    const int x0 = wrap_coord((int)fx0, w), y0 = wrap_coord((int)fy0, h);
    const int x1 = wrap_coord((int)fx0 + 1, w), y1 = wrap_coord((int)fy0 + 1, h);
    const float* r0 = d + (size_t)y0 * w;
    const float* r1 = d + (size_t)y1 * w;
    const float t00 = __ldg(r0 + x0), t10 = __ldg(r0 + x1), t01 = __ldg(r1 + x0), t11 = __ldg(r1 + x1);
    const float top = t00 + fx * (t10 - t00);
    const float bot = t01 + fx * (t11 - t01);
    return top + fy * (bot - top);
}

__global__ void __launch_bounds__(128) k_volumetric_screen(const VolumetricArgs a) {
    const FrameConst& fc = a.fc;
    const uint32_t x = blockIdx.x * 32 + (threadIdx.x & 31);
    const uint32_t r = blockIdx.y * 4 + (threadIdx.x >> 5);
    if (x >= fc.width || r >= a.rows.rows) return;
    const uint32_t y = band_row(fc, a.rows, blockIdx.z, r);
    const int W = (int)fc.width, H = (int)fc.height;

    const float pu = (float)x / (float)W, pv = (float)y / (float)H; // pixelPos / imageSize (:27)
    const uchar4 bn8 = __ldg(a.blue_noise + (size_t)(y % fc.bn_h) * fc.bn_w + (x % fc.bn_w));
    const float bn_r = (float)bn8.x / 255.0f;
    float3 radiance = f3(0.0f, 0.0f, 0.0f);

    for (int li = 0; li < a.n_lights; li++) {
        const VolLight L = a.lights[li]; // uniform address: one broadcast load per warp
        if (L.volumetric_type != LUZW_VOLUMETRIC_SCREEN_SPACE) continue;
        const float3 lpos = f3(L.position_type.x, L.position_type.y, L.position_type.z);
        const int type = __float_as_int(L.position_type.w);
        float4 lp = mat_mul(fc.view_proj, f4(lpos.x, lpos.y, lpos.z, 1.0f));
        if (type == LUZW_LIGHT_DIRECTIONAL)
            lp = mat_mul(fc.view_proj, f4(-L.direction_absorption.x * 10000.0f, -L.direction_absorption.y * 10000.0f,
                                          -L.direction_absorption.z * 10000.0f, 1.0f));
        const float lu = (lp.x / lp.w) * 0.5f + 0.5f, lv = (lp.y / lp.w) * 0.5f + 0.5f;
        const int samples = L.samples;
        const float absorption = L.direction_absorption.w / 1000.0f;
        const float inv_n = 1.0f / (float)samples;
        const float du = (pu - lu) * inv_n, dv = (pv - lv) * inv_n;
        const float3 base = f3(L.color_intensity.x, L.color_intensity.y, L.color_intensity.z) * L.color_intensity.w * absorption;
        const float j0 = (fractf(bn_r + kGoldenRatio * (float)fc.frame_mod) * 2.0f - 1.0f) * 0.003f;
        float su = pu + j0, sv = pv + j0;
        for (int i = 0; i < samples; i++) {
            const float ji = (fractf(bn_r + kGoldenRatio * (float)(128 * (i + 1) + fc.frame_mod)) * 2.0f - 1.0f) * 0.003f;
            su -= du + ji;
            sv -= dv + ji;
            const bool inside = su >= 0.0f && su <= 1.0f && sv >= 0.0f && sv <= 1.0f;
            if (!inside) continue;
            const float sd = depth_bilinear(a.depth, W, H, su, sv);
            if (sd != 1.0f) continue;
            float3 sr = base;
            if (type == LUZW_LIGHT_POINT) {
                const float3 wp = depth_to_world(fc, su, sv, sd);
                sr = sr * (5.0f / length3(wp - lpos));
            }
            radiance = radiance + sr;
        }
    }
    float4* px = a.light + (size_t)storage_row(fc, y) * fc.width + x;
    float4 v = *px;
    v.x += radiance.x;
    v.y += radiance.y;
    v.z += radiance.z;
    *px = v;
}

// shadowMapVolumetricLight.comp:42-74, dispatched by DeferredRenderer::ShadowMapVolumetricLightPass
// (DeferredRenderer.cpp:309-322) right after the screen-space pass: per pixel and per light with volumetricType ==
// VOLUMETRIC_TYPE_SHADOW_MAP, 128 steps along the segment camera -> (1.094 x) the pixel's world position, each adding
// (1 - shadow) * color * intensity * 0.000005, shadow from the light's shadow map.  weight / samples / density are
// the shader's hard-coded locals (:57-60), not the LightBlock fields of the same names.
__global__ void __launch_bounds__(128) k_volumetric_shadow_map(const VolumetricArgs a) {
    const FrameConst& fc = a.fc;
    const uint32_t x = blockIdx.x * 32 + (threadIdx.x & 31);
    const uint32_t r = blockIdx.y * 4 + (threadIdx.x >> 5);
    if (x >= fc.width || r >= a.rows.rows) return;
    const uint32_t y = band_row(fc, a.rows, blockIdx.z, r);
    const int W = (int)fc.width, H = (int)fc.height;

    const float pu = (float)x / (float)W, pv = (float)y / (float)H; // :48
    const float pixelDepth = depth_bilinear(a.depth, W, H, pu, pv);  // texture() at a texel corner: a 2x2 average
    const float3 worldPos = depth_to_world(fc, pu, pv, pixelDepth);
    const float3 camPos = f3(fc.cam_pos[0], fc.cam_pos[1], fc.cam_pos[2]);
    const uchar4 bn8 = __ldg(a.blue_noise + (size_t)(y % fc.bn_h) * fc.bn_w + (x % fc.bn_w));
    const float bn_r = (float)bn8.x / 255.0f;
    const float noise0 = fractf(bn_r + kGoldenRatio * (float)fc.frame_mod);
    float3 radiance = f3(0.0f, 0.0f, 0.0f);

    for (int li = 0; li < a.n_lights; li++) {
        const VolLight L = a.lights[li];
        if (L.volumetric_type != LUZW_VOLUMETRIC_SHADOW_MAP) continue;
        const ShadowMapRec& m = a.shadow_maps[L.light_index];
        const float3 lpos = f3(L.position_type.x, L.position_type.y, L.position_type.z);
        const int type = __float_as_int(L.position_type.w);
        const float3 lcol = f3(L.color_intensity.x, L.color_intensity.y, L.color_intensity.z);
        const float weight = 0.000005f, decay = 1.0f, density = 1.094f;
        const int samples = 128;
        const float3 deltaPos = (camPos - worldPos) * density * (1.0f / (float)samples);
        const float off = noise0 * length3(deltaPos);
        float3 samplePos = f3(camPos.x + off, camPos.y + off, camPos.z + off);
        for (int i = 0; i < samples; i++) {
            samplePos = samplePos - deltaPos;
            const float sh = shadow_map_factor(m, type, lpos, samplePos, samplePos);
            radiance = radiance + (1.0f - sh) * lcol * L.color_intensity.w * weight * decay;
        }
    }
    float4* px = a.light + (size_t)storage_row(fc, y) * fc.width + x;
    float4 v = *px;
    v.x += radiance.x;
    v.y += radiance.y;
    v.z += radiance.z;
    *px = v;
}

} // namespace

cudaError_t launch_volumetric_shadow_map(cudaStream_t stream, const VolumetricArgs& args) {
    if (args.rows.rows == 0 || args.rows.n_bands == 0 || args.fc.width == 0 || args.n_lights == 0) return cudaSuccess;
    const dim3 grid((args.fc.width + 31) / 32, (args.rows.rows + 3) / 4, args.rows.n_bands);
    k_volumetric_shadow_map<<<grid, 128, 0, stream>>>(args);
    return cudaGetLastError();
}

cudaError_t launch_volumetric_screen(cudaStream_t stream, const VolumetricArgs& args) {
    if (args.rows.rows == 0 || args.rows.n_bands == 0 || args.fc.width == 0 || args.n_lights == 0) return cudaSuccess;
    const dim3 grid((args.fc.width + 31) / 32, (args.rows.rows + 3) / 4, args.rows.n_bands);
    k_volumetric_screen<<<grid, 128, 0, stream>>>(args);
    return cudaGetLastError();
}

} // namespace luz

// shadow_map.cuh -- per-light shadow maps: the record the kernels read and the two texture() look-ups that
// light.frag:147-165 and shadowMapVolumetricLight.comp:22-40 make into them.
//
// A map is what DeferredRenderer::ShadowMapPass (DeferredRenderer.cpp:268-291) renders with shadowMap.vert /
// .geom / .frag into a D32F image of scene->shadowResolution^2 texels (GPUScene.cpp:317-325):
//   point light       6 layers in cube-face order +X -X +Y -Y +Z -Z, texel = |light.position - fragPos| / zFar
//                     (shadowMap.frag:14-15), cleared to 1.0;
//   spot/directional  1 layer, texel = gl_FragCoord.z under light.viewProj[0] (an orthographic matrix), cleared to 1.0.
// Both look-ups go through the one LINEAR / REPEAT sampler (VulkanWrapper.cpp:2429-2461); bilinear taps are nested
// lerps a + w * (b - a) (a constant neighbourhood is returned exactly).  The 2-D tap wraps; the cube tap selects
// the face by the Vulkan rules (major axis, z over y over x on ties; sc / tc table) and clamps its 2x2 footprint
// to the face -- texels of the adjacent face are not blended in at face borders (documented deviation, it affects
// look-ups within half a texel of a cube edge).
#pragma once

#include "common.cuh"

namespace luz {

struct __align__(16) ShadowMapRec {
    float view_proj[16]; // light.viewProj[0], column-major (spot / directional look-up)
    const float* data;   // layers * res * res floats, nullptr: this light has no map
    uint32_t res;
    uint32_t layers; // 6: cube (point light), 1: 2-D
    float z_far;
    uint32_t pad[3];
};
static_assert(sizeof(ShadowMapRec) == 96, "ShadowMapRec");

__device__ __forceinline__ float shadow_lerp(float a, float b, float w) { return a + w * (b - a); }

__device__ __forceinline__ int shadow_wrap(int i, int n) {
    i %= n;
    return i < 0 ? i + n : i;
}

// texture(textures[light.shadowMap], uv).r, REPEAT addressing
__device__ __forceinline__ float shadow_tap_2d(const float* __restrict__ d, int res, float u, float v) {
    const float x = u * (float)res - 0.5f, y = v * (float)res - 0.5f;
    const float fx0 = floorf(x), fy0 = floorf(y);
    const float fx = x - fx0, fy = y - fy0;
    // uv is unbounded here (a fragment far outside the light's frustum): clamp before the int conversion
    const int ix = (int)fminf(fmaxf(fx0, -1.0e9f), 1.0e9f), iy = (int)fminf(fmaxf(fy0, -1.0e9f), 1.0e9f);
    const int x0 = shadow_wrap(ix, res), y0 = shadow_wrap(iy, res);
    const int x1 = shadow_wrap(ix + 1, res), y1 = shadow_wrap(iy + 1, res);
    const float* r0 = d + (size_t)y0 * res;
    const float* r1 = d + (size_t)y1 * res;
    const float top = shadow_lerp(__ldg(r0 + x0), __ldg(r0 + x1), fx);
    const float bot = shadow_lerp(__ldg(r1 + x0), __ldg(r1 + x1), fx);
    return shadow_lerp(top, bot, fy);
}

// texture(cubeTextures[light.shadowMap], r).r
__device__ __forceinline__ float shadow_tap_cube(const float* __restrict__ d, int res, float3 r) {
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
    const int x0 = min(max((int)fx0, 0), res - 1), y0 = min(max((int)fy0, 0), res - 1);
    const int x1 = min(max((int)fx0 + 1, 0), res - 1), y1 = min(max((int)fy0 + 1, 0), res - 1);
    const float* f = d + (size_t)face * res * res;
    const float* r0 = f + (size_t)y0 * res;
    const float* r1 = f + (size_t)y1 * res;
    const float top = shadow_lerp(__ldg(r0 + x0), __ldg(r0 + x1), fx);
    const float bot = shadow_lerp(__ldg(r1 + x0), __ldg(r1 + x1), fx);
    return shadow_lerp(top, bot, fy);
}

// The shadow-map branch of EvaluateShadow (light.frag:147-165; identical in shadowMapVolumetricLight.comp:22-40
// with samplePos in place of both fragPos and shadowOrigin): 1 = in shadow.
__device__ __forceinline__ float shadow_map_factor(const ShadowMapRec& m, int light_type, float3 light_pos, float3 frag_pos,
                                                   float3 shadow_origin) {
    if (light_type == LUZW_LIGHT_POINT) {
        const float3 lightToFrag = frag_pos - light_pos;
        const float shadowDepth = shadow_tap_cube(m.data, (int)m.res, lightToFrag);
        return (length3(lightToFrag) - 0.05f >= shadowDepth * m.z_far) ? 1.0f : 0.0f;
    }
    const float4 fragInLight = mat_mul(m.view_proj, f4(shadow_origin.x, shadow_origin.y, shadow_origin.z, 1.0f));
    const float shadowDepth = shadow_tap_2d(m.data, (int)m.res, fragInLight.x * 0.5f + 0.5f, fragInLight.y * 0.5f + 0.5f);
    return (fragInLight.z >= shadowDepth) ? 1.0f : 0.0f;
}

} // namespace luz

// shade_kernel.cuh -- k_light_shade: light.frag:171-235 (main) with the BRDF helpers :17-49 and the SHADOW_TYPE_MAP branch
// of EvaluateShadow :147-165, evaluated with the occluded fractions the ray kernels left in the visibility masks.
//
// Instantiated twice.  FAST == false (light_pass.cu, compiled with -fmad=false): every operation of the shader in the
// shader's order, IEEE division and square root -- the build the oracle is compared with value by value.  FAST == true
// (relaxed.cu, compiled with FMA contraction): MUFU reciprocal / rsqrt, pow(x, 5) by multiplication, the kernel a host
// gets unless it sets LUZRT_DEBUG_EXACT_MATH; inside the contract's tolerance (max-abs 1e-3 in linear HDR or PSNR >= 50 dB).
// Both builds hoist what does not depend on the pixel out of the light loop when the lights are staged in shared memory
// (normalize(-direction), inner - outer, color * intensity: the same operations on the same operands, done once per
// CTA instead of once per pixel), and take albedo^2.2 from a 256-entry table of powf() values (an RGBA8 channel has
// 256 values).
#pragma once

#include "passes.h"

namespace luz {
namespace {

constexpr float kShadePI = 3.14159265359f; // LuzCommon.h:11
constexpr int kShadeChunk = 256;

template <bool FAST>
__device__ __forceinline__ float sdiv(float a, float b) {
    return FAST ? __fdividef(a, b) : a / b;
}
template <bool FAST>
__device__ __forceinline__ float3 snormalize(float3 a) {
    if (FAST) {
        const float s = rsqrtf(dot3(a, a));
        return a * s;
    }
    return normalize3(a);
}
template <bool FAST>
__device__ __forceinline__ float slength(float3 a) {
    if (FAST) {
        float r;
        asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(dot3(a, a)));
        return r;
    }
    return length3(a);
}

// light.frag:17-26
template <bool FAST>
__device__ __forceinline__ float distribution_ggx(float3 N, float3 H, float roughness) {
    const float a = roughness * roughness;
    const float a2 = a * a;
    const float NdotH = fmaxf(dot3(N, H), 0.0f);
    const float NdotH2 = NdotH * NdotH;
    float denom = (NdotH2 * (a2 - 1.0f) + 1.0f);
    denom = kShadePI * denom * denom;
    return sdiv<FAST>(a2, denom);
}
// light.frag:28-36
template <bool FAST>
__device__ __forceinline__ float geometry_schlick_ggx(float NdotV, float roughness) {
    const float r = roughness + 1.0f;
    const float k = (r * r) / 8.0f;
    return sdiv<FAST>(NdotV, NdotV * (1.0f - k) + k);
}

// number of set bits among bits [b0, b0 + n) of a pixel's mask words
__device__ __forceinline__ uint32_t count_bits(const uint32_t* __restrict__ m, uint32_t b0, int n) {
    uint32_t c = 0;
    while (n > 0) {
        const uint32_t w = __ldg(m + (b0 >> 5)), s = b0 & 31u;
        const uint32_t take = min((uint32_t)n, 32u - s);
        const uint32_t sel = take == 32u ? 0xFFFFFFFFu : ((1u << take) - 1u);
        c += __popc((w >> s) & sel);
        b0 += take;
        n -= (int)take;
    }
    return c;
}

// one light as the shading loop wants it (64 bytes in shared memory): the pixel-independent parts of light.frag:193-210
struct __align__(16) ShadeLight {
    float4 pos_type;     // light.position, type (int bits)
    float4 ndir_outer;   // normalize(-light.direction), outerAngle
    float4 col_eps;      // light.color * light.intensity, innerAngle - outerAngle
    int num_shadow_samples;
    int shadow_map;
    int pad[2];
};
static_assert(sizeof(ShadeLight) == 64, "ShadeLight");

// SMAP: compiled with the shadow-map branch of EvaluateShadow (light.frag:147-165).
template <bool SMAP, bool FAST>
__global__ void __launch_bounds__(128) k_light_shade(const LightArgs a) {
    __shared__ ShadeLight s_lights[kShadeChunk];
    __shared__ float s_pow22[256]; // (c / 255)^2.2 for the 256 values of an RGBA8 channel (light.frag:172)
    const FrameConst& fc = a.fc;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t x = blockIdx.x * 32 + lane;
    const uint32_t r = blockIdx.y * 4 + warp;
    const bool in_image = x < fc.width && r < a.rows.rows;
    const uint32_t y = in_image ? band_row(fc, a.rows, blockIdx.z, r) : 0u;
    const size_t pix = (size_t)y * fc.width + x;                   // G-buffer, masks: natural row order
    const size_t opix = (size_t)storage_row(fc, y) * fc.width + x; // light image: banded storage order
    for (int k = threadIdx.x; k < 256; k += blockDim.x) s_pow22[k] = __ldg(a.pow22 + k);

    // ---- G-buffer fetch (light.frag:172-176; texel loads) ----
    float3 N = f3(0.0f, 0.0f, 0.0f);
    uchar4 a8 = make_uchar4(0, 0, 0, 0), m8 = a8, e8 = a8;
    float depth = 1.0f;
    if (in_image) {
        const float4 n4 = __ldg(a.normal + pix);
        N = f3(n4.x, n4.y, n4.z);
        a8 = __ldg(a.albedo + pix);
        m8 = __ldg(a.material + pix);
        e8 = __ldg(a.emission + pix);
        depth = __ldg(a.depth + pix);
    }
    const float3 ambientLight = f3(fc.ambient[0], fc.ambient[1], fc.ambient[2]);
    const bool lit = in_image && (length3(N) != 0.0f); // :178
    if (in_image && !lit) a.out[opix] = make_float4(ambientLight.x, ambientLight.y, ambientLight.z, 1.0f);
    { // lit pixels of the rows this rank owns (halo rows are recomputation): the frame's ray count follows from it,
      // whichever ray kernels ran (shadow rays only, AO rays only, both, none)
        const bool counted = r >= a.count_row_begin && r < a.count_row_end;
        const unsigned int lit_warp = __popc(__ballot_sync(0xFFFFFFFFu, lit && counted));
        if (lane == 0 && lit_warp)
            atomicAdd(a.lit_counters + 16 * ((blockIdx.x + (blockIdx.y * 4u + warp + blockIdx.z * 7u) * 29u) & 63u),
                      (unsigned long long)lit_warp);
    }
    __syncthreads();
    const float3 albedo = f3(s_pow22[a8.x], s_pow22[a8.y], s_pow22[a8.z]);
    const float roughness = (float)m8.x / 255.0f, metallic = (float)m8.y / 255.0f, occlusion = (float)m8.z / 255.0f;
    const float u = ((float)x + 0.5f) / (float)fc.width, v = ((float)y + 0.5f) / (float)fc.height;
    float3 fragPos;
    if (FAST) {
        const float4 clip = f4(u * 2.0f - 1.0f, v * 2.0f - 1.0f, depth, 1.0f);
        float4 view = mat_mul(fc.inverse_proj, clip);
        view = view * __fdividef(1.0f, view.w);
        const float4 world = mat_mul(fc.inverse_view, view);
        fragPos = f3(world.x, world.y, world.z);
    } else {
        fragPos = depth_to_world(fc, u, v, depth);
    }
    const float3 camPos = f3(fc.cam_pos[0], fc.cam_pos[1], fc.cam_pos[2]);
    const float3 V = snormalize<FAST>(camPos - fragPos);
    const float3 F0 = f3(0.04f, 0.04f, 0.04f) * (1.0f - metallic) + albedo * metallic;
    const float camDist = length3(fragPos - camPos);
    const float NdotV = fmaxf(dot3(N, V), 0.0f);
    const float ggxV = geometry_schlick_ggx<FAST>(NdotV, roughness);
    const uint32_t* smask = a.shadow_mask + pix * a.shadow_words;
    const uint32_t* amask = a.ao_mask + pix * a.ao_words;

    float3 Lo = f3(0.0f, 0.0f, 0.0f);
    uint32_t shadow_bit = 0;
    for (int base = 0; base < fc.num_lights; base += kShadeChunk) {
        const int chunk = min(kShadeChunk, fc.num_lights - base);
        __syncthreads();
        for (int k = threadIdx.x; k < chunk; k += blockDim.x) { // stage + hoist: once per CTA, the shader's own operations
            const LightRec L4 = a.lights[base + k];
            const float3 ldir = f3(L4.direction_outer.x, L4.direction_outer.y, L4.direction_outer.z);
            const float3 nd = normalize3(-ldir); // normalize(-light.direction) :198, :203
            ShadeLight sl;
            sl.pos_type = f4(L4.position_inner.x, L4.position_inner.y, L4.position_inner.z, __int_as_float(L4.type));
            sl.ndir_outer = f4(nd.x, nd.y, nd.z, L4.direction_outer.w);
            sl.col_eps = f4(L4.color_intensity.x * L4.color_intensity.w, L4.color_intensity.y * L4.color_intensity.w,
                            L4.color_intensity.z * L4.color_intensity.w, L4.position_inner.w - L4.direction_outer.w);
            sl.num_shadow_samples = L4.num_shadow_samples;
            sl.shadow_map = L4.shadow_map;
            sl.pad[0] = sl.pad[1] = 0;
            s_lights[k] = sl;
        }
        __syncthreads();
        if (!lit) continue;
        for (int li = 0; li < chunk; li++) {
            const ShadeLight& sl = s_lights[li];
            const float4 pt = sl.pos_type, nd = sl.ndir_outer, ce = sl.col_eps;
            const int type = __float_as_int(pt.w);
            const float3 lpos = f3(pt.x, pt.y, pt.z);
            const float3 Lvec = lpos - fragPos;
            float3 L;
            float attenuation = 1.0f;
            if (FAST) {
                const float d2 = dot3(Lvec, Lvec);
                const float inv = rsqrtf(d2);
                L = Lvec * inv; // normalize(L_)
                if (type == LUZW_LIGHT_DIRECTIONAL) {
                    L = f3(nd.x, nd.y, nd.z);
                } else if (type == LUZW_LIGHT_SPOT) {
                    attenuation = inv * inv; // 1 / (dist * dist)
                    const float theta = dot3(L, f3(nd.x, nd.y, nd.z));
                    attenuation *= clampf(__fdividef(theta - nd.w, ce.w), 0.0f, 1.0f);
                } else if (type == LUZW_LIGHT_POINT) {
                    attenuation = inv * inv;
                }
            } else {
                const float dist = length3(Lvec);
                L = Lvec / dist; // normalize(L_)
                if (type == LUZW_LIGHT_DIRECTIONAL) {
                    L = f3(nd.x, nd.y, nd.z);
                } else if (type == LUZW_LIGHT_SPOT) {
                    attenuation = 1.0f / (dist * dist);
                    const float theta = dot3(L, f3(nd.x, nd.y, nd.z));
                    attenuation *= clampf((theta - nd.w) / ce.w, 0.0f, 1.0f);
                } else if (type == LUZW_LIGHT_POINT) {
                    attenuation = 1.0f / (dist * dist);
                }
            }
            // shadow factor: RT with samples -> occluded fraction; RT with 0 samples -> 0; otherwise 1 (:166-168)
            const int n_samples = fc.shadow_type == LUZW_SHADOW_RAYTRACING ? sl.num_shadow_samples : 0;
            float shadowFactor = fc.shadow_type != LUZW_SHADOW_RAYTRACING ? 1.0f : 0.0f;
            if (n_samples > 0) {
                shadowFactor = sdiv<FAST>((float)count_bits(smask, shadow_bit, n_samples), (float)n_samples);
                shadow_bit += (uint32_t)n_samples;
            }
            if (SMAP && fc.shadow_type == LUZW_SHADOW_MAP && sl.shadow_map != -1) { // light.frag:147-165
                const float3 O = fragPos + N * fmaxf(camDist * 0.01f, 0.05f);
                shadowFactor = shadow_map_factor(a.shadow_maps[base + li], type, lpos, fragPos, O);
            }
            // light.color * light.intensity * attenuation * (1.0 - shadowFactor), left to right
            const float3 radiance = f3(ce.x, ce.y, ce.z) * attenuation * (1.0f - shadowFactor);

            const float3 H = snormalize<FAST>(V + L);
            const float NDF = distribution_ggx<FAST>(N, H, roughness);
            const float NdotL = fmaxf(dot3(N, L), 0.0f);
            const float G = geometry_schlick_ggx<FAST>(NdotL, roughness) * ggxV; // GeometrySmith :38-45
            const float c1 = clampf(1.0f - clampf(dot3(H, V), 0.0f, 1.0f), 0.0f, 1.0f);
            float fp;
            if (FAST) {
                const float c2 = c1 * c1;
                fp = c2 * c2 * c1;
            } else {
                fp = powf(c1, 5.0f);
            }
            const float3 F = F0 + (f3(1.0f, 1.0f, 1.0f) - F0) * fp; // FresnelSchlick :47-49
            const float3 num = NDF * G * F;
            const float denom = 4.0f * NdotV * NdotL + 0.0001f;
            float3 kD = f3(1.0f, 1.0f, 1.0f) - F;
            kD = kD * (1.0f - metallic);
            if (FAST) {
                const float rden = __fdividef(1.0f, denom);
                Lo = Lo + (kD * albedo * (1.0f / kShadePI) + num * rden) * radiance * NdotL;
            } else {
                const float3 spec = num / denom;
                Lo = Lo + (kD * albedo / kShadePI + spec) * radiance * NdotL;
            }
        }
    }
    if (lit) {
        float rayTracedAo = 1.0f;
        const int n_ao = fc.ao_num_samples;
        if (n_ao != 0) rayTracedAo = sdiv<FAST>((float)n_ao - (float)count_bits(amask, 0u, n_ao), (float)n_ao); // ao / aoNumSamples
        const float3 emission = f3((float)e8.x / 255.0f, (float)e8.y / 255.0f, (float)e8.z / 255.0f);
        const float3 ambient = ambientLight * albedo * occlusion * rayTracedAo;
        const float3 color = ambient + Lo + emission;
        a.out[opix] = make_float4(color.x, color.y, color.z, 1.0f);
    }
}

} // namespace
} // namespace luz

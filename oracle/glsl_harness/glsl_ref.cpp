// glsl_ref.cpp -- runs the reference's OWN light.frag and taa.comp (generated *.inc, see preprocess.py) on the host
// (TEST INFRASTRUCTURE; built into oracle/_ref/libglsl_ref.so by oracle/Makefile, only where /root/reference exists).
// tests/test_glsl_pin.py compares oracle/luz_oracle.cpp with it; tests/golden/make_glsl_golden.py stores its outputs as
// fixtures for machines without the reference.
#include "glsl_prelude.hpp"

#include "../luz_oracle.h"

namespace light_frag {
using namespace glsl;
#include "light_frag.inc"
} // namespace light_frag

// LuzCommon.h's resource macros would otherwise rewrite struct members of the second copy (`int vertexBuffer;`)
#undef scene
#undef tlas
#undef model
#undef vertexBuffer
#undef lineBlocks

namespace taa_comp {
using namespace glsl;
#include "taa_comp.inc"
} // namespace taa_comp

static_assert(sizeof(light_frag::SceneBlock) == 31200 && sizeof(light_frag::LightBlock) == 480, "std430 layout == C++ layout");

namespace {
struct TraceCtx {
    const orc_world* world;
    int exhaustive;
};
bool trace_cb(void* user, const float* o, const float* d, float tmin, float tmax) {
    const TraceCtx* c = (const TraceCtx*)user;
    uint8_t hit = 0;
    orc_trace_any(c->world, 1, o, d, &tmin, &tmax, c->exhaustive, &hit);
    return hit != 0;
}
} // namespace

extern "C" {

// light.frag main() for the listed pixels (x, y pairs).  out: n * 4 floats.  shadow_mask / ao_mask: n * words u32, ray i of
// the shadow loops (lights in order) / of the AO loop occluded = bit i (may be NULL).  Returns 0, or -1 for > 64 lights.
int glsl_light_frag(const void* scene_block, uint32_t width, uint32_t height, const orc_gbuffer* gb, uint32_t frame,
                    const uint8_t* blue_noise_rgba8, uint32_t bn_w, uint32_t bn_h, const orc_world* world, int exhaustive,
                    const uint32_t* pixels_xy, uint32_t n_pixels, float* out_rgba, uint32_t* shadow_mask,
                    uint32_t shadow_words, uint32_t* ao_mask, uint32_t ao_words) {
    using namespace light_frag;
    std::memcpy(&sceneBuffers[0].block, scene_block, sizeof(SceneBlock));
    SceneBlock& s = sceneBuffers[0].block;
    if (s.numLights > LUZ_MAX_LIGHTS) return -1;
    s.blueNoiseTexture = 5;
    s.tlasRid = 0;
    const int W = (int)width, H = (int)height;
    textures[0] = sampler2D{gb->albedo, W, H, TEX_RGBA8_UNORM};
    textures[1] = sampler2D{gb->normal, W, H, TEX_RGBA32F};
    textures[2] = sampler2D{gb->material, W, H, TEX_RGBA8_UNORM};
    textures[3] = sampler2D{gb->emission, W, H, TEX_RGBA8_UNORM};
    textures[4] = sampler2D{gb->depth, W, H, TEX_R32F};
    textures[5] = sampler2D{blue_noise_rgba8, (int)bn_w, (int)bn_h, TEX_RGBA8_UNORM};
    ctx = LightConstants{0, 0, (int)frame, 0, 1, 2, 3, 4};
    TraceCtx tc{world, exhaustive};
    uint32_t shadow_rays = 0; // rays the shadow loops fire per lit pixel (light.frag:86-108, :141-146)
    if (s.shadowType == SHADOW_TYPE_RAYTRACING)
        for (int i = 0; i < s.numLights; i++) shadow_rays += (uint32_t)(s.lights[i].numShadowSamples > 0 ? s.lights[i].numShadowSamples : 0);
    for (uint32_t k = 0; k < n_pixels; k++) {
        const uint32_t x = pixels_xy[2 * k], y = pixels_xy[2 * k + 1];
        gl_FragCoord = vec4((float)x + 0.5f, (float)y + 0.5f, 0.0f, 1.0f);
        // light.vert:16 + the viewport (VulkanWrapper.cpp:1216-1222): the interpolated uv of the full-screen triangle
        fragTexCoord = vec2(((float)x + 0.5f) / (float)width, ((float)y + 0.5f) / (float)height);
        g_rays = RayHook{};
        g_rays.fn = trace_cb;
        g_rays.user = &tc;
        outColor = vec4(0.0f);
        shader_main();
        out_rgba[4 * k + 0] = outColor.x, out_rgba[4 * k + 1] = outColor.y, out_rgba[4 * k + 2] = outColor.z, out_rgba[4 * k + 3] = outColor.w;
        for (uint32_t w = 0; shadow_mask && w < shadow_words; w++) shadow_mask[(size_t)k * shadow_words + w] = 0;
        for (uint32_t w = 0; ao_mask && w < ao_words; w++) ao_mask[(size_t)k * ao_words + w] = 0;
        for (uint32_t r = 0; r < g_rays.n_rays && r < 512; r++) {
            if (!((g_rays.hit_bits[r >> 6] >> (r & 63)) & 1ull)) continue;
            if (r < shadow_rays) {
                if (shadow_mask && (r >> 5) < shadow_words) shadow_mask[(size_t)k * shadow_words + (r >> 5)] |= 1u << (r & 31);
            } else {
                const uint32_t a = r - shadow_rays;
                if (ao_mask && (a >> 5) < ao_words) ao_mask[(size_t)k * ao_words + (a >> 5)] |= 1u << (a & 31);
            }
        }
    }
    return 0;
}

// taa.comp main() for the listed invocations (x, y pairs; invocations outside the image return like :274-276).
// out_rgba: the full-frame RGBA32F image imageStore writes to.
int glsl_taa_comp(const void* scene_block, uint32_t width, uint32_t height, const float* light_in, const float* history,
                  const float* depth, int reconstruct, const uint32_t* pixels_xy, uint32_t n_pixels, float* out_rgba) {
    using namespace taa_comp;
    std::memcpy(&sceneBuffers[0].block, scene_block, sizeof(SceneBlock));
    const int W = (int)width, H = (int)height;
    textures[0] = sampler2D{light_in, W, H, TEX_RGBA32F};
    textures[1] = sampler2D{history, W, H, TEX_RGBA32F};
    textures[2] = sampler2D{depth, W, H, TEX_R32F};
    images[0] = image2D{out_rgba, W, H};
    ctx = PostProcessingConstants{};
    ctx.lightInputRID = 0, ctx.lightOutputRID = 0, ctx.lightHistoryRID = 1, ctx.depthRID = 2;
    ctx.size = vec2((float)width, (float)height);
    ctx.sceneBufferIndex = 0;
    ctx.reconstruct = reconstruct;
    for (uint32_t k = 0; k < n_pixels; k++) {
        gl_GlobalInvocationID = uvec3(pixels_xy[2 * k], pixels_xy[2 * k + 1], 0u);
        shader_main();
    }
    return 0;
}

} // extern "C"

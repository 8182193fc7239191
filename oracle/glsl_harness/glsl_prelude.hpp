// glsl_prelude.hpp -- the GLSL environment the reference's shader text is compiled in (TEST INFRASTRUCTURE).
//
// Types and built-in functions come from the reference's own glm (deps/glm, function swizzles); this header only adds
// what glm has no counterpart for: the implicit int -> float vector conversions GLSL performs, and the fixed-function
// units the shaders call into -- the texture unit (texture / texelFetch / textureSize / imageStore) and the ray query.
// Those are driver territory, not shader arithmetic: the texture unit is implemented here from the
// sampler the reference creates (LINEAR, REPEAT, LOD 0; VulkanWrapper.cpp:2429-2461), the ray query is routed to a
// callback (the oracle's exhaustive tracer in the pinning tests).
#pragma once
#define GLM_FORCE_SWIZZLE
#define GLM_FORCE_DEPTH_ZERO_TO_ONE
#include <glm.hpp>

#include <cmath>
#include <cstdint>
#include <cstring>

namespace glsl {
using namespace glm;
using uint = unsigned int;

#define GLSL_GLOBAL static thread_local

// ---- texture unit ------------------------------------------------------------------------------------------------
enum TexFormat { TEX_NONE = 0, TEX_RGBA8_UNORM, TEX_RGBA32F, TEX_R32F };
struct sampler2D {
    const void* data = nullptr;
    int w = 0, h = 0;
    TexFormat format = TEX_NONE;
};
struct samplerCube {
    int unused = 0;
};
struct image2D {
    float* data = nullptr; // RGBA32F
    int w = 0, h = 0;
};
inline int wrap_repeat(int i, int n) {
    int r = i % n;
    return r < 0 ? r + n : r;
}
inline vec4 fetch_texel(const sampler2D& s, int x, int y) {
    const size_t i = (size_t)wrap_repeat(y, s.h) * s.w + wrap_repeat(x, s.w);
    switch (s.format) {
        case TEX_RGBA8_UNORM: {
            const uint8_t* p = (const uint8_t*)s.data + 4 * i;
            return vec4((float)p[0] / 255.0f, (float)p[1] / 255.0f, (float)p[2] / 255.0f, (float)p[3] / 255.0f);
        }
        case TEX_RGBA32F: {
            const float* p = (const float*)s.data + 4 * i;
            return vec4(p[0], p[1], p[2], p[3]);
        }
        case TEX_R32F: {
            const float v = ((const float*)s.data)[i];
            return vec4(v, 0.0f, 0.0f, 1.0f);
        }
        default: return vec4(0.0f);
    }
}
inline ivec2 textureSize(const sampler2D& s, int) { return ivec2(s.w, s.h); }
inline vec4 texelFetch(const sampler2D& s, ivec2 p, int) { return fetch_texel(s, p.x, p.y); }
// texture(): LINEAR mag/min filter, REPEAT addressing, LOD 0 (the one global sampler).  The unnormalised coordinate is
// taken to the texel grid with 8 sub-texel bits (VkPhysicalDeviceLimits::subTexelPrecisionBits, the value every desktop
// part reports): a tap that lands on a texel centre at that precision IS that texel (this is what makes the G-buffer
// fetches at uv = (p + 0.5) / size texel loads); any other tap interpolates its four texels with full fp32 weights.
inline vec4 texture(const sampler2D& s, vec2 uv) {
    const float x = uv.x * (float)s.w - 0.5f, y = uv.y * (float)s.h - 0.5f;
    const float rx = std::floor(x + 0.5f), ry = std::floor(y + 0.5f);
    const bool on_x = std::fabs(x - rx) <= 1.0f / 512.0f, on_y = std::fabs(y - ry) <= 1.0f / 512.0f;
    if (on_x && on_y) return fetch_texel(s, (int)rx, (int)ry);
    const float fx0 = std::floor(x), fy0 = std::floor(y);
    const float fx = x - fx0, fy = y - fy0;
    const int x0 = (int)fx0, y0 = (int)fy0;
    const vec4 top = fetch_texel(s, x0, y0) * (1.0f - fx) + fetch_texel(s, x0 + 1, y0) * fx;
    const vec4 bot = fetch_texel(s, x0, y0 + 1) * (1.0f - fx) + fetch_texel(s, x0 + 1, y0 + 1) * fx;
    return top * (1.0f - fy) + bot * fy;
}
inline vec4 texture(const samplerCube&, vec3) { return vec4(1.0f); } // shadow-map path: not exercised by the harness
inline void imageStore(image2D& im, ivec2 p, vec4 v) {
    if (p.x < 0 || p.y < 0 || p.x >= im.w || p.y >= im.h) return;
    float* q = im.data + 4 * ((size_t)p.y * im.w + p.x);
    q[0] = v.x, q[1] = v.y, q[2] = v.z, q[3] = v.w;
}

// ---- ray query (GL_EXT_ray_query): any-hit, terminate on first hit -----------------------------------------------------
struct accelerationStructureEXT {
    int unused = 0;
};
typedef bool (*TraceFn)(void* user, const float* origin3, const float* dir3, float tmin, float tmax);
struct RayHook {
    TraceFn fn = nullptr;
    void* user = nullptr;
    uint32_t n_rays = 0;       // rays fired by the current invocation, in call order
    uint64_t hit_bits[8] = {}; // bit i set = ray i occluded (up to 512 rays per pixel)
};
GLSL_GLOBAL RayHook g_rays;
struct rayQueryEXT {
    bool hit = false;
};
constexpr uint gl_RayFlagsTerminateOnFirstHitEXT = 4u;
constexpr uint gl_RayQueryCommittedIntersectionNoneEXT = 0u;
inline void rayQueryInitializeEXT(rayQueryEXT& rq, const accelerationStructureEXT&, uint, uint, vec3 o, float tmin, vec3 d, float tmax) {
    const float oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z};
    rq.hit = g_rays.fn ? g_rays.fn(g_rays.user, oo, dd, tmin, tmax) : false;
    if (rq.hit && g_rays.n_rays < 512) g_rays.hit_bits[g_rays.n_rays >> 6] |= 1ull << (g_rays.n_rays & 63);
    g_rays.n_rays++;
}
inline bool rayQueryProceedEXT(rayQueryEXT&) { return false; }
inline uint rayQueryGetIntersectionTypeEXT(const rayQueryEXT& rq, bool) { return rq.hit ? 1u : 0u; }

// ---- GLSL's implicit conversions where glm wants identical types -----------------------------------------------------
inline bvec2 greaterThanEqual(ivec2 a, vec2 b) { return glm::greaterThanEqual(vec2(a), b); }
inline vec2 operator/(ivec2 a, vec2 b) { return vec2(a) / b; }

// ---- built-in variables ---------------------------------------------------------------------------------------------
GLSL_GLOBAL vec4 gl_FragCoord;
GLSL_GLOBAL uvec3 gl_GlobalInvocationID;
} // namespace glsl

// shadow_map.cu -- renders a light's shadow map by casting one closest-hit ray per texel through the same
// BLAS/TLAS the lighting rays use (there is no rasteriser on this path).
//
// Replaces DeferredRenderer::ShadowMapPass (DeferredRenderer.cpp:268-291) with shadowMap.vert:15-17,
// shadowMap.geom:17-38 and shadowMap.frag:12-18:
//   * the pipeline culls FRONT faces, front = counter-clockwise in framebuffer space (DeferredRenderer.cpp:90-101
//     `.cullFront = true`; VulkanWrapper.cpp:941-946), depth test LESS against a clear value of 1.0
//     (VulkanWrapper.cpp:963, :1212), no depth bias, no depth clamp (:936, :947);
//   * a fragment exists where the texel centre is covered, so the ray goes through the texel centre;
//   * point lights: layer f is rendered with light.viewProj[f] = perspective(90 deg, 1, 0, zFar) * lookAt(pos,
//     pos + axis_f, up_f) (GPUScene.cpp:268-276), which maps a direction r to exactly the (sc, tc) / |ma| of the
//     Vulkan cube face table (tests/test_host_golden.py checks the six matrices), so the texel-centre ray of layer
//     f is that table inverted; the stored value is |light.position - fragPos| / zFar (shadowMap.frag:14-15);
//   * spot / directional lights: one layer under the orthographic light.viewProj[0] (GPUScene.cpp:278-310); the
//     ray runs from NDC z = 0 to z = 1 through the texel centre, its parameter t IS gl_FragCoord.z, and fragments
//     outside 0 <= z <= 1 are clipped.
// Facing: with clip coordinates c_i of a triangle's vertices, the framebuffer area has the sign of
// -det[c_0; c_1; c_2] over (x, y, w), so a triangle is front-facing iff that determinant is negative; along a ray d
// that reaches the triangle this equals s_view * sign(det M_instance) * dot(d, (v1 - v0) x (v2 - v0)) < 0, where
// s_view = sign of the determinant of the view's linear part (rows x, y, w for a perspective face; x, y, z for an
// orthographic view), computed by the host per layer.
#include "passes.h"
#include "traverse.cuh"

namespace luz {

namespace {

__global__ void __launch_bounds__(128) k_shadow_map(const ShadowMapArgs a) {
    // one warp = 8 x 4 texels (neighbouring rays stay together), one CTA = 16 x 8
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t x = blockIdx.x * 16 + (warp & 1) * 8 + (lane & 7);
    const uint32_t y = blockIdx.y * 8 + (warp >> 1) * 4 + (lane >> 3);
    const uint32_t layer = blockIdx.z;
    if (x >= a.res || y >= a.res) return;
    const float sc = ((float)x + 0.5f) / (float)a.res * 2.0f - 1.0f;
    const float tc = ((float)y + 0.5f) / (float)a.res * 2.0f - 1.0f;
    float3 o, d;
    float tmax;
    if (a.is_cube) {
        o = f3(a.eye[0], a.eye[1], a.eye[2]);
        switch (layer) { // inverse of the cube face table: (sc, tc, ma = 1) -> direction
            case 0: d = f3(1.0f, -tc, -sc); break;
            case 1: d = f3(-1.0f, -tc, sc); break;
            case 2: d = f3(sc, 1.0f, tc); break;
            case 3: d = f3(sc, -1.0f, -tc); break;
            case 4: d = f3(sc, -tc, 1.0f); break;
            default: d = f3(-sc, -tc, -1.0f); break;
        }
        tmax = 3.0e38f;
    } else {
        const float4 o4 = mat_mul(a.inv_view_proj, f4(sc, tc, 0.0f, 1.0f));
        o = f3(o4.x, o4.y, o4.z);
        d = f3(a.inv_view_proj[8], a.inv_view_proj[9], a.inv_view_proj[10]);
        tmax = 1.0f;
    }
    uint2 stack[LUZ_STACK_SIZE];
    HitInfo h;
    h.t = 0.0f;
    h.cull_sign = a.cull_sign[layer];
    float depth = 1.0f; // the clear value
    if (trace_ray<true, false, true>(a.scene, o, d, 0.0f, tmax, &h, nullptr, stack)) {
        const float z = a.is_cube ? (h.t * length3(d)) / a.z_far : h.t;
        if (z < 1.0f) depth = z; // VK_COMPARE_OP_LESS against the cleared 1.0
    }
    a.out[((size_t)layer * a.res + y) * a.res + x] = depth;
}

} // namespace

cudaError_t launch_shadow_map(cudaStream_t stream, const ShadowMapArgs& args) {
    if (args.res == 0 || args.layers == 0) return cudaSuccess;
    const dim3 grid((args.res + 15) / 16, (args.res + 7) / 8, args.layers);
    k_shadow_map<<<grid, 128, 0, stream>>>(args);
    return cudaGetLastError();
}

} // namespace luz

// gbuffer.cu -- G-buffer producer by primary visibility.
//
// The reference rasterises every model into four MRTs + depth with opaque.vert:21-31 /
// opaque.frag:21-59 (draw loop source/Core/main.cpp:242-258, attachment formats
// DeferredRenderer.cpp:176-238, clears colour 0 / depth 1 VulkanWrapper.cpp:1194-1196, :1212).
// Here the same five attachments come from the closest hit of the primary ray through each pixel
// centre, traced through the same TLAS/BLAS as the shadow rays; the fragment-stage arithmetic
// (material fetches, alpha test, TBN normal) follows opaque.frag.  The Opaque Pipeline culls back
// faces (VK_CULL_MODE_BACK_BIT, front = counter-clockwise; VulkanWrapper.cpp:941-946): the primary ray
// ignores back-facing triangles the same way (trace_ray<FACE_CULL>, sign from the camera's viewProj and
// the instance determinant), so a camera inside a mesh, single-sided planes seen from behind and
// mirrored instances give the reference's coverage.  What remains different from the rasteriser is
// coverage on triangle edges.
#include "passes.h"
#include "traverse.cuh"

namespace luz {

namespace {

__device__ __forceinline__ int wrapi(int i, int n) {
    int r = i % n;
    return r < 0 ? r + n : r;
}

__device__ __forceinline__ unsigned char unorm8(float v) {
    if (!(v > 0.0f)) return 0;
    if (v > 1.0f) v = 1.0f;
    return (unsigned char)(int)floorf(v * 255.0f + 0.5f);
}

// bilinear REPEAT fetch of an RGBA8 texture at LOD 0 (the global sampler, VulkanWrapper.cpp:2429-2461)
__device__ float4 tex_rgba8(const uchar4* data, uint2 size, float u, float v) {
    const float x = u * (float)size.x - 0.5f, y = v * (float)size.y - 0.5f;
    const float fx0 = floorf(x), fy0 = floorf(y);
    const float fx = x - fx0, fy = y - fy0;
    const int x0 = (int)fx0, y0 = (int)fy0;
    auto at = [&](int xi, int yi) -> float4 {
        const uchar4 p = __ldg(data + (size_t)wrapi(yi, (int)size.y) * size.x + wrapi(xi, (int)size.x));
        return f4((float)p.x / 255.0f, (float)p.y / 255.0f, (float)p.z / 255.0f, (float)p.w / 255.0f);
    };
    const float4 top = at(x0, y0) * (1.0f - fx) + at(x0 + 1, y0) * fx;
    const float4 bot = at(x0, y0 + 1) * (1.0f - fx) + at(x0 + 1, y0 + 1) * fx;
    return top * (1.0f - fy) + bot * fy;
}

__device__ __forceinline__ bool tex_valid(const GbufferArgs& a, int rid) {
    return rid >= 0 && (uint32_t)rid < a.n_textures && a.tex_data[rid] != nullptr;
}

// (A^-1)^T * n, where the InstanceRec rows hold A^-1: transpose(inverse(mat3(model))) * n
__device__ __forceinline__ float3 normal_xform(const float4 r0, const float4 r1, const float4 r2, float3 n) {
    return f3(r0.x * n.x + r1.x * n.y + r2.x * n.z, r0.y * n.x + r1.y * n.y + r2.y * n.z,
              r0.z * n.x + r1.z * n.y + r2.z * n.z);
}

__global__ void __launch_bounds__(128) k_gbuffer(const GbufferArgs a) {
    const FrameConst& fc = a.fc;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t x = blockIdx.x * 16 + (warp & 1) * 8 + (lane & 7);
    const uint32_t r = blockIdx.y * 8 + (warp >> 1) * 4 + (lane >> 3);
    if (x >= fc.width || r >= a.rows.rows) return;
    const uint32_t y = band_row(fc, a.rows, blockIdx.z, r);
    const size_t pix = (size_t)y * fc.width + x;

    const float u = ((float)x + 0.5f) / (float)fc.width, v = ((float)y + 0.5f) / (float)fc.height;
    const float3 pn = depth_to_world(fc, u, v, 0.0f);
    const float3 pf = depth_to_world(fc, u, v, 1.0f);
    const float3 d = pf - pn;

    float tmin = 0.0f;
    bool have = false;
    uint2 stack[LUZ_STACK_SIZE];
    float4 albedo = f4(0, 0, 0, 0), emission = f4(0, 0, 0, 0);
    float roughness = 0.0f, metallic = 0.0f, occl = 1.0f, depth = 1.0f;
    float3 N = f3(0, 0, 0);
    for (int iter = 0; iter < 16 && !have; iter++) {
        HitInfo h;
        h.cull_sign = a.cull_sign;
        if (!trace_ray<true, false, true>(a.scene, pn, d, tmin, 1.0f, &h, nullptr, stack)) break;
        const InstanceMeta im = a.inst_meta[h.inst];
        if (im.custom_index >= a.n_models) break;
        const luzw_model_block& mb = a.models[im.custom_index];
        const BlasAttr ba = a.blas_attr[im.blas_slot];
        const InstanceRec* rec = a.scene.instances + h.inst;
        const float4 r0 = __ldg(&rec->r0), r1 = __ldg(&rec->r1), r2 = __ldg(&rec->r2);
        const float b0 = h.bu, b1 = h.bv, b2 = 1.0f - h.bu - h.bv;
        float3 n0 = f3(0, 0, 0), n1 = n0, n2 = n0;
        float4 tg0 = f4(0, 0, 0, 0), tg1 = tg0, tg2 = tg0;
        float uvx = 0.0f, uvy = 0.0f;
        if (ba.has_attr) {
            const uint32_t i0 = ba.indices[3 * h.prim], i1 = ba.indices[3 * h.prim + 1], i2 = ba.indices[3 * h.prim + 2];
            const float* a0 = reinterpret_cast<const float*>(ba.vertices + (size_t)i0 * ba.stride) + 3;
            const float* a1 = reinterpret_cast<const float*>(ba.vertices + (size_t)i1 * ba.stride) + 3;
            const float* a2 = reinterpret_cast<const float*>(ba.vertices + (size_t)i2 * ba.stride) + 3;
            n0 = f3(a0[0], a0[1], a0[2]);
            n1 = f3(a1[0], a1[1], a1[2]);
            n2 = f3(a2[0], a2[1], a2[2]);
            tg0 = f4(a0[3], a0[4], a0[5], a0[6]);
            tg1 = f4(a1[3], a1[4], a1[5], a1[6]);
            tg2 = f4(a2[3], a2[4], a2[5], a2[6]);
            uvx = a0[7] * b0 + a1[7] * b1 + a2[7] * b2;
            uvy = a0[8] * b0 + a1[8] * b1 + a2[8] * b2;
        }
        albedo = f4(mb.color[0], mb.color[1], mb.color[2], mb.color[3]);
        if (tex_valid(a, mb.color_map)) {
            const float4 t = tex_rgba8(a.tex_data[mb.color_map], a.tex_size[mb.color_map], uvx, uvy);
            albedo = f4(albedo.x * t.x, albedo.y * t.y, albedo.z * t.z, albedo.w * t.w);
        }
        if (albedo.w < 0.5f) { // discard (opaque.frag:32-34): look behind this surface
            tmin = h.t;
            continue;
        }
        roughness = mb.roughness;
        metallic = mb.metallic;
        occl = 1.0f;
        emission = f4(mb.emission[0], mb.emission[1], mb.emission[2], 1.0f);
        float3 normalSample = f3(1, 1, 1);
        if (tex_valid(a, mb.metallic_roughness_map)) {
            const float4 t = tex_rgba8(a.tex_data[mb.metallic_roughness_map], a.tex_size[mb.metallic_roughness_map], uvx, uvy);
            roughness *= t.y;
            metallic *= t.z;
        }
        if (tex_valid(a, mb.ao_map)) occl = tex_rgba8(a.tex_data[mb.ao_map], a.tex_size[mb.ao_map], uvx, uvy).x;
        if (tex_valid(a, mb.normal_map)) {
            const float4 t = tex_rgba8(a.tex_data[mb.normal_map], a.tex_size[mb.normal_map], uvx, uvy);
            normalSample = f3(t.x, t.y, t.z);
        }
        if (tex_valid(a, mb.emission_map)) {
            const float4 t = tex_rgba8(a.tex_data[mb.emission_map], a.tex_size[mb.emission_map], uvx, uvy);
            emission = f4(emission.x * t.x, emission.y * t.y, emission.z * t.z, emission.w * t.w);
        }
        // opaque.vert:25-30 per vertex, interpolated with the hit's barycentrics
        float3 fn[3], ft[3], fb[3];
        const float3 nn[3] = {n0, n1, n2};
        const float4 tt[3] = {tg0, tg1, tg2};
#pragma unroll
        for (int k = 0; k < 3; k++) {
            ft[k] = normalize3(normal_xform(r0, r1, r2, f3(tt[k].x, tt[k].y, tt[k].z)));
            fn[k] = normalize3(normal_xform(r0, r1, r2, nn[k]));
            ft[k] = normalize3(ft[k] - dot3(ft[k], fn[k]) * fn[k]);
            fb[k] = cross3(fn[k], ft[k]) * tt[k].w;
        }
        const float3 fragNormal = fn[0] * b0 + fn[1] * b1 + fn[2] * b2;
        const float3 fragTangent = ft[0] * b0 + ft[1] * b1 + ft[2] * b2;
        const float3 fragBitan = fb[0] * b0 + fb[1] * b1 + fb[2] * b2;
        const bool tangent_zero = fragTangent.x == 0.0f && fragTangent.y == 0.0f && fragTangent.z == 0.0f;
        const bool ns_one = normalSample.x == 1.0f && normalSample.y == 1.0f && normalSample.z == 1.0f;
        if (tangent_zero || ns_one) {
            N = normalize3(fragNormal);
        } else {
            const float3 ts = normalize3(normalSample * 2.0f - f3(1, 1, 1));
            N = normalize3(fragTangent * ts.x + fragBitan * ts.y + fragNormal * ts.z);
        }
        const float3 wp = pn + d * h.t;
        const float4 clip = mat_mul(fc.view_proj, f4(wp.x, wp.y, wp.z, 1.0f));
        depth = clip.z / clip.w;
        have = true;
    }
    if (!have) {
        a.albedo[pix] = make_uchar4(0, 0, 0, 0);
        a.normal[pix] = make_float4(0, 0, 0, 0);
        a.material[pix] = make_uchar4(0, 0, 0, 0);
        a.emission[pix] = make_uchar4(0, 0, 0, 0);
        a.depth[pix] = 1.0f;
        return;
    }
    a.albedo[pix] = make_uchar4(unorm8(albedo.x), unorm8(albedo.y), unorm8(albedo.z), unorm8(albedo.w));
    a.normal[pix] = make_float4(N.x, N.y, N.z, 1.0f);
    a.material[pix] = make_uchar4(unorm8(roughness), unorm8(metallic), unorm8(occl), 255);
    a.emission[pix] = make_uchar4(unorm8(emission.x), unorm8(emission.y), unorm8(emission.z), unorm8(emission.w));
    a.depth[pix] = depth;
}

} // namespace

cudaError_t launch_gbuffer_pass(cudaStream_t stream, const GbufferArgs& args) {
    if (args.rows.rows == 0 || args.rows.n_bands == 0 || args.fc.width == 0) return cudaSuccess;
    const dim3 grid((args.fc.width + 15) / 16, (args.rows.rows + 7) / 8, args.rows.n_bands);
    k_gbuffer<<<grid, 128, 0, stream>>>(args);
    return cudaGetLastError();
}

} // namespace luz

// bvh.h -- host-side interface of the BVH builder (bvh_build.cu).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "common.cuh"

namespace luz {

struct BoxF {
    float lox, loy, loz, hix, hiy, hiz;
};

// A built 8-wide BVH living in device memory.
struct WideBvh {
    WideNode* nodes = nullptr;      // n_nodes, breadth-first: level l occupies levels[l]
    BoxF* node_bounds = nullptr;    // n_nodes exact boxes (kept for refit)
    uint32_t* prim_order = nullptr; // n_prims: leaf order -> index of the input primitive
    uint32_t n_nodes = 0;
    uint32_t n_prims = 0;
    std::vector<uint2> levels;      // (first node, node count) per level, root first
    size_t node_capacity = 0, prim_capacity = 0;
};

// Grow-only scratch shared by every build on a ctx.
struct BuildScratch {
    void* mem = nullptr;
    size_t bytes = 0;
    uint64_t* host_pair = nullptr; // pinned, 2 x u64
    uint32_t* host_levels = nullptr; // pinned level table of the single-CTA collapse
    ~BuildScratch();
};

// Builds a WideBvh over n boxes (device pointer).  max_leaf = primitives per leaf slot (3 for
// triangles, 1 for instances).  Deterministic.  Synchronises the stream once (trees of up to 2^17
// primitives: TLAS rebuilds) or once per tree level (large meshes, built once).
// Returns cudaSuccess or the failing CUDA error.  *launches counts kernels launched.
cudaError_t build_wide_bvh(cudaStream_t stream, BuildScratch& scratch, const BoxF* d_boxes, uint32_t n,
                           uint32_t max_leaf, WideBvh& out, uint64_t* launches);

// Recomputes every node of `bvh` bottom-up from new primitive boxes (same count, same topology).
cudaError_t refit_wide_bvh(cudaStream_t stream, const BoxF* d_boxes, WideBvh& bvh, uint64_t* launches);

void free_wide_bvh(WideBvh& bvh);

// Triangle boxes of an indexed mesh (positions at the start of each `stride`-byte vertex).
cudaError_t launch_triangle_boxes(cudaStream_t stream, const uint8_t* d_vertices, uint32_t stride,
                                  const uint32_t* d_indices, uint32_t n_tris, BoxF* d_boxes);
// Writes the 48-byte triangles of a BLAS in leaf order.
cudaError_t launch_gather_triangles(cudaStream_t stream, const uint8_t* d_vertices, uint32_t stride,
                                    const uint32_t* d_indices, const uint32_t* d_prim_order, uint32_t n_tris,
                                    WideTri* d_tris);

// Per-instance input for the TLAS kernels (device copy of what luzrt_tlas_build receives).
struct InstanceIn {
    float m[16];
    const WideNode* nodes;
    const WideTri* tris;
    BoxF blas_bounds;
    uint32_t custom_index;
    uint32_t blas_slot;
    uint32_t n_tris; // triangles of the BLAS (stored in the w of the instance's world box: bounds of per-ray occluder hints)
    uint32_t pad;
};
// World boxes + inverse transforms for n instances (input order).
cudaError_t launch_instance_prepare(cudaStream_t stream, const InstanceIn* d_in, uint32_t n, BoxF* d_boxes,
                                    InstanceRec* d_recs_in_order, InstanceMeta* d_meta_in_order);
// Permutes records into TLAS leaf order.
cudaError_t launch_instance_gather(cudaStream_t stream, const InstanceIn* d_in, const InstanceRec* d_recs_in, const InstanceMeta* d_meta_in,
                                   const BoxF* d_boxes_in, const uint32_t* d_prim_order, uint32_t n,
                                   InstanceRec* d_recs_out, InstanceMeta* d_meta_out, float4* d_boxes_out,
                                   uint32_t* d_leaf_of_input /* n: inverse of d_prim_order */);
// Adds the content hash of `bytes` of device data to *d_acc (a zeroed u64 on the device): of every stride_bytes-long
// record only the first take_bytes count (stride_bytes == take_bytes: everything).  Order independent, deterministic.
cudaError_t launch_hash_words(cudaStream_t stream, const void* data, size_t bytes, uint64_t seed, size_t stride_bytes,
                              size_t take_bytes, unsigned long long* d_acc);
// Boxes permuted into leaf order (for refit).
cudaError_t launch_box_gather(cudaStream_t stream, const BoxF* d_in, const uint32_t* d_prim_order, uint32_t n,
                              BoxF* d_out);

} // namespace luz

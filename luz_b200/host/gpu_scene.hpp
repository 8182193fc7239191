// gpu_scene.hpp -- mirror of the reference's GPUScene (source/Graphics/GPUScene.hpp:17-62) and of
// the two DeferredRenderer passes on the lighting path (DeferredRenderer.hpp:17-54), routed through
// the C ABI of libluzrt.so instead of vkw::.
#pragma once

#include <unordered_map>
#include <vector>

#include "../../include/luzrt.h"
#include "scene.hpp"

namespace luzhost {

struct GPUMesh { // GPUScene.hpp:17-23
    uint32_t vertexCount = 0, indexCount = 0;
    luzrt_blas blas = 0;
};
struct GPUTexture { // GPUScene.hpp:25-27
    int32_t rid = -1;
};
struct GPUModel { // GPUScene.hpp:29-33
    GPUMesh mesh;
    uint32_t modelRID = 0;
    Ref<MeshNode> node;
};

struct GPUScene {
    explicit GPUScene(luzrt_ctx* rt) : rt(rt) {}

    int AddMesh(const Ref<MeshAsset>& asset);       // GPUScene.cpp:108-138
    int AddTexture(const Ref<TextureAsset>& asset); // :140-156
    int AddAssets(const AssetManager& assets);      // :164-179
    void ClearAssets();                             // :158-162 (also destroys the BLASes)

    // :181-346, CPU only: scene graph + camera -> SceneBlock, ModelBlock[]
    void UpdateResources(const Ref<SceneAsset>& scene, const Ref<CameraNode>& camera);
    // :348-366: uploads SceneBlock (+ lights beyond LUZ_MAX_LIGHTS) and (re)builds the TLAS.
    // tlasMode 0 = rebuild (what the reference does every frame), 1 = refit.
    int UpdateResourcesGPU(int tlasMode = 0);

    std::vector<GPUModel>& GetMeshModels() { return meshModels; }
    const luzw_scene_block& GetSceneBlock() const { return sceneBlock; }
    const std::vector<luzw_model_block>& GetModelsBlock() const { return modelsBlock; }
    const std::vector<luzw_light_block>& GetExtraLights() const { return extraLights; }
    const std::vector<luzrt_instance>& GetInstances() const { return instances; }

    luzrt_ctx* rt;
    luzw_scene_block sceneBlock{};
    std::vector<luzw_model_block> modelsBlock;
    std::vector<luzw_light_block> extraLights;
    std::vector<GPUModel> meshModels;
    std::vector<luzrt_instance> instances;
    std::unordered_map<UUID, GPUMesh> meshes;
    std::unordered_map<UUID, GPUTexture> textures;
    static void ShadowViewProj(const LightNode& light, const CameraNode& camera, const lm::mat4& sceneView,
                               float (*out)[16]); // GPUScene.cpp:266-311
    bool anyVolumetricLight = false; // GPUScene.cpp:16, :403-405
    bool AnyVolumetricLight() const { return anyVolumetricLight; }
    bool anyShadowMapVolumetric = false;
    bool AnyShadowMapVolumetric() const { return anyShadowMapVolumetric; }
    bool firstFrame = true;
    int32_t nextCpuRid = 0;
};

struct LightConstants { // DeferredRenderer.hpp:17-26 (the RIDs are meaningless without bindless sets)
    int sceneBufferIndex = 0, modelBufferIndex = 0, frameID = 0;
};

struct DeferredRenderer {
    explicit DeferredRenderer(luzrt_ctx* rt) : rt(rt) {}
    int CreateImages(uint32_t width, uint32_t height);                      // DeferredRenderer.cpp:175-248
    int OpaquePass(GPUScene& gpuScene);                                     // main.cpp:242-258 (G-buffer)
    int LightPass(LightConstants constants);                                // DeferredRenderer.cpp:324-345
    int ShadowMapPass(const Ref<SceneAsset>& scene);                        // :268-291, for every light (main.cpp:260-264)
    int ScreenSpaceVolumetricLightPass(GPUScene& gpuScene, int frame);      // :294-307 (+ :309-322)
    int TAAPass(GPUScene& gpuScene, const Ref<SceneAsset>& scene);          // :425-445
    int ComposePass(const Ref<SceneAsset>& scene);                          // :347-370
    int SwapLightHistory();                                                 // :469-471
    luzrt_ctx* rt;
};

} // namespace luzhost

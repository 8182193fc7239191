// gpu_scene.cpp -- GPUScene mirror: flattens the scene graph into SceneBlock / ModelBlock[] exactly
// as source/Graphics/GPUScene.cpp:181-346 does, and hands meshes / instances / blocks to libluzrt.
#include "gpu_scene.hpp"

#include <cstring>

namespace luzhost {

using lm::mat4;
using lm::vec3;
using lm::vec4;

static void put_mat(float* dst, const mat4& m) { memcpy(dst, m.data(), 64); }
static void put_vec3(float* dst, vec3 v) {
    dst[0] = v.x;
    dst[1] = v.y;
    dst[2] = v.z;
}

int GPUScene::AddMesh(const Ref<MeshAsset>& asset) {
    GPUMesh& mesh = meshes[asset->uuid];
    if (mesh.blas && rt) luzrt_blas_destroy(rt, mesh.blas);
    mesh.vertexCount = (uint32_t)asset->vertices.size();
    mesh.indexCount = (uint32_t)asset->indices.size();
    if (!rt) return LUZRT_OK; // CPU-only use of the mirror (block fill without a device)
    return luzrt_blas_create(rt, asset->vertices.data(), mesh.vertexCount, (uint32_t)sizeof(MeshAsset::MeshVertex),
                             asset->indices.data(), mesh.indexCount, &mesh.blas);
}

int GPUScene::AddTexture(const Ref<TextureAsset>& asset) {
    if (asset->channels != 4) return LUZRT_E_INVALID; // ASSERT in the reference (GPUScene.cpp:142)
    GPUTexture& t = textures[asset->uuid];
    if (!rt) { // CPU-only: RIDs are handed out in creation order, like luzrt_texture_create does
        t.rid = nextCpuRid++;
        return LUZRT_OK;
    }
    return luzrt_texture_create(rt, asset->data.data(), (uint32_t)asset->width, (uint32_t)asset->height, &t.rid);
}

void GPUScene::ClearAssets() {
    for (auto& kv : meshes)
        if (kv.second.blas && rt) luzrt_blas_destroy(rt, kv.second.blas);
    meshes.clear();
    textures.clear();
    meshModels.clear();
}

int GPUScene::AddAssets(const AssetManager& assets) {
    for (auto& mesh : assets.GetAll<MeshAsset>(ObjectType::MeshAsset)) {
        if (mesh->gpuDirty) {
            const int rc = AddMesh(mesh);
            if (rc != LUZRT_OK) return rc;
            mesh->gpuDirty = false;
        }
    }
    for (auto& tex : assets.GetAll<TextureAsset>(ObjectType::TextureAsset)) {
        if (tex->gpuDirty) {
            const int rc = AddTexture(tex);
            if (rc != LUZRT_OK) return rc;
            tex->gpuDirty = false;
        }
    }
    return LUZRT_OK;
}

void GPUScene::UpdateResources(const Ref<SceneAsset>& scene, const Ref<CameraNode>& camera) {
    std::vector<Ref<MeshNode>> meshNodes;
    for (auto& n : scene->nodes) { // SceneAsset::GetAll order: pre-order over the node tree
        if (n->type == ObjectType::MeshNode) meshNodes.emplace_back(std::dynamic_pointer_cast<MeshNode>(n));
        n->GetAll<MeshNode>(ObjectType::MeshNode, meshNodes);
    }
    modelsBlock.clear();
    meshModels.clear();
    instances.clear();
    for (const auto& node : meshNodes) {
        if (!node->mesh) continue; // the reference would dereference null; skip instead
        GPUModel gm;
        gm.mesh = meshes[node->mesh->uuid];
        gm.modelRID = (uint32_t)modelsBlock.size();
        gm.node = node;
        meshModels.push_back(gm);
        luzw_model_block block{}; // defaultModelBlock, GPUScene.cpp:19-30
        const mat4 ident(1.0f);
        put_mat(block.model_mat, ident);
        block.color[0] = block.color[1] = block.color[2] = block.color[3] = 1.0f;
        block.metallic = 0.0f;
        block.roughness = 0.5f;
        block.ao_map = block.color_map = block.normal_map = block.emission_map = block.metallic_roughness_map = -1;
        if (const Ref<MaterialAsset>& material = node->material) {
            for (int k = 0; k < 4; k++) block.color[k] = material->color[k];
            put_vec3(block.emission, material->emission);
            block.metallic = material->metallic;
            block.roughness = material->roughness;
            auto rid = [&](const Ref<TextureAsset>& t) { return textures[t->uuid].rid; };
            if (material->colorMap) block.color_map = rid(material->colorMap);
            if (material->normalMap) block.normal_map = rid(material->normalMap);
            if (material->metallicRoughnessMap) block.metallic_roughness_map = rid(material->metallicRoughnessMap);
            if (material->emissionMap) block.emission_map = rid(material->emissionMap);
            // (the reference never binds aoMap: GPUScene.cpp:200-211)
        }
        block.vertex_buffer = block.index_buffer = 0;
        const mat4 world = node->GetWorldTransform();
        put_mat(block.model_mat, world);
        modelsBlock.push_back(block);
        // the BLASInstance list of UpdateResourcesGPU (GPUScene.cpp:355-363)
        luzrt_instance inst{};
        inst.blas = gm.mesh.blas;
        put_mat(inst.model_mat, world);
        inst.custom_index = gm.modelRID;
        instances.push_back(inst);
    }

    luzw_scene_block& s = sceneBlock;
    s.num_lights = 0;
    extraLights.clear();
    // camera->eye is a by-product of GetView(); like the reference (GPUScene.cpp:224) this reads the value
    // left by the previous GetView() call (or by the project file), i.e. camPos lags a moving camera by a frame
    put_vec3(s.cam_pos, camera->eye);
    memcpy(s.prev_view_proj, s.view_proj, 64); // :225
    const lm::vec2 pj = camera->GetJitter();
    s.prev_jitter[0] = pj.x;
    s.prev_jitter[1] = pj.y;
    camera->NextJitter();
    const lm::vec2 jt = camera->GetJitter();
    s.jitter[0] = jt.x;
    s.jitter[1] = jt.y;
    const mat4 proj = camera->GetProjJittered();
    const mat4 view = camera->GetView();
    put_mat(s.proj, proj);
    put_mat(s.view, view);
    put_mat(s.view_proj, proj * view);
    put_mat(s.inverse_proj, lm::inverse(proj));
    put_mat(s.inverse_view, lm::inverse(view));
    if (firstFrame) {
        // the reference leaves prevViewProj uninitialised on its very first frame (GPUScene.cpp:14, :43);
        // this path defines frame 0 as "no motion": prevViewProj = viewProj, prevJitter = jitter
        memcpy(s.prev_view_proj, s.view_proj, 64);
        s.prev_jitter[0] = s.jitter[0];
        s.prev_jitter[1] = s.jitter[1];
        firstFrame = false;
    }

    anyVolumetricLight = false; // GPUScene.cpp:237
    anyShadowMapVolumetric = false;
    int nLights = 0;
    for (const auto& light : scene->GetAll<LightNode>(ObjectType::LightNode)) {
        luzw_light_block local{};
        luzw_light_block* block;
        if (s.num_lights < LUZW_MAX_LIGHTS) {
            block = &s.lights[s.num_lights++];
            memset(block, 0, sizeof(*block));
        } else { // the reference writes past the array here (GPUScene.cpp:240); keep the overflow separately
            extraLights.push_back(local);
            block = &extraLights.back();
        }
        put_vec3(block->color, light->color);
        block->intensity = light->intensity;
        put_vec3(block->position, light->GetWorldPosition());
        block->inner_angle = lm::radians(light->innerAngle);
        put_vec3(block->direction, light->GetWorldFront()); // world * (0,-1,0,0), unnormalised
        block->outer_angle = lm::radians(light->outerAngle);
        block->type = light->lightType;
        block->num_shadow_samples = scene->shadowType == ShadowRayTraced ? scene->lightSamples : 0;
        block->radius = light->radius;
        block->z_far = light->shadowMapFar;
        // the reference stores the bindless RID of the light's shadow-map image here (GPUScene.cpp:315-329), never
        // -1 after this function; luzrt only distinguishes -1 from "has a map" and indexes maps by light
        block->shadow_map = (int32_t)nLights;
        ShadowViewProj(*light, *camera, view, block->view_proj);
        nLights++;
        block->volumetric_type = light->volumetricType;
        if (light->volumetricType == LightNode::ScreenSpace) {
            block->volumetric_samples = light->volumetricScreenSamples;
            block->volumetric_absorption = light->volumetricScreenAbsorption;
            anyVolumetricLight = true;
        } else if (light->volumetricType == LightNode::ShadowMapVolumetric) {
            block->volumetric_weight = light->volumetricShadowWeight;
            block->volumetric_samples = light->volumetricShadowSamples;
            block->volumetric_density = light->volumetricShadowDensity;
            block->volumetric_absorption = light->volumetricShadowAbsorption;
            anyVolumetricLight = true;
            anyShadowMapVolumetric = true;
        }
    }
    put_vec3(s.ambient_light_color, scene->ambientLightColor);
    s.ambient_light_intensity = scene->ambientLight;
    s.ao_max = scene->aoMax;
    s.ao_min = scene->aoMin;
    s.ao_num_samples = scene->aoSamples > 0 ? scene->aoSamples : 0;
    s.exposure = scene->exposure;
    s.tlas_rid = 0;
    s.blue_noise_texture = 0;
    s.white_texture = -1;
    s.black_texture = -1;
    s.shadow_type = scene->shadowType;
}

// light.viewProj[] as GPUScene.cpp:266-311 fills it: six perspective(90 deg, 1, 0, shadowMapFar) * lookAt cube
// faces for a point light; for the others one orthographic matrix fitted to the part of the camera frustum
// within farDistance / shadowMapRange, seen along the light's front vector.  Restated literally, including the
// bounds being seeded with the first corner in WORLD space (:299-300) before the view-space corners are merged.
void GPUScene::ShadowViewProj(const LightNode& light, const CameraNode& camera, const lm::mat4& sceneView,
                              float (*out)[16]) {
    using lm::vec3;
    using lm::vec4;
    const vec3 pos = light.GetWorldPosition();
    if (light.lightType == LightNode::Point) {
        const mat4 proj = lm::perspective(lm::radians(90.0f), 1.0f, 0.0f, light.shadowMapFar);
        const vec3 axis[6] = {vec3(1, 0, 0), vec3(-1, 0, 0), vec3(0, 1, 0), vec3(0, -1, 0), vec3(0, 0, 1), vec3(0, 0, -1)};
        const vec3 up[6] = {vec3(0, -1, 0), vec3(0, -1, 0), vec3(0, 0, 1), vec3(0, 0, -1), vec3(0, -1, 0), vec3(0, -1, 0)};
        for (int f = 0; f < 6; f++) put_mat(out[f], proj * lm::look_at(pos, pos + axis[f], up[f]));
        return;
    }
    const mat4 camProjView = lm::inverse(camera.GetProj(camera.nearDistance, camera.farDistance / light.shadowMapRange) * sceneView);
    vec4 corners[8];
    int n = 0;
    for (int i = 0; i < 2; i++)
        for (int j = 0; j < 2; j++)
            for (int k = 0; k < 2; k++) {
                const vec4 pt = camProjView * vec4(2.0f * i - 1.0f, 2.0f * j - 1.0f, 2.0f * k - 1.0f, 1.0f);
                corners[n++] = vec4(pt.x / pt.w, pt.y / pt.w, pt.z / pt.w, pt.w / pt.w);
            }
    vec3 centre(0.0f);
    for (const vec4& p : corners) centre = centre + vec3(p.x, p.y, p.z);
    centre = vec3(centre.x / 8.0f, centre.y / 8.0f, centre.z / 8.0f);
    const float r = light.shadowMapRange;
    const mat4 view = lm::look_at(centre + light.GetWorldFront(), centre, vec3(0.0f, 1.0f, 0.0f));
    vec3 mn(corners[0].x, corners[0].y, corners[0].z), mx = mn;
    auto vmin = [](float a, float b) { return b < a ? b : a; }; // glm::min(x, y) = y < x ? y : x
    auto vmax = [](float a, float b) { return a < b ? b : a; }; // glm::max(x, y) = x < y ? y : x
    for (const vec4& p : corners) {
        const vec4 q = view * p;
        mn = vec3(vmin(q.x, mn.x), vmin(q.y, mn.y), vmin(q.z, mn.z));
        mx = vec3(vmax(q.x, mx.x), vmax(q.y, mx.y), vmax(q.z, mx.z));
    }
    mn.z = mn.z < 0 ? mn.z * r : mn.z / r;
    mx.z = mx.z < 0 ? mx.z / r : mx.z * r;
    put_mat(out[0], lm::ortho(mn.x, mx.x, mn.y, mx.y, mx.z, mn.z) * view);
}

int GPUScene::UpdateResourcesGPU(int tlasMode) {
    if (!rt) return LUZRT_E_NODEVICE;
    int rc = luzrt_set_scene(rt, &sceneBlock, extraLights.empty() ? nullptr : extraLights.data(),
                             (uint32_t)extraLights.size());
    if (rc != LUZRT_OK) return rc;
    // the reference returns early when there are no models (GPUScene.cpp:349-351) and keeps the old TLAS;
    // an empty TLAS is the well-defined equivalent
    return luzrt_tlas_build(rt, instances.empty() ? nullptr : instances.data(), (uint32_t)instances.size(), tlasMode);
}

// ---- DeferredRenderer -------------------------------------------------------------------------------
int DeferredRenderer::CreateImages(uint32_t width, uint32_t height) { return luzrt_resize(rt, width, height); }
int DeferredRenderer::OpaquePass(GPUScene& g) {
    return luzrt_gbuffer_pass(rt, g.modelsBlock.empty() ? nullptr : g.modelsBlock.data(), (uint32_t)g.modelsBlock.size());
}
int DeferredRenderer::LightPass(LightConstants c) { return luzrt_light_pass(rt, (uint32_t)c.frameID); }
int DeferredRenderer::ShadowMapPass(const Ref<SceneAsset>& scene) { return luzrt_shadow_map_pass(rt, scene->shadowResolution); }
int DeferredRenderer::ScreenSpaceVolumetricLightPass(GPUScene&, int frame) { return luzrt_volumetric_pass(rt, (uint32_t)frame); }
int DeferredRenderer::TAAPass(GPUScene&, const Ref<SceneAsset>& scene) {
    if (!scene->taaEnabled) return LUZRT_OK; // DeferredRenderer.cpp:426
    int rc = luzrt_taa_pass(rt, scene->taaReconstruct ? 1 : 0);
    if (rc != LUZRT_OK) return rc;
    return luzrt_gather(rt); // no-op on one GPU; the assembled frame is next frame's history
}
int DeferredRenderer::ComposePass(const Ref<SceneAsset>& scene) { return luzrt_compose_pass(rt, scene->exposure); }
int DeferredRenderer::SwapLightHistory() { return luzrt_swap_light_history(rt); }

} // namespace luzhost

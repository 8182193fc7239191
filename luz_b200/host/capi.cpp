// capi.cpp -- flat C interface over the host mirror, for harnesses that cannot include C++ headers
// (the Python tests / bench.py).  A C++ Luz host uses gpu_scene.hpp / scene.hpp directly.
// The frame loop is main.cpp's RenderFrame (source/Core/main.cpp:223-311) restricted to this path.
#include <cstdio>
#include <cstring>
#include <string>

#include "gpu_scene.hpp"
#include "import.hpp"

using namespace luzhost;

#define LUZHOST_API extern "C" __attribute__((visibility("default")))

struct luzhost_app {
    luzrt_ctx* rt = nullptr;
    AssetManager assets;
    Ref<SceneAsset> scene;
    Ref<CameraNode> camera;
    GPUScene* gpuScene = nullptr;
    DeferredRenderer* renderer = nullptr;
    int frameCount = 0;
    std::string error;
    std::vector<Ref<MeshNode>> meshNodes;
    std::vector<Ref<LightNode>> lightNodes;
    std::vector<Ref<MeshAsset>> meshAssets;
};

enum { LUZHOST_FRAME_OPAQUE = 1, LUZHOST_FRAME_COMPOSE = 2, LUZHOST_FRAME_TLAS_REFIT = 4, LUZHOST_FRAME_NO_UPDATE = 8 };

static int fail(luzhost_app* a, int code, const std::string& msg) {
    a->error = msg;
    return code;
}
static int rtfail(luzhost_app* a, int rc) {
    if (rc != LUZRT_OK) a->error = a->rt ? luzrt_last_error(a->rt) : "no device context (CPU-only host)";
    return rc;
}

// rt may be NULL: loading, transforms and UpdateResources (all CPU) still work, GPU calls fail.
LUZHOST_API luzhost_app* luzhost_create(luzrt_ctx* rt) {
    luzhost_app* a = new luzhost_app();
    a->rt = rt;
    a->gpuScene = new GPUScene(rt);
    a->renderer = new DeferredRenderer(rt);
    return a;
}
LUZHOST_API void luzhost_destroy(luzhost_app* a) {
    if (!a) return;
    delete a->gpuScene;
    delete a->renderer;
    delete a;
}
LUZHOST_API const char* luzhost_last_error(luzhost_app* a) { return a ? a->error.c_str() : "null app"; }

static void refresh_lists(luzhost_app* a) {
    a->meshNodes.clear();
    for (auto& n : a->scene->nodes) {
        if (n->type == ObjectType::MeshNode) a->meshNodes.emplace_back(std::dynamic_pointer_cast<MeshNode>(n));
        n->GetAll<MeshNode>(ObjectType::MeshNode, a->meshNodes);
    }
    a->lightNodes = a->scene->GetAll<LightNode>(ObjectType::LightNode);
    a->meshAssets = a->assets.GetAll<MeshAsset>(ObjectType::MeshAsset);
}

// == Setup(): AssetManager::LoadProject + GetInitialScene + GetMainCamera (main.cpp:74-93)
LUZHOST_API int luzhost_load_project(luzhost_app* a, const char* path, const char* bin_path) {
    if (!a->assets.LoadProject(path, bin_path)) return fail(a, -1, a->assets.error);
    a->scene = a->assets.GetInitialScene();
    if (!a->scene) return fail(a, -1, "initialScene not found");
    a->camera = a->assets.GetMainCamera(a->scene);
    refresh_lists(a);
    return 0;
}
// == AssetIO::Import (AssetIO.cpp:73-82) as the editor's drag-and-drop does it (Editor::AssetsPanel): the file's
// assets join the manager.  as_scene != 0 also makes the imported scene the one being rendered (its nodes are what
// GPUScene flattens); otherwise its top-level nodes are added to the current scene, like dropping a model into it.
LUZHOST_API int luzhost_import(luzhost_app* a, const char* path, int as_scene) {
    a->assets.error.clear();
    const UUID id = AssetIO::Import(path, a->assets);
    if (!id) return fail(a, -1, a->assets.error.empty() ? std::string("nothing to import from ") + path : a->assets.error);
    Ref<SceneAsset> imported = a->assets.Get<SceneAsset>(id);
    if (!imported) return 0; // a texture
    if (as_scene || !a->scene) {
        a->scene = imported;
        a->assets.initialScene = id;
        a->camera = a->assets.GetMainCamera(a->scene);
    } else {
        for (auto& n : imported->nodes) a->scene->Add(n);
    }
    refresh_lists(a);
    return 0;
}
// Test hook: import `path` into a fresh manager and write what was imported in the JSON layout of oracle/ref_import.cpp.
LUZHOST_API int luzhost_import_dump(const char* path, const char* out_json, char* err, uint32_t err_cap) {
    AssetManager m;
    const UUID id = AssetIO::Import(path, m);
    if (err && err_cap) snprintf(err, err_cap, "%s", m.error.c_str());
    if (!id) return -1;
    const std::string text = AssetIO::DumpImportedScene(m, id);
    FILE* f = fopen(out_json, "w");
    if (!f) return -2;
    fwrite(text.data(), 1, text.size(), f);
    fclose(f);
    return 0;
}
LUZHOST_API int luzhost_save_project(luzhost_app* a, const char* path, const char* bin_path) {
    return a->assets.SaveProject(path, bin_path) ? 0 : fail(a, -1, a->assets.error);
}
// == camera->extent = viewportSize + DeferredRenderer::CreateImages (main.cpp:93, :339-343)
LUZHOST_API int luzhost_set_extent(luzhost_app* a, uint32_t w, uint32_t h, int create_images) {
    if (!a->camera) return fail(a, -1, "no project loaded");
    a->camera->extent = lm::vec2{(float)w, (float)h};
    if (create_images) return a->rt ? rtfail(a, a->renderer->CreateImages(w, h)) : fail(a, LUZRT_E_NODEVICE, "no device context");
    return 0;
}
LUZHOST_API int luzhost_add_assets(luzhost_app* a) { return rtfail(a, a->gpuScene->AddAssets(a->assets)); }
LUZHOST_API int luzhost_update_resources(luzhost_app* a) {
    if (!a->scene) return fail(a, -1, "no project loaded");
    a->gpuScene->UpdateResources(a->scene, a->camera);
    return 0;
}
LUZHOST_API int luzhost_update_resources_gpu(luzhost_app* a, int tlas_mode) {
    return rtfail(a, a->gpuScene->UpdateResourcesGPU(tlas_mode));
}

// One RenderFrame + frame counter advance (main.cpp:223-311, :320-321).
LUZHOST_API int luzhost_render_frame(luzhost_app* a, uint32_t flags) {
    if (!a->scene) return fail(a, -1, "no project loaded");
    if (!a->rt) return fail(a, LUZRT_E_NODEVICE, "no device context: the lighting path has no CPU fallback");
    int rc;
    if (!(flags & LUZHOST_FRAME_NO_UPDATE)) {
        a->gpuScene->UpdateResources(a->scene, a->camera);
        if ((rc = a->gpuScene->UpdateResourcesGPU((flags & LUZHOST_FRAME_TLAS_REFIT) ? 1 : 0)) != LUZRT_OK) return rtfail(a, rc);
    }
    if (flags & LUZHOST_FRAME_OPAQUE)
        if ((rc = a->renderer->OpaquePass(*a->gpuScene)) != LUZRT_OK) return rtfail(a, rc);
    // main.cpp:260-264 renders every light's map every frame; luzrt renders the ones that get sampled
    if (a->scene->shadowType == ShadowMap || a->gpuScene->AnyShadowMapVolumetric())
        if ((rc = a->renderer->ShadowMapPass(a->scene)) != LUZRT_OK) return rtfail(a, rc);
    LightConstants lc;
    lc.frameID = a->frameCount;
    if ((rc = a->renderer->LightPass(lc)) != LUZRT_OK) return rtfail(a, rc);
    if (a->gpuScene->AnyVolumetricLight()) // main.cpp:274-279
        if ((rc = a->renderer->ScreenSpaceVolumetricLightPass(*a->gpuScene, a->frameCount)) != LUZRT_OK) return rtfail(a, rc);
    if ((rc = a->renderer->TAAPass(*a->gpuScene, a->scene)) != LUZRT_OK) return rtfail(a, rc);
    if (flags & LUZHOST_FRAME_COMPOSE)
        if ((rc = a->renderer->ComposePass(a->scene)) != LUZRT_OK) return rtfail(a, rc);
    if ((rc = a->renderer->SwapLightHistory()) != LUZRT_OK) return rtfail(a, rc);
    a->frameCount = (a->frameCount + 1) % (1 << 15);
    return 0;
}
LUZHOST_API int luzhost_frame_count(luzhost_app* a) { return a->frameCount; }
LUZHOST_API void luzhost_set_frame_count(luzhost_app* a, int f) { a->frameCount = f % (1 << 15); }

// ---- read access to what UpdateResources produced -----------------------------------------------
LUZHOST_API const luzw_scene_block* luzhost_scene_block(luzhost_app* a) { return &a->gpuScene->sceneBlock; }
LUZHOST_API const luzw_model_block* luzhost_models(luzhost_app* a, uint32_t* n) {
    *n = (uint32_t)a->gpuScene->modelsBlock.size();
    return a->gpuScene->modelsBlock.data();
}
LUZHOST_API const luzw_light_block* luzhost_extra_lights(luzhost_app* a, uint32_t* n) {
    *n = (uint32_t)a->gpuScene->extraLights.size();
    return a->gpuScene->extraLights.data();
}
LUZHOST_API const luzrt_instance* luzhost_instances(luzhost_app* a, uint32_t* n) {
    *n = (uint32_t)a->gpuScene->instances.size();
    return a->gpuScene->instances.data();
}
// index of the mesh asset (in luzhost_mesh order) each instance uses
LUZHOST_API int luzhost_instance_mesh(luzhost_app* a, uint32_t i) {
    if (i >= a->gpuScene->meshModels.size()) return -1;
    const auto& node = a->gpuScene->meshModels[i].node;
    for (size_t k = 0; k < a->meshAssets.size(); k++)
        if (a->meshAssets[k]->uuid == node->mesh->uuid) return (int)k;
    return -1;
}
LUZHOST_API uint32_t luzhost_mesh_count(luzhost_app* a) { return (uint32_t)a->meshAssets.size(); }
LUZHOST_API int luzhost_mesh(luzhost_app* a, uint32_t i, const void** verts, uint32_t* n_verts, const uint32_t** idx,
                             uint32_t* n_idx, uint64_t* uuid) {
    if (i >= a->meshAssets.size()) return -1;
    const auto& m = a->meshAssets[i];
    *verts = m->vertices.data();
    *n_verts = (uint32_t)m->vertices.size();
    *idx = m->indices.data();
    *n_idx = (uint32_t)m->indices.size();
    if (uuid) *uuid = m->uuid;
    return 0;
}
LUZHOST_API uint32_t luzhost_texture_count(luzhost_app* a) {
    return (uint32_t)a->assets.GetAll<TextureAsset>(ObjectType::TextureAsset).size();
}
// textures in RID order (the order AddAssets created them)
LUZHOST_API int luzhost_texture(luzhost_app* a, uint32_t i, const uint8_t** data, uint32_t* w, uint32_t* h) {
    auto all = a->assets.GetAll<TextureAsset>(ObjectType::TextureAsset);
    if (i >= all.size()) return -1;
    *data = all[i]->data.data();
    *w = (uint32_t)all[i]->width;
    *h = (uint32_t)all[i]->height;
    return 0;
}

// ---- scene edits (what the editor panels do) -----------------------------------------------------
LUZHOST_API uint32_t luzhost_mesh_node_count(luzhost_app* a) { return (uint32_t)a->meshNodes.size(); }
LUZHOST_API uint32_t luzhost_light_count(luzhost_app* a) { return (uint32_t)a->lightNodes.size(); }
LUZHOST_API int luzhost_mesh_node_set_transform(luzhost_app* a, uint32_t i, const float* pos, const float* rot,
                                                const float* scale) {
    if (i >= a->meshNodes.size()) return fail(a, -1, "mesh node index out of range");
    Node& n = *a->meshNodes[i];
    if (pos) n.position = lm::vec3(pos[0], pos[1], pos[2]);
    if (rot) n.rotation = lm::vec3(rot[0], rot[1], rot[2]);
    if (scale) n.scale = lm::vec3(scale[0], scale[1], scale[2]);
    return 0;
}
LUZHOST_API int luzhost_mesh_node_get_transform(luzhost_app* a, uint32_t i, float* pos, float* rot, float* scale) {
    if (i >= a->meshNodes.size()) return fail(a, -1, "mesh node index out of range");
    const Node& n = *a->meshNodes[i];
    for (int k = 0; k < 3; k++) {
        pos[k] = n.position[k];
        rot[k] = n.rotation[k];
        scale[k] = n.scale[k];
    }
    return 0;
}
// bulk variants for animated benchmark scenes: n x 3 floats each (NULL = leave)
LUZHOST_API int luzhost_mesh_nodes_set_transforms(luzhost_app* a, uint32_t first, uint32_t n, const float* pos,
                                                  const float* rot, const float* scale) {
    if ((size_t)first + n > a->meshNodes.size()) return fail(a, -1, "mesh node range out of bounds");
    for (uint32_t i = 0; i < n; i++) {
        Node& nd = *a->meshNodes[first + i];
        if (pos) nd.position = lm::vec3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
        if (rot) nd.rotation = lm::vec3(rot[3 * i], rot[3 * i + 1], rot[3 * i + 2]);
        if (scale) nd.scale = lm::vec3(scale[3 * i], scale[3 * i + 1], scale[3 * i + 2]);
    }
    return 0;
}
// Scene panel settings (Editor.cpp:314-350): pass a negative value to leave a field unchanged
LUZHOST_API int luzhost_scene_settings(luzhost_app* a, int light_samples, int ao_samples, int shadow_type,
                                       int taa_enabled, int taa_reconstruct) {
    if (!a->scene) return fail(a, -1, "no project loaded");
    if (light_samples >= 0) a->scene->lightSamples = light_samples;
    if (ao_samples >= 0) a->scene->aoSamples = ao_samples;
    if (shadow_type >= 0) a->scene->shadowType = shadow_type;
    if (taa_enabled >= 0) a->scene->taaEnabled = taa_enabled != 0;
    if (taa_reconstruct >= 0) a->scene->taaReconstruct = taa_reconstruct != 0;
    return 0;
}
LUZHOST_API int luzhost_scene_get_settings(luzhost_app* a, int* out5, float* out_ao_min_max_exposure_ambient) {
    if (!a->scene) return fail(a, -1, "no project loaded");
    out5[0] = a->scene->lightSamples;
    out5[1] = a->scene->aoSamples;
    out5[2] = a->scene->shadowType;
    out5[3] = a->scene->taaEnabled;
    out5[4] = a->scene->taaReconstruct;
    out_ao_min_max_exposure_ambient[0] = a->scene->aoMin;
    out_ao_min_max_exposure_ambient[1] = a->scene->aoMax;
    out_ao_min_max_exposure_ambient[2] = a->scene->exposure;
    out_ao_min_max_exposure_ambient[3] = a->scene->ambientLight;
    return 0;
}
LUZHOST_API int luzhost_camera_set_orbit(luzhost_app* a, const float* center, const float* rotation, float zoom) {
    if (!a->camera) return fail(a, -1, "no project loaded");
    if (center) a->camera->center = lm::vec3(center[0], center[1], center[2]);
    if (rotation) a->camera->camRotation = lm::vec3(rotation[0], rotation[1], rotation[2]);
    if (zoom > 0) a->camera->zoom = zoom;
    return 0;
}
LUZHOST_API int luzhost_camera_use_jitter(luzhost_app* a, int on) {
    if (!a->camera) return fail(a, -1, "no project loaded");
    a->camera->useJitter = on != 0;
    return 0;
}

// ---- pure functions exported for the host-parity tests (no state) -----------------------------------
LUZHOST_API float luzhost_halton(uint32_t i, uint32_t b) { return Halton(i, b); }
LUZHOST_API void luzhost_compose_transform(const float* pos, const float* rot, const float* scale,
                                           const float* parent16, float* out16) {
    lm::mat4 p(1.0f);
    if (parent16) memcpy(p.data(), parent16, 64);
    const lm::mat4 m = Node::ComposeTransform(lm::vec3(pos[0], pos[1], pos[2]), lm::vec3(rot[0], rot[1], rot[2]),
                                              lm::vec3(scale[0], scale[1], scale[2]), p);
    memcpy(out16, m.data(), 64);
}
// glm entry points the shadow-map matrices are built from (GPUScene.cpp:266-311)
LUZHOST_API void luzhost_perspective(float fovy, float aspect, float n, float f, float* out16) {
    const lm::mat4 m = lm::perspective(fovy, aspect, n, f);
    memcpy(out16, m.data(), 64);
}
LUZHOST_API void luzhost_ortho(float l, float r, float b, float t, float n, float f, float* out16) {
    const lm::mat4 m = lm::ortho(l, r, b, t, n, f);
    memcpy(out16, m.data(), 64);
}
LUZHOST_API void luzhost_look_at(const float* eye, const float* center, const float* up, float* out16) {
    const lm::mat4 m = lm::look_at(lm::vec3(eye[0], eye[1], eye[2]), lm::vec3(center[0], center[1], center[2]),
                                   lm::vec3(up[0], up[1], up[2]));
    memcpy(out16, m.data(), 64);
}
LUZHOST_API void luzhost_mat4_mul(const float* a16, const float* b16, float* out16) {
    lm::mat4 a(1.0f), b(1.0f);
    memcpy(a.data(), a16, 64);
    memcpy(b.data(), b16, 64);
    const lm::mat4 m = a * b;
    memcpy(out16, m.data(), 64);
}
LUZHOST_API int luzhost_camera_proj(luzhost_app* a, float n, float f, float* out16) {
    if (!a->camera) return fail(a, -1, "no camera");
    const lm::mat4 m = a->camera->GetProj(n, f);
    memcpy(out16, m.data(), 64);
    return 0;
}
LUZHOST_API void luzhost_mat4_inverse(const float* in16, float* out16) {
    lm::mat4 m(1.0f);
    memcpy(m.data(), in16, 64);
    const lm::mat4 r = lm::inverse(m);
    memcpy(out16, r.data(), 64);
}

// ref_dump.cpp -- drives the REFERENCE'S OWN compiled host code (source/Resources/*,
// source/Core/Util.*, glm) to produce golden vectors for everything on the lighting path that
// runs on the host: .luz/.luzbin loading, node world transforms, LightBlock inputs, the camera's
// view / projection / Halton jitter sequence, glm::inverse, and the byte layout of the wire
// structs in LuzCommon.h.  TEST INFRASTRUCTURE: built by oracle/Makefile into oracle/_ref/ from
// the sources where they lie under /root/reference; tests/golden/make_golden.py runs it and
// commits its JSON output.  It contains no reference code itself, it only calls it.
//
// usage: ref_dump <project.luz> <project.luzbin> <width> <height> <frames> <out.json>
// (the reference logger owns stdout and writes Luz.log into the cwd: run it from a scratch dir)
#include "AssetManager.hpp"
#include "LuzCommon.h"

#include <cstddef>
#include <cstdio>
#include <random>

static FILE* g_out = nullptr;
#define printf(...) fprintf(g_out, __VA_ARGS__)

static void put_floats(const char* name, const float* v, int n, bool comma = true) {
    printf("\"%s\":[", name);
    for (int i = 0; i < n; i++) printf("%s%.9g", i ? "," : "", v[i]);
    printf("]%s", comma ? "," : "");
}
static void put_mat(const char* name, const glm::mat4& m, bool comma = true) { put_floats(name, &m[0][0], 16, comma); }

int main(int argc, char** argv) {
    if (argc < 7) {
        fprintf(stderr, "usage: ref_dump project.luz project.luzbin width height frames out.json\n");
        return 2;
    }
    const float width = (float)atoi(argv[3]), height = (float)atoi(argv[4]);
    const int frames = atoi(argv[5]);
    g_out = fopen(argv[6], "w");
    if (!g_out) return 3;
    Logger::Init();
    AssetManager assets;
    assets.LoadProject(argv[1], argv[2]);
    Ref<SceneAsset> scene = assets.GetInitialScene();
    Ref<CameraNode> camera = assets.GetMainCamera(scene);
    camera->extent = {width, height}; // main.cpp:93

    printf("{");
    // ---- struct layouts (LuzCommon.h) -----------------------------------------------------
    printf("\"layout\":{");
    printf("\"LightBlock\":%zu,\"ModelBlock\":%zu,\"SceneBlock\":%zu,\"LightConstants\":%zu,"
           "\"PostProcessingConstants\":%zu,\"MeshVertex\":%zu,",
           sizeof(LightBlock), sizeof(ModelBlock), sizeof(SceneBlock), sizeof(LightConstants),
           sizeof(PostProcessingConstants), sizeof(MeshAsset::MeshVertex));
    printf("\"SceneBlock.ambientLightColor\":%zu,\"SceneBlock.proj\":%zu,\"SceneBlock.view\":%zu,"
           "\"SceneBlock.viewProj\":%zu,\"SceneBlock.prevViewProj\":%zu,\"SceneBlock.inverseProj\":%zu,"
           "\"SceneBlock.inverseView\":%zu,\"SceneBlock.jitter\":%zu,\"SceneBlock.prevJitter\":%zu,"
           "\"SceneBlock.camPos\":%zu,\"SceneBlock.numLights\":%zu,\"SceneBlock.aoMin\":%zu,\"SceneBlock.aoMax\":%zu,"
           "\"SceneBlock.exposure\":%zu,\"SceneBlock.aoNumSamples\":%zu,\"SceneBlock.blueNoiseTexture\":%zu,"
           "\"SceneBlock.tlasRid\":%zu,\"SceneBlock.shadowType\":%zu,",
           offsetof(SceneBlock, ambientLightColor), offsetof(SceneBlock, proj), offsetof(SceneBlock, view),
           offsetof(SceneBlock, viewProj), offsetof(SceneBlock, prevViewProj), offsetof(SceneBlock, inverseProj),
           offsetof(SceneBlock, inverseView), offsetof(SceneBlock, jitter), offsetof(SceneBlock, prevJitter),
           offsetof(SceneBlock, camPos), offsetof(SceneBlock, numLights), offsetof(SceneBlock, aoMin),
           offsetof(SceneBlock, aoMax), offsetof(SceneBlock, exposure), offsetof(SceneBlock, aoNumSamples),
           offsetof(SceneBlock, blueNoiseTexture), offsetof(SceneBlock, tlasRid), offsetof(SceneBlock, shadowType));
    printf("\"LightBlock.position\":%zu,\"LightBlock.direction\":%zu,\"LightBlock.type\":%zu,"
           "\"LightBlock.numShadowSamples\":%zu,\"LightBlock.radius\":%zu,\"LightBlock.viewProj\":%zu,"
           "\"LightBlock.zFar\":%zu,\"ModelBlock.color\":%zu,\"ModelBlock.roughness\":%zu,\"ModelBlock.colorMap\":%zu,"
           "\"PostProcessingConstants.size\":%zu,\"PostProcessingConstants.reconstruct\":%zu},",
           offsetof(LightBlock, position), offsetof(LightBlock, direction), offsetof(LightBlock, type),
           offsetof(LightBlock, numShadowSamples), offsetof(LightBlock, radius), offsetof(LightBlock, viewProj),
           offsetof(LightBlock, zFar), offsetof(ModelBlock, color), offsetof(ModelBlock, roughness),
           offsetof(ModelBlock, colorMap), offsetof(PostProcessingConstants, size),
           offsetof(PostProcessingConstants, reconstruct));

    // ---- scene settings -------------------------------------------------------------------
    printf("\"scene\":{\"name\":\"%s\",\"aoSamples\":%d,\"lightSamples\":%d,\"shadowType\":%d,\"taaEnabled\":%d,"
           "\"taaReconstruct\":%d,",
           scene->name.c_str(), scene->aoSamples, scene->lightSamples, (int)scene->shadowType,
           (int)scene->taaEnabled, (int)scene->taaReconstruct);
    {
        float v[4] = {scene->aoMin, scene->aoMax, scene->exposure, scene->ambientLight};
        put_floats("aoMin_aoMax_exposure_ambientLight", v, 4);
        put_floats("ambientLightColor", &scene->ambientLightColor.x, 3, false);
    }
    printf("},");

    // ---- mesh nodes in GPUScene::UpdateResources order (GPUScene.cpp:183-184) ------------------
    std::vector<Ref<MeshNode>> meshNodes;
    scene->GetAll<MeshNode>(ObjectType::MeshNode, meshNodes);
    printf("\"meshNodes\":[");
    for (size_t i = 0; i < meshNodes.size(); i++) {
        auto& n = meshNodes[i];
        printf("%s{\"name\":\"%s\",\"meshUuid\":%llu,\"vertexCount\":%zu,\"indexCount\":%zu,", i ? "," : "",
               n->name.c_str(), (unsigned long long)n->mesh->uuid, n->mesh->vertices.size(), n->mesh->indices.size());
        if (n->material) {
            put_floats("color", &n->material->color.x, 4);
            put_floats("emission", &n->material->emission.x, 3);
            float mr[2] = {n->material->metallic, n->material->roughness};
            put_floats("metallic_roughness", mr, 2);
            printf("\"colorMapUuid\":%llu,",
                   (unsigned long long)(n->material->colorMap ? n->material->colorMap->uuid : 0));
        }
        // a few raw vertex floats + indices to pin the blob decoding
        put_floats("vertex0", (const float*)&n->mesh->vertices[0], 12);
        printf("\"indices\":[");
        for (size_t k = 0; k < n->mesh->indices.size(); k++) printf("%s%u", k ? "," : "", n->mesh->indices[k]);
        printf("],");
        put_mat("world", n->GetWorldTransform(), false);
        printf("}");
    }
    printf("],");

    // ---- lights: the inputs GPUScene.cpp:239-251 writes into LightBlock --------------------------
    printf("\"lights\":[");
    {
        auto lights = scene->GetAll<LightNode>(ObjectType::LightNode);
        for (size_t i = 0; i < lights.size(); i++) {
            auto& l = lights[i];
            glm::vec3 pos = l->GetWorldPosition();
            glm::vec3 dir = l->GetWorldTransform() * glm::vec4(0, -1, 0, 0);
            float ang[2] = {glm::radians(l->innerAngle), glm::radians(l->outerAngle)};
            float misc[3] = {l->intensity, l->radius, l->shadowMapFar};
            printf("%s{\"type\":%d,\"volumetricType\":%d,", i ? "," : "", (int)l->lightType, (int)l->volumetricType);
            put_floats("color", &l->color.x, 3);
            put_floats("position", &pos.x, 3);
            put_floats("direction", &dir.x, 3);
            put_floats("inner_outer_radians", ang, 2);
            put_floats("intensity_radius_zfar", misc, 3, false);
            printf("}");
        }
    }
    printf("],");

    // ---- camera: the per-frame sequence of GPUScene.cpp:225-234 ------------------------------------
    printf("\"camera\":{");
    {
        float c[6] = {camera->zoom, camera->farDistance, camera->nearDistance, camera->horizontalFov, width, height};
        put_floats("zoom_far_near_fov_w_h", c, 6);
        put_floats("center", &camera->center.x, 3);
        put_floats("rotation", &camera->rotation.x, 3);
        printf("\"mode\":%d,\"cameraType\":%d,", (int)camera->mode, (int)camera->cameraType);
    }
    printf("\"frames\":[");
    for (int f = 0; f < frames; f++) {
        glm::vec2 prevJitter = camera->GetJitter();
        camera->NextJitter();
        glm::vec2 jitter = camera->GetJitter();
        glm::mat4 proj = camera->GetProjJittered();
        glm::mat4 view = camera->GetView();
        glm::mat4 viewProj = camera->GetProjJittered() * camera->GetView();
        glm::mat4 invProj = glm::inverse(camera->GetProjJittered());
        glm::mat4 invView = glm::inverse(camera->GetView());
        printf("%s{", f ? "," : "");
        put_floats("prevJitter", &prevJitter.x, 2);
        put_floats("jitter", &jitter.x, 2);
        put_floats("camPos", &camera->eye.x, 3);
        put_mat("proj", proj);
        put_mat("view", view);
        put_mat("viewProj", viewProj);
        put_mat("inverseProj", invProj);
        put_mat("inverseView", invView, false);
        printf("}");
    }
    printf("]},");

    // ---- Halton (Util.hpp:23-34) -----------------------------------------------------------
    {
        float h2[32], h3[32];
        for (int i = 0; i < 32; i++) {
            h2[i] = Halton(i, 2);
            h3[i] = Halton(i, 3);
        }
        put_floats("halton2", h2, 32);
        put_floats("halton3", h3, 32);
    }

    // ---- ComposeTransform / glm::inverse on seeded inputs (AssetManager.cpp:38-43) ----------------
    printf("\"transforms\":[");
    {
        std::mt19937 rng(20240229u);
        std::uniform_real_distribution<float> up(-50.0f, 50.0f), ur(-360.0f, 360.0f), us(0.01f, 9.0f);
        for (int i = 0; i < 24; i++) {
            glm::vec3 p(up(rng), up(rng), up(rng)), r(ur(rng), ur(rng), ur(rng)), s(us(rng), us(rng), us(rng));
            if (i == 0) {
                p = glm::vec3(0);
                r = glm::vec3(0);
                s = glm::vec3(1);
            }
            glm::mat4 parent = (i % 3 == 2) ? Node::ComposeTransform(glm::vec3(1, 2, 3), glm::vec3(10, 20, 30), glm::vec3(2, 2, 2))
                                            : glm::mat4(1);
            glm::mat4 m = Node::ComposeTransform(p, r, s, parent);
            glm::mat4 inv = glm::inverse(m);
            printf("%s{", i ? "," : "");
            put_floats("pos", &p.x, 3);
            put_floats("rot", &r.x, 3);
            put_floats("scale", &s.x, 3);
            put_mat("parent", parent);
            put_mat("mat", m);
            put_mat("inverse", inv, false);
            printf("}");
        }
    }
    printf("],");

    // ---- the glm / camera entry points GPUScene.cpp:266-311 builds light.viewProj[] from, called on seeded
    //      inputs: perspective(90 deg, 1, 0, far) (zNear = 0), lookAt with the six cube-face axes / ups, ortho with
    //      zNear > zFar, CameraNode::GetProj(near, far / range), inverse and products of those ------------------
    printf("\"shadowGlm\":{");
    {
        put_mat("persp90_far2000", glm::perspective(glm::radians(90.0f), 1.0f, 0.0f, 2000.0f));
        put_mat("persp90_far37", glm::perspective(glm::radians(90.0f), 1.0f, 0.0f, 37.5f));
        const glm::vec3 pos(3.139984130859375f, 6.1400675773620605f, -3.6627626419067383f);
        put_floats("pos", &pos.x, 3);
        const glm::vec3 axis[6] = {{1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};
        const glm::vec3 up[6] = {{0, -1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}, {0, -1, 0}, {0, -1, 0}};
        printf("\"faces\":[");
        for (int f = 0; f < 6; f++) {
            glm::mat4 v = glm::lookAt(pos, pos + axis[f], up[f]);
            glm::mat4 vp = glm::perspective(glm::radians(90.0f), 1.0f, 0.0f, 2000.0f) * v;
            printf("%s{", f ? "," : "");
            put_mat("lookAt", v);
            put_mat("viewProj", vp, false);
            printf("}");
        }
        printf("],");
        put_mat("ortho", glm::ortho(-3.25f, 5.5f, -2.125f, 7.75f, 11.5f, -9.25f));
        glm::mat4 cp = camera->GetProj(camera->nearDistance, camera->farDistance / 3.0f);
        put_mat("camProj_far_over_3", cp);
        glm::mat4 cv = camera->GetView();
        put_mat("camView", cv);
        put_mat("inverse_camProjView", glm::inverse(cp * cv));
        const glm::vec3 centre(0.5f, -1.25f, 2.0f), front(0.3f, -1.0f, 0.2f);
        put_mat("lookAt_front", glm::lookAt(centre + front, centre, glm::vec3(.0f, 1.0f, .0f)), false);
    }
    printf("}}\n");
    fclose(g_out);
    return 0;
}

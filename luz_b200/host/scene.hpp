// scene.hpp -- Luz's scene data model and .luz/.luzbin project format, as the lighting path needs it.
//
// Mirrors (same names, fields, defaults and (de)serialisation keys) the reference's
// source/Resources/AssetManager.hpp:14-353 and Serializer.hpp:65-158, written from scratch on top of
// lmath.hpp / ljson.hpp instead of glm / nlohmann.  Scene import (glTF / GLB / OBJ / PNG) is import.hpp; everything
// editor-related (ImGui, cloning) is out of scope.
#pragma once

#include <cstdint>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "ljson.hpp"
#include "lmath.hpp"

namespace luzhost {

using UUID = uint64_t;
template <class T>
using Ref = std::shared_ptr<T>;

enum class ObjectType { // AssetManager.hpp:14-25
    Invalid,
    TextureAsset,
    MeshAsset,
    MaterialAsset,
    SceneAsset,
    Node,
    MeshNode,
    LightNode,
    CameraNode,
    Count
};

enum ShadowType { ShadowDisabled = 0, ShadowRayTraced = 1, ShadowMap = 2 }; // AssetManager.hpp:40-45

struct AssetManager;
struct SceneAsset;

// what Serializer.hpp carries around: the JSON node being read/written + the binary blob
struct Serializer {
    lj::Value& j;
    std::vector<uint8_t>& blob;
    AssetManager& manager;
    bool saving;
};

struct Object { // AssetManager.hpp:49-66
    std::string name = "Unintialized";
    UUID uuid = 0;
    ObjectType type = ObjectType::Invalid;
    bool gpuDirty = true;
    virtual ~Object() = default;
    virtual void Serialize(Serializer& s) = 0;
};
struct Asset : Object {};

struct TextureAsset : Asset { // :73-81
    std::vector<uint8_t> data;
    int channels = 0, width = 0, height = 0;
    TextureAsset() { type = ObjectType::TextureAsset; }
    void Serialize(Serializer& s) override;
};

struct MeshAsset : Asset { // :83-99
    struct MeshVertex {
        lm::vec3 position, normal;
        lm::vec4 tangent;
        lm::vec2 texCoord;
    };
    std::vector<MeshVertex> vertices;
    std::vector<uint32_t> indices;
    MeshAsset() { type = ObjectType::MeshAsset; }
    void Serialize(Serializer& s) override;
};
static_assert(sizeof(MeshAsset::MeshVertex) == 48, "MeshVertex is 48 bytes");

struct MaterialAsset : Asset { // :101-114
    lm::vec4 color{1, 1, 1, 1};
    lm::vec3 emission{0, 0, 0};
    float metallic = 0, roughness = 0.5f;
    Ref<TextureAsset> aoMap, colorMap, normalMap, emissionMap, metallicRoughnessMap;
    MaterialAsset() { type = ObjectType::MaterialAsset; }
    void Serialize(Serializer& s) override;
};

struct Node : Object, std::enable_shared_from_this<Node> { // :116-176
    Node* parent = nullptr; // the reference holds a Ref; a raw back-pointer avoids the cycle
    std::vector<Ref<Node>> children;
    lm::vec3 position{0, 0, 0}, rotation{0, 0, 0}, scale{1, 1, 1};
    Node() { type = ObjectType::Node; }
    void Serialize(Serializer& s) override;
    void SerializeNodeFields(Serializer& s);

    template <class T>
    void GetAll(ObjectType t, std::vector<Ref<T>>& all) {
        for (auto& n : children) {
            if (n->type == t) all.emplace_back(std::dynamic_pointer_cast<T>(n));
            n->GetAll(t, all);
        }
    }
    lm::mat4 GetLocalTransform() const;
    lm::mat4 GetWorldTransform() const;
    lm::mat4 GetParentTransform() const;
    lm::vec3 GetWorldPosition() const;
    lm::vec3 GetWorldFront() const;
    static lm::mat4 ComposeTransform(lm::vec3 pos, lm::vec3 rot, lm::vec3 scl, const lm::mat4& parent = lm::mat4(1.0f));
    static void SetParent(const Ref<Node>& child, const Ref<Node>& parent);
};

struct MeshNode : Node { // :178-184
    Ref<MeshAsset> mesh;
    Ref<MaterialAsset> material;
    MeshNode() { type = ObjectType::MeshNode; }
    void Serialize(Serializer& s) override;
};

struct LightNode : Node { // :186-227
    enum LightType { Point = 0, Spot = 1, Directional = 2 };
    enum VolumetricType { Disabled = 0, ScreenSpace = 1, ShadowMapVolumetric = 2 };
    lm::vec3 color{1, 1, 1};
    float intensity = 10.0f;
    int lightType = Point;
    float radius = 2.0f, innerAngle = 60.0f, outerAngle = 50.0f;
    float shadowMapRange = 3.0f, shadowMapFar = 2000.0f;
    float volumetricScreenAbsorption = 0.5f;
    int volumetricScreenSamples = 128;
    float volumetricShadowWeight = 0.0001f, volumetricShadowAbsorption = 1.0f, volumetricShadowDensity = 1.094f;
    int volumetricShadowSamples = 128;
    int volumetricType = ScreenSpace;
    LightNode() { type = ObjectType::LightNode; }
    void Serialize(Serializer& s) override;
};

struct CameraNode : Node { // :229-275
    enum CameraMode { Orbit, Fly };
    enum CameraType { Perspective, Orthographic };
    int cameraType = Perspective, mode = Orbit;
    lm::vec3 eye{0, 0, 0}, center{0, 0, 0}, camRotation{0, 0, 0}; // CameraNode::rotation shadows Node::rotation
    bool useJitter = true;
    float zoom = 10.0f, farDistance = 1000.0f, nearDistance = 0.01f, horizontalFov = 60.0f;
    float orthoFarDistance = 10.0f, orthoNearDistance = -100.0f;
    lm::vec2 extent{1.0f, 1.0f};
    CameraNode() { type = ObjectType::CameraNode; }
    void Serialize(Serializer& s) override;

    lm::mat4 GetView();
    lm::mat4 GetProj() const { return GetProj(nearDistance, farDistance); }
    lm::mat4 GetProjJittered() const;
    lm::mat4 GetProj(float zNear, float zFar) const;
    lm::vec2 GetJitter() const { return jitter; }
    void NextJitter();

private:
    lm::vec2 jitter{0, 0};
    uint32_t jitterIndex = 0;
};

struct SceneAsset : Asset { // :277-351
    std::vector<Ref<Node>> nodes;
    lm::vec3 ambientLightColor{1, 1, 1};
    float ambientLight = 0.01f;
    int aoSamples = 4, lightSamples = 2;
    float aoMin = 0.0001f, aoMax = 1.0f, exposure = 2.0f;
    int shadowType = ShadowRayTraced;
    uint32_t shadowResolution = 1024;
    Ref<CameraNode> mainCamera;
    bool taaEnabled = true, taaReconstruct = true;
    SceneAsset() { type = ObjectType::SceneAsset; }
    void Serialize(Serializer& s) override;

    void Add(const Ref<Node>& n) { nodes.push_back(n); }
    template <class T>
    Ref<T> Get(UUID id) { // top-level nodes only, like the reference (:313-321)
        for (auto& n : nodes)
            if (n->uuid == id) return std::dynamic_pointer_cast<T>(n);
        return {};
    }
    template <class T>
    std::vector<Ref<T>> GetAll(ObjectType t) {
        std::vector<Ref<T>> all;
        for (auto& n : nodes) {
            if (n->type == t) all.emplace_back(std::dynamic_pointer_cast<T>(n));
            n->GetAll(t, all);
        }
        return all;
    }
    void UpdateParents();
};

float Halton(uint32_t i, uint32_t b); // source/Core/Util.hpp:23-34

struct AssetManager { // AssetManager.hpp:353-
    std::unordered_map<UUID, Ref<Asset>> assets;
    std::vector<UUID> load_order;
    UUID initialScene = 0;
    std::string error; // the reference logs and carries on; the mirror records the message

    Ref<Object> CreateObject(ObjectType type, const std::string& name, UUID uuid);
    template <class T>
    Ref<T> Get(UUID id) {
        auto it = assets.find(id);
        return it == assets.end() ? Ref<T>() : std::dynamic_pointer_cast<T>(it->second);
    }
    template <class T>
    std::vector<Ref<T>> GetAll(ObjectType t) const {
        std::vector<Ref<T>> all;
        for (UUID u : load_order) {
            auto it = assets.find(u);
            if (it != assets.end() && it->second->type == t) all.emplace_back(std::dynamic_pointer_cast<T>(it->second));
        }
        return all;
    }
    // AssetManager.cpp:182-214 / :216-257.  Return false (and set `error`) on I/O or parse failure.
    bool LoadProject(const std::string& path, const std::string& binPath);
    bool SaveProject(const std::string& path, const std::string& binPath);
    Ref<SceneAsset> GetInitialScene();
    Ref<CameraNode> GetMainCamera(const Ref<SceneAsset>& scene);
    UUID NewUUID();

private:
    uint64_t uuid_state = 0x9E3779B97F4A7C15ull;
};

} // namespace luzhost

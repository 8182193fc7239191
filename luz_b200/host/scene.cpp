// scene.cpp -- scene data model, transforms, camera and the .luz/.luzbin reader/writer.
// Reference: source/Resources/AssetManager.cpp (:38-63 transforms, :73-144 + :318-332 Serialize
// methods, :182-257 Load/SaveProject, :334-392 camera), Serializer.hpp:65-158, Util.hpp:23-34.
#include "scene.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>

namespace luzhost {

using lm::mat4;
using lm::vec2;
using lm::vec3;
using lm::vec4;

// ---- Util.hpp:23-34 ---------------------------------------------------------------------------
float Halton(uint32_t i, uint32_t b) {
    float f = 1.0f, r = 0.0f;
    while (i > 0) {
        f /= static_cast<float>(b);
        r = r + f * static_cast<float>(i % b);
        i = static_cast<uint32_t>(floorf(static_cast<float>(i) / static_cast<float>(b)));
    }
    return r;
}

// ---- field (de)serialisation helpers (Serializer.hpp:84-158) -------------------------------------
namespace {

void io(Serializer& s, const char* k, float& v) {
    if (s.saving) s.j[k] = lj::Value::number((double)v);
    else if (s.j.contains(k) && s.j.at(k).is_number()) v = (float)s.j.at(k).as_double();
}
void io(Serializer& s, const char* k, int& v) {
    if (s.saving) s.j[k] = lj::Value::integer(v);
    else if (s.j.contains(k)) v = (int)s.j.at(k).as_int();
}
void io(Serializer& s, const char* k, uint32_t& v) {
    if (s.saving) s.j[k] = lj::Value::uinteger(v);
    else if (s.j.contains(k)) v = (uint32_t)s.j.at(k).as_uint();
}
void io(Serializer& s, const char* k, bool& v) {
    if (s.saving) s.j[k] = lj::Value::boolean(v);
    else if (s.j.contains(k)) v = s.j.at(k).as_bool();
}
void io(Serializer& s, const char* k, vec3& v) {
    if (s.saving) {
        lj::Value a = lj::Value::array();
        for (int i = 0; i < 3; i++) a.a->push_back(lj::Value::number((double)v[i]));
        s.j[k] = a;
    } else if (s.j.contains(k)) {
        const lj::Value& a = s.j.at(k);
        if (a.is_array() && a.size() == 3) // Serializer.hpp:26-32: other shapes leave the default
            for (int i = 0; i < 3; i++) v[i] = (float)(*a.a)[i].as_double();
    }
}
void io(Serializer& s, const char* k, vec4& v) {
    if (s.saving) {
        lj::Value a = lj::Value::array();
        for (int i = 0; i < 4; i++) a.a->push_back(lj::Value::number((double)v[i]));
        s.j[k] = a;
    } else if (s.j.contains(k)) {
        const lj::Value& a = s.j.at(k);
        if (a.is_array() && a.size() == 4)
            for (int i = 0; i < 4; i++) v[i] = (float)(*a.a)[i].as_double();
    }
}

// Serializer::Vector (Serializer.hpp:93-109): {offset,size} byte range of the .luzbin blob
template <class T>
void io_blob(Serializer& s, const char* k, std::vector<T>& v) {
    if (s.saving) {
        const uint32_t size = (uint32_t)(v.size() * sizeof(T));
        const uint32_t offset = (uint32_t)s.blob.size();
        s.blob.resize(s.blob.size() + size);
        if (size) memcpy(s.blob.data() + offset, v.data(), size);
        lj::Value o = lj::Value::object();
        o["offset"] = lj::Value::uinteger(offset);
        o["size"] = lj::Value::uinteger(size);
        s.j[k] = o;
    } else if (s.j.contains(k)) {
        const lj::Value& o = s.j.at(k);
        const uint64_t size = o.at("size").as_uint(), offset = o.at("offset").as_uint();
        if (offset + size > s.blob.size()) { // the reference would read out of bounds; fail softly
            s.manager.error = std::string("blob range out of bounds for field ") + k;
            v.clear();
            return;
        }
        v.resize(size / sizeof(T));
        if (size) memcpy(v.data(), s.blob.data() + offset, v.size() * sizeof(T));
    }
}

// Serializer::Asset (Serializer.hpp:133-144): uuid reference, 0 = none
template <class T>
void io_asset(Serializer& s, const char* k, Ref<T>& obj) {
    if (s.saving) s.j[k] = lj::Value::uinteger(obj ? obj->uuid : 0);
    else if (s.j.contains(k) && s.j.at(k).as_uint() != 0) obj = s.manager.Get<T>(s.j.at(k).as_uint());
}

// Serializer::Serialize(Ref<T>&) (Serializer.hpp:65-82)
Ref<Object> load_object(lj::Value& j, std::vector<uint8_t>& blob, AssetManager& m) {
    if (!(j.contains("type") && j.contains("name") && j.contains("uuid"))) {
        m.error = "object without type/name/uuid";
        return {};
    }
    const ObjectType type = (ObjectType)j.at("type").as_int();
    Ref<Object> obj = m.CreateObject(type, j.at("name").s, j.at("uuid").as_uint());
    if (!obj) return {};
    Serializer s{j, blob, m, false};
    obj->Serialize(s);
    return obj;
}
void save_object(const Ref<Object>& obj, lj::Value& j, std::vector<uint8_t>& blob, AssetManager& m) {
    if (!j.is_object()) j = lj::Value::object();
    j["type"] = lj::Value::integer((int)obj->type);
    j["name"] = lj::Value::string(obj->name);
    j["uuid"] = lj::Value::uinteger(obj->uuid);
    Serializer s{j, blob, m, true};
    obj->Serialize(s);
}

// Serializer::VectorRef (Serializer.hpp:111-131)
void io_children(Serializer& s, const char* k, std::vector<Ref<Node>>& v) {
    if (s.saving) {
        lj::Value arr = lj::Value::array();
        for (auto& n : v) {
            arr.a->emplace_back();
            save_object(n, arr.a->back(), s.blob, s.manager);
        }
        s.j[k] = arr;
    } else if (s.j.contains(k) && s.j.at(k).is_array()) {
        for (lj::Value& cj : *s.j.o->at(k).a) {
            Ref<Node> child = std::dynamic_pointer_cast<Node>(load_object(cj, s.blob, s.manager));
            if (child) v.push_back(child);
        }
    }
}

bool read_file(const std::string& path, std::string& out) {
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    std::ostringstream ss;
    ss << f.rdbuf();
    out = ss.str();
    return true;
}

} // namespace

// ---- Serialize methods: same keys as the reference ---------------------------------------------
void TextureAsset::Serialize(Serializer& s) { // AssetManager.cpp:73-78
    io_blob(s, "data", data);
    io(s, "width", width);
    io(s, "height", height);
    io(s, "channels", channels);
}
void MeshAsset::Serialize(Serializer& s) { // :80-83
    io_blob(s, "vertices", vertices);
    io_blob(s, "indices", indices);
}
void MaterialAsset::Serialize(Serializer& s) { // :85-95
    io(s, "color", color);
    io(s, "emission", emission);
    io(s, "metallic", metallic);
    io(s, "roughness", roughness);
    io_asset(s, "colorMap", colorMap);
    io_asset(s, "aoMap", aoMap);
    io_asset(s, "emissionMap", emissionMap);
    io_asset(s, "normalMap", normalMap);
    io_asset(s, "metallicRoughnessMap", metallicRoughnessMap);
}
void SceneAsset::Serialize(Serializer& s) { // :97-110
    io_children(s, "nodes", nodes);
    io(s, "ambientLight", ambientLight);
    io(s, "ambientLightColor", ambientLightColor);
    io(s, "lightSamples", lightSamples);
    io(s, "aoSamples", aoSamples);
    io(s, "aoMin", aoMin);
    io(s, "aoMax", aoMax);
    io(s, "exposure", exposure);
    io(s, "shadowType", shadowType);
    io(s, "taaEnabled", taaEnabled);
    io(s, "taaReconstruct", taaReconstruct);
    // Serializer::Node (Serializer.hpp:146-157): main camera by uuid among the top-level nodes
    if (s.saving) s.j["mainCamera"] = lj::Value::uinteger(mainCamera ? mainCamera->uuid : 0);
    else if (s.j.contains("mainCamera") && s.j.at("mainCamera").as_uint() != 0)
        mainCamera = Get<CameraNode>(s.j.at("mainCamera").as_uint());
}
void Node::SerializeNodeFields(Serializer& s) { // :112-117
    io_children(s, "children", children);
    io(s, "position", position);
    io(s, "rotation", rotation);
    io(s, "scale", scale);
}
void Node::Serialize(Serializer& s) { SerializeNodeFields(s); }
void MeshNode::Serialize(Serializer& s) { // :119-123
    SerializeNodeFields(s);
    io_asset(s, "mesh", mesh);
    io_asset(s, "material", material);
}
void LightNode::Serialize(Serializer& s) { // :125-144
    SerializeNodeFields(s);
    io(s, "color", color);
    io(s, "intensity", intensity);
    io(s, "lightType", lightType);
    io(s, "innerAngle", innerAngle);
    io(s, "outerAngle", outerAngle);
    io(s, "radius", radius);
    io(s, "shadowMapRange", shadowMapRange);
    io(s, "shadowMapFar", shadowMapFar);
    io(s, "volumetricType", volumetricType);
    io(s, "volumetricScreenAbsorption", volumetricScreenAbsorption);
    io(s, "volumetricScreenSamples", volumetricScreenSamples);
    io(s, "volumetricShadowWeight", volumetricShadowWeight);
    io(s, "volumetricShadowAbsorption", volumetricShadowAbsorption);
    io(s, "volumetricShadowDensity", volumetricShadowDensity);
    io(s, "volumetricShadowSamples", volumetricShadowSamples);
}
void CameraNode::Serialize(Serializer& s) { // :318-332
    SerializeNodeFields(s);
    io(s, "cameraType", cameraType);
    io(s, "mode", mode);
    io(s, "eye", eye);
    io(s, "center", center);
    // CameraNode::rotation shadows Node::rotation; both are (de)serialised under "rotation".  On save
    // the camera's value is written last, so it wins, exactly as in the reference.
    io(s, "rotation", camRotation);
    io(s, "zoom", zoom);
    io(s, "farDistance", farDistance);
    io(s, "nearDistance", nearDistance);
    io(s, "horizontalFov", horizontalFov);
    io(s, "orthoFarDistance", orthoFarDistance);
    io(s, "orthoNearDistance", orthoNearDistance);
}

// ---- transforms (AssetManager.cpp:38-63) -------------------------------------------------------
mat4 Node::ComposeTransform(vec3 pos, vec3 rot, vec3 scl, const mat4& parent) {
    const mat4 rotationMat = lm::mat4_cast(lm::quat_from_euler(lm::radians(rot)));
    const mat4 translationMat = lm::translate(pos);
    const mat4 scaleMat = lm::scale(scl);
    return parent * (translationMat * rotationMat * scaleMat);
}
mat4 Node::GetLocalTransform() const { return ComposeTransform(position, rotation, scale); }
mat4 Node::GetParentTransform() const { return parent ? parent->GetWorldTransform() : mat4(1.0f); }
mat4 Node::GetWorldTransform() const { return GetParentTransform() * GetLocalTransform(); }
vec3 Node::GetWorldPosition() const {
    const vec4 p = GetParentTransform() * vec4(position, 1.0f);
    return {p.x, p.y, p.z};
}
vec3 Node::GetWorldFront() const {
    const vec4 p = GetWorldTransform() * vec4(0, -1, 0, 0);
    return {p.x, p.y, p.z};
}
void Node::SetParent(const Ref<Node>& child, const Ref<Node>& parent) {
    if (child->parent) {
        auto& sib = child->parent->children;
        sib.erase(std::remove_if(sib.begin(), sib.end(), [&](const Ref<Node>& n) { return n->uuid == child->uuid; }),
                  sib.end());
    }
    child->parent = parent.get();
    parent->children.push_back(child);
}
static void update_children_parent(Node* n) {
    for (auto& c : n->children) {
        c->parent = n;
        update_children_parent(c.get());
    }
}
void SceneAsset::UpdateParents() { // AssetManager.hpp:343-348
    for (auto& n : nodes) {
        n->parent = nullptr;
        update_children_parent(n.get());
    }
}

// ---- camera (AssetManager.cpp:334-392) ---------------------------------------------------------
mat4 CameraNode::GetView() {
    if (mode == Orbit) {
        const vec3 rads = lm::radians(camRotation + vec3(90.0f, 90.0f, 0.0f));
        vec3 viewDir;
        viewDir.x = std::cos(-rads.y) * std::sin(rads.x);
        viewDir.z = std::sin(-rads.y) * std::sin(rads.x);
        viewDir.y = std::cos(rads.x);
        if (cameraType == Perspective) viewDir = viewDir * zoom;
        eye = center - viewDir;
        return lm::look_at(eye, center, vec3(0.0f, 1.0f, 0.0f));
    }
    const vec3 rads = lm::radians(camRotation + vec3(0.0f, 180.0f, 0.0f));
    mat4 rot(1.0f);
    rot = lm::rotate(rot, rads.z, vec3(0.0f, 0.0f, 1.0f));
    rot = lm::rotate(rot, rads.y, vec3(0.0f, 1.0f, 0.0f));
    rot = lm::rotate(rot, rads.x, vec3(1.0f, 0.0f, 0.0f));
    return lm::look_at(eye, eye + vec3(rot[2].x, rot[2].y, rot[2].z), vec3(0.0f, 1.0f, 0.0f));
}
mat4 CameraNode::GetProjJittered() const { return lm::translate(vec3(jitter.x, jitter.y, 0.0f)) * GetProj(); }
mat4 CameraNode::GetProj(float zNear, float zFar) const {
    mat4 proj;
    if (cameraType == Perspective) {
        // the value called horizontalFov is passed as glm's fovy
        proj = lm::perspective(lm::radians(horizontalFov), extent.x / extent.y, zNear, zFar);
    } else {
        const float sx = 1.0f * zoom, sy = (extent.y / extent.x) * zoom;
        proj = lm::ortho(-sx, sx, -sy, sy, orthoNearDistance, orthoFarDistance);
    }
    proj[1][1] *= -1.0f;
    return proj;
}
void CameraNode::NextJitter() {
    if (useJitter) {
        jitterIndex = (jitterIndex + 1) % 16;
        jitter = vec2{Halton(jitterIndex + 1, 2), Halton(jitterIndex + 1, 3)};
        jitter.x = 2.0f * jitter.x - 1.0f;
        jitter.y = 2.0f * jitter.y - 1.0f;
        jitter.x /= extent.x;
        jitter.y /= extent.y;
    } else {
        jitter = vec2{0, 0};
    }
}

// ---- AssetManager ---------------------------------------------------------------------------------
UUID AssetManager::NewUUID() { // AssetManager.cpp:268-274: uniform in [2^61, 2^62]; here splitmix64
    uuid_state += 0x9E3779B97F4A7C15ull;
    uint64_t z = uuid_state;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (1ull << 61) + (z % (1ull << 61));
}

template <class T>
static Ref<T> make_named(const std::string& name, UUID uuid) {
    Ref<T> a = std::make_shared<T>();
    a->name = name;
    a->uuid = uuid;
    return a;
}

Ref<Object> AssetManager::CreateObject(ObjectType type, const std::string& name, UUID uuid) {
    if (uuid == 0) uuid = NewUUID();
    Ref<Asset> asset;
    switch (type) {
        case ObjectType::TextureAsset: asset = make_named<TextureAsset>(name, uuid); break;
        case ObjectType::MaterialAsset: asset = make_named<MaterialAsset>(name, uuid); break;
        case ObjectType::MeshAsset: asset = make_named<MeshAsset>(name, uuid); break;
        case ObjectType::SceneAsset: asset = make_named<SceneAsset>(name, uuid); break;
        case ObjectType::Node: return make_named<Node>(name, uuid);
        case ObjectType::MeshNode: return make_named<MeshNode>(name, uuid);
        case ObjectType::LightNode: return make_named<LightNode>(name, uuid);
        case ObjectType::CameraNode: return make_named<CameraNode>(name, uuid);
        default: error = "invalid object type " + std::to_string((int)type); return {};
    }
    if (assets.find(uuid) == assets.end()) load_order.push_back(uuid);
    assets[uuid] = asset;
    if (asset->type == ObjectType::SceneAsset && !initialScene) initialScene = uuid;
    return asset;
}

bool AssetManager::LoadProject(const std::string& path, const std::string& binPath) {
    std::string text, bin;
    if (!read_file(path, text)) {
        error = "Project file not found: " + path;
        return false;
    }
    if (!read_file(binPath, bin)) {
        error = "Project blob not found: " + binPath;
        return false;
    }
    std::vector<uint8_t> blob(bin.begin(), bin.end());
    lj::Value j;
    try {
        j = lj::parse(text);
    } catch (const std::exception& e) {
        error = e.what();
        return false;
    }
    error.clear();
    if (j.contains("assets") && j.at("assets").is_array())
        for (lj::Value& aj : *j.o->at("assets").a) load_object(aj, blob, *this);
    if (j.contains("scenes")) {
        lj::Value& sc = j.o->at("scenes");
        if (sc.is_object())
            for (auto& kv : *sc.o) load_object(kv.second, blob, *this);
        else if (sc.is_array())
            for (lj::Value& sj : *sc.a) load_object(sj, blob, *this);
    }
    if (j.contains("initialScene")) initialScene = j.at("initialScene").as_uint();
    for (auto& sc : GetAll<SceneAsset>(ObjectType::SceneAsset)) sc->UpdateParents();
    return error.empty();
}

bool AssetManager::SaveProject(const std::string& path, const std::string& binPath) {
    std::vector<uint8_t> blob;
    lj::Value j = lj::Value::object();
    j["scenes"] = lj::Value::object();
    lj::Value arr = lj::Value::array();
    std::vector<Ref<Asset>> ordered;
    for (UUID u : load_order) ordered.push_back(assets[u]);
    std::stable_sort(ordered.begin(), ordered.end(),
                     [](const Ref<Asset>& a, const Ref<Asset>& b) { return (int)a->type < (int)b->type; });
    for (auto& a : ordered) {
        if (a->type == ObjectType::SceneAsset) continue;
        arr.a->emplace_back();
        save_object(a, arr.a->back(), blob, *this);
    }
    j["assets"] = arr;
    for (auto& sc : GetAll<SceneAsset>(ObjectType::SceneAsset))
        save_object(sc, (*j.o->at("scenes").o)[std::to_string(sc->uuid)], blob, *this);
    j["initialScene"] = lj::Value::uinteger(initialScene);
    std::ofstream fb(binPath, std::ios::binary);
    std::ofstream fj(path, std::ios::binary);
    if (!fb || !fj) {
        error = "cannot write project files";
        return false;
    }
    fb.write((const char*)blob.data(), (std::streamsize)blob.size());
    const std::string text = lj::dump(j);
    fj.write(text.data(), (std::streamsize)text.size());
    return true;
}

Ref<SceneAsset> AssetManager::GetInitialScene() {
    if (!initialScene) CreateObject(ObjectType::SceneAsset, "DefaultScene", 0);
    return Get<SceneAsset>(initialScene);
}

Ref<CameraNode> AssetManager::GetMainCamera(const Ref<SceneAsset>& scene) {
    if (!scene->mainCamera) {
        auto cam = make_named<CameraNode>("Default Camera", NewUUID());
        scene->Add(cam);
        scene->mainCamera = cam;
    }
    return scene->mainCamera;
}

} // namespace luzhost

// ref_import.cpp -- drives the REFERENCE'S OWN scene importers (source/Resources/AssetIO.cpp: glTF / GLB through
// tiny_gltf, OBJ through tiny_obj_loader, textures through stb_image) and dumps what they produce -- meshes, materials,
// textures, the node tree -- as JSON with every float written as its bit pattern, so that luz_b200/host/import.cpp can
// be checked bit for bit (tests/test_host_import.py).  TEST INFRASTRUCTURE: built by oracle/Makefile into oracle/_ref/
// from the sources where they lie under /root/reference; tests/golden/make_import_golden.py runs it and commits its
// output.  It contains no reference code itself, it only calls it.
//
// usage: ref_import <scene.glb|.gltf|.obj> <out.json>      (run from a scratch dir: the reference logger writes Luz.log)
#include "AssetIO.hpp"
#include "AssetManager.hpp"

#include <cstdio>
#include <cstring>
#include <map>

static FILE* g_out = nullptr;
static std::vector<Ref<MeshAsset>> g_meshes;
static std::vector<Ref<MaterialAsset>> g_materials;
static std::vector<Ref<TextureAsset>> g_textures;

template <class T>
static int index_of(std::vector<Ref<T>>& v, const Ref<T>& p) {
    if (!p) return -1;
    for (size_t i = 0; i < v.size(); i++)
        if (v[i] == p) return (int)i;
    v.push_back(p);
    return (int)v.size() - 1;
}
static unsigned bits(float f) {
    unsigned u;
    memcpy(&u, &f, 4);
    return u;
}
static void put_str(const std::string& s) {
    fputc('"', g_out);
    for (char c : s) {
        if (c == '"' || c == '\\') fputc('\\', g_out);
        fputc(c, g_out);
    }
    fputc('"', g_out);
}
static void put_vec(const char* name, const float* v, int n) {
    fprintf(g_out, "\"%s\":[", name);
    for (int i = 0; i < n; i++) fprintf(g_out, "%s%u", i ? "," : "", bits(v[i]));
    fprintf(g_out, "]");
}
static void put_node(const Ref<Node>& n) {
    fprintf(g_out, "{\"name\":");
    put_str(n->name);
    fprintf(g_out, ",\"type\":%d,", (int)n->type);
    put_vec("position", &n->position.x, 3);
    fprintf(g_out, ",");
    put_vec("rotation", &n->rotation.x, 3);
    fprintf(g_out, ",");
    put_vec("scale", &n->scale.x, 3);
    int mesh = -1, material = -1;
    if (n->type == ObjectType::MeshNode) {
        auto mn = std::dynamic_pointer_cast<MeshNode>(n);
        mesh = index_of(g_meshes, mn->mesh);
        material = index_of(g_materials, mn->material);
    }
    fprintf(g_out, ",\"mesh\":%d,\"material\":%d,\"children\":[", mesh, material);
    for (size_t i = 0; i < n->children.size(); i++) {
        if (i) fputc(',', g_out);
        put_node(n->children[i]);
    }
    fprintf(g_out, "]}");
}
static unsigned long long fnv(const unsigned char* p, size_t n) {
    unsigned long long h = 1469598103934665603ull;
    for (size_t i = 0; i < n; i++) h = (h ^ p[i]) * 1099511628211ull;
    return h;
}

int main(int argc, char** argv) {
    if (argc < 3) {
        fprintf(stderr, "usage: ref_import scene.(glb|gltf|obj) out.json\n");
        return 2;
    }
    g_out = fopen(argv[2], "w");
    if (!g_out) return 3;
    Logger::Init();
    AssetManager assets;
    const UUID id = AssetIO::Import(argv[1], assets);
    Ref<SceneAsset> scene = assets.Get<SceneAsset>(id);
    if (!scene) {
        fprintf(g_out, "{\"error\":\"import failed\"}\n");
        fclose(g_out);
        return 0;
    }
    fprintf(g_out, "{\"scene\":");
    put_str(scene->name);
    fprintf(g_out, ",\"nodes\":[");
    for (size_t i = 0; i < scene->nodes.size(); i++) {
        if (i) fputc(',', g_out);
        put_node(scene->nodes[i]);
    }
    fprintf(g_out, "],\"materials\":[");
    for (size_t i = 0; i < g_materials.size(); i++) { // may grow g_textures only
        const auto& m = g_materials[i];
        if (i) fputc(',', g_out);
        fprintf(g_out, "{\"name\":");
        put_str(m->name);
        fprintf(g_out, ",");
        put_vec("color", &m->color.x, 4);
        fprintf(g_out, ",");
        put_vec("emission", &m->emission.x, 3);
        // textures are numbered in this order of first use (sequenced explicitly: argument evaluation order is not)
        const int ao = index_of(g_textures, m->aoMap);
        const int col = index_of(g_textures, m->colorMap);
        const int nrm = index_of(g_textures, m->normalMap);
        const int emi = index_of(g_textures, m->emissionMap);
        const int mr = index_of(g_textures, m->metallicRoughnessMap);
        fprintf(g_out, ",\"metallic\":%u,\"roughness\":%u,\"aoMap\":%d,\"colorMap\":%d,\"normalMap\":%d,\"emissionMap\":%d,"
                       "\"metallicRoughnessMap\":%d}",
                bits(m->metallic), bits(m->roughness), ao, col, nrm, emi, mr);
    }
    fprintf(g_out, "],\"textures\":[");
    for (size_t i = 0; i < g_textures.size(); i++) {
        const auto& t = g_textures[i];
        if (i) fputc(',', g_out);
        fprintf(g_out, "{\"name\":");
        put_str(t->name);
        fprintf(g_out, ",\"width\":%d,\"height\":%d,\"channels\":%d,\"bytes\":%zu,\"fnv1a\":\"%016llx\"}", t->width, t->height,
                t->channels, t->data.size(), fnv(t->data.data(), t->data.size()));
    }
    fprintf(g_out, "],\"meshes\":[");
    for (size_t i = 0; i < g_meshes.size(); i++) {
        const auto& m = g_meshes[i];
        if (i) fputc(',', g_out);
        fprintf(g_out, "{\"name\":");
        put_str(m->name);
        fprintf(g_out, ",\"vertex_count\":%zu,\"vertices\":[", m->vertices.size());
        const float* f = reinterpret_cast<const float*>(m->vertices.data());
        for (size_t k = 0; k < m->vertices.size() * 12; k++) fprintf(g_out, "%s%u", k ? "," : "", bits(f[k]));
        fprintf(g_out, "],\"indices\":[");
        for (size_t k = 0; k < m->indices.size(); k++) fprintf(g_out, "%s%u", k ? "," : "", m->indices[k]);
        fprintf(g_out, "]}");
    }
    fprintf(g_out, "]}\n");
    fclose(g_out);
    return 0;
}

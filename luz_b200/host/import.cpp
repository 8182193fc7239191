// import.cpp -- see import.hpp.  glTF 2.0 (Khronos specification), Wavefront OBJ / MTL and PNG readers written from the
// format descriptions, feeding the same asset construction the reference performs in source/Resources/AssetIO.cpp.
// Where the reference's third-party parsers define behaviour that shows in the result, it is restated and cited:
//   tiny_gltf.h (deps/, v2.x): Accessor::ByteStride, Parameter::ColorFactor, texture / material value maps
//   tiny_obj_loader.h (deps/, v2.0.0): decimal parsing (tryParseDouble), index fix-up, quad split by the shorter
//   diagonal, shape / per-face material bookkeeping of `o`, `g`, `usemtl`
// Build with -ffp-contract=off: the tangent generation is compared bit for bit.
#include "import.hpp"

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iterator>
#include <map>
#include <sstream>

#include "png.hpp"

namespace luzhost {
namespace AssetIO {

using lm::vec2;
using lm::vec3;
using lm::vec4;

namespace {

// ---- small file / path helpers ---------------------------------------------------------------------------------
bool read_file(const std::string& path, std::vector<uint8_t>& out) {
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    out.assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
    return true;
}
std::string extension(const std::string& path) { // std::filesystem::path::extension
    const size_t slash = path.find_last_of('/');
    const size_t dot = path.find_last_of('.');
    if (dot == std::string::npos || (slash != std::string::npos && dot < slash) || dot == (slash == std::string::npos ? 0 : slash + 1))
        return "";
    return path.substr(dot);
}
std::string stem(const std::string& path) { // std::filesystem::path::stem
    const size_t slash = path.find_last_of('/');
    const std::string file = slash == std::string::npos ? path : path.substr(slash + 1);
    const size_t dot = file.find_last_of('.');
    return (dot == std::string::npos || dot == 0) ? file : file.substr(0, dot);
}
std::string parent_path(const std::string& path) {
    const size_t slash = path.find_last_of('/');
    return slash == std::string::npos ? "" : path.substr(0, slash);
}

template <class T>
Ref<T> create(AssetManager& m, ObjectType type, const std::string& name) {
    return std::dynamic_pointer_cast<T>(m.CreateObject(type, name, 0));
}

bool decode_texture_bytes(const std::vector<uint8_t>& bytes, const Ref<TextureAsset>& t, std::string& err) {
    if (!decode_png(bytes.data(), bytes.size(), t->data, t->width, t->height, err)) return false;
    t->channels = 4;
    return true;
}

// ---- base64 (data: URIs of .gltf buffers and images) -----------------------------------------------------------
bool base64_decode(const std::string& in, size_t from, std::vector<uint8_t>& out) {
    uint32_t acc = 0;
    int bits = 0;
    for (size_t i = from; i < in.size(); i++) {
        const char c = in[i];
        int v;
        if (c >= 'A' && c <= 'Z') v = c - 'A';
        else if (c >= 'a' && c <= 'z') v = c - 'a' + 26;
        else if (c >= '0' && c <= '9') v = c - '0' + 52;
        else if (c == '+' || c == '-') v = 62;
        else if (c == '/' || c == '_') v = 63;
        else if (c == '=' || c == '\n' || c == '\r') continue;
        else return false;
        acc = (acc << 6) | (uint32_t)v;
        bits += 6;
        if (bits >= 8) {
            bits -= 8;
            out.push_back((uint8_t)((acc >> bits) & 0xFF));
        }
    }
    return true;
}

// ---- glTF accessors ----------------------------------------------------------------------------------------------
struct GltfModel {
    lj::Value json;
    std::vector<std::vector<uint8_t>> buffers;
    std::string base_dir;
};

int component_size(int component_type) {
    switch (component_type) {
        case 5120: case 5121: return 1; // BYTE, UNSIGNED_BYTE
        case 5122: case 5123: return 2; // SHORT, UNSIGNED_SHORT
        case 5125: case 5126: return 4; // UNSIGNED_INT, FLOAT
        default: return 0;
    }
}
int type_components(const std::string& t) {
    if (t == "SCALAR") return 1;
    if (t == "VEC2") return 2;
    if (t == "VEC3") return 3;
    if (t == "VEC4" || t == "MAT2") return 4;
    if (t == "MAT3") return 9;
    if (t == "MAT4") return 16;
    return 0;
}

struct View { // one accessor resolved against its bufferView and buffer
    const uint8_t* data = nullptr; // &buffer[view.byteOffset + accessor.byteOffset]
    size_t available = 0;          // bytes from `data` to the end of the buffer
    int stride_bytes = 0;          // tiny_gltf Accessor::ByteStride: the view's byteStride, else the element size
    int component_type = 0;
    uint32_t count = 0;
};

const lj::Value* member(const lj::Value& o, const char* k) {
    if (!o.is_object()) return nullptr;
    auto it = o.o->find(k);
    return it == o.o->end() ? nullptr : &it->second;
}
int64_t int_or(const lj::Value& o, const char* k, int64_t d) {
    const lj::Value* v = member(o, k);
    return (v && v->is_number()) ? v->as_int() : d;
}
std::string str_or(const lj::Value& o, const char* k, const std::string& d) {
    const lj::Value* v = member(o, k);
    return (v && v->kind == lj::Value::String) ? v->s : d;
}
const lj::Value* element(const lj::Value& root, const char* array, int64_t i) {
    const lj::Value* a = member(root, array);
    if (!a || !a->is_array() || i < 0 || (size_t)i >= a->size()) return nullptr;
    return &(*a->a)[(size_t)i];
}

bool resolve(const GltfModel& m, int64_t accessor_index, View& out, std::string& err) {
    const lj::Value* acc = element(m.json, "accessors", accessor_index);
    if (!acc) { err = "accessor index out of range"; return false; }
    const lj::Value* bv = element(m.json, "bufferViews", int_or(*acc, "bufferView", -1));
    if (!bv) { err = "accessor without a bufferView (sparse accessors are not supported)"; return false; }
    const int64_t buffer = int_or(*bv, "buffer", -1);
    if (buffer < 0 || (size_t)buffer >= m.buffers.size()) { err = "bufferView.buffer out of range"; return false; }
    const std::vector<uint8_t>& b = m.buffers[(size_t)buffer];
    // every quantity is validated as a signed 64-bit value before it is cast: negative offsets / strides / counts in a
    // crafted file must not wrap into range
    const int64_t view_off = int_or(*bv, "byteOffset", 0), acc_off = int_or(*acc, "byteOffset", 0);
    const int64_t count = int_or(*acc, "count", 0);
    if (view_off < 0 || acc_off < 0 || (uint64_t)view_off > b.size() || (uint64_t)acc_off > b.size()) {
        err = "accessor starts outside its buffer";
        return false;
    }
    if (count < 0 || count > (int64_t)UINT32_MAX) { err = "bad accessor count"; return false; }
    const uint64_t off = (uint64_t)view_off + (uint64_t)acc_off;
    out.component_type = (int)int_or(*acc, "componentType", 0);
    const int csize = component_size(out.component_type), ncomp = type_components(str_or(*acc, "type", ""));
    if (!csize || !ncomp) { err = "bad accessor type"; return false; }
    const int64_t view_stride = int_or(*bv, "byteStride", 0);
    // glTF 2.0: byteStride, when present, lies in [4, 252] and holds at least one element
    if (view_stride && (view_stride < (int64_t)csize * ncomp || view_stride > 252)) {
        err = "byteStride out of range";
        return false;
    }
    out.stride_bytes = view_stride ? (int)view_stride : csize * ncomp;
    if (view_stride && view_stride % csize) { err = "byteStride is not a multiple of the component size"; return false; }
    out.count = (uint32_t)count;
    if (off > b.size()) { err = "accessor starts outside its buffer"; return false; }
    out.data = b.data() + off;
    out.available = b.size() - off;
    if (out.count && (uint64_t)(out.count - 1) * (uint64_t)out.stride_bytes + (uint64_t)csize * ncomp > out.available) {
        err = "accessor runs past the end of its buffer";
        return false;
    }
    return true;
}

float read_f32(const uint8_t* p) {
    float f;
    memcpy(&f, p, 4);
    return f;
}

bool load_gltf_file(const std::string& path, GltfModel& m, std::string& err) {
    std::vector<uint8_t> file;
    if (!read_file(path, file)) { err = "cannot read " + path; return false; }
    m.base_dir = parent_path(path);
    std::string json_text;
    std::vector<uint8_t> bin_chunk;
    bool have_bin = false;
    if (extension(path) == ".gltf") {
        json_text.assign(file.begin(), file.end());
    } else { // binary glTF: 12-byte header, then chunks {u32 length, u32 type, data}
        if (file.size() < 20 || memcmp(file.data(), "glTF", 4) != 0) { err = "not a GLB file"; return false; }
        uint32_t version, length;
        memcpy(&version, file.data() + 4, 4);
        memcpy(&length, file.data() + 8, 4);
        if (version != 2 || length > file.size()) { err = "unsupported GLB header"; return false; }
        size_t off = 12;
        while (off + 8 <= length) {
            uint32_t clen, ctype;
            memcpy(&clen, file.data() + off, 4);
            memcpy(&ctype, file.data() + off + 4, 4);
            if (off + 8 + (size_t)clen > length) { err = "GLB chunk runs past the end of the file"; return false; }
            if (ctype == 0x4E4F534Au) json_text.assign((const char*)file.data() + off + 8, clen);
            else if (ctype == 0x004E4942u && !have_bin) bin_chunk.assign(file.data() + off + 8, file.data() + off + 8 + clen), have_bin = true;
            off += 8 + (size_t)clen;
        }
    }
    try {
        m.json = lj::parse(json_text);
    } catch (const std::exception& e) {
        err = std::string("glTF JSON: ") + e.what();
        return false;
    }
    if (!m.json.is_object()) { err = "glTF JSON is not an object"; return false; }
    if (const lj::Value* bufs = member(m.json, "buffers")) {
        for (size_t i = 0; bufs->is_array() && i < bufs->size(); i++) {
            const lj::Value& b = (*bufs->a)[i];
            std::vector<uint8_t> data;
            const std::string uri = str_or(b, "uri", "");
            if (uri.empty()) {
                if (i != 0 || !have_bin) { err = "buffer without uri and no GLB binary chunk"; return false; }
                data = bin_chunk;
            } else if (uri.compare(0, 5, "data:") == 0) {
                const size_t comma = uri.find(',');
                if (comma == std::string::npos || uri.find(";base64") == std::string::npos || !base64_decode(uri, comma + 1, data)) {
                    err = "bad data URI in buffer";
                    return false;
                }
            } else if (!read_file(m.base_dir.empty() ? uri : m.base_dir + "/" + uri, data)) {
                err = "cannot read buffer " + uri;
                return false;
            }
            const uint64_t want = (uint64_t)int_or(b, "byteLength", 0);
            if (data.size() < want) { err = "buffer shorter than its byteLength"; return false; }
            m.buffers.push_back(std::move(data));
        }
    }
    return true;
}

// glm::eulerAngles(quat) = (pitch, yaw, roll) (glm/gtc/quaternion.inl) and glm::degrees
vec3 euler_degrees(const lm::quat& q) {
    const float eps = 1.1920928955078125e-07f;
    float pitch;
    {
        const float y = 2.0f * (q.y * q.z + q.w * q.x);
        const float x = q.w * q.w - q.x * q.x - q.y * q.y + q.z * q.z;
        if (std::fabs(x) <= eps && std::fabs(y) <= eps) pitch = 2.0f * std::atan2(q.x, q.w);
        else pitch = std::atan2(y, x);
    }
    float s = -2.0f * (q.x * q.z - q.w * q.y);
    s = std::min(std::max(s, -1.0f), 1.0f);
    const float yaw = std::asin(s);
    const float roll = std::atan2(2.0f * (q.x * q.y + q.w * q.z), q.w * q.w + q.x * q.x - q.y * q.y - q.z * q.z);
    const float k = 57.295779513082320876798154814105f;
    return vec3(pitch * k, yaw * k, roll * k);
}

float length3(vec3 v) { return std::sqrt(lm::dot(v, v)); }

// glm::decompose (glm/gtx/matrix_decompose.inl) for affine matrices: scale, rotation, translation.  Returns false
// (outputs untouched) where glm does: m[3][3] == 0 or a singular upper 3x3.
bool decompose_affine(const lm::mat4& in, vec3& scale, lm::quat& q, vec3& translation, bool& has_perspective) {
    const float eps = 1.1920928955078125e-07f;
    lm::mat4 L = in;
    has_perspective = false;
    if (std::fabs(L[3][3]) < eps) return false;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) L[i][j] /= L[3][3];
    // singularity test on the matrix with the perspective partition cleared
    lm::mat4 P = L;
    for (int i = 0; i < 3; i++) P[i][3] = 0.0f;
    P[3][3] = 1.0f;
    const float det = P[0][0] * (P[1][1] * P[2][2] - P[2][1] * P[1][2]) - P[1][0] * (P[0][1] * P[2][2] - P[2][1] * P[0][2]) +
                      P[2][0] * (P[0][1] * P[1][2] - P[1][1] * P[0][2]);
    if (std::fabs(det) < eps) return false;
    if (std::fabs(L[0][3]) >= eps || std::fabs(L[1][3]) >= eps || std::fabs(L[2][3]) >= eps) {
        has_perspective = true; // glm solves for the perspective vector and clears the partition; the rest is the same
        L[0][3] = L[1][3] = L[2][3] = 0.0f;
        L[3][3] = 1.0f;
    }
    translation = vec3(L[3][0], L[3][1], L[3][2]);
    vec3 row[3];
    for (int i = 0; i < 3; i++) row[i] = vec3(L[i][0], L[i][1], L[i][2]);
    auto rescale = [](vec3 v, float desired) { return v * (desired / length3(v)); };       // detail::scale
    auto combine = [](vec3 a, vec3 b, float as, float bs) { return (a * as) + (b * bs); }; // detail::combine
    vec3 skew;
    scale.x = length3(row[0]);
    row[0] = rescale(row[0], 1.0f);
    skew.z = lm::dot(row[0], row[1]);
    row[1] = combine(row[1], row[0], 1.0f, -skew.z);
    scale.y = length3(row[1]);
    row[1] = rescale(row[1], 1.0f);
    skew.z /= scale.y;
    skew.y = lm::dot(row[0], row[2]);
    row[2] = combine(row[2], row[0], 1.0f, -skew.y);
    skew.x = lm::dot(row[1], row[2]);
    row[2] = combine(row[2], row[1], 1.0f, -skew.x);
    scale.z = length3(row[2]);
    row[2] = rescale(row[2], 1.0f);
    const vec3 pdum = lm::cross(row[1], row[2]);
    if (lm::dot(row[0], pdum) < 0.0f) {
        for (int i = 0; i < 3; i++) {
            scale[i] *= -1.0f;
            row[i] = row[i] * -1.0f;
        }
    }
    float o[4] = {0, 0, 0, 0}; // x y z w
    const float trace = row[0].x + row[1].y + row[2].z;
    if (trace > 0.0f) {
        float root = std::sqrt(trace + 1.0f);
        o[3] = 0.5f * root;
        root = 0.5f / root;
        o[0] = root * (row[1].z - row[2].y);
        o[1] = root * (row[2].x - row[0].z);
        o[2] = root * (row[0].y - row[1].x);
    } else {
        static const int next[3] = {1, 2, 0};
        int i = 0;
        if (row[1].y > row[0].x) i = 1;
        if (row[2].z > row[i][i]) i = 2;
        const int j = next[i], k = next[j];
        float root = std::sqrt(row[i][i] - row[j][j] - row[k][k] + 1.0f);
        o[i] = 0.5f * root;
        root = 0.5f / root;
        o[j] = root * (row[i][j] + row[j][i]);
        o[k] = root * (row[i][k] + row[k][i]);
        o[3] = root * (row[j][k] - row[k][j]);
    }
    q.x = o[0], q.y = o[1], q.z = o[2], q.w = o[3];
    return true;
}

// ---- tangents (AssetIO.cpp:313-346): per-triangle s / t directions accumulated per vertex, Gram-Schmidt against the
// normal, handedness in w.  The reference accumulates into `new glm::vec3[...]`, which is not zero-initialised there;
// zero is what it relies on (fresh heap pages) and what the golden run enforces (tests/golden/make_import_golden.py).
void generate_tangents(MeshAsset& mesh, uint32_t vertex_count) {
    std::vector<vec3> tan1(mesh.vertices.size(), vec3(0.0f)), tan2(mesh.vertices.size(), vec3(0.0f));
    (void)vertex_count;
    for (size_t id = 0; id + 2 < mesh.indices.size(); id += 3) {
        const uint32_t i1 = mesh.indices[id + 0], i2 = mesh.indices[id + 2], i3 = mesh.indices[id + 1];
        if (i1 >= mesh.vertices.size() || i2 >= mesh.vertices.size() || i3 >= mesh.vertices.size()) continue;
        const auto &v1 = mesh.vertices[i1], &v2 = mesh.vertices[i2], &v3 = mesh.vertices[i3];
        const vec3 e1 = v2.position - v1.position, e2 = v3.position - v1.position;
        const vec2 duv1{v2.texCoord.x - v1.texCoord.x, v2.texCoord.y - v1.texCoord.y};
        const vec2 duv2{v3.texCoord.x - v1.texCoord.x, v3.texCoord.y - v1.texCoord.y};
        const float f = 1.0f / (duv1.x * duv2.y - duv2.x * duv1.y);
        const vec3 sdir = ((e1 * duv2.y) - (e2 * duv1.y)) * f;
        const vec3 tdir = ((e2 * duv1.x) - (e1 * duv2.x)) * f;
        tan1[i1] = tan1[i1] + sdir;
        tan1[i2] = tan1[i2] + sdir;
        tan1[i3] = tan1[i3] + sdir;
        tan2[i1] = tan2[i1] + tdir;
        tan2[i2] = tan2[i2] + tdir;
        tan2[i3] = tan2[i3] + tdir;
    }
    for (size_t a = 0; a < mesh.vertices.size(); a++) {
        const vec3 t = tan1[a];
        auto& v = mesh.vertices[a];
        const vec3 n = v.normal;
        v.tangent = vec4(lm::normalize(t - n * lm::dot(t, n)), 1.0f);
        v.tangent.w = (lm::dot(lm::cross(n, t), tan2[a]) < 0.0f) ? -1.0f : 1.0f;
    }
}

// ---- OBJ ---------------------------------------------------------------------------------------------------------
// Decimal -> double the way tiny_obj_loader's tryParseDouble does it (digit by digit in double arithmetic, fraction
// digits weighted by a power-of-ten table, exponent applied as ldexp(m * 5^e, e)), then narrowed to float.
bool parse_double(const char* s, const char* end, double& result) {
    if (s >= end) return false;
    double mantissa = 0.0;
    int exponent = 0, read = 0;
    char sign = '+', exp_sign = '+';
    const char* c = s;
    bool leading_dot = false;
    if (*c == '+' || *c == '-') {
        sign = *c++;
        if (c != end && *c == '.') leading_dot = true;
    } else if (*c >= '0' && *c <= '9') {
    } else if (*c == '.') {
        leading_dot = true;
    } else {
        return false;
    }
    bool more = c != end;
    if (!leading_dot) {
        while (more && *c >= '0' && *c <= '9') {
            mantissa *= 10;
            mantissa += (int)(*c - '0');
            c++, read++;
            more = c != end;
        }
        if (read == 0) return false;
    }
    if (more && *c == '.') {
        static const double lut[] = {1.0, 0.1, 0.01, 0.001, 0.0001, 0.00001, 0.000001, 0.0000001};
        c++;
        read = 1;
        more = c != end;
        while (more && *c >= '0' && *c <= '9') {
            mantissa += (int)(*c - '0') * (read < 8 ? lut[read] : std::pow(10.0, -read));
            read++, c++;
            more = c != end;
        }
    } else if (more && !(*c == 'e' || *c == 'E')) {
        more = false; // anything else ends the number
    }
    if (more && (*c == 'e' || *c == 'E')) {
        c++;
        more = c != end;
        if (more && (*c == '+' || *c == '-')) exp_sign = *c++;
        else if (more && *c >= '0' && *c <= '9') {}
        else return false;
        read = 0;
        more = c != end;
        while (more && *c >= '0' && *c <= '9') {
            if (exponent > 2147483647 / 10) return false;
            exponent = exponent * 10 + (int)(*c - '0');
            c++, read++;
            more = c != end;
        }
        if (exp_sign == '-') exponent = -exponent;
        if (read == 0) return false;
    }
    result = (sign == '+' ? 1 : -1) * (exponent ? std::ldexp(mantissa * std::pow(5.0, exponent), exponent) : mantissa);
    return true;
}
float parse_real(const char*& tok, double def = 0.0) {
    tok += strspn(tok, " \t");
    const char* end = tok + strcspn(tok, " \t\r");
    double v = def;
    parse_double(tok, end, v);
    tok = end;
    return (float)v;
}
std::string parse_string(const char*& tok) {
    tok += strspn(tok, " \t");
    const size_t n = strcspn(tok, " \t\r");
    std::string s(tok, n);
    tok += n;
    return s;
}

struct ObjIndex {
    int v = -1, vt = -1, vn = -1;
};
struct ObjMaterial {
    std::string name;
    float diffuse[3] = {0, 0, 0}, specular[3] = {0, 0, 0}, emission[3] = {0, 0, 0};
    float roughness = 0, metallic = 0;
    std::string diffuse_texname, normal_texname;
};
struct ObjShape {
    std::string name;
    std::vector<ObjIndex> indices; // triangulated
    std::vector<int> material_ids; // per triangle
};

bool fix_index(int idx, int n, int& out) { // 1-based, negative = relative to the end; 0 is invalid
    if (idx > 0) { out = idx - 1; return true; }
    if (idx == 0) return false;
    out = n + idx;
    return true;
}
bool parse_triple(const char*& tok, int nv, int nvn, int nvt, ObjIndex& out) {
    ObjIndex vi;
    if (!fix_index(atoi(tok), nv, vi.v)) return false;
    tok += strcspn(tok, "/ \t\r");
    if (tok[0] != '/') { out = vi; return true; }
    tok++;
    if (tok[0] == '/') { // v//vn
        tok++;
        if (!fix_index(atoi(tok), nvn, vi.vn)) return false;
        tok += strcspn(tok, "/ \t\r");
        out = vi;
        return true;
    }
    if (!fix_index(atoi(tok), nvt, vi.vt)) return false; // v/vt[/vn]
    tok += strcspn(tok, "/ \t\r");
    if (tok[0] != '/') { out = vi; return true; }
    tok++;
    if (!fix_index(atoi(tok), nvn, vi.vn)) return false;
    tok += strcspn(tok, "/ \t\r");
    out = vi;
    return true;
}

void load_mtl(const std::string& path, std::vector<ObjMaterial>& materials, std::map<std::string, int>& by_name) {
    std::ifstream f(path);
    if (!f) return;
    ObjMaterial cur;
    bool have = false;
    std::string line;
    auto flush = [&]() {
        if (!have) return;
        by_name[cur.name] = (int)materials.size();
        materials.push_back(cur);
    };
    while (std::getline(f, line)) {
        while (!line.empty() && (line.back() == '\n' || line.back() == '\r')) line.pop_back();
        const char* tok = line.c_str();
        tok += strspn(tok, " \t");
        if (!tok[0] || tok[0] == '#') continue;
        auto key = [&](const char* k) {
            const size_t n = strlen(k);
            return strncmp(tok, k, n) == 0 && (tok[n] == ' ' || tok[n] == '\t');
        };
        if (key("newmtl")) {
            flush();
            cur = ObjMaterial();
            have = true;
            tok += 7;
            cur.name = parse_string(tok); // tiny_obj takes the rest of the line; names with spaces are not supported here
            continue;
        }
        if (key("Kd")) { tok += 2; for (int k = 0; k < 3; k++) cur.diffuse[k] = parse_real(tok); continue; }
        if (key("Ks")) { tok += 2; for (int k = 0; k < 3; k++) cur.specular[k] = parse_real(tok); continue; }
        if (key("Ke")) { tok += 2; for (int k = 0; k < 3; k++) cur.emission[k] = parse_real(tok); continue; }
        if (key("Pr")) { tok += 2; cur.roughness = parse_real(tok); continue; }
        if (key("Pm")) { tok += 2; cur.metallic = parse_real(tok); continue; }
        if (key("map_Kd")) { tok += 7; cur.diffuse_texname = parse_string(tok); continue; }
        if (key("norm")) { tok += 5; cur.normal_texname = parse_string(tok); continue; }
    }
    flush();
}

// crossing-number point-in-polygon test (W. R. Franklin's pnpoly, as tiny_obj_loader uses it on one triangle)
bool point_in_triangle(const float* px, const float* py, float tx, float ty) {
    bool inside = false;
    for (int i = 0, j = 2; i < 3; j = i++)
        if (((py[i] > ty) != (py[j] > ty)) && (tx < (px[j] - px[i]) * (ty - py[i]) / (py[j] - py[i]) + px[i])) inside = !inside;
    return inside;
}

// Polygons with more than four corners: tiny_obj_loader's built-in ear clipping.  The polygon is projected onto the
// coordinate plane picked from its first non-degenerate corner; starting at corner `guess`, a corner triple is cut
// off as a triangle if it turns the same way as the (first-edge) area sign and no other remaining corner lies inside
// it, otherwise the start moves on by one; the cut corner is removed and the search continues from the same index.
// It gives up after a full round without progress and emits what is left if exactly three corners remain.
void triangulate_polygon(const std::vector<ObjIndex>& face, const std::vector<float>& v, ObjShape& shape, int material) {
    size_t n = face.size();
    size_t axes[2] = {1, 2};
    for (size_t k = 0; k < n; ++k) {
        const size_t a = (size_t)face[k % n].v, b = (size_t)face[(k + 1) % n].v, c = (size_t)face[(k + 2) % n].v;
        if (3 * a + 2 >= v.size() || 3 * b + 2 >= v.size() || 3 * c + 2 >= v.size()) continue;
        const float e0x = v[b * 3] - v[a * 3], e0y = v[b * 3 + 1] - v[a * 3 + 1], e0z = v[b * 3 + 2] - v[a * 3 + 2];
        const float e1x = v[c * 3] - v[b * 3], e1y = v[c * 3 + 1] - v[b * 3 + 1], e1z = v[c * 3 + 2] - v[b * 3 + 2];
        const float cx = std::fabs(e0y * e1z - e0z * e1y), cy = std::fabs(e0z * e1x - e0x * e1z), cz = std::fabs(e0x * e1y - e0y * e1x);
        const float eps = 1.1920928955078125e-07f;
        if (cx > eps || cy > eps || cz > eps) {
            if (!(cx > cy && cx > cz)) {
                axes[0] = 0;
                if (cz > cx && cz > cy) axes[1] = 1;
            }
            break;
        }
    }
    std::vector<ObjIndex> rest = face;
    size_t guess = 0, iterations = face.size(), previous = rest.size();
    while (rest.size() > 3 && iterations > 0) {
        n = rest.size();
        if (guess >= n) guess -= n;
        if (previous != n) {
            previous = n;
            iterations = n;
        } else {
            iterations--;
        }
        ObjIndex ind[3];
        float vx[3], vy[3];
        for (size_t k = 0; k < 3; k++) {
            ind[k] = rest[(guess + k) % n];
            const size_t vi = (size_t)ind[k].v;
            const bool ok = vi * 3 + axes[0] < v.size() && vi * 3 + axes[1] < v.size();
            vx[k] = ok ? v[vi * 3 + axes[0]] : 0.0f;
            vy[k] = ok ? v[vi * 3 + axes[1]] : 0.0f;
        }
        const float e0x = vx[1] - vx[0], e0y = vy[1] - vy[0], e1x = vx[2] - vx[1], e1y = vy[2] - vy[1];
        const float cross = e0x * e1y - e0y * e1x;
        const float area = (vx[0] * vy[1] - vy[0] * vx[1]) * 0.5f;
        if (cross * area < 0.0f) {
            guess += 1;
            continue;
        }
        bool overlap = false;
        for (size_t other = 3; other < n; ++other) {
            const size_t idx = (guess + other) % n;
            if (idx >= rest.size()) continue;
            const size_t ovi = (size_t)rest[idx].v;
            if (ovi * 3 + axes[0] >= v.size() || ovi * 3 + axes[1] >= v.size()) continue;
            if (point_in_triangle(vx, vy, v[ovi * 3 + axes[0]], v[ovi * 3 + axes[1]])) {
                overlap = true;
                break;
            }
        }
        if (overlap) {
            guess += 1;
            continue;
        }
        shape.indices.push_back(ind[0]);
        shape.indices.push_back(ind[1]);
        shape.indices.push_back(ind[2]);
        shape.material_ids.push_back(material);
        rest.erase(rest.begin() + (ptrdiff_t)((guess + 1) % n));
    }
    if (rest.size() == 3) {
        shape.indices.insert(shape.indices.end(), rest.begin(), rest.end());
        shape.material_ids.push_back(material);
    }
}

// flushes the faces gathered since the last flush into `shape` (tiny_obj_loader exportGroupsToShape, triangulate = true)
bool export_faces(ObjShape& shape, std::vector<std::vector<ObjIndex>>& faces, int material, const std::string& name,
                  const std::vector<float>& v, std::string& err) {
    if (faces.empty()) return false;
    shape.name = name;
    for (const auto& face : faces) {
        const size_t n = face.size();
        if (n < 3) continue;
        if (n == 3) {
            shape.indices.insert(shape.indices.end(), face.begin(), face.end());
            shape.material_ids.push_back(material);
        } else if (n == 4) {
            bool ok = true;
            for (int k = 0; k < 4; k++) ok = ok && face[k].v >= 0 && (size_t)(3 * face[k].v + 2) < v.size();
            if (!ok) continue;
            auto p = [&](int k, int c) { return v[(size_t)face[k].v * 3 + c]; };
            const float e02x = p(2, 0) - p(0, 0), e02y = p(2, 1) - p(0, 1), e02z = p(2, 2) - p(0, 2);
            const float e13x = p(3, 0) - p(1, 0), e13y = p(3, 1) - p(1, 1), e13z = p(3, 2) - p(1, 2);
            const float sqr02 = e02x * e02x + e02y * e02y + e02z * e02z, sqr13 = e13x * e13x + e13y * e13y + e13z * e13z;
            static const int a[6] = {0, 1, 2, 0, 2, 3}, b[6] = {0, 1, 3, 1, 2, 3};
            const int* order = sqr02 < sqr13 ? a : b; // split along the shorter diagonal
            for (int k = 0; k < 6; k++) shape.indices.push_back(face[order[k]]);
            shape.material_ids.push_back(material);
            shape.material_ids.push_back(material);
        } else {
            triangulate_polygon(face, v, shape, material);
        }
    }
    return true;
}

} // namespace

bool IsTexture(const std::string& path) {
    const std::string e = extension(path);
    return e == ".jpg" || e == ".png" || e == ".jpeg" || e == ".tga" || e == ".bmp";
}
bool IsScene(const std::string& path) {
    const std::string e = extension(path);
    return e == ".obj" || e == ".gltf" || e == ".glb";
}

UUID Import(const std::string& path, AssetManager& manager) {
    if (IsTexture(path)) return ImportTexture(path, manager);
    if (IsScene(path)) return ImportScene(path, manager);
    return 0;
}

static bool import_texture_into(const std::string& path, const Ref<TextureAsset>& t, AssetManager& manager) {
    std::vector<uint8_t> bytes;
    if (!read_file(path, bytes)) {
        manager.error = "cannot read texture " + path;
        return false;
    }
    std::string err;
    if (!decode_texture_bytes(bytes, t, err)) {
        manager.error = "texture " + path + ": " + err + " (only PNG is decoded)";
        return false;
    }
    return true;
}

UUID ImportTexture(const std::string& path, AssetManager& manager) {
    auto t = create<TextureAsset>(manager, ObjectType::TextureAsset, stem(path));
    if (!import_texture_into(path, t, manager)) return 0;
    return t->uuid;
}

UUID ImportScene(const std::string& path, AssetManager& manager) {
    const std::string e = extension(path);
    if (e == ".gltf" || e == ".glb") return ImportSceneGLTF(path, manager);
    if (e == ".obj") return ImportSceneOBJ(path, manager);
    return 0;
}

UUID ImportSceneGLTF(const std::string& path, AssetManager& manager) {
    GltfModel model;
    std::string err;
    if (!load_gltf_file(path, model, err)) {
        manager.error = "Failed to parse glTF: " + err;
        return 0;
    }
    const lj::Value& J = model.json;
    auto array_size = [&](const char* k) {
        const lj::Value* a = member(J, k);
        return (a && a->is_array()) ? a->size() : (size_t)0;
    };

    // textures (AssetIO.cpp:146-169): one TextureAsset per glTF texture, named after the texture, always RGBA8
    std::vector<Ref<TextureAsset>> loadedTextures(array_size("textures"));
    for (size_t i = 0; i < loadedTextures.size(); i++) {
        const lj::Value& tex = *element(J, "textures", (int64_t)i);
        const lj::Value* img = element(J, "images", int_or(tex, "source", -1));
        if (!img) { manager.error = "texture without a valid image source"; return 0; }
        std::vector<uint8_t> bytes;
        const std::string uri = str_or(*img, "uri", "");
        if (member(*img, "bufferView")) {
            const lj::Value* bv = element(J, "bufferViews", int_or(*img, "bufferView", -1));
            const int64_t b = bv ? int_or(*bv, "buffer", -1) : -1;
            if (!bv || b < 0 || (size_t)b >= model.buffers.size()) { manager.error = "image bufferView out of range"; return 0; }
            const uint64_t off = (uint64_t)int_or(*bv, "byteOffset", 0), len = (uint64_t)int_or(*bv, "byteLength", 0);
            if (off + len > model.buffers[(size_t)b].size()) { manager.error = "image bufferView runs past its buffer"; return 0; }
            bytes.assign(model.buffers[(size_t)b].begin() + (ptrdiff_t)off, model.buffers[(size_t)b].begin() + (ptrdiff_t)(off + len));
        } else if (uri.compare(0, 5, "data:") == 0) {
            const size_t comma = uri.find(',');
            if (comma == std::string::npos || !base64_decode(uri, comma + 1, bytes)) { manager.error = "bad data URI in image"; return 0; }
        } else if (!read_file(model.base_dir.empty() ? uri : model.base_dir + "/" + uri, bytes)) {
            manager.error = "cannot read image " + uri;
            return 0;
        }
        loadedTextures[i] = create<TextureAsset>(manager, ObjectType::TextureAsset, str_or(tex, "name", ""));
        std::vector<uint16_t> deep;
        Ref<TextureAsset>& t = loadedTextures[i];
        if (!decode_png(bytes.data(), bytes.size(), t->data, t->width, t->height, err, &deep)) {
            manager.error = "glTF image " + std::to_string(i) + ": " + err + " (only PNG is decoded)";
            return 0;
        }
        t->channels = 4;
        if (!deep.empty()) {
            // tiny_gltf hands 16-bit PNGs over as 16-bit samples and the reference copies the first width * height * 4
            // BYTES of them (AssetIO.cpp:150-163), i.e. the first half of the image as little-endian u16: mirrored
            memcpy(t->data.data(), deep.data(), t->data.size());
        }
    }
    auto texture_of = [&](const lj::Value& holder, const char* key) -> Ref<TextureAsset> {
        const lj::Value* t = member(holder, key);
        if (!t) return {};
        const int64_t idx = int_or(*t, "index", -1);
        return (idx >= 0 && (size_t)idx < loadedTextures.size()) ? loadedTextures[(size_t)idx] : Ref<TextureAsset>();
    };

    // materials (AssetIO.cpp:171-209)
    std::vector<Ref<MaterialAsset>> materials(array_size("materials"));
    for (size_t i = 0; i < materials.size(); i++) {
        const lj::Value& mat = *element(J, "materials", (int64_t)i);
        materials[i] = create<MaterialAsset>(manager, ObjectType::MaterialAsset, str_or(mat, "name", ""));
        if (const lj::Value* pbr = member(mat, "pbrMetallicRoughness")) {
            if (member(*pbr, "baseColorTexture")) materials[i]->colorMap = texture_of(*pbr, "baseColorTexture");
            if (member(*pbr, "metallicRoughnessTexture")) materials[i]->metallicRoughnessMap = texture_of(*pbr, "metallicRoughnessTexture");
            if (const lj::Value* f = member(*pbr, "baseColorFactor")) { // Parameter::ColorFactor: rgb + (a or 1)
                if (f->is_array() && f->size() >= 3) {
                    for (int k = 0; k < 3; k++) materials[i]->color[k] = (float)(*f->a)[k].as_double();
                    materials[i]->color[3] = f->size() > 3 ? (float)(*f->a)[3].as_double() : 1.0f;
                }
            }
            if (const lj::Value* f = member(*pbr, "roughnessFactor")) materials[i]->roughness = (float)f->as_double();
            if (const lj::Value* f = member(*pbr, "metallicFactor")) materials[i]->metallic = (float)f->as_double();
        }
        if (member(mat, "normalTexture")) materials[i]->normalMap = texture_of(mat, "normalTexture");
        if (member(mat, "emissiveTexture")) materials[i]->emissionMap = texture_of(mat, "emissiveTexture");
        if (member(mat, "occlusionTexture")) materials[i]->aoMap = texture_of(mat, "occlusionTexture");
        if (const lj::Value* f = member(mat, "emissiveFactor"))
            if (f->is_array() && f->size() >= 3)
                for (int k = 0; k < 3; k++) materials[i]->emission[k] = (float)(*f->a)[k].as_double();
    }

    // meshes: one MeshAsset per primitive (AssetIO.cpp:211-347)
    std::vector<Ref<MeshAsset>> loadedMeshes;
    std::vector<int> loadedMeshMaterials;
    for (size_t mi = 0; mi < array_size("meshes"); mi++) {
        const lj::Value& mesh = *element(J, "meshes", (int64_t)mi);
        const lj::Value* prims = member(mesh, "primitives");
        for (size_t i = 0; prims && prims->is_array() && i < prims->size(); i++) {
            const lj::Value& prim = (*prims->a)[i];
            const std::string mesh_name = str_or(mesh, "name", "");
            const std::string name = (mesh_name != "" ? mesh_name : stem(path)) + "_" + std::to_string(i);
            Ref<MeshAsset> desc = create<MeshAsset>(manager, ObjectType::MeshAsset, name);
            loadedMeshes.push_back(desc);
            loadedMeshMaterials.push_back((int)int_or(prim, "material", -1));
            const lj::Value* attrs = member(prim, "attributes");
            auto attribute = [&](const char* semantic, View& view, bool& present) -> bool {
                present = false;
                const lj::Value* a = attrs ? member(*attrs, semantic) : nullptr;
                if (!a) return true;
                if (!resolve(model, a->as_int(), view, err)) return false;
                if (view.component_type != 5126) { err = std::string(semantic) + " is not FLOAT"; return false; }
                present = true;
                return true;
            };
            View pos, nrm, tan, uv;
            bool hasPos, hasNrm, hasTan, hasUV;
            if (!attribute("POSITION", pos, hasPos) || !attribute("NORMAL", nrm, hasNrm) || !attribute("TANGENT", tan, hasTan) ||
                !attribute("TEXCOORD_0", uv, hasUV)) {
                manager.error = "glTF primitive " + name + ": " + err;
                return 0;
            }
            if (!hasPos) { manager.error = "Primitive don't have position attribute"; return 0; }
            const uint32_t vertexCount = pos.count;
            if ((hasNrm && nrm.count < vertexCount) || (hasTan && tan.count < vertexCount) || (hasUV && uv.count < vertexCount)) {
                manager.error = "glTF primitive " + name + ": attribute accessors shorter than POSITION";
                return 0;
            }
            desc->vertices.reserve(vertexCount);
            for (uint32_t v = 0; v < vertexCount; v++) {
                MeshAsset::MeshVertex vertex{};
                const uint8_t* p = pos.data + (size_t)v * pos.stride_bytes;
                vertex.position = vec3(read_f32(p), read_f32(p + 4), read_f32(p + 8));
                if (hasNrm) {
                    const uint8_t* q = nrm.data + (size_t)v * nrm.stride_bytes;
                    vertex.normal = vec3(read_f32(q), read_f32(q + 4), read_f32(q + 8));
                }
                if (hasUV) {
                    const uint8_t* q = uv.data + (size_t)v * uv.stride_bytes;
                    vertex.texCoord.x = read_f32(q);
                    vertex.texCoord.y = read_f32(q + 4);
                }
                if (hasTan) {
                    const uint8_t* q = tan.data + (size_t)v * tan.stride_bytes;
                    vertex.tangent = vec4(read_f32(q), read_f32(q + 4), read_f32(q + 8), read_f32(q + 12));
                }
                desc->vertices.push_back(vertex);
            }
            if (!member(prim, "indices")) { manager.error = "Non indexed primitive not supported!"; return 0; }
            View idx;
            if (!resolve(model, int_or(prim, "indices", -1), idx, err)) { manager.error = "glTF primitive " + name + ": " + err; return 0; }
            desc->indices.reserve(idx.count);
            for (uint32_t k = 0; k < idx.count; k++) { // packed, like the reference's pointer walk (AssetIO.cpp:291-309)
                uint32_t value;
                if (idx.component_type == 5125) { memcpy(&value, idx.data + 4 * (size_t)k, 4); }
                else if (idx.component_type == 5123) { uint16_t s; memcpy(&s, idx.data + 2 * (size_t)k, 2); value = s; }
                else if (idx.component_type == 5121) { value = idx.data[k]; }
                else { manager.error = "Index type not supported!"; return 0; }
                desc->indices.push_back(value);
            }
            if (!hasTan) generate_tangents(*desc, vertexCount);
        }
    }

    // nodes (AssetIO.cpp:349-396): every glTF node becomes a group Node; its mesh a MeshNode child.
    // The reference indexes loadedMeshes (one entry per PRIMITIVE) with node.mesh (a MESH index): identical whenever
    // every mesh has one primitive; mirrored literally otherwise.
    std::vector<Ref<Node>> loadedNodes;
    for (size_t ni = 0; ni < array_size("nodes"); ni++) {
        const lj::Value& node = *element(J, "nodes", (int64_t)ni);
        const std::string node_name = str_or(node, "name", "");
        Ref<Node> groupNode = create<Node>(manager, ObjectType::Node, node_name);
        const int64_t mesh = int_or(node, "mesh", -1);
        if (mesh >= 0) {
            if ((size_t)mesh >= loadedMeshes.size()) { manager.error = "node.mesh out of range"; return 0; }
            Ref<MeshNode> meshNode = create<MeshNode>(manager, ObjectType::MeshNode, node_name);
            meshNode->mesh = loadedMeshes[(size_t)mesh];
            const int matId = loadedMeshMaterials[(size_t)mesh];
            if (matId >= 0 && (size_t)matId < materials.size()) meshNode->material = materials[(size_t)matId];
            Node::SetParent(meshNode, groupNode);
        }
        if (const lj::Value* ext = member(node, "extensions"))
            if (const lj::Value* kl = member(*ext, "KHR_lights_punctual"))
                if (int_or(*kl, "light", -1) >= 0) {
                    Ref<LightNode> lightNode = create<LightNode>(manager, ObjectType::LightNode, node_name);
                    Node::SetParent(lightNode, groupNode);
                }
        auto numbers = [&](const char* k, size_t n, double* out) {
            const lj::Value* a = member(node, k);
            if (!a || !a->is_array() || a->size() != n) return false;
            for (size_t i = 0; i < n; i++) out[i] = (*a->a)[i].as_double();
            return true;
        };
        double d[16];
        if (numbers("translation", 3, d)) groupNode->position = vec3((float)d[0], (float)d[1], (float)d[2]);
        if (numbers("rotation", 4, d)) {
            lm::quat q;
            q.w = (float)d[3], q.x = (float)d[0], q.y = (float)d[1], q.z = (float)d[2];
            groupNode->rotation = euler_degrees(q);
        }
        if (numbers("scale", 3, d)) groupNode->scale = vec3((float)d[0], (float)d[1], (float)d[2]);
        if (numbers("matrix", 16, d)) {
            lm::mat4 m(0.0f);
            for (int i = 0; i < 16; i++) m.data()[i] = (float)d[i];
            lm::quat q;
            q.w = q.x = q.y = q.z = 0.0f; // `glm::quat quat = {}` in the reference
            bool persp = false;
            decompose_affine(m, groupNode->scale, q, groupNode->position, persp);
            groupNode->rotation = euler_degrees(q);
        }
        loadedNodes.push_back(groupNode);
    }
    for (size_t ni = 0; ni < loadedNodes.size(); ni++) {
        const lj::Value* ch = member(*element(J, "nodes", (int64_t)ni), "children");
        for (size_t k = 0; ch && ch->is_array() && k < ch->size(); k++) {
            const int64_t c = (*ch->a)[k].as_int();
            if (c < 0 || (size_t)c >= loadedNodes.size()) { manager.error = "node child out of range"; return 0; }
            Node::SetParent(loadedNodes[(size_t)c], loadedNodes[ni]);
        }
    }

    std::vector<Ref<SceneAsset>> loadedScenes;
    for (size_t si = 0; si < array_size("scenes"); si++) {
        const lj::Value& scene = *element(J, "scenes", (int64_t)si);
        Ref<SceneAsset> s = create<SceneAsset>(manager, ObjectType::SceneAsset, str_or(scene, "name", ""));
        loadedScenes.push_back(s);
        const lj::Value* ns = member(scene, "nodes");
        for (size_t k = 0; ns && ns->is_array() && k < ns->size(); k++) {
            const int64_t n = (*ns->a)[k].as_int();
            if (n < 0 || (size_t)n >= loadedNodes.size()) { manager.error = "scene node out of range"; return 0; }
            s->Add(loadedNodes[(size_t)n]);
        }
    }
    return loadedScenes.size() ? loadedScenes[0]->uuid : 0;
}

UUID ImportSceneOBJ(const std::string& path, AssetManager& manager) {
    const std::string filename = stem(path);
    const std::string parentPath = parent_path(path) + "/";
    std::ifstream in(path);
    if (!in) {
        manager.error = "Failed to load obj file " + path;
        return 0;
    }
    // ---- parse (tiny_obj_loader LoadObj, triangulate = true) ----
    std::vector<float> v, vn, vt;
    std::vector<ObjShape> shapes;
    std::vector<ObjMaterial> objMaterials;
    std::map<std::string, int> material_map;
    ObjShape shape;
    std::vector<std::vector<ObjIndex>> faces;
    std::string name, line, err;
    int material = -1;
    while (std::getline(in, line)) {
        while (!line.empty() && (line.back() == '\n' || line.back() == '\r')) line.pop_back();
        const char* tok = line.c_str();
        tok += strspn(tok, " \t");
        if (tok[0] == '\0' || tok[0] == '#') continue;
        const auto space = [](char c) { return c == ' ' || c == '\t'; };
        if (tok[0] == 'v' && space(tok[1])) {
            tok += 2;
            for (int k = 0; k < 3; k++) v.push_back(parse_real(tok));
            continue;
        }
        if (tok[0] == 'v' && tok[1] == 'n' && space(tok[2])) {
            tok += 3;
            for (int k = 0; k < 3; k++) vn.push_back(parse_real(tok));
            continue;
        }
        if (tok[0] == 'v' && tok[1] == 't' && space(tok[2])) {
            tok += 3;
            for (int k = 0; k < 2; k++) vt.push_back(parse_real(tok));
            continue;
        }
        if (tok[0] == 'f' && space(tok[1])) {
            tok += 2;
            tok += strspn(tok, " \t");
            std::vector<ObjIndex> face;
            bool ok = true;
            while (tok[0] != '\0' && tok[0] != '\r' && tok[0] != '\n') {
                ObjIndex vi;
                if (!parse_triple(tok, (int)(v.size() / 3), (int)(vn.size() / 3), (int)(vt.size() / 2), vi)) {
                    ok = false;
                    break;
                }
                face.push_back(vi);
                tok += strspn(tok, " \t\r");
            }
            if (!ok) {
                manager.error = "Failed parse `f' line(e.g. zero value for face index)";
                return 0;
            }
            faces.push_back(std::move(face));
            continue;
        }
        if (strncmp(tok, "usemtl", 6) == 0) {
            tok += 6;
            const std::string namebuf = parse_string(tok);
            auto it = material_map.find(namebuf);
            const int newMaterialId = it != material_map.end() ? it->second : -1;
            if (newMaterialId != material) { // per-face materials: the shape goes on
                if (!export_faces(shape, faces, material, name, v, err) && !err.empty()) { manager.error = err; return 0; }
                faces.clear();
                material = newMaterialId;
            }
            continue;
        }
        if (strncmp(tok, "mtllib", 6) == 0 && space(tok[6])) {
            tok += 7;
            std::istringstream names(tok);
            std::string fn;
            while (names >> fn) {
                const size_t before = objMaterials.size();
                load_mtl(parentPath + fn, objMaterials, material_map);
                if (objMaterials.size() != before) break;
            }
            continue;
        }
        if ((tok[0] == 'g' || tok[0] == 'o') && space(tok[1])) {
            if (!export_faces(shape, faces, material, name, v, err) && !err.empty()) { manager.error = err; return 0; }
            if (!shape.indices.empty()) shapes.push_back(shape);
            shape = ObjShape();
            faces.clear();
            if (tok[0] == 'o') {
                name = std::string(tok + 2);
            } else {
                std::vector<std::string> names;
                while (tok[0] != '\0' && tok[0] != '\r' && tok[0] != '\n') {
                    names.push_back(parse_string(tok));
                    tok += strspn(tok, " \t\r");
                }
                name = "";
                for (size_t i = 1; i < names.size(); i++) name += (i > 1 ? " " : "") + names[i];
            }
            continue;
        }
        // s, l, p, t, vw, unknown: ignored
    }
    {
        const bool ret = export_faces(shape, faces, material, name, v, err);
        if (!ret && !err.empty()) { manager.error = err; return 0; }
        if (ret || !shape.indices.empty()) shapes.push_back(shape);
    }

    // ---- materials (AssetIO.cpp:425-458) ----
    std::vector<Ref<MaterialAsset>> materialAssets;
    std::map<std::string, Ref<TextureAsset>> textureAssets;
    auto texture_for = [&](const std::string& texname) -> Ref<TextureAsset> {
        auto it = textureAssets.find(texname);
        if (it != textureAssets.end()) return it->second;
        Ref<TextureAsset> t = create<TextureAsset>(manager, ObjectType::TextureAsset, texname);
        import_texture_into(parentPath + texname, t, manager); // failure leaves an empty texture and manager.error set
        textureAssets[texname] = t;
        return t;
    };
    for (const ObjMaterial& m : objMaterials) {
        Ref<MaterialAsset> asset = create<MaterialAsset>(manager, ObjectType::MaterialAsset, filename + ":" + m.name);
        asset->color = vec4(m.diffuse[0], m.diffuse[1], m.diffuse[2], 1.0f);
        asset->emission = vec3(m.emission[0], m.emission[1], m.emission[2]);
        asset->metallic = m.metallic;
        // `if (materials[i].specular != 0)` tests the address of the array in the reference: always true
        asset->roughness = 1.0f - (m.specular[0] + m.specular[1] + m.specular[2]) / 3.0f;
        if (m.diffuse_texname != "") asset->colorMap = texture_for(m.diffuse_texname);
        if (m.normal_texname != "") asset->normalMap = texture_for(m.normal_texname);
        materialAssets.push_back(asset);
    }

    // ---- shapes -> meshes and nodes (AssetIO.cpp:460-525) ----
    Ref<SceneAsset> scene = create<SceneAsset>(manager, ObjectType::SceneAsset, filename);
    Ref<Node> parentNode = create<Node>(manager, ObjectType::Node, filename);
    scene->Add(parentNode);
    for (const ObjShape& s : shapes) {
        if (s.indices.empty()) continue;
        std::map<std::array<uint32_t, 8>, uint32_t> uniqueVertices;
        const int splittedShapeIndex = 0; // never incremented in the reference
        size_t j = 0;
        long long lastMaterialId = s.material_ids.size() > 0 ? s.material_ids[0] : -1;
        Ref<MeshAsset> asset = create<MeshAsset>(manager, ObjectType::MeshAsset, filename + ":" + s.name);
        for (const ObjIndex& index : s.indices) {
            MeshAsset::MeshVertex vertex{};
            if (index.v < 0 || (size_t)(3 * index.v + 2) >= v.size()) { manager.error = "Vertex indices out of bounds"; return 0; }
            vertex.position = vec3(v[3 * (size_t)index.v], v[3 * (size_t)index.v + 1], v[3 * (size_t)index.v + 2]);
            if (index.vn != -1) {
                if (index.vn < 0 || (size_t)(3 * index.vn + 2) >= vn.size()) { manager.error = "Vertex normal indices out of bounds"; return 0; }
                vertex.normal = vec3(vn[3 * (size_t)index.vn], vn[3 * (size_t)index.vn + 1], vn[3 * (size_t)index.vn + 2]);
            }
            if (index.vt != -1) {
                if (index.vt < 0 || (size_t)(2 * index.vt + 1) >= vt.size()) { manager.error = "Vertex texcoord indices out of bounds"; return 0; }
                vertex.texCoord.x = vt[2 * (size_t)index.vt];
                vertex.texCoord.y = 1.0f - vt[2 * (size_t)index.vt + 1]; // the v-flip
            }
            // MeshVertex::operator== compares position, normal, texCoord (AssetManager.hpp:88-90): -0 == +0, NaN != NaN
            std::array<uint32_t, 8> key;
            const float comps[8] = {vertex.position.x, vertex.position.y, vertex.position.z, vertex.normal.x,
                                    vertex.normal.y,   vertex.normal.z,   vertex.texCoord.x, vertex.texCoord.y};
            bool nan = false;
            for (int k = 0; k < 8; k++) {
                const float c = comps[k] == 0.0f ? 0.0f : comps[k];
                nan = nan || c != c;
                memcpy(&key[k], &c, 4);
            }
            uint32_t at;
            auto it = nan ? uniqueVertices.end() : uniqueVertices.find(key);
            if (it == uniqueVertices.end()) {
                at = (uint32_t)asset->vertices.size();
                if (!nan) uniqueVertices[key] = at;
                asset->vertices.push_back(vertex);
            } else {
                at = it->second;
            }
            asset->indices.push_back(at);
            j += 1;
            if (j % 3 == 0) {
                const size_t faceId = j / 3;
                if (faceId >= s.material_ids.size() || s.material_ids[faceId] != lastMaterialId) {
                    asset->name += "_" + std::to_string(splittedShapeIndex);
                    Ref<MeshNode> model = create<MeshNode>(manager, ObjectType::MeshNode, asset->name);
                    Node::SetParent(model, parentNode);
                    model->mesh = asset;
                    if (lastMaterialId != -1 && (size_t)lastMaterialId < materialAssets.size()) model->material = materialAssets[(size_t)lastMaterialId];
                    if (faceId < s.material_ids.size()) lastMaterialId = s.material_ids[faceId];
                    uniqueVertices.clear();
                }
            }
        }
    }
    return scene->uuid;
}

// ---- dump (test hook; same layout as oracle/ref_import.cpp) -------------------------------------------------------
namespace {
struct Dumper {
    std::ostringstream o;
    std::vector<Ref<MeshAsset>> meshes;
    std::vector<Ref<MaterialAsset>> materials;
    std::vector<Ref<TextureAsset>> textures;
    template <class T>
    static int index_of(std::vector<Ref<T>>& v, const Ref<T>& p) {
        if (!p) return -1;
        for (size_t i = 0; i < v.size(); i++)
            if (v[i] == p) return (int)i;
        v.push_back(p);
        return (int)v.size() - 1;
    }
    static uint32_t bits(float f) {
        uint32_t u;
        memcpy(&u, &f, 4);
        return u;
    }
    void str(const std::string& s) {
        o << '"';
        for (char c : s) {
            if (c == '"' || c == '\\') o << '\\';
            o << c;
        }
        o << '"';
    }
    void vec(const char* name, const float* v, int n) {
        o << '"' << name << "\":[";
        for (int i = 0; i < n; i++) o << (i ? "," : "") << bits(v[i]);
        o << "]";
    }
    void node(const Ref<Node>& n) {
        o << "{\"name\":";
        str(n->name);
        o << ",\"type\":" << (int)n->type << ",";
        vec("position", &n->position.x, 3);
        o << ",";
        vec("rotation", &n->rotation.x, 3);
        o << ",";
        vec("scale", &n->scale.x, 3);
        int mesh = -1, material = -1;
        if (n->type == ObjectType::MeshNode) {
            auto mn = std::dynamic_pointer_cast<MeshNode>(n);
            mesh = index_of(meshes, mn->mesh);
            material = index_of(materials, mn->material);
        }
        o << ",\"mesh\":" << mesh << ",\"material\":" << material << ",\"children\":[";
        for (size_t i = 0; i < n->children.size(); i++) {
            if (i) o << ",";
            node(n->children[i]);
        }
        o << "]}";
    }
};
} // namespace

std::string DumpImportedScene(AssetManager& manager, UUID id) {
    Ref<SceneAsset> scene = manager.Get<SceneAsset>(id);
    if (!scene) return "{\"error\":\"import failed\"}\n";
    Dumper d;
    d.o << "{\"scene\":";
    d.str(scene->name);
    d.o << ",\"nodes\":[";
    for (size_t i = 0; i < scene->nodes.size(); i++) {
        if (i) d.o << ",";
        d.node(scene->nodes[i]);
    }
    d.o << "],\"materials\":[";
    for (size_t i = 0; i < d.materials.size(); i++) {
        const auto m = d.materials[i];
        if (i) d.o << ",";
        d.o << "{\"name\":";
        d.str(m->name);
        d.o << ",";
        d.vec("color", &m->color.x, 4);
        d.o << ",";
        d.vec("emission", &m->emission.x, 3);
        d.o << ",\"metallic\":" << Dumper::bits(m->metallic) << ",\"roughness\":" << Dumper::bits(m->roughness);
        const int ao = Dumper::index_of(d.textures, m->aoMap), col = Dumper::index_of(d.textures, m->colorMap);
        const int nrm = Dumper::index_of(d.textures, m->normalMap), emi = Dumper::index_of(d.textures, m->emissionMap);
        const int mr = Dumper::index_of(d.textures, m->metallicRoughnessMap);
        d.o << ",\"aoMap\":" << ao << ",\"colorMap\":" << col << ",\"normalMap\":" << nrm << ",\"emissionMap\":" << emi
            << ",\"metallicRoughnessMap\":" << mr << "}";
    }
    d.o << "],\"textures\":[";
    for (size_t i = 0; i < d.textures.size(); i++) {
        const auto& t = d.textures[i];
        if (i) d.o << ",";
        unsigned long long h = 1469598103934665603ull;
        for (uint8_t b : t->data) h = (h ^ b) * 1099511628211ull;
        char hex[32];
        snprintf(hex, sizeof hex, "%016llx", h);
        d.o << "{\"name\":";
        d.str(t->name);
        d.o << ",\"width\":" << t->width << ",\"height\":" << t->height << ",\"channels\":" << t->channels
            << ",\"bytes\":" << t->data.size() << ",\"fnv1a\":\"" << hex << "\"}";
    }
    d.o << "],\"meshes\":[";
    for (size_t i = 0; i < d.meshes.size(); i++) {
        const auto& m = d.meshes[i];
        if (i) d.o << ",";
        d.o << "{\"name\":";
        d.str(m->name);
        d.o << ",\"vertex_count\":" << m->vertices.size() << ",\"vertices\":[";
        const float* f = reinterpret_cast<const float*>(m->vertices.data());
        for (size_t k = 0; k < m->vertices.size() * 12; k++) d.o << (k ? "," : "") << Dumper::bits(f[k]);
        d.o << "],\"indices\":[";
        for (size_t k = 0; k < m->indices.size(); k++) d.o << (k ? "," : "") << m->indices[k];
        d.o << "]}";
    }
    d.o << "]}\n";
    return d.o.str();
}

} // namespace AssetIO
} // namespace luzhost

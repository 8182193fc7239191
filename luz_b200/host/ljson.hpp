// ljson.hpp -- a small JSON value + parser + writer for .luz project files.
// The reference uses nlohmann::json (deps/json.hpp); this is an independent minimal
// implementation with the properties the .luz format needs: exact 64-bit integers (uuids are u64
// in [2^61, 2^62], AssetManager.cpp:268-274), objects with alphabetically ordered keys (what
// nlohmann's default std::map storage produces on dump), doubles printed round-trip exact.
#pragma once

#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace lj {

struct Value;
using Array = std::vector<Value>;
using Object = std::map<std::string, Value>;

struct Value {
    enum Kind { Null, Bool, Int, UInt, Double, String, Arr, Obj } kind = Null;
    bool b = false;
    int64_t i = 0;
    uint64_t u = 0;
    double d = 0.0;
    std::string s;
    std::shared_ptr<Array> a;
    std::shared_ptr<Object> o;

    Value() = default;
    static Value boolean(bool v) { Value r; r.kind = Bool; r.b = v; return r; }
    static Value integer(int64_t v) { Value r; r.kind = Int; r.i = v; return r; }
    static Value uinteger(uint64_t v) { Value r; r.kind = UInt; r.u = v; return r; }
    static Value number(double v) { Value r; r.kind = Double; r.d = v; return r; }
    static Value string(const std::string& v) { Value r; r.kind = String; r.s = v; return r; }
    static Value array() { Value r; r.kind = Arr; r.a = std::make_shared<Array>(); return r; }
    static Value object() { Value r; r.kind = Obj; r.o = std::make_shared<Object>(); return r; }

    bool is_number() const { return kind == Int || kind == UInt || kind == Double; }
    bool is_array() const { return kind == Arr; }
    bool is_object() const { return kind == Obj; }
    double as_double() const { return kind == Double ? d : kind == Int ? (double)i : kind == UInt ? (double)u : 0.0; }
    int64_t as_int() const { return kind == Int ? i : kind == UInt ? (int64_t)u : kind == Double ? (int64_t)d : (kind == Bool ? (b ? 1 : 0) : 0); }
    uint64_t as_uint() const { return kind == UInt ? u : kind == Int ? (uint64_t)i : kind == Double ? (uint64_t)d : 0; }
    bool as_bool() const { return kind == Bool ? b : as_int() != 0; }
    bool contains(const std::string& k) const { return kind == Obj && o->find(k) != o->end(); }
    const Value& at(const std::string& k) const;
    Value& operator[](const std::string& k); // creates (object)
    size_t size() const { return kind == Arr ? a->size() : kind == Obj ? o->size() : 0; }
};

// Throws std::runtime_error with position on malformed input.
Value parse(const std::string& text);
std::string dump(const Value& v);

} // namespace lj

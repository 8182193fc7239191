// json.cpp -- parser / writer of ljson.hpp.
#include "ljson.hpp"

#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>

namespace lj {

const Value& Value::at(const std::string& k) const {
    static const Value null_value;
    if (kind != Obj) return null_value;
    auto it = o->find(k);
    return it == o->end() ? null_value : it->second;
}

Value& Value::operator[](const std::string& k) {
    if (kind != Obj) {
        *this = Value::object();
    }
    return (*o)[k];
}

namespace {

struct Parser {
    const char* p;
    const char* end;
    const char* begin;

    [[noreturn]] void fail(const char* what) {
        char buf[128];
        snprintf(buf, sizeof(buf), "json: %s at offset %zu", what, (size_t)(p - begin));
        throw std::runtime_error(buf);
    }
    void ws() {
        while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) p++;
    }
    bool lit(const char* s) {
        size_t n = strlen(s);
        if ((size_t)(end - p) >= n && memcmp(p, s, n) == 0) {
            p += n;
            return true;
        }
        return false;
    }
    static void put_utf8(std::string& out, unsigned cp) {
        if (cp < 0x80) out += (char)cp;
        else if (cp < 0x800) {
            out += (char)(0xC0 | (cp >> 6));
            out += (char)(0x80 | (cp & 0x3F));
        } else if (cp < 0x10000) {
            out += (char)(0xE0 | (cp >> 12));
            out += (char)(0x80 | ((cp >> 6) & 0x3F));
            out += (char)(0x80 | (cp & 0x3F));
        } else {
            out += (char)(0xF0 | (cp >> 18));
            out += (char)(0x80 | ((cp >> 12) & 0x3F));
            out += (char)(0x80 | ((cp >> 6) & 0x3F));
            out += (char)(0x80 | (cp & 0x3F));
        }
    }
    unsigned hex4() {
        if (end - p < 4) fail("bad \\u escape");
        unsigned v = 0;
        for (int k = 0; k < 4; k++) {
            char c = *p++;
            v <<= 4;
            if (c >= '0' && c <= '9') v |= (unsigned)(c - '0');
            else if (c >= 'a' && c <= 'f') v |= (unsigned)(c - 'a' + 10);
            else if (c >= 'A' && c <= 'F') v |= (unsigned)(c - 'A' + 10);
            else fail("bad hex digit");
        }
        return v;
    }
    std::string str() {
        if (p >= end || *p != '"') fail("expected string");
        p++;
        std::string out;
        while (true) {
            if (p >= end) fail("unterminated string");
            char c = *p++;
            if (c == '"') break;
            if (c == '\\') {
                if (p >= end) fail("bad escape");
                char e = *p++;
                switch (e) {
                    case '"': out += '"'; break;
                    case '\\': out += '\\'; break;
                    case '/': out += '/'; break;
                    case 'b': out += '\b'; break;
                    case 'f': out += '\f'; break;
                    case 'n': out += '\n'; break;
                    case 'r': out += '\r'; break;
                    case 't': out += '\t'; break;
                    case 'u': {
                        unsigned cp = hex4();
                        if (cp >= 0xD800 && cp <= 0xDBFF && end - p >= 6 && p[0] == '\\' && p[1] == 'u') {
                            p += 2;
                            unsigned lo = hex4();
                            cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                        }
                        put_utf8(out, cp);
                        break;
                    }
                    default: fail("unknown escape");
                }
            } else {
                out += c;
            }
        }
        return out;
    }
    Value number() {
        const char* s = p;
        bool neg = false, is_float = false;
        if (p < end && *p == '-') {
            neg = true;
            p++;
        }
        if (p >= end || !(*p >= '0' && *p <= '9')) fail("bad number");
        while (p < end && *p >= '0' && *p <= '9') p++;
        if (p < end && *p == '.') {
            is_float = true;
            p++;
            while (p < end && *p >= '0' && *p <= '9') p++;
        }
        if (p < end && (*p == 'e' || *p == 'E')) {
            is_float = true;
            p++;
            if (p < end && (*p == '+' || *p == '-')) p++;
            while (p < end && *p >= '0' && *p <= '9') p++;
        }
        std::string tok(s, p);
        if (!is_float) {
            errno = 0;
            if (neg) {
                long long v = strtoll(tok.c_str(), nullptr, 10);
                if (errno == 0) return Value::integer(v);
            } else {
                unsigned long long v = strtoull(tok.c_str(), nullptr, 10);
                if (errno == 0) return Value::uinteger(v);
            }
        }
        return Value::number(strtod(tok.c_str(), nullptr));
    }
    Value value(int depth) {
        if (depth > 256) fail("nesting too deep");
        ws();
        if (p >= end) fail("unexpected end");
        char c = *p;
        if (c == '{') {
            p++;
            Value v = Value::object();
            ws();
            if (p < end && *p == '}') {
                p++;
                return v;
            }
            while (true) {
                ws();
                std::string k = str();
                ws();
                if (p >= end || *p != ':') fail("expected ':'");
                p++;
                (*v.o)[k] = value(depth + 1);
                ws();
                if (p < end && *p == ',') {
                    p++;
                    continue;
                }
                if (p < end && *p == '}') {
                    p++;
                    break;
                }
                fail("expected ',' or '}'");
            }
            return v;
        }
        if (c == '[') {
            p++;
            Value v = Value::array();
            ws();
            if (p < end && *p == ']') {
                p++;
                return v;
            }
            while (true) {
                v.a->push_back(value(depth + 1));
                ws();
                if (p < end && *p == ',') {
                    p++;
                    continue;
                }
                if (p < end && *p == ']') {
                    p++;
                    break;
                }
                fail("expected ',' or ']'");
            }
            return v;
        }
        if (c == '"') return Value::string(str());
        if (lit("true")) return Value::boolean(true);
        if (lit("false")) return Value::boolean(false);
        if (lit("null")) return Value();
        return number();
    }
};

void dump_string(const std::string& s, std::string& out) {
    out += '"';
    for (unsigned char c : s) {
        switch (c) {
            case '"': out += "\\\""; break;
            case '\\': out += "\\\\"; break;
            case '\b': out += "\\b"; break;
            case '\f': out += "\\f"; break;
            case '\n': out += "\\n"; break;
            case '\r': out += "\\r"; break;
            case '\t': out += "\\t"; break;
            default:
                if (c < 0x20) {
                    char buf[8];
                    snprintf(buf, sizeof(buf), "\\u%04x", c);
                    out += buf;
                } else {
                    out += (char)c;
                }
        }
    }
    out += '"';
}

void dump_double(double d, std::string& out) {
    if (!std::isfinite(d)) {
        out += "null";
        return;
    }
    char buf[40];
    // shortest representation that round-trips
    for (int prec = 1; prec <= 17; prec++) {
        snprintf(buf, sizeof(buf), "%.*g", prec, d);
        if (strtod(buf, nullptr) == d) break;
    }
    out += buf;
    if (!strpbrk(buf, ".eE")) out += ".0";
}

void dump_value(const Value& v, std::string& out) {
    char buf[32];
    switch (v.kind) {
        case Value::Null: out += "null"; break;
        case Value::Bool: out += v.b ? "true" : "false"; break;
        case Value::Int:
            snprintf(buf, sizeof(buf), "%lld", (long long)v.i);
            out += buf;
            break;
        case Value::UInt:
            snprintf(buf, sizeof(buf), "%llu", (unsigned long long)v.u);
            out += buf;
            break;
        case Value::Double: dump_double(v.d, out); break;
        case Value::String: dump_string(v.s, out); break;
        case Value::Arr: {
            out += '[';
            bool first = true;
            for (const Value& e : *v.a) {
                if (!first) out += ',';
                first = false;
                dump_value(e, out);
            }
            out += ']';
            break;
        }
        case Value::Obj: {
            out += '{';
            bool first = true;
            for (const auto& kv : *v.o) {
                if (!first) out += ',';
                first = false;
                dump_string(kv.first, out);
                out += ':';
                dump_value(kv.second, out);
            }
            out += '}';
            break;
        }
    }
}

} // namespace

Value parse(const std::string& text) {
    Parser ps{text.data(), text.data() + text.size(), text.data()};
    Value v = ps.value(0);
    ps.ws();
    if (ps.p != ps.end) ps.fail("trailing characters");
    return v;
}

std::string dump(const Value& v) {
    std::string out;
    dump_value(v, out);
    return out;
}

} // namespace lj

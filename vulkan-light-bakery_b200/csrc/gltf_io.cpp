// gltf_io.cpp — on-disk format of the baked light: writer = LightBaker::serialize
// (src/baker/light_baker.cpp:375-402, base64 at :330-373), reader = Scene_t::loadBakedLight
// (src/scene_manager.cpp:613-648, base64 at :556-611). Host-only, untimed.
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>

#include "vlb_context.h"
#include "vlb_json.h"

namespace vlb {

// ---------------------------------------------------------------- JSON ---------------------
namespace {
struct Parser {
    const std::string& t;
    size_t p = 0;
    explicit Parser(const std::string& text) : t(text) {}
    [[noreturn]] void fail(const char* what) {
        throw std::runtime_error(std::string("json: ") + what + " at byte " + std::to_string(p));
    }
    void ws() { while (p < t.size() && (t[p] == ' ' || t[p] == '\t' || t[p] == '\n' || t[p] == '\r')) ++p; }
    bool eat(char c) { ws(); if (p < t.size() && t[p] == c) { ++p; return true; } return false; }
    void append_utf8(std::string& out, unsigned cp) {
        if (cp < 0x80) out += (char)cp;
        else if (cp < 0x800) { out += (char)(0xC0 | (cp >> 6)); out += (char)(0x80 | (cp & 0x3F)); }
        else if (cp < 0x10000) { out += (char)(0xE0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
        else { out += (char)(0xF0 | (cp >> 18)); out += (char)(0x80 | ((cp >> 12) & 0x3F)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
    }
    unsigned hex4() {
        if (p + 4 > t.size()) fail("truncated \\u escape");
        unsigned v = 0;
        for (int i = 0; i < 4; ++i) {
            const char c = t[p++];
            v <<= 4;
            if (c >= '0' && c <= '9') v |= c - '0';
            else if (c >= 'a' && c <= 'f') v |= c - 'a' + 10;
            else if (c >= 'A' && c <= 'F') v |= c - 'A' + 10;
            else fail("bad \\u escape");
        }
        return v;
    }
    std::string str() {
        if (!eat('"')) fail("expected string");
        std::string out;
        for (;;) {
            if (p >= t.size()) fail("unterminated string");
            const char c = t[p++];
            if (c == '"') break;
            if (c != '\\') { out += c; continue; }
            if (p >= t.size()) fail("unterminated escape");
            const char e = t[p++];
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
                    if (cp >= 0xD800 && cp <= 0xDBFF && p + 1 < t.size() && t[p] == '\\' && t[p + 1] == 'u') {
                        p += 2;
                        const unsigned lo = hex4();
                        cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                    }
                    append_utf8(out, cp);
                    break;
                }
                default: fail("bad escape");
            }
        }
        return out;
    }
    Json value(int depth) {
        if (depth > 256) fail("nesting too deep");
        ws();
        if (p >= t.size()) fail("unexpected end");
        const char c = t[p];
        Json j;
        if (c == '{') {
            ++p; j.type = Json::Object;
            if (eat('}')) return j;
            do {
                ws();
                std::string k = str();
                if (!eat(':')) fail("expected ':'");
                j.o[k] = value(depth + 1);
            } while (eat(','));
            if (!eat('}')) fail("expected '}'");
        } else if (c == '[') {
            ++p; j.type = Json::Array;
            if (eat(']')) return j;
            do { j.a.push_back(value(depth + 1)); } while (eat(','));
            if (!eat(']')) fail("expected ']'");
        } else if (c == '"') {
            j.type = Json::String; j.s = str();
        } else if (t.compare(p, 4, "true") == 0) { p += 4; j.type = Json::Bool; j.b = true; }
        else if (t.compare(p, 5, "false") == 0) { p += 5; j.type = Json::Bool; j.b = false; }
        else if (t.compare(p, 4, "null") == 0) { p += 4; j.type = Json::Null; }
        else {
            const size_t s0 = p;
            if (p < t.size() && (t[p] == '-' || t[p] == '+')) ++p;
            while (p < t.size() && (isdigit((unsigned char)t[p]) || t[p] == '.' || t[p] == 'e' || t[p] == 'E' || t[p] == '-' || t[p] == '+')) ++p;
            if (p == s0) fail("unexpected character");
            j.type = Json::Number; j.s = t.substr(s0, p - s0);
            char* end = nullptr;
            strtod(j.s.c_str(), &end);
            if (!end || *end) fail("bad number");
        }
        return j;
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
                if (c < 0x20) { char b[8]; snprintf(b, sizeof b, "\\u%04x", c); out += b; }
                else out += (char)c;
        }
    }
    out += '"';
}

void dump_rec(const Json& j, int indent, int level, std::string& out) {
    const std::string pad(indent > 0 ? (size_t)indent * (level + 1) : 0, ' ');
    const std::string pad_close(indent > 0 ? (size_t)indent * level : 0, ' ');
    const char* nl = indent > 0 ? "\n" : "";
    switch (j.type) {
        case Json::Null: out += "null"; break;
        case Json::Bool: out += j.b ? "true" : "false"; break;
        case Json::Number: out += j.s; break;
        case Json::String: dump_string(j.s, out); break;
        case Json::Array:
            if (j.a.empty()) { out += "[]"; break; }
            out += "["; out += nl;
            for (size_t i = 0; i < j.a.size(); ++i) {
                out += pad; dump_rec(j.a[i], indent, level + 1, out);
                if (i + 1 < j.a.size()) out += ",";
                out += nl;
            }
            out += pad_close; out += "]";
            break;
        case Json::Object: {
            if (j.o.empty()) { out += "{}"; break; }
            out += "{"; out += nl;
            size_t i = 0;
            for (const auto& kv : j.o) {
                out += pad; dump_string(kv.first, out); out += indent > 0 ? ": " : ":";
                dump_rec(kv.second, indent, level + 1, out);
                if (++i < j.o.size()) out += ",";
                out += nl;
            }
            out += pad_close; out += "}";
            break;
        }
    }
}
}  // namespace

Json Json::number(double v) {
    Json j; j.type = Number;
    if (!std::isfinite(v)) { j.type = Null; return j; }
    char buf[40];
    for (int prec = 1; prec <= 17; ++prec) {      // shortest representation that round-trips
        snprintf(buf, sizeof buf, "%.*g", prec, v);
        if (strtod(buf, nullptr) == v) break;
    }
    j.s = buf;
    if (j.s.find_first_of(".eEn") == std::string::npos) j.s += ".0";   // keep it a float token
    return j;
}
Json Json::integer(long long v) { Json j; j.type = Number; j.s = std::to_string(v); return j; }
double Json::num() const { return type == Number ? strtod(s.c_str(), nullptr) : 0.0; }
long long Json::integer_value() const { return type == Number ? (long long)strtod(s.c_str(), nullptr) : 0; }

Json json_parse(const std::string& text) {
    Parser ps(text);
    Json j = ps.value(0);
    ps.ws();
    if (ps.p != text.size()) ps.fail("trailing characters");
    return j;
}
std::string json_dump(const Json& j, int indent) {
    std::string out;
    dump_rec(j, indent, 0, out);
    return out;
}

// ---------------------------------------------------------------- base64 -------------------
static const char kB64[] = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";

std::string base64_encode(const uint8_t* d, size_t n) {
    std::string out;
    out.reserve((n + 2) / 3 * 4);
    size_t i = 0;
    for (; i + 3 <= n; i += 3) {
        const uint32_t v = (d[i] << 16) | (d[i + 1] << 8) | d[i + 2];
        out += kB64[(v >> 18) & 63]; out += kB64[(v >> 12) & 63]; out += kB64[(v >> 6) & 63]; out += kB64[v & 63];
    }
    if (n - i == 1) {
        const uint32_t v = d[i] << 16;
        out += kB64[(v >> 18) & 63]; out += kB64[(v >> 12) & 63]; out += "==";
    } else if (n - i == 2) {
        const uint32_t v = (d[i] << 16) | (d[i + 1] << 8);
        out += kB64[(v >> 18) & 63]; out += kB64[(v >> 12) & 63]; out += kB64[(v >> 6) & 63]; out += '=';
    }
    return out;
}

std::vector<uint8_t> base64_decode(const std::string& s) {
    int8_t lut[256];
    memset(lut, -1, sizeof lut);
    for (int i = 0; i < 64; ++i) lut[(unsigned char)kB64[i]] = (int8_t)i;
    std::vector<uint8_t> out;
    out.reserve(s.size() / 4 * 3);
    uint32_t acc = 0; int bits = 0;
    for (unsigned char c : s) {
        if (c == '=') break;                 // the reference's decoder stops at the first '='
        const int v = lut[c];
        if (v < 0) break;                    // ... or at the first non-alphabet character
        acc = (acc << 6) | (uint32_t)v; bits += 6;
        if (bits >= 8) { bits -= 8; out.push_back((uint8_t)((acc >> bits) & 0xFF)); }
    }
    return out;
}

static bool read_file(const char* path, std::string& out) {
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    std::ostringstream ss; ss << f.rdbuf();
    out = ss.str();
    return true;
}

}  // namespace vlb

using namespace vlb;

extern "C" {

static const char kHeader[] = "data:application/octet-stream;base64,";

int vlb_bake_serialize_gltf(const char* in_path, const char* out_path, const float* coeffs, uint64_t n_probes,
                            const vlb_bake_settings* s) {
    if (!in_path || !out_path || !coeffs || !s) { set_thread_error("vlb_bake_serialize_gltf: NULL argument"); return VLB_ERR_INVALID; }
    std::string text;
    if (!read_file(in_path, text)) { set_thread_error("vlb_bake_serialize_gltf: cannot read the input glTF"); return VLB_ERR_IO; }
    try {
        Json json = json_parse(text);
        if (!json.is(Json::Object)) throw std::runtime_error("json: glTF root is not an object");
        const uint64_t bytes = n_probes * VLB_SH_STRIDE * sizeof(float);
        // light_baker.cpp:381-395
        Json buffer = Json::object();
        buffer["byteLength"] = Json::integer((long long)bytes);
        buffer["uri"] = Json::string(std::string(kHeader) + base64_encode(reinterpret_cast<const uint8_t*>(coeffs), bytes));
        Json& buffers = json["buffers"];
        if (buffers.is(Json::Null)) buffers = Json::array();
        buffers.a.push_back(buffer);
        Json view = Json::object();
        view["buffer"] = Json::integer((long long)buffers.a.size() - 1);
        view["byteLength"] = Json::integer((long long)bytes);
        view["byteOffset"] = Json::integer(0);
        Json& views = json["bufferViews"];
        if (views.is(Json::Null)) views = Json::array();
        views.a.push_back(view);
        Json& light = json["light"];
        light = Json::object();
        Json step = Json::object();                               // glm::to_json, light_baker.cpp:12-15
        step["x"] = Json::number((double)s->step[0]); step["y"] = Json::number((double)s->step[1]); step["z"] = Json::number((double)s->step[2]);
        light["gridStep"] = step;
        light["lmax"] = Json::integer(16);                        // light_baker.cpp:396 (coefficient count)
        light["bufferView"] = Json::integer((long long)views.a.size() - 1);
        // extra keys the reference reader ignores: what a correct consumer needs
        Json org = Json::object();
        org["x"] = Json::number((double)s->origin[0]); org["y"] = Json::number((double)s->origin[1]); org["z"] = Json::number((double)s->origin[2]);
        light["origin"] = org;
        Json cnt = Json::object();
        cnt["x"] = Json::integer(s->probes[0]); cnt["y"] = Json::integer(s->probes[1]); cnt["z"] = Json::integer(s->probes[2]);
        light["count"] = cnt;
        light["order"] = Json::integer(s->sh_order);
        light["stride"] = Json::integer(16);
        light["probeOrder"] = Json::string((s->flags & VLB_BAKE_REFERENCE_PROBE_ORDER) ? "reference" : "x-fastest");
        std::ofstream o(out_path, std::ios::binary);
        if (!o) { set_thread_error("vlb_bake_serialize_gltf: cannot open the output file"); return VLB_ERR_IO; }
        o << json_dump(json, 4) << std::endl;                     // std::setw(4), light_baker.cpp:400
        if (!o) { set_thread_error("vlb_bake_serialize_gltf: write failed"); return VLB_ERR_IO; }
    } catch (const std::exception& e) {
        set_thread_error(e.what());
        return VLB_ERR_IO;
    }
    return VLB_OK;
}

int vlb_bake_deserialize_gltf(const char* path, float* coeffs, uint64_t capacity_floats, uint64_t* n_floats_out,
                              float grid_step_out[3]) {
    if (!path) { set_thread_error("vlb_bake_deserialize_gltf: NULL path"); return VLB_ERR_INVALID; }
    std::string text;
    if (!read_file(path, text)) { set_thread_error("vlb_bake_deserialize_gltf: cannot read the file"); return VLB_ERR_IO; }
    try {
        const Json json = json_parse(text);
        const Json* light = json.find("light");
        if (!light || !light->is(Json::Object)) { set_thread_error("no \"light\" object in the glTF"); return VLB_ERR_STATE; }
        const Json* step = light->find("gridStep");
        if (grid_step_out && step) {
            const char* k[3] = {"x", "y", "z"};
            for (int i = 0; i < 3; ++i) { const Json* c = step->find(k[i]); grid_step_out[i] = c ? (float)c->num() : 0.f; }
        }
        const Json* bv = light->find("bufferView");
        const Json* views = json.find("bufferViews");
        const Json* buffers = json.find("buffers");
        if (!bv || !views || !buffers) throw std::runtime_error("light.bufferView / bufferViews / buffers missing");
        const long long vi = bv->integer_value();
        if (vi < 0 || (size_t)vi >= views->a.size()) throw std::runtime_error("light.bufferView out of range");
        const Json& view = views->a[(size_t)vi];
        const Json* bi = view.find("buffer");
        if (!bi || bi->integer_value() < 0 || (size_t)bi->integer_value() >= buffers->a.size())
            throw std::runtime_error("bufferView.buffer out of range");
        const Json& buffer = buffers->a[(size_t)bi->integer_value()];
        const Json* uri = buffer.find("uri");
        const Json* len = buffer.find("byteLength");
        if (!uri || !uri->is(Json::String) || !len) throw std::runtime_error("buffer.uri / byteLength missing");
        std::string u = uri->s;
        if (u.compare(0, sizeof(kHeader) - 1, kHeader) != 0) throw std::runtime_error("buffer.uri is not an octet-stream data URI");
        u.erase(0, sizeof(kHeader) - 1);
        const std::vector<uint8_t> bytes = base64_decode(u);
        const uint64_t size = (uint64_t)len->integer_value();
        if (bytes.size() < size) throw std::runtime_error("decoded buffer shorter than byteLength");
        if (n_floats_out) *n_floats_out = size / sizeof(float);
        if (coeffs) {
            if (capacity_floats < size / sizeof(float)) { set_thread_error("vlb_bake_deserialize_gltf: output too small"); return VLB_ERR_INVALID; }
            memcpy(coeffs, bytes.data(), size / sizeof(float) * sizeof(float));
        }
    } catch (const std::exception& e) {
        set_thread_error(e.what());
        return VLB_ERR_IO;
    }
    return VLB_OK;
}

}  // extern "C"

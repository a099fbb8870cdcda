// vlb_json.h — minimal JSON DOM (parse + dump) for the glTF side of the bake path.
// The reference uses nlohmann::json (src/baker/light_baker.cpp:375-402); that library is not part
// of this build, so this is a from-scratch reader/writer with the same observable output style:
// object keys sorted (nlohmann's default std::map), `dump(4)` indentation, untouched numbers
// re-emitted verbatim.
#pragma once

#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace vlb {

struct Json {
    enum Type { Null, Bool, Number, String, Array, Object } type = Null;
    bool b = false;
    std::string s;                       // String value, or the raw token of a Number
    std::vector<Json> a;
    std::map<std::string, Json> o;

    static Json number(double v);
    static Json integer(long long v);
    static Json string(const std::string& v) { Json j; j.type = String; j.s = v; return j; }
    static Json array() { Json j; j.type = Array; return j; }
    static Json object() { Json j; j.type = Object; return j; }

    bool is(Type t) const { return type == t; }
    double num() const;
    long long integer_value() const;
    const Json* find(const std::string& k) const {
        if (type != Object) return nullptr;
        auto it = o.find(k);
        return it == o.end() ? nullptr : &it->second;
    }
    Json& operator[](const std::string& k) { if (type == Null) type = Object; return o[k]; }
};

// Throws std::runtime_error with a position on malformed input.
Json json_parse(const std::string& text);
std::string json_dump(const Json& j, int indent);

std::string base64_encode(const uint8_t* data, size_t n);
// Ignores characters outside the alphabet (as the reference's decoder stops at '=').
std::vector<uint8_t> base64_decode(const std::string& s);

}  // namespace vlb

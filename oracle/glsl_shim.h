// glsl_shim.h — just enough GLSL for the REFERENCE's own shader sources to compile as C++ (TEST INFRASTRUCTURE).
//
// oracle/make_ref_shaders.py takes the shader files from where they lie under /root/reference/shaders, strips
// only what is interface syntax (#version / #extension lines, `layout(...)` declarations, hitAttributeEXT) and
// renames main() -> shader_main(); every function body (sRGB, getBaseColor, dir2SkyboxUV, the main() of
// env_map.rgen / env_map.rchit / main.rmiss / shadow.rmiss / sh.comp / skybox_sh.comp and the whole of sh_common.h
// and structures.h) is compiled from the reference's text behind this header. Outputs go to oracle/_ref/ only.
// Nothing here is product code; the product never links it.
//
// Semantics notes (what GLSL leaves to the implementation is called out, because parity stays unpinned there):
//   * floats are IEEE binary32, no contraction (-ffp-contract=off), literals are fp32 (-fsingle-precision-constant);
//   * sin/cos/pow/acos/atan are libm's float versions (GLSL only bounds their error);
//   * texture(): bilinear, repeat, LOD 0 (the reference's sampler: src/application.hpp:45-52; anisotropy ignored);
//   * imageStore to an rgba8 image: clamp + round-to-nearest-even to n/255 (Vulkan float -> UNORM conversion);
//   * traceRayEXT: forwarded to a callback (intersection and acceleration structure are the driver's in the reference).
#pragma once

#include <cmath>
#include <cstdint>

namespace glsl {

typedef unsigned int uint;

struct vec2; struct vec3; struct vec4; struct ivec2; struct ivec3; struct uvec2;

// swizzle proxies: overlay the owning vector's storage (N floats), convert on read
template <int N, int A, int B> struct sw2 { float d[N]; inline operator vec2() const; };
template <int N, int A, int B, int C> struct sw3 { float d[N]; inline operator vec3() const; };
template <int N, int A, int B> struct usw2 { uint d[N]; inline operator uvec2() const; };

struct vec2 {
    union { struct { float x, y; }; struct { float r, g; }; };
    vec2() : x(0), y(0) {}
    explicit vec2(float a) : x(a), y(a) {}
    vec2(float a, float b) : x(a), y(b) {}
};
struct vec3 {
    union {
        struct { float x, y, z; }; struct { float r, g, b; };
        sw3<3, 0, 1, 2> xyz; sw3<3, 0, 1, 2> rgb; sw3<3, 0, 2, 1> xzy; sw2<3, 0, 1> xy;
    };
    vec3() : x(0), y(0), z(0) {}
    explicit vec3(float a) : x(a), y(a), z(a) {}
    vec3(float a, float b, float c) : x(a), y(b), z(c) {}
    explicit inline vec3(const vec4& v);
};
struct vec4 {
    union {
        struct { float x, y, z, w; }; struct { float r, g, b, a; };
        sw3<4, 0, 1, 2> xyz; sw3<4, 0, 1, 2> rgb; sw3<4, 0, 2, 1> xzy; sw2<4, 0, 1> xy;
    };
    vec4() : x(0), y(0), z(0), w(0) {}
    explicit vec4(float s) : x(s), y(s), z(s), w(s) {}
    vec4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
    vec4(const vec3& v, float d) : x(v.x), y(v.y), z(v.z), w(d) {}
};
inline vec3::vec3(const vec4& v) : x(v.x), y(v.y), z(v.z) {}
template <int N, int A, int B> inline sw2<N, A, B>::operator vec2() const { return vec2(d[A], d[B]); }
template <int N, int A, int B, int C> inline sw3<N, A, B, C>::operator vec3() const { return vec3(d[A], d[B], d[C]); }
static_assert(sizeof(vec2) == 8 && sizeof(vec3) == 12 && sizeof(vec4) == 16, "scalar block layout sizes");

struct uvec2 { uint x, y; uvec2(uint a, uint b) : x(a), y(b) {} };
template <int N, int A, int B> inline usw2<N, A, B>::operator uvec2() const { return uvec2(d[A], d[B]); }
struct uvec3 {
    union { struct { uint x, y, z; }; usw2<3, 0, 1> xy; };
    uvec3() : x(0), y(0), z(0) {}
};
struct ivec2 {
    int x, y;
    ivec2() : x(0), y(0) {}
    ivec2(int a, int b) : x(a), y(b) {}
    ivec2(uint a, uint b) : x((int)a), y((int)b) {}
    explicit ivec2(const uvec2& v) : x((int)v.x), y((int)v.y) {}
};
struct ivec3 {
    int x, y, z;
    ivec3() : x(0), y(0), z(0) {}
    ivec3(int a, int b, int c) : x(a), y(b), z(c) {}
    explicit ivec3(const vec3& v) : x((int)v.x), y((int)v.y), z((int)v.z) {}      // GLSL float -> int: truncation
};
static_assert(sizeof(ivec3) == 12, "Indices buffer stride");
struct bvec4 { bool x, y, z, w; };
struct bvec3 { bool x, y, z; };

// ---- arithmetic -------------------------------------------------------------------------------------------
#define GLSL_OPS(V, EXPR2, EXPRS, EXPRS_L)                                                     \
    inline V operator+(const V& a, const V& b) { return EXPR2(+); }                            \
    inline V operator-(const V& a, const V& b) { return EXPR2(-); }                            \
    inline V operator*(const V& a, const V& b) { return EXPR2(*); }                            \
    inline V operator/(const V& a, const V& b) { return EXPR2(/); }                            \
    inline V operator*(const V& a, float s) { return EXPRS(*); }                               \
    inline V operator/(const V& a, float s) { return EXPRS(/); }                               \
    inline V operator*(float s, const V& a) { return EXPRS_L(*); }                             \
    inline V& operator+=(V& a, const V& b) { a = a + b; return a; }
#define E2_2(op) vec2(a.x op b.x, a.y op b.y)
#define ES_2(op) vec2(a.x op s, a.y op s)
#define EL_2(op) vec2(s op a.x, s op a.y)
#define E2_3(op) vec3(a.x op b.x, a.y op b.y, a.z op b.z)
#define ES_3(op) vec3(a.x op s, a.y op s, a.z op s)
#define EL_3(op) vec3(s op a.x, s op a.y, s op a.z)
#define E2_4(op) vec4(a.x op b.x, a.y op b.y, a.z op b.z, a.w op b.w)
#define ES_4(op) vec4(a.x op s, a.y op s, a.z op s, a.w op s)
#define EL_4(op) vec4(s op a.x, s op a.y, s op a.z, s op a.w)
GLSL_OPS(vec2, E2_2, ES_2, EL_2)
GLSL_OPS(vec3, E2_3, ES_3, EL_3)
GLSL_OPS(vec4, E2_4, ES_4, EL_4)
#undef GLSL_OPS
inline vec3 operator-(const vec3& a) { return vec3(-a.x, -a.y, -a.z); }
inline bool operator!=(const vec4& a, const vec4& b) { return a.x != b.x || a.y != b.y || a.z != b.z || a.w != b.w; }
inline bool operator!=(const vec3& a, const vec3& b) { return a.x != b.x || a.y != b.y || a.z != b.z; }

// ---- built-in functions -----------------------------------------------------------------------------------
inline float sin(float x) { return std::sin(x); }
inline float cos(float x) { return std::cos(x); }
inline float acos(float x) { return std::acos(x); }
inline float atan(float y, float x) { return std::atan2(y, x); }
inline float pow(float x, float y) { return std::pow(x, y); }
inline float floor(float x) { return std::floor(x); }
inline float max(float a, float b) { return a < b ? b : a; }          // GLSL: max(x, y) = y if x < y else x
inline float min(float a, float b) { return b < a ? b : a; }
inline float clamp(float x, float lo, float hi) { return min(max(x, lo), hi); }
inline float mod(float x, float y) { return x - y * std::floor(x / y); }
inline float dot(const vec3& a, const vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float length(const vec3& v) { return std::sqrt(dot(v, v)); }
inline vec3 normalize(const vec3& v) { const float l = length(v); return vec3(v.x / l, v.y / l, v.z / l); }
inline vec3 reflect(const vec3& i, const vec3& n) { return i - 2.0f * dot(n, i) * n; }
inline vec4 pow(const vec4& a, const vec4& b) { return vec4(pow(a.x, b.x), pow(a.y, b.y), pow(a.z, b.z), pow(a.w, b.w)); }
inline bvec4 lessThan(const vec4& a, const vec4& b) { return bvec4{a.x < b.x, a.y < b.y, a.z < b.z, a.w < b.w}; }
inline vec3 pow(const vec3& a, const vec3& b) { return vec3(pow(a.x, b.x), pow(a.y, b.y), pow(a.z, b.z)); }
inline bvec3 lessThan(const vec3& a, const vec3& b) { return bvec3{a.x < b.x, a.y < b.y, a.z < b.z}; }
inline vec3 mix(const vec3& x, const vec3& y, const bvec3& a) { return vec3(a.x ? y.x : x.x, a.y ? y.y : x.y, a.z ? y.z : x.z); }
inline vec3 floor(const vec3& v) { return vec3(std::floor(v.x), std::floor(v.y), std::floor(v.z)); }
inline vec4 mix(const vec4& x, const vec4& y, const bvec4& a) { return vec4(a.x ? y.x : x.x, a.y ? y.y : x.y, a.z ? y.z : x.z, a.w ? y.w : x.w); }

// mat4x3: 4 columns of vec3 (gl_WorldToObjectEXT). v * M = row vector times matrix = (dot(v, column j))_j.
struct mat4x3 { vec3 c[4]; };
inline vec4 operator*(const vec3& v, const mat4x3& m) { return vec4(dot(v, m.c[0]), dot(v, m.c[1]), dot(v, m.c[2]), dot(v, m.c[3])); }

// ---- resources --------------------------------------------------------------------------------------------
struct sampler2D {          // RGBA32F texels, row-major; bilinear + repeat
    const float* texels = nullptr;
    int w = 0, h = 0;
};
inline int wrapi(int i, int n) { const int r = i % n; return r < 0 ? r + n : r; }
inline vec4 texelFetch(const sampler2D& s, const ivec2& p, int /*lod*/) {
    const float* t = s.texels + 4 * ((size_t)p.y * s.w + p.x);
    return vec4(t[0], t[1], t[2], t[3]);
}
inline vec4 texture(const sampler2D& s, const vec2& uv) {
    const float fx = uv.x * (float)s.w - 0.5f, fy = uv.y * (float)s.h - 0.5f;
    const float flx = std::floor(fx), fly = std::floor(fy);
    const float ax = fx - flx, ay = fy - fly;
    const int x0 = wrapi((int)flx, s.w), x1 = wrapi((int)flx + 1, s.w), y0 = wrapi((int)fly, s.h), y1 = wrapi((int)fly + 1, s.h);
    const vec4 p00 = texelFetch(s, ivec2(x0, y0), 0), p10 = texelFetch(s, ivec2(x1, y0), 0);
    const vec4 p01 = texelFetch(s, ivec2(x0, y1), 0), p11 = texelFetch(s, ivec2(x1, y1), 0);
    const vec4 top = p00 + (p10 - p00) * ax, bot = p01 + (p11 - p01) * ax;
    return top + (bot - top) * ay;
}
struct image2D {            // float RGBA storage; rgba8 = quantise on store as an RGBA8_UNORM image does
    float* texels = nullptr;
    int w = 0, h = 0;
    bool rgba8 = true;
};
inline vec4 imageLoad(const image2D& im, const ivec2& p) {
    const float* t = im.texels + 4 * ((size_t)p.y * im.w + p.x);
    return vec4(t[0], t[1], t[2], t[3]);
}
inline float to_unorm8(float c) { return std::nearbyint(clamp(c, 0.0f, 1.0f) * 255.0f) / 255.0f; }
inline void imageStore(image2D& im, const ivec2& p, const vec4& v) {
    float* t = im.texels + 4 * ((size_t)p.y * im.w + p.x);
    t[0] = im.rgba8 ? to_unorm8(v.x) : v.x; t[1] = im.rgba8 ? to_unorm8(v.y) : v.y;
    t[2] = im.rgba8 ? to_unorm8(v.z) : v.z; t[3] = im.rgba8 ? to_unorm8(v.w) : v.w;
}

// SH accumulator of the projection shaders: `vec3 coeffs[16]` with `+=` in DOUBLE, because the reference's
// fp32 non-atomic += is a data race whose intended value is the full sum (a sequential fp32 sum
// of 2 M texels is itself 5e-4 off).
struct dvec3acc {
    double x = 0, y = 0, z = 0;
    dvec3acc& operator+=(const vec3& v) { x += v.x; y += v.y; z += v.z; return *this; }
};

// ---- ray tracing ------------------------------------------------------------------------------------------
struct accelerationStructureEXT {};
const uint gl_RayFlagsOpaqueEXT = 1u, gl_RayFlagsTerminateOnFirstHitEXT = 4u, gl_RayFlagsSkipClosestHitShaderEXT = 8u;
// set by oracle/ref_pipeline.cpp: runs the intersection and the hit / miss shader the pipeline binds
typedef void (*trace_fn)(uint flags, uint miss_index, const float* origin, float tmin, const float* dir, float tmax, int payload);
extern trace_fn g_trace;
inline void traceRayEXT(const accelerationStructureEXT&, uint flags, uint /*mask*/, uint /*sbtOffset*/, uint /*sbtStride*/,
                        uint miss_index, const vec3& origin, float tmin, const vec3& dir, float tmax, int payload) {
    const float o[3] = {origin.x, origin.y, origin.z}, d[3] = {dir.x, dir.y, dir.z};
    g_trace(flags, miss_index, o, tmin, d, tmax, payload);
}

}  // namespace glsl

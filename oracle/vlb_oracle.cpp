// vlb_oracle.cpp — CPU ORACLE for the bake path of Reefufui/vulkan-light-bakery.
//
// THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE. Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / `--impl reference` legs may load it; nothing under
// vulkan-light-bakery_b200/ links, imports or calls it. It restates, in plain C++ on the CPU,
// what the reference's shaders compute (citations are paths relative to the reference root):
//
//   sh_basis / x2phi / y2theta / toVector ...... shaders/sh_common.h:1-224
//   vo_skybox_project ........................... shaders/skybox_sh.comp:23-42
//   vo_envmap_project ........................... shaders/sh.comp:23-42
//   probe ray generation ........................ shaders/env_map.rgen:18-28
//   hit shading + shadow ray .................... shaders/env_map.rchit:25-102
//   miss / sky lookup + sRGB .................... shaders/main.rmiss:9-41
//   probe grid .................................. src/baker/light_baker.cpp:80-101
//   output packing float[probe][16][3] .......... src/baker/light_baker.cpp:294-322
//   geometry flattening ......................... src/scene_manager.cpp:257-337,385-477
//
// PARITY PINNING. The reference ships no tests, golden vectors or KATs, and it
// cannot be built or run here (no Vulkan loader/ICD, glslang, glm, tinygltf). What CAN be pinned
// is pinned: oracle/_ref/libvlb_refsh.so is compiled from the reference's own
// shaders/sh_common.h (recipe: oracle/build_oracle.py) and tests/test_oracle_ref.py checks the
// SH basis, x2phi, y2theta and toVector of this file against it; tests/golden/ holds vectors
// generated from that library so the check also runs where /root/reference is absent.
// PARITY UNPINNED (driver-defined in the reference, no reference test constrains it): the
// ray/triangle intersection + acceleration structure (Vulkan KHR ray tracing), bilinear texture
// filtering, and the sin/cos/pow precision of the GLSL implementation. For those this file IS
// the specification the CUDA path is held to: Moller-Trumbore with the exact operation order
// of `intersect_tri` below (explicit fmaf, no contraction), closest hit = smallest t then
// smallest flat triangle id, and trigonometry evaluated in double and rounded to float.
// Accumulation is in double (sequential fp32 accumulation is off by 5e-4).
//
// Build: g++ -O3 -march=native -ffp-contract=off -fopenmp -shared -fPIC (oracle/build_oracle.py).

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/vlb_bake.h"

namespace {

constexpr float PI_F = 3.1415926538f;  // sh_common.h:1 (rounds to 3.14159274f)

struct V3 { float x, y, z; };
static inline V3 v3(float x, float y, float z) { return V3{x, y, z}; }
static inline V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
static inline float dot3(V3 a, V3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }

// Trigonometry: evaluated in double, rounded once to float (see header: GLSL precision unpinned).
static inline float sin_f(float x) { return (float)std::sin((double)x); }
static inline float cos_f(float x) { return (float)std::cos((double)x); }

// sh_common.h:14-22
static inline float x2phi(int x, int w) { return 2.f * PI_F * ((float)x + 0.5f) / (float)w; }
static inline float y2theta(int y, int h) { return PI_F * ((float)y + 0.5f) / (float)h; }

// normalize(): v / sqrt(dot(v,v)) with the fixed operation order shared with the CUDA path.
static inline V3 normalize3(V3 v) {
    float l2 = fmaf(v.z, v.z, fmaf(v.y, v.y, v.x * v.x));
    float len = sqrtf(l2);
    return v3(v.x / len, v.y / len, v.z / len);
}

// sh_common.h:8-12 given sin/cos of theta and phi.
static inline V3 to_vector_sc(float st, float ct, float cp, float sp) {
    return normalize3(v3(st * cp, st * sp, ct));
}
static inline V3 to_vector(float phi, float theta) {
    return to_vector_sc(sin_f(theta), cos_f(theta), cos_f(phi), sin_f(phi));
}

// sh_common.h:26-152, index i = l(l+1)+m. 6-digit constants verbatim, fp32 like GLSL.
static inline void sh_basis25(V3 d, float* o) {
    const float x = d.x, y = d.y, z = d.z;
    o[0] = 0.282095f;
    o[1] = -0.488603f * y;
    o[2] = 0.488603f * z;
    o[3] = -0.488603f * x;
    o[4] = 1.092548f * x * y;
    o[5] = -1.092548f * y * z;
    o[6] = 0.315392f * (-x * x - y * y + 2.0f * z * z);
    o[7] = -1.092548f * x * z;
    o[8] = 0.546274f * (x * x - y * y);
    o[9] = -0.590044f * y * (3.0f * x * x - y * y);
    o[10] = 2.890611f * x * y * z;
    o[11] = -0.457046f * y * (4.0f * z * z - x * x - y * y);
    o[12] = 0.373176f * z * (2.0f * z * z - 3.0f * x * x - 3.0f * y * y);
    o[13] = -0.457046f * x * (4.0f * z * z - x * x - y * y);
    o[14] = 1.445306f * z * (x * x - y * y);
    o[15] = -0.590044f * x * (x * x - 3.0f * y * y);
    o[16] = 2.503343f * x * y * (x * x - y * y);
    o[17] = -1.770131f * y * z * (3.0f * x * x - y * y);
    o[18] = 0.946175f * x * y * (7.0f * z * z - 1.0f);
    o[19] = -0.669047f * y * z * (7.0f * z * z - 3.0f);
    {
        float z2 = z * z;
        o[20] = 0.105786f * (35.0f * z2 * z2 - 30.0f * z2 + 3.0f);
    }
    o[21] = -0.669047f * x * z * (7.0f * z * z - 3.0f);
    o[22] = 0.473087f * (x * x - y * y) * (7.0f * z * z - 1.0f);
    o[23] = -1.770131f * x * z * (x * x - 3.0f * y * y);
    {
        float x2 = x * x, y2 = y * y;
        o[24] = 0.625836f * (x2 * (x2 - 3.0f * y2) - y2 * (3.0f * x2 - y2));
    }
}

static inline int n_coeffs(int order) { return (order + 1) * (order + 1); }

// env_map.rchit:27-34 / main.rmiss:9-16, one channel.
static inline float srgb1(float c) {
    return c < 0.0031308f ? c * 12.92f : 1.055f * (float)std::pow((double)c, 1.0 / 2.4) - 0.055f;
}

static inline void fetch_texel(const void* texels, int fmt, int w, int x, int y, float rgb[3]) {
    if (fmt == VLB_FMT_RGBA32F) {
        const float* p = (const float*)texels + ((size_t)y * w + x) * 4;
        rgb[0] = p[0]; rgb[1] = p[1]; rgb[2] = p[2];
    } else {
        const uint8_t* p = (const uint8_t*)texels + ((size_t)y * w + x) * 4;
        rgb[0] = (float)p[0] / 255.0f; rgb[1] = (float)p[1] / 255.0f; rgb[2] = (float)p[2] / 255.0f;
    }
}

// Common body of skybox_sh.comp:25-41 (skybox=true) and sh.comp:25-41 (skybox=false).
static void project_image(const void* texels, int fmt, int W, int H, int order, bool skybox,
                          float* out48) {
    const int K = n_coeffs(order);
    std::vector<double> acc(48, 0.0);
    const float pixelArea = (2.0f * PI_F / (float)W) * (PI_F / (float)H);
#pragma omp parallel
    {
        std::vector<double> loc(48, 0.0);
#pragma omp for schedule(static)
        for (int y = 0; y < H; ++y) {
            const float theta = y2theta(y, H);
            const float st = sin_f(theta), ct = cos_f(theta);
            const float weight = pixelArea * st;
            for (int x = 0; x < W; ++x) {
                float phi = x2phi(x, W);
                if (skybox) phi = phi - PI_F / 2.0f;                 // skybox_sh.comp:28
                V3 d = to_vector_sc(st, ct, cos_f(phi), sin_f(phi));
                V3 s = skybox ? v3(d.x, d.z, d.y) : d;               // skybox_sh.comp:39 dir.xzy
                float b[25];
                sh_basis25(s, b);
                float rgb[3];
                fetch_texel(texels, fmt, W, x, y, rgb);
                for (int i = 0; i < K; ++i) {
                    const double bw = (double)b[i] * (double)weight;
                    loc[i * 3 + 0] += bw * rgb[0];
                    loc[i * 3 + 1] += bw * rgb[1];
                    loc[i * 3 + 2] += bw * rgb[2];
                }
            }
        }
#pragma omp critical
        for (int i = 0; i < 48; ++i) acc[i] += loc[i];
    }
    for (int i = 0; i < 48; ++i) out48[i] = (float)acc[i];
}

// ------------------------------------------------------------------------------------------
// Scene: flattened world-space triangles + a simple CPU BVH.
// ------------------------------------------------------------------------------------------
struct Tri {
    V3 v0, e1, e2;     // world space
    V3 n0, n1, n2;     // OBJECT-space vertex normals (env_map.rchit:64 interpolates these)
    float uv[6];       // Vertex::uv0 of the three corners (env_map.rchit:65)
    uint32_t inst;
};

struct Inst {
    float m[12];       // object->world 3x4 row-major
    float nm[9];       // inverse(M3x3) rows, so that (nrm * W2O) = nm^T * nrm; see xform_normal
    uint32_t material;
};

struct BNode {
    float lo[3], hi[3];
    int32_t left, right;   // children (internal) ; left = -1 -> leaf
    int32_t first, count;  // leaf range into order[]
};

struct Scene {
    std::vector<Tri> tris;
    std::vector<Inst> insts;
    std::vector<vlb_material> mats;
    std::vector<BNode> nodes;
    std::vector<uint32_t> order;
    float ref_bounds[6];
    float tight_bounds[6];
    // sky
    std::vector<float> sky;  // RGBA32F copy (RGBA8 is converted value/255)
    int skyW = 0, skyH = 0;
    // textures (Scene_t::loadTextures, src/scene_manager.cpp:941-973): RGBA8 level 0 + sampler state
    struct Tex { std::vector<uint8_t> texels; int W = 0, H = 0, wrap_u = 0, wrap_v = 0, filter = 0; };
    std::vector<Tex> textures;
};

static inline V3 xform_point(const float* m, V3 p) {
    V3 r;
    r.x = fmaf(m[2], p.z, fmaf(m[1], p.y, fmaf(m[0], p.x, m[3])));
    r.y = fmaf(m[6], p.z, fmaf(m[5], p.y, fmaf(m[4], p.x, m[7])));
    r.z = fmaf(m[10], p.z, fmaf(m[9], p.y, fmaf(m[8], p.x, m[11])));
    return r;
}

// inverse of the upper 3x3 (double, rounded to float). nm[r*3+c] = (M^-1)[r][c].
static void inverse3x3(const float* m, float* nm) {
    const double a = m[0], b = m[1], c = m[2], d = m[4], e = m[5], f = m[6], g = m[8], h = m[9],
                 i = m[10];
    const double A = e * i - f * h, B = -(d * i - f * g), C = d * h - e * g;
    const double det = a * A + b * B + c * C;
    const double id = det != 0.0 ? 1.0 / det : 0.0;
    nm[0] = (float)(A * id);               nm[1] = (float)(-(b * i - c * h) * id); nm[2] = (float)((b * f - c * e) * id);
    nm[3] = (float)(B * id);               nm[4] = (float)((a * i - c * g) * id);  nm[5] = (float)(-(a * f - c * d) * id);
    nm[6] = (float)(C * id);               nm[7] = (float)(-(a * h - b * g) * id); nm[8] = (float)((a * e - b * d) * id);
}

// vec3(nrm * gl_WorldToObjectEXT) (env_map.rchit:68): component j = dot(nrm, column j of W2O)
// = sum_r nrm[r] * Minv[r][j].
static inline V3 xform_normal(const float* nm, V3 n) {
    V3 r;
    r.x = fmaf(nm[6], n.z, fmaf(nm[3], n.y, nm[0] * n.x));
    r.y = fmaf(nm[7], n.z, fmaf(nm[4], n.y, nm[1] * n.x));
    r.z = fmaf(nm[8], n.z, fmaf(nm[5], n.y, nm[2] * n.x));
    return r;
}

// THE intersection specification (see header). Returns true and (t,u,v) if the ray hits.
static inline bool intersect_tri(const Tri& tr, V3 o, V3 d, float& t, float& u, float& v) {
    const V3 e1 = tr.e1, e2 = tr.e2;
    const float px = fmaf(d.y, e2.z, -(d.z * e2.y));
    const float py = fmaf(d.z, e2.x, -(d.x * e2.z));
    const float pz = fmaf(d.x, e2.y, -(d.y * e2.x));
    const float det = fmaf(e1.z, pz, fmaf(e1.y, py, e1.x * px));
    if (det == 0.0f) return false;
    const float inv = 1.0f / det;
    const float tx = o.x - tr.v0.x, ty = o.y - tr.v0.y, tz = o.z - tr.v0.z;
    u = fmaf(tz, pz, fmaf(ty, py, tx * px)) * inv;
    if (!(u >= 0.0f) || u > 1.0f) return false;
    const float qx = fmaf(ty, e1.z, -(tz * e1.y));
    const float qy = fmaf(tz, e1.x, -(tx * e1.z));
    const float qz = fmaf(tx, e1.y, -(ty * e1.x));
    v = fmaf(d.z, qz, fmaf(d.y, qy, d.x * qx)) * inv;
    if (!(v >= 0.0f) || u + v > 1.0f) return false;
    t = fmaf(e2.z, qz, fmaf(e2.y, qy, e2.x * qx)) * inv;
    return true;
}

struct Hit { int32_t id; float t, u, v; };

// --- CPU BVH (binned SAH, top-down). Boxes are padded so the float slab test is conservative.
static void tri_bounds(const Tri& t, float lo[3], float hi[3]) {
    const V3 a = t.v0, b = t.v0 + t.e1, c = t.v0 + t.e2;
    lo[0] = std::min(a.x, std::min(b.x, c.x)); hi[0] = std::max(a.x, std::max(b.x, c.x));
    lo[1] = std::min(a.y, std::min(b.y, c.y)); hi[1] = std::max(a.y, std::max(b.y, c.y));
    lo[2] = std::min(a.z, std::min(b.z, c.z)); hi[2] = std::max(a.z, std::max(b.z, c.z));
}

struct Builder {
    Scene& s;
    std::vector<float> tlo, thi, cen;
    float pad;
    explicit Builder(Scene& sc) : s(sc) {}

    float area(const float* lo, const float* hi) {
        float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        return 2.f * (dx * dy + dy * dz + dz * dx);
    }

    int build(int first, int count) {
        BNode n;
        for (int k = 0; k < 3; ++k) { n.lo[k] = 1e30f; n.hi[k] = -1e30f; }
        float clo[3] = {1e30f, 1e30f, 1e30f}, chi[3] = {-1e30f, -1e30f, -1e30f};
        for (int i = first; i < first + count; ++i) {
            uint32_t t = s.order[i];
            for (int k = 0; k < 3; ++k) {
                n.lo[k] = std::min(n.lo[k], tlo[t * 3 + k]);
                n.hi[k] = std::max(n.hi[k], thi[t * 3 + k]);
                clo[k] = std::min(clo[k], cen[t * 3 + k]);
                chi[k] = std::max(chi[k], cen[t * 3 + k]);
            }
        }
        n.left = -1; n.right = -1; n.first = first; n.count = count;
        const int idx = (int)s.nodes.size();
        s.nodes.push_back(n);
        if (count <= 4) { pad_node(idx); return idx; }
        // binned SAH over the widest centroid axis
        int axis = 0;
        float ext = chi[0] - clo[0];
        for (int k = 1; k < 3; ++k) if (chi[k] - clo[k] > ext) { ext = chi[k] - clo[k]; axis = k; }
        int mid = first + count / 2;
        if (ext > 0.f) {
            const int NB = 16;
            int cnt[NB] = {0};
            float blo[NB][3], bhi[NB][3];
            for (int b = 0; b < NB; ++b) for (int k = 0; k < 3; ++k) { blo[b][k] = 1e30f; bhi[b][k] = -1e30f; }
            const float scale = (float)NB / ext;
            auto bin_of = [&](uint32_t t) {
                int b = (int)((cen[t * 3 + axis] - clo[axis]) * scale);
                return std::min(std::max(b, 0), NB - 1);
            };
            for (int i = first; i < first + count; ++i) {
                uint32_t t = s.order[i];
                int b = bin_of(t);
                cnt[b]++;
                for (int k = 0; k < 3; ++k) {
                    blo[b][k] = std::min(blo[b][k], tlo[t * 3 + k]);
                    bhi[b][k] = std::max(bhi[b][k], thi[t * 3 + k]);
                }
            }
            float la[NB], ra[NB]; int lc[NB], rc[NB];
            float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
            int c = 0;
            for (int b = 0; b < NB; ++b) {
                c += cnt[b];
                for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], blo[b][k]); hi[k] = std::max(hi[k], bhi[b][k]); }
                la[b] = c ? area(lo, hi) : 0.f; lc[b] = c;
            }
            for (int k = 0; k < 3; ++k) { lo[k] = 1e30f; hi[k] = -1e30f; }
            c = 0;
            for (int b = NB - 1; b >= 0; --b) {
                c += cnt[b];
                for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], blo[b][k]); hi[k] = std::max(hi[k], bhi[b][k]); }
                ra[b] = c ? area(lo, hi) : 0.f; rc[b] = c;
            }
            float best = 1e30f; int bs = -1;
            for (int b = 0; b < NB - 1; ++b) {
                if (lc[b] == 0 || rc[b + 1] == 0) continue;
                float cost = la[b] * lc[b] + ra[b + 1] * rc[b + 1];
                if (cost < best) { best = cost; bs = b; }
            }
            if (bs >= 0) {
                auto it = std::partition(s.order.begin() + first, s.order.begin() + first + count,
                                         [&](uint32_t t) { return bin_of(t) <= bs; });
                mid = (int)(it - s.order.begin());
            }
        }
        if (mid == first || mid == first + count) {
            mid = first + count / 2;
            std::nth_element(s.order.begin() + first, s.order.begin() + mid,
                             s.order.begin() + first + count, [&](uint32_t a, uint32_t b) {
                                 return cen[a * 3 + axis] < cen[b * 3 + axis];
                             });
        }
        int l = build(first, mid - first);
        int r = build(mid, first + count - mid);
        s.nodes[idx].left = l; s.nodes[idx].right = r;
        pad_node(idx);
        return idx;
    }
    void pad_node(int idx) {
        BNode& n = s.nodes[idx];
        for (int k = 0; k < 3; ++k) {
            float m = std::max(std::fabs(n.lo[k]), std::fabs(n.hi[k]));
            float p = pad + m * 4e-6f;
            n.lo[k] -= p; n.hi[k] += p;
        }
    }
    void run() {
        const size_t n = s.tris.size();
        tlo.resize(n * 3); thi.resize(n * 3); cen.resize(n * 3);
        float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
        for (size_t i = 0; i < n; ++i) {
            tri_bounds(s.tris[i], &tlo[i * 3], &thi[i * 3]);
            for (int k = 0; k < 3; ++k) {
                cen[i * 3 + k] = 0.5f * (tlo[i * 3 + k] + thi[i * 3 + k]);
                lo[k] = std::min(lo[k], tlo[i * 3 + k]); hi[k] = std::max(hi[k], thi[i * 3 + k]);
            }
        }
        for (int k = 0; k < 3; ++k) { s.tight_bounds[k] = n ? lo[k] : 0.f; s.tight_bounds[3 + k] = n ? hi[k] : 0.f; }
        float ext = 0.f;
        for (int k = 0; k < 3; ++k) ext = std::max(ext, hi[k] - lo[k]);
        pad = n ? ext * 1e-6f : 0.f;
        s.order.resize(n);
        for (size_t i = 0; i < n; ++i) s.order[i] = (uint32_t)i;
        s.nodes.clear();
        s.nodes.reserve(n ? n : 1);
        if (n) build(0, (int)n);
    }
};

static inline bool box_hit(const BNode& n, V3 o, V3 id, float tmin, float tmax) {
    float t0 = (n.lo[0] - o.x) * id.x, t1 = (n.hi[0] - o.x) * id.x;
    float tn = std::min(t0, t1), tf = std::max(t0, t1);
    t0 = (n.lo[1] - o.y) * id.y; t1 = (n.hi[1] - o.y) * id.y;
    tn = std::max(tn, std::min(t0, t1)); tf = std::min(tf, std::max(t0, t1));
    t0 = (n.lo[2] - o.z) * id.z; t1 = (n.hi[2] - o.z) * id.z;
    tn = std::max(tn, std::min(t0, t1)); tf = std::min(tf, std::max(t0, t1));
    tn = std::max(tn, tmin); tf = std::min(tf, tmax);
    return tn <= tf;   // boxes are padded (Builder::pad_node), so the plain test is conservative
}

static inline float safe_inv(float d) {
    const float eps = 1e-30f;
    if (std::fabs(d) < eps) d = std::copysign(eps, d);
    return 1.0f / d;
}

static inline void consider(const Scene& s, uint32_t tid, V3 o, V3 d, float tmin, Hit& best) {
    float t, u, v;
    if (!intersect_tri(s.tris[tid], o, d, t, u, v)) return;
    if (!(t > tmin)) return;
    if (t < best.t || (t == best.t && best.id >= 0 && (int32_t)tid < best.id)) {
        best.id = (int32_t)tid; best.t = t; best.u = u; best.v = v;
    }
}

// closest hit with tmin < t < tmax. best.t starts at tmax so only t < tmax is accepted.
static Hit trace_closest(const Scene& s, V3 o, V3 d, float tmin, float tmax, bool brute) {
    Hit best{-1, tmax, 0.f, 0.f};
    if (s.tris.empty()) return best;
    if (brute) {
        for (uint32_t i = 0; i < s.tris.size(); ++i) consider(s, i, o, d, tmin, best);
        return best;
    }
    const V3 id = v3(safe_inv(d.x), safe_inv(d.y), safe_inv(d.z));
    int stack[128]; int sp = 0; stack[sp++] = 0;
    while (sp) {
        const BNode& n = s.nodes[stack[--sp]];
        if (!box_hit(n, o, id, tmin, best.t)) continue;
        if (n.left < 0) {
            for (int i = n.first; i < n.first + n.count; ++i) consider(s, s.order[i], o, d, tmin, best);
        } else {
            stack[sp++] = n.left; stack[sp++] = n.right;
        }
    }
    return best;
}

// any hit with tmin < t < tmax (shadow rays: gl_RayFlagsTerminateOnFirstHitEXT, env_map.rchit:86)
static bool trace_any(const Scene& s, V3 o, V3 d, float tmin, float tmax, bool brute, Hit* out) {
    if (s.tris.empty()) return false;
    float t, u, v;
    if (brute) {
        for (uint32_t i = 0; i < s.tris.size(); ++i)
            if (intersect_tri(s.tris[i], o, d, t, u, v) && t > tmin && t < tmax) {
                if (out) *out = Hit{(int32_t)i, t, u, v};
                return true;
            }
        return false;
    }
    const V3 id = v3(safe_inv(d.x), safe_inv(d.y), safe_inv(d.z));
    int stack[128]; int sp = 0; stack[sp++] = 0;
    while (sp) {
        const BNode& n = s.nodes[stack[--sp]];
        if (!box_hit(n, o, id, tmin, tmax)) continue;
        if (n.left < 0) {
            for (int i = n.first; i < n.first + n.count; ++i) {
                uint32_t tid = s.order[i];
                if (intersect_tri(s.tris[tid], o, d, t, u, v) && t > tmin && t < tmax) {
                    if (out) *out = Hit{(int32_t)tid, t, u, v};
                    return true;
                }
            }
        } else {
            stack[sp++] = n.left; stack[sp++] = n.right;
        }
    }
    return false;
}

// main.rmiss:18-35 + bilinear/repeat lookup (VkSampler of src/application.hpp:45-52; LOD 0).
static inline float glsl_mod(float x, float y) { return x - y * std::floor(x / y); }
static inline float clampf(float x, float a, float b) { return std::min(std::max(x, a), b); }

// dir2SkyboxUV, main.rmiss:18-35
static void dir_to_sky_uv(V3 dir, float* u_out, float* v_out) {
    float theta = (float)std::acos((double)clampf(dir.y, -1.0f, 1.0f));
    float phi = (float)std::atan2((double)dir.x, (double)dir.z);
    theta = glsl_mod(theta, 2.0f * PI_F);
    theta = clampf(theta, 0.0f, 2.0f * PI_F);
    if (theta > PI_F) { theta = 2.0f * PI_F - theta; phi += PI_F; }
    phi = glsl_mod(phi, 2.0f * PI_F);
    phi = clampf(phi, 0.0f, 2.0f * PI_F);
    *u_out = phi / (2.0f * PI_F); *v_out = theta / PI_F;
}

static void sky_lookup(const Scene& s, V3 dir, float rgb[3]) {
    float u, v;
    dir_to_sky_uv(dir, &u, &v);
    const int W = s.skyW, H = s.skyH;
    const float fx = u * (float)W - 0.5f, fy = v * (float)H - 0.5f;
    const float flx = std::floor(fx), fly = std::floor(fy);
    const float ax = fx - flx, ay = fy - fly;
    auto wrap = [](int i, int n) { int r = i % n; return r < 0 ? r + n : r; };
    const int x0 = wrap((int)flx, W), x1 = wrap((int)flx + 1, W);
    const int y0 = wrap((int)fly, H), y1 = wrap((int)fly + 1, H);
    const float* p00 = &s.sky[((size_t)y0 * W + x0) * 4];
    const float* p10 = &s.sky[((size_t)y0 * W + x1) * 4];
    const float* p01 = &s.sky[((size_t)y1 * W + x0) * 4];
    const float* p11 = &s.sky[((size_t)y1 * W + x1) * 4];
    for (int c = 0; c < 3; ++c) {
        const float top = p00[c] + (p10[c] - p00[c]) * ax;
        const float bot = p01[c] + (p11[c] - p01[c]) * ax;
        rgb[c] = top + (bot - top) * ay;
    }
}

// VkSamplerAddressMode of one texel index (loadSamplers, src/scene_manager.cpp:654-668).
static inline int wrap_texel(int i, int n, int mode) {
    auto wrap = [](int i, int n) { int r = i % n; return r < 0 ? r + n : r; };
    if (mode == VLB_WRAP_CLAMP_TO_EDGE) return i < 0 ? 0 : (i >= n ? n - 1 : i);
    if (mode == VLB_WRAP_MIRRORED_REPEAT) { const int m = wrap(i, 2 * n); return m < n ? m : 2 * n - 1 - m; }
    return wrap(i, n);
}

// texture(textures[idx], uv).rgb (env_map.rchit:42) at the base level ("parity unpinned": the driver's
// filter arithmetic is not specified bit for bit; this restatement is): unnormalised coordinate
// u * W - 0.5, four neighbours under the address modes, fp32 lerps in the order of sky_lookup.
static void tex_sample(const Scene::Tex& t, float u, float v, float rgb[3]) {
    auto texel = [&](int x, int y, float c[3]) {
        const uint8_t* p = &t.texels[((size_t)y * t.W + x) * 4];
        c[0] = (float)p[0] / 255.0f; c[1] = (float)p[1] / 255.0f; c[2] = (float)p[2] / 255.0f;
    };
    if (t.filter == VLB_FILTER_NEAREST) {
        texel(wrap_texel((int)std::floor(u * (float)t.W), t.W, t.wrap_u), wrap_texel((int)std::floor(v * (float)t.H), t.H, t.wrap_v), rgb);
        return;
    }
    const float fx = u * (float)t.W - 0.5f, fy = v * (float)t.H - 0.5f;
    const float flx = std::floor(fx), fly = std::floor(fy);
    const float ax = fx - flx, ay = fy - fly;
    const int x0 = wrap_texel((int)flx, t.W, t.wrap_u), x1 = wrap_texel((int)flx + 1, t.W, t.wrap_u);
    const int y0 = wrap_texel((int)fly, t.H, t.wrap_v), y1 = wrap_texel((int)fly + 1, t.H, t.wrap_v);
    float p00[3], p10[3], p01[3], p11[3];
    texel(x0, y0, p00); texel(x1, y0, p10); texel(x0, y1, p01); texel(x1, y1, p11);
    for (int c = 0; c < 3; ++c) {
        const float top = p00[c] + (p10[c] - p00[c]) * ax;
        const float bot = p01[c] + (p11[c] - p01[c]) * ax;
        rgb[c] = top + (bot - top) * ay;
    }
}

// getBaseColor, env_map.rchit:36-49: texture if the material names one, else the factor if non-zero, else 1.
static void base_color(const Scene& s, const vlb_material& m, float u, float v, float out[4]) {
    const float* f = m.base_color_factor;
    const int ti = m.base_color.index;
    if (ti >= 0 && ti < (int)s.textures.size()) {
        tex_sample(s.textures[ti], u, v, out);
        out[3] = 1.0f;
    } else if (f[0] != 0.f || f[1] != 0.f || f[2] != 0.f || f[3] != 0.f) {
        out[0] = f[0]; out[1] = f[1]; out[2] = f[2]; out[3] = f[3];
    } else {
        out[0] = out[1] = out[2] = out[3] = 1.0f;
    }
}

// ---- multi-bounce gather: the reference's run-time operator, shaders/main.rchit:124-163 with the
// probe lookup of shaders/sh.rmiss:20-36, restated for an arbitrary grid. The reference never runs this
// inside the bake (it bakes direct light only), so as a bake pass this restatement is the specification of
// BASELINE configs[3]; as an operator it is pinned to the reference's own main.rchit + sh.rmiss, compiled from
// /root/reference/shaders (tests/test_ref_shaders.py: *_viewer_hit_shader, under the conditions in which the
// deviations below vanish). Deviations from the literal shader, all documented in
// include/vlb_bake.h: the grid origin is honoured (main.rchit:126 assumes 0); the cell is clamped into
// the grid (the shader hard-codes 7x7x7, sh.rmiss:22); each corner reads ITS probe (sh.rmiss:25 reads
// the cell's base probe for all eight -- evident defect); weights are clamped at 0 and an empty
// weight sum yields 0 (the shader divides 0/0); the SH argument is the normal in the frame the bake
// stored the coefficients in.
struct GatherCtx {
    const float* prev;                 // [Nx*Ny*Nz][48], x-fastest
    const float* px; const float* py; const float* pz;
    int Nx, Ny, Nz;
};

static inline int cell_of(float p, float origin, float step, int n) {
    if (n < 2) return 0;
    const float g = std::floor((p - origin) / step);                 // main.rchit:126
    if (!(g > 0.0f)) return 0;                                       // also NaN (step == 0)
    return g >= (float)(n - 2) ? n - 2 : (int)g;
}

static void gather_indirect(const Scene& s, const vlb_bake_settings& st, const GatherCtx& g, V3 P, V3 N,
                            V3 so, bool brute, float out[3]) {
    const int K = n_coeffs(st.sh_order);
    const int ci = cell_of(P.x, st.origin[0], st.step[0], g.Nx);
    const int cj = cell_of(P.y, st.origin[1], st.step[1], g.Ny);
    const int ck = cell_of(P.z, st.origin[2], st.step[2], g.Nz);
    const float weightMax = sqrtf(dot3(v3(st.step[0], st.step[1], st.step[2]), v3(st.step[0], st.step[1], st.step[2])));  // :141
    float b[25];
    sh_basis25((st.flags & VLB_BAKE_SH_WORLD_FRAME) ? N : v3(N.x, N.z, N.y), b);
    float sum[3] = {0.f, 0.f, 0.f}, wsum = 0.f;
    for (int c = 0; c < 8; ++c) {                                    // gridVertices order, :128-137
        const int i = std::min(ci + ((c >> 2) & 1), g.Nx - 1), j = std::min(cj + ((c >> 1) & 1), g.Ny - 1),
                  k = std::min(ck + (c & 1), g.Nz - 1);
        const V3 d = v3(g.px[i] - P.x, g.py[j] - P.y, g.pz[k] - P.z);    // :145
        const float tmax = sqrtf(dot3(d, d));                            // :154
        const float w = std::max(weightMax - tmax, 0.0f);                // :156
        bool occluded = false;
        if (tmax > 0.0f) occluded = trace_any(s, so, v3(d.x / tmax, d.y / tmax, d.z / tmax), 0.0f, tmax, brute, nullptr);  // :155
        if (occluded) continue;
        const float* sh = g.prev + ((size_t)i + (size_t)g.Nx * ((size_t)j + (size_t)g.Ny * k)) * 48;   // sh.rmiss:25
        float v[3] = {0.f, 0.f, 0.f};
        for (int q = 0; q < K; ++q) {                                    // sh.rmiss:27-34
            v[0] = fmaf(sh[3 * q + 0], b[q], v[0]); v[1] = fmaf(sh[3 * q + 1], b[q], v[1]); v[2] = fmaf(sh[3 * q + 2], b[q], v[2]);
        }
        sum[0] = fmaf(w, v[0], sum[0]); sum[1] = fmaf(w, v[1], sum[1]); sum[2] = fmaf(w, v[2], sum[2]);   // :160
        wsum += w;                                                        // :161
    }
    for (int c = 0; c < 3; ++c) out[c] = wsum > 0.0f ? st.indirect_gain * (sum[c] / wsum) : 0.0f;        // :164-165
}

// env_map.rchit:51-102
static void shade_hit(const Scene& s, const vlb_bake_settings& st, const Hit& h, V3 o, V3 r,
                      bool brute, float rgb[3], uint64_t* shadow_rays, const GatherCtx* g = nullptr) {
    const Tri& tr = s.tris[h.id];
    const Inst& in = s.insts[tr.inst];
    const float b0 = 1.0f - h.u - h.v, b1 = h.u, b2 = h.v;
    const V3 nrm = tr.n0 * b0 + tr.n1 * b1 + tr.n2 * b2;
    const V3 P = v3(fmaf(r.x, h.t, o.x), fmaf(r.y, h.t, o.y), fmaf(r.z, h.t, o.z));
    const V3 N = normalize3(xform_normal(in.nm, nrm));
    float bc[4];
    const float tu = (tr.uv[0] * b0 + tr.uv[2] * b1) + tr.uv[4] * b2;       // env_map.rchit:65
    const float tv = (tr.uv[1] * b0 + tr.uv[3] * b1) + tr.uv[5] * b2;
    base_color(s, s.mats[in.material], tu, tv, bc);
    const V3 L = v3(st.light_pos[0] - P.x, st.light_pos[1] - P.y, st.light_pos[2] - P.z);
    const float llen = sqrtf(dot3(L, L));
    const V3 Ln = v3(L.x / llen, L.y / llen, L.z / llen);
    const float sDotN = std::max(dot3(Ln, N), 0.0f);
    float diffuse = 0.f, specular = 0.f;
    bool inShadow = true;
    if (sDotN != 0.0f) {
        if (st.flags & VLB_BAKE_SHADOW_RAYS) {
            const V3 so = v3(fmaf(st.shadow_bias, N.x, P.x), fmaf(st.shadow_bias, N.y, P.y),
                             fmaf(st.shadow_bias, N.z, P.z));
            if (shadow_rays) ++*shadow_rays;
            inShadow = trace_any(s, so, Ln, 0.0f, llen, brute, nullptr);
        } else {
            inShadow = false;
        }
    }
    if (!inShadow) {
        diffuse = st.c_diffuse * sDotN;
        const float dn = dot3(N, Ln);                                  // reflect(I,N) = I - 2 dot(N,I) N
        const V3 R = v3(Ln.x - 2.0f * dn * N.x, Ln.y - 2.0f * dn * N.y, Ln.z - 2.0f * dn * N.z);
        const float rd = std::max(dot3(R, r), 0.0f);
        specular = st.c_specular * (float)std::pow((double)rd, (double)st.gloss);
    }
    const float k = st.ambient + diffuse + specular;
    float ind[3] = {0.f, 0.f, 0.f};
    if (g) {                                                           // main.rchit:102,124-165
        const V3 so = v3(fmaf(st.shadow_bias, N.x, P.x), fmaf(st.shadow_bias, N.y, P.y), fmaf(st.shadow_bias, N.z, P.z));
        gather_indirect(s, st, *g, P, N, so, brute, ind);
    }
    for (int c = 0; c < 3; ++c) {
        float v = bc[c] * (k + ind[c]);
        rgb[c] = (st.flags & VLB_BAKE_SRGB_ENCODE) ? srgb1(v) : v;
    }
}

static inline float quant8(float c) {
    float x = clampf(c, 0.0f, 1.0f);
    return std::nearbyint(x * 255.0f) / 255.0f;   // round-half-even, as UNORM conversion
}

struct DirTable {
    int W, H;
    std::vector<V3> t;       // un-swizzled toVector per texel (sh.comp:30)
    std::vector<float> w;    // weight per row (sh.comp:32-33)
};
static DirTable make_dirs(int W, int H) {
    DirTable d; d.W = W; d.H = H; d.t.resize((size_t)W * H); d.w.resize(H);
    const float pixelArea = (2.0f * PI_F / (float)W) * (PI_F / (float)H);
    std::vector<float> cp(W), sp(W);
    for (int x = 0; x < W; ++x) { float p = x2phi(x, W); cp[x] = cos_f(p); sp[x] = sin_f(p); }
    for (int y = 0; y < H; ++y) {
        const float th = y2theta(y, H); const float st = sin_f(th), ct = cos_f(th);
        d.w[y] = pixelArea * st;
        for (int x = 0; x < W; ++x) d.t[(size_t)y * W + x] = to_vector_sc(st, ct, cp[x], sp[x]);
    }
    return d;
}

// LightBaker::probePositionsFromBoudingBox coordinates are separable per axis: coordinate i of
// an axis is bounds_min + step added i times in fp32 (light_baker.cpp:92-96).
static void axis_coords(float origin, float step, int n, std::vector<float>& out) {
    out.resize(n);
    float p = origin;
    for (int i = 0; i < n; ++i) { out[i] = p; p += step; }
}

// writer order of light_baker.cpp:80-101: output slot of grid cell (i,j,k).
static inline size_t ref_order_index(int i, int j, int k, int Nx, int Ny, int Nz) {
    size_t q = (j == 0) ? (size_t)i : (size_t)Nx + (size_t)i * (Ny - 1) + (j - 1);
    return (k == 0) ? q : (size_t)Nx * Ny + q * (Nz - 1) + (k - 1);
}

static void bake_one(const Scene& s, const vlb_bake_settings& st, const DirTable& dt, V3 pos,
                     bool brute, double* acc48, float* image_rgb, uint64_t* shadow_rays, const GatherCtx* g = nullptr) {
    const int K = n_coeffs(st.sh_order);
    for (int y = 0; y < dt.H; ++y) {
        for (int x = 0; x < dt.W; ++x) {
            const V3 t = dt.t[(size_t)y * dt.W + x];
            const V3 r = v3(t.x, t.z, t.y);                            // env_map.rgen:21 .xzy
            float rgb[3] = {0.f, 0.f, 0.f};                            // env_map.rgen:25
            const Hit h = trace_closest(s, pos, r, st.tmin, st.tmax, brute);
            if (h.id >= 0) {
                shade_hit(s, st, h, pos, r, brute, rgb, shadow_rays, g);
            } else if ((st.flags & VLB_BAKE_SKYBOX_ON_MISS) && s.skyW > 0) {
                sky_lookup(s, r, rgb);
                if (st.flags & VLB_BAKE_SRGB_ENCODE) for (int c = 0; c < 3; ++c) rgb[c] = srgb1(rgb[c]);
            }
            if (st.flags & VLB_BAKE_QUANTIZE_RGBA8) for (int c = 0; c < 3; ++c) rgb[c] = quant8(rgb[c]);
            if (image_rgb) {
                float* p = image_rgb + ((size_t)y * dt.W + x) * 3;
                p[0] = rgb[0]; p[1] = rgb[1]; p[2] = rgb[2];
            }
            float b[25];
            sh_basis25((st.flags & VLB_BAKE_SH_WORLD_FRAME) ? r : t, b);
            const float w = dt.w[y];
            for (int i = 0; i < K; ++i) {
                const double bw = (double)b[i] * (double)w;
                acc48[i * 3 + 0] += bw * rgb[0];
                acc48[i * 3 + 1] += bw * rgb[1];
                acc48[i * 3 + 2] += bw * rgb[2];
            }
        }
    }
}

}  // namespace

// ==========================================================================================
extern "C" {

int vo_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void vo_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

// n directions (xyz) -> n x 25 basis values (sh_common.h:26-152)
void vo_sh_basis(const float* dirs, uint64_t n, float* out25) {
    for (uint64_t i = 0; i < n; ++i) sh_basis25(v3(dirs[i * 3], dirs[i * 3 + 1], dirs[i * 3 + 2]), out25 + i * 25);
}
float vo_x2phi(int x, int w) { return x2phi(x, w); }
float vo_y2theta(int y, int h) { return y2theta(y, h); }
void vo_to_vector(float phi, float theta, float* out3) {
    V3 v = to_vector(phi, theta);
    out3[0] = v.x; out3[1] = v.y; out3[2] = v.z;
}
float vo_srgb(float c) { return srgb1(c); }
void vo_dir2uv(const float dir[3], float uv[2]) { dir_to_sky_uv(v3(dir[0], dir[1], dir[2]), &uv[0], &uv[1]); }

// probe direction table: t (un-swizzled), ray dir r = t.xzy, weight per texel
void vo_probe_dirs(int W, int H, float* t_out, float* r_out, float* w_out) {
    DirTable d = make_dirs(W, H);
    for (size_t i = 0; i < d.t.size(); ++i) {
        if (t_out) { t_out[i * 3] = d.t[i].x; t_out[i * 3 + 1] = d.t[i].y; t_out[i * 3 + 2] = d.t[i].z; }
        if (r_out) { r_out[i * 3] = d.t[i].x; r_out[i * 3 + 1] = d.t[i].z; r_out[i * 3 + 2] = d.t[i].y; }
        if (w_out) w_out[i] = d.w[i / W];
    }
}

void vo_skybox_project(const void* texels, int fmt, int W, int H, int order, float* out48) {
    project_image(texels, fmt, W, H, order, true, out48);
}
void vo_envmap_project(const void* texels, int fmt, int W, int H, int order, float* out48) {
    project_image(texels, fmt, W, H, order, false, out48);
}

// sh_sum.comp:35-52 without the x1250 debug gain: reconstruct radiance from coefficients.
void vo_sh_reconstruct(const float* coeffs48, int order, int W, int H, float* rgb_out) {
    const int K = n_coeffs(order);
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            V3 d = to_vector(x2phi(x, W), y2theta(y, H));
            float b[25]; sh_basis25(d, b);
            double c[3] = {0, 0, 0};
            for (int i = 0; i < K; ++i) for (int k = 0; k < 3; ++k) c[k] += (double)coeffs48[i * 3 + k] * b[i];
            for (int k = 0; k < 3; ++k) rgb_out[((size_t)y * W + x) * 3 + k] = (float)c[k];
        }
}

void vo_probe_positions(const vlb_bake_settings* st, float* out_xyz) {
    std::vector<float> px, py, pz;
    axis_coords(st->origin[0], st->step[0], st->probes[0], px);
    axis_coords(st->origin[1], st->step[1], st->probes[1], py);
    axis_coords(st->origin[2], st->step[2], st->probes[2], pz);
    const int Nx = st->probes[0], Ny = st->probes[1], Nz = st->probes[2];
    for (int k = 0; k < Nz; ++k) for (int j = 0; j < Ny; ++j) for (int i = 0; i < Nx; ++i) {
        size_t idx = (st->flags & VLB_BAKE_REFERENCE_PROBE_ORDER) ? ref_order_index(i, j, k, Nx, Ny, Nz)
                                                                 : (size_t)i + (size_t)j * Nx + (size_t)k * Nx * Ny;
        out_xyz[idx * 3] = px[i]; out_xyz[idx * 3 + 1] = py[j]; out_xyz[idx * 3 + 2] = pz[k];
    }
}

// literal restatement of light_baker.cpp:80-101 (vector doubling per dimension), used to pin
// ref_order_index + axis_coords above.
uint64_t vo_probe_positions_literal(const float bounds[6], const int counts[3], float* out_xyz, float step_out[3]) {
    std::vector<V3> positions;
    positions.push_back(v3(bounds[0], bounds[1], bounds[2]));
    float step[3];
    for (int d = 0; d < 3; ++d) step[d] = (bounds[3 + d] - bounds[d]) / ((float)counts[d] - 1.f);
    for (int dim = 0; dim < 3; ++dim) {
        std::vector<V3> copy = positions;
        for (V3 p : copy) {
            for (int i = 0; i < counts[dim] - 1; ++i) {
                float* c = dim == 0 ? &p.x : (dim == 1 ? &p.y : &p.z);
                *c += step[dim];
                positions.push_back(p);
            }
        }
    }
    for (size_t i = 0; i < positions.size(); ++i) {
        out_xyz[i * 3] = positions[i].x; out_xyz[i * 3 + 1] = positions[i].y; out_xyz[i * 3 + 2] = positions[i].z;
    }
    if (step_out) { step_out[0] = step[0]; step_out[1] = step[1]; step_out[2] = step[2]; }
    return positions.size();
}

void* vo_scene_create(const vlb_vertex* verts, uint64_t n_verts, const uint32_t* indices,
                      uint64_t n_indices, const vlb_instance* insts, uint32_t n_insts,
                      const vlb_material* mats, uint32_t n_mats) {
    (void)n_verts; (void)n_indices;
    Scene* s = new Scene();
    s->mats.assign(mats, mats + n_mats);
    for (int k = 0; k < 6; ++k) s->ref_bounds[k] = 0.f;            // scene_manager.hpp:170 bounds{}
    for (uint32_t ii = 0; ii < n_insts; ++ii) {
        const vlb_instance& vi = insts[ii];
        Inst in;
        std::memcpy(in.m, vi.transform, sizeof(in.m));
        inverse3x3(in.m, in.nm);
        in.material = vi.material_index < n_mats ? vi.material_index : (n_mats ? n_mats - 1 : 0);
        s->insts.push_back(in);
        // reference bounds: local AABB corners through the node matrix (scene_manager.cpp:497-507)
        if (vi.vertex_count) {
            float lo[3] = {3.4e38f, 3.4e38f, 3.4e38f}, hi[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
            for (uint32_t v = 0; v < vi.vertex_count; ++v) {
                const float* p = verts[vi.first_vertex + v].position;
                for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], p[k]); hi[k] = std::max(hi[k], p[k]); }
            }
            V3 a = xform_point(in.m, v3(lo[0], lo[1], lo[2])), b = xform_point(in.m, v3(hi[0], hi[1], hi[2]));
            s->ref_bounds[0] = std::min(s->ref_bounds[0], a.x); s->ref_bounds[1] = std::min(s->ref_bounds[1], a.y);
            s->ref_bounds[2] = std::min(s->ref_bounds[2], a.z);
            s->ref_bounds[3] = std::max(s->ref_bounds[3], b.x); s->ref_bounds[4] = std::max(s->ref_bounds[4], b.y);
            s->ref_bounds[5] = std::max(s->ref_bounds[5], b.z);
        }
        const uint32_t ntri = vi.index_count / 3;
        for (uint32_t t = 0; t < ntri; ++t) {
            const uint32_t* ix = indices + vi.first_index + (size_t)t * 3;
            const vlb_vertex& a = verts[vi.first_vertex + ix[0]];
            const vlb_vertex& b = verts[vi.first_vertex + ix[1]];
            const vlb_vertex& c = verts[vi.first_vertex + ix[2]];
            Tri tr;
            const V3 p0 = xform_point(in.m, v3(a.position[0], a.position[1], a.position[2]));
            const V3 p1 = xform_point(in.m, v3(b.position[0], b.position[1], b.position[2]));
            const V3 p2 = xform_point(in.m, v3(c.position[0], c.position[1], c.position[2]));
            tr.v0 = p0; tr.e1 = p1 - p0; tr.e2 = p2 - p0;
            tr.n0 = v3(a.normal[0], a.normal[1], a.normal[2]);
            tr.n1 = v3(b.normal[0], b.normal[1], b.normal[2]);
            tr.n2 = v3(c.normal[0], c.normal[1], c.normal[2]);
            tr.uv[0] = a.uv0[0]; tr.uv[1] = a.uv0[1]; tr.uv[2] = b.uv0[0]; tr.uv[3] = b.uv0[1];
            tr.uv[4] = c.uv0[0]; tr.uv[5] = c.uv0[1];
            tr.inst = ii;
            s->tris.push_back(tr);
        }
    }
    Builder(*s).run();
    return s;
}
void vo_scene_destroy(void* h) { delete (Scene*)h; }
uint64_t vo_scene_num_triangles(void* h) { return ((Scene*)h)->tris.size(); }
void vo_scene_bounds(void* h, int tight, float out6[6]) {
    Scene* s = (Scene*)h;
    std::memcpy(out6, tight ? s->tight_bounds : s->ref_bounds, 6 * sizeof(float));
}
void vo_scene_set_textures(void* h, const vlb_texture* tex, uint32_t n) {
    Scene* s = (Scene*)h;
    s->textures.assign(n, Scene::Tex());
    for (uint32_t i = 0; i < n; ++i) {
        Scene::Tex& t = s->textures[i];
        t.W = tex[i].width; t.H = tex[i].height; t.wrap_u = tex[i].wrap_u; t.wrap_v = tex[i].wrap_v; t.filter = tex[i].filter;
        const uint8_t* p = static_cast<const uint8_t*>(tex[i].texels);
        t.texels.assign(p, p + (size_t)t.W * t.H * 4);
    }
}
void vo_tex_sample(void* h, uint32_t tex, const float* uv, uint64_t n, float* rgb_out) {
    Scene* s = (Scene*)h;
    for (uint64_t i = 0; i < n; ++i) tex_sample(s->textures[tex], uv[2 * i], uv[2 * i + 1], rgb_out + 3 * i);
}
void vo_scene_set_skybox(void* h, const void* texels, int fmt, int W, int H) {
    Scene* s = (Scene*)h;
    s->skyW = W; s->skyH = H; s->sky.resize((size_t)W * H * 4);
    for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) {
        float rgb[3]; fetch_texel(texels, fmt, W, x, y, rgb);
        float* p = &s->sky[((size_t)y * W + x) * 4];
        p[0] = rgb[0]; p[1] = rgb[1]; p[2] = rgb[2]; p[3] = 1.f;
    }
}
// world-space triangle soup as the oracle flattened it: v0,e1,e2 (9 floats per triangle)
// texture(skybox, dir2SkyboxUV(dir)) of main.rmiss:39 (before sRGB) and getBaseColor of env_map.rchit:36-49
void vo_sky_lookup(void* h, const float dir[3], float rgb[3]) { sky_lookup(*(Scene*)h, v3(dir[0], dir[1], dir[2]), rgb); }
void vo_base_color(void* h, uint32_t material, float u, float v, float out4[4]) {
    Scene* s = (Scene*)h;
    base_color(*s, s->mats[material], u, v, out4);
}

void vo_scene_triangles(void* h, float* out9) {
    Scene* s = (Scene*)h;
    for (size_t i = 0; i < s->tris.size(); ++i) {
        const Tri& t = s->tris[i];
        float* o = out9 + i * 9;
        o[0] = t.v0.x; o[1] = t.v0.y; o[2] = t.v0.z; o[3] = t.e1.x; o[4] = t.e1.y; o[5] = t.e1.z;
        o[6] = t.e2.x; o[7] = t.e2.y; o[8] = t.e2.z;
    }
}

void vo_trace_rays(void* h, const float* origins, const float* dirs, uint64_t n, float tmin,
                   float tmax, int accel, int kind, int32_t* ids, float* tuv) {
    Scene* s = (Scene*)h;
    const bool brute = accel == VLB_TRACE_BRUTE_FORCE;
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        V3 o = v3(origins[i * 3], origins[i * 3 + 1], origins[i * 3 + 2]);
        V3 d = v3(dirs[i * 3], dirs[i * 3 + 1], dirs[i * 3 + 2]);
        Hit hit{-1, tmax, 0.f, 0.f};
        if (kind == VLB_TRACE_ANY) {
            Hit a;
            if (trace_any(*s, o, d, tmin, tmax, brute, &a)) hit = a;
        } else {
            hit = trace_closest(*s, o, d, tmin, tmax, brute);
        }
        ids[i] = hit.id;
        if (tuv) { tuv[i * 3] = hit.t; tuv[i * 3 + 1] = hit.u; tuv[i * 3 + 2] = hit.v; }
    }
}

// Bakes the probes listed in probe_ids (indices in x-fastest grid order, i + j*Nx + k*Nx*Ny) or,
// when probe_ids == NULL, the slices slab_k0, slab_k0 + slab_stride, ... < slab_k1. out = n x 48 floats in list order
// (slab mode: output order as the settings' flags say, relative to the slab start).
// Returns the number of shadow rays traced.
static uint64_t bake_probes_impl(void* h, const vlb_bake_settings* st, const float* prev_full, const int64_t* probe_ids,
                                 uint64_t n_ids, int brute, float* out) {
    Scene* s = (Scene*)h;
    const int Nx = st->probes[0], Ny = st->probes[1], Nz = st->probes[2];
    std::vector<float> px, py, pz;
    axis_coords(st->origin[0], st->step[0], Nx, px);
    axis_coords(st->origin[1], st->step[1], Ny, py);
    axis_coords(st->origin[2], st->step[2], Nz, pz);
    const DirTable dt = make_dirs(st->dir_w, st->dir_h);
    const int k0 = st->slab_k1 < 0 ? 0 : st->slab_k0, k1 = st->slab_k1 < 0 ? Nz : st->slab_k1;
    const int kstride = st->slab_stride > 1 ? st->slab_stride : 1;      // slices k0, k0 + stride, ... < k1
    const uint64_t n = probe_ids ? n_ids : (uint64_t)Nx * Ny * ((k1 - k0 + kstride - 1) / kstride);
    const bool ref_order = !probe_ids && (st->flags & VLB_BAKE_REFERENCE_PROBE_ORDER);
    uint64_t shadow_total = 0;
    GatherCtx gctx{prev_full, px.data(), py.data(), pz.data(), Nx, Ny, Nz};
    std::vector<double> all((size_t)n * 48, 0.0);
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : shadow_total)
    for (int64_t q = 0; q < (int64_t)n; ++q) {
        const int64_t nxy = (int64_t)Nx * Ny;
        const int64_t g = probe_ids ? probe_ids[q] : q % nxy + (k0 + (q / nxy) * kstride) * nxy;
        const int i = (int)(g % Nx), j = (int)((g / Nx) % Ny), k = (int)(g / ((int64_t)Nx * Ny));
        uint64_t sr = 0;
        size_t slot = (size_t)q;
        if (ref_order) slot = ref_order_index(i, j, k, Nx, Ny, Nz);     // whole-grid only
        bake_one(*s, *st, dt, v3(px[i], py[j], pz[k]), brute != 0, &all[slot * 48], nullptr, &sr, prev_full ? &gctx : nullptr);
        shadow_total += sr;
    }
    if (st->flags & VLB_BAKE_ACCUMULATE_ACROSS_PROBES) {               // light_baker.cpp:110-121 literal
        for (uint64_t q = 1; q < n; ++q) for (int c = 0; c < 48; ++c) all[q * 48 + c] += all[(q - 1) * 48 + c];
    }
    for (size_t c = 0; c < all.size(); ++c) out[c] = (float)all[c];
    return shadow_total;
}

uint64_t vo_bake_probes(void* h, const vlb_bake_settings* st, const int64_t* probe_ids,
                        uint64_t n_ids, int brute, float* out) {
    return bake_probes_impl(h, st, nullptr, probe_ids, n_ids, brute, out);
}

// One gather pass (include/vlb_bake.h: vlb_bake_gather_device): as vo_bake_probes, every hit adding the
// gather of prev_full ([Nx*Ny*Nz][48], x-fastest). prev_full == NULL is the direct pass.
uint64_t vo_bake_gather(void* h, const vlb_bake_settings* st, const float* prev_full, const int64_t* probe_ids,
                        uint64_t n_ids, int brute, float* out) {
    return bake_probes_impl(h, st, prev_full, probe_ids, n_ids, brute, out);
}

// One probe's environment image (what env_map.rgen writes), W*H*3 floats, plus its SH.
void vo_probe_envmap(void* h, const vlb_bake_settings* st, const float pos[3], int brute,
                     float* image_rgb, float* out48) {
    Scene* s = (Scene*)h;
    const DirTable dt = make_dirs(st->dir_w, st->dir_h);
    double acc[48] = {0};
    bake_one(*s, *st, dt, v3(pos[0], pos[1], pos[2]), brute != 0, acc, image_rgb, nullptr);
    if (out48) for (int c = 0; c < 48; ++c) out48[c] = (float)acc[c];
}

// The same for a gather pass: every hit adds the gather of prev_full over the settings' grid (the image the viewer's
// main.rchit would shade from this position; tests/test_ref_shaders.py holds it to that shader).
void vo_probe_envmap_gather(void* h, const vlb_bake_settings* st, const float pos[3], const float* prev_full, int brute,
                            float* image_rgb, float* out48) {
    Scene* s = (Scene*)h;
    std::vector<float> px, py, pz;
    axis_coords(st->origin[0], st->step[0], st->probes[0], px);
    axis_coords(st->origin[1], st->step[1], st->probes[1], py);
    axis_coords(st->origin[2], st->step[2], st->probes[2], pz);
    GatherCtx gctx{prev_full, px.data(), py.data(), pz.data(), st->probes[0], st->probes[1], st->probes[2]};
    const DirTable dt = make_dirs(st->dir_w, st->dir_h);
    double acc[48] = {0};
    bake_one(*s, *st, dt, v3(pos[0], pos[1], pos[2]), brute != 0, acc, image_rgb, nullptr, prev_full ? &gctx : nullptr);
    if (out48) for (int c = 0; c < 48; ++c) out48[c] = (float)acc[c];
}

}  // extern "C"

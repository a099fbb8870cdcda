// ref_pipeline.cpp — the reference's bake ray-tracing pipeline re-assembled around its OWN shader code (TEST
// INFRASTRUCTURE, built into oracle/_ref/libvlb_refshaders.so by oracle/make_ref_shaders.py).
//
// What is reference code here: every shader body (env_map.rgen, env_map.rchit, main.rmiss, shadow.rmiss, sh.comp),
// compiled from /root/reference/shaders behind oracle/glsl_shim.h. What is NOT: the glue below, which plays the part of
// the Vulkan ray-tracing pipeline the reference builds in src/baker/env_map_generator.cpp:200-290 -- raygen =
// env_map.rgen, miss 0 = main.rmiss, miss 1 = shadow.rmiss, hit group = env_map.rchit -- and of the driver's
// traversal: ray / triangle intersection is delegated to the caller's callback (the CPU oracle's vo_trace_rays), because
// in the reference it is the Vulkan driver's and pinned by no reference test ("parity unpinned").
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../include/vlb_bake.h"

namespace glsl {
typedef unsigned int uint;
typedef void (*trace_fn)(uint flags, uint miss_index, const float* origin, float tmin, const float* dir, float tmax, int payload);
trace_fn g_trace = nullptr;
}

extern "C" {
// the glue translation units (oracle/ref_glue/*.cpp)
void ref_rchit_set_light(const float p[3]);
void ref_rchit_set_in_shadow(int v);
void ref_rchit_run(const void* instance_infos, const void* materials, const void* samplers, int instance, int primitive,
                   const float bary_uv[2], const float origin[3], const float dir[3], float t, const float w2o[12], float color_out[3]);
void ref_rmiss_run(const float dir[3], const float* sky_texels, int w, int h, float rgb_out[3]);
int ref_shadow_rmiss_run(void);
void ref_rgen_set_color(const float rgb[3]);
void ref_rgen_run(int x, int y, int W, int H, const float origin[3], float* image_texels, int rgba8);
void ref_sh_comp_dispatch(float* texels, int W, int H, double* out48);
// the viewer's closest-hit shader (main.rchit) and the miss shader of its probe-visibility rays (sh.rmiss)
void ref_main_rchit_set_constants(const float grid_step[3], unsigned lmax, const float light[3], float shadow_bias, float ambient,
                                  float c_diffuse, float c_specular, float c_gloss);
void ref_main_rchit_set_in_shadow(int v);
void ref_main_rchit_get_sh_payload(float sum[3], float normal[3], int ijk[3], unsigned* lmax, int* occluded);
void ref_main_rchit_set_sh_payload(const float sum[3], int occluded);
void ref_main_rchit_run(const void* instance_infos, const void* materials, const void* samplers, int instance, int primitive,
                        const float bary_uv[2], const float origin[3], const float dir[3], float t, const float w2o[12], float color_out[3]);
void ref_sh_rmiss_run(float sum[3], const float normal[3], const int ijk[3], unsigned lmax, int* occluded, const float* sh);
}

typedef void (*vo_trace_fn)(void* h, const float* o, const float* d, uint64_t n, float tmin, float tmax, int accel, int kind,
                            int32_t* ids, float* tuv);

namespace {

struct Sampler { const float* texels; int w, h; };      // layout of glsl::sampler2D

struct Pipeline {
    std::vector<uint64_t> instance_info;                 // shader::InstanceInfo: {vertex address, index address, material}
    std::vector<uint32_t> tri_offset;                    // flat triangle id of each instance's first primitive
    std::vector<float> w2o;                              // 12 floats per instance: 4 columns of gl_WorldToObjectEXT
    const vlb_material* materials = nullptr;
    std::vector<Sampler> samplers;
    const float* sky = nullptr; int sky_w = 0, sky_h = 0;
    vo_trace_fn trace = nullptr; void* oracle_scene = nullptr;
    uint32_t flags = 0;
    uint64_t shadow_rays = 0;
    bool viewer_hit = false;                             // closest-hit shader: env_map.rchit (the bake) or main.rchit (the viewer's gather)
    const float* sh_coeffs = nullptr;                    // SHCoeffs buffer of sh.rmiss: (lmax+1)^2 vec3 per probe of the 7x7x7 grid
};
Pipeline* g_cur = nullptr;

// inverse of the affine 3x4 row-major object->world matrix, as the 4 columns of the 3-row world->object matrix
void inverse_columns(const float* m, float* out12) {
    const double a = m[0], b = m[1], c = m[2], d = m[4], e = m[5], f = m[6], g = m[8], h = m[9], i = m[10];
    const double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
    const double inv[9] = {(e * i - f * h) / det, (c * h - b * i) / det, (b * f - c * e) / det,
                           (f * g - d * i) / det, (a * i - c * g) / det, (c * d - a * f) / det,
                           (d * h - e * g) / det, (b * g - a * h) / det, (a * e - b * d) / det};   // row-major
    const double t[3] = {m[3], m[7], m[11]};
    for (int col = 0; col < 3; ++col)
        for (int row = 0; row < 3; ++row) out12[3 * col + row] = (float)inv[3 * row + col];
    for (int row = 0; row < 3; ++row) out12[9 + row] = (float)-(inv[3 * row] * t[0] + inv[3 * row + 1] * t[1] + inv[3 * row + 2] * t[2]);
}

// traceRayEXT of the shaders: intersection by the oracle, then the shader the pipeline's binding table selects
void pipeline_trace(unsigned flags, unsigned miss_index, const float* o, float tmin, const float* d, float tmax, int /*payload*/) {
    Pipeline* p = g_cur;
    int32_t id = -1; float tuv[3] = {0.f, 0.f, 0.f};
    if (flags & 4u) {                                    // gl_RayFlagsTerminateOnFirstHitEXT | SkipClosestHitShader: shadow ray
        ++p->shadow_rays;
        p->trace(p->oracle_scene, o, d, 1, tmin, tmax, VLB_TRACE_BVH, VLB_TRACE_ANY, &id, tuv);
        if (id >= 0) return;                             // occluded: no shader runs (SkipClosestHitShader)
        if (miss_index == 1) {                           // shadow.rmiss -> payload 1 of whichever hit shader is bound
            if (p->viewer_hit) ref_main_rchit_set_in_shadow(ref_shadow_rmiss_run());
            else ref_rchit_set_in_shadow(ref_shadow_rmiss_run());
        } else if (miss_index == 2 && p->viewer_hit) {   // sh.rmiss -> payload 2 (main.rchit:153)
            float sum[3], normal[3]; int ijk[3], occ; unsigned lmax;
            ref_main_rchit_get_sh_payload(sum, normal, ijk, &lmax, &occ);
            ref_sh_rmiss_run(sum, normal, ijk, lmax, &occ, p->sh_coeffs);
            ref_main_rchit_set_sh_payload(sum, occ);
        }                                                // miss 3 (skybox_sh.rmiss, main.rchit:121): only read when gridStep == 0
        return;
    }
    p->trace(p->oracle_scene, o, d, 1, tmin, tmax, VLB_TRACE_BVH, VLB_TRACE_CLOSEST, &id, tuv);
    float rgb[3] = {0.f, 0.f, 0.f};
    if (id >= 0) {
        size_t inst = 0;
        while (inst + 1 < p->tri_offset.size() && p->tri_offset[inst + 1] <= (uint32_t)id) ++inst;
        const float bary[2] = {tuv[1], tuv[2]};
        (p->viewer_hit ? ref_main_rchit_run : ref_rchit_run)(p->instance_info.data(), p->materials, p->samplers.data(), (int)inst,
                                                             id - (int)p->tri_offset[inst], bary, o, d, tuv[0], &p->w2o[12 * inst], rgb);
        ref_rgen_set_color(rgb);
    } else if ((p->flags & VLB_BAKE_SKYBOX_ON_MISS) && p->sky && miss_index == 0) {
        ref_rmiss_run(d, p->sky, p->sky_w, p->sky_h, rgb);                                  // main.rmiss -> payload 0
        ref_rgen_set_color(rgb);
    }
}

}  // namespace

extern "C" {

void* rp_create(const vlb_vertex* verts, const uint32_t* indices, const vlb_instance* insts, uint32_t n_insts,
                const vlb_material* mats, vo_trace_fn trace, void* oracle_scene) {
    Pipeline* p = new Pipeline();
    p->materials = mats; p->trace = trace; p->oracle_scene = oracle_scene;
    uint32_t off = 0;
    for (uint32_t i = 0; i < n_insts; ++i) {
        p->instance_info.push_back((uint64_t)(uintptr_t)(verts + insts[i].first_vertex));
        p->instance_info.push_back((uint64_t)(uintptr_t)(indices + insts[i].first_index));
        p->instance_info.push_back(insts[i].material_index);
        p->tri_offset.push_back(off);
        off += insts[i].index_count / 3;
        float c[12];
        inverse_columns(insts[i].transform, c);
        p->w2o.insert(p->w2o.end(), c, c + 12);
    }
    p->tri_offset.push_back(off);
    return p;
}
void rp_destroy(void* h) { delete static_cast<Pipeline*>(h); }
void rp_set_skybox(void* h, const float* texels, int w, int height) {
    Pipeline* p = static_cast<Pipeline*>(h);
    p->sky = texels; p->sky_w = w; p->sky_h = height;
}
void rp_set_textures(void* h, const float* const* texels, const int* wh, uint32_t n) {
    Pipeline* p = static_cast<Pipeline*>(h);
    p->samplers.clear();
    for (uint32_t i = 0; i < n; ++i) p->samplers.push_back(Sampler{texels[i], wh[2 * i], wh[2 * i + 1]});
}

// One probe as LightBaker::bake does it (light_baker.cpp:298-325): getMap (a W x H launch of env_map.rgen into the
// environment image) then dispatchBakingKernel (sh.comp over that image). image: W*H*4 floats. Returns the shadow rays.
uint64_t rp_bake_probe(void* h, const float origin[3], int W, int H, uint32_t flags, const float light[3], float* image, double* coeffs48) {
    Pipeline* p = static_cast<Pipeline*>(h);
    g_cur = p; glsl::g_trace = pipeline_trace;
    p->flags = flags; p->shadow_rays = 0;
    ref_rchit_set_light(light);
    std::memset(image, 0, sizeof(float) * 4 * (size_t)W * H);
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) ref_rgen_run(x, y, W, H, origin, image, (flags & VLB_BAKE_QUANTIZE_RGBA8) ? 1 : 0);
    ref_sh_comp_dispatch(image, W, H, coeffs48);
    g_cur = nullptr;
    return p->shadow_rays;
}

// The same probe with the VIEWER's closest-hit shader bound instead of the bake's: main.rchit shades every hit with direct
// light plus its gather over the probe grid (main.rchit:124-167; probe-visibility rays with sh.rmiss as their miss shader).
// This is the reference code the multi-bounce passes of BASELINE configs[3] iterate. sh_coeffs: the SHCoeffs buffer the
// shader indexes -- (lmax + 1)^2 vec3 per probe, 7 x 7 x 7 probes, x-fastest (sh.rmiss:22-25). Push constants as
// main.rchit:38-48.
uint64_t rp_bake_probe_viewer_hit(void* h, const float origin[3], int W, int H, uint32_t flags, const float light[3], const float grid_step[3],
                                  unsigned lmax, float shadow_bias, float ambient, float c_diffuse, float c_specular, float c_gloss,
                                  const float* sh_coeffs, float* image, double* coeffs48) {
    Pipeline* p = static_cast<Pipeline*>(h);
    g_cur = p; glsl::g_trace = pipeline_trace;
    p->flags = flags; p->shadow_rays = 0; p->viewer_hit = true; p->sh_coeffs = sh_coeffs;
    ref_main_rchit_set_constants(grid_step, lmax, light, shadow_bias, ambient, c_diffuse, c_specular, c_gloss);
    std::memset(image, 0, sizeof(float) * 4 * (size_t)W * H);
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) ref_rgen_run(x, y, W, H, origin, image, (flags & VLB_BAKE_QUANTIZE_RGBA8) ? 1 : 0);
    ref_sh_comp_dispatch(image, W, H, coeffs48);
    p->viewer_hit = false; p->sh_coeffs = nullptr;
    g_cur = nullptr;
    return p->shadow_rays;
}

}  // extern "C"

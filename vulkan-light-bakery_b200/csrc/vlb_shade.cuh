// vlb_shade.cuh — radiance of one probe ray: the CUDA restatement of what the reference does in
// shaders/env_map.rgen:18-28 (ray), env_map.rchit:51-102 (hit shading + shadow ray),
// main.rmiss:18-41 (sky lookup) and shadow.rmiss:6-9 (visibility).
#pragma once

#include "vlb_bvh.cuh"

namespace vlb {

// Device-resident shading inputs (all 16-byte records):
//   tri_shade[3*id+0..2]  id = flat triangle id: (n0.xyz, bits(instance)), (n1.xyz, 0), (n2.xyz, 0)
//                         OBJECT-space vertex normals, as env_map.rchit:59-64 fetches them
//   inst[3*k+0..2]        (Minv row0.xyz, bits(material)), (Minv row1.xyz, 0), (Minv row2.xyz, 0)
//   base_color[m]         resolved baseColor of material m (env_map.rchit:36-49, factor path)
struct ShadeView {
    const float4* tri_shade;
    const float4* inst;
    const float4* base_color;   // rgb = resolved factor; w = bits(baseColor texture index, -1 = none)
    const float4* tri_uv;       // 2 float4 per flat triangle: (u0, v0, u1, v1), (u2, v2, 0, 0) = Vertex::uv0
    const int4* tex_desc;       // per texture: (first texel, width, height, wrap_u | wrap_v << 2 | filter << 4)
    const uchar4* tex_texels;   // RGBA8 texels of all textures, back to back
    const float4* sky;       // RGBA32F texels, NULL if no skybox set
    int sky_w, sky_h;
};

struct BakeConsts {
    float light[3];
    float shadow_bias, c_diffuse, c_specular, gloss, ambient;
    float tmin, tmax;
    uint32_t flags;
};

VLB_HD float glsl_mod(float x, float y) { return x - y * floorf(f_div_c(x, y)); }
VLB_HD float clampf(float x, float a, float b) { return fminf(fmaxf(x, a), b); }
VLB_HD int wrapi(int i, int n) { int r = i % n; return r < 0 ? r + n : r; }
VLB_HD int mini(int a, int b) { return a < b ? a : b; }
// wrapi for an index no further than n outside [0, n) -- the sky lookup's texels are in [-1, n] -- without the integer
// division (~20 instructions, four sites); anything else (NaN directions) is clamped into the map instead of wrapped.
VLB_HD int wrap_near(int i, int n) {
    i = i < 0 ? i + n : i;
    i = i >= n ? i - n : i;
    return i < 0 ? 0 : (i >= n ? n - 1 : i);
}

// main.rmiss:18-35 + bilinear / repeat lookup at LOD 0 (sampler: src/application.hpp:45-52)
VLB_HD void sky_lookup(const ShadeView& s, Vec3 dir, float rgb[3]) {
    float theta = acosf(clampf(dir.y, -1.0f, 1.0f));
    float phi = atan2f(dir.x, dir.z);
    theta = glsl_mod(theta, 2.0f * kPi);
    theta = clampf(theta, 0.0f, 2.0f * kPi);
    if (theta > kPi) { theta = 2.0f * kPi - theta; phi += kPi; }
    phi = glsl_mod(phi, 2.0f * kPi);
    phi = clampf(phi, 0.0f, 2.0f * kPi);
    const float u = f_div_c(phi, 2.0f * kPi), v = f_div_c(theta, kPi);
    const int W = s.sky_w, H = s.sky_h;
    const float fx = u * (float)W - 0.5f, fy = v * (float)H - 0.5f;
    const float flx = floorf(fx), fly = floorf(fy);
    const float ax = fx - flx, ay = fy - fly;
    const int x0 = wrap_near((int)flx, W), x1 = wrap_near((int)flx + 1, W);
    const int y0 = wrap_near((int)fly, H), y1 = wrap_near((int)fly + 1, H);
    const float4 p00 = ld4_stream(s.sky + (size_t)y0 * W + x0);
    const float4 p10 = ld4_stream(s.sky + (size_t)y0 * W + x1);
    const float4 p01 = ld4_stream(s.sky + (size_t)y1 * W + x0);
    const float4 p11 = ld4_stream(s.sky + (size_t)y1 * W + x1);
    const float t0 = p00.x + (p10.x - p00.x) * ax, b0 = p01.x + (p11.x - p01.x) * ax;
    const float t1 = p00.y + (p10.y - p00.y) * ax, b1 = p01.y + (p11.y - p01.y) * ax;
    const float t2 = p00.z + (p10.z - p00.z) * ax, b2 = p01.z + (p11.z - p01.z) * ax;
    rgb[0] = t0 + (b0 - t0) * ay;
    rgb[1] = t1 + (b1 - t1) * ay;
    rgb[2] = t2 + (b2 - t2) * ay;
}

// Texel index along one axis under a Vulkan address mode (VkSamplerAddressMode; loadSamplers,
// src/scene_manager.cpp:654-668). mode: VLB_WRAP_REPEAT / CLAMP_TO_EDGE / MIRRORED_REPEAT.
VLB_HD int wrap_texel(int i, int n, int mode) {
    if (mode == 1) return i < 0 ? 0 : (i >= n ? n - 1 : i);
    if (mode == 2) {
        const int m = wrapi(i, 2 * n);
        return m < n ? m : 2 * n - 1 - m;
    }
    return wrapi(i, n);
}

VLB_HD void texel_rgb(const ShadeView& s, int base, int W, int x, int y, float c[3]) {
#ifdef __CUDA_ARCH__
    const uchar4 t = __ldg(s.tex_texels + (size_t)base + (size_t)y * W + x);
#else
    const uchar4 t = s.tex_texels[(size_t)base + (size_t)y * W + x];
#endif
    c[0] = (float)t.x / 255.0f; c[1] = (float)t.y / 255.0f; c[2] = (float)t.z / 255.0f;   // unorm8
}

// texture(textures[tex], uv).rgb at the base level (env_map.rchit:42): unnormalised coordinate u * W - 0.5,
// the four neighbours under the sampler's address modes, fp32 weights (same form as sky_lookup).
VLB_HD void tex_sample(const ShadeView& s, int tex, float u, float v, float rgb[3]) {
#ifdef __CUDA_ARCH__
    const int4 d = __ldg(s.tex_desc + tex);
#else
    const int4 d = s.tex_desc[tex];
#endif
    const int W = d.y, H = d.z, mu = d.w & 3, mv = (d.w >> 2) & 3;
    if ((d.w >> 4) & 1) {                                                // nearest
        const int x = wrap_texel((int)floorf(f_mul(u, (float)W)), W, mu), y = wrap_texel((int)floorf(f_mul(v, (float)H)), H, mv);
        texel_rgb(s, d.x, W, x, y, rgb);
        return;
    }
    const float fx = f_sub(f_mul(u, (float)W), 0.5f), fy = f_sub(f_mul(v, (float)H), 0.5f);
    const float flx = floorf(fx), fly = floorf(fy);
    const float ax = fx - flx, ay = fy - fly;
    const int x0 = wrap_texel((int)flx, W, mu), x1 = wrap_texel((int)flx + 1, W, mu);
    const int y0 = wrap_texel((int)fly, H, mv), y1 = wrap_texel((int)fly + 1, H, mv);
    float p00[3], p10[3], p01[3], p11[3];
    texel_rgb(s, d.x, W, x0, y0, p00); texel_rgb(s, d.x, W, x1, y0, p10);
    texel_rgb(s, d.x, W, x0, y1, p01); texel_rgb(s, d.x, W, x1, y1, p11);
    for (int c = 0; c < 3; ++c) {
        const float top = p00[c] + (p10[c] - p00[c]) * ax;
        const float bot = p01[c] + (p11[c] - p01[c]) * ax;
        rgb[c] = top + (bot - top) * ay;
    }
}

// baseColor of a hit on a textured material: uv0 interpolated as env_map.rchit:65 does, then tex_sample.
// Out of line on the device: only hits on textured materials pay for it, and its registers stay out of the
// bake kernel's traversal loop (inlined it raised k_bake_stream's spills from 60 to 160 bytes, -2.5 %).
#ifdef __CUDA_ARCH__
__device__ __noinline__
#else
inline
#endif
float3 textured_base_color(const float4* tri_uv, const int4* tex_desc, const uchar4* tex_texels, int tex, int tri,
                           float b0, float b1, float b2) {
    ShadeView s{};
    s.tex_desc = tex_desc; s.tex_texels = tex_texels;
    const float4 ua = ld4_stream(tri_uv + 2 * (size_t)tri), ub = ld4_stream(tri_uv + 2 * (size_t)tri + 1);
    const float tu = f_add(f_add(f_mul(ua.x, b0), f_mul(ua.z, b1)), f_mul(ub.x, b2));   // :65
    const float tv = f_add(f_add(f_mul(ua.y, b0), f_mul(ua.w, b1)), f_mul(ub.y, b2));
    float rgb[3];
    tex_sample(s, tex, tu, tv, rgb);
    return make_float3(rgb[0], rgb[1], rgb[2]);
}

// imageStore to rgba8 (src/baker/env_map_generator.hpp:39): clamp, round-half-even to n/255. One out-of-line copy on the
// device (an IEEE division, nine call sites in the bake kernel's cold shading code; see pow_f).
#if defined(__CUDA_ARCH__) && VLB_POW_OUTLINE
static __device__ __noinline__
#else
VLB_HD
#endif
float quant8(float c) {
    return rintf(clampf(c, 0.0f, 1.0f) * 255.0f) / 255.0f;
}

// Hit shading up to the point where the shadow ray is needed. Returns false when no light
// reaches the point regardless of occlusion (sDotN == 0), i.e. no shadow ray is traced.
struct ShadePrelude {
    Vec3 N, Ln, so;     // shading normal, unit light vector, biased shadow-ray origin
    Vec3 P;             // hit position (gather passes only)
    float llen, sDotN;
    float bc[3];
};

// Source of a gather pass (include/vlb_bake.h: vlb_bake_gather_device): the previous pass over the
// whole grid and the grid itself. prev == NULL: direct pass.
struct GatherView {
    const float* prev;                                   // [Nx*Ny*Nz][48], x-fastest
    const float* px; const float* py; const float* pz;   // probe axis coordinates
    int Nx, Ny, Nz;
    float origin[3], step[3];
    float gain;
    int world_frame;
    // bake kernel only: per grid cell, the node (or leaf ref, or kNoChild) below which every triangle that reaches into
    // the cell lies (k_cell_roots, bake.cu); visibility rays of hits inside the cell start there instead of at the root.
    // NULL: start at the root.
    const int* cell_root;
    float cell_margin;
};

// TEX = false compiles the texture branch out: the bake kernel is instantiated both ways and scenes without
// textured materials run the lean one (the branch costs 2.1 % on C3 even when never taken: registers).
template <bool TEX = true>
VLB_HD bool shade_prelude(const ShadeView& s, const BakeConsts& c, const HitRec& h, Vec3 o, Vec3 r,
                          ShadePrelude& p) {
    const float4 a0 = ld4_stream(s.tri_shade + 3 * (size_t)h.id + 0);
    const float4 a1 = ld4_stream(s.tri_shade + 3 * (size_t)h.id + 1);
    const float4 a2 = ld4_stream(s.tri_shade + 3 * (size_t)h.id + 2);
    const int inst = f2i(a0.w);
    const float4 m0 = ld4(s.inst + 3 * (size_t)inst + 0);
    const float4 m1 = ld4(s.inst + 3 * (size_t)inst + 1);
    const float4 m2 = ld4(s.inst + 3 * (size_t)inst + 2);
    const float4 bc = ld4(s.base_color + f2i(m0.w));
    p.bc[0] = bc.x; p.bc[1] = bc.y; p.bc[2] = bc.z;
    const float b0 = 1.0f - h.u - h.v, b1 = h.u, b2 = h.v;              // env_map.rchit:63
    const int tex = f2i(bc.w);
    if (TEX && tex >= 0) {                                              // getBaseColor, texture branch (:40-43)
        const float3 c = textured_base_color(s.tri_uv, s.tex_desc, s.tex_texels, tex, h.id, b0, b1, b2);
        p.bc[0] = c.x; p.bc[1] = c.y; p.bc[2] = c.z;
    }
    const Vec3 nrm = mk3(a0.x * b0 + a1.x * b1 + a2.x * b2, a0.y * b0 + a1.y * b1 + a2.y * b2,
                         a0.z * b0 + a1.z * b1 + a2.z * b2);            // :64
    const float nm[9] = {m0.x, m0.y, m0.z, m1.x, m1.y, m1.z, m2.x, m2.y, m2.z};
    p.N = normalize_exact(xform_normal(nm, nrm));                       // :68
    const Vec3 P = mk3(f_fma(r.x, h.t, o.x), f_fma(r.y, h.t, o.y), f_fma(r.z, h.t, o.z));  // :67
    const Vec3 L = mk3(c.light[0] - P.x, c.light[1] - P.y, c.light[2] - P.z);             // :73
    p.llen = f_sqrt_c(dot_exact(L, L));
    p.Ln = mk3(f_div_c(L.x, p.llen), f_div_c(L.y, p.llen), f_div_c(L.z, p.llen));
    p.sDotN = fmaxf(dot_exact(p.Ln, p.N), 0.0f);                        // :79
    p.so = mk3(f_fma(c.shadow_bias, p.N.x, P.x), f_fma(c.shadow_bias, p.N.y, P.y),
               f_fma(c.shadow_bias, p.N.z, P.z));                       // :82
    p.P = P;
    return p.sDotN != 0.0f;                                             // :84
}

// Grid cell of coordinate p along one axis, clamped so that the cell's far corner exists
// (shaders/main.rchit:126 `floor(hitPosition / gridStep)`, for an arbitrary grid).
VLB_HD int gather_cell(float p, float origin, float step, int n) {
    if (n < 2) return 0;
    const float g = floorf(f_div_c(f_sub(p, origin), step));
    if (!(g > 0.0f)) return 0;                                          // also NaN (step == 0)
    return g >= (float)(n - 2) ? n - 2 : (int)g;
}

// Corner c (gridVertices order, shaders/main.rchit:128-137) of the grid cell (ci, cj, ck) around a hit at P: probe
// indices, vector from P to the probe and its length (= the visibility ray's direction and tmax, :145,154).
VLB_HD void gather_corner(const GatherView& g, Vec3 P, int ci, int cj, int ck, int c, int& i, int& j, int& k, Vec3& d, float& tmax) {
    i = mini(ci + ((c >> 2) & 1), g.Nx - 1); j = mini(cj + ((c >> 1) & 1), g.Ny - 1); k = mini(ck + (c & 1), g.Nz - 1);
    d = mk3(f_sub(g.px[i], P.x), f_sub(g.py[j], P.y), f_sub(g.pz[k], P.z));
    tmax = f_sqrt_c(dot_exact(d, d));
}

// The reference's run-time gather (shaders/main.rchit:124-163, probe lookup shaders/sh.rmiss:20-36)
// over the previous pass: visibility-weighted interpolation of the 8 probes around the hit, each
// probe's SH evaluated on the shading normal. Operation order = oracle gather_indirect. `occluded` has bit c set
// when the visibility ray to corner c was blocked (:155); the rays themselves are traced by the caller -- inline
// (gather_indirect below) or, in the bake kernel, as one warp-wide batch per 32 shaded hits (bake.cu).
template <int K>
VLB_HD void gather_accumulate(const GatherView& g, const ShadePrelude& p, unsigned occluded, float out[3]) {
    const int ci = gather_cell(p.P.x, g.origin[0], g.step[0], g.Nx);
    const int cj = gather_cell(p.P.y, g.origin[1], g.step[1], g.Ny);
    const int ck = gather_cell(p.P.z, g.origin[2], g.step[2], g.Nz);
    const float weight_max = f_sqrt_c(dot_exact(mk3(g.step[0], g.step[1], g.step[2]), mk3(g.step[0], g.step[1], g.step[2])));  // :141
    float basis[K];
    sh_basis<K>(g.world_frame ? p.N : mk3(p.N.x, p.N.z, p.N.y), basis);
    float sum[3] = {0.f, 0.f, 0.f}, wsum = 0.f;
#pragma unroll 1                                                        // one copy of the corner body: cold code (bake.cu: instruction cache)
    for (int c = 0; c < 8; ++c) {                                       // gridVertices order, :128-137
        if ((occluded >> c) & 1u) continue;
        int i, j, k; Vec3 d; float tmax;
        gather_corner(g, p.P, ci, cj, ck, c, i, j, k, d, tmax);
        const float w = fmaxf(f_sub(weight_max, tmax), 0.0f);           // :156
        const float* shp = g.prev + ((size_t)i + (size_t)g.Nx * ((size_t)j + (size_t)g.Ny * (size_t)k)) * 48;   // sh.rmiss:25
        // the probe's K x 3 coefficients as 16-byte loads (a probe record is 192 bytes, 16-byte aligned)
        constexpr int NQ = (K * 3 + 3) / 4;
        float sh[NQ * 4];
#pragma unroll
        for (int q4 = 0; q4 < NQ; ++q4) {
            const float4 t = ld4(reinterpret_cast<const float4*>(shp) + q4);
            sh[4 * q4] = t.x; sh[4 * q4 + 1] = t.y; sh[4 * q4 + 2] = t.z; sh[4 * q4 + 3] = t.w;
        }
        float v0 = 0.f, v1 = 0.f, v2 = 0.f;
#pragma unroll
        for (int q = 0; q < K; ++q) {                                   // sh.rmiss:27-34
            v0 = f_fma(sh[3 * q + 0], basis[q], v0); v1 = f_fma(sh[3 * q + 1], basis[q], v1); v2 = f_fma(sh[3 * q + 2], basis[q], v2);
        }
        sum[0] = f_fma(w, v0, sum[0]); sum[1] = f_fma(w, v1, sum[1]); sum[2] = f_fma(w, v2, sum[2]);   // :160
        wsum = f_add(wsum, w);                                          // :161
    }
    for (int c = 0; c < 3; ++c) out[c] = wsum > 0.0f ? f_mul(g.gain, f_div_c(sum[c], wsum)) : 0.0f;      // :164-165
}

// Gather with the 8 visibility rays traced right here by the calling thread (tests/emu, probe_ray_radiance).
template <int K, bool COUNT>
VLB_HD void gather_indirect(const BvhView& b, const GatherView& g, const ShadePrelude& p, float out[3], TraceCounters* cnt) {
    const int ci = gather_cell(p.P.x, g.origin[0], g.step[0], g.Nx);
    const int cj = gather_cell(p.P.y, g.origin[1], g.step[1], g.Ny);
    const int ck = gather_cell(p.P.z, g.origin[2], g.step[2], g.Nz);
    unsigned occluded = 0;
    for (int c = 0; c < 8; ++c) {
        int i, j, k; Vec3 d; float tmax;
        gather_corner(g, p.P, ci, cj, ck, c, i, j, k, d, tmax);
        if (tmax > 0.0f &&
            trace_any<COUNT>(b, p.so, mk3(f_div(d.x, tmax), f_div(d.y, tmax), f_div(d.z, tmax)), 0.0f, tmax, cnt, nullptr))  // :155
            occluded |= 1u << c;
    }
    gather_accumulate<K>(g, p, occluded, out);
}

// `ind` = gathered indirect term per channel (zeros in the direct pass: k + 0 == k, so the direct
// pass is bit-identical with or without it).
VLB_HD void shade_finish(const BakeConsts& c, const ShadePrelude& p, Vec3 r, bool in_shadow, const float ind[3], float rgb[3]) {
    float diffuse = 0.f, specular = 0.f;
    if (!in_shadow) {                                                   // env_map.rchit:90-99
        diffuse = c.c_diffuse * p.sDotN;
        const float dn = dot_exact(p.N, p.Ln);
        const Vec3 R = mk3(p.Ln.x - 2.0f * dn * p.N.x, p.Ln.y - 2.0f * dn * p.N.y, p.Ln.z - 2.0f * dn * p.N.z);
        const float rd = fmaxf(dot_exact(R, r), 0.0f);
        specular = c.c_specular * pow_f(rd, c.gloss);
    }
    const float k = c.ambient + diffuse + specular;
    for (int ch = 0; ch < 3; ++ch) {
        const float v = p.bc[ch] * (k + ind[ch]);
        rgb[ch] = (c.flags & 4u) ? srgb_encode(v) : v;                  // :101 (VLB_BAKE_SRGB_ENCODE)
    }
}

// Full radiance of one ray (primary + shadow), used by the bake kernel and by tests/emu.
template <bool COUNT, int K = 9>
VLB_HD void probe_ray_radiance(const BvhView& b, const ShadeView& s, const BakeConsts& c, Vec3 o, Vec3 r,
                               float rgb[3], TraceCounters* cnt, uint32_t* shadow_rays, const GatherView* g = nullptr) {
    rgb[0] = rgb[1] = rgb[2] = 0.0f;                                    // env_map.rgen:25
    const HitRec h = trace_closest<COUNT>(b, o, r, c.tmin, c.tmax, cnt);
    if (h.id >= 0) {
        ShadePrelude p;
        bool lit = shade_prelude(s, c, h, o, r, p);
        bool in_shadow = true;                                          // env_map.rchit:83
        if (lit) {
            if (c.flags & 1u) {                                         // VLB_BAKE_SHADOW_RAYS
                if (shadow_rays) ++*shadow_rays;
                in_shadow = trace_any<COUNT>(b, p.so, p.Ln, 0.0f, p.llen, cnt, nullptr);  // :87
            } else {
                in_shadow = false;
            }
        }
        float ind[3] = {0.f, 0.f, 0.f};
        if (g && g->prev) gather_indirect<K, COUNT>(b, *g, p, ind, cnt);
        shade_finish(c, p, r, in_shadow, ind, rgb);
    } else if ((c.flags & 2u) && s.sky) {                               // VLB_BAKE_SKYBOX_ON_MISS
        sky_lookup(s, r, rgb);
        if (c.flags & 4u) { rgb[0] = srgb_encode(rgb[0]); rgb[1] = srgb_encode(rgb[1]); rgb[2] = srgb_encode(rgb[2]); }
    }
    if (c.flags & 8u) { rgb[0] = quant8(rgb[0]); rgb[1] = quant8(rgb[1]); rgb[2] = quant8(rgb[2]); }  // QUANTIZE_RGBA8
}

}  // namespace vlb

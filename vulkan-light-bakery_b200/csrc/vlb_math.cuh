// vlb_math.cuh — arithmetic shared by every kernel of the bake path.
//
// Everything here is `__host__ __device__` so that tests/emu can run the very same code on the
// CPU to debug logic before GPU time is spent (TEST-ONLY; the product library never executes
// these on the host). Where results must be bit-identical between kernels (BVH traversal vs the
// brute-force intersector) and with the CPU oracle, operations are written with explicit
// round-to-nearest intrinsics so the compiler cannot contract or reassociate them.
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define VLB_HD __host__ __device__ __forceinline__
#else
#define VLB_HD inline
#endif

namespace vlb {

constexpr float kPi = 3.1415926538f;  // shaders/sh_common.h:1

// ---- exactly-rounded primitives ---------------------------------------------------------
VLB_HD float f_mul(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fmul_rn(a, b);
#else
    volatile float r = a * b; return r;
#endif
}
VLB_HD float f_add(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fadd_rn(a, b);
#else
    volatile float r = a + b; return r;
#endif
}
VLB_HD float f_sub(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fsub_rn(a, b);
#else
    volatile float r = a - b; return r;
#endif
}
VLB_HD float f_fma(float a, float b, float c) {
#ifdef __CUDA_ARCH__
    return __fmaf_rn(a, b, c);
#else
    return fmaf(a, b, c);
#endif
}
VLB_HD float f_div(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fdiv_rn(a, b);
#else
    volatile float r = a / b; return r;
#endif
}
VLB_HD float f_sqrt(float a) {
#ifdef __CUDA_ARCH__
    return __fsqrt_rn(a);
#else
    return sqrtf(a);
#endif
}

// The same division / square root for the cold code of the bake kernel (hit shading, sky lookup, gather): one
// out-of-line copy each on the device instead of ~10 inlined instructions per site (see pow_f below: cold straight-line
// code pays for every instruction-cache line). The traversal (intersect_tri, safe_inv) keeps the inlined forms.
#ifndef VLB_POW_OUTLINE
#define VLB_POW_OUTLINE 1
#endif
#ifndef VLB_COLD_OUTLINE
#define VLB_COLD_OUTLINE VLB_POW_OUTLINE
#endif
#if defined(__CUDA_ARCH__) && VLB_COLD_OUTLINE
static __device__ __noinline__ float f_div_c(float a, float b) { return __fdiv_rn(a, b); }
static __device__ __noinline__ float f_sqrt_c(float a) { return __fsqrt_rn(a); }
#else
VLB_HD float f_div_c(float a, float b) { return f_div(a, b); }
VLB_HD float f_sqrt_c(float a) { return f_sqrt(a); }
#endif

struct Vec3 {
    float x, y, z;
};
VLB_HD Vec3 mk3(float x, float y, float z) { Vec3 v; v.x = x; v.y = y; v.z = z; return v; }
VLB_HD float dot_exact(Vec3 a, Vec3 b) { return f_fma(a.z, b.z, f_fma(a.y, b.y, f_mul(a.x, b.x))); }

// normalize(): v / sqrt(dot(v,v)), fixed operation order (oracle: normalize3).
#ifndef VLB_NORMALIZE_OUTLINE
#define VLB_NORMALIZE_OUTLINE 0
#endif
#if defined(__CUDA_ARCH__) && VLB_NORMALIZE_OUTLINE
static __device__ __noinline__
#else
VLB_HD
#endif
Vec3 normalize_exact(Vec3 v) {
    const float l2 = f_fma(v.z, v.z, f_fma(v.y, v.y, f_mul(v.x, v.x)));
    const float len = f_sqrt_c(l2);
    return mk3(f_div_c(v.x, len), f_div_c(v.y, len), f_div_c(v.z, len));
}

// shaders/sh_common.h:8-12 from tabulated sin/cos (tables are built on the host in double and
// rounded to float, src: host_tables.cpp), so directions are bit-identical everywhere.
VLB_HD Vec3 to_vector_sc(float st, float ct, float cp, float sp) {
    return normalize_exact(mk3(f_mul(st, cp), f_mul(st, sp), ct));
}

// ---- SH basis, shaders/sh_common.h:26-104, index l(l+1)+m, 6-digit constants verbatim ------
template <int K>
VLB_HD void sh_basis(Vec3 d, float* o) {
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
    if (K > 9) {
        o[9] = -0.590044f * y * (3.0f * x * x - y * y);
        o[10] = 2.890611f * x * y * z;
        o[11] = -0.457046f * y * (4.0f * z * z - x * x - y * y);
        o[12] = 0.373176f * z * (2.0f * z * z - 3.0f * x * x - 3.0f * y * y);
        o[13] = -0.457046f * x * (4.0f * z * z - x * x - y * y);
        o[14] = 1.445306f * z * (x * x - y * y);
        o[15] = -0.590044f * x * (x * x - 3.0f * y * y);
    }
}

// powf of the shading code: ONE out-of-line copy on the device. The hit shading calls it up to 14 times (sRGB of three
// channels for the lit, the occluded and the sky outcome, the specular lobe), and inlined those expansions were a fifth of
// k_bake_stream's instructions -- cold straight-line code that missed the instruction cache on every line (ncu round 2:
// `no_instruction` = 45 % of the stall samples of the kernel's cold code). Same code, same bits; C3 148.7 -> 144.8 ms
// (profiles/r02_bake_icache_ab.log). VLB_POW_OUTLINE=0 builds the inlined form.
#if defined(__CUDA_ARCH__) && VLB_POW_OUTLINE
static __device__ __noinline__ float pow_f(float a, float b) { return powf(a, b); }
#else
VLB_HD float pow_f(float a, float b) { return powf(a, b); }
#endif

// shaders/env_map.rchit:27-34, one channel
VLB_HD float srgb_encode(float c) {
    return c < 0.0031308f ? c * 12.92f : 1.055f * pow_f(c, 1.0f / 2.4f) - 0.055f;
}

// ---- ray / triangle intersection: THE specification (oracle: intersect_tri) ------------------
// Triangle = v0, e1 = v1 - v0, e2 = v2 - v0 in world space. Returns true and (t,u,v).
VLB_HD bool intersect_tri(const float4 v0, const float4 e1, const float4 e2, Vec3 o, Vec3 d,
                          float& t, float& u, float& v) {
    const float px = f_fma(d.y, e2.z, -f_mul(d.z, e2.y));
    const float py = f_fma(d.z, e2.x, -f_mul(d.x, e2.z));
    const float pz = f_fma(d.x, e2.y, -f_mul(d.y, e2.x));
    const float det = f_fma(e1.z, pz, f_fma(e1.y, py, f_mul(e1.x, px)));
    if (det == 0.0f) return false;
    const float inv = f_div(1.0f, det);
    const float tx = f_sub(o.x, v0.x), ty = f_sub(o.y, v0.y), tz = f_sub(o.z, v0.z);
    u = f_mul(f_fma(tz, pz, f_fma(ty, py, f_mul(tx, px))), inv);
    if (!(u >= 0.0f) || u > 1.0f) return false;
    const float qx = f_fma(ty, e1.z, -f_mul(tz, e1.y));
    const float qy = f_fma(tz, e1.x, -f_mul(tx, e1.z));
    const float qz = f_fma(tx, e1.y, -f_mul(ty, e1.x));
    v = f_mul(f_fma(d.z, qz, f_fma(d.y, qy, f_mul(d.x, qx))), inv);
    if (!(v >= 0.0f) || f_add(u, v) > 1.0f) return false;
    t = f_mul(f_fma(e2.z, qz, f_fma(e2.y, qy, f_mul(e2.x, qx))), inv);
    return true;
}

VLB_HD float safe_inv(float d) {
    const float eps = 1e-30f;
    if (fabsf(d) < eps) d = copysignf(eps, d);
    return 1.0f / d;
}

// object->world point transform, fixed operation order (oracle: xform_point)
VLB_HD Vec3 xform_point(const float* m, Vec3 p) {
    Vec3 r;
    r.x = f_fma(m[2], p.z, f_fma(m[1], p.y, f_fma(m[0], p.x, m[3])));
    r.y = f_fma(m[6], p.z, f_fma(m[5], p.y, f_fma(m[4], p.x, m[7])));
    r.z = f_fma(m[10], p.z, f_fma(m[9], p.y, f_fma(m[8], p.x, m[11])));
    return r;
}
// vec3(nrm * gl_WorldToObjectEXT) (shaders/env_map.rchit:68), nm = inverse 3x3 row-major
VLB_HD Vec3 xform_normal(const float* nm, Vec3 n) {
    Vec3 r;
    r.x = f_fma(nm[6], n.z, f_fma(nm[3], n.y, f_mul(nm[0], n.x)));
    r.y = f_fma(nm[7], n.z, f_fma(nm[4], n.y, f_mul(nm[1], n.x)));
    r.z = f_fma(nm[8], n.z, f_fma(nm[5], n.y, f_mul(nm[2], n.x)));
    return r;
}

}  // namespace vlb

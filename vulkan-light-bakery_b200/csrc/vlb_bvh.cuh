// vlb_bvh.cuh — software LBVH: data layout in HBM, per-thread build steps and traversal.
//
// Replaces the driver-built BLAS/TLAS of the reference (src/scene_manager.cpp:339-443) on a GPU
// without RT cores: Morton codes -> radix sort -> Karras-2012 hierarchy -> bottom-up AABB refit ->
// emission of 64-byte two-child traversal nodes read as four 16-byte loads.
//
// HBM layout (all float4 arrays, 16-byte aligned):
//   tri[3*j+0..2]   j = position in Morton order:  (v0.xyz, bits(flat id)), (e1.xyz, 0), (e2.xyz, 0)
//   node[4*i+0..3]  i = Karras internal node id:   (c0.lo.x, c0.hi.x, c0.lo.y, c0.hi.y)
//                                                  (c1.lo.x, c1.hi.x, c1.lo.y, c1.hi.y)
//                                                  (c0.lo.z, c0.hi.z, c1.lo.z, c1.hi.z)
//                                                  bits(ref0, ref1, first, last)
//   ref >= 0 : internal node index.  ref < 0 : leaf, ~ref = (first_tri << 3) | (count - 1).
//
// Per-thread bodies are `__host__ __device__` (see vlb_math.cuh) so tests/emu can execute them
// serially on the CPU; the kernels in bvh_build.cu / bake.cu are thin wrappers.
#pragma once

#include <string.h>

#include "vlb_math.cuh"

namespace vlb {

constexpr int kMaxLeaf = 8;
constexpr int kStackSize = 64;
// Culling slack: a node is skipped only if its entry distance exceeds best_t * kCullSlack, so
// that two triangles whose computed t differ by rounding are both reached and the
// (t, flat id) tie-break decides exactly as in the brute-force intersector.
constexpr float kCullSlack = 1.0001f;

struct BvhView {
    const float4* nodes;
    const float4* tris;
    uint32_t n_tris;
};

VLB_HD float i2f(int v) {
#ifdef __CUDA_ARCH__
    return __int_as_float(v);
#else
    float f; memcpy(&f, &v, 4); return f;
#endif
}
VLB_HD int f2i(float f) {
#ifdef __CUDA_ARCH__
    return __float_as_int(f);
#else
    int v; memcpy(&v, &f, 4); return v;
#endif
}

VLB_HD int clz32(uint32_t x) {
#ifdef __CUDA_ARCH__
    return __clz((int)x);
#else
    return x ? __builtin_clz(x) : 32;
#endif
}
VLB_HD int clz64(uint64_t x) {
#ifdef __CUDA_ARCH__
    return __clzll((long long)x);
#else
    return x ? __builtin_clzll(x) : 64;
#endif
}

// 21 bits per axis -> 63-bit Morton code
VLB_HD uint64_t expand21(uint32_t v) {
    uint64_t x = v & 0x1fffffull;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}
VLB_HD uint64_t morton63(float nx, float ny, float nz) {  // inputs in [0,1]
    const float s = 2097152.0f;                           // 2^21
    uint32_t ix = (uint32_t)fminf(fmaxf(nx * s, 0.0f), s - 1.0f);
    uint32_t iy = (uint32_t)fminf(fmaxf(ny * s, 0.0f), s - 1.0f);
    uint32_t iz = (uint32_t)fminf(fmaxf(nz * s, 0.0f), s - 1.0f);
    return (expand21(ix) << 2) | (expand21(iy) << 1) | expand21(iz);
}

// Karras 2012: common-prefix length of sorted keys i and j, ties broken by index.
VLB_HD int delta(const uint64_t* keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    const uint64_t x = keys[i] ^ keys[j];
    if (x == 0) return 64 + clz32((uint32_t)i ^ (uint32_t)j);
    return clz64(x);
}

// One internal node i in [0, n-2]. Children are encoded: >= 0 internal index, < 0 : ~leaf index.
VLB_HD void karras_node(const uint64_t* keys, int n, int i, int* left, int* right, int* first,
                        int* last, int* parent_internal, int* parent_leaf) {
    const int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) < 0 ? -1 : 1;
    const int dmin = delta(keys, n, i, i - d);
    int lmax = 2;
    while (delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = delta(keys, n, i, j);
    int s = 0;
    int t = l;
    do {
        t = (t + 1) >> 1;
        if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    const int gamma = i + s * d + (d < 0 ? -1 : 0);
    const int lo = i < j ? i : j, hi = i < j ? j : i;
    first[i] = lo;
    last[i] = hi;
    if (lo == gamma) { left[i] = ~gamma; parent_leaf[gamma] = i; }
    else             { left[i] = gamma;  parent_internal[gamma] = i; }
    if (hi == gamma + 1) { right[i] = ~(gamma + 1); parent_leaf[gamma + 1] = i; }
    else                 { right[i] = gamma + 1;    parent_internal[gamma + 1] = i; }
}

VLB_HD void tri_aabb(const float4 v0, const float4 e1, const float4 e2, float4* lo, float4* hi) {
    const float ax = v0.x, ay = v0.y, az = v0.z;
    const float bx = f_add(v0.x, e1.x), by = f_add(v0.y, e1.y), bz = f_add(v0.z, e1.z);
    const float cx = f_add(v0.x, e2.x), cy = f_add(v0.y, e2.y), cz = f_add(v0.z, e2.z);
    *lo = make_float4(fminf(ax, fminf(bx, cx)), fminf(ay, fminf(by, cy)), fminf(az, fminf(bz, cz)), 0.f);
    *hi = make_float4(fmaxf(ax, fmaxf(bx, cx)), fmaxf(ay, fmaxf(by, cy)), fmaxf(az, fmaxf(bz, cz)), 0.f);
}

// Conservative padding of an AABB before it is stored in a traversal node: the slab test uses
// fma(lo, idir, -o*idir), whose error is ~|coord| * 2^-23 in space.
VLB_HD void pad_box(float4* lo, float4* hi, float abs_pad) {
    const float r = 4e-6f;
    lo->x -= abs_pad + r * fabsf(lo->x); lo->y -= abs_pad + r * fabsf(lo->y); lo->z -= abs_pad + r * fabsf(lo->z);
    hi->x += abs_pad + r * fabsf(hi->x); hi->y += abs_pad + r * fabsf(hi->y); hi->z += abs_pad + r * fabsf(hi->z);
}

VLB_HD int leaf_ref(int first_tri, int count) { return ~((first_tri << 3) | (count - 1)); }

// Emit traversal node i from the Karras arrays and the refitted boxes. Subtrees of at most
// `max_leaf` triangles become one leaf (their triangles are contiguous in Morton order).
VLB_HD void emit_node(int i, const int* left, const int* right, const int* first, const int* last,
                      const float4* ibox, const float4* lbox, int max_leaf, float abs_pad,
                      float4* nodes) {
    int refs[2];
    float4 lo[2], hi[2];
    const int ch[2] = {left[i], right[i]};
    for (int c = 0; c < 2; ++c) {
        if (ch[c] < 0) {
            const int leaf = ~ch[c];
            refs[c] = leaf_ref(leaf, 1);
            lo[c] = lbox[2 * leaf]; hi[c] = lbox[2 * leaf + 1];
        } else {
            const int k = ch[c];
            const int cnt = last[k] - first[k] + 1;
            refs[c] = cnt <= max_leaf ? leaf_ref(first[k], cnt) : k;
            lo[c] = ibox[2 * k]; hi[c] = ibox[2 * k + 1];
        }
        pad_box(&lo[c], &hi[c], abs_pad);
    }
    nodes[4 * i + 0] = make_float4(lo[0].x, hi[0].x, lo[0].y, hi[0].y);
    nodes[4 * i + 1] = make_float4(lo[1].x, hi[1].x, lo[1].y, hi[1].y);
    nodes[4 * i + 2] = make_float4(lo[0].z, hi[0].z, lo[1].z, hi[1].z);
    float4 meta;
    meta.x = i2f(refs[0]);
    meta.y = i2f(refs[1]);
    meta.z = i2f(first[i]);
    meta.w = i2f(last[i]);
    nodes[4 * i + 3] = meta;
}


struct HitRec {
    int id;       // flat triangle id, -1 = miss
    float t, u, v;
};

struct TraceCounters {
    uint32_t nodes, tris;
};

VLB_HD float4 ld4(const float4* p) {
#ifdef __CUDA_ARCH__
    return __ldg(p);
#else
    return *p;
#endif
}

// Two-child slab test. Returns entry distances (tn0, tn1) and hit flags.
VLB_HD void test_children(const float4 n0xy, const float4 n1xy, const float4 nz, Vec3 idir, Vec3 ood,
                          float tmin, float tcull, float& tn0, float& tn1, bool& h0, bool& h1) {
    const float c0lox = f_fma(n0xy.x, idir.x, -ood.x), c0hix = f_fma(n0xy.y, idir.x, -ood.x);
    const float c0loy = f_fma(n0xy.z, idir.y, -ood.y), c0hiy = f_fma(n0xy.w, idir.y, -ood.y);
    const float c0loz = f_fma(nz.x, idir.z, -ood.z), c0hiz = f_fma(nz.y, idir.z, -ood.z);
    const float c1lox = f_fma(n1xy.x, idir.x, -ood.x), c1hix = f_fma(n1xy.y, idir.x, -ood.x);
    const float c1loy = f_fma(n1xy.z, idir.y, -ood.y), c1hiy = f_fma(n1xy.w, idir.y, -ood.y);
    const float c1loz = f_fma(nz.z, idir.z, -ood.z), c1hiz = f_fma(nz.w, idir.z, -ood.z);
    tn0 = fmaxf(fmaxf(fminf(c0lox, c0hix), fminf(c0loy, c0hiy)), fmaxf(fminf(c0loz, c0hiz), tmin));
    const float tf0 = fminf(fminf(fmaxf(c0lox, c0hix), fmaxf(c0loy, c0hiy)), fminf(fmaxf(c0loz, c0hiz), tcull));
    tn1 = fmaxf(fmaxf(fminf(c1lox, c1hix), fminf(c1loy, c1hiy)), fmaxf(fminf(c1loz, c1hiz), tmin));
    const float tf1 = fminf(fminf(fmaxf(c1lox, c1hix), fmaxf(c1loy, c1hiy)), fminf(fmaxf(c1loz, c1hiz), tcull));
    h0 = tn0 <= tf0;
    h1 = tn1 <= tf1;
}

// Closest hit with tmin < t < tmax; ties: smaller t, then smaller flat id.
template <bool COUNT>
VLB_HD HitRec trace_closest(const BvhView& b, Vec3 o, Vec3 d, float tmin, float tmax, TraceCounters* cnt) {
    HitRec best;
    best.id = -1; best.t = tmax; best.u = 0.f; best.v = 0.f;
    if (b.n_tris == 0) return best;
    const Vec3 idir = mk3(safe_inv(d.x), safe_inv(d.y), safe_inv(d.z));
    const Vec3 ood = mk3(o.x * idir.x, o.y * idir.y, o.z * idir.z);
    int stack[kStackSize];
    int sp = 0;
    int cur = 0;
    float tcull = tmax;
    for (;;) {
        if (cur >= 0) {
            const float4 n0xy = ld4(b.nodes + 4 * cur + 0);
            const float4 n1xy = ld4(b.nodes + 4 * cur + 1);
            const float4 nz = ld4(b.nodes + 4 * cur + 2);
            const float4 meta = ld4(b.nodes + 4 * cur + 3);
            if (COUNT) cnt->nodes++;
            float tn0, tn1; bool h0, h1;
            test_children(n0xy, n1xy, nz, idir, ood, tmin, tcull, tn0, tn1, h0, h1);
            const int r0 = f2i(meta.x), r1 = f2i(meta.y);
            if (h0 && h1) {
                const bool swap = tn1 < tn0;
                cur = swap ? r1 : r0;
                if (sp < kStackSize) stack[sp++] = swap ? r0 : r1;
                continue;
            } else if (h0) { cur = r0; continue; }
            else if (h1) { cur = r1; continue; }
        } else {
            const int x = ~cur;
            const int first = x >> 3, count = (x & 7) + 1;
            for (int k = 0; k < count; ++k) {
                const float4 v0 = ld4(b.tris + 3 * (first + k) + 0);
                const float4 e1 = ld4(b.tris + 3 * (first + k) + 1);
                const float4 e2 = ld4(b.tris + 3 * (first + k) + 2);
                if (COUNT) cnt->tris++;
                float t, u, v;
                if (intersect_tri(v0, e1, e2, o, d, t, u, v) && t > tmin) {
                    const int id = f2i(v0.w);
                    if (t < best.t || (t == best.t && best.id >= 0 && id < best.id)) {
                        best.id = id; best.t = t; best.u = u; best.v = v;
                        tcull = t * kCullSlack;
                    }
                }
            }
        }
        if (sp == 0) break;
        cur = stack[--sp];
    }
    return best;
}

// Any hit with tmin < t < tmax (shadow rays; gl_RayFlagsTerminateOnFirstHitEXT).
template <bool COUNT>
VLB_HD bool trace_any(const BvhView& b, Vec3 o, Vec3 d, float tmin, float tmax, TraceCounters* cnt, HitRec* out) {
    if (b.n_tris == 0) return false;
    const Vec3 idir = mk3(safe_inv(d.x), safe_inv(d.y), safe_inv(d.z));
    const Vec3 ood = mk3(o.x * idir.x, o.y * idir.y, o.z * idir.z);
    int stack[kStackSize];
    int sp = 0;
    int cur = 0;
    for (;;) {
        if (cur >= 0) {
            const float4 n0xy = ld4(b.nodes + 4 * cur + 0);
            const float4 n1xy = ld4(b.nodes + 4 * cur + 1);
            const float4 nz = ld4(b.nodes + 4 * cur + 2);
            const float4 meta = ld4(b.nodes + 4 * cur + 3);
            if (COUNT) cnt->nodes++;
            float tn0, tn1; bool h0, h1;
            test_children(n0xy, n1xy, nz, idir, ood, tmin, tmax, tn0, tn1, h0, h1);
            const int r0 = f2i(meta.x), r1 = f2i(meta.y);
            if (h0 && h1) {
                cur = r0;
                if (sp < kStackSize) stack[sp++] = r1;
                continue;
            } else if (h0) { cur = r0; continue; }
            else if (h1) { cur = r1; continue; }
        } else {
            const int x = ~cur;
            const int first = x >> 3, count = (x & 7) + 1;
            for (int k = 0; k < count; ++k) {
                const float4 v0 = ld4(b.tris + 3 * (first + k) + 0);
                const float4 e1 = ld4(b.tris + 3 * (first + k) + 1);
                const float4 e2 = ld4(b.tris + 3 * (first + k) + 2);
                if (COUNT) cnt->tris++;
                float t, u, v;
                if (intersect_tri(v0, e1, e2, o, d, t, u, v) && t > tmin && t < tmax) {
                    if (out) { out->id = f2i(v0.w); out->t = t; out->u = u; out->v = v; }
                    return true;
                }
            }
        }
        if (sp == 0) break;
        cur = stack[--sp];
    }
    return false;
}

}  // namespace vlb

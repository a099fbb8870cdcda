// vlb_bvh.cuh — software LBVH: data layout in HBM, per-thread build steps and traversal.
//
// Replaces the driver-built BLAS/TLAS of the reference (src/scene_manager.cpp:339-443) on a GPU
// without RT cores: Morton codes -> radix sort -> Karras-2012 binary hierarchy -> bottom-up AABB
// refit -> collapse of every second level into 4-wide traversal nodes of 128 bytes, read as seven
// 16-byte loads. Four children per step halve the number of dependent memory round trips per ray,
// which is what bounds traversal on L2-resident data.
//
// HBM layout (all float4 arrays, 16-byte aligned):
//   tri[3*j+0..2]   j = position in Morton order:  (v0.xyz, bits(flat id)), (e1.xyz, 0), (e2.xyz, 0)
//   node[Q*i+0..Q-1] i = Karras id of a binary node at EVEN depth (odd-depth nodes are absorbed by
//                   their parent; their slots stay unused). Q = kNodeQuads.
//                   fp32 format (Q = 7 or 8):
//                   [0] lo.x of children 0..3   [1] hi.x   [2] lo.y   [3] hi.y   [4] lo.z   [5] hi.z
//                   [6] bits(ref of children 0..3)         [7] (Q = 8 only) pads the node to a 128-byte line
//                   8-bit format (Q = 4, VLB_NODE_Q8): plane = origin + byte * step, step = 2^k per axis
//                   [0] origin.x, origin.y, origin.z, step.x * 2^15
//                   [1] step.y * 2^15, step.z * 2^15, lo.x bytes of children 0..3, hi.x bytes
//                   [2] lo.y bytes, hi.y bytes, lo.z bytes, hi.z bytes        [3] bits(ref of children 0..3)
//   ref >= 0 : node index.  ref < 0 : leaf, ~ref = (first_tri << 3) | (count - 1).
//   kNoChild (0x80000000) marks an empty slot; its box is (+inf, -inf) and can never be hit.
//
// Per-thread bodies are `__host__ __device__` (see vlb_math.cuh) so tests/emu can execute them
// serially on the CPU; the kernels in bvh_build.cu / bake.cu are thin wrappers.
#pragma once

#include <string.h>

#include "vlb_math.cuh"

namespace vlb {

constexpr int kMaxLeaf = 8;
constexpr int kStackSize = 96;              // up to three pushes per 4-wide node
// Node formats (compile-time):
//   VLB_NODE_Q8 = 0  (default) fp32 planes, VLB_NODE_QUADS float4s per node. 7 = packed 112-byte nodes: the
//                    same field of different nodes then falls into different L1 data banks, measured 4 %
//                    faster on B200 than 8 (one 128-byte line per node).
//   VLB_NODE_Q8 = 1  64-byte nodes (4 float4): child planes quantised to 8 bits in the node's own frame
//                    (origin + power-of-two step per axis), always conservatively (the quantised box contains
//                    the padded fp32 box; hit ids stay bit-exact, +1.6 % node visits). 4 instead of 7
//                    sixteen-byte loads per node step take the L1 data pipe from 87 % to 53 % busy, but the
//                    decode moves the bound to the ALU pipe (55 % -> 72 %): 5.7 % slower on C3, so it is a
//                    build option for scenes whose fp32 nodes do not fit L2
//                    (profiles/r01_bake_kernel_ab_experiments.log).
#ifndef VLB_NODE_Q8
#define VLB_NODE_Q8 0
#endif
#ifndef VLB_NODE_QUADS
#define VLB_NODE_QUADS 7
#endif
// VLB_NODE_ORDER1D = 1 (8-bit nodes only): the children of a node are stored sorted along the axis on which their
// centroids spread most, the axis is kept in the node, and a closest-hit ray visits the hit children in slot order
// (ascending or descending by the sign of its direction on that axis) instead of sorting them by entry distance: the
// 25 ALU-pipe instructions of the sort network pay for the 8-bit decode. Traversal order never changes a result
// (culling uses the hit distance with kCullSlack, ties go by flat id), only the number of nodes visited.
#ifndef VLB_NODE_ORDER1D
#define VLB_NODE_ORDER1D 0
#endif
// VLB_BVH8 = 1: 8-wide nodes of 96 bytes (6 float4) with 8-bit planes, the layout of a compressed wide BVH
// (Ylitie, Karras, Laine 2017) kept with explicit child refs so that the Morton-ordered triangle array stays as it is:
//   [0] origin.xyz, bits(ex | ey << 8 | ez << 16): plane = origin + byte * 2^(e - 127 - 15)
//   [1] lo.x bytes of children 0..3, of 4..7, hi.x bytes of 0..3, of 4..7     [2] the same for y     [3] for z
//   [4] refs of children 0..3      [5] refs of children 4..7   (NOT read by the node step: only the ref of the child
//                                                                that is entered is fetched, one 4-byte load)
// Children sit in octant-ordered slots (slot s lies towards ((s&1)?+:-, (s&2)?+:-, (s&4)?+:-) of the node's centre), so
// a ray visits the hit slots in the order flip ^ 0, flip ^ 1, ..., flip ^ 7 (flip = its three sign bits): no distance
// sort. The stack holds one (node, remaining hit mask) group per node, not one entry per child.
#ifndef VLB_BVH8
#define VLB_BVH8 0
#endif
constexpr bool kBvh8 = VLB_BVH8 != 0;
constexpr bool kNodeQ8 = VLB_NODE_Q8 != 0;
constexpr bool kOrder1D = VLB_NODE_Q8 != 0 && VLB_NODE_ORDER1D != 0;
constexpr int kNodeQuads = kBvh8 ? 6 : (kNodeQ8 ? 4 : VLB_NODE_QUADS);  // float4s per traversal node
constexpr int kRefQuad = kBvh8 ? 4 : (kNodeQ8 ? 3 : 6);                 // first float4 of the child refs
constexpr int kWide = kBvh8 ? 8 : 4;                                    // children per node
constexpr int kNoChild = (int)0x80000000;   // empty child slot / "no node": never a valid leaf ref (n_tris < 2^28)
// Culling slack: a node is skipped only if its entry distance exceeds best_t * kCullSlack, so
// that two triangles whose computed t differ by rounding are both reached and the
// (t, flat id) tie-break decides exactly as in the brute-force intersector.
constexpr float kCullSlack = 1.0001f;

struct BvhView {
    const float4* nodes;
    const float4* tris;
    uint32_t n_tris;
    unsigned int* overflow;   // set to 1 if a traversal stack overflowed (device pointer, may be NULL)
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

// Depth of binary internal node i (root = 0), by walking the parent chain.
VLB_HD int node_depth(const int* parent_internal, int i) {
    int d = 0;
    for (int k = parent_internal[i]; k >= 0; k = parent_internal[k]) ++d;
    return d;
}

// One child slot of a 4-wide node from a binary child reference `c` (>= 0 internal Karras id,
// < 0 : ~leaf index). Subtrees of at most `max_leaf` triangles become one leaf (their triangles are
// contiguous in Morton order). Returns true if `c` is a large internal node (to be expanded or
// referenced).
VLB_HD bool classify_child(int c, const int* first, const int* last, const float4* ibox, const float4* lbox,
                           int max_leaf, int* ref, float4* lo, float4* hi) {
    if (c < 0) {
        const int leaf = ~c;
        *ref = leaf_ref(leaf, 1);
        *lo = lbox[2 * leaf]; *hi = lbox[2 * leaf + 1];
        return false;
    }
    const int cnt = last[c] - first[c] + 1;
    *lo = ibox[2 * c]; *hi = ibox[2 * c + 1];
    if (cnt <= max_leaf) { *ref = leaf_ref(first[c], cnt); return false; }
    *ref = c;
    return true;
}

// Smallest power of two >= x (x > 0, finite).
VLB_HD float pow2_ceil(float x) {
    int e;
    const float m = frexpf(x, &e);          // x = m * 2^e, m in [0.5, 1)
    return ldexpf(1.0f, m == 0.5f ? e - 1 : e);
}

// 8-bit frame of one axis of a node: origin and power-of-two step such that every valid child's (padded)
// [lo, hi] maps into bytes 0..255 with room for the rounding margins of quant_lo / quant_hi.
VLB_HD void quant_frame(const float* lo, const float* hi, int n, float* origin, float* step) {
    float mn = lo[0], mx = hi[0];
    for (int k = 1; k < n; ++k) { mn = fminf(mn, lo[k]); mx = fmaxf(mx, hi[k]); }
    const float mag = fmaxf(fabsf(mn), fabsf(mx));
    // the step never goes below 4 ulp of the coordinates (finer planes are not representable in fp32 anyway)
    const float s = pow2_ceil(fmaxf(fmaxf((mx - mn) * (1.0f / 250.0f), mag * 4.8e-7f), 1e-30f));
    float o = mn - 0.125f * s;
    while ((double)o > (double)mn - 0.0625 * (double)s) o = nextafterf(o, -INFINITY);
    *origin = o; *step = s;
}
// Largest byte whose plane origin + byte * step lies at least 1/64 step below `v` (exact in double), and the
// mirror image for upper planes. The 1/64 step covers the traversal's extra rounding (< 1/256 step, bvh4_step).
VLB_HD uint32_t quant_lo(float v, float o, float s) {
    const double x = floor(((double)v - (double)o) / (double)s - 1.0 / 64.0);
    return x < 0.0 ? 0u : (x > 255.0 ? 255u : (uint32_t)x);
}
VLB_HD uint32_t quant_hi(float v, float o, float s) {
    const double x = ceil(((double)v - (double)o) / (double)s + 1.0 / 64.0);
    return x < 0.0 ? 0u : (x > 255.0 ? 255u : (uint32_t)x);
}

// Writes one traversal node: `n` valid children (padded boxes lo/hi, refs), the remaining slots empty.
VLB_HD void store_node4(float4* q, const int* refs_in, const float4* lo_in, const float4* hi_in, int n) {
    int refs[4]; float4 lo[4], hi[4];
    for (int k = 0; k < 4; ++k) { refs[k] = k < n ? refs_in[k] : kNoChild; if (k < n) { lo[k] = lo_in[k]; hi[k] = hi_in[k]; } }
    int order_axis = 0;
    if (kOrder1D && n > 1) {
        // axis on which the child centroids spread most; children sorted ascending along it (insertion sort of <= 4)
        float cmin[3] = {INFINITY, INFINITY, INFINITY}, cmax[3] = {-INFINITY, -INFINITY, -INFINITY};
        for (int k = 0; k < n; ++k) {
            const float c[3] = {lo[k].x + hi[k].x, lo[k].y + hi[k].y, lo[k].z + hi[k].z};
            for (int a = 0; a < 3; ++a) { cmin[a] = fminf(cmin[a], c[a]); cmax[a] = fmaxf(cmax[a], c[a]); }
        }
        const float e[3] = {cmax[0] - cmin[0], cmax[1] - cmin[1], cmax[2] - cmin[2]};
        order_axis = e[0] >= e[1] ? (e[0] >= e[2] ? 0 : 2) : (e[1] >= e[2] ? 1 : 2);
        for (int i = 1; i < n; ++i)
            for (int j = i; j > 0; --j) {
                const float a = order_axis == 0 ? lo[j - 1].x + hi[j - 1].x : (order_axis == 1 ? lo[j - 1].y + hi[j - 1].y : lo[j - 1].z + hi[j - 1].z);
                const float b = order_axis == 0 ? lo[j].x + hi[j].x : (order_axis == 1 ? lo[j].y + hi[j].y : lo[j].z + hi[j].z);
                if (!(b < a)) break;
                const float4 tl = lo[j], th = hi[j]; const int tr = refs[j];
                lo[j] = lo[j - 1]; hi[j] = hi[j - 1]; refs[j] = refs[j - 1];
                lo[j - 1] = tl; hi[j - 1] = th; refs[j - 1] = tr;
            }
    }
    if (kNodeQ8) {
        float o[3], st[3];
        uint32_t lob[3] = {0, 0, 0}, hib[3] = {0, 0, 0};
        for (int a = 0; a < 3; ++a) {
            float l[4], h[4];
            for (int k = 0; k < n; ++k) {
                l[k] = a == 0 ? lo[k].x : (a == 1 ? lo[k].y : lo[k].z);
                h[k] = a == 0 ? hi[k].x : (a == 1 ? hi[k].y : hi[k].z);
            }
            quant_frame(l, h, n, &o[a], &st[a]);
            for (int k = 0; k < 4; ++k) {
                // empty slots: inverted (255, 0); bvh4_step also rejects them by their kNoChild ref
                const uint32_t bl = k < n ? quant_lo(l[k], o[a], st[a]) : 255u;
                const uint32_t bh = k < n ? quant_hi(h[k], o[a], st[a]) : 0u;
                lob[a] |= bl << (8 * k); hib[a] |= bh << (8 * k);
            }
        }
        // the ordering axis rides in the two lowest mantissa bits of step.x * 2^15 (a power of two: they are free; the
        // traversal multiplies with them in place, 3 ulp of the scale against a margin of 1/64 step)
        q[0] = make_float4(o[0], o[1], o[2], kOrder1D ? i2f(f2i(st[0] * 32768.0f) | order_axis) : st[0] * 32768.0f);
        q[1] = make_float4(st[1] * 32768.0f, st[2] * 32768.0f, i2f((int)lob[0]), i2f((int)hib[0]));
        q[2] = make_float4(i2f((int)lob[1]), i2f((int)hib[1]), i2f((int)lob[2]), i2f((int)hib[2]));
        q[3] = make_float4(i2f(refs[0]), i2f(refs[1]), i2f(refs[2]), i2f(refs[3]));
        return;
    }
    const float inf = INFINITY;
    float4 l[4], h[4];
    for (int k = 0; k < 4; ++k) {
        l[k] = k < n ? lo[k] : make_float4(inf, inf, inf, 0.f);
        h[k] = k < n ? hi[k] : make_float4(-inf, -inf, -inf, 0.f);
    }
    q[0] = make_float4(l[0].x, l[1].x, l[2].x, l[3].x);
    q[1] = make_float4(h[0].x, h[1].x, h[2].x, h[3].x);
    q[2] = make_float4(l[0].y, l[1].y, l[2].y, l[3].y);
    q[3] = make_float4(h[0].y, h[1].y, h[2].y, h[3].y);
    q[4] = make_float4(l[0].z, l[1].z, l[2].z, l[3].z);
    q[5] = make_float4(h[0].z, h[1].z, h[2].z, h[3].z);
    q[6] = make_float4(i2f(refs[0]), i2f(refs[1]), i2f(refs[2]), i2f(refs[3]));
    if (kNodeQuads > 7) q[7] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// Half surface area of an AABB (the SAH weight of a node).
VLB_HD float box_half_area(const float4 lo, const float4 hi) {
    const float dx = hi.x - lo.x, dy = hi.y - lo.y, dz = hi.z - lo.z;
    return dx * dy + dy * dz + dz * dx;
}

// Emit the 4-wide traversal node rooted at binary node i (i spans more than max_leaf triangles, or is the
// root). Its children are chosen greedily by surface area, as wide-BVH builders collapse a binary BVH: start
// from i's two children and keep opening the largest child that is still a big internal node until there are
// four. An LBVH is unbalanced, so taking the grandchildren level by level leaves many slots empty (2.6 children
// per node on the atrium); the greedy collapse fills them and opens the boxes a ray is most likely to enter.
// Children that are big internal nodes become roots of the next wide nodes: they are appended to `next`
// (atomic counter `n_next`; both may be NULL when the caller walks the tree itself and reads the refs back).
VLB_HD void emit_node4(int i, const int* left, const int* right, const int* first, const int* last,
                       const float4* ibox, const float4* lbox, int max_leaf, float abs_pad, float4* nodes,
                       int* next, unsigned int* n_next) {
    int c[4] = {left[i], right[i], 0, 0};
    int n = 2;
    while (n < 4) {
        int best = -1;
        float best_area = -1.0f;
        for (int k = 0; k < n; ++k) {
            if (c[k] < 0 || last[c[k]] - first[c[k]] + 1 <= max_leaf) continue;       // a leaf stays closed
            const float area = box_half_area(ibox[2 * c[k]], ibox[2 * c[k] + 1]);
            if (area > best_area) { best_area = area; best = k; }
        }
        if (best < 0) break;
        const int open = c[best];
        c[best] = left[open];
        c[n++] = right[open];
    }
    int refs[4] = {kNoChild, kNoChild, kNoChild, kNoChild};
    float4 lo[4], hi[4];
    for (int k = 0; k < n; ++k) {
        if (classify_child(c[k], first, last, ibox, lbox, max_leaf, &refs[k], &lo[k], &hi[k]) && next) {
#ifdef __CUDA_ARCH__
            next[atomicAdd(n_next, 1u)] = c[k];
#else
            next[(*n_next)++] = c[k];
#endif
        }
        pad_box(&lo[k], &hi[k], abs_pad);
    }
    store_node4(nodes + (size_t)kNodeQuads * i, refs, lo, hi, n);
}

// Single-triangle scene: one node with one leaf child.
VLB_HD void emit_single4(const float4* lbox, float abs_pad, float4* nodes) {
    float4 lo[4], hi[4];
    lo[0] = lbox[0]; hi[0] = lbox[1];
    pad_box(&lo[0], &hi[0], abs_pad);
    const int refs[4] = {leaf_ref(0, 1), kNoChild, kNoChild, kNoChild};
    store_node4(nodes, refs, lo, hi, 1);
}

// ---- 8-wide nodes (VLB_BVH8) ----
// Writes one 8-wide node: `n` valid children (padded boxes, refs) assigned to octant slots, the rest empty.
VLB_HD void store_node8(float4* q, const int* refs_in, const float4* lo_in, const float4* hi_in, int n) {
    // slot assignment: greedy maximum of dot(child centroid - node centre, diagonal of the slot)
    float cen[8][3], mid[3] = {0.f, 0.f, 0.f};
    for (int k = 0; k < n; ++k) {
        cen[k][0] = 0.5f * (lo_in[k].x + hi_in[k].x); cen[k][1] = 0.5f * (lo_in[k].y + hi_in[k].y); cen[k][2] = 0.5f * (lo_in[k].z + hi_in[k].z);
        for (int a = 0; a < 3; ++a) mid[a] += cen[k][a] / (float)n;
    }
    int slot_of[8], child_in[8];
    for (int k = 0; k < 8; ++k) { slot_of[k] = -1; child_in[k] = -1; }
    for (int it = 0; it < n; ++it) {
        float best = -INFINITY; int bc = -1, bs = -1;
        for (int c = 0; c < n; ++c) {
            if (slot_of[c] >= 0) continue;
            for (int sl = 0; sl < 8; ++sl) {
                if (child_in[sl] >= 0) continue;
                const float sc = ((sl & 1) ? 1.f : -1.f) * (cen[c][0] - mid[0]) + ((sl & 2) ? 1.f : -1.f) * (cen[c][1] - mid[1]) +
                                 ((sl & 4) ? 1.f : -1.f) * (cen[c][2] - mid[2]);
                if (sc > best) { best = sc; bc = c; bs = sl; }
            }
        }
        slot_of[bc] = bs; child_in[bs] = bc;
    }
    float o[3], st[3];
    uint32_t lob[3][2] = {{0, 0}, {0, 0}, {0, 0}}, hib[3][2] = {{0, 0}, {0, 0}, {0, 0}};
    uint32_t ebits = 0;
    for (int a = 0; a < 3; ++a) {
        float l[8], h[8];
        for (int k = 0; k < n; ++k) {
            l[k] = a == 0 ? lo_in[k].x : (a == 1 ? lo_in[k].y : lo_in[k].z);
            h[k] = a == 0 ? hi_in[k].x : (a == 1 ? hi_in[k].y : hi_in[k].z);
        }
        quant_frame(l, h, n, &o[a], &st[a]);
        ebits |= (((uint32_t)f2i(st[a] * 32768.0f) >> 23) & 0xffu) << (8 * a);      // step is a power of two: only its exponent is kept
        for (int sl = 0; sl < 8; ++sl) {
            const int c = child_in[sl];
            // empty slots: inverted (255, 0): near > far for either ray sign
            const uint32_t bl = c >= 0 ? quant_lo(l[c], o[a], st[a]) : 255u;
            const uint32_t bh = c >= 0 ? quant_hi(h[c], o[a], st[a]) : 0u;
            lob[a][sl >> 2] |= bl << (8 * (sl & 3)); hib[a][sl >> 2] |= bh << (8 * (sl & 3));
        }
    }
    q[0] = make_float4(o[0], o[1], o[2], i2f((int)ebits));
    for (int a = 0; a < 3; ++a)
        q[1 + a] = make_float4(i2f((int)lob[a][0]), i2f((int)lob[a][1]), i2f((int)hib[a][0]), i2f((int)hib[a][1]));
    int r[8];
    for (int sl = 0; sl < 8; ++sl) r[sl] = child_in[sl] >= 0 ? refs_in[child_in[sl]] : kNoChild;
    q[4] = make_float4(i2f(r[0]), i2f(r[1]), i2f(r[2]), i2f(r[3]));
    q[5] = make_float4(i2f(r[4]), i2f(r[5]), i2f(r[6]), i2f(r[7]));
}

// emit_node4's 8-wide twin: open the largest big child until there are eight.
VLB_HD void emit_node8(int i, const int* left, const int* right, const int* first, const int* last,
                       const float4* ibox, const float4* lbox, int max_leaf, float abs_pad, float4* nodes,
                       int* next, unsigned int* n_next) {
    int c[8] = {left[i], right[i], 0, 0, 0, 0, 0, 0};
    int n = 2;
    while (n < 8) {
        int best = -1;
        float best_area = -1.0f;
        for (int k = 0; k < n; ++k) {
            if (c[k] < 0 || last[c[k]] - first[c[k]] + 1 <= max_leaf) continue;
            const float area = box_half_area(ibox[2 * c[k]], ibox[2 * c[k] + 1]);
            if (area > best_area) { best_area = area; best = k; }
        }
        if (best < 0) break;
        const int open = c[best];
        c[best] = left[open];
        c[n++] = right[open];
    }
    int refs[8];
    float4 lo[8], hi[8];
    for (int k = 0; k < n; ++k) {
        if (classify_child(c[k], first, last, ibox, lbox, max_leaf, &refs[k], &lo[k], &hi[k]) && next) {
#ifdef __CUDA_ARCH__
            next[atomicAdd(n_next, 1u)] = c[k];
#else
            next[(*n_next)++] = c[k];
#endif
        }
        pad_box(&lo[k], &hi[k], abs_pad);
    }
    store_node8(nodes + (size_t)kNodeQuads * i, refs, lo, hi, n);
}

// The node emitters of the configured width.
VLB_HD void emit_wide_node(int i, const int* left, const int* right, const int* first, const int* last,
                           const float4* ibox, const float4* lbox, int max_leaf, float abs_pad, float4* nodes,
                           int* next, unsigned int* n_next) {
    if (kBvh8) emit_node8(i, left, right, first, last, ibox, lbox, max_leaf, abs_pad, nodes, next, n_next);
    else emit_node4(i, left, right, first, last, ibox, lbox, max_leaf, abs_pad, nodes, next, n_next);
}
VLB_HD void emit_wide_single(const float4* lbox, float abs_pad, float4* nodes) {
    if (kBvh8) {
        float4 lo[1] = {lbox[0]}, hi[1] = {lbox[1]};
        pad_box(&lo[0], &hi[0], abs_pad);
        const int refs[1] = {leaf_ref(0, 1)};
        store_node8(nodes, refs, lo, hi, 1);
    } else {
        emit_single4(lbox, abs_pad, nodes);
    }
}

struct HitRec {
    int id;       // flat triangle id, -1 = miss
    float t, u, v;
};

struct TraceCounters {
    uint32_t nodes, tris;
};

// A 16-byte load of data with little reuse inside an SM (skybox texels, per-triangle shading records): served through
// L2 without taking an L1 line away from the BVH nodes and triangles (VLB_L1_HINTS, A/B in DESIGN.md 4.2).
#ifndef VLB_L1_HINTS
#define VLB_L1_HINTS 0
#endif
VLB_HD float4 ld4_stream(const float4* p) {
#if defined(__CUDA_ARCH__) && VLB_L1_HINTS
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
#elif defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}

VLB_HD float4 ld4(const float4* p) {
#ifdef __CUDA_ARCH__
    return __ldg(p);
#else
    return *p;
#endif
}

// 1 + byte k of `w` * 2^-15: the byte dropped into bits 8..15 of 1.0f (one PRMT on the device).
VLB_HD float byte_to_unit(uint32_t w, int k) {
#ifdef __CUDA_ARCH__
    return __uint_as_float(__byte_perm(w, 0x3F800000u, 0x7604u | ((uint32_t)k << 4)));
#else
    return i2f((int)(0x3F800000u | (((w >> (8 * k)) & 0xffu) << 8)));
#endif
}

VLB_HD void order2(float& ta, int& ra, float& tb, int& rb) {
    const bool sw = tb < ta;
    const float t0 = sw ? tb : ta, t1 = sw ? ta : tb;
    const int r0 = sw ? rb : ra, r1 = sw ? ra : rb;
    ta = t0; tb = t1; ra = r0; rb = r1;
}

// three-input min / max: one FMNMX3 on sm_100
VLB_HD float max3(float a, float b, float c) {
#ifdef __CUDA_ARCH__
    float r; asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r;
#else
    return fmaxf(fmaxf(a, b), c);
#endif
}
VLB_HD float min3(float a, float b, float c) {
#ifdef __CUDA_ARCH__
    float r; asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r;
#else
    return fminf(fminf(a, b), c);
#endif
}

#ifndef VLB_FFMA2
#define VLB_FFMA2 1
#endif
VLB_HD float byte_to_unit(uint32_t w, int k);
// 8-bit node format: the four plane distances v_k * S + C of one plane word (v_k = 1 + byte k * 2^-15). On the device
// four PRMTs and two packed FFMA2, per-element IEEE, so bit-identical to the scalar host form.
VLB_HD void q8_planes(uint32_t w, float S, float C, float4& out) {
#if defined(__CUDA_ARCH__) && VLB_FFMA2
    unsigned long long a0, a1, ss, cc;
    asm("mov.b64 %0, {%1, %2};" : "=l"(a0) : "f"(byte_to_unit(w, 0)), "f"(byte_to_unit(w, 1)));
    asm("mov.b64 %0, {%1, %2};" : "=l"(a1) : "f"(byte_to_unit(w, 2)), "f"(byte_to_unit(w, 3)));
    asm("mov.b64 %0, {%1, %1};" : "=l"(ss) : "f"(S));
    asm("mov.b64 %0, {%1, %1};" : "=l"(cc) : "f"(C));
    asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a0) : "l"(ss), "l"(cc));
    asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a1) : "l"(ss), "l"(cc));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(out.x), "=f"(out.y) : "l"(a0));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(out.z), "=f"(out.w) : "l"(a1));
#else
    out.x = f_fma(byte_to_unit(w, 0), S, C); out.y = f_fma(byte_to_unit(w, 1), S, C);
    out.z = f_fma(byte_to_unit(w, 2), S, C); out.w = f_fma(byte_to_unit(w, 3), S, C);
#endif
}
#if defined(__CUDA_ARCH__)
// v = v * s - c on all four components with two packed fp32x2 FMAs (fma.rn.f32x2 -> FFMA2 on sm_100a)
__device__ __forceinline__ void fma2_planes(float4& v, float s, float c) {
    unsigned long long a0, a1, ss, cc;
    const float nc = -c;
    asm("mov.b64 %0, {%1, %2};" : "=l"(a0) : "f"(v.x), "f"(v.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(a1) : "f"(v.z), "f"(v.w));
    asm("mov.b64 %0, {%1, %1};" : "=l"(ss) : "f"(s));
    asm("mov.b64 %0, {%1, %1};" : "=l"(cc) : "f"(nc));
    asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a0) : "l"(ss), "l"(cc));
    asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a1) : "l"(ss), "l"(cc));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(v.x), "=f"(v.y) : "l"(a0));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(v.z), "=f"(v.w) : "l"(a1));
}
#endif

// Traversal stack of one ray as a per-thread array (local memory on the device). The bake kernel uses the
// warp-shared-memory stack of bake.cu instead; bvh4_step takes either.
// VLB_STACK_CULL = 1: every stack entry carries the entry distance of its box; a popped entry whose box starts beyond
// the current culling distance is dropped without a node step (closest-hit rays: after the first hit most pending far
// children are farther than the hit). Hit ids stay bit-exact: a triangle inside the box is hit no earlier than the box
// is entered, and tcull already carries kCullSlack.
#ifndef VLB_STACK_CULL
#define VLB_STACK_CULL 0
#endif
VLB_HD int ctz32(uint32_t x) {
#ifdef __CUDA_ARCH__
    return __ffs((int)x) - 1;
#else
    return __builtin_ctz(x);
#endif
}
// ref of child `slot` of 8-wide node `node`: the one 4-byte load a descent costs besides the node's four plane loads
VLB_HD int ref_of8(const BvhView& b, int node, int slot) {
    const int* r = reinterpret_cast<const int*>(b.nodes + (size_t)kNodeQuads * node + kRefQuad) + slot;
#ifdef __CUDA_ARCH__
    return __ldg(r);
#else
    return *r;
#endif
}

struct LocalStack {
    int a[kStackSize];
#if VLB_STACK_CULL
    float t[kStackSize];
#endif
#if VLB_BVH8
    unsigned m[kStackSize];     // remaining hit mask of the group (priority order)
    // 8-wide: enter the first child of node `cur` in priority order (mask bit p = slot p ^ flip was hit), keep the rest
    // of the group on the stack; with an empty mask continue with the newest group on the stack.
    VLB_HD int descend(const BvhView& b, int cur, unsigned mask, int flip) {
        if (mask == 0u) return pop_group(b, flip);
        const int p = ctz32(mask);
        mask &= mask - 1u;
        if (mask != 0u) {
            if (!room(1)) { if (b.overflow) *b.overflow = 1u; }
            else { a[sp] = cur; m[sp] = mask; ++sp; }
        }
        return ref_of8(b, cur, p ^ flip);
    }
    VLB_HD int pop_group(const BvhView& b, int flip) {
        if (sp == 0) return kNoChild;
        const int node = a[sp - 1];
        unsigned mask = m[sp - 1];
        const int p = ctz32(mask);
        mask &= mask - 1u;
        if (mask != 0u) m[sp - 1] = mask; else --sp;
        return ref_of8(b, node, p ^ flip);
    }
#endif
    int sp;
    VLB_HD void clear() { sp = 0; }
    VLB_HD bool empty() const { return sp == 0; }
    VLB_HD bool room(int n) const { return sp + n <= kStackSize; }
    VLB_HD void push(int v, float tn) {
#if VLB_STACK_CULL
        t[sp] = tn;
#endif
        a[sp++] = v;
    }
    // next pending node / leaf whose box starts within tcull, or kNoChild
    VLB_HD int pop(float tcull) {
#if VLB_STACK_CULL
        while (sp > 0) {
            --sp;
            if (t[sp] <= tcull) return a[sp];
        }
        return kNoChild;
#else
        (void)tcull;
        return sp > 0 ? a[--sp] : kNoChild;
#endif
    }
    // End of an ordered node step: r0..r3 = the hit children, nearest first, misses (kNoChild) last, t1..t3 the entry
    // distances of r1..r3. Returns the next node / leaf of the ray (r0, or the top of the stack when nothing was hit, or
    // kNoChild when the ray is done) and pushes the other hit children so that they pop near to far.
    VLB_HD int advance(int r0, int r1, int r2, int r3, float t1, float t2, float t3, float tcull, unsigned int* overflow) {
        if (r0 == kNoChild) return pop(tcull);
        if (r1 != kNoChild) {
            // a full stack never drops subtrees silently: *overflow is set (the caller reports VLB_ERR_UNSUPPORTED)
            if (!room(3)) { if (overflow) *overflow = 1u; return r0; }
            if (r3 != kNoChild) push(r3, t3);
            if (r2 != kNoChild) push(r2, t2);
            push(r1, t1);
        }
        return r0;
    }
    // End of an unsorted node step (VLB_NODE_ORDER1D): a0..a3 = the children in visiting order, kNoChild where the ray
    // misses. Returns the first hit one (or the top of the stack / kNoChild) and pushes the later ones so that they pop
    // in visiting order.
    VLB_HD int advance_unsorted(int a0, int a1, int a2, int a3, float tcull, unsigned int* overflow) {
        const int a[4] = {a0, a1, a2, a3};
        int f = 0;
        while (f < 4 && a[f] == kNoChild) ++f;
        if (f == 4) return pop(tcull);
        if (!room(3)) { if (overflow) *overflow = 1u; return a[f]; }
        for (int k = 3; k > f; --k) if (a[k] != kNoChild) push(a[k], 0.0f);
        return a[f];
    }
};

// One step through 4-wide node `cur`: slab-tests the four children against (tmin, tcull). ORDERED
// (closest-hit rays): returns the nearest hit child and pushes the others so that they pop in
// near-to-far order. Unordered (any-hit rays): returns the first hit child, pushes the rest. With
// no hit child it pops. Returns kNoChild when the traversal is finished. THE node step of every
// traversal in the library (closest hit, any hit, the bake kernel).
template <bool ORDERED, class STACK>
VLB_HD int bvh4_step(const BvhView& b, int cur, Vec3 idir, Vec3 ood, float tmin, float tcull, STACK& stk) {
    const float4* q = b.nodes + (size_t)kNodeQuads * cur;
    // near / far planes by the sign of the direction: no per-child min/max pairs
    const int sx = idir.x < 0.0f, sy = idir.y < 0.0f, sz = idir.z < 0.0f;
    float tn[4]; int r[4];
    const float inf = INFINITY;
#if VLB_NODE_Q8
    const float4 q0 = ld4(q), q1 = ld4(q + 1), q2 = ld4(q + 2), rf = ld4(q + 3);
    // t(byte) = (origin + byte * step - o) * idir = v * S + C with v = 1 + byte * 2^-15 (the byte dropped into
    // the mantissa of 1.0f), S = step * 2^15 * idir (exact: step is a power of two), C = (origin * idir - o * idir) - S.
    // Extra rounding against the fp32 format: < |S| * 2^-23 in t, i.e. < 1/256 step in space (margin: 1/64 step).
    const float Sx = f_mul(q0.w, idir.x), Sy = f_mul(q1.x, idir.y), Sz = f_mul(q1.y, idir.z);
    const float Cx = f_sub(f_fma(q0.x, idir.x, -ood.x), Sx), Cy = f_sub(f_fma(q0.y, idir.y, -ood.y), Sy),
                Cz = f_sub(f_fma(q0.z, idir.z, -ood.z), Sz);
    const uint32_t nxw = (uint32_t)f2i(sx ? q1.w : q1.z), fxw = (uint32_t)f2i(sx ? q1.z : q1.w);
    const uint32_t nyw = (uint32_t)f2i(sy ? q2.y : q2.x), fyw = (uint32_t)f2i(sy ? q2.x : q2.y);
    const uint32_t nzw = (uint32_t)f2i(sz ? q2.w : q2.z), fzw = (uint32_t)f2i(sz ? q2.z : q2.w);
    float4 tnx, tfx, tny, tfy, tnz, tfz;       // plane distances of the four children, near and far, per axis
    q8_planes(nxw, Sx, Cx, tnx); q8_planes(fxw, Sx, Cx, tfx);
    q8_planes(nyw, Sy, Cy, tny); q8_planes(fyw, Sy, Cy, tfy);
    q8_planes(nzw, Sz, Cz, tnz); q8_planes(fzw, Sz, Cz, tfz);
#define VLB_SLAB(k, c)                                                                                           \
    {                                                                                                            \
        const float a = fmaxf(max3(tnx.c, tny.c, tnz.c), tmin);                                                  \
        const float e = fminf(min3(tfx.c, tfy.c, tfz.c), tcull);                                                 \
        const int ref = f2i(rf.c);                                                                               \
        const bool hit = a <= e && ref != kNoChild;                                                              \
        tn[k] = hit ? a : inf;                                                                                   \
        r[k] = hit ? ref : kNoChild;                                                                             \
    }
#else
    float4 nx = ld4(q + sx), fx = ld4(q + 1 - sx);
    float4 ny = ld4(q + 2 + sy), fy = ld4(q + 3 - sy);
    float4 nz = ld4(q + 4 + sz), fz = ld4(q + 5 - sz);
    const float4 rf = ld4(q + 6);
#if defined(__CUDA_ARCH__) && VLB_FFMA2
    // the 24 plane distances plane * idir - o * idir as 12 packed fp32x2 FMAs (FFMA2, sm_100): per-element IEEE
    // round-to-nearest, so bit-identical to the scalar form
    fma2_planes(nx, idir.x, ood.x); fma2_planes(fx, idir.x, ood.x);
    fma2_planes(ny, idir.y, ood.y); fma2_planes(fy, idir.y, ood.y);
    fma2_planes(nz, idir.z, ood.z); fma2_planes(fz, idir.z, ood.z);
#define VLB_SLAB(k, c)                                                                                           \
    {                                                                                                            \
        const float a = fmaxf(max3(nx.c, ny.c, nz.c), tmin);                                                     \
        const float e = fminf(min3(fx.c, fy.c, fz.c), tcull);                                                    \
        const bool hit = a <= e;                                                                                 \
        tn[k] = hit ? a : inf;                                                                                   \
        r[k] = hit ? f2i(rf.c) : kNoChild;                                                                       \
    }
#else
#define VLB_SLAB(k, c)                                                                                           \
    {                                                                                                            \
        const float a = fmaxf(max3(f_fma(nx.c, idir.x, -ood.x), f_fma(ny.c, idir.y, -ood.y),                      \
                                   f_fma(nz.c, idir.z, -ood.z)), tmin);                                          \
        const float e = fminf(min3(f_fma(fx.c, idir.x, -ood.x), f_fma(fy.c, idir.y, -ood.y),                      \
                                   f_fma(fz.c, idir.z, -ood.z)), tcull);                                         \
        const bool hit = a <= e;                                                                                 \
        tn[k] = hit ? a : inf;                                                                                   \
        r[k] = hit ? f2i(rf.c) : kNoChild;                                                                       \
    }
#endif
#endif
    VLB_SLAB(0, x) VLB_SLAB(1, y) VLB_SLAB(2, z) VLB_SLAB(3, w)
#undef VLB_SLAB
#if VLB_NODE_Q8 && VLB_NODE_ORDER1D
    if (ORDERED) {
        // slot order = ascending along the node's ordering axis (store_node4): a ray running against that axis visits
        // the slots in descending order. No distance sort.
        const int ax = f2i(q0.w) & 3;
        const bool rev = (ax == 0 ? sx : (ax == 1 ? sy : sz)) != 0;
        return stk.advance_unsorted(rev ? r[3] : r[0], rev ? r[2] : r[1], rev ? r[1] : r[2], rev ? r[0] : r[3], tcull, b.overflow);
    }
#endif
    if (ORDERED) {
        // sort the (entry distance, ref) pairs ascending; misses carry +inf and sink to the end
        order2(tn[0], r[0], tn[1], r[1]);
        order2(tn[2], r[2], tn[3], r[3]);
        order2(tn[0], r[0], tn[2], r[2]);
        order2(tn[1], r[1], tn[3], r[3]);
        order2(tn[1], r[1], tn[2], r[2]);
        return stk.advance(r[0], r[1], r[2], r[3], tn[1], tn[2], tn[3], tcull, b.overflow);
    }
    if (!stk.room(4)) {
        if (b.overflow) *b.overflow = 1u;
    } else {
        if (r[3] != kNoChild) stk.push(r[3], tn[3]);
        if (r[2] != kNoChild) stk.push(r[2], tn[2]);
        if (r[1] != kNoChild) stk.push(r[1], tn[1]);
        if (r[0] != kNoChild) stk.push(r[0], tn[0]);
    }
    return stk.pop(tcull);
}

#if VLB_BVH8
// One step through 8-wide node `cur` (closest-hit and any-hit rays alike): slab-tests the eight children, enters the
// first hit one in octant order and leaves the others as one (node, mask) group on the stack.
template <class STACK>
VLB_HD int bvh8_step(const BvhView& b, int cur, Vec3 idir, Vec3 ood, float tmin, float tcull, STACK& stk) {
    const float4* q = b.nodes + (size_t)kNodeQuads * cur;
    const float4 q0 = ld4(q), q1 = ld4(q + 1), q2 = ld4(q + 2), q3 = ld4(q + 3);
    const int sx = idir.x < 0.0f, sy = idir.y < 0.0f, sz = idir.z < 0.0f;
    const uint32_t eb = (uint32_t)f2i(q0.w);
    // t(byte) = v * S + C with v = 1 + byte * 2^-15, S = step * 2^15 * idir, C = (origin * idir - o * idir) - S (see the 4-wide 8-bit format)
    const float Sx = f_mul(i2f((int)((eb & 0xffu) << 23)), idir.x), Sy = f_mul(i2f((int)(((eb >> 8) & 0xffu) << 23)), idir.y),
                Sz = f_mul(i2f((int)(((eb >> 16) & 0xffu) << 23)), idir.z);
    const float Cx = f_sub(f_fma(q0.x, idir.x, -ood.x), Sx), Cy = f_sub(f_fma(q0.y, idir.y, -ood.y), Sy),
                Cz = f_sub(f_fma(q0.z, idir.z, -ood.z), Sz);
    float4 nx0, nx1, fx0, fx1, ny0, ny1, fy0, fy1, nz0, nz1, fz0, fz1;   // near / far plane distances, children 0..3 and 4..7
    q8_planes((uint32_t)f2i(sx ? q1.z : q1.x), Sx, Cx, nx0); q8_planes((uint32_t)f2i(sx ? q1.w : q1.y), Sx, Cx, nx1);
    q8_planes((uint32_t)f2i(sx ? q1.x : q1.z), Sx, Cx, fx0); q8_planes((uint32_t)f2i(sx ? q1.y : q1.w), Sx, Cx, fx1);
    q8_planes((uint32_t)f2i(sy ? q2.z : q2.x), Sy, Cy, ny0); q8_planes((uint32_t)f2i(sy ? q2.w : q2.y), Sy, Cy, ny1);
    q8_planes((uint32_t)f2i(sy ? q2.x : q2.z), Sy, Cy, fy0); q8_planes((uint32_t)f2i(sy ? q2.y : q2.w), Sy, Cy, fy1);
    q8_planes((uint32_t)f2i(sz ? q3.z : q3.x), Sz, Cz, nz0); q8_planes((uint32_t)f2i(sz ? q3.w : q3.y), Sz, Cz, nz1);
    q8_planes((uint32_t)f2i(sz ? q3.x : q3.z), Sz, Cz, fz0); q8_planes((uint32_t)f2i(sz ? q3.y : q3.w), Sz, Cz, fz1);
    unsigned m = 0u;      // bit k: the ray enters child slot k (empty slots have inverted boxes and never hit)
#define VLB_SLAB8(k, h, c)                                                                                       \
    {                                                                                                            \
        const float a = fmaxf(max3(nx##h.c, ny##h.c, nz##h.c), tmin);                                            \
        const float e = fminf(min3(fx##h.c, fy##h.c, fz##h.c), tcull);                                           \
        m |= (a <= e) ? (1u << (k)) : 0u;                                                                        \
    }
    VLB_SLAB8(0, 0, x) VLB_SLAB8(1, 0, y) VLB_SLAB8(2, 0, z) VLB_SLAB8(3, 0, w)
    VLB_SLAB8(4, 1, x) VLB_SLAB8(5, 1, y) VLB_SLAB8(6, 1, z) VLB_SLAB8(7, 1, w)
#undef VLB_SLAB8
    // slot order -> priority order: bit p of the result = bit (p ^ flip) of m (an XOR on bit indices = three swap stages)
    if (sx) m = ((m & 0x55u) << 1) | ((m >> 1) & 0x55u);
    if (sy) m = ((m & 0x33u) << 2) | ((m >> 2) & 0x33u);
    if (sz) m = ((m & 0x0fu) << 4) | ((m >> 4) & 0x0fu);
    return stk.descend(b, cur, m, sx | (sy << 1) | (sz << 2));
}
#endif

// The node step / the "what next" after a leaf of the configured tree width. Every traversal in the library goes
// through these two (closest hit, any hit, the bake kernel's main rays and its visibility-ray batches).
template <bool ORDERED, class STACK>
VLB_HD int traverse_step(const BvhView& b, int cur, Vec3 idir, Vec3 ood, float tmin, float tcull, STACK& stk) {
#if VLB_BVH8
    return bvh8_step(b, cur, idir, ood, tmin, tcull, stk);
#else
    return bvh4_step<ORDERED>(b, cur, idir, ood, tmin, tcull, stk);
#endif
}
template <class STACK>
VLB_HD int traverse_pop(const BvhView& b, Vec3 idir, float tcull, STACK& stk) {
#if VLB_BVH8
    (void)tcull;
    return stk.pop_group(b, (idir.x < 0.0f ? 1 : 0) | (idir.y < 0.0f ? 2 : 0) | (idir.z < 0.0f ? 4 : 0));
#else
    (void)b; (void)idir;
    return stk.pop(tcull);
#endif
}

// Intersects the triangles of leaf `ref` (< 0). Closest-hit rays update `best` and the culling
// distance; any-hit rays return true at the first triangle inside (tmin, tcull).
template <bool ANY, bool COUNT>
VLB_HD bool leaf_step(const BvhView& b, int ref, Vec3 o, Vec3 d, float tmin, float& tcull, HitRec& best, TraceCounters* cnt) {
    const int x = ~ref;
    const int first = x >> 3, count = (x & 7) + 1;
    for (int k = 0; k < count; ++k) {
        const float4 v0 = ld4(b.tris + 3 * (first + k) + 0);
        const float4 e1 = ld4(b.tris + 3 * (first + k) + 1);
        const float4 e2 = ld4(b.tris + 3 * (first + k) + 2);
        if (COUNT) cnt->tris++;
        float t, u, v;
        if (intersect_tri(v0, e1, e2, o, d, t, u, v) && t > tmin) {
            const int id = f2i(v0.w);
            if (ANY) {
                if (t < tcull) { best.id = id; best.t = t; best.u = u; best.v = v; return true; }
            } else if (t < best.t || (t == best.t && best.id >= 0 && id < best.id)) {
                best.id = id; best.t = t; best.u = u; best.v = v;
                tcull = t * kCullSlack;
            }
        }
    }
    return false;
}

// Closest hit with tmin < t < tmax; ties: smaller t, then smaller flat id.
template <bool COUNT>
VLB_HD HitRec trace_closest(const BvhView& b, Vec3 o, Vec3 d, float tmin, float tmax, TraceCounters* cnt) {
    HitRec best;
    best.id = -1; best.t = tmax; best.u = 0.f; best.v = 0.f;
    if (b.n_tris == 0) return best;
    const Vec3 idir = mk3(safe_inv(d.x), safe_inv(d.y), safe_inv(d.z));
    const Vec3 ood = mk3(o.x * idir.x, o.y * idir.y, o.z * idir.z);
    LocalStack stk;
    stk.clear();
    int cur = 0;
    float tcull = tmax;
    while (cur != kNoChild) {
        if (cur >= 0) {
            if (COUNT) cnt->nodes++;
            cur = traverse_step<true>(b, cur, idir, ood, tmin, tcull, stk);
        } else {
            leaf_step<false, COUNT>(b, cur, o, d, tmin, tcull, best, cnt);
            cur = traverse_pop(b, idir, tcull, stk);
        }
    }
    return best;
}

// Any hit with tmin < t < tmax (shadow rays; gl_RayFlagsTerminateOnFirstHitEXT).
template <bool COUNT>
VLB_HD bool trace_any(const BvhView& b, Vec3 o, Vec3 d, float tmin, float tmax, TraceCounters* cnt, HitRec* out) {
    if (b.n_tris == 0) return false;
    const Vec3 idir = mk3(safe_inv(d.x), safe_inv(d.y), safe_inv(d.z));
    const Vec3 ood = mk3(o.x * idir.x, o.y * idir.y, o.z * idir.z);
    LocalStack stk;
    stk.clear();
    int cur = 0;
    float tcull = tmax;
    HitRec h; h.id = -1; h.t = tmax; h.u = 0.f; h.v = 0.f;
    while (cur != kNoChild) {
        if (cur >= 0) {
            if (COUNT) cnt->nodes++;
            cur = traverse_step<false>(b, cur, idir, ood, tmin, tcull, stk);
        } else {
            if (leaf_step<true, COUNT>(b, cur, o, d, tmin, tcull, h, cnt)) {
                if (out) *out = h;
                return true;
            }
            cur = traverse_pop(b, idir, tcull, stk);
        }
    }
    return false;
}

}  // namespace vlb

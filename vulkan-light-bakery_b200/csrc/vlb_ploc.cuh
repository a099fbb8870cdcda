// vlb_ploc.cuh — PLOC (parallel locally-ordered clustering, Meister & Bittner 2018): the "prefer fast trace" builder
// (the reference asks its driver for VK_BUILD_ACCELERATION_STRUCTURE_PREFER_FAST_TRACE_BIT,
// src/scene_manager.cpp:346-347) next to the Karras LBVH. A binary hierarchy over the Morton-sorted
// triangles built by agglomeration; the wide-node collapse (emit_wide_node) and the traversal are the same for both
// builders, and hit ids stay bit-exact with any tree. Selected with vlb_bvh_set_builder / VLB_BVH_BUILDER=ploc;
// bvh_build.cu runs the rounds in one cooperative kernel, tests/emu runs the same per-item bodies serially.
// Measured on the bake's own ray distribution (profiles/r02_bvh_builder_ab.log): 2.5-4.6 % fewer node visits than
// the LBVH on the 262,144-triangle atrium (radius 8-64), -0.7..+3.7 % at 1 M triangles.
//
// Clusters start as the n leaves in Morton order. Every round, each cluster looks `radius` places left and right
// in the CURRENT order for the neighbour whose merged box has the smallest surface area (ties: smaller index);
// pairs that chose each other merge into a new internal node which takes the lower member's place, the rest stay;
// the order is compacted and the round repeats until one cluster is left. The pair of globally smallest merged
// area is always mutual (smallest-index tie-break), so every round makes progress, and every decision is a pure
// function of the inputs: the tree is deterministic.
//
// Node references inside the builder: leaf i (position in Morton order) = i, internal node k (creation order) =
// n + k; boxes live in one array of 2 float4 per reference. ploc_finish_* turn that into what emit_node4
// (vlb_bvh.cuh) consumes: internal ids renumbered so that the root is 0, leaves renumbered into depth-first order
// (every subtree then owns a contiguous triangle range, which is how leaf_ref addresses its triangles), and
// first / last per internal node.
//
// Per-item bodies are `__host__ __device__` so that tests/emu runs the same code serially on the CPU.
#pragma once

#include "vlb_bvh.cuh"

namespace vlb {

// Everything another block wrote before the last grid-wide barrier -- the cluster order, the neighbour choices, and
// the boxes / counts of nodes created in earlier rounds (a new node shares its 128-byte line with older ones, so a line
// cached before the node existed would be stale) -- is read past the non-coherent L1 on the device.
VLB_HD int ld_cg(const int* p) {
#ifdef __CUDA_ARCH__
    return __ldcg(p);
#else
    return *p;
#endif
}

VLB_HD float4 ld_cg4(const float4* p) {
#ifdef __CUDA_ARCH__
    return __ldcg(p);
#else
    return *p;
#endif
}

VLB_HD float merged_half_area(const float4 alo, const float4 ahi, const float4 blo, const float4 bhi) {
    const float dx = fmaxf(ahi.x, bhi.x) - fminf(alo.x, blo.x);
    const float dy = fmaxf(ahi.y, bhi.y) - fminf(alo.y, blo.y);
    const float dz = fmaxf(ahi.z, bhi.z) - fminf(alo.z, blo.z);
    return f_add(f_add(f_mul(dx, dy), f_mul(dy, dz)), f_mul(dz, dx));      // no contraction: same value on host and device
}

// Nearest neighbour of cluster i among positions [i - radius, i + radius] of the current order `C` (m clusters).
VLB_HD int ploc_nearest(const int* C, int m, int i, int radius, const float4* box) {
    const int ci = ld_cg(C + i);
    const float4 lo = ld_cg4(box + 2 * (size_t)ci), hi = ld_cg4(box + 2 * (size_t)ci + 1);
    const int j0 = i - radius < 0 ? 0 : i - radius, j1 = i + radius > m - 1 ? m - 1 : i + radius;
    int best = -1;
    float best_cost = INFINITY;
    for (int j = j0; j <= j1; ++j) {
        if (j == i) continue;
        const int cj = ld_cg(C + j);
        const float c = merged_half_area(lo, hi, ld_cg4(box + 2 * (size_t)cj), ld_cg4(box + 2 * (size_t)cj + 1));
        if (c < best_cost) { best_cost = c; best = j; }
    }
    return best;
}

// What position i does this round: 0 = stays, 1 = lower member of a mutual pair (creates the node), 2 = upper member (leaves).
VLB_HD int ploc_role(const int* nn, int i) {
    const int j = ld_cg(nn + i);
    if (j < 0 || ld_cg(nn + j) != i) return 0;
    return i < j ? 1 : 2;
}

// Lower member i of a mutual pair creates internal node `k` (creation index) from C[i] (left) and C[nn[i]] (right).
VLB_HD int ploc_merge(const int* C, const int* nn, int i, int k, int n, float4* box, int* left, int* right, int* parent, int* count,
                      int* leftmost) {
    const int a = ld_cg(C + i), b = ld_cg(C + ld_cg(nn + i)), node = n + k;
    left[k] = a; right[k] = b;
    parent[a] = node; parent[b] = node;
    const float4 alo = ld_cg4(box + 2 * (size_t)a), ahi = ld_cg4(box + 2 * (size_t)a + 1), blo = ld_cg4(box + 2 * (size_t)b), bhi = ld_cg4(box + 2 * (size_t)b + 1);
    box[2 * (size_t)node] = make_float4(fminf(alo.x, blo.x), fminf(alo.y, blo.y), fminf(alo.z, blo.z), 0.f);
    box[2 * (size_t)node + 1] = make_float4(fmaxf(ahi.x, bhi.x), fmaxf(ahi.y, bhi.y), fmaxf(ahi.z, bhi.z), 0.f);
    count[node] = ld_cg(count + a) + ld_cg(count + b);
    leftmost[node] = ld_cg(leftmost + a);
    return node;
}

// Depth-first position of leaf i: the triangles left of it = sum, over the ancestors it hangs under on the RIGHT, of
// the left sibling's triangle count.
VLB_HD int ploc_leaf_position(int i, int n, const int* left, const int* parent, const int* count) {
    int pos = 0, c = i;
    for (int p = parent[c]; p >= 0; p = parent[p]) {
        const int l = left[p - n];
        if (l != c) pos += count[l];
        c = p;
    }
    return pos;
}

// Builder reference -> emit_node4 child encoding: internal creation index k -> id (n - 2) - k (the root, created
// last, becomes 0); leaf i -> ~(its depth-first position).
VLB_HD int ploc_emit_ref(int ref, int n, const int* leaf_pos) { return ref >= n ? (n - 2) - (ref - n) : ~leaf_pos[ref]; }

// Internal node k (creation index) in emit_node4's arrays (indexed by the renumbered id).
VLB_HD void ploc_finish_node(int k, int n, const int* left, const int* right, const int* count, const int* leftmost, const int* leaf_pos,
                             const float4* box, int* e_left, int* e_right, int* e_first, int* e_last, float4* e_ibox) {
    const int id = (n - 2) - k, node = n + k;
    e_left[id] = ploc_emit_ref(left[k], n, leaf_pos);
    e_right[id] = ploc_emit_ref(right[k], n, leaf_pos);
    e_first[id] = leaf_pos[leftmost[node]];
    e_last[id] = e_first[id] + count[node] - 1;
    e_ibox[2 * (size_t)id] = box[2 * (size_t)node];
    e_ibox[2 * (size_t)id + 1] = box[2 * (size_t)node + 1];
}

}  // namespace vlb

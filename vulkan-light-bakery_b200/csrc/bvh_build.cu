// bvh_build.cu — software LBVH build on the GPU; replaces the driver acceleration-structure
// build of the reference (Scene_t::buildAccelerationStructures, src/scene_manager.cpp:385-443).
//
//   k_scene_bounds   tight AABB + centroid AABB of the flattened triangles   (reads 48 B/tri)
//   k_morton         63-bit Morton key of each centroid                       (reads 48 B/tri)
//   radix sort       (key, flat id) pairs, 8 passes of 8 bits (radix_sort.cuh, hand-written)
//   k_gather         triangles into Morton order + leaf AABBs                 (48 B in, 80 B out)
//   k_karras         Karras-2012 hierarchy, one thread per internal node
//   k_refit          bottom-up AABB refit, atomic arrival flags
//   k_emit_levels    4-wide traversal nodes (vlb_bvh.cuh), top-down in one cooperative launch: children chosen greedily by
//                    surface area, small subtrees -> leaves
// Optional second builder for the binary hierarchy (vlb_bvh_set_builder / VLB_BVH_BUILDER=ploc; replaces k_karras + k_refit):
//   k_ploc_rounds    PLOC agglomeration (vlb_ploc.cuh), all rounds in one cooperative launch
//   k_ploc_leaf_pos, k_ploc_permute, k_ploc_finish   depth-first leaf order, triangles moved into it, arrays for k_emit_levels

#include "vlb_bvh.cuh"
#include "vlb_ploc.cuh"
#include <cooperative_groups.h>

#include "radix_sort.cuh"
#include "vlb_context.h"

namespace cg = cooperative_groups;

namespace vlb {

__device__ __forceinline__ void atomic_min_float(float* addr, float v) {
    if (v >= 0.f) atomicMin(reinterpret_cast<int*>(addr), __float_as_int(v));
    else atomicMax(reinterpret_cast<unsigned*>(addr), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
    if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
    else atomicMin(reinterpret_cast<unsigned*>(addr), __float_as_uint(v));
}

// scratch[0..2] tight lo, [3..5] tight hi, [6..8] centroid lo, [9..11] centroid hi, [12] nodes emitted (u32),
// [13] traversal stack overflow flag (u32, set by any traversal of this BVH)
__global__ void k_init_bounds(float* scratch) {
    const int i = threadIdx.x;
    if (i < 12) scratch[i] = ((i / 3) & 1) ? -INFINITY : INFINITY;
    if (i >= 12 && i < 16) scratch[i] = 0.f;   // [12] emitted-node counter, [13] traversal stack overflow flag
}

__global__ void k_scene_bounds(const float4* __restrict__ tri_flat, uint32_t n, float* scratch) {
    float v[12];
    for (int k = 0; k < 12; ++k) v[k] = ((k / 3) & 1) ? -INFINITY : INFINITY;
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
        float4 lo, hi;
        tri_aabb(tri_flat[3ull * t], tri_flat[3ull * t + 1], tri_flat[3ull * t + 2], &lo, &hi);
        const float c[3] = {0.5f * (lo.x + hi.x), 0.5f * (lo.y + hi.y), 0.5f * (lo.z + hi.z)};
        v[0] = fminf(v[0], lo.x); v[1] = fminf(v[1], lo.y); v[2] = fminf(v[2], lo.z);
        v[3] = fmaxf(v[3], hi.x); v[4] = fmaxf(v[4], hi.y); v[5] = fmaxf(v[5], hi.z);
        for (int k = 0; k < 3; ++k) { v[6 + k] = fminf(v[6 + k], c[k]); v[9 + k] = fmaxf(v[9 + k], c[k]); }
    }
    for (int k = 0; k < 12; ++k) {
        const bool is_max = (k / 3) & 1;
        for (int off = 16; off > 0; off >>= 1) {
            const float o = __shfl_xor_sync(0xffffffffu, v[k], off);
            v[k] = is_max ? fmaxf(v[k], o) : fminf(v[k], o);
        }
    }
    // one set of 12 atomics per BLOCK (per warp they were 98 k atomics on 12 addresses at 262 k triangles: the kernel spent
    // 68 us on a 12 MB read)
    __shared__ float s_v[12][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = (blockDim.x + 31) >> 5;
    if (lane == 0) for (int k = 0; k < 12; ++k) s_v[k][warp] = v[k];
    __syncthreads();
    if (threadIdx.x < 12) {
        const int k = threadIdx.x;
        const bool is_max = (k / 3) & 1;
        float r = s_v[k][0];
        for (int w = 1; w < n_warps; ++w) r = is_max ? fmaxf(r, s_v[k][w]) : fminf(r, s_v[k][w]);
        if (is_max) atomic_max_float(scratch + k, r);
        else atomic_min_float(scratch + k, r);
    }
}

// cubic: all three axes are normalised by the LARGEST centroid extent, so Morton cells are cubes and a flat scene
// spends no key bits (= no tree levels) on splits along its short axes; else every axis is stretched to [0, 1].
__global__ void k_morton(const float4* __restrict__ tri_flat, uint32_t n, const float* __restrict__ scratch, int cubic,
                         uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    float4 lo, hi;
    tri_aabb(tri_flat[3ull * t], tri_flat[3ull * t + 1], tri_flat[3ull * t + 2], &lo, &hi);
    const float c[3] = {0.5f * (lo.x + hi.x), 0.5f * (lo.y + hi.y), 0.5f * (lo.z + hi.z)};
    float nrm[3];
    const float ext_max = fmaxf(scratch[9] - scratch[6], fmaxf(scratch[10] - scratch[7], scratch[11] - scratch[8]));
    for (int k = 0; k < 3; ++k) {
        const float ext = cubic ? ext_max : scratch[9 + k] - scratch[6 + k];
        nrm[k] = ext > 0.f ? (c[k] - scratch[6 + k]) / ext : 0.f;
    }
    keys[t] = morton63(nrm[0], nrm[1], nrm[2]);
    vals[t] = t;
}

__global__ void k_gather(const float4* __restrict__ tri_flat, const uint32_t* __restrict__ order, uint32_t n,
                         float4* __restrict__ tris, float4* __restrict__ lbox) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint32_t t = order[j];
    const float4 v0 = tri_flat[3ull * t], e1 = tri_flat[3ull * t + 1], e2 = tri_flat[3ull * t + 2];
    tris[3ull * j] = v0; tris[3ull * j + 1] = e1; tris[3ull * j + 2] = e2;
    float4 lo, hi;
    tri_aabb(v0, e1, e2, &lo, &hi);
    lbox[2ull * j] = lo; lbox[2ull * j + 1] = hi;
}

__global__ void k_karras(const uint64_t* __restrict__ keys, int n, int* left, int* right, int* first, int* last,
                         int* parent_i, int* parent_l) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    if (i == 0) parent_i[0] = -1;
    karras_node(keys, n, i, left, right, first, last, parent_i, parent_l);
}

__global__ void k_refit(int n, const int* __restrict__ left, const int* __restrict__ right,
                        const int* __restrict__ parent_i, const int* __restrict__ parent_l,
                        const float4* __restrict__ lbox, float4* ibox, int* flags) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    int cur = parent_l[j];
    while (cur >= 0) {
        if (atomicAdd(&flags[cur], 1) == 0) return;   // first arrival: the sibling subtree is not done yet
        __threadfence();
        float4 lo[2], hi[2];
        const int ch[2] = {left[cur], right[cur]};
        for (int c = 0; c < 2; ++c) {
            if (ch[c] < 0) { lo[c] = lbox[2 * (~ch[c])]; hi[c] = lbox[2 * (~ch[c]) + 1]; }
            else { lo[c] = __ldcg(&ibox[2 * ch[c]]); hi[c] = __ldcg(&ibox[2 * ch[c] + 1]); }
        }
        ibox[2 * cur] = make_float4(fminf(lo[0].x, lo[1].x), fminf(lo[0].y, lo[1].y), fminf(lo[0].z, lo[1].z), 0.f);
        ibox[2 * cur + 1] = make_float4(fmaxf(hi[0].x, hi[1].x), fmaxf(hi[0].y, hi[1].y), fmaxf(hi[0].z, hi[1].z), 0.f);
        __threadfence();
        cur = parent_i[cur];
    }
}

// Top-down collapse into 4-wide nodes in ONE cooperative launch: the grid walks the wide tree level by level
// (grid.sync between levels, no host round trips). Every thread takes wide-node roots of the current frontier,
// emits their nodes (emit_node4) and appends the roots of the next level. Three rotating counters: the level
// being read, the one being filled, and the one being cleared for the level after.
__global__ void __launch_bounds__(256) k_emit_levels(int* __restrict__ frontier_a, int* __restrict__ frontier_b, unsigned int* cnt /*[3]*/,
                                                     const int* __restrict__ left, const int* __restrict__ right,
                                                     const int* __restrict__ first, const int* __restrict__ last,
                                                     const float4* __restrict__ ibox, const float4* __restrict__ lbox, int max_leaf,
                                                     const float* __restrict__ scratch, float4* __restrict__ nodes,
                                                     unsigned int* n_emitted) {
    cg::grid_group grid = cg::this_grid();
    const float ext = fmaxf(scratch[3] - scratch[0], fmaxf(scratch[4] - scratch[1], scratch[5] - scratch[2]));
    const unsigned int gtid = blockIdx.x * blockDim.x + threadIdx.x, gsize = gridDim.x * blockDim.x;
    unsigned int total = 0;
    for (int level = 0;; ++level) {
        const unsigned int n_front = *reinterpret_cast<volatile unsigned int*>(cnt + level % 3);
        if (n_front == 0) break;
        total += n_front;
        const int* cur = (level & 1) ? frontier_b : frontier_a;
        int* next = (level & 1) ? frontier_a : frontier_b;
        if (gtid == 0) cnt[(level + 2) % 3] = 0;
        for (unsigned int t = gtid; t < n_front; t += gsize)
            emit_wide_node(__ldcg(cur + t), left, right, first, last, ibox, lbox, max_leaf, ext * 1e-6f, nodes, next, cnt + (level + 1) % 3);
        __threadfence();
        grid.sync();
    }
    if (gtid == 0) *n_emitted = total;
}

// ---- PLOC (vlb_ploc.cuh) -------------------------------------------------------------------------------------
__global__ void k_ploc_init(int n, int* __restrict__ C, int* __restrict__ leftmost, int* __restrict__ count, int* __restrict__ parent) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 2 * n - 1) { parent[i] = -1; count[i] = 1; leftmost[i] = i; }
    if (i < n) C[i] = i;
}

constexpr int kPlocTail = 1024;     // cluster count below which one block finishes the agglomeration alone
constexpr int kPlocMaxRadius = 64;

// ploc_nearest for the 256 positions [t0, t0 + 256) of the current order, through shared memory: the boxes of positions
// [t0 - radius, t0 + 256 + radius) are staged once (one coalesced read of the order + one gather of the boxes, instead of
// 2 * radius dependent trips to L2 per thread), then every thread scans its window there. Same candidates, same
// arithmetic, same tie-break as ploc_nearest. All threads of the block must call it.
__device__ __forceinline__ void ploc_nearest_tile(const int* cur, int m, int t0, int radius, const float4* box, int* nn,
                                                  float4* s_lo, float4* s_hi) {
    const int tid = threadIdx.x, first = t0 - radius, span = (int)blockDim.x + 2 * radius;
    for (int k = tid; k < span; k += blockDim.x) {
        const int j = first + k;
        if (j >= 0 && j < m) {
            const int cj = ld_cg(cur + j);
            s_lo[k] = ld_cg4(box + 2 * (size_t)cj); s_hi[k] = ld_cg4(box + 2 * (size_t)cj + 1);
        }
    }
    __syncthreads();
    const int i = t0 + tid;
    if (i < m) {
        const float4 lo = s_lo[tid + radius], hi = s_hi[tid + radius];
        const int j0 = i - radius < 0 ? 0 : i - radius, j1 = i + radius > m - 1 ? m - 1 : i + radius;
        int best = -1;
        float best_cost = INFINITY;
        for (int j = j0; j <= j1; ++j) {
            if (j == i) continue;
            const float c = merged_half_area(lo, hi, s_lo[j - first], s_hi[j - first]);
            if (c < best_cost) { best_cost = c; best = j; }
        }
        nn[i] = best;
    }
    __syncthreads();
}

// All rounds of the agglomeration in one cooperative launch. A round: (1) every cluster picks its nearest neighbour
// within `radius` places of the current order; (2) every block counts, over its contiguous chunk of the order, the
// clusters that stay or create a node and the nodes created (no barrier between (1) and (2): see the halo below);
// (3) with the block totals every block knows where its
// chunk lands in the next order and which creation indices it hands out, and an ordered block scan places each
// cluster. Output position and creation index of a cluster are "how many before it", exactly the serial loop of
// tests/emu, so the tree is the same whatever the grid size.
__global__ void __launch_bounds__(256) k_ploc_rounds(int n, int radius, int tail, int* C, int* Cn, int* nn, float4* box, int* left, int* right,
                                                     int* parent, int* count, int* leftmost, int2* block_tot) {
    cg::grid_group grid = cg::this_grid();
    __shared__ int s_w[4][8];
    __shared__ float4 s_lo[256 + 2 * kPlocMaxRadius], s_hi[256 + 2 * kPlocMaxRadius];
    const int nb = gridDim.x, b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned lt = (1u << lane) - 1u;
    int m = n, created = 0;
    int* cur = C; int* nxt = Cn;
    while (m > 1) {
        if (m <= tail) {
            // The last rounds (about half of them) involve a few hundred clusters: block 0 finishes them alone, with block
            // barriers instead of grid-wide ones (a grid.sync costs microseconds, these rounds nothing).
            if (b != 0) return;
            while (m > 1) {
                for (int t0 = 0; t0 < m; t0 += blockDim.x) ploc_nearest_tile(cur, m, t0, radius, box, nn, s_lo, s_hi);
                int run_k = 0, run_g = created;
                for (int base = 0; base < m; base += blockDim.x) {
                    const int i = base + tid;
                    const int role = i < m ? ploc_role(nn, i) : 2;
                    const bool fk = role != 2, fg = role == 1;
                    const unsigned bk = __ballot_sync(0xffffffffu, fk), bg = __ballot_sync(0xffffffffu, fg);
                    if (lane == 0) { s_w[0][warp] = __popc(bk); s_w[1][warp] = __popc(bg); }
                    __syncthreads();
                    int wk = 0, wg = 0, totk = 0, totg = 0;
                    for (int w = 0; w < 8; ++w) {
                        if (w < warp) { wk += s_w[0][w]; wg += s_w[1][w]; }
                        totk += s_w[0][w]; totg += s_w[1][w];
                    }
                    if (fk) {
                        const int out = run_k + wk + __popc(bk & lt);
                        nxt[out] = fg ? ploc_merge(cur, nn, i, run_g + wg + __popc(bg & lt), n, box, left, right, parent, count, leftmost)
                                      : ld_cg(cur + i);
                    }
                    run_k += totk; run_g += totg;
                    __syncthreads();
                }
                created = run_g; m = run_k;
                int* t = cur; cur = nxt; nxt = t;
                __threadfence_block();
                __syncthreads();
            }
            return;
        }
        const int chunk = (m + nb - 1) / nb, lo = min(m, b * chunk), hi = min(m, lo + chunk);
        // The neighbour choices of this block's chunk AND of `radius` positions on either side: a cluster's role needs its
        // neighbour's choice, which lies within the radius, so the block can count its chunk without waiting for the
        // others (two grid-wide barriers per round instead of three). Positions in the overlap are computed by two
        // blocks and written twice with the same value.
        if (lo < hi)
            for (int t0 = max(0, lo - radius); t0 < min(m, hi + radius); t0 += blockDim.x) ploc_nearest_tile(cur, m, t0, radius, box, nn, s_lo, s_hi);
        __syncthreads();
        int kept = 0, merged = 0;
        for (int i = lo + tid; i < hi; i += blockDim.x) { const int role = ploc_role(nn, i); kept += role != 2; merged += role == 1; }
        for (int off = 16; off > 0; off >>= 1) { kept += __shfl_xor_sync(0xffffffffu, kept, off); merged += __shfl_xor_sync(0xffffffffu, merged, off); }
        if (lane == 0) { s_w[0][warp] = kept; s_w[1][warp] = merged; }
        __syncthreads();
        if (tid == 0) {
            int k = 0, g = 0;
            for (int w = 0; w < 8; ++w) { k += s_w[0][w]; g += s_w[1][w]; }
            block_tot[b] = make_int2(k, g);
        }
        __threadfence();
        grid.sync();
        // totals of the round and of the blocks before this one
        int ok = 0, og = 0, tk = 0, tg = 0;
        for (int j = tid; j < nb; j += blockDim.x) {
            const int2 t = __ldcg(block_tot + j);
            tk += t.x; tg += t.y;
            if (j < b) { ok += t.x; og += t.y; }
        }
        for (int off = 16; off > 0; off >>= 1) {
            ok += __shfl_xor_sync(0xffffffffu, ok, off); og += __shfl_xor_sync(0xffffffffu, og, off);
            tk += __shfl_xor_sync(0xffffffffu, tk, off); tg += __shfl_xor_sync(0xffffffffu, tg, off);
        }
        __syncthreads();
        if (lane == 0) { s_w[0][warp] = ok; s_w[1][warp] = og; s_w[2][warp] = tk; s_w[3][warp] = tg; }
        __syncthreads();
        ok = og = tk = tg = 0;
        for (int w = 0; w < 8; ++w) { ok += s_w[0][w]; og += s_w[1][w]; tk += s_w[2][w]; tg += s_w[3][w]; }
        __syncthreads();
        int run_k = ok, run_g = created + og;
        for (int base = lo; base < hi; base += blockDim.x) {
            const int i = base + tid;
            const int role = i < hi ? ploc_role(nn, i) : 2;
            const bool fk = role != 2, fg = role == 1;
            const unsigned bk = __ballot_sync(0xffffffffu, fk), bg = __ballot_sync(0xffffffffu, fg);
            if (lane == 0) { s_w[0][warp] = __popc(bk); s_w[1][warp] = __popc(bg); }
            __syncthreads();
            int wk = 0, wg = 0, totk = 0, totg = 0;
            for (int w = 0; w < 8; ++w) {
                if (w < warp) { wk += s_w[0][w]; wg += s_w[1][w]; }
                totk += s_w[0][w]; totg += s_w[1][w];
            }
            if (fk) {
                const int out = run_k + wk + __popc(bk & lt);
                nxt[out] = fg ? ploc_merge(cur, nn, i, run_g + wg + __popc(bg & lt), n, box, left, right, parent, count, leftmost)
                              : ld_cg(cur + i);
            }
            run_k += totk; run_g += totg;
            __syncthreads();
        }
        m = tk; created += tg;
        int* t = cur; cur = nxt; nxt = t;
        __threadfence();
        grid.sync();
    }
}

__global__ void k_ploc_leaf_pos(int n, const int* __restrict__ left, const int* __restrict__ parent, const int* __restrict__ count, int* __restrict__ pos) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) pos[i] = ploc_leaf_position(i, n, left, parent, count);
}
__global__ void k_ploc_permute(int n, const int* __restrict__ pos, const float4* __restrict__ tris_in, const float4* __restrict__ lbox_in,
                               float4* __restrict__ tris_out, float4* __restrict__ lbox_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const size_t p = (size_t)pos[i];
    for (int k = 0; k < 3; ++k) tris_out[3 * p + k] = tris_in[3 * (size_t)i + k];
    lbox_out[2 * p] = lbox_in[2 * (size_t)i]; lbox_out[2 * p + 1] = lbox_in[2 * (size_t)i + 1];
}
__global__ void k_ploc_finish(int n, const int* __restrict__ left, const int* __restrict__ right, const int* __restrict__ count,
                              const int* __restrict__ leftmost, const int* __restrict__ pos, const float4* __restrict__ box,
                              int* e_left, int* e_right, int* e_first, int* e_last, float4* e_ibox) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n - 1) ploc_finish_node(k, n, left, right, count, leftmost, pos, box, e_left, e_right, e_first, e_last, e_ibox);
}

// The PLOC hierarchy over the Morton-ordered triangles in ctx->d_tris / d_lbox; leaves d_left / d_right / d_first / d_last /
// d_ibox as k_emit_levels reads them and the triangles + leaf boxes in the tree's depth-first order.
static int ploc_build(vlb_ctx* ctx, uint32_t n_u, cudaStream_t st) {
    const int n = (int)n_u, B = 256;
    const size_t n_ref = 2 * (size_t)n - 1;
    int per_sm = 0;
    VLB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_ploc_rounds, 256, 0));
    int want_per_sm = 2;
    if (const char* v = getenv("VLB_PLOC_BLOCKS_PER_SM")) want_per_sm = std::max(1, atoi(v));
    const unsigned int grid_c = (unsigned int)std::max(1, std::min(per_sm, want_per_sm) * ctx->sm_count);
    // one scratch allocation, carved: box | C | Cn | nn | left | right | parent | count | leftmost | pos | block totals
    size_t off = 0;
    auto carve = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) & ~size_t(255); return o; };
    const size_t o_box = carve(2 * n_ref * sizeof(float4)), o_C = carve(n * sizeof(int)), o_Cn = carve(n * sizeof(int)), o_nn = carve(n * sizeof(int)),
                 o_left = carve(n * sizeof(int)), o_right = carve(n * sizeof(int)), o_parent = carve(n_ref * sizeof(int)),
                 o_count = carve(n_ref * sizeof(int)), o_lm = carve(n_ref * sizeof(int)), o_pos = carve(n * sizeof(int)),
                 o_tot = carve(grid_c * sizeof(int2));
    VLB_CUDA(ctx, ctx->d_ploc.reserve(off));
    VLB_CUDA(ctx, ctx->d_tris_alt.reserve(3ull * n * sizeof(float4)));
    VLB_CUDA(ctx, ctx->d_lbox_alt.reserve(2ull * n * sizeof(float4)));
    char* base = static_cast<char*>(ctx->d_ploc.p);
    float4* box = reinterpret_cast<float4*>(base + o_box);
    int* C = reinterpret_cast<int*>(base + o_C); int* Cn = reinterpret_cast<int*>(base + o_Cn); int* nn = reinterpret_cast<int*>(base + o_nn);
    int* left = reinterpret_cast<int*>(base + o_left); int* right = reinterpret_cast<int*>(base + o_right);
    int* parent = reinterpret_cast<int*>(base + o_parent); int* count = reinterpret_cast<int*>(base + o_count);
    int* leftmost = reinterpret_cast<int*>(base + o_lm); int* pos = reinterpret_cast<int*>(base + o_pos);
    int2* tot = reinterpret_cast<int2*>(base + o_tot);
    VLB_CUDA(ctx, cudaMemcpyAsync(box, ctx->d_lbox.p, 2ull * n * sizeof(float4), cudaMemcpyDeviceToDevice, st));
    k_ploc_init<<<(unsigned)((n_ref + B - 1) / B), B, 0, st>>>(n, C, leftmost, count, parent);
    VLB_LAUNCH_CHECK(ctx);
    int radius = std::max(1, std::min(kPlocMaxRadius, ctx->ploc_radius));
    int n_arg = n, tail = kPlocTail;
    if (const char* v = getenv("VLB_PLOC_TAIL")) tail = std::max(2, atoi(v));
    void* args[] = {&n_arg, &radius, &tail, &C, &Cn, &nn, &box, &left, &right, &parent, &count, &leftmost, &tot};
    VLB_CUDA(ctx, cudaLaunchCooperativeKernel(reinterpret_cast<void*>(k_ploc_rounds), dim3(grid_c), dim3(256), args, 0, st));
    VLB_LAUNCH_CHECK(ctx);
    const unsigned grid_n = (unsigned)((n + B - 1) / B);
    k_ploc_leaf_pos<<<grid_n, B, 0, st>>>(n, left, parent, count, pos);
    VLB_LAUNCH_CHECK(ctx);
    k_ploc_permute<<<grid_n, B, 0, st>>>(n, pos, ctx->d_tris.as<float4>(), ctx->d_lbox.as<float4>(), ctx->d_tris_alt.as<float4>(), ctx->d_lbox_alt.as<float4>());
    VLB_LAUNCH_CHECK(ctx);
    std::swap(ctx->d_tris, ctx->d_tris_alt);        // the permuted arrays become the BVH's triangles and leaf boxes
    std::swap(ctx->d_lbox, ctx->d_lbox_alt);
    k_ploc_finish<<<grid_n, B, 0, st>>>(n, left, right, count, leftmost, pos, box, ctx->d_left.as<int>(), ctx->d_right.as<int>(),
                                        ctx->d_first.as<int>(), ctx->d_last.as<int>(), ctx->d_ibox.as<float4>());
    VLB_LAUNCH_CHECK(ctx);
    return VLB_OK;
}

// n == 1: a single node whose only child is the one-triangle leaf.
__global__ void k_emit_single(const float4* __restrict__ lbox, const float* __restrict__ scratch, float4* nodes) {
    const float ext = fmaxf(scratch[3] - scratch[0], fmaxf(scratch[4] - scratch[1], scratch[5] - scratch[2]));
    emit_wide_single(lbox, ext * 1e-6f, nodes);
}

int bvh_build(vlb_ctx* ctx, vlb_bvh_stats* stats) {
    const uint32_t n = (uint32_t)ctx->n_tris;
    cudaStream_t st = ctx->stream;
    float sort_ms = 0.f, build_ms = 0.f;
    if (const char* v = getenv("VLB_BVH_MAX_LEAF")) ctx->max_leaf = std::max(1, std::min(kMaxLeaf, atoi(v)));
    VLB_CUDA(ctx, ctx->d_scratch.reserve(64 * sizeof(float)));
    VLB_CUDA(ctx, cudaEventRecord(ctx->ev[0], st));
    k_init_bounds<<<1, 32, 0, st>>>(ctx->d_scratch.as<float>());
    VLB_LAUNCH_CHECK(ctx);
    ctx->n_nodes = 0;
    if (n > 0) {
        const int B = 256;
        const unsigned grid_n = (n + B - 1) / B;
        VLB_CUDA(ctx, ctx->d_tris.reserve(3ull * n * sizeof(float4)));
        VLB_CUDA(ctx, ctx->d_nodes.reserve((size_t)kNodeQuads * std::max<uint32_t>(n - 1, 1) * sizeof(float4)));
        VLB_CUDA(ctx, ctx->d_keys.reserve(n * sizeof(uint64_t)));
        VLB_CUDA(ctx, ctx->d_keys_sorted.reserve(n * sizeof(uint64_t)));
        VLB_CUDA(ctx, ctx->d_vals.reserve(n * sizeof(uint32_t)));
        VLB_CUDA(ctx, ctx->d_vals_sorted.reserve(n * sizeof(uint32_t)));
        VLB_CUDA(ctx, ctx->d_lbox.reserve(2ull * n * sizeof(float4)));
        for (DevBuf* b : {&ctx->d_left, &ctx->d_right, &ctx->d_first, &ctx->d_last, &ctx->d_parent_i, &ctx->d_parent_l, &ctx->d_flags})
            VLB_CUDA(ctx, b->reserve(n * sizeof(int)));
        VLB_CUDA(ctx, ctx->d_ibox.reserve(2ull * n * sizeof(float4)));

        const unsigned red_grid = std::min<unsigned>(grid_n, (unsigned)ctx->sm_count * 4u);
        k_scene_bounds<<<red_grid, B, 0, st>>>(ctx->d_tri_flat.as<float4>(), n, ctx->d_scratch.as<float>());
        VLB_LAUNCH_CHECK(ctx);
        const char* cubic_env = getenv("VLB_BVH_CUBIC");
        k_morton<<<grid_n, B, 0, st>>>(ctx->d_tri_flat.as<float4>(), n, ctx->d_scratch.as<float>(), cubic_env && *cubic_env ? atoi(cubic_env) : 0,
                                      ctx->d_keys.as<uint64_t>(), ctx->d_vals.as<uint32_t>());
        VLB_LAUNCH_CHECK(ctx);

        VLB_CUDA(ctx, ctx->d_sort_tmp.reserve(radix_sort_scratch_bytes(n)));
        VLB_CUDA(ctx, cudaEventRecord(ctx->ev[2], st));
        const int where = radix_sort_pairs(ctx->d_keys.as<uint64_t>(), ctx->d_vals.as<uint32_t>(), ctx->d_keys_sorted.as<uint64_t>(),
                                           ctx->d_vals_sorted.as<uint32_t>(), n, 8, ctx->d_sort_tmp.p, st, nullptr);
        ctx->launches += 3 * 8 - 1;   // three kernels per pass; the check below counts the last one
        VLB_LAUNCH_CHECK(ctx);
        const uint64_t* keys_sorted = where ? ctx->d_keys_sorted.as<uint64_t>() : ctx->d_keys.as<uint64_t>();
        const uint32_t* vals_sorted = where ? ctx->d_vals_sorted.as<uint32_t>() : ctx->d_vals.as<uint32_t>();
        VLB_CUDA(ctx, cudaEventRecord(ctx->ev[3], st));

        k_gather<<<grid_n, B, 0, st>>>(ctx->d_tri_flat.as<float4>(), vals_sorted, n,
                                      ctx->d_tris.as<float4>(), ctx->d_lbox.as<float4>());
        VLB_LAUNCH_CHECK(ctx);
        if (n == 1) {
            k_emit_single<<<1, 1, 0, st>>>(ctx->d_lbox.as<float4>(), ctx->d_scratch.as<float>(), ctx->d_nodes.as<float4>());
            VLB_LAUNCH_CHECK(ctx);
            ctx->n_nodes = 1;
        } else {
            int builder = ctx->bvh_builder;
            if (const char* v = getenv("VLB_BVH_BUILDER")) builder = (*v == 'p' || *v == 'P' || *v == '1') ? VLB_BVH_BUILDER_PLOC : VLB_BVH_BUILDER_LBVH;
            if (const char* v = getenv("VLB_PLOC_RADIUS")) ctx->ploc_radius = atoi(v);
            if (builder == VLB_BVH_BUILDER_PLOC) {
                if (int r = ploc_build(ctx, n, st)) return r;
            } else {
                VLB_CUDA(ctx, cudaMemsetAsync(ctx->d_flags.p, 0, n * sizeof(int), st));
                k_karras<<<grid_n, B, 0, st>>>(keys_sorted, (int)n, ctx->d_left.as<int>(), ctx->d_right.as<int>(),
                                              ctx->d_first.as<int>(), ctx->d_last.as<int>(), ctx->d_parent_i.as<int>(), ctx->d_parent_l.as<int>());
                VLB_LAUNCH_CHECK(ctx);
                k_refit<<<grid_n, B, 0, st>>>((int)n, ctx->d_left.as<int>(), ctx->d_right.as<int>(), ctx->d_parent_i.as<int>(),
                                             ctx->d_parent_l.as<int>(), ctx->d_lbox.as<float4>(), ctx->d_ibox.as<float4>(), ctx->d_flags.as<int>());
                VLB_LAUNCH_CHECK(ctx);
            }
            // top-down collapse into 4-wide nodes: one cooperative launch (grid.sync per level of the wide tree)
            VLB_CUDA(ctx, ctx->d_frontier[0].reserve(n * sizeof(int)));
            VLB_CUDA(ctx, ctx->d_frontier[1].reserve(n * sizeof(int)));
            VLB_CUDA(ctx, ctx->d_frontier_n.reserve(4 * sizeof(unsigned int)));
            const unsigned int cnt_init[4] = {1u, 0u, 0u, 0u};
            VLB_CUDA(ctx, cudaMemcpyAsync(ctx->d_frontier_n.p, cnt_init, sizeof cnt_init, cudaMemcpyHostToDevice, st));   // pageable: staged at once
            VLB_CUDA(ctx, cudaMemsetAsync(ctx->d_frontier[0].p, 0, sizeof(int), st));       // frontier 0 = {root}
            int per_sm = 0;
            VLB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_emit_levels, 256, 0));
            const unsigned int grid_c = (unsigned int)std::max(1, std::min(per_sm, 4) * ctx->sm_count);
            int* fa = ctx->d_frontier[0].as<int>(); int* fb = ctx->d_frontier[1].as<int>();
            unsigned int* cnt = ctx->d_frontier_n.as<unsigned int>();
            const int* a_left = ctx->d_left.as<int>(); const int* a_right = ctx->d_right.as<int>();
            const int* a_first = ctx->d_first.as<int>(); const int* a_last = ctx->d_last.as<int>();
            const float4* a_ibox = ctx->d_ibox.as<float4>(); const float4* a_lbox = ctx->d_lbox.as<float4>();
            int a_max_leaf = ctx->max_leaf;
            const float* a_scratch = ctx->d_scratch.as<float>(); float4* a_nodes = ctx->d_nodes.as<float4>();
            unsigned int* a_emitted = reinterpret_cast<unsigned int*>(ctx->d_scratch.as<float>() + 12);
            void* args[] = {&fa, &fb, &cnt, &a_left, &a_right, &a_first, &a_last, &a_ibox, &a_lbox, &a_max_leaf, &a_scratch, &a_nodes, &a_emitted};
            VLB_CUDA(ctx, cudaLaunchCooperativeKernel(reinterpret_cast<void*>(k_emit_levels), dim3(grid_c), dim3(256), args, 0, st));
            VLB_LAUNCH_CHECK(ctx);
        }
    }
    VLB_CUDA(ctx, cudaEventRecord(ctx->ev[1], st));
    float h[13];
    VLB_CUDA(ctx, cudaMemcpyAsync(h, ctx->d_scratch.p, sizeof h, cudaMemcpyDeviceToHost, st));
    VLB_CUDA(ctx, cudaStreamSynchronize(st));
    if (n > 1) { unsigned int ne; memcpy(&ne, &h[12], 4); ctx->n_nodes = ne; }
    VLB_CUDA(ctx, cudaEventElapsedTime(&build_ms, ctx->ev[0], ctx->ev[1]));
    if (n > 0) VLB_CUDA(ctx, cudaEventElapsedTime(&sort_ms, ctx->ev[2], ctx->ev[3]));
    for (int k = 0; k < 6; ++k) ctx->tight_bounds[k] = n ? h[k] : 0.f;
    ctx->have_tight_bounds = true;
    ctx->have_bvh = true;
    if (stats) {
        stats->n_triangles = n; stats->n_nodes = ctx->n_nodes; stats->max_leaf_size = (uint32_t)ctx->max_leaf;
        stats->reserved = 0;
        for (int k = 0; k < 6; ++k) stats->bounds[k] = ctx->tight_bounds[k];
        stats->build_ms = build_ms; stats->sort_ms = sort_ms;
    }
    return VLB_OK;
}

}  // namespace vlb

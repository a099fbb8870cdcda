// bake.cu — the probe bake: replaces the reference's per-probe pair
//   EnvMapGenerator::getMap  (src/baker/env_map_generator.cpp:325-359; shaders env_map.rgen/.rchit,
//                             main.rmiss, shadow.rmiss)          and
//   LightBaker::dispatchBakingKernel (src/baker/light_baker.cpp:269-285; shaders/sh.comp)
// with ONE persistent launch over all probes of the slab: every warp pulls (probe, direction
// chunk) items from a global counter, fires the probe's fixed equirect directions plus shadow
// rays through the software LBVH, and projects the radiance onto SH in registers. Per-ray
// radiance never goes to HBM; the only global writes are 192 bytes per probe.
#include <algorithm>

#include "vlb_context.h"
#include "vlb_shade.cuh"
#include "vlb_warp.cuh"

namespace vlb {

constexpr int kBakeBlock = 128;
// One warp-sized direction tile = 2^lw x 2^(5-lw) adjacent texels of the equirect direction grid. The shape is picked
// per direction grid so that the tile is as square as possible in ANGLE (a texel spans 360/W x 180/H degrees): 4x8
// texels for W = H (the 32x32 / 64x64 grids of the BASELINE configs; measured +1.4 % over 8x4 on C3), 8x4 for
// W >= 2H (the reference's 3141x1000). Rays in flight in a warp then cover the smallest solid angle.
__host__ __device__ inline int tile_x(int tile, int w, int tiles_x, int lw) { return ((tile % tiles_x) << lw) + (w & ((1 << lw) - 1)); }
__host__ __device__ inline int tile_y(int tile, int w, int tiles_x, int lw) { return ((tile / tiles_x) << (5 - lw)) + (w >> lw); }

struct WarpQueues;
struct BakeParams {
    BvhView bvh;
    ShadeView shade;
    BakeConsts c;
    const float* px; const float* py; const float* pz;   // probe axis coordinates
    const float2* row_sc;                                // (sin, cos) theta per direction row
    const float2* col_cs;                                // (cos, sin) phi per direction column
    int Nx, Ny, Nz, k0, kstride;   // the call bakes z-slices k0, k0 + kstride, ...
    int W, H, tiles_x, n_tiles, tile_lw;   // tile_lw: log2 of the direction tile's width
    int chunks, tiles_per_chunk;
    uint32_t n_items;
    float pixel_area;
    float* out;                  // chunks == 1: final [slot][48]; else partials [item][48]
    unsigned int* work_counter;
    unsigned long long* stats;   // [0] shadow rays, [1] nodes visited, [2] triangles tested
    int ref_order, world_frame;
    WarpQueues* stream_scratch;  // k_bake_stream: per-warp radiance tile + ray queues, [grid * warps per block]
    int node_min;                // k_bake_stream: the node loop yields to the leaf phase below this many lanes
    GatherView g;                // gather pass source (g.prev == NULL: direct pass)
};

__device__ __forceinline__ size_t out_slot(const BakeParams& p, uint32_t q) {
    if (!p.ref_order) return q;
    const int i = q % p.Nx, j = (q / p.Nx) % p.Ny, k = q / (p.Nx * p.Ny);
    // writer order of LightBaker::probePositionsFromBoudingBox (light_baker.cpp:87-98)
    const size_t s = (j == 0) ? (size_t)i : (size_t)p.Nx + (size_t)i * (p.Ny - 1) + (j - 1);
    return (k == 0) ? s : (size_t)p.Nx * p.Ny + s * (p.Nz - 1) + (k - 1);
}

template <int K, bool COUNT>
__global__ void __launch_bounds__(kBakeBlock) k_bake(const BakeParams p) {
    constexpr int V = (K * 3 <= 32) ? 32 : 64;
    const int lane = threadIdx.x & 31;
    TraceCounters cnt; cnt.nodes = 0; cnt.tris = 0;
    uint32_t shadow = 0;
    for (;;) {
        unsigned item = 0;
        if (lane == 0) item = atomicAdd(p.work_counter, 1u);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= p.n_items) break;
        const uint32_t q = item / (uint32_t)p.chunks, chunk = item % (uint32_t)p.chunks;
        const uint32_t nxy = (uint32_t)(p.Nx * p.Ny), in_slice = q % nxy;
        const Vec3 o = mk3(p.px[in_slice % p.Nx], p.py[in_slice / p.Nx], p.pz[p.k0 + (q / nxy) * p.kstride]);
        float acc[V];
#pragma unroll
        for (int i = 0; i < V; ++i) acc[i] = 0.f;
        const int t0 = chunk * p.tiles_per_chunk;
        const int t1 = min(t0 + p.tiles_per_chunk, p.n_tiles);
        for (int tile = t0; tile < t1; ++tile) {
            const int x = tile_x(tile, lane, p.tiles_x, p.tile_lw);
            const int y = tile_y(tile, lane, p.tiles_x, p.tile_lw);
            if (x < p.W && y < p.H) {
                const float2 row = __ldg(p.row_sc + y), col = __ldg(p.col_cs + x);
                const Vec3 t = to_vector_sc(row.x, row.y, col.x, col.y);   // sh_common.h:8-12
                const Vec3 r = mk3(t.x, t.z, t.y);                         // env_map.rgen:21 .xzy
                float rgb[3];
                probe_ray_radiance<COUNT, K>(p.bvh, p.shade, p.c, o, r, rgb, &cnt, &shadow, &p.g);
                const float w = p.pixel_area * row.x;                      // sh.comp:32-33
                float b[K];
                sh_basis<K>(p.world_frame ? r : t, b);                     // sh.comp:30,39
#pragma unroll
                for (int i = 0; i < K; ++i) {
                    const float bw = b[i] * w;
                    acc[3 * i + 0] = fmaf(bw, rgb[0], acc[3 * i + 0]);
                    acc[3 * i + 1] = fmaf(bw, rgb[1], acc[3 * i + 1]);
                    acc[3 * i + 2] = fmaf(bw, rgb[2], acc[3 * i + 2]);
                }
            }
            __syncwarp();
        }
        warp_transpose_reduce<V>(acc, lane);
        float* dst = p.out + (p.chunks == 1 ? out_slot(p, q) : (size_t)item) * VLB_SH_STRIDE;
        if (V == 32) {
            if (lane < K * 3) dst[lane] = acc[0];
            if (lane + 32 < VLB_SH_STRIDE) dst[lane + 32] = 0.f;
            if (lane >= K * 3) dst[lane] = 0.f;
        } else {
            if (2 * lane < VLB_SH_STRIDE) {
                dst[2 * lane] = acc[0];
                dst[2 * lane + 1] = acc[1];
            }
        }
    }
    // statistics: one atomic per warp
    unsigned long long s = shadow, nn = cnt.nodes, nt = cnt.tris;
    for (int off = 16; off > 0; off >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, off);
        if (COUNT) { nn += __shfl_xor_sync(0xffffffffu, nn, off); nt += __shfl_xor_sync(0xffffffffu, nt, off); }
    }
    if (lane == 0) {
        atomicAdd(p.stats + 0, s);
        if (COUNT) { atomicAdd(p.stats + 1, nn); atomicAdd(p.stats + 2, nt); }
    }
}

// =========================================================================================
// k_bake_stream — the production bake kernel: a warp is a small wavefront path tracer.
//
// A warp owns a work item = (probe, run of 256-direction chunks). Inside a chunk every LANE is a ray
// slot: an idle lane first takes a queued shadow ray, otherwise the next direction of the chunk
// (ballot + popc ranks; directions in tile order (32-texel tiles, square in angle: tile_x / tile_y) so the rays in flight stay angularly
// adjacent), and traverses the LBVH. Primary (closest-hit) and shadow (any-hit) rays share one
// "while-while" loop: all lanes step through internal nodes until most of them hold a leaf, then
// the leaves are intersected. A finished primary ray only pushes its hit record into a per-warp
// queue and frees its lane; once 32 records are queued the whole warp shades them
// together (env_map.rchit / main.rmiss arithmetic at full SIMD width) and pushes the shadow rays
// that are needed into a second queue. A finished shadow ray selects the lit or dark radiance
// computed at shading time. Radiance goes to a shared-memory tile indexed by direction; when the
// chunk is drained the warp projects its 256 radiances onto SH cooperatively (lane l takes
// directions l, l+32, ... in a fixed order, so the result is bitwise reproducible whatever the
// run-time ray scheduling was), reduces with shuffles and keeps one or two running coefficients per
// lane. Radiance never leaves the SM; HBM sees 192 bytes per item.
// The per-ray arithmetic is exactly that of probe_ray_radiance (vlb_shade.cuh).
// =========================================================================================
constexpr int kChunkTiles = 8;
constexpr int kChunkDirs = kChunkTiles * 32;
constexpr int kRayDone = kNoChild;          // traversal finished
constexpr int kHitCap = 64;                 // queued hit records per warp (<= 31 waiting + 32 arriving)
constexpr int kShadowCap = 64;              // queued shadow rays per warp (shading needs 32 free slots)
constexpr int kStreamWarps = kBakeBlock / 32;

struct WarpQueues {
    float rad[3][kChunkDirs];      // radiance of the current chunk, by direction
    // hit queue (SoA): flat triangle id (-1: miss), t, u, v, direction index
    int hq_id[kHitCap]; float hq_t[kHitCap], hq_u[kHitCap], hq_v[kHitCap]; int hq_dir[kHitCap];
    // shadow-ray queue: origin, unit direction, length, direction index, radiance if lit / if occluded
    float sq_o[3][kShadowCap], sq_d[3][kShadowCap], sq_len[kShadowCap]; int sq_dir[kShadowCap];
    float sq_rgb[6][kShadowCap];
    float lane_rgb[6][32];         // the same two radiances for the shadow ray a lane is tracing
};

#ifndef VLB_BAKE_MIN_BLOCKS
#define VLB_BAKE_MIN_BLOCKS 8      // resident 128-thread blocks per SM the register allocation is held to (8 -> 64 registers)
#endif
template <int K, bool COUNT, bool GATHER, bool TEX>
__global__ void __launch_bounds__(kBakeBlock, GATHER ? 6 : VLB_BAKE_MIN_BLOCKS) k_bake_stream(const BakeParams p) {
    constexpr int V = (K * 3 <= 32) ? 32 : 64;
    // The per-warp queues live in a global scratch buffer (L1/L2 resident, ~9 KB per resident warp),
    // not in shared memory: measured on B200, leaving the SM's 228 KB to the L1 cache (BVH nodes)
    // and fitting 8 blocks per SM is worth more than shared-memory latency for the queue traffic
    // (about 100 bytes per ray against ~1.5 KB of node and triangle reads).
    WarpQueues* s_all = p.stream_scratch + (size_t)blockIdx.x * kStreamWarps;
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    WarpQueues& S = s_all[warp];
    const BvhView& bvh = p.bvh;
    const bool want_shadow = (p.c.flags & 1u) != 0;
    TraceCounters cnt; cnt.nodes = 0; cnt.tris = 0;
    uint32_t shadow = 0;
    int stack[kStackSize];

    for (;;) {
        unsigned item = 0;
        if (lane == 0) item = atomicAdd(p.work_counter, 1u);
        item = __shfl_sync(full, item, 0);
        if (item >= p.n_items) break;
        const uint32_t q = item / (uint32_t)p.chunks, part = item % (uint32_t)p.chunks;
        const uint32_t nxy = (uint32_t)(p.Nx * p.Ny), in_slice = q % nxy;
        const Vec3 po = mk3(p.px[in_slice % p.Nx], p.py[in_slice / p.Nx], p.pz[p.k0 + (q / nxy) * p.kstride]);
        const int tile_begin = part * p.tiles_per_chunk;
        const int tile_end = min(tile_begin + p.tiles_per_chunk, p.n_tiles);
        float coef0 = 0.f, coef1 = 0.f;   // running sums of coefficient `lane` (V=32) / 2*lane, 2*lane+1 (V=64)

        for (int base_tile = tile_begin; base_tile < tile_end; base_tile += kChunkTiles) {
            const int n_dirs = min(kChunkTiles, tile_end - base_tile) * 32;
            // ------------------------------ trace the chunk ------------------------------
            int next = 0, n_hit = 0, n_sh = 0;   // warp-uniform: directions handed out, queue fills
            bool busy = false;
            int kind = 0;                        // 0 primary (closest hit), 1 shadow (any hit)
            int my_dir = 0, cur = kRayDone, sp = 0;
            Vec3 ro = po, rd = po, idir = po, ood = po;
            float tmin = 0.f, tcull = 0.f;
            HitRec best; best.id = -1; best.t = 0.f; best.u = 0.f; best.v = 0.f;
            for (;;) {
                // ---- 1. refill idle lanes: queued shadow rays first, then new directions ----
                const unsigned idle = __ballot_sync(full, !busy);
                if (idle != 0u && (n_sh > 0 || next < n_dirs)) {
                    const int n_idle = __popc(idle), rank = __popc(idle & lt_mask);
                    const int take_sh = min(n_idle, n_sh);
                    const int take_new = min(n_idle - take_sh, n_dirs - next);
                    if (!busy && rank < take_sh) {
                        const int e = n_sh - 1 - rank;
                        ro = mk3(S.sq_o[0][e], S.sq_o[1][e], S.sq_o[2][e]);
                        rd = mk3(S.sq_d[0][e], S.sq_d[1][e], S.sq_d[2][e]);
                        tmin = 0.0f; tcull = S.sq_len[e];                               // env_map.rchit:87
                        my_dir = S.sq_dir[e];
#pragma unroll
                        for (int k = 0; k < 6; ++k) S.lane_rgb[k][lane] = S.sq_rgb[k][e];
                        kind = 1; busy = true;
                    } else if (!busy && rank - take_sh < take_new) {
                        const int cand = next + rank - take_sh;
                        const int tile = base_tile + (cand >> 5), w = cand & 31;
                        const int x = tile_x(tile, w, p.tiles_x, p.tile_lw);
                        const int y = tile_y(tile, w, p.tiles_x, p.tile_lw);
                        if (x < p.W && y < p.H) {
                            const float2 row = __ldg(p.row_sc + y), col = __ldg(p.col_cs + x);
                            const Vec3 t = to_vector_sc(row.x, row.y, col.x, col.y);   // sh_common.h:8-12
                            rd = mk3(t.x, t.z, t.y);                                   // env_map.rgen:21 .xzy
                            ro = po;
                            tmin = p.c.tmin; tcull = p.c.tmax;
                            my_dir = cand; kind = 0; busy = true;
                        } else {
                            S.rad[0][cand] = 0.f; S.rad[1][cand] = 0.f; S.rad[2][cand] = 0.f;   // outside the direction grid
                        }
                    }
                    if (busy && cur == kRayDone) {     // freshly started ray
                        idir = mk3(safe_inv(rd.x), safe_inv(rd.y), safe_inv(rd.z));
                        ood = mk3(ro.x * idir.x, ro.y * idir.y, ro.z * idir.z);
                        best.id = -1; best.t = tcull; best.u = 0.f; best.v = 0.f;
                        sp = 0;
                        cur = bvh.n_tris ? 0 : kRayDone;
                    }
                    n_sh -= take_sh;
                    next += take_new;
                    __syncwarp();
                }
                const unsigned running = __ballot_sync(full, busy);
                // ---- 2. shade queued hits: a full warp of them, or whatever is left when nothing runs ----
                if ((n_hit >= 32 && n_sh <= kShadowCap - 32) || (running == 0u && n_hit > 0)) {
                    const int take = min(n_hit, 32);
                    const int e = n_hit - take + lane;
                    bool push = false;
                    float lit_rgb[3] = {0.f, 0.f, 0.f}, dark_rgb[3] = {0.f, 0.f, 0.f};
                    ShadePrelude pre;
                    int dir = 0;
                    if (lane < take) {
                        HitRec h; h.id = S.hq_id[e]; h.t = S.hq_t[e]; h.u = S.hq_u[e]; h.v = S.hq_v[e];
                        dir = S.hq_dir[e];
                        const int tile = base_tile + (dir >> 5), w = dir & 31;
                        const int x = tile_x(tile, w, p.tiles_x, p.tile_lw);
                        const int y = tile_y(tile, w, p.tiles_x, p.tile_lw);
                        const float2 row = __ldg(p.row_sc + y), col = __ldg(p.col_cs + x);
                        const Vec3 t = to_vector_sc(row.x, row.y, col.x, col.y);
                        const Vec3 r = mk3(t.x, t.z, t.y);
                        float rgb[3] = {0.f, 0.f, 0.f};                                 // env_map.rgen:25
                        if (h.id >= 0) {
                            const bool lit = shade_prelude<TEX>(p.shade, p.c, h, po, r, pre);
                            float ind[3] = {0.f, 0.f, 0.f};
                            // gather passes: the 8 short visibility rays to the surrounding probes are traced
                            // right here, per lane (main.rchit:143-163); only the sun shadow ray is queued
                            if (GATHER) gather_indirect<K, COUNT>(bvh, p.g, pre, ind, &cnt);
                            if (lit && want_shadow) {
                                // radiance for both outcomes now, the shadow ray decides (env_map.rchit:83-99)
                                shade_finish(p.c, pre, r, false, ind, lit_rgb);
                                shade_finish(p.c, pre, r, true, ind, dark_rgb);
                                push = true;
                            } else {
                                shade_finish(p.c, pre, r, !lit, ind, rgb);
                            }
                        } else if ((p.c.flags & 2u) && p.shade.sky) {                   // VLB_BAKE_SKYBOX_ON_MISS
                            sky_lookup(p.shade, r, rgb);
                            if (p.c.flags & 4u) { rgb[0] = srgb_encode(rgb[0]); rgb[1] = srgb_encode(rgb[1]); rgb[2] = srgb_encode(rgb[2]); }
                        }
                        if (!push) {
                            if (p.c.flags & 8u) { rgb[0] = quant8(rgb[0]); rgb[1] = quant8(rgb[1]); rgb[2] = quant8(rgb[2]); }
                            S.rad[0][dir] = rgb[0]; S.rad[1][dir] = rgb[1]; S.rad[2][dir] = rgb[2];
                        }
                    }
                    const unsigned pm = __ballot_sync(full, push);
                    if (push) {
                        const int d = n_sh + __popc(pm & lt_mask);
                        S.sq_o[0][d] = pre.so.x; S.sq_o[1][d] = pre.so.y; S.sq_o[2][d] = pre.so.z;   // env_map.rchit:82
                        S.sq_d[0][d] = pre.Ln.x; S.sq_d[1][d] = pre.Ln.y; S.sq_d[2][d] = pre.Ln.z;
                        S.sq_len[d] = pre.llen; S.sq_dir[d] = dir;
#pragma unroll
                        for (int k = 0; k < 3; ++k) { S.sq_rgb[k][d] = lit_rgb[k]; S.sq_rgb[3 + k][d] = dark_rgb[k]; }
                        ++shadow;
                    }
                    n_sh += __popc(pm);
                    n_hit -= take;
                    __syncwarp();
                    continue;
                }
                if (running == 0u) {
                    if (n_sh == 0 && next >= n_dirs) break;
                    continue;
                }
                // ---- 3. internal nodes: lanes step until most of them hold a leaf ----
                for (;;) {
                    const bool at_node = busy && cur >= 0;
                    const unsigned nm = __ballot_sync(full, at_node);
                    if (nm == 0u || (__popc(nm) < p.node_min && __popc(running) - __popc(nm) >= p.node_min)) break;
                    if (at_node) {
                        if (COUNT) cnt.nodes++;
                        cur = bvh4_step<true>(bvh, cur, idir, ood, tmin, tcull, stack, sp);   // one code path for both ray kinds
                    }
                }
                // ---- 4. leaves ----
                if (busy && cur < 0 && cur != kRayDone) {
                    const bool terminated = kind ? leaf_step<true, COUNT>(bvh, cur, ro, rd, tmin, tcull, best, &cnt)
                                                 : leaf_step<false, COUNT>(bvh, cur, ro, rd, tmin, tcull, best, &cnt);
                    cur = (terminated || sp == 0) ? kRayDone : stack[--sp];
                }
                // ---- 5. finished rays free their lane ----
                const bool fin = busy && cur == kRayDone;
                const unsigned fp = __ballot_sync(full, fin && kind == 0);
                if (fin) {
                    if (kind == 0) {
                        const int e = n_hit + __popc(fp & lt_mask);
                        S.hq_id[e] = best.id; S.hq_t[e] = best.t; S.hq_u[e] = best.u; S.hq_v[e] = best.v; S.hq_dir[e] = my_dir;
                    } else {
                        const int o3 = best.id >= 0 ? 3 : 0;                            // occluded -> dark
                        float rgb[3] = {S.lane_rgb[o3 + 0][lane], S.lane_rgb[o3 + 1][lane], S.lane_rgb[o3 + 2][lane]};
                        if (p.c.flags & 8u) { rgb[0] = quant8(rgb[0]); rgb[1] = quant8(rgb[1]); rgb[2] = quant8(rgb[2]); }  // QUANTIZE_RGBA8
                        S.rad[0][my_dir] = rgb[0]; S.rad[1][my_dir] = rgb[1]; S.rad[2][my_dir] = rgb[2];
                    }
                    busy = false;
                }
                n_hit += __popc(fp);
                if (n_hit > kHitCap) __trap();     // invariant: <= 31 waiting + 32 arriving
                __syncwarp();
            }
            __syncwarp();
            // ------------------------------ project the chunk ------------------------------
            float acc[V];
#pragma unroll
            for (int i = 0; i < V; ++i) acc[i] = 0.f;
            for (int tt = 0; tt * 32 < n_dirs; ++tt) {
                const int tile = base_tile + tt;
                const int x = tile_x(tile, lane, p.tiles_x, p.tile_lw);
                const int y = tile_y(tile, lane, p.tiles_x, p.tile_lw);
                if (x < p.W && y < p.H) {
                    const float2 row = __ldg(p.row_sc + y), col = __ldg(p.col_cs + x);
                    const Vec3 t = to_vector_sc(row.x, row.y, col.x, col.y);
                    const float w = p.pixel_area * row.x;                               // sh.comp:32-33
                    float b[K];
                    sh_basis<K>(p.world_frame ? mk3(t.x, t.z, t.y) : t, b);             // sh.comp:30,39
                    const float r0 = S.rad[0][tt * 32 + lane], r1 = S.rad[1][tt * 32 + lane], r2 = S.rad[2][tt * 32 + lane];
#pragma unroll
                    for (int i = 0; i < K; ++i) {
                        const float bw = b[i] * w;
                        acc[3 * i + 0] = fmaf(bw, r0, acc[3 * i + 0]);
                        acc[3 * i + 1] = fmaf(bw, r1, acc[3 * i + 1]);
                        acc[3 * i + 2] = fmaf(bw, r2, acc[3 * i + 2]);
                    }
                }
            }
            warp_transpose_reduce<V>(acc, lane);
            coef0 += acc[0];
            if (V == 64) coef1 += acc[1];
            __syncwarp();
        }
        float* dst = p.out + (p.chunks == 1 ? out_slot(p, q) : (size_t)item) * VLB_SH_STRIDE;
        if (V == 32) {
            dst[lane] = lane < K * 3 ? coef0 : 0.f;
            if (lane + 32 < VLB_SH_STRIDE) dst[lane + 32] = 0.f;
        } else {
            if (2 * lane < VLB_SH_STRIDE) { dst[2 * lane] = coef0; dst[2 * lane + 1] = coef1; }
        }
    }
    // statistics: one atomic per warp
    unsigned long long s = shadow, nn = cnt.nodes, nt = cnt.tris;
    for (int off = 16; off > 0; off >>= 1) {
        s += __shfl_xor_sync(full, s, off);
        if (COUNT) { nn += __shfl_xor_sync(full, nn, off); nt += __shfl_xor_sync(full, nt, off); }
    }
    if (lane == 0) {
        atomicAdd(p.stats + 0, s);
        if (COUNT) { atomicAdd(p.stats + 1, nn); atomicAdd(p.stats + 2, nt); }
    }
}

// chunks > 1: per-probe sum of the chunk partials in chunk order (fixed order => deterministic).
__global__ void k_sum_partials(const BakeParams p, const float* __restrict__ partials, uint32_t n_probes, float* out) {
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_probes * VLB_SH_STRIDE) return;
    const uint32_t q = idx / VLB_SH_STRIDE, c = idx % VLB_SH_STRIDE;
    float s = 0.f;
    for (int k = 0; k < p.chunks; ++k) s += partials[((size_t)q * p.chunks + k) * VLB_SH_STRIDE + c];
    out[out_slot(p, q) * VLB_SH_STRIDE + c] = s;
}

// VLB_BAKE_ACCUMULATE_ACROSS_PROBES: the reference never re-zeroes its SSBO (light_baker.cpp:110-121),
// so probe i holds the running total of probes 0..i in bake order. Forensic mode only.
__global__ void k_running_total(float* out, uint32_t n_probes) {
    const int c = threadIdx.x;
    if (c >= VLB_SH_STRIDE) return;
    float s = 0.f;
    for (uint32_t q = 0; q < n_probes; ++q) { s += out[(size_t)q * VLB_SH_STRIDE + c]; out[(size_t)q * VLB_SH_STRIDE + c] = s; }
}

static int env_flag(const char* name, int dflt) {
    const char* v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

int bake_device(vlb_ctx* ctx, const vlb_bake_settings* s, const float* d_prev_full, float* d_out) {
    cudaStream_t st = ctx->stream;
    const int Nx = s->probes[0], Ny = s->probes[1], Nz = s->probes[2];
    const int k0 = s->slab_k1 < 0 ? 0 : s->slab_k0, k1 = s->slab_k1 < 0 ? Nz : s->slab_k1;
    const int kstride = s->slab_stride > 1 ? s->slab_stride : 1;
    const uint64_t n_probes = (uint64_t)Nx * Ny * (uint64_t)((k1 - k0 + kstride - 1) / kstride);
    if (int r = bake_collect_stats(ctx)) return r;     // a previous asynchronous bake still owns the stats buffers
    if (int r = sky_upload_join(ctx)) return r;        // vlb_skybox_set_async: the bake samples the new texels
    ctx->last_bake = vlb_bake_stats{};
    if (n_probes == 0) return VLB_OK;
    const int W = s->dir_w, H = s->dir_h;

    // tables: probe axis coordinates and the equirect sin/cos tables (host_tables.cpp)
    std::vector<float> axis((size_t)Nx + Ny + Nz), row(2 * (size_t)H), col(2 * (size_t)W);
    host_axis_coords(s->origin[0], s->step[0], Nx, axis.data());
    host_axis_coords(s->origin[1], s->step[1], Ny, axis.data() + Nx);
    host_axis_coords(s->origin[2], s->step[2], Nz, axis.data() + Nx + Ny);
    if (ctx->dir_w != W || ctx->dir_h != H) {
        host_dir_tables(W, H, 0.f, row.data(), col.data());
        VLB_CUDA(ctx, ctx->d_row_sc.reserve(row.size() * sizeof(float)));
        VLB_CUDA(ctx, ctx->d_col_sc.reserve(col.size() * sizeof(float)));
        VLB_CUDA(ctx, cudaMemcpyAsync(ctx->d_row_sc.p, row.data(), row.size() * sizeof(float), cudaMemcpyHostToDevice, st));
        VLB_CUDA(ctx, cudaMemcpyAsync(ctx->d_col_sc.p, col.data(), col.size() * sizeof(float), cudaMemcpyHostToDevice, st));
        VLB_CUDA(ctx, cudaStreamSynchronize(st));
        ctx->dir_w = W; ctx->dir_h = H;
    }
    VLB_CUDA(ctx, ctx->d_axis.reserve(axis.size() * sizeof(float)));
    VLB_CUDA(ctx, ctx->d_work_counter.reserve(16));
    VLB_CUDA(ctx, ctx->d_stats.reserve(4 * sizeof(unsigned long long)));
    VLB_CUDA(ctx, cudaMemcpyAsync(ctx->d_axis.p, axis.data(), axis.size() * sizeof(float), cudaMemcpyHostToDevice, st));
    VLB_CUDA(ctx, cudaMemsetAsync(ctx->d_work_counter.p, 0, 16, st));
    VLB_CUDA(ctx, cudaMemsetAsync(ctx->d_stats.p, 0, 4 * sizeof(unsigned long long), st));

    BakeParams p{};
    p.bvh.nodes = ctx->d_nodes.as<float4>(); p.bvh.tris = ctx->d_tris.as<float4>(); p.bvh.n_tris = (uint32_t)ctx->n_tris;
    p.bvh.overflow = reinterpret_cast<unsigned int*>(ctx->d_scratch.as<float>() + 13);
    p.shade.tri_shade = ctx->d_tri_shade.as<float4>(); p.shade.inst = ctx->d_inst.as<float4>();
    p.shade.base_color = ctx->d_base_color.as<float4>();
    p.shade.tri_uv = ctx->d_tri_uv.as<float4>(); p.shade.tex_desc = ctx->d_tex_desc.as<int4>();
    p.shade.tex_texels = ctx->d_tex_texels.as<uchar4>();
    if (ctx->max_tex_index >= (int)ctx->n_textures)
        return ctx->fail(VLB_ERR_STATE, "bake: a material names baseColor texture %d but %u textures are set (vlb_scene_set_textures)",
                         ctx->max_tex_index, ctx->n_textures);
    p.shade.sky = ctx->sky_w ? ctx->d_sky.as<float4>() : nullptr; p.shade.sky_w = ctx->sky_w; p.shade.sky_h = ctx->sky_h;
    for (int k = 0; k < 3; ++k) p.c.light[k] = s->light_pos[k];
    p.c.shadow_bias = s->shadow_bias; p.c.c_diffuse = s->c_diffuse; p.c.c_specular = s->c_specular;
    p.c.gloss = s->gloss; p.c.ambient = s->ambient; p.c.tmin = s->tmin; p.c.tmax = s->tmax; p.c.flags = s->flags;
    p.px = ctx->d_axis.as<float>(); p.py = p.px + Nx; p.pz = p.py + Ny;
    p.row_sc = ctx->d_row_sc.as<float2>(); p.col_cs = ctx->d_col_sc.as<float2>();
    p.Nx = Nx; p.Ny = Ny; p.Nz = Nz; p.k0 = k0; p.kstride = kstride; p.W = W; p.H = H;
    p.tile_lw = env_flag("VLB_BAKE_TILE_LW", (long long)W < 2ll * H ? 2 : 3);     // a function of the direction grid only
    p.tile_lw = std::max(0, std::min(5, p.tile_lw));
    const int tile_w = 1 << p.tile_lw, tile_h = 32 >> p.tile_lw;
    p.tiles_x = (W + tile_w - 1) / tile_w;
    p.n_tiles = p.tiles_x * ((H + tile_h - 1) / tile_h);
    // Work decomposition: a function of the WHOLE grid and the direction grid only (never of the
    // slab), so that a probe's coefficients are bit-identical however the grid is sharded. An item is
    // one 256-direction chunk of one probe (8 tiles: measured best on B200 for large grids); for
    // small grids the chunk is halved down to one tile until the whole grid has >= 2^18 items, i.e.
    // every GPU of an 8-way shard still sees several waves of warps.
    const uint64_t total_probes = (uint64_t)Nx * Ny * Nz;
    int tiles_per_item = kChunkTiles;
    while (tiles_per_item > 1 && total_probes * (uint64_t)((p.n_tiles + tiles_per_item - 1) / tiles_per_item) < (1ull << 18))
        tiles_per_item >>= 1;
    int chunks = (p.n_tiles + tiles_per_item - 1) / tiles_per_item;
    chunks = env_flag("VLB_BAKE_CHUNKS", chunks);
    chunks = std::max(1, std::min(chunks, p.n_tiles));
    p.tiles_per_chunk = (p.n_tiles + chunks - 1) / chunks;
    p.chunks = (p.n_tiles + p.tiles_per_chunk - 1) / p.tiles_per_chunk;
    if (n_probes * (uint64_t)p.chunks >= (1ull << 32)) return ctx->fail(VLB_ERR_UNSUPPORTED, "bake: too many work items");
    p.n_items = (uint32_t)(n_probes * (uint64_t)p.chunks);
    p.pixel_area = (2.0f * kPi / (float)W) * (kPi / (float)H);            // sh.comp:32
    p.work_counter = ctx->d_work_counter.as<unsigned int>();
    p.stats = ctx->d_stats.as<unsigned long long>();
    p.ref_order = (s->flags & VLB_BAKE_REFERENCE_PROBE_ORDER) ? 1 : 0;
    p.world_frame = (s->flags & VLB_BAKE_SH_WORLD_FRAME) ? 1 : 0;
    p.node_min = std::max(1, std::min(32, env_flag("VLB_BAKE_NODE_MIN", 8)));
    p.g.prev = d_prev_full; p.g.px = p.px; p.g.py = p.py; p.g.pz = p.pz; p.g.Nx = Nx; p.g.Ny = Ny; p.g.Nz = Nz;
    for (int k = 0; k < 3; ++k) { p.g.origin[k] = s->origin[k]; p.g.step[k] = s->step[k]; }
    p.g.gain = s->indirect_gain; p.g.world_frame = p.world_frame;
    const bool gather = d_prev_full != nullptr;
    if (p.chunks > 1) {
        VLB_CUDA(ctx, ctx->d_partials.reserve((size_t)p.n_items * VLB_SH_STRIDE * sizeof(float)));
        p.out = ctx->d_partials.as<float>();
    } else {
        p.out = d_out;
    }

    const bool count = env_flag("VLB_BAKE_COUNTERS", 0) != 0;
    const int K = s->sh_order == 2 ? 9 : 16;
    void (*kern)(const BakeParams) = nullptr;
    if (env_flag("VLB_BAKE_KERNEL", 2) == 1) {      // the round-1 tile-per-warp kernel, kept for A/B runs
        if (K == 9) kern = count ? k_bake<9, true> : k_bake<9, false>;
        else        kern = count ? k_bake<16, true> : k_bake<16, false>;
    } else {
        // TEX: only scenes with a textured material pay for the texture branch of the hit shading
        const bool tex = ctx->max_tex_index >= 0;
#define VLB_PICK(KK) (gather ? (tex ? (count ? k_bake_stream<KK, true, true, true> : k_bake_stream<KK, false, true, true>)     \
                                    : (count ? k_bake_stream<KK, true, true, false> : k_bake_stream<KK, false, true, false>))   \
                             : (tex ? (count ? k_bake_stream<KK, true, false, true> : k_bake_stream<KK, false, false, true>)   \
                                    : (count ? k_bake_stream<KK, true, false, false> : k_bake_stream<KK, false, false, false>)))
        kern = K == 9 ? VLB_PICK(9) : VLB_PICK(16);
#undef VLB_PICK
    }
    // the kernel uses no shared memory: ask for the whole 228 KB as L1 (BVH nodes, per-warp queues)
    if (env_flag("VLB_BAKE_CARVEOUT", 0) >= 0)
        VLB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, env_flag("VLB_BAKE_CARVEOUT", 0)));
    int per_sm = 0;
    VLB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kBakeBlock, 0));
    per_sm = std::max(per_sm, 1);
    const uint32_t warps_needed = p.n_items;
    uint32_t grid = (uint32_t)(ctx->sm_count * per_sm);
    grid = std::max(1u, std::min(grid, (warps_needed + (kBakeBlock / 32) - 1) / (kBakeBlock / 32)));

    VLB_CUDA(ctx, ctx->d_stream_scratch.reserve((size_t)grid * kStreamWarps * sizeof(WarpQueues)));
    p.stream_scratch = ctx->d_stream_scratch.as<WarpQueues>();
    VLB_CUDA(ctx, cudaEventRecord(ctx->ev[0], st));
    VLB_CUDA(ctx, cudaEventRecord(ctx->ev[2], st));
    kern<<<grid, kBakeBlock, 0, st>>>(p);
    VLB_LAUNCH_CHECK(ctx);
    VLB_CUDA(ctx, cudaEventRecord(ctx->ev[3], st));
    if (p.chunks > 1) {
        const uint32_t n = (uint32_t)n_probes * VLB_SH_STRIDE;
        k_sum_partials<<<(n + 255) / 256, 256, 0, st>>>(p, ctx->d_partials.as<float>(), (uint32_t)n_probes, d_out);
        VLB_LAUNCH_CHECK(ctx);
    }
    if (s->flags & VLB_BAKE_ACCUMULATE_ACROSS_PROBES) {
        k_running_total<<<1, 64, 0, st>>>(d_out, (uint32_t)n_probes);
        VLB_LAUNCH_CHECK(ctx);
    }
    VLB_CUDA(ctx, cudaEventRecord(ctx->ev[1], st));
    // statistics + overflow flag travel to pinned host memory asynchronously; bake_collect_stats() waits for them.
    // (The axis table above was copied from pageable memory, which is staged before cudaMemcpyAsync returns.)
    if (!ctx->h_bake_stats) VLB_CUDA(ctx, cudaHostAlloc(reinterpret_cast<void**>(&ctx->h_bake_stats), 4 * sizeof(unsigned long long), cudaHostAllocDefault));
    ctx->h_bake_stats[3] = 0;
    VLB_CUDA(ctx, cudaMemcpyAsync(ctx->h_bake_stats, ctx->d_stats.p, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    VLB_CUDA(ctx, cudaMemcpyAsync(ctx->h_bake_stats + 3, ctx->d_scratch.as<float>() + 13, 4, cudaMemcpyDeviceToHost, st));
    VLB_CUDA(ctx, cudaEventRecord(ctx->ev_done, st));
    ctx->last_bake.n_probes = n_probes;
    ctx->last_bake.n_primary_rays = n_probes * (uint64_t)W * H;
    ctx->bake_pending = true;
    return VLB_OK;
}

// Waits for the last enqueued bake and fills ctx->last_bake; reports a traversal stack overflow.
int bake_collect_stats(vlb_ctx* ctx) {
    if (!ctx->bake_pending) return VLB_OK;
    ctx->bake_pending = false;
    VLB_CUDA(ctx, cudaEventSynchronize(ctx->ev_done));
    vlb_bake_stats& b = ctx->last_bake;
    b.n_shadow_rays = ctx->h_bake_stats[0]; b.n_nodes_visited = ctx->h_bake_stats[1]; b.n_tris_tested = ctx->h_bake_stats[2];
    VLB_CUDA(ctx, cudaEventElapsedTime(&b.kernel_ms, ctx->ev[2], ctx->ev[3]));
    VLB_CUDA(ctx, cudaEventElapsedTime(&b.total_ms, ctx->ev[0], ctx->ev[1]));
    if (ctx->h_bake_stats[3])
        return ctx->fail(VLB_ERR_UNSUPPORTED, "bake: BVH traversal stack overflow (more than %d pending nodes on a ray)", kStackSize);
    return VLB_OK;
}

// =========================================================================================
// Validation: ray casts through the BVH and through a brute-force intersector that shares
// intersect_tri(), so hit ids must agree bit for bit (BASELINE.json north_star).
// =========================================================================================
__global__ void k_trace_bvh(BvhView b, const float* __restrict__ o, const float* __restrict__ d, uint32_t n, float tmin,
                            float tmax, int kind, int* __restrict__ ids, float* __restrict__ tuv) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Vec3 ro = mk3(o[3 * i], o[3 * i + 1], o[3 * i + 2]), rd = mk3(d[3 * i], d[3 * i + 1], d[3 * i + 2]);
    HitRec h; h.id = -1; h.t = tmax; h.u = 0.f; h.v = 0.f;
    if (kind == VLB_TRACE_ANY) {
        HitRec a;
        if (trace_any<false>(b, ro, rd, tmin, tmax, nullptr, &a)) h = a;
    } else {
        h = trace_closest<false>(b, ro, rd, tmin, tmax, nullptr);
    }
    ids[i] = h.id;
    if (tuv) { tuv[3 * i] = h.t; tuv[3 * i + 1] = h.u; tuv[3 * i + 2] = h.v; }
}

constexpr int kBruteSeg = 2048;   // triangles per thread block column

__global__ void k_brute_init(unsigned long long* keys, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keys[i] = ~0ull;
}

// grid = (ray blocks, triangle segments): every thread tests its ray against one segment and
// publishes (t bits, flat id) with a 64-bit atomicMin: smallest t, then smallest id.
__global__ void k_brute(const float4* __restrict__ tri_flat, uint32_t n_tris, const float* __restrict__ o,
                        const float* __restrict__ d, uint32_t n, float tmin, float tmax, unsigned long long* keys) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Vec3 ro = mk3(o[3 * i], o[3 * i + 1], o[3 * i + 2]), rd = mk3(d[3 * i], d[3 * i + 1], d[3 * i + 2]);
    const uint32_t s0 = blockIdx.y * kBruteSeg, s1 = min(s0 + kBruteSeg, n_tris);
    unsigned long long best = ~0ull;
    for (uint32_t k = s0; k < s1; ++k) {
        float t, u, v;
        if (intersect_tri(ld4(tri_flat + 3ull * k), ld4(tri_flat + 3ull * k + 1), ld4(tri_flat + 3ull * k + 2), ro, rd, t, u, v) &&
            t > tmin && t < tmax) {
            const unsigned long long key = ((unsigned long long)__float_as_uint(t) << 32) | k;
            best = key < best ? key : best;
        }
    }
    if (best != ~0ull) atomicMin(keys + i, best);
}

__global__ void k_brute_finish(const float4* __restrict__ tri_flat, const float* __restrict__ o, const float* __restrict__ d,
                               uint32_t n, float tmax, const unsigned long long* __restrict__ keys, int* ids, float* tuv) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long key = keys[i];
    int id = -1; float t = tmax, u = 0.f, v = 0.f;
    if (key != ~0ull) {
        id = (int)(key & 0xffffffffu);
        const Vec3 ro = mk3(o[3 * i], o[3 * i + 1], o[3 * i + 2]), rd = mk3(d[3 * i], d[3 * i + 1], d[3 * i + 2]);
        intersect_tri(tri_flat[3ull * id], tri_flat[3ull * id + 1], tri_flat[3ull * id + 2], ro, rd, t, u, v);
    }
    ids[i] = id;
    if (tuv) { tuv[3 * i] = t; tuv[3 * i + 1] = u; tuv[3 * i + 2] = v; }
}

int trace_rays(vlb_ctx* ctx, const float* o, const float* d, uint64_t n64, float tmin, float tmax, int accel, int kind,
               int32_t* ids, float* tuv) {
    if (n64 >= (1ull << 31)) return ctx->fail(VLB_ERR_UNSUPPORTED, "vlb_trace_rays: too many rays in one call");
    if (tmin < 0.f) return ctx->fail(VLB_ERR_INVALID, "vlb_trace_rays: tmin must be >= 0");
    const uint32_t n = (uint32_t)n64;
    cudaStream_t st = ctx->stream;
    VLB_CUDA(ctx, ctx->d_ray_o.reserve(3ull * n * sizeof(float)));
    VLB_CUDA(ctx, ctx->d_ray_d.reserve(3ull * n * sizeof(float)));
    VLB_CUDA(ctx, ctx->d_hit_id.reserve(n * sizeof(int)));
    VLB_CUDA(ctx, ctx->d_hit_tuv.reserve(3ull * n * sizeof(float)));
    VLB_CUDA(ctx, cudaMemcpyAsync(ctx->d_ray_o.p, o, 3ull * n * sizeof(float), cudaMemcpyHostToDevice, st));
    VLB_CUDA(ctx, cudaMemcpyAsync(ctx->d_ray_d.p, d, 3ull * n * sizeof(float), cudaMemcpyHostToDevice, st));
    const unsigned grid = (n + 127) / 128;
    if (accel == VLB_TRACE_BVH) {
        BvhView b; b.nodes = ctx->d_nodes.as<float4>(); b.tris = ctx->d_tris.as<float4>(); b.n_tris = (uint32_t)ctx->n_tris;
        b.overflow = reinterpret_cast<unsigned int*>(ctx->d_scratch.as<float>() + 13);
        k_trace_bvh<<<grid, 128, 0, st>>>(b, ctx->d_ray_o.as<float>(), ctx->d_ray_d.as<float>(), n, tmin, tmax, kind,
                                          ctx->d_hit_id.as<int>(), ctx->d_hit_tuv.as<float>());
        VLB_LAUNCH_CHECK(ctx);
    } else if (accel == VLB_TRACE_BRUTE_FORCE) {
        VLB_CUDA(ctx, ctx->d_hit_key.reserve(n * sizeof(unsigned long long)));
        k_brute_init<<<(n + 255) / 256, 256, 0, st>>>(ctx->d_hit_key.as<unsigned long long>(), n);
        VLB_LAUNCH_CHECK(ctx);
        const uint32_t nt = (uint32_t)ctx->n_tris;
        if (nt) {
            const dim3 g(grid, (nt + kBruteSeg - 1) / kBruteSeg);
            if (g.y > 65535) return ctx->fail(VLB_ERR_UNSUPPORTED, "brute force: scene too large");
            k_brute<<<g, 128, 0, st>>>(ctx->d_tri_flat.as<float4>(), nt, ctx->d_ray_o.as<float>(), ctx->d_ray_d.as<float>(), n,
                                       tmin, tmax, ctx->d_hit_key.as<unsigned long long>());
            VLB_LAUNCH_CHECK(ctx);
        }
        k_brute_finish<<<(n + 255) / 256, 256, 0, st>>>(ctx->d_tri_flat.as<float4>(), ctx->d_ray_o.as<float>(), ctx->d_ray_d.as<float>(),
                                                        n, tmax, ctx->d_hit_key.as<unsigned long long>(), ctx->d_hit_id.as<int>(),
                                                        ctx->d_hit_tuv.as<float>());
        VLB_LAUNCH_CHECK(ctx);
    } else {
        return ctx->fail(VLB_ERR_INVALID, "vlb_trace_rays: unknown accel");
    }
    VLB_CUDA(ctx, cudaMemcpyAsync(ids, ctx->d_hit_id.p, n * sizeof(int), cudaMemcpyDeviceToHost, st));
    if (tuv) VLB_CUDA(ctx, cudaMemcpyAsync(tuv, ctx->d_hit_tuv.p, 3ull * n * sizeof(float), cudaMemcpyDeviceToHost, st));
    unsigned int overflow = 0;
    if (accel == VLB_TRACE_BVH) VLB_CUDA(ctx, cudaMemcpyAsync(&overflow, ctx->d_scratch.as<float>() + 13, 4, cudaMemcpyDeviceToHost, st));
    VLB_CUDA(ctx, cudaStreamSynchronize(st));
    if (overflow) return ctx->fail(VLB_ERR_UNSUPPORTED, "vlb_trace_rays: BVH traversal stack overflow");
    return VLB_OK;
}

}  // namespace vlb

// skybox_sh.cu — equirect -> SH projection; replaces one dispatch of shaders/skybox_sh.comp
// (Skybox_t::computeSH, src/skybox_manager.cpp:107-130) and, with variant 1, of shaders/sh.comp
// on a caller-supplied environment image (src/baker/light_baker.cpp:68-73, 269-285).
//
// The reference evaluates 16 SH polynomials per texel and adds them, racily, into one SSBO.
// Here the sum is restructured so that the kernel is bound by HBM, not by issue slots: every
// real SH function of band <= 3 is a polynomial of degree <= 3 in the unit vector
// d = (S cos(phi), S sin(phi), C) with S = sin(theta), C = cos(theta), so
//
//   sum_texels w * rgb * dx^a dy^b dz^c  =  sum_x cos^a(phi_x) sin^b(phi_x) * H[a+b][c](x)
//   H[p][c](x) = sum_y w_y S_y^p C_y^c rgb(x,y)
//
// and with C^2 = 1 - S^2 only 5 (band <= 2) or 7 (band <= 3) row-weighted column sums
// G[p][q] (q in {0,1}) are needed: per texel one 16-byte read and 15 / 21 FMAs with warp-uniform
// row factors; the phi factors and the SH polynomials are applied once per column at the end.
//
// Two kernels share that math:
//   k_project_tma  (main path) persistent CTAs, one per SM. A producer warp streams 64-column x
//                  32-row texel tiles into a shared-memory ring with cp.async.bulk (TMA bulk copy,
//                  SASS UBLKCP) completing on mbarriers; 8 consumer warps read the tiles with
//                  conflict-free 16-byte shared loads. Bytes in flight are bounded by shared memory
//                  (192 KB per SM), not by registers, so one wave keeps ~25 MB outstanding.
//   k_project_ldg  fallback for rows that are not 16-byte multiples (e.g. RGBA8 with W % 4 != 0):
//                  one column per thread, 8 coalesced loads in flight per thread.
// Reduction (both): per-block 48-float partial -> the last block of each map sums the partials in
// a fixed order. No float atomics: results are bitwise reproducible for a given launch geometry.
//
// Algorithmic bytes per texel: 16 (RGBA32F) or 4 (RGBA8) read; 192 bytes written per map.
#include <cuda.h>   // CUtensorMap types only; the encoder is fetched with cudaGetDriverEntryPoint

#include <algorithm>

#include "vlb_context.h"
#include "vlb_math.cuh"
#include "vlb_warp.cuh"

namespace vlb {

constexpr int kProjBlock = 256;
constexpr int kProjUnroll = 8;
constexpr int kFinGroups = 21;        // 21 groups x 12 float4 lanes = 252 threads in the final sum

struct ProjParams {
    const void* texels;
    uint64_t map_stride;      // bytes between maps
    uint32_t n_maps;
    int W, H;
    int strips, row_blocks, rows_per_block;
    const float4* row_tab;    // 2 float4 per row: {w, wS, wS^2, wC}, {wSC, wS^3, wS^2C, 0}
    const float2* col_cs;     // (cos phi, sin phi) per column (phi already shifted for skyboxes)
    float* partials;          // [map][strips*row_blocks][48]
    unsigned int* counters;   // [map], self-resetting
    float* out;               // [map][48]
    int variant;              // 0: skybox_sh.comp (SH argument d.xzy), 1: sh.comp (SH argument d)
    // k_project_tiles only
    uint32_t n_tiles, tiles_per_map, row_tiles;
    uint32_t split_q, split_r;   // CTA b owns q + (b < r) consecutive tiles
    int n_stages;
#ifdef VLB_PROJ_TIMING
    unsigned long long* timing;  // [grid][8] %globaltimer stamps (debug builds only)
#endif
};

#ifdef VLB_PROJ_TIMING
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define VLB_STAMP(k) do { if ((threadIdx.x & 255) == 0 && (k) < 8) p.timing[blockIdx.x * 8 + (k)] = gtime(); } while (0)
#else
#define VLB_STAMP(k) do {} while (0)
#endif

// Moment of dx^a dy^b dz^c for one channel from the column sums G and the column's phi factors.
// G index: (p,q) -> {(0,0):0,(1,0):1,(2,0):2,(0,1):3,(1,1):4,(3,0):5,(2,1):6}
template <int NG>
struct ColumnMoments {
    float g[NG];
    float cp[4], sp[4];   // powers of cos(phi), sin(phi)
    __device__ __forceinline__ float G(int p, int q) const {
        if (q == 0) return p == 0 ? g[0] : (p == 1 ? g[1] : (p == 2 ? g[2] : g[NG > 5 ? 5 : 0]));
        return p == 0 ? g[3] : (p == 1 ? g[4] : g[NG > 6 ? 6 : 0]);
    }
    __device__ __forceinline__ float H(int p, int c) const {
        if (c == 0) return G(p, 0);
        if (c == 1) return G(p, 1);
        if (c == 2) return G(p, 0) - G(p + 2, 0);       // C^2 = 1 - S^2
        return G(p, 1) - G(p + 2, 1);                    // C^3 = C - S^2 C
    }
    __device__ __forceinline__ float Md(int a, int b, int c) const { return cp[a] * sp[b] * H(a + b, c); }
    // moment of sx^a sy^b sz^c where s = d.xzy (variant 0) or s = d (variant 1)
    __device__ __forceinline__ float M(int variant, int a, int b, int c) const {
        return variant == 0 ? Md(a, c, b) : Md(a, b, c);
    }
    __device__ __forceinline__ void set_phi(float2 cs) {
        cp[0] = 1.f; cp[1] = cs.x; cp[2] = cs.x * cs.x; cp[3] = cp[2] * cs.x;
        sp[0] = 1.f; sp[1] = cs.y; sp[2] = cs.y * cs.y; sp[3] = sp[2] * cs.y;
    }
};

// SH coefficients from monomial moments: shaders/sh_common.h:26-104 with every monomial replaced
// by its moment (the polynomials are linear in the monomials).
template <int K, int NG>
__device__ __forceinline__ void sh_from_moments(const ColumnMoments<NG>& m, int v, float* o /*K*/) {
    o[0] = 0.282095f * m.M(v, 0, 0, 0);
    o[1] = -0.488603f * m.M(v, 0, 1, 0);
    o[2] = 0.488603f * m.M(v, 0, 0, 1);
    o[3] = -0.488603f * m.M(v, 1, 0, 0);
    o[4] = 1.092548f * m.M(v, 1, 1, 0);
    o[5] = -1.092548f * m.M(v, 0, 1, 1);
    const float xx = m.M(v, 2, 0, 0), yy = m.M(v, 0, 2, 0), zz = m.M(v, 0, 0, 2);
    o[6] = 0.315392f * (-xx - yy + 2.0f * zz);
    o[7] = -1.092548f * m.M(v, 1, 0, 1);
    o[8] = 0.546274f * (xx - yy);
    if (K > 9) {
        const float xxy = m.M(v, 2, 1, 0), yyy = m.M(v, 0, 3, 0), yzz = m.M(v, 0, 1, 2);
        const float zzz = m.M(v, 0, 0, 3), xxz = m.M(v, 2, 0, 1), yyz = m.M(v, 0, 2, 1);
        const float xzz = m.M(v, 1, 0, 2), xxx = m.M(v, 3, 0, 0), xyy = m.M(v, 1, 2, 0);
        o[9] = -0.590044f * (3.0f * xxy - yyy);
        o[10] = 2.890611f * m.M(v, 1, 1, 1);
        o[11] = -0.457046f * (4.0f * yzz - xxy - yyy);
        o[12] = 0.373176f * (2.0f * zzz - 3.0f * xxz - 3.0f * yyz);
        o[13] = -0.457046f * (4.0f * xzz - xxx - xyy);
        o[14] = 1.445306f * (xxz - yyz);
        o[15] = -0.590044f * (xxx - 3.0f * xyy);
    }
}

struct FinishSmem {
    float4 grp[kFinGroups][12];
    int last;
};

// Publishes this block's 48-float partial (value held by threads tid < 48) and, in the block that
// arrives last for `map`, sums all P partials of the map in a fixed order and writes the result.
// SYNC() must be a barrier over exactly the threads that call this function.
template <class SYNC>
__device__ __forceinline__ void publish_and_finish(const ProjParams& p, uint32_t map, size_t base_slot, uint32_t slot,
                                                   uint32_t P, float my_value, int tid, FinishSmem& fs, SYNC sync) {
    if (tid < VLB_SH_STRIDE) {
        p.partials[(base_slot + slot) * VLB_SH_STRIDE + tid] = my_value;
        __threadfence();
    }
    sync();
    if (tid == 0) {
        const unsigned prev = atomicAdd(p.counters + map, 1u);
        fs.last = prev == P - 1;
        if (fs.last) p.counters[map] = 0;   // self-reset for the next launch
    }
    sync();
    if (fs.last) {
        __threadfence();
        const float4* base = reinterpret_cast<const float4*>(p.partials + base_slot * VLB_SH_STRIDE);
        if (tid < kFinGroups * 12) {
            const int g = tid / 12, c4 = tid % 12;
            float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
            // batches of 8 independent loads (one L2 round trip covers P <= 168 partials)
            for (uint32_t k0 = g; k0 < P; k0 += 8 * kFinGroups) {
                float4 v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint32_t k = k0 + j * kFinGroups;
                    v[j] = k < P ? __ldcg(base + (size_t)k * 12 + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) { s.x += v[j].x; s.y += v[j].y; s.z += v[j].z; s.w += v[j].w; }
            }
            fs.grp[g][c4] = s;
        }
        sync();
        if (tid < VLB_SH_STRIDE) {
            const float* col = reinterpret_cast<const float*>(&fs.grp[0][0]) + tid;
            float s = 0.f;
#pragma unroll
            for (int g = 0; g < kFinGroups; ++g) s += col[g * VLB_SH_STRIDE];
            p.out[(size_t)map * VLB_SH_STRIDE + tid] = s;
        }
    }
    sync();
}

template <int FMT>
__device__ __forceinline__ float4 load_texel(const void* base, size_t idx);
template <>
__device__ __forceinline__ float4 load_texel<VLB_FMT_RGBA32F>(const void* base, size_t idx) {
    return __ldcs(reinterpret_cast<const float4*>(base) + idx);   // streaming: read once
}
template <>
__device__ __forceinline__ float4 load_texel<VLB_FMT_RGBA8>(const void* base, size_t idx) {
    const uchar4 c = __ldcs(reinterpret_cast<const uchar4*>(base) + idx);
    return make_float4((float)c.x, (float)c.y, (float)c.z, (float)c.w);   // /255 applied at the end
}

// =========================================================================================
// Fallback kernel: one column per thread, LDG.128 / LDG.32.
// =========================================================================================
template <int K, int FMT>
__global__ void __launch_bounds__(kProjBlock) k_project_ldg(const ProjParams p) {
    constexpr int NG = K > 9 ? 7 : 5;
    constexpr int V = (K * 3 <= 32) ? 32 : 64;
    constexpr int R = V / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t bid = blockIdx.x;
    const uint32_t P = p.strips * p.row_blocks;
    const int rb = bid % p.row_blocks;
    const int strip = (bid / p.row_blocks) % p.strips;
    const uint32_t map = bid / P;
    const int x = strip * kProjBlock + tid;
    const bool active = x < p.W;
    const int y0 = rb * p.rows_per_block, y1 = min(p.H, y0 + p.rows_per_block);
    const char* map_base = reinterpret_cast<const char*>(p.texels) + (size_t)map * p.map_stride;

    float g[NG][3];
#pragma unroll
    for (int i = 0; i < NG; ++i) g[i][0] = g[i][1] = g[i][2] = 0.f;

    if (active) {
        for (int y = y0; y < y1; y += kProjUnroll) {
            float4 tx[kProjUnroll];
#pragma unroll
            for (int u = 0; u < kProjUnroll; ++u)
                if (y + u < y1) tx[u] = load_texel<FMT>(map_base, (size_t)(y + u) * p.W + x);
#pragma unroll
            for (int u = 0; u < kProjUnroll; ++u) {
                if (y + u < y1) {
                    const float4 ra = __ldg(p.row_tab + 2 * (y + u));
                    const float4 rb4 = __ldg(p.row_tab + 2 * (y + u) + 1);
                    const float t[7] = {ra.x, ra.y, ra.z, ra.w, rb4.x, rb4.y, rb4.z};
#pragma unroll
                    for (int i = 0; i < NG; ++i) {
                        g[i][0] = fmaf(t[i], tx[u].x, g[i][0]);
                        g[i][1] = fmaf(t[i], tx[u].y, g[i][1]);
                        g[i][2] = fmaf(t[i], tx[u].z, g[i][2]);
                    }
                }
            }
        }
    }

    // per-thread end stage: phi factors of this column, then the SH polynomials
    float acc[V];
#pragma unroll
    for (int i = 0; i < V; ++i) acc[i] = 0.f;
    if (active) {
        const float2 cs = __ldg(p.col_cs + x);
        const float scale = FMT == VLB_FMT_RGBA8 ? (1.0f / 255.0f) : 1.0f;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            ColumnMoments<NG> m;
#pragma unroll
            for (int i = 0; i < NG; ++i) m.g[i] = g[i][ch] * scale;
            m.set_phi(cs);
            float o[K];
            sh_from_moments<K, NG>(m, p.variant, o);
#pragma unroll
            for (int i = 0; i < K; ++i) acc[3 * i + ch] = o[i];
        }
    }

    __shared__ float s_red[kProjBlock / 32][V];
    __shared__ FinishSmem s_fin;
    warp_transpose_reduce<V>(acc, lane);
#pragma unroll
    for (int r = 0; r < R; ++r) s_red[warp][R * lane + r] = acc[r];
    __syncthreads();
    float mine = 0.f;
    if (tid < K * 3) {
#pragma unroll
        for (int w = 0; w < kProjBlock / 32; ++w) mine += s_red[w][tid];
    }
    publish_and_finish(p, map, (size_t)map * P, bid % P, P, mine, tid, s_fin, [] { __syncthreads(); });
}

// =========================================================================================
// Main kernel: persistent CTAs (one per SM), TMA tensor-map tiles into a shared-memory ring.
//
// Tiles of 64 columns x 32 rows are linearised as ((map * strips + strip) * row_tiles + row_tile)
// and split evenly over the CTAs (tile counts differ by at most one). One elected producer thread
// issues, per tile, ONE cp.async.bulk.tensor.3d (SASS UTMALDG; out-of-range rows / columns are
// zero-filled by the TMA unit, so there are no edge predicates in the consumers) plus one 1 KB
// cp.async.bulk (UBLKCP) of the tile's 32 row-weight records; both complete on the stage's
// mbarrier. 8 consumer warps read the stage with conflict-free LDS.128 (texels) and broadcast
// LDS.128 (weights). Per-column sums live in registers across the row tiles of one strip; at a
// strip boundary they are folded into the SH polynomials, at a map boundary the CTA's 48-float
// partial is published; the last CTA of a map sums that map's partials in CTA order.
// =========================================================================================
constexpr int kTCols = 64;                 // columns per tile
constexpr int kTPhases = 4;                // row phases: 64 columns x 4 phases = 256 consumer threads
constexpr int kTConsumers = kTCols * kTPhases;
constexpr int kTThreads = kTConsumers + 32;
constexpr int kTMaxStages = 8;
constexpr int kTCtasPerSm = 2;             // a long launch fills two CTA slots per SM (5 stages each): one streams while the other flushes
constexpr int kTRegSlots = 3;              // registers are held to 72 so that THREE short launches (3 stages each) of a chain share an SM
// rows per tile (= pipeline stage): 16 KB of texels per TMA instruction in either format
__host__ __device__ constexpr int tile_rows(int fmt) { return fmt == VLB_FMT_RGBA32F ? 16 : 64; }
constexpr int kTRowsMax = 64;
constexpr int kScratchSets = 1 + VLB_MAX_LANES;   // the ctx stream's set + one per auxiliary lane

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    for (uint32_t spins = 0; !done; ++spins) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (!done && spins > (1u << 24)) __trap();   // a lost transaction must abort, not hang the GPU
    }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_tile_g2s(void* dst, const CUtensorMap* tmap, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(dst)), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kTConsumers) : "memory"); }

// CTA that owns tile t under the even split (first r CTAs own q + 1 tiles, the rest q).
__device__ __forceinline__ uint32_t cta_of_tile(const ProjParams& p, uint32_t t) {
    const uint32_t big = p.split_r * (p.split_q + 1);
    return t < big ? t / (p.split_q + 1) : p.split_r + (t - big) / p.split_q;
}

template <int K, int FMT>
__global__ void __launch_bounds__(kTThreads, kTRegSlots) k_project_tiles(const __grid_constant__ CUtensorMap tmap, const ProjParams p) {
    constexpr int NG = K > 9 ? 7 : 5;
    constexpr int BPT = FMT == VLB_FMT_RGBA32F ? 16 : 4;
    constexpr int TR = tile_rows(FMT);
    constexpr int TILE_BYTES = TR * kTCols * BPT;
    constexpr int WEIGHT_BYTES = TR * 32;            // 8 floats per row
    constexpr int STAGE_BYTES = TILE_BYTES + WEIGHT_BYTES;
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full[kTMaxStages], empty[kTMaxStages];
    __shared__ float s_red[6][32];
    __shared__ FinishSmem s_fin;
    // TMA tile destinations must be 128-byte aligned
    unsigned char* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    float* s_g = reinterpret_cast<float*>(smem + (size_t)p.n_stages * STAGE_BYTES);   // [4][64][NG*3]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int S = p.n_stages;
    const uint32_t b = blockIdx.x;
    const uint32_t t0 = b * p.split_q + min(b, p.split_r);
    const uint32_t nt = p.split_q + (b < p.split_r ? 1u : 0u);
    VLB_STAMP(0);
    // Programmatic dependent launch (project_sh_device chains back-to-back projections on the ctx's own stream): the next
    // launch of the chain may be scheduled as soon as every CTA of this one has started, and streams its texels while this
    // launch reduces and retires. Texels are never written by a projection, so only the shared scratch (partials, counters)
    // and the output need ordering: griddep_wait() below, right before the first of those writes. Both instructions are
    // no-ops for an ordinary launch.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (tid == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kTConsumers / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == kTConsumers / 32) {
        // ===== producer: one elected thread, two copies per tile =====
        if (lane == 0) {
            for (uint32_t i = 0; i < nt; ++i) {
                const uint32_t t = t0 + i;
                const uint32_t rt = t % p.row_tiles, ms = t / p.row_tiles;
                const uint32_t strip = ms % p.strips, map = ms / p.strips;
                const int s = i % S;
                mbar_wait(&empty[s], ((i / S) & 1) ^ 1);
                unsigned char* stage = smem + (size_t)s * STAGE_BYTES;
                mbar_arrive_expect_tx(&full[s], STAGE_BYTES);
                tma_tile_g2s(stage, &tmap, (int)(strip * kTCols * 4), (int)(rt * TR), (int)map, &full[s]);
                bulk_g2s(stage + TILE_BYTES, reinterpret_cast<const char*>(p.row_tab) + (size_t)rt * WEIGHT_BYTES,
                         WEIGHT_BYTES, &full[s]);
            }
        }
        return;
    }

    // ===== consumer warps =====
    const int col = tid & (kTCols - 1), ph = tid / kTCols;
    float g[NG][3];
#pragma unroll
    for (int i = 0; i < NG; ++i) g[i][0] = g[i][1] = g[i][2] = 0.f;
    float macc = 0.f;   // threads tid < K*3: this CTA's running partial of coefficient tid of the current map
    for (uint32_t i = 0; i < nt; ++i) {
        const uint32_t t = t0 + i;
        const int s = i % S;
        mbar_wait(&full[s], (i / S) & 1);
        if (i == 0) VLB_STAMP(1);
        const unsigned char* stage = smem + (size_t)s * STAGE_BYTES;
        const float4* wts = reinterpret_cast<const float4*>(stage + TILE_BYTES);
#pragma unroll
        for (int j = 0; j < TR / kTPhases; ++j) {
            const int rr = j * kTPhases + ph;
            float4 tx;
            if (FMT == VLB_FMT_RGBA32F) {
                tx = reinterpret_cast<const float4*>(stage)[rr * kTCols + col];
            } else {
                const uchar4 c = reinterpret_cast<const uchar4*>(stage)[rr * kTCols + col];
                tx = make_float4((float)c.x, (float)c.y, (float)c.z, 0.f);
            }
            const float4 ra = wts[2 * rr], rb4 = wts[2 * rr + 1];
            const float tw[7] = {ra.x, ra.y, ra.z, ra.w, rb4.x, rb4.y, rb4.z};
#pragma unroll
            for (int k = 0; k < NG; ++k) {
                g[k][0] = fmaf(tw[k], tx.x, g[k][0]);
                g[k][1] = fmaf(tw[k], tx.y, g[k][1]);
                g[k][2] = fmaf(tw[k], tx.z, g[k][2]);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);

        const bool last = i + 1 == nt;
        if (i == 1) VLB_STAMP(6);
        if (i == 3) VLB_STAMP(7);
        if (last) VLB_STAMP(2);
        if (last || (t + 1) % p.row_tiles == 0) {
            // ---- strip boundary: merge the 4 row phases, apply the phi factors, reduce over columns ----
            const uint32_t ms = t / p.row_tiles, strip = ms % p.strips;
            float* mine_g = s_g + ((size_t)ph * kTCols + col) * (NG * 3);
#pragma unroll
            for (int k = 0; k < NG; ++k) {
                mine_g[3 * k] = g[k][0]; mine_g[3 * k + 1] = g[k][1]; mine_g[3 * k + 2] = g[k][2];
                g[k][0] = g[k][1] = g[k][2] = 0.f;
            }
            consumer_sync();
            if (tid < 3 * kTCols) {
                const int c = tid & (kTCols - 1), ch = tid / kTCols;
                ColumnMoments<NG> m;
                const float scale = FMT == VLB_FMT_RGBA8 ? (1.0f / 255.0f) : 1.0f;
#pragma unroll
                for (int k = 0; k < NG; ++k) {
                    float v = 0.f;
#pragma unroll
                    for (int q = 0; q < kTPhases; ++q) v += s_g[((size_t)q * kTCols + c) * (NG * 3) + 3 * k + ch];
                    m.g[k] = v * scale;
                }
                float o[32];
#pragma unroll
                for (int k = 0; k < 32; ++k) o[k] = 0.f;
                const int x = (int)strip * kTCols + c;
                if (x < p.W) {
                    m.set_phi(__ldg(p.col_cs + x));
                    sh_from_moments<K, NG>(m, p.variant, o);
                }
                warp_transpose_reduce<32>(o, lane);        // lane l: sum over this warp's 32 columns of coefficient l
                s_red[warp][lane] = o[0];
            }
            consumer_sync();
            if (tid < K * 3) {
                const int k = tid / 3, ch = tid % 3;
                macc += s_red[2 * ch][k] + s_red[2 * ch + 1][k];
            }
            if (last) VLB_STAMP(3);
            if (last || (t + 1) % p.tiles_per_map == 0) {
                // ---- map boundary: publish this CTA's partial of the map ----
                const uint32_t map = t / p.tiles_per_map;
                const uint32_t b_first = cta_of_tile(p, map * p.tiles_per_map);
                const uint32_t b_last = cta_of_tile(p, (map + 1) * p.tiles_per_map - 1);
                const uint32_t P = b_last - b_first + 1;
                asm volatile("griddepcontrol.wait;" ::: "memory");      // the previous launch of the chain has retired: scratch and output are ours
                if (P == 1) {
                    if (tid < VLB_SH_STRIDE) p.out[(size_t)map * VLB_SH_STRIDE + tid] = macc;
                } else {
                    // (CTA, map) pairs form a monotone staircase, so slot b + map is unique
                    publish_and_finish(p, map, (size_t)b_first + map, b - b_first, P, macc, tid, s_fin, [] { consumer_sync(); });
                }
                macc = 0.f;
                if (last) VLB_STAMP(4);
            }
        }
    }
    VLB_STAMP(5);
}

static int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

// Tuning knobs read once (a projection of one map is a 5 microsecond launch: no getenv on that path).
struct ProjEnv { int tma, ctas_per_sm, stages, pdl, blocks_per_sm, rows; };
static const ProjEnv& proj_env() {
    static const ProjEnv e = {env_int("VLB_PROJ_TMA", 1), env_int("VLB_PROJ_CTAS_PER_SM", 0), env_int("VLB_PROJ_STAGES", 0),
                              env_int("VLB_PROJ_PDL", 1), env_int("VLB_PROJ_BLOCKS_PER_SM", 2), env_int("VLB_PROJ_ROWS", 0)};
    return e;
}

#ifdef VLB_PROJ_TIMING
static DevBuf g_tbuf;
static unsigned g_tgrid = 0;
int proj_timing_dump(vlb_ctx* ctx, int n_launches) {
    if (!env_int("VLB_PROJ_TIMING", 0) || !g_tbuf.p) return VLB_OK;
    std::vector<unsigned long long> h((size_t)kScratchSets * 512 * 8);
    VLB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    VLB_CUDA(ctx, cudaMemcpy(h.data(), g_tbuf.p, h.size() * 8, cudaMemcpyDeviceToHost));
    unsigned long long t0 = ~0ull;
    for (unsigned b = 0; b < g_tgrid; ++b) t0 = std::min(t0, h[b * 8]);
    static const char* names[8] = {"start", "first", "last", "flushed", "published", "exit", "tile1", "tile3"};
    for (int l = 0; l < n_launches && l < kScratchSets; ++l) {
        fprintf(stderr, "[launch %d us]", l);
        for (int k = 0; k < 8; ++k) {
            double mn = 1e30, mx = 0, av = 0;
            for (unsigned b = 0; b < g_tgrid; ++b) {
                const double d = (double)(long long)(h[((size_t)l * 512 + b) * 8 + k] - t0) * 1e-3;
                mn = std::min(mn, d); mx = std::max(mx, d); av += d / g_tgrid;
            }
            fprintf(stderr, " %s %.1f/%.1f/%.1f |", names[k], mn, av, mx);
        }
        fprintf(stderr, "\n");
    }
    return VLB_OK;
}
#endif

template <int K, int FMT>
static cudaError_t launch_tiles(const CUtensorMap& tmap, const ProjParams& p, unsigned grid, size_t smem, cudaStream_t st, bool chained) {
    static size_t allowed = 0;   // per instantiation; raising the limit is idempotent, so a race is harmless
    if (smem > allowed) {
        cudaError_t e = cudaFuncSetAttribute(k_project_tiles<K, FMT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        allowed = smem;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kTThreads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = chained ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, k_project_tiles<K, FMT>, tmap, p);
}

// Partials + arrival counters of scratch set `set` (0: ctx stream, 1 + lane: auxiliary lanes). Growing a
// buffer frees the old one, which synchronises the device, so it is safe while other lanes run; the
// counters are self-resetting and only need zeroing when (re)allocated.
static int scratch_set(vlb_ctx* ctx, int set, size_t n_slots, size_t n_maps, float** partials, unsigned** counters) {
    if (ctx->d_proj_partials.cap < kScratchSets * n_slots * VLB_SH_STRIDE * sizeof(float) ||
        ctx->d_proj_counters.cap < kScratchSets * n_maps * sizeof(unsigned)) {
        VLB_CUDA(ctx, cudaDeviceSynchronize());
        VLB_CUDA(ctx, ctx->d_proj_partials.reserve(kScratchSets * n_slots * VLB_SH_STRIDE * sizeof(float)));
        VLB_CUDA(ctx, ctx->d_proj_counters.reserve(kScratchSets * n_maps * sizeof(unsigned)));
        VLB_CUDA(ctx, cudaMemset(ctx->d_proj_counters.p, 0, ctx->d_proj_counters.cap));
    }
    // sets are laid out by the CURRENT call's sizes; a set is only ever used by one stream at a time
    *partials = ctx->d_proj_partials.as<float>() + (size_t)set * n_slots * VLB_SH_STRIDE;
    *counters = ctx->d_proj_counters.as<unsigned>() + (size_t)set * n_maps;
    return VLB_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// cuTensorMapEncodeTiled through the runtime's driver entry point query: libvlb_bake.so links only libcudart.
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            f = nullptr;
        return reinterpret_cast<EncodeTiledFn>(f);
    }();
    return fn;
}

// lane < 0: an ordinary launch on the ctx stream. lane >= 0: a launch on the ctx's auxiliary stream
// `lane` with that lane's scratch set (vlb_skybox_project_sh_device_ptrs: independent maps alternate
// over the lanes, so one map streams from HBM while another flushes, reduces and retires).
int project_sh_device(vlb_ctx* ctx, const void* d_texels, uint64_t map_stride, uint32_t n_maps, int fmt, int W, int H,
                      int order, int variant, float* d_out, int lane, bool chain_in_lane) {
    cudaStream_t st = lane < 0 ? ctx->stream : ctx->lane_stream[lane];
    const ProjEnv& env = proj_env();
    // chain onto the previous projection launch (programmatic dependent launch, see k_project_tiles) only when this
    // ctx knows that nothing else was enqueued on the stream in between: its own stream, previous call = a projection;
    // or (chain_in_lane) the previous launch on this lane belongs to the same vlb_skybox_project_sh_device_ptrs call
    bool chained = ((lane < 0 && ctx->proj_chain && st == ctx->own_stream) || (lane >= 0 && chain_in_lane)) && env.pdl != 0;
    ctx->proj_chain = false;
    // tables (cached per size/variant): per-row quadrature factors (zero-padded to whole tiles), per-column cos/sin(phi)
    if (ctx->tab_w != W || ctx->tab_h != H || ctx->tab_variant != variant) {
        chained = false;
        const size_t h_pad = ((size_t)H + kTRowsMax - 1) / kTRowsMax * kTRowsMax;
        std::vector<float> row_tab(8 * h_pad, 0.f), row_sc(2 * (size_t)H), col(2 * (size_t)W);
        host_proj_row_table(W, H, row_tab.data());
        host_dir_tables(W, H, variant == 0 ? kPi / 2.0f : 0.f, row_sc.data(), col.data());   // skybox_sh.comp:28
        VLB_CUDA(ctx, ctx->d_row_tab.reserve(row_tab.size() * sizeof(float)));
        VLB_CUDA(ctx, ctx->d_col_tab.reserve(col.size() * sizeof(float)));
        VLB_CUDA(ctx, cudaMemcpyAsync(ctx->d_row_tab.p, row_tab.data(), row_tab.size() * sizeof(float), cudaMemcpyHostToDevice, st));
        VLB_CUDA(ctx, cudaMemcpyAsync(ctx->d_col_tab.p, col.data(), col.size() * sizeof(float), cudaMemcpyHostToDevice, st));
        VLB_CUDA(ctx, cudaStreamSynchronize(st));
        ctx->tab_w = W; ctx->tab_h = H; ctx->tab_variant = variant;
    }
    ProjParams p{};
    p.texels = d_texels; p.map_stride = map_stride; p.n_maps = n_maps; p.W = W; p.H = H;
    p.row_tab = ctx->d_row_tab.as<float4>(); p.col_cs = ctx->d_col_tab.as<float2>();
    p.out = d_out; p.variant = variant;
    const int bpt = fmt == VLB_FMT_RGBA32F ? 16 : 4;
    // TMA needs a 16-byte aligned base and 16-byte multiples for the row and map pitches
    const bool aligned = ((size_t)W * bpt) % 16 == 0 && (n_maps == 1 || map_stride % 16 == 0) &&
                         (reinterpret_cast<uintptr_t>(d_texels) % 16) == 0 && (uint64_t)W * 4 < (1ull << 32);
    const bool use_tma = aligned && env.tma != 0 && encode_tiled_fn() != nullptr;

    if (use_tma) {
        p.strips = (W + kTCols - 1) / kTCols;
        const int TR = tile_rows(fmt);
        p.row_tiles = (uint32_t)((H + TR - 1) / TR);
        const uint64_t tiles_per_map = (uint64_t)p.strips * p.row_tiles;
        const uint64_t n_tiles = tiles_per_map * n_maps;
        if (n_tiles >= (1ull << 31)) return ctx->fail(VLB_ERR_UNSUPPORTED, "project_sh: too many tiles");
        p.tiles_per_map = (uint32_t)tiles_per_map; p.n_tiles = (uint32_t)n_tiles;
        // Long launches (many maps) fill both CTA slots of every SM. Short ones (a single map) take one
        // slot per SM, so that a launch on another lane / stream runs in the other slot and streams
        // while this one flushes, publishes and sums.
        const bool long_launch = n_tiles >= (uint64_t)ctx->sm_count * kTCtasPerSm * 16;
        const int per_sm = env.ctas_per_sm > 0 ? env.ctas_per_sm : (long_launch ? kTCtasPerSm : 1);
        const unsigned grid = (unsigned)std::min<uint64_t>(n_tiles, (uint64_t)ctx->sm_count * std::max(1, std::min(per_sm, kTCtasPerSm)));
        p.split_q = p.n_tiles / grid; p.split_r = p.n_tiles % grid;

        // the encoded tensor map is cached per (pointer, shape): encoding costs about as much as the launch itself
        vlb_ctx::TmapEntry* hit = nullptr;
        for (vlb_ctx::TmapEntry& e : ctx->tmap_cache)
            if (e.ptr == d_texels && e.W == W && e.H == H && e.fmt == fmt && e.n_maps == n_maps && e.stride == map_stride) { hit = &e; break; }
        if (!hit) {
            hit = &ctx->tmap_cache[ctx->tmap_next++ % (sizeof ctx->tmap_cache / sizeof ctx->tmap_cache[0])];
            static_assert(sizeof(CUtensorMap) <= sizeof hit->map, "tensor map cache slot too small");
            CUtensorMap* tm = reinterpret_cast<CUtensorMap*>(hit->map);
            const cuuint64_t gdim[3] = {(cuuint64_t)W * 4, (cuuint64_t)H, (cuuint64_t)n_maps};
            const cuuint64_t gstride[2] = {(cuuint64_t)W * bpt, n_maps == 1 ? (cuuint64_t)W * bpt * H : (cuuint64_t)map_stride};
            const cuuint32_t box[3] = {(cuuint32_t)kTCols * 4, (cuuint32_t)TR, 1};
            const cuuint32_t estride[3] = {1, 1, 1};
            const CUresult cr = encode_tiled_fn()(tm, bpt == 16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 3,
                                                  const_cast<void*>(d_texels), gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (cr != CUDA_SUCCESS) { hit->ptr = nullptr; return ctx->fail(VLB_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)cr); }
            hit->ptr = d_texels; hit->stride = map_stride; hit->n_maps = n_maps; hit->W = W; hit->H = H; hit->fmt = fmt;
        }
        const CUtensorMap& tmap = *reinterpret_cast<const CUtensorMap*>(hit->map);

        const int stage_bytes = TR * kTCols * bpt + TR * 32;
        const int ng3 = (order == 2 ? 5 : 7) * 3;
        const size_t merge_bytes = (size_t)kTPhases * kTCols * ng3 * sizeof(float);
        // lanes: 3 stages (52 KB in flight per CTA) so that three launches fit on an SM side by side
        // one launch takes one CTA slot per SM (unless it is a long batch): with 3 stages (52 KB in flight per CTA) three
        // launches of a chain / of the lanes fit on an SM side by side, one streaming while its neighbours start up or retire
        int stages = env.stages > 0 ? env.stages : (long_launch ? 5 : 3);
        stages = std::max(2, std::min(stages, kTMaxStages));
        p.n_stages = stages;
        const size_t smem = (size_t)stages * stage_bytes + merge_bytes + 128;
        const size_t n_slots = (size_t)grid + n_maps;
        const int set = lane < 0 ? 0 : 1 + lane;
        if (int r = scratch_set(ctx, set, n_slots, n_maps, &p.partials, &p.counters)) return r;
#ifdef VLB_PROJ_TIMING
        VLB_CUDA(ctx, g_tbuf.reserve((size_t)kScratchSets * 512 * 8 * sizeof(unsigned long long)));
        p.timing = g_tbuf.as<unsigned long long>() + (size_t)set * 512 * 8;
        g_tgrid = grid;
#endif
        cudaError_t e;
        if (order == 2) e = fmt == VLB_FMT_RGBA32F ? launch_tiles<9, VLB_FMT_RGBA32F>(tmap, p, grid, smem, st, chained) : launch_tiles<9, VLB_FMT_RGBA8>(tmap, p, grid, smem, st, chained);
        else            e = fmt == VLB_FMT_RGBA32F ? launch_tiles<16, VLB_FMT_RGBA32F>(tmap, p, grid, smem, st, chained) : launch_tiles<16, VLB_FMT_RGBA8>(tmap, p, grid, smem, st, chained);
        VLB_CUDA(ctx, e);
        VLB_LAUNCH_CHECK(ctx);
        ctx->proj_chain = lane < 0 && st == ctx->own_stream;     // the next device projection may chain onto this launch
        return VLB_OK;
    }

    p.strips = (W + kProjBlock - 1) / kProjBlock;
    // launch geometry: about 2 blocks per SM for a single map, whole columns for big batches
    const long long target = (long long)ctx->sm_count * env.blocks_per_sm;
    long long rbw = std::max<long long>(1, target / std::max<long long>(1, (long long)n_maps * p.strips));
    int rows = (int)((H + rbw - 1) / rbw);
    rows = std::max(rows, kProjUnroll);
    if (env.rows > 0) rows = env.rows;
    rows = std::max(1, std::min(rows, H));
    p.rows_per_block = rows;
    p.row_blocks = (H + rows - 1) / rows;
    const uint64_t P = (uint64_t)p.strips * p.row_blocks;
    const uint64_t n_blocks = P * n_maps;
    if (n_blocks >= (1ull << 31)) return ctx->fail(VLB_ERR_UNSUPPORTED, "project_sh: launch too large");
    if (int r = scratch_set(ctx, lane < 0 ? 0 : 1 + lane, n_blocks, n_maps, &p.partials, &p.counters)) return r;
    const unsigned grid = (unsigned)n_blocks;
    if (order == 2) {
        if (fmt == VLB_FMT_RGBA32F) k_project_ldg<9, VLB_FMT_RGBA32F><<<grid, kProjBlock, 0, st>>>(p);
        else                        k_project_ldg<9, VLB_FMT_RGBA8><<<grid, kProjBlock, 0, st>>>(p);
    } else {
        if (fmt == VLB_FMT_RGBA32F) k_project_ldg<16, VLB_FMT_RGBA32F><<<grid, kProjBlock, 0, st>>>(p);
        else                        k_project_ldg<16, VLB_FMT_RGBA8><<<grid, kProjBlock, 0, st>>>(p);
    }
    VLB_LAUNCH_CHECK(ctx);
    return VLB_OK;
}

}  // namespace vlb

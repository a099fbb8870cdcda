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
    // k_project_tma only
    uint32_t n_units;
    int n_stages, bpt;
};

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
__device__ __forceinline__ void publish_and_finish(const ProjParams& p, uint32_t map, uint32_t slot, uint32_t P,
                                                   float my_value, int tid, FinishSmem& fs, SYNC sync) {
    if (tid < VLB_SH_STRIDE) {
        p.partials[((size_t)map * P + slot) * VLB_SH_STRIDE + tid] = my_value;
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
        const float4* base = reinterpret_cast<const float4*>(p.partials + (size_t)map * P * VLB_SH_STRIDE);
        if (tid < kFinGroups * 12) {
            const int g = tid / 12, c4 = tid % 12;
            float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
            for (uint32_t k = g; k < P; k += kFinGroups) {
                const float4 v = __ldcg(base + (size_t)k * 12 + c4);
                s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
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
    publish_and_finish(p, map, bid % P, P, mine, tid, s_fin, [] { __syncthreads(); });
}

// =========================================================================================
// Main kernel: persistent CTAs, TMA bulk copies into a shared-memory ring.
// =========================================================================================
constexpr int kTCols = 64;                 // columns per tile
constexpr int kTRows = 32;                 // rows per pipeline stage
constexpr int kTPhases = 4;                // row phases: 64 columns x 4 phases = 256 consumer threads
constexpr int kTConsumers = kTCols * kTPhases;
constexpr int kTThreads = kTConsumers + 32;
constexpr int kTMaxStages = 8;

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
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kTConsumers) : "memory"); }

struct TmaUnit {
    uint32_t map;
    int strip, r0, r1, c0, ncols;
};
__device__ __forceinline__ TmaUnit decode_unit(const ProjParams& p, uint32_t u) {
    TmaUnit t;
    const int rb = u % p.row_blocks;
    t.strip = (u / p.row_blocks) % p.strips;
    t.map = u / (p.row_blocks * p.strips);
    t.r0 = rb * p.rows_per_block;
    t.r1 = min(p.H, t.r0 + p.rows_per_block);
    t.c0 = t.strip * kTCols;
    t.ncols = min(kTCols, p.W - t.c0);
    return t;
}

template <int K, int FMT>
__global__ void __launch_bounds__(kTThreads, 1) k_project_tma(const ProjParams p) {
    constexpr int NG = K > 9 ? 7 : 5;
    constexpr int BPT = FMT == VLB_FMT_RGBA32F ? 16 : 4;
    constexpr int STAGE_BYTES = kTRows * kTCols * BPT;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t full[kTMaxStages], empty[kTMaxStages];
    __shared__ float s_red[6][32];
    __shared__ FinishSmem s_fin;
    float* s_g = reinterpret_cast<float*>(smem + (size_t)p.n_stages * STAGE_BYTES);   // [4][64][NG*3]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int S = p.n_stages;
    if (tid == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kTConsumers / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == kTConsumers / 32) {
        // ===== producer warp: one bulk copy per tile row, 32 rows per stage =====
        uint32_t it = 0;
        for (uint32_t u = blockIdx.x; u < p.n_units; u += gridDim.x) {
            const TmaUnit t = decode_unit(p, u);
            const uint32_t row_bytes = (uint32_t)t.ncols * BPT;
            const char* gbase = reinterpret_cast<const char*>(p.texels) + (size_t)t.map * p.map_stride +
                                ((size_t)t.r0 * p.W + t.c0) * BPT;
            for (int r = t.r0; r < t.r1; r += kTRows, ++it) {
                const int s = it % S;
                mbar_wait(&empty[s], ((it / S) & 1) ^ 1);
                const int nr = min(kTRows, t.r1 - r);
                if (lane == 0) mbar_arrive_expect_tx(&full[s], (uint32_t)nr * row_bytes);
                __syncwarp();
                if (lane < nr)
                    bulk_g2s(smem + (size_t)s * STAGE_BYTES + (size_t)lane * kTCols * BPT,
                             gbase + (size_t)(r - t.r0 + lane) * p.W * BPT, row_bytes, &full[s]);
            }
        }
        return;
    }

    // ===== consumer warps =====
    const int col = tid & (kTCols - 1), ph = tid / kTCols;
    uint32_t it = 0;
    for (uint32_t u = blockIdx.x; u < p.n_units; u += gridDim.x) {
        const TmaUnit t = decode_unit(p, u);
        float g[NG][3];
#pragma unroll
        for (int i = 0; i < NG; ++i) g[i][0] = g[i][1] = g[i][2] = 0.f;
        for (int r = t.r0; r < t.r1; r += kTRows, ++it) {
            const int s = it % S;
            mbar_wait(&full[s], (it / S) & 1);
            const int nr = min(kTRows, t.r1 - r);
            const unsigned char* stage = smem + (size_t)s * STAGE_BYTES;
            if (col < t.ncols) {
#pragma unroll
                for (int j = 0; j < kTRows / kTPhases; ++j) {
                    const int rr = ph + j * kTPhases;
                    if (rr < nr) {
                        float4 tx;
                        if (FMT == VLB_FMT_RGBA32F) {
                            tx = reinterpret_cast<const float4*>(stage)[rr * kTCols + col];
                        } else {
                            const uchar4 c = reinterpret_cast<const uchar4*>(stage)[rr * kTCols + col];
                            tx = make_float4((float)c.x, (float)c.y, (float)c.z, 0.f);
                        }
                        const float4 ra = __ldg(p.row_tab + 2 * (r + rr));
                        const float4 rb4 = __ldg(p.row_tab + 2 * (r + rr) + 1);
                        const float tw[7] = {ra.x, ra.y, ra.z, ra.w, rb4.x, rb4.y, rb4.z};
#pragma unroll
                        for (int i = 0; i < NG; ++i) {
                            g[i][0] = fmaf(tw[i], tx.x, g[i][0]);
                            g[i][1] = fmaf(tw[i], tx.y, g[i][1]);
                            g[i][2] = fmaf(tw[i], tx.z, g[i][2]);
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }
        // ---- flush this unit: merge the 4 row phases, apply the phi factors, reduce over columns ----
        float* mine_g = s_g + ((size_t)ph * kTCols + col) * (NG * 3);
#pragma unroll
        for (int i = 0; i < NG; ++i) { mine_g[3 * i] = g[i][0]; mine_g[3 * i + 1] = g[i][1]; mine_g[3 * i + 2] = g[i][2]; }
        consumer_sync();
        if (tid < 3 * kTCols) {
            const int c = tid & (kTCols - 1), ch = tid / kTCols;
            ColumnMoments<NG> m;
            const float scale = FMT == VLB_FMT_RGBA8 ? (1.0f / 255.0f) : 1.0f;
#pragma unroll
            for (int i = 0; i < NG; ++i) {
                float v = 0.f;
#pragma unroll
                for (int q = 0; q < kTPhases; ++q) v += s_g[((size_t)q * kTCols + c) * (NG * 3) + 3 * i + ch];
                m.g[i] = v * scale;
            }
            float o[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = 0.f;
            if (c < t.ncols) {
                m.set_phi(__ldg(p.col_cs + t.c0 + c));
                sh_from_moments<K, NG>(m, p.variant, o);
            }
            warp_transpose_reduce<32>(o, lane);        // lane l: sum over this warp's 32 columns of coefficient l
            s_red[warp][lane] = o[0];
        }
        consumer_sync();
        float mine = 0.f;
        if (tid < K * 3) {
            const int i = tid / 3, ch = tid % 3;
            mine = s_red[2 * ch][i] + s_red[2 * ch + 1][i];
        }
        const uint32_t P = p.strips * p.row_blocks;
        publish_and_finish(p, t.map, u % P, P, mine, tid, s_fin, [] { consumer_sync(); });
    }
}

static int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

template <int K, int FMT>
static cudaError_t launch_tma(const ProjParams& p, unsigned grid, size_t smem, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(k_project_tma<K, FMT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k_project_tma<K, FMT><<<grid, kTThreads, smem, st>>>(p);
    return cudaSuccess;
}

int project_sh_device(vlb_ctx* ctx, const void* d_texels, uint64_t map_stride, uint32_t n_maps, int fmt, int W, int H,
                      int order, int variant, float* d_out) {
    cudaStream_t st = ctx->stream;
    // tables (cached per size/variant): per-row quadrature factors, per-column cos/sin(phi)
    if (ctx->tab_w != W || ctx->tab_h != H || ctx->tab_variant != variant) {
        std::vector<float> row_tab(8 * (size_t)H), row_sc(2 * (size_t)H), col(2 * (size_t)W);
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
    p.bpt = bpt;
    const bool aligned = ((size_t)W * bpt) % 16 == 0 && map_stride % 16 == 0 && (reinterpret_cast<uintptr_t>(d_texels) % 16) == 0;
    const bool use_tma = aligned && env_int("VLB_PROJ_TMA", 1) != 0;

    if (use_tma) {
        p.strips = (W + kTCols - 1) / kTCols;
        const long long n_ms = (long long)n_maps * p.strips;
        const int sms = ctx->sm_count;
        int rows = H;
        if (n_ms < sms) {
            const int rbk = std::max<long long>(1, sms / n_ms);
            rows = (H + rbk - 1) / rbk;
            rows = std::max(kTPhases, (rows + kTPhases - 1) / kTPhases * kTPhases);
        }
        rows = env_int("VLB_PROJ_ROWS", rows);
        rows = std::max(1, std::min(rows, H));
        p.rows_per_block = rows;
        p.row_blocks = (H + rows - 1) / rows;
        const uint64_t P = (uint64_t)p.strips * p.row_blocks;
        const uint64_t n_units = P * n_maps;
        if (n_units >= (1ull << 31)) return ctx->fail(VLB_ERR_UNSUPPORTED, "project_sh: too many tiles");
        p.n_units = (uint32_t)n_units;
        const int stage_bytes = kTRows * kTCols * bpt;
        const int ng3 = (order == 2 ? 5 : 7) * 3;
        const size_t merge_bytes = (size_t)kTPhases * kTCols * ng3 * sizeof(float);
        int stages = env_int("VLB_PROJ_STAGES", bpt == 16 ? 6 : kTMaxStages);
        stages = std::max(2, std::min(stages, kTMaxStages));
        p.n_stages = stages;
        const size_t smem = (size_t)stages * stage_bytes + merge_bytes;
        const unsigned grid = (unsigned)std::min<uint64_t>(n_units, (uint64_t)sms);
        VLB_CUDA(ctx, ctx->d_proj_partials.reserve(n_units * VLB_SH_STRIDE * sizeof(float)));
        if (ctx->d_proj_counters.cap < n_maps * sizeof(unsigned)) {
            VLB_CUDA(ctx, ctx->d_proj_counters.reserve(n_maps * sizeof(unsigned)));
            VLB_CUDA(ctx, cudaMemsetAsync(ctx->d_proj_counters.p, 0, ctx->d_proj_counters.cap, st));
        }
        p.partials = ctx->d_proj_partials.as<float>(); p.counters = ctx->d_proj_counters.as<unsigned>();
        cudaError_t e;
        if (order == 2) e = fmt == VLB_FMT_RGBA32F ? launch_tma<9, VLB_FMT_RGBA32F>(p, grid, smem, st) : launch_tma<9, VLB_FMT_RGBA8>(p, grid, smem, st);
        else            e = fmt == VLB_FMT_RGBA32F ? launch_tma<16, VLB_FMT_RGBA32F>(p, grid, smem, st) : launch_tma<16, VLB_FMT_RGBA8>(p, grid, smem, st);
        VLB_CUDA(ctx, e);
        VLB_LAUNCH_CHECK(ctx);
        return VLB_OK;
    }

    p.strips = (W + kProjBlock - 1) / kProjBlock;
    // launch geometry: about 2 blocks per SM for a single map, whole columns for big batches
    const long long target = (long long)ctx->sm_count * env_int("VLB_PROJ_BLOCKS_PER_SM", 2);
    long long rbw = std::max<long long>(1, target / std::max<long long>(1, (long long)n_maps * p.strips));
    int rows = (int)((H + rbw - 1) / rbw);
    rows = std::max(rows, kProjUnroll);
    rows = env_int("VLB_PROJ_ROWS", rows);
    rows = std::max(1, std::min(rows, H));
    p.rows_per_block = rows;
    p.row_blocks = (H + rows - 1) / rows;
    const uint64_t P = (uint64_t)p.strips * p.row_blocks;
    const uint64_t n_blocks = P * n_maps;
    if (n_blocks >= (1ull << 31)) return ctx->fail(VLB_ERR_UNSUPPORTED, "project_sh: launch too large");
    VLB_CUDA(ctx, ctx->d_proj_partials.reserve(n_blocks * VLB_SH_STRIDE * sizeof(float)));
    if (ctx->d_proj_counters.cap < n_maps * sizeof(unsigned)) {
        VLB_CUDA(ctx, ctx->d_proj_counters.reserve(n_maps * sizeof(unsigned)));
        VLB_CUDA(ctx, cudaMemsetAsync(ctx->d_proj_counters.p, 0, ctx->d_proj_counters.cap, st));
    }
    p.partials = ctx->d_proj_partials.as<float>(); p.counters = ctx->d_proj_counters.as<unsigned>();
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

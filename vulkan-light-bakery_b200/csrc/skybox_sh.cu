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
// G[p][q] (q in {0,1}) are needed. A thread owns ONE column: per texel it does one 16-byte load
// and 15 / 21 FMAs with warp-uniform row factors; the phi factors and the SH polynomials are
// applied once per thread at the end. Reduction: halving-exchange warp reduce -> shared memory ->
// one 48-float partial per block -> the last block of each map sums the partials in a fixed order
// (no float atomics: results are bitwise reproducible for a given launch geometry).
//
// Algorithmic bytes per texel: 16 (RGBA32F) or 4 (RGBA8) read; 192 bytes written per map.
#include <algorithm>

#include "vlb_context.h"
#include "vlb_math.cuh"
#include "vlb_warp.cuh"

namespace vlb {

constexpr int kProjBlock = 256;
constexpr int kProjUnroll = 8;

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
    // moment in d-space
    __device__ __forceinline__ float Md(int a, int b, int c) const { return cp[a] * sp[b] * H(a + b, c); }
    // moment of sx^a sy^b sz^c where s = d.xzy (variant 0) or s = d (variant 1)
    __device__ __forceinline__ float M(int variant, int a, int b, int c) const {
        return variant == 0 ? Md(a, c, b) : Md(a, b, c);
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

template <int K, int FMT>
__global__ void __launch_bounds__(kProjBlock) k_project(const ProjParams p) {
    constexpr int NG = K > 9 ? 7 : 5;
    constexpr int V = (K * 3 <= 32) ? 32 : 64;
    constexpr int R = V / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t bid = blockIdx.x;
    const int rb = bid % p.row_blocks;
    const int strip = (bid / p.row_blocks) % p.strips;
    const uint32_t map = bid / (p.row_blocks * p.strips);
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
            m.cp[0] = 1.f; m.cp[1] = cs.x; m.cp[2] = cs.x * cs.x; m.cp[3] = m.cp[2] * cs.x;
            m.sp[0] = 1.f; m.sp[1] = cs.y; m.sp[2] = cs.y * cs.y; m.sp[3] = m.sp[2] * cs.y;
            float o[K];
            sh_from_moments<K, NG>(m, p.variant, o);
#pragma unroll
            for (int i = 0; i < K; ++i) acc[3 * i + ch] = o[i];
        }
    }

    // block reduction: warp halving exchange, then the 8 warp rows in shared memory
    __shared__ float s_red[kProjBlock / 32][V];
    __shared__ float s_grp[4][VLB_SH_STRIDE];
    __shared__ bool s_last;
    warp_transpose_reduce<V>(acc, lane);
#pragma unroll
    for (int r = 0; r < R; ++r) s_red[warp][R * lane + r] = acc[r];
    __syncthreads();
    const uint32_t P = p.strips * p.row_blocks;
    float* my_partial = p.partials + ((size_t)map * P + (bid % P)) * VLB_SH_STRIDE;
    if (tid < VLB_SH_STRIDE) {
        float s = 0.f;
        if (tid < K * 3) {
#pragma unroll
            for (int w = 0; w < kProjBlock / 32; ++w) s += s_red[w][tid];
        }
        my_partial[tid] = s;
        __threadfence();
    }
    __syncthreads();
    if (tid == 0) {
        const unsigned prev = atomicAdd(p.counters + map, 1u);
        s_last = prev == P - 1;
        if (s_last) p.counters[map] = 0;   // self-reset for the next launch
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // last block of this map: sum the P partials in a fixed order (4 interleaved groups)
    const float* base = p.partials + (size_t)map * P * VLB_SH_STRIDE;
    if (tid < 4 * VLB_SH_STRIDE) {
        const int grp = tid / VLB_SH_STRIDE, c = tid % VLB_SH_STRIDE;
        float s = 0.f;
        for (uint32_t k = grp; k < P; k += 4) s += __ldcg(base + (size_t)k * VLB_SH_STRIDE + c);
        s_grp[grp][c] = s;
    }
    __syncthreads();
    if (tid < VLB_SH_STRIDE)
        p.out[(size_t)map * VLB_SH_STRIDE + tid] = (s_grp[0][tid] + s_grp[1][tid]) + (s_grp[2][tid] + s_grp[3][tid]);
}

static int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v && *v ? atoi(v) : dflt;
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
    p.row_tab = ctx->d_row_tab.as<float4>(); p.col_cs = ctx->d_col_tab.as<float2>();
    p.partials = ctx->d_proj_partials.as<float>(); p.counters = ctx->d_proj_counters.as<unsigned>();
    p.out = d_out; p.variant = variant;
    const unsigned grid = (unsigned)n_blocks;
    if (order == 2) {
        if (fmt == VLB_FMT_RGBA32F) k_project<9, VLB_FMT_RGBA32F><<<grid, kProjBlock, 0, st>>>(p);
        else                        k_project<9, VLB_FMT_RGBA8><<<grid, kProjBlock, 0, st>>>(p);
    } else {
        if (fmt == VLB_FMT_RGBA32F) k_project<16, VLB_FMT_RGBA32F><<<grid, kProjBlock, 0, st>>>(p);
        else                        k_project<16, VLB_FMT_RGBA8><<<grid, kProjBlock, 0, st>>>(p);
    }
    VLB_LAUNCH_CHECK(ctx);
    return VLB_OK;
}

}  // namespace vlb

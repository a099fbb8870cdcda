// diag.cu -- measured ceilings for the roofline of the traversal kernel ("bytes per ray divided into achieved
// L2 bandwidth"). k_bake_stream is bound by the L1 data pipe, not by HBM, and its loads are scattered 16-byte pieces of
// L1/L2-resident nodes, so neither the HBM copy peak nor a nominal 128 B/clk/SM says what the hardware can deliver to such
// a kernel. Three micro-kernels measure it on the device the context owns:
//   L2 stream     every SM streams one L2-resident buffer (48 MB) with 16-byte ld.global.cg (L1 bypassed)
//   L1 stream     every block re-reads its own 32 KB (L1-resident) with coalesced 16-byte loads
//   L1 scattered  the traversal's access shape: per load instruction the 32 lanes fall on 8 distinct 128-byte lines of an
//                 L1-resident region (4 lanes share one 16-byte address, like rays on the same node), lines picked
//                 pseudo-randomly -- what ncu counts as 8 wavefronts per request
// Diagnostics only: nothing on the bake or projection path calls this file.
#include "vlb_context.h"

namespace vlb {

__device__ __forceinline__ float4 ld_ca(const float4* p) {
    float4 v;
    asm volatile("ld.global.ca.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_cg(const float4* p) {
    float4 v;
    asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}

// grid-stride stream over n_quads float4s, `passes` times, L1 bypassed
__global__ void __launch_bounds__(512) k_diag_l2_stream(const float4* __restrict__ buf, uint32_t n_quads, int passes, float* sink) {
    float acc = 0.f;
    const uint32_t stride = gridDim.x * blockDim.x;
    for (int it = 0; it < passes; ++it) {
        uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
        for (; i + 7u * stride < n_quads; i += 8u * stride) {
            float4 v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = ld_cg(buf + i + k * stride);
#pragma unroll
            for (int k = 0; k < 8; ++k) acc += v[k].x + v[k].w;
        }
        for (; i < n_quads; i += stride) acc += ld_cg(buf + i).x;
    }
    if (acc == 123.456f) *sink = acc;
}

constexpr int kDiagRegionQuads = 2048;        // 32 KB per block
// every block re-reads its own region, coalesced (thread t reads quads t, t + blockDim, ...)
__global__ void __launch_bounds__(512) k_diag_l1_stream(const float4* __restrict__ buf, int iters, float* sink) {
    const float4* reg = buf + (size_t)blockIdx.x * kDiagRegionQuads;
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
        float4 v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = ld_ca(reg + ((threadIdx.x + k * 512 + it * 64) & (kDiagRegionQuads - 1)));
#pragma unroll
        for (int k = 0; k < 4; ++k) acc += v[k].x + v[k].w;
    }
    if (acc == 123.456f) *sink = acc;
}

// the traversal's access shape: `lines` distinct 128-byte lines per load instruction, 32 / lines lanes per address
__global__ void __launch_bounds__(128) k_diag_l1_scatter(const float4* __restrict__ buf, int iters, int lines, float* sink) {
    const float4* reg = buf + (size_t)blockIdx.x * kDiagRegionQuads;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t group = (uint32_t)(lane * lines) >> 5;             // lanes of one group share an address
    uint32_t h = (blockIdx.x * 4u + warp) * 2654435761u + group * 40503u;
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
        float4 v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            h = h * 1664525u + 1013904223u;
            const uint32_t line = (h >> 12) & (kDiagRegionQuads / 8 - 1), quad = (h >> 7) & 7u;
            v[k] = ld_ca(reg + line * 8 + quad);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) acc += v[k].x + v[k].w;
    }
    if (acc == 123.456f) *sink = acc;
}

}  // namespace vlb

using namespace vlb;

extern "C" int vlb_diag_cache_peaks(vlb_ctx* ctx, vlb_cache_peaks* out) {
    if (!ctx || !out) return VLB_ERR_INVALID;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return ctx->fail(VLB_ERR_CUDA, "vlb_diag_cache_peaks: cudaSetDevice failed");
    *out = vlb_cache_peaks{};
    cudaStream_t st = ctx->stream;
    const size_t l2_bytes = 48ull << 20;
    const uint32_t n_quads = (uint32_t)(l2_bytes / 16);
    struct Scratch {                 // released on every path out of this function
        DevBuf buf;
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        ~Scratch() { if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1); buf.release(); }
    } sc;
    DevBuf& buf = sc.buf;
    VLB_CUDA(ctx, buf.reserve(l2_bytes + 64));
    VLB_CUDA(ctx, cudaMemsetAsync(buf.p, 0, l2_bytes, st));
    float* sink = reinterpret_cast<float*>(static_cast<char*>(buf.p) + l2_bytes);
    VLB_CUDA(ctx, cudaEventCreate(&sc.e0));
    VLB_CUDA(ctx, cudaEventCreate(&sc.e1));
    cudaEvent_t e0 = sc.e0, e1 = sc.e1;
    float ms = 0.f;
    const int sms = ctx->sm_count;
    auto timed = [&](auto&& launch) -> cudaError_t {
        launch();                                   // warm-up: brings the data into the cache under test
        cudaError_t e = cudaEventRecord(e0, st);
        launch();
        if (e == cudaSuccess) e = cudaEventRecord(e1, st);
        if (e == cudaSuccess) e = cudaEventSynchronize(e1);
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
        if (e == cudaSuccess) e = cudaGetLastError();
        return e;
    };
    // L2 -> SM, streaming
    const int passes = 40;
    VLB_CUDA(ctx, timed([&] { k_diag_l2_stream<<<sms * 4, 512, 0, st>>>(buf.as<float4>(), n_quads, passes, sink); }));
    out->l2_read_gbs = (double)l2_bytes * passes / (ms * 1e-3) / 1e9;
    // L1 -> SM, streaming: 4 blocks of 512 threads per SM, 32 KB each
    const int it1 = 4000;
    VLB_CUDA(ctx, timed([&] { k_diag_l1_stream<<<sms * 4, 512, 0, st>>>(buf.as<float4>(), it1, sink); }));
    out->l1_read_gbs = (double)sms * 4 * 512 * 4 * 16 * it1 / (ms * 1e-3) / 1e9;
    // L1 -> SM, the traversal's scattered shape: 8 blocks of 128 threads per SM (as k_bake_stream), 8 lines per request
    const int it2 = 2000, lines = 8;
    VLB_CUDA(ctx, timed([&] { k_diag_l1_scatter<<<sms * 8, 128, 0, st>>>(buf.as<float4>(), it2, lines, sink); }));
    const double requests = (double)sms * 8 * 4 * 8 * it2;        // warp-level load instructions
    out->l1_scatter_lines_per_request = lines;
    out->l1_scatter_requests_per_s = requests / (ms * 1e-3);
    out->l1_scatter_wavefronts_per_s = out->l1_scatter_requests_per_s * lines;
    out->l1_scatter_gbs = out->l1_scatter_requests_per_s * 32 * 16 / 1e9;   // bytes the 32 lanes receive
    ctx->launches += 6;
    return VLB_OK;
}

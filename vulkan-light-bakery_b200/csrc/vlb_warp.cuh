// vlb_warp.cuh — warp-level reduction shared by the bake epilogue and the skybox projection.
#pragma once

namespace vlb {

// Sum V = 32*R per-lane values over the 32 lanes of a warp with V-R shuffles (halving
// exchange): afterwards lane l holds the warp-wide sums of original indices R*l .. R*l+R-1 in
// v[0..R). The exchange pattern is fixed, so the result is bitwise reproducible.
template <int V>
__device__ __forceinline__ void warp_transpose_reduce(float (&v)[V], int lane) {
    int n = V;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const int h = n >> 1;
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < V / 2; ++i) {
            if (i < h) {
                const float a = v[i], b = v[i + h];
                const float send = upper ? a : b;
                const float keep = upper ? b : a;
                v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
            }
        }
        n = h;
    }
}

}  // namespace vlb

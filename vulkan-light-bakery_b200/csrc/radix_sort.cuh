// radix_sort.cuh — stable LSD radix sort of (64-bit key, 32-bit value) pairs for the LBVH build
// (Morton code, triangle id): the "radix sort" step of Scene_t::buildAccelerationStructures' replacement
// (csrc/bvh_build.cu). Hand-written for sm_100a; no library on the path.
//
// 8 bits per pass. Per pass, three launches over tiles of 2,048 pairs:
//   k_rs_hist     per-tile digit histogram (shared-memory atomics) -> hist[digit][tile], and the digit totals
//   k_rs_scan     one block per digit: exclusive scan of its row of tile counts, offset by the totals of all
//                 smaller digits -> the first output slot of every (digit, tile) pair
//   k_rs_scatter  re-reads the tile, ranks every pair among the tile's pairs of the same digit in input order
//                 (warp match_any + a scan over the tile's 64 warp-rows), and writes it to its slot
// HBM traffic per pass: 12 B read (histogram) + 12 B read + 12 B written (scatter) per pair. The sort is stable,
// so equal Morton codes keep their triangle order and the build is deterministic.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace vlb {

constexpr int kRsThreads = 256;
constexpr int kRsItems = 8;                          // pairs per thread
constexpr int kRsTile = kRsThreads * kRsItems;       // 2,048 pairs per tile
constexpr int kRsRows = kRsItems * (kRsThreads / 32);   // warp-rows of 32 consecutive pairs per tile: 64
constexpr int kRsDigits = 256;

__device__ __forceinline__ uint32_t rs_digit(uint64_t key, int shift) { return (uint32_t)(key >> shift) & 255u; }

__global__ void __launch_bounds__(kRsThreads) k_rs_hist(const uint64_t* __restrict__ keys, uint32_t n, int shift, uint32_t n_tiles,
                                                         uint32_t* __restrict__ hist, uint32_t* __restrict__ totals) {
    __shared__ uint32_t s_hist[kRsDigits];
    const uint32_t tile = blockIdx.x;
    s_hist[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t base = tile * kRsTile;
#pragma unroll
    for (int i = 0; i < kRsItems; ++i) {
        const uint32_t idx = base + i * kRsThreads + threadIdx.x;
        if (idx < n) atomicAdd(&s_hist[rs_digit(keys[idx], shift)], 1u);
    }
    __syncthreads();
    const uint32_t c = s_hist[threadIdx.x];
    hist[(size_t)threadIdx.x * n_tiles + tile] = c;
    if (c) atomicAdd(&totals[threadIdx.x], c);
}

// Block d: hist[d][t] <- (pairs with a smaller digit) + (pairs with digit d in tiles before t).
__global__ void __launch_bounds__(kRsThreads) k_rs_scan(uint32_t* __restrict__ hist, const uint32_t* __restrict__ totals, uint32_t n_tiles) {
    __shared__ uint32_t s_warp[kRsThreads / 32];
    __shared__ uint32_t s_carry;
    const uint32_t d = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // totals of the smaller digits
    uint32_t v = tid < d ? totals[tid] : 0u;
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if (lane == 0) s_warp[warp] = v;
    __syncthreads();
    if (tid == 0) {
        uint32_t s = 0;
        for (int w = 0; w < kRsThreads / 32; ++w) s += s_warp[w];
        s_carry = s;
    }
    __syncthreads();
    uint32_t* row = hist + (size_t)d * n_tiles;
    for (uint32_t t0 = 0; t0 < n_tiles; t0 += kRsThreads) {
        const uint32_t t = t0 + tid;
        const uint32_t c = t < n_tiles ? row[t] : 0u;
        uint32_t x = c;                                   // inclusive scan inside the warp
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, off);
            if (lane >= (uint32_t)off) x += y;
        }
        if (lane == 31) s_warp[warp] = x;
        __syncthreads();
        uint32_t before = s_carry;
        for (uint32_t w = 0; w < warp; ++w) before += s_warp[w];
        if (t < n_tiles) row[t] = before + x - c;
        __syncthreads();
        if (tid == kRsThreads - 1) s_carry = before + x;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kRsThreads) k_rs_scatter(const uint64_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                                                            uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, uint32_t n,
                                                            int shift, uint32_t n_tiles, const uint32_t* __restrict__ offsets) {
    // s_cnt[row][digit]: pairs of that digit in warp-row `row`, then (after the scan) in the rows before it
    __shared__ uint16_t s_cnt[kRsRows][kRsDigits];
    __shared__ uint32_t s_base[kRsDigits];
    const uint32_t tile = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < kRsRows * kRsDigits / 2; i += kRsThreads) reinterpret_cast<uint32_t*>(&s_cnt[0][0])[i] = 0u;
    s_base[tid] = offsets[(size_t)tid * n_tiles + tile];
    __syncthreads();
    const uint32_t base = tile * kRsTile;
    uint64_t key[kRsItems];
    uint32_t rank_in_row[kRsItems];
#pragma unroll
    for (int i = 0; i < kRsItems; ++i) {
        const uint32_t idx = base + i * kRsThreads + tid;
        const bool live = idx < n;
        key[i] = live ? keys_in[idx] : ~0ull;
        const uint32_t dg = rs_digit(key[i], shift);
        // lanes of this warp-row holding the same digit (dead lanes vote in their own group)
        const unsigned peers = __match_any_sync(0xffffffffu, live ? dg : 256u + lane);
        rank_in_row[i] = __popc(peers & ((1u << lane) - 1u));
        if (live && rank_in_row[i] == 0) s_cnt[i * (kRsThreads / 32) + warp][dg] = (uint16_t)__popc(peers);
    }
    __syncthreads();
    {   // thread = digit: exclusive scan down the 64 rows
        uint32_t run = 0;
#pragma unroll 8
        for (int r = 0; r < kRsRows; ++r) {
            const uint32_t c = s_cnt[r][tid];
            s_cnt[r][tid] = (uint16_t)run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kRsItems; ++i) {
        const uint32_t idx = base + i * kRsThreads + tid;
        if (idx < n) {
            const uint32_t dg = rs_digit(key[i], shift);
            const uint32_t dst = s_base[dg] + s_cnt[i * (kRsThreads / 32) + warp][dg] + rank_in_row[i];
            keys_out[dst] = key[i];
            vals_out[dst] = vals_in[idx];
        }
    }
}

// Scratch needed by radix_sort_pairs for n pairs, in bytes.
inline size_t radix_sort_scratch_bytes(uint32_t n) {
    const size_t n_tiles = ((size_t)n + kRsTile - 1) / kRsTile;
    return (kRsDigits * n_tiles + 8 * kRsDigits) * sizeof(uint32_t);
}

// Sorts n pairs by bits [0, 8 * n_passes) of the key, ping-ponging between the (a) and (b) buffers; the input is in
// (a). Returns 0 if the result is in (a), 1 if it is in (b). Everything is enqueued on `st`; nothing synchronises.
inline int radix_sort_pairs(uint64_t* keys_a, uint32_t* vals_a, uint64_t* keys_b, uint32_t* vals_b, uint32_t n, int n_passes,
                            void* scratch, cudaStream_t st, uint64_t* launches) {
    if (n == 0) return 0;
    const uint32_t n_tiles = (n + kRsTile - 1) / kRsTile;
    uint32_t* hist = static_cast<uint32_t*>(scratch);
    uint32_t* totals = hist + (size_t)kRsDigits * n_tiles;             // [pass][256]
    cudaMemsetAsync(totals, 0, (size_t)n_passes * kRsDigits * sizeof(uint32_t), st);
    int cur = 0;
    for (int pass = 0; pass < n_passes; ++pass) {
        const uint64_t* kin = cur ? keys_b : keys_a; const uint32_t* vin = cur ? vals_b : vals_a;
        uint64_t* kout = cur ? keys_a : keys_b;      uint32_t* vout = cur ? vals_a : vals_b;
        uint32_t* tot = totals + (size_t)pass * kRsDigits;
        k_rs_hist<<<n_tiles, kRsThreads, 0, st>>>(kin, n, 8 * pass, n_tiles, hist, tot);
        k_rs_scan<<<kRsDigits, kRsThreads, 0, st>>>(hist, tot, n_tiles);
        k_rs_scatter<<<n_tiles, kRsThreads, 0, st>>>(kin, vin, kout, vout, n, 8 * pass, n_tiles, hist);
        if (launches) *launches += 3;
        cur ^= 1;
    }
    return cur;
}

}  // namespace vlb

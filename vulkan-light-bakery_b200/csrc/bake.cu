// bake.cu — the probe bake: replaces the reference's per-probe pair
//   EnvMapGenerator::getMap  (src/baker/env_map_generator.cpp:325-359; shaders env_map.rgen/.rchit,
//                             main.rmiss, shadow.rmiss)          and
//   LightBaker::dispatchBakingKernel (src/baker/light_baker.cpp:269-285; shaders/sh.comp)
// with ONE persistent launch over all probes of the slab: every warp pulls (probe, direction
// chunk) items from a global counter, fires the probe's fixed equirect directions plus shadow
// rays through the software LBVH, and projects the radiance onto SH in registers. Per-ray
// radiance never goes to HBM; the only global writes are 192 bytes per probe.
//
// Two translation units are made of this file. bake.cu itself holds the direct-pass instantiations of k_bake_stream and
// everything on the host side; bake_gather.cu includes it with VLB_BAKE_GATHER_TU to compile the gather-pass
// instantiations (GATHER = true) alone, with the cold shading code's divisions and square roots out of line
// (VLB_COLD_OUTLINE): the gather kernel is half as large again and gains 6 % from the smaller code, the direct kernel
// loses 1.6 % to the calls (profiles/r02_bake_icache_ab.log, same box).
#ifndef VLB_BAKE_GATHER_TU
#define VLB_BAKE_GATHER_TU 0
#endif
#ifndef VLB_COLD_OUTLINE
#define VLB_COLD_OUTLINE VLB_BAKE_GATHER_TU
#endif
// Per-direction tables (k_dir_tables): compiled into the direct-pass kernels only. C3 -0.85 %, C2 -2.3 %, the direct pass of
// C4 -2.6 %; the gather kernel, whose time is its visibility rays, lost 0.7 % to the extra code (profiles/r02_bake_dir_tables_ab.log).
#ifndef VLB_DIR_TABLES
#define VLB_DIR_TABLES (!VLB_BAKE_GATHER_TU)
#endif
// Two small layout choices measured together (profiles/r02_bake_dir_tables_ab.log): shadow-ray queue records as three float4s
// instead of eleven scalars, and 1 / direction carried by the direction table. Each alone lost ~0.7 % at C3, both together
// gain 1.25 % (136.6 -> 134.9 ms) -- the kernel's register allocation under the 64-register cap decides, not the idea.
#ifndef VLB_SQ_PACKED
#define VLB_SQ_PACKED 1
#endif
#ifndef VLB_IDIR_TAB
#define VLB_IDIR_TAB 1
#endif
#include <algorithm>
#include <cstring>

#include "vlb_context.h"
#include "vlb_shade.cuh"
#include "vlb_warp.cuh"

namespace vlb {

constexpr int kBakeBlock = 128;
// One warp-sized direction tile = 2^lw x 2^(5-lw) adjacent texels of the equirect direction grid. The shape is picked
// per direction grid so that the tile is as square as possible in ANGLE (a texel spans 360/W x 180/H degrees): 4x8
// texels for W = H (the 32x32 / 64x64 grids of the BASELINE configs; measured +1.4 % over 8x4 on C3), 8x4 for
// W >= 2H (the reference's 3141x1000). Rays in flight in a warp then cover the smallest solid angle.
// texel (x, y) of lane slot w of a tile: x = ((tile % tiles_x) << lw) + (w & (2^lw - 1)), y = ((tile / tiles_x) << (5 - lw)) + (w >> lw)
// -- tile_xy() below, which divides with the host's reciprocal (an integer division is ~20 instructions, and the kernel
// needs the pair in four places, one of them per ray).

struct WarpSpill;
struct BakeParams {
    BvhView bvh;
    ShadeView shade;
    BakeConsts c;
    const float* px; const float* py; const float* pz;   // probe axis coordinates
    const float2* row_sc;                                // (sin, cos) theta per direction row
    const float2* col_cs;                                // (cos, sin) phi per direction column
    int Nx, Ny, Nz, k0, kstride;   // the call bakes z-slices k0, k0 + kstride, ...
    int W, H, tiles_x, n_tiles, tile_lw;   // tile_lw: log2 of the direction tile's width
    // Probe-independent per-direction work, hoisted into tables (k_dir_tables; NULL: computed in place, large direction grids):
    const float4* dir_tab;                 // [n_tiles * 32]: unit ray direction (env_map.rgen:21) of slot w of a tile, w = 1 inside the grid / 0 outside
    const float* proj_tab;                 // [n_tiles][K][32]: SH basis x quadrature weight of the slot's direction (sh.comp:30-39)
    unsigned long long tiles_x_rcp;        // floor(2^64 / tiles_x) + 1: umul64hi(tile, rcp) == tile / tiles_x for every 32-bit tile (tiles_x >= 2)
    int chunks, tiles_per_chunk;
    uint32_t n_items;
    uint32_t n_whole;            // items [0, n_whole) are whole probes; item n_whole + j is chunk run j % chunks of probe n_whole + j / chunks
    float pixel_area;
    float* out;                  // final [slot][48]
    float* partials;             // [item - n_whole][48] of the chunk-run items
    unsigned int* work_counter;
    unsigned long long* stats;   // [0] shadow rays, [1] nodes visited, [2] triangles tested, [4..11] COUNT builds: phase utilisation
    int ref_order, world_frame;
    void* stream_scratch;         // k_bake_stream: per-warp direction slots (hit record, then radiance) + shadow-ray queue, [grid * warps per block]
    WarpSpill* stream_spill;     // k_bake_stream: per-warp stack overflow rows (rarely touched), same indexing
    int node_min;                // k_bake_stream: the node loop yields to the leaf phase below this many lanes
    int leaf_min;                // ... and at least this many lanes wait at a leaf
    int refill_min;              // k_bake_stream: idle lanes are refilled once there are this many of them (or all)
    int refill_order;            // k_bake_stream: 0 shadow rays first, 1 homogeneous refill batches
    int vis_refill_min;          // gather passes: the same threshold inside a visibility-ray batch
    GatherView g;                // gather pass source (g.prev == NULL: direct pass)
    int* vis_ovf;                // gather passes: stack overflow slab of the visibility-ray batches, [grid * warps][kOvfStack][32]
};

constexpr int kDirTabQuads = VLB_IDIR_TAB ? 2 : 1;   // float4s per slot of the direction table
__device__ __forceinline__ void tile_xy(const BakeParams& p, int tile, int w, int& x, int& y) {
    // exact: rcp = (2^64 + e) / d with 0 < e <= d, so tile * rcp / 2^64 = tile / d + tile * e / (d * 2^64) and tile * e < 2^64
    const uint32_t row = p.tiles_x == 1 ? (uint32_t)tile : (uint32_t)__umul64hi((unsigned long long)(uint32_t)tile, p.tiles_x_rcp);
    const uint32_t col = (uint32_t)tile - row * (uint32_t)p.tiles_x;
    x = (int)(col << p.tile_lw) + (w & ((1 << p.tile_lw) - 1));
    y = (int)(row << (5 - p.tile_lw)) + (w >> p.tile_lw);
}

// unit ray direction of slot w of a tile (env_map.rgen:19-21): from the table, or from the sin / cos tables
__device__ __forceinline__ Vec3 slot_direction(const BakeParams& p, int tile, int w) {
    if (VLB_DIR_TABLES && p.dir_tab) {
        const float4 dt = __ldg(p.dir_tab + ((size_t)tile * 32 + w) * kDirTabQuads);
        return mk3(dt.x, dt.y, dt.z);
    }
    int x, y;
    tile_xy(p, tile, w, x, y);
    const float2 row = __ldg(p.row_sc + y), col = __ldg(p.col_cs + x);
    const Vec3 t = to_vector_sc(row.x, row.y, col.x, col.y);   // sh_common.h:8-12
    return mk3(t.x, t.z, t.y);                                 // .xzy
}

#if !VLB_BAKE_GATHER_TU
// The per-direction work that does not depend on the probe -- the ray direction and the projection weights basis x
// sin(theta) x pixel area -- is the same for all probes of a call (131,072 at C3): computed once here with the very
// expressions the kernel would use in place, so the bits are the same.
template <int K>
__global__ void k_dir_tables(const BakeParams p, float4* dir_tab, float* proj_tab) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= p.n_tiles * 32) return;
    const int tile = d >> 5, w = d & 31;
    int x, y;
    tile_xy(p, tile, w, x, y);
    const bool inside = x < p.W && y < p.H;
    float4 dt = make_float4(0.f, 0.f, 0.f, 0.f);
    float b[K];
#pragma unroll
    for (int i = 0; i < K; ++i) b[i] = 0.f;
    float wgt = 0.f;
    if (inside) {
        const float2 row = __ldg(p.row_sc + y), col = __ldg(p.col_cs + x);
        const Vec3 t = to_vector_sc(row.x, row.y, col.x, col.y);
        dt = make_float4(t.x, t.z, t.y, 1.0f);
        wgt = p.pixel_area * row.x;                                         // sh.comp:32-33
        sh_basis<K>(p.world_frame ? mk3(t.x, t.z, t.y) : t, b);             // sh.comp:30,39
    }
    dir_tab[(size_t)d * kDirTabQuads] = dt;
#if VLB_IDIR_TAB
    dir_tab[(size_t)d * kDirTabQuads + 1] = make_float4(safe_inv(dt.x), safe_inv(dt.y), safe_inv(dt.z), 0.f);
#endif
#pragma unroll
    for (int i = 0; i < K; ++i) proj_tab[(size_t)tile * (K * 32) + 32 * i + w] = b[i] * wgt;
}
#endif

__device__ __forceinline__ size_t out_slot(const BakeParams& p, uint32_t q) {
    if (!p.ref_order) return q;
    const int i = q % p.Nx, j = (q / p.Nx) % p.Ny, k = q / (p.Nx * p.Ny);
    // writer order of LightBaker::probePositionsFromBoudingBox (light_baker.cpp:87-98)
    const size_t s = (j == 0) ? (size_t)i : (size_t)p.Nx + (size_t)i * (p.Ny - 1) + (j - 1);
    return (k == 0) ? s : (size_t)p.Nx * p.Ny + s * (p.Nz - 1) + (k - 1);
}

// =========================================================================================
// k_bake_stream — the production bake kernel: a warp is a small wavefront path tracer.
//
// A warp owns a work item = (probe, run of direction chunks; a chunk = kChunkTiles tiles of 32 directions). Inside a
// chunk every LANE is a ray slot; idle lanes are refilled in batches (ballot + popc ranks; directions in tile order
// (32-texel tiles, square in angle: tile_x / tile_y) so the rays in flight stay angularly adjacent) and traverse the
// BVH. Primary (closest-hit) and shadow (any-hit) rays share one "while-while" loop: all lanes step through internal
// nodes until most of them hold a leaf, then the leaves are intersected. A chunk of full size runs in two phases
// (refill order 3): A -- the chunk's closest-hit rays, nothing else in flight; a finished ray only drops its hit record
// into its direction's slot of the warp's scratch, queues the direction index and frees its lane. B -- 32 queued hits
// at a time are shaded by the whole warp (env_map.rchit / main.rmiss arithmetic at full SIMD width) whenever the shadow
// queue cannot fill every idle lane; shading pushes the shadow rays that are needed, and the any-hit batches run as
// full and as homogeneous as the closest-hit ones. (Short chunks interleave the two kinds, order 1: the drain between
// phases would cost more than the mix.) Shading stores the radiance of the OCCLUDED outcome in the direction's slot and
// hands the lit radiance to the shadow ray, which overwrites the slot when it escapes. When the chunk is drained the
// warp projects its radiances onto SH cooperatively (lane l takes directions l, l+32, ... in a fixed order, so the
// result is bitwise reproducible whatever the run-time ray scheduling was), reduces with shuffles and keeps one or two
// running coefficients per lane. Radiance never goes to HBM by design; the result is 192 bytes per probe.
//
// Where the per-ray state lives (round 2; round 1 kept all of it in local / global memory behind the
// L1 LSU pipe, which ncu showed to be the binding unit):
//   traversal stack   the top kSmemStack entries of every lane in SHARED memory, laid out [entry][lane] so
//                     that lane l only ever touches bank l: a push or pop is one conflict-free wavefront
//                     however far the lanes' depths have diverged (a per-thread local array costs one
//                     wavefront per distinct depth). Deeper entries overflow to a global scratch slab
//                     (rare: the short stack covers > 99 % of pushes on the BASELINE scenes).
//   hit queue         the finishing ORDER (16-bit direction indices) in shared memory.
//   direction slots   global scratch, L1/L2-resident, 16 bytes per direction of the chunk: the hit record until the
//                     hit is shaded, the radiance from then on (one 128-bit access either way).
//   shadow-ray queue  global scratch, L1/L2-resident, coalesced accesses.
// The per-ray arithmetic is exactly that of probe_ray_radiance (vlb_shade.cuh).
// =========================================================================================
#ifndef VLB_BAKE_CHUNK_TILES
#define VLB_BAKE_CHUNK_TILES 16             // direction tiles (of 32 rays) traced between two drains of the warp
#endif
constexpr int kChunkTiles = VLB_BAKE_CHUNK_TILES;
constexpr int kChunkDirs = kChunkTiles * 32;
constexpr int kRayDone = kNoChild;          // traversal finished
constexpr int kHitCap = kChunkDirs + 32;    // queued hits per warp: every direction of the chunk (deferred shading) + the 32 the invariant check allows for
#ifndef VLB_BAKE_SHADOW_CAP
#define VLB_BAKE_SHADOW_CAP 64
#endif
constexpr int kShadowCap = VLB_BAKE_SHADOW_CAP;   // queued shadow rays per warp (shading needs 32 free slots)
constexpr int kStreamWarps = kBakeBlock / 32;
#ifndef VLB_BAKE_SMEM_STACK
#define VLB_BAKE_SMEM_STACK 12              // stack entries per lane in shared memory; 0 = round-1 per-thread local array
#endif
#ifndef VLB_BAKE_FAST_PUSH
#define VLB_BAKE_FAST_PUSH 1                // branch-free end of a node step in the shared short stack (WarpStack::advance)
#endif
constexpr int kSmemStack = VLB_BAKE_SMEM_STACK;
constexpr int kOvfStack = kSmemStack > 0 ? kStackSize - kSmemStack : 1;

// Hit queue: the ORDER in which the chunk's closest-hit rays finished, as direction indices (16 bits each, shared
// memory); the hit record itself -- flat triangle id (-1: miss), t, u, v -- waits in the direction's slot of the warp's
// global scratch (WarpQueues::slot) until it is shaded, and the radiance then takes its place.
struct HitQueue {
    unsigned short dir[kHitCap];
};
#ifndef VLB_BAKE_DISCARD
#define VLB_BAKE_DISCARD 1                  // discard.global.L2 of the direction slots once a chunk is projected
#endif
// Loads of the per-warp scratch (written by other lanes of the warp): with VLB_L1_HINTS through L2 only, so that the
// queues and direction slots do not compete with the BVH for L1 lines.
template <class T> __device__ __forceinline__ T ld_scratch(const T* p) {
#if VLB_L1_HINTS
    return __ldcg(p);
#else
    return *p;
#endif
}
// PACKED: the shadow-ray queue records as three float4s (the direct-pass kernels) or as eleven scalar arrays (the gather
// kernels, which measured 2 % slower with the packed form); everything else is the same.
template <bool PACKED> struct WarpQueuesT;
template <> struct alignas(128) WarpQueuesT<true> {
    static constexpr bool kPacked = true;
    // one slot per direction of the current chunk: first the hit record of its closest-hit ray (bits(id), t, u, v), from
    // the moment the hit is shaded its radiance (r, g, b, -). Whole 128-byte lines: see the discard below.
    float4 slot[kChunkDirs];
    // shadow-ray queue: (origin, length), (unit direction, bits(direction index)), (radiance if the light is visible, -)
    float4 sq_a[kShadowCap], sq_b[kShadowCap], sq_c[kShadowCap];
};
template <> struct alignas(128) WarpQueuesT<false> {
    static constexpr bool kPacked = false;
    float4 slot[kChunkDirs];
    // shadow-ray queue: origin, unit direction, length, direction index, radiance if the light is visible
    float sq_o[3][kShadowCap], sq_d[3][kShadowCap], sq_len[kShadowCap]; int sq_dir[kShadowCap];
    float sq_rgb[3][kShadowCap];
};
constexpr size_t kWarpQueuesBytes = sizeof(WarpQueuesT<true>) > sizeof(WarpQueuesT<false>) ? sizeof(WarpQueuesT<true>) : sizeof(WarpQueuesT<false>);
// The cold part of a warp's scratch, in a slab of its own so that the hot part above is one dense range of a few tens
// of MB (the L2 access-policy window of the launch, bake_device).
struct alignas(128) WarpSpill {
    int ovf[kOvfStack * ((VLB_STACK_CULL || VLB_BVH8) ? 2 : 1)][32];   // stack entries beyond the shared-memory short stack, [entry][word][lane]
};

// Short stack in shared memory + overflow in global scratch; same interface as LocalStack (vlb_bvh.cuh).
// `sm` is the 32-bit shared-window address of this lane's entry 0; it is produced by an opaque asm move so that
// the compiler keeps it in a register instead of re-deriving it (S2R tid, shifts, IMAD: ten instructions) at every
// push and pop, which is what it does with a plain pointer under the kernel's 64-register cap.
// An entry is kStackWords words: the ref and (VLB_STACK_CULL) the entry distance of its box, one 128-byte row each.
constexpr int kStackWords = (VLB_STACK_CULL || VLB_BVH8) ? 2 : 1;   // 8-wide: an entry is a (node, remaining hit mask) group
constexpr uint32_t kEntryBytes = 128u * kStackWords;
struct WarpStack {
    uint32_t sm;  // shared address of s_stack[warp][0][0][lane]; entry e lives kEntryBytes * e further
    int* ovf;     // &scratch.ovf[0][lane]: overflow entries, kStackWords rows of 32 lanes each
    int sp;
    __device__ __forceinline__ void bind(const int* entry0, int* overflow) {
        const uint32_t a = (uint32_t)__cvta_generic_to_shared(entry0);
        asm volatile("mov.u32 %0, %1;" : "=r"(sm) : "r"(a));
        ovf = overflow;
    }
    __device__ __forceinline__ void clear() { sp = 0; }
    __device__ __forceinline__ bool empty() const { return sp == 0; }
    __device__ __forceinline__ bool room(int n) const { return sp + n <= kStackSize; }
    __device__ __forceinline__ void sts(uint32_t addr, int v, float tn) {
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
#if VLB_STACK_CULL
        asm volatile("st.shared.b32 [%0+128], %1;" ::"r"(addr), "f"(tn) : "memory");
#endif
    }
    __device__ __forceinline__ void push(int v, float tn) {
        if (sp < kSmemStack) sts(sm + kEntryBytes * (uint32_t)sp, v, tn);
        else {
            ovf[32 * kStackWords * (sp - kSmemStack)] = v;
#if VLB_STACK_CULL
            ovf[32 * kStackWords * (sp - kSmemStack) + 32] = __float_as_int(tn);
#endif
        }
        ++sp;
    }
    // next pending node / leaf whose box starts within tcull, or kNoChild
    __device__ __forceinline__ int pop(float tcull) {
        while (sp > 0) {
            --sp;
            int v;
            float tn = 0.f;
            if (sp < kSmemStack) {
                asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(sm + kEntryBytes * (uint32_t)sp) : "memory");
#if VLB_STACK_CULL
                asm volatile("ld.shared.b32 %0, [%1+128];" : "=f"(tn) : "r"(sm + kEntryBytes * (uint32_t)sp) : "memory");
#endif
            } else {
                v = ovf[32 * kStackWords * (sp - kSmemStack)];
#if VLB_STACK_CULL
                tn = __int_as_float(ovf[32 * kStackWords * (sp - kSmemStack) + 32]);
#endif
            }
            if (!VLB_STACK_CULL || tn <= tcull) return v;
        }
        return kNoChild;
    }
    // End of an ordered node step (see LocalStack::advance). While the three slots above the top are inside the shared
    // short stack the pushes are branch-free: three unconditional stores -- entry r_k goes to slot sp + max(n - k, 0)
    // with n = number of valid refs, so an invalid r3 / r2 is written first to the slot the next valid one overwrites
    // (or, for n = 0 / 1, to slots above the new top, which are dead). Without VLB_STACK_CULL the pop is branch-free too
    // (one predicated load): the node loop then has no taken branch besides its back edge (ncu round 2:
    // no_instruction + branch_resolving = 17 % of the stall samples).
    __device__ __forceinline__ int advance(int r0, int r1, int r2, int r3, float t1, float t2, float t3, float tcull, unsigned int* overflow) {
#if VLB_BAKE_FAST_PUSH
        if (sp + 3 <= kSmemStack) {
            const int v3 = r3 != kNoChild, v2 = r2 != kNoChild, v1 = r1 != kNoChild;
            const uint32_t base = sm + kEntryBytes * (uint32_t)sp;
            sts(base, r3, t3);
            sts(base + kEntryBytes * (uint32_t)v3, r2, t2);
            sts(base + kEntryBytes * (uint32_t)(v3 + v2), r1, t1);
            sp += v1 + v2 + v3;
#if VLB_STACK_CULL
            return r0 == kNoChild ? pop(tcull) : r0;                             // nothing hit => nothing was pushed
#else
            const bool need_pop = r0 == kNoChild, can_pop = need_pop && sp > 0;   // nothing hit => nothing was pushed
            int top = kNoChild;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\t@p ld.shared.b32 %0, [%1];\n\t}"
                         : "+r"(top) : "r"(base - 128u), "r"((int)can_pop) : "memory");
            sp -= can_pop ? 1 : 0;
            return need_pop ? top : r0;
#endif
        }
#endif
        if (r0 == kNoChild) return pop(tcull);
        if (r1 != kNoChild) {
            if (!room(3)) { if (overflow) *overflow = 1u; return r0; }
            if (r3 != kNoChild) push(r3, t3);
            if (r2 != kNoChild) push(r2, t2);
            push(r1, t1);
        }
        return r0;
    }
#if VLB_BVH8
    // 8-wide (see LocalStack::descend / pop_group): the group's two words sit in the entry's two rows.
    __device__ __forceinline__ void group_store(int e, int node, unsigned mask, bool both) {
        if (e < kSmemStack) {
            const uint32_t addr = sm + kEntryBytes * (uint32_t)e;
            if (both) asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(node) : "memory");
            asm volatile("st.shared.b32 [%0+128], %1;" ::"r"(addr), "r"(mask) : "memory");
        } else {
            if (both) ovf[64 * (e - kSmemStack)] = node;
            ovf[64 * (e - kSmemStack) + 32] = (int)mask;
        }
    }
    __device__ __forceinline__ int descend(const BvhView& b, int cur, unsigned mask, int flip) {
        if (mask == 0u) return pop_group(b, flip);
        const int p = __ffs((int)mask) - 1;
        mask &= mask - 1u;
        if (mask != 0u) {
            if (!room(1)) { if (b.overflow) *b.overflow = 1u; }
            else { group_store(sp, cur, mask, true); ++sp; }
        }
        return ref_of8(b, cur, p ^ flip);
    }
    __device__ __forceinline__ int pop_group(const BvhView& b, int flip) {
        if (sp == 0) return kNoChild;
        const int e = sp - 1;
        int node; unsigned mask;
        if (e < kSmemStack) {
            const uint32_t addr = sm + kEntryBytes * (uint32_t)e;
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(node) : "r"(addr) : "memory");
            asm volatile("ld.shared.b32 %0, [%1+128];" : "=r"(mask) : "r"(addr) : "memory");
        } else {
            node = ovf[64 * (e - kSmemStack)];
            mask = (unsigned)ovf[64 * (e - kSmemStack) + 32];
        }
        const int p = __ffs((int)mask) - 1;
        mask &= mask - 1u;
        if (mask != 0u) group_store(e, node, mask, false); else --sp;
        return ref_of8(b, node, p ^ flip);
    }
#endif
    // End of an unsorted node step (see LocalStack::advance_unsorted): a0..a3 in visiting order, kNoChild = missed.
    // Fast path, branch-free: a3, a2, a1 are stored at positions that advance only past valid refs (an invalid one is
    // overwritten by the next store or left above the new top); the ray continues with a0 if it was hit, else with
    // the new top of the stack.
    __device__ __forceinline__ int advance_unsorted(int a0, int a1, int a2, int a3, float tcull, unsigned int* overflow) {
        static_assert(!(VLB_NODE_ORDER1D && VLB_STACK_CULL), "unsorted node steps carry no entry distances");
        if (sp + 3 <= kSmemStack) {
            const int v3 = a3 != kNoChild, v2 = a2 != kNoChild, v1 = a1 != kNoChild;
            const uint32_t base = sm + kEntryBytes * (uint32_t)sp;
            sts(base, a3, 0.f);
            sts(base + kEntryBytes * (uint32_t)v3, a2, 0.f);
            sts(base + kEntryBytes * (uint32_t)(v3 + v2), a1, 0.f);
            sp += v1 + v2 + v3;
            const bool need_pop = a0 == kNoChild, can_pop = need_pop && sp > 0;
            int top = kNoChild;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\t@p ld.shared.b32 %0, [%1];\n\t}"
                         : "+r"(top) : "r"(sm + kEntryBytes * (uint32_t)(sp - 1)), "r"((int)can_pop) : "memory");
            sp -= can_pop ? 1 : 0;
            return need_pop ? top : a0;
        }
        const int a[4] = {a0, a1, a2, a3};
        int f = 0;
        while (f < 4 && a[f] == kNoChild) ++f;
        if (f == 4) return pop(tcull);
        if (!room(3)) { if (overflow) *overflow = 1u; return a[f]; }
        for (int k = 3; k > f; --k) if (a[k] != kNoChild) push(a[k], 0.0f);
        return a[f];
    }
};

#if VLB_BAKE_SMEM_STACK > 0
using RayStack = WarpStack;
#else
using RayStack = LocalStack;
#endif

#ifndef VLB_VIS_CORNER_MAJOR
#define VLB_VIS_CORNER_MAJOR 0
#endif
// Per-warp exchange area of a visibility-ray batch (gather passes): what the ray lanes need to know about the 32 hits
// being shaded -- hit position, biased ray origin (env_map.rchit:82 / main.rchit:143), grid cell -- and the result.
struct VisExchange {
    float P[3][32], so[3][32];
    int cell[3][32];
    unsigned occluded[32];       // bit c: the ray from hit `lane` to corner c of its cell was blocked
    int list[32];                // the lanes that hold a hit, in lane order (ray r belongs to hit list[r / 8])
    int root[32];                // where the 8 rays of hit `lane` enter the tree: the root, or their cell's common ancestor (k_cell_roots)
};

// Entry points of the visibility rays (gather passes). A visibility ray runs from a hit inside grid cell (i, j, k) to
// one of the cell's corners, so it stays inside the cell's box (grown by `margin` for the biased origin): the only
// triangles it can meet are those reaching into that box. One thread per cell walks down from the root for as long as
// exactly ONE child box overlaps the cell's box; the node it stops at has every such triangle below it, and the ray may
// start there -- on a scene much larger than a cell that skips the upper half of every descent. out[cell]: a node
// index, a leaf ref (< 0) or kNoChild when nothing reaches into the cell at all. fp32 4-wide nodes only.
#if !VLB_BAKE_GATHER_TU
__global__ void k_cell_roots(BvhView b, const float* __restrict__ px, const float* __restrict__ py, const float* __restrict__ pz,
                             int cx, int cy, int cz, float margin, int* __restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cx * cy * cz) return;
    const int i = c % cx, j = (c / cx) % cy, k = c / (cx * cy);
    const float lo[3] = {fminf(px[i], px[i + 1]) - margin, fminf(py[j], py[j + 1]) - margin, fminf(pz[k], pz[k + 1]) - margin};
    const float hi[3] = {fmaxf(px[i], px[i + 1]) + margin, fmaxf(py[j], py[j + 1]) + margin, fmaxf(pz[k], pz[k + 1]) + margin};
    int cur = b.n_tris ? 0 : kNoChild;
#if !VLB_NODE_Q8 && !VLB_BVH8
    while (cur >= 0) {
        const float4* q = b.nodes + (size_t)kNodeQuads * cur;
        const float4 lx = ld4(q), hx = ld4(q + 1), ly = ld4(q + 2), hy = ld4(q + 3), lz = ld4(q + 4), hz = ld4(q + 5), rf = ld4(q + 6);
        const float clx[4] = {lx.x, lx.y, lx.z, lx.w}, chx[4] = {hx.x, hx.y, hx.z, hx.w}, cly[4] = {ly.x, ly.y, ly.z, ly.w},
                    chy[4] = {hy.x, hy.y, hy.z, hy.w}, clz[4] = {lz.x, lz.y, lz.z, lz.w}, chz[4] = {hz.x, hz.y, hz.z, hz.w};
        const int ref[4] = {f2i(rf.x), f2i(rf.y), f2i(rf.z), f2i(rf.w)};
        int n_over = 0, only = kNoChild;
        for (int s = 0; s < 4; ++s) {      // empty slots have (+inf, -inf) boxes and never overlap
            if (clx[s] <= hi[0] && chx[s] >= lo[0] && cly[s] <= hi[1] && chy[s] >= lo[1] && clz[s] <= hi[2] && chz[s] >= lo[2]) { ++n_over; only = ref[s]; }
        }
        if (n_over == 0) { cur = kNoChild; break; }
        if (n_over > 1) break;
        cur = only;
    }
#endif
    out[c] = cur;
}
#endif

// The 8 visibility rays of each of the (up to 32) hits a warp shades together (shaders/main.rchit:143-163), traced as
// ONE batch by the whole warp: ray r = (hit r / 8, corner r % 8) goes to whichever lane is idle (the corner-major order,
// VLB_VIS_CORNER_MAJOR, measured 2 % slower), lanes step through the
// tree in the same while-while loop as the main rays and refill as they finish, so the warp stays full although the
// rays are short and of very different lengths. (Round 1 traced the 8 rays of a hit one after the other inside the
// shading lane: 32 lanes in lockstep on unrelated rays, a gather pass cost 5.5 direct passes.) `hm`: lanes holding a hit.
// Entry point of the 8 visibility rays of a hit with biased origin `ro` in grid cell (ci, cj, ck): below the root when the
// rays lie inside their cell's (grown) box -- origin inside, targets = the cell's corners -- else the root.
__device__ __forceinline__ int vis_entry(const GatherView& g, Vec3 ro, int ci, int cj, int ck) {
    if (!g.cell_root) return 0;
    const float m = g.cell_margin;
    const bool inside = g.Nx > 1 && g.Ny > 1 && g.Nz > 1 &&
        ro.x >= fminf(g.px[ci], g.px[ci + 1]) - m && ro.x <= fmaxf(g.px[ci], g.px[ci + 1]) + m &&
        ro.y >= fminf(g.py[cj], g.py[cj + 1]) - m && ro.y <= fmaxf(g.py[cj], g.py[cj + 1]) + m &&
        ro.z >= fminf(g.pz[ck], g.pz[ck + 1]) - m && ro.z <= fmaxf(g.pz[ck], g.pz[ck + 1]) + m;
    return inside ? __ldg(g.cell_root + ci + (g.Nx - 1) * (cj + (g.Ny - 1) * ck)) : 0;
}

template <bool COUNT>
__device__ __forceinline__ void trace_vis_batch(const BvhView& bvh, const GatherView& g, RayStack& stk, VisExchange& X, unsigned hm,
                                                int lane, int node_min, int refill_min, TraceCounters& cnt) {
    const unsigned full = 0xffffffffu, lt_mask = (1u << lane) - 1u;
    const int n_hits = __popc(hm), n_rays = 8 * n_hits;
    int next = 0, cur = kRayDone, tag = 0;
    bool busy = false;
    Vec3 ro = mk3(0.f, 0.f, 0.f), rd = ro, idir = ro, ood = ro;
    float tcull = 0.f;
    for (;;) {
        const unsigned idle = __ballot_sync(full, !busy);
        if ((__popc(idle) >= refill_min || idle == full) && next < n_rays) {      // batched refills, as in the main loop
            const int cand = next + __popc(idle & lt_mask);
            if (!busy && cand < n_rays) {
#if VLB_VIS_CORNER_MAJOR
                // corner-major order: neighbouring lanes trace the rays of neighbouring hits towards the SAME corner of their
                // cells (near-parallel rays from near-by origins) instead of the eight diverging rays of one hit
                const int c = cand / n_hits, h = X.list[cand - c * n_hits];
#else
                const int h = X.list[cand >> 3], c = cand & 7;     // (a table instead of __fns: ~50 instructions per ray)
#endif
                const Vec3 P = mk3(X.P[0][h], X.P[1][h], X.P[2][h]);
                int i, j, k; Vec3 d; float tmax;
                gather_corner(g, P, X.cell[0][h], X.cell[1][h], X.cell[2][h], c, i, j, k, d, tmax);
                if (tmax > 0.0f) {                                                     // main.rchit:154-155
                    ro = mk3(X.so[0][h], X.so[1][h], X.so[2][h]);
                    rd = mk3(f_div_c(d.x, tmax), f_div_c(d.y, tmax), f_div_c(d.z, tmax));
                    idir = mk3(safe_inv(rd.x), safe_inv(rd.y), safe_inv(rd.z));
                    ood = mk3(ro.x * idir.x, ro.y * idir.y, ro.z * idir.z);
                    tcull = tmax; tag = (h << 3) | c;
                    stk.clear();
                    cur = X.root[h];                 // the root, or below it (vis_entry: once per hit, not per ray)
                    busy = cur != kRayDone;          // kNoChild: nothing reaches into the cell, the corner is visible
                }
            }
            next = min(n_rays, next + __popc(idle));
        }
        const unsigned running = __ballot_sync(full, busy);
        if (running == 0u) {
            if (next >= n_rays) break;
            continue;
        }
        for (;;) {
            const bool at_node = busy && cur >= 0;
            const unsigned nm = __ballot_sync(full, at_node);
            if (nm == 0u || (__popc(nm) < node_min && __popc(running) - __popc(nm) >= node_min)) break;
            if (at_node) {
                if (COUNT) cnt.nodes++;
                cur = traverse_step<true>(bvh, cur, idir, ood, 0.0f, tcull, stk);
            }
        }
        if (busy && cur < 0 && cur != kRayDone) {
            HitRec occ; occ.id = -1; occ.t = tcull; occ.u = 0.f; occ.v = 0.f;
            if (leaf_step<true, COUNT>(bvh, cur, ro, rd, 0.0f, tcull, occ, &cnt)) {
                atomicOr(&X.occluded[tag >> 3], 1u << (tag & 7));
                cur = kRayDone;
            } else {
                cur = traverse_pop(bvh, idir, tcull, stk);
            }
        }
        if (busy && cur == kRayDone) busy = false;
    }
    __syncwarp();
}

#ifndef VLB_BAKE_MIN_BLOCKS
#define VLB_BAKE_MIN_BLOCKS 8      // resident 128-thread blocks per SM the register allocation is held to (8 -> 64 registers)
#endif
constexpr size_t kBakeSmemPerBlock = (kSmemStack > 0 ? (size_t)kStreamWarps * kSmemStack * kStackWords * 32 * sizeof(int) : 0) +
                                     (size_t)kStreamWarps * sizeof(HitQueue);
// gather passes add the second short stack and the exchange area of the visibility-ray batches
constexpr size_t kBakeSmemPerBlockGather = kBakeSmemPerBlock + (kSmemStack > 0 ? (size_t)kStreamWarps * kSmemStack * kStackWords * 32 * sizeof(int) : 0) +
                                           (size_t)kStreamWarps * sizeof(VisExchange);

template <int K, bool COUNT, bool GATHER, bool TEX>
__global__ void __launch_bounds__(kBakeBlock, GATHER ? 6 : VLB_BAKE_MIN_BLOCKS) k_bake_stream(const BakeParams p) {
    constexpr int V = (K * 3 <= 32) ? 32 : 64;
    __shared__ int s_stack[kSmemStack > 0 ? kStreamWarps : 1][kSmemStack > 0 ? kSmemStack : 1][kStackWords][32];
    __shared__ HitQueue s_hq[kStreamWarps];
    __shared__ int s_stack2[GATHER && kSmemStack > 0 ? kStreamWarps : 1][GATHER && kSmemStack > 0 ? kSmemStack : 1][kStackWords][32];
    __shared__ VisExchange s_vis[GATHER ? kStreamWarps : 1];
    using WarpQueues = WarpQueuesT<(VLB_SQ_PACKED != 0) && !GATHER>;
    // (the host sizes the slab for the larger form; each kernel strides it by its own)
    WarpQueues* s_all = static_cast<WarpQueues*>(p.stream_scratch) + (size_t)blockIdx.x * kStreamWarps;
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    WarpQueues& S = s_all[warp];
    WarpSpill& SP = p.stream_spill[(size_t)blockIdx.x * kStreamWarps + warp];
    HitQueue& HQ = s_hq[warp];
    const BvhView& bvh = p.bvh;
    const bool want_shadow = (p.c.flags & 1u) != 0;
    TraceCounters cnt; cnt.nodes = 0; cnt.tris = 0;
    uint32_t shadow = 0;
    // COUNT builds: warp-level utilisation of the three phases (iterations, active lanes summed)
    uint32_t u_node_it = 0, u_node_ln = 0, u_leaf_it = 0, u_leaf_ln = 0, u_shade_it = 0, u_shade_ln = 0, u_outer = 0, u_ovf = 0;
#if VLB_BAKE_SMEM_STACK > 0
    WarpStack stk;
    stk.bind(&s_stack[warp][0][0][lane], &SP.ovf[0][lane]);
#else
    LocalStack stk;
#endif
    stk.clear();
    RayStack stk2;               // visibility-ray batches of the gather passes (the main rays keep `stk` while a batch runs)
#if VLB_BAKE_SMEM_STACK > 0
    if (GATHER) stk2.bind(&s_stack2[warp][0][0][lane], p.vis_ovf + ((size_t)blockIdx.x * kStreamWarps + warp) * kOvfStack * kStackWords * 32 + lane);
#endif
    stk2.clear();

    for (;;) {
        unsigned item = 0;
        if (lane == 0) item = atomicAdd(p.work_counter, 1u);
        item = __shfl_sync(full, item, 0);
        if (item >= p.n_items) break;
        // items [0, n_whole): one whole probe each, written straight to the output; the rest: one chunk run of a
        // probe each, written as a partial (bake_device explains why the two give bit-identical sums)
        uint32_t q, part;
        const bool whole = item < p.n_whole;
        if (whole) { q = item; part = 0; }
        else { const uint32_t j = item - p.n_whole; q = p.n_whole + j / (uint32_t)p.chunks; part = j % (uint32_t)p.chunks; }
        const uint32_t nxy = (uint32_t)(p.Nx * p.Ny), in_slice = q % nxy;
        const Vec3 po = mk3(p.px[in_slice % p.Nx], p.py[in_slice / p.Nx], p.pz[p.k0 + (q / nxy) * p.kstride]);
        const int tile_begin = whole ? 0 : part * p.tiles_per_chunk;
        const int tile_end = whole ? p.n_tiles : min(tile_begin + p.tiles_per_chunk, p.n_tiles);
        float coef0 = 0.f, coef1 = 0.f;   // running sums of coefficient `lane` (V=32) / 2*lane, 2*lane+1 (V=64)

        for (int base_tile = tile_begin; base_tile < tile_end; base_tile += kChunkTiles) {
            const int n_dirs = min(kChunkTiles, tile_end - base_tile) * 32;
            // ------------------------------ trace the chunk ------------------------------
            int next = 0, n_hit = 0, n_sh = 0;   // warp-uniform: directions handed out, queue fills
            bool busy = false;
            int kind = 0;                        // 0 primary (closest hit), 1 shadow (any hit)
            int my_dir = 0, cur = kRayDone;
            Vec3 ro = po, rd = po, idir = po, ood = po;
            float tmin = 0.f, tcull = 0.f;
            // primary ray: the closest hit so far. Shadow ray: id = -1 until occluded, (t, u, v) = the radiance of
            // the direction if the light turns out to be visible (an any-hit ray has no use for them).
            HitRec best; best.id = -1; best.t = 0.f; best.u = 0.f; best.v = 0.f;
            for (;;) {
                if (COUNT) ++u_outer;
                // ---- 1. refill idle lanes: queued shadow rays first, then new directions ----
                const unsigned idle = __ballot_sync(full, !busy);
                // p.refill_order 3 (needs a hit queue as large as a chunk): the hits of the chunk's directions are only queued while
                // directions are left (phase A: closest-hit rays only); then (phase B) they are shaded 32 at a time whenever the
                // shadow queue cannot fill every idle lane, so the any-hit batches are homogeneous and full as well
                const bool deferred = p.refill_order == 3, phase_b = deferred && next >= n_dirs;
                const bool shade_first = phase_b && n_hit > 0 && n_sh < __popc(idle);
                if ((__popc(idle) >= p.refill_min || idle == full) && (n_sh > 0 || next < n_dirs) && !shade_first) {
                    const int n_idle = __popc(idle), rank = __popc(idle & lt_mask);
                    // p.refill_order 0: queued shadow rays first, new directions for the lanes that are left. 1: homogeneous batches --
                    // shadow rays only when they fill every idle lane (or must be drained before the next shading), else new
                    // directions first and shadow rays for the rest.
                    // 2 (needs a shadow queue as large as a chunk): all directions of the chunk first, then its shadow rays
                    const bool sh_first = p.refill_order == 0 || (p.refill_order == 1 && n_sh >= n_idle) || n_sh > kShadowCap - 32 || next >= n_dirs;
                    const int take_sh = sh_first ? min(n_idle, n_sh) : min(n_sh, max(0, n_idle - (n_dirs - next)));
                    const int take_new = min(n_idle - take_sh, n_dirs - next);
                    bool fresh = false;
                    bool tab_idir = false;      // VLB_IDIR_TAB: idir came from the direction table
                    if (!busy && rank < take_sh) {
                        const int e = n_sh - 1 - rank;
                        if constexpr (WarpQueues::kPacked) {
                            const float4 qa = ld_scratch(&S.sq_a[e]), qb = ld_scratch(&S.sq_b[e]), qc = ld_scratch(&S.sq_c[e]);
                            ro = mk3(qa.x, qa.y, qa.z); rd = mk3(qb.x, qb.y, qb.z);
                            tmin = 0.0f; tcull = qa.w;                                      // env_map.rchit:87
                            my_dir = __float_as_int(qb.w);
                            best.id = -1; best.t = qc.x; best.u = qc.y; best.v = qc.z;
                        } else {
                            ro = mk3(ld_scratch(&S.sq_o[0][e]), ld_scratch(&S.sq_o[1][e]), ld_scratch(&S.sq_o[2][e]));
                            rd = mk3(ld_scratch(&S.sq_d[0][e]), ld_scratch(&S.sq_d[1][e]), ld_scratch(&S.sq_d[2][e]));
                            tmin = 0.0f; tcull = ld_scratch(&S.sq_len[e]);                  // env_map.rchit:87
                            my_dir = ld_scratch(&S.sq_dir[e]);
                            best.id = -1; best.t = ld_scratch(&S.sq_rgb[0][e]); best.u = ld_scratch(&S.sq_rgb[1][e]); best.v = ld_scratch(&S.sq_rgb[2][e]);
                        }
                        kind = 1; busy = true; fresh = true;
                    } else if (!busy && rank - take_sh < take_new) {
                        const int cand = next + rank - take_sh;
                        const int tile = base_tile + (cand >> 5), w = cand & 31;
                        bool inside;
                        if (VLB_DIR_TABLES && p.dir_tab) {
                            const float4 dt = __ldg(p.dir_tab + ((size_t)tile * 32 + w) * kDirTabQuads);
                            rd = mk3(dt.x, dt.y, dt.z);
                            inside = dt.w != 0.0f;
#if VLB_IDIR_TAB
                            if (inside) {
                                const float4 it = __ldg(p.dir_tab + ((size_t)tile * 32 + w) * kDirTabQuads + 1);
                                idir = mk3(it.x, it.y, it.z);
                                tab_idir = true;
                            }
#endif
                        } else {
                            int x, y;
                            tile_xy(p, tile, w, x, y);
                            inside = x < p.W && y < p.H;
                            if (inside) {
                                const float2 row = __ldg(p.row_sc + y), col = __ldg(p.col_cs + x);
                                const Vec3 t = to_vector_sc(row.x, row.y, col.x, col.y);   // sh_common.h:8-12
                                rd = mk3(t.x, t.z, t.y);                                   // env_map.rgen:21 .xzy
                            }
                        }
                        if (inside) {
                            ro = po;
                            tmin = p.c.tmin; tcull = p.c.tmax;
                            best.id = -1; best.t = tcull; best.u = 0.f; best.v = 0.f;
                            my_dir = cand; kind = 0; busy = true; fresh = true;
                        } else {
                            S.slot[cand] = make_float4(0.f, 0.f, 0.f, 0.f);                    // outside the direction grid
                        }
                    }
                    if (fresh) {
                        if (!(VLB_IDIR_TAB && tab_idir)) idir = mk3(safe_inv(rd.x), safe_inv(rd.y), safe_inv(rd.z));
                        ood = mk3(ro.x * idir.x, ro.y * idir.y, ro.z * idir.z);
                        stk.clear();
                        cur = bvh.n_tris ? 0 : kRayDone;
                    }
                    n_sh -= take_sh;
                    next += take_new;
                    __syncwarp();
                }
                const unsigned running = __ballot_sync(full, busy);
                // ---- 2. shade queued hits: a full warp of them, or whatever is left when nothing runs ----
                const bool shade_now = deferred ? ((shade_first && (__popc(idle) >= p.refill_min || idle == full) && n_sh <= kShadowCap - 32) ||
                                                   (n_hit > kHitCap - 32 && n_sh <= kShadowCap - 32))
                                                : (n_hit >= 32 && n_sh <= kShadowCap - 32);
                if (shade_now || (running == 0u && n_hit > 0)) {
                    const int take = min(n_hit, 32);
                    const int e = n_hit - take + lane;
                    if (COUNT) { ++u_shade_it; u_shade_ln += take; }
                    bool push = false;
                    float lit_rgb[3] = {0.f, 0.f, 0.f};
                    ShadePrelude pre;
                    int dir = 0;
                    unsigned occluded = 0;
                    if (GATHER) {
                        // gather passes: first the visibility rays of all hits of this batch (trace_vis_batch); the hit
                        // shading below then only needs the 8-bit result per hit
                        VisExchange& X = s_vis[warp];
                        bool has_hit = false;
                        const int hd = lane < take ? (int)HQ.dir[e] : 0;
                        const float4 hrec = lane < take ? ld_scratch(&S.slot[hd]) : make_float4(__int_as_float(-1), 0.f, 0.f, 0.f);
                        if (lane < take && __float_as_int(hrec.x) >= 0) {
                            HitRec h; h.id = __float_as_int(hrec.x); h.t = hrec.y; h.u = hrec.z; h.v = hrec.w;
                            const int tile = base_tile + (hd >> 5), w = hd & 31;
                            ShadePrelude q;
                            shade_prelude<TEX>(p.shade, p.c, h, po, slot_direction(p, tile, w), q);
                            X.P[0][lane] = q.P.x; X.P[1][lane] = q.P.y; X.P[2][lane] = q.P.z;
                            X.so[0][lane] = q.so.x; X.so[1][lane] = q.so.y; X.so[2][lane] = q.so.z;
                            X.cell[0][lane] = gather_cell(q.P.x, p.g.origin[0], p.g.step[0], p.g.Nx);
                            X.cell[1][lane] = gather_cell(q.P.y, p.g.origin[1], p.g.step[1], p.g.Ny);
                            X.cell[2][lane] = gather_cell(q.P.z, p.g.origin[2], p.g.step[2], p.g.Nz);
                            X.root[lane] = vis_entry(p.g, q.so, X.cell[0][lane], X.cell[1][lane], X.cell[2][lane]);
                            has_hit = true;
                        }
                        X.occluded[lane] = 0u;
                        const unsigned hm = __ballot_sync(full, has_hit);
                        if (has_hit) X.list[__popc(hm & lt_mask)] = lane;
                        __syncwarp();
                        if (hm != 0u && bvh.n_tris) trace_vis_batch<COUNT>(bvh, p.g, stk2, X, hm, lane, p.node_min, p.vis_refill_min, cnt);
                        occluded = X.occluded[lane];
                        __syncwarp();
                    }
                    if (lane < take) {
                        dir = HQ.dir[e];
                        const float4 hrec = ld_scratch(&S.slot[dir]);
                        HitRec h; h.id = __float_as_int(hrec.x); h.t = hrec.y; h.u = hrec.z; h.v = hrec.w;
                        const int tile = base_tile + (dir >> 5), w = dir & 31;
                        const Vec3 r = slot_direction(p, tile, w);
                        float rgb[3] = {0.f, 0.f, 0.f};                                 // env_map.rgen:25
                        if (h.id >= 0) {
                            const bool lit = shade_prelude<TEX>(p.shade, p.c, h, po, r, pre);
                            float ind[3] = {0.f, 0.f, 0.f};
                            if (GATHER) gather_accumulate<K>(p.g, pre, occluded, ind);         // main.rchit:156-165
                            // With a shadow ray to come, radiance for both outcomes now and the ray decides (env_map.rchit:83-99):
                            // the slot gets the occluded one, the ray carries the lit one. (Two inlined copies of
                            // shade_finish, not three: this is cold code that pays for every instruction-cache line.)
                            push = lit && want_shadow;
                            shade_finish(p.c, pre, r, !lit || want_shadow, ind, rgb);
                            if (push) shade_finish(p.c, pre, r, false, ind, lit_rgb);
                        } else if ((p.c.flags & 2u) && p.shade.sky) {                   // VLB_BAKE_SKYBOX_ON_MISS
                            sky_lookup(p.shade, r, rgb);
                            if (p.c.flags & 4u) { rgb[0] = srgb_encode(rgb[0]); rgb[1] = srgb_encode(rgb[1]); rgb[2] = srgb_encode(rgb[2]); }
                        }
                        if (p.c.flags & 8u) { rgb[0] = quant8(rgb[0]); rgb[1] = quant8(rgb[1]); rgb[2] = quant8(rgb[2]); }
                        S.slot[dir] = make_float4(rgb[0], rgb[1], rgb[2], 0.f);
                    }
                    const unsigned pm = __ballot_sync(full, push);
                    if (push) {
                        const int d = n_sh + __popc(pm & lt_mask);
                        if constexpr (WarpQueues::kPacked) {
                            S.sq_a[d] = make_float4(pre.so.x, pre.so.y, pre.so.z, pre.llen);             // env_map.rchit:82
                            S.sq_b[d] = make_float4(pre.Ln.x, pre.Ln.y, pre.Ln.z, __int_as_float(dir));
                            S.sq_c[d] = make_float4(lit_rgb[0], lit_rgb[1], lit_rgb[2], 0.f);
                        } else {
                            S.sq_o[0][d] = pre.so.x; S.sq_o[1][d] = pre.so.y; S.sq_o[2][d] = pre.so.z;   // env_map.rchit:82
                            S.sq_d[0][d] = pre.Ln.x; S.sq_d[1][d] = pre.Ln.y; S.sq_d[2][d] = pre.Ln.z;
                            S.sq_len[d] = pre.llen; S.sq_dir[d] = dir;
#pragma unroll
                            for (int k = 0; k < 3; ++k) S.sq_rgb[k][d] = lit_rgb[k];
                        }
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
                    if (nm == 0u || (__popc(nm) < p.node_min && __popc(running) - __popc(nm) >= p.leaf_min)) break;
                    if (COUNT) { ++u_node_it; u_node_ln += __popc(nm); }
                    if (at_node) {
                        if (COUNT) { cnt.nodes++; if (stk.sp > kSmemStack) ++u_ovf; }
                        cur = traverse_step<true>(bvh, cur, idir, ood, tmin, tcull, stk);   // one code path for both ray kinds
                    }
                }
                // ---- 4. leaves ----
                const bool at_leaf = busy && cur < 0 && cur != kRayDone;
                if (COUNT) { const unsigned lm = __ballot_sync(full, at_leaf); if (lm) { ++u_leaf_it; u_leaf_ln += __popc(lm); } }
                if (at_leaf) {
                    bool terminated;
                    if (kind) {
                        HitRec occ; occ.id = -1; occ.t = tcull; occ.u = 0.f; occ.v = 0.f;
                        terminated = leaf_step<true, COUNT>(bvh, cur, ro, rd, tmin, tcull, occ, &cnt);
                        if (terminated) best.id = 0;                                    // occluded (shadow.rmiss did not run)
                    } else {
                        terminated = leaf_step<false, COUNT>(bvh, cur, ro, rd, tmin, tcull, best, &cnt);
                    }
                    cur = terminated ? kRayDone : traverse_pop(bvh, idir, tcull, stk);
                }
                // ---- 5. finished rays free their lane ----
                const bool fin = busy && cur == kRayDone;
                const unsigned fp = __ballot_sync(full, fin && kind == 0);
                if (fin) {
                    if (kind == 0) {
                        const int e = n_hit + __popc(fp & lt_mask);
                        S.slot[my_dir] = make_float4(__int_as_float(best.id), best.t, best.u, best.v);
                        HQ.dir[e] = (unsigned short)my_dir;
                    } else if (best.id < 0) {                                           // light visible: the lit radiance replaces the occluded one
                        float rgb[3] = {best.t, best.u, best.v};
                        if (p.c.flags & 8u) { rgb[0] = quant8(rgb[0]); rgb[1] = quant8(rgb[1]); rgb[2] = quant8(rgb[2]); }  // QUANTIZE_RGBA8
                        S.slot[my_dir] = make_float4(rgb[0], rgb[1], rgb[2], 0.f);
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
                if (VLB_DIR_TABLES && p.proj_tab) {
                    if (__ldg(p.dir_tab + ((size_t)tile * 32 + lane) * kDirTabQuads).w != 0.0f) {
                        const float* bwp = p.proj_tab + (size_t)tile * (K * 32) + lane;
                        const float4 rad = ld_scratch(&S.slot[tt * 32 + lane]);
                        const float r0 = rad.x, r1 = rad.y, r2 = rad.z;
#pragma unroll
                        for (int i = 0; i < K; ++i) {
                            const float bw = __ldg(bwp + 32 * i);
                            acc[3 * i + 0] = fmaf(bw, r0, acc[3 * i + 0]);
                            acc[3 * i + 1] = fmaf(bw, r1, acc[3 * i + 1]);
                            acc[3 * i + 2] = fmaf(bw, r2, acc[3 * i + 2]);
                        }
                    }
                    continue;
                }
                int x, y;
                tile_xy(p, tile, lane, x, y);
                if (x < p.W && y < p.H) {
                    const float2 row = __ldg(p.row_sc + y), col = __ldg(p.col_cs + x);
                    const Vec3 t = to_vector_sc(row.x, row.y, col.x, col.y);
                    const float w = p.pixel_area * row.x;                               // sh.comp:32-33
                    float b[K];
                    sh_basis<K>(p.world_frame ? mk3(t.x, t.z, t.y) : t, b);             // sh.comp:30,39
                    const float4 rad = ld_scratch(&S.slot[tt * 32 + lane]);
                    const float r0 = rad.x, r1 = rad.y, r2 = rad.z;
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
#if VLB_BAKE_DISCARD
            // The tile's radiances are dead now (the next chunk writes every entry before it reads any), but their
            // lines sit dirty in L2 and would be written back to HBM when the BVH / skybox traffic evicts them:
            // round 2 measured 4.2 GB of such write-backs per C3 launch. Tell L2 to drop them instead.
            static_assert(sizeof(S.slot) % 128 == 0, "the direction slots must cover whole 128-byte lines");
            for (uint32_t off = 128u * lane; off < sizeof(S.slot); off += 32u * 128u)
                asm volatile("discard.global.L2 [%0], 128;" ::"l"(reinterpret_cast<char*>(&S.slot[0]) + off) : "memory");
            __syncwarp();
#endif
        }
        float* dst = whole ? p.out + out_slot(p, q) * VLB_SH_STRIDE : p.partials + (size_t)(item - p.n_whole) * VLB_SH_STRIDE;
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
    if (COUNT) {
        unsigned long long ov = u_ovf;
        for (int off = 16; off > 0; off >>= 1) ov += __shfl_xor_sync(full, ov, off);
        if (lane == 0) {
            atomicAdd(p.stats + 4, (unsigned long long)u_node_it); atomicAdd(p.stats + 5, (unsigned long long)u_node_ln);
            atomicAdd(p.stats + 6, (unsigned long long)u_leaf_it); atomicAdd(p.stats + 7, (unsigned long long)u_leaf_ln);
            atomicAdd(p.stats + 8, (unsigned long long)u_shade_it); atomicAdd(p.stats + 9, (unsigned long long)u_shade_ln);
            atomicAdd(p.stats + 10, (unsigned long long)u_outer); atomicAdd(p.stats + 11, ov);
        }
    }
    if (lane == 0) {
        atomicAdd(p.stats + 0, s);
        if (COUNT) { atomicAdd(p.stats + 1, nn); atomicAdd(p.stats + 2, nt); }
    }
}

using BakeKernel = void (*)(const BakeParams);
#if VLB_BAKE_GATHER_TU
// bake_gather.cu: the gather-pass instantiation for (sh coefficients K, instrumented build, textured scene)
BakeKernel bake_gather_kernel(int K, bool count, bool tex) {
#define VLB_PICK(KK) (tex ? (count ? k_bake_stream<KK, true, true, true> : k_bake_stream<KK, false, true, true>)   \
                          : (count ? k_bake_stream<KK, true, true, false> : k_bake_stream<KK, false, true, false>))
    return K == 9 ? VLB_PICK(9) : VLB_PICK(16);
#undef VLB_PICK
}
#else
BakeKernel bake_gather_kernel(int K, bool count, bool tex);   // bake_gather.cu

// Probes baked as chunk-run items: per-probe sum of the partials in chunk order (fixed order => deterministic).
__global__ void k_sum_partials(const BakeParams p, uint32_t n_probes) {
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (n_probes - p.n_whole) * VLB_SH_STRIDE) return;
    const uint32_t j = idx / VLB_SH_STRIDE, c = idx % VLB_SH_STRIDE;
    float s = 0.f;
    for (int k = 0; k < p.chunks; ++k) s += p.partials[((size_t)j * p.chunks + k) * VLB_SH_STRIDE + c];
    p.out[out_slot(p, p.n_whole + j) * VLB_SH_STRIDE + c] = s;
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
    VLB_CUDA(ctx, ctx->d_stats.reserve(16 * sizeof(unsigned long long)));
    VLB_CUDA(ctx, cudaMemcpyAsync(ctx->d_axis.p, axis.data(), axis.size() * sizeof(float), cudaMemcpyHostToDevice, st));
    VLB_CUDA(ctx, cudaMemsetAsync(ctx->d_work_counter.p, 0, 16, st));
    VLB_CUDA(ctx, cudaMemsetAsync(ctx->d_stats.p, 0, 16 * sizeof(unsigned long long), st));

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
    p.tiles_x_rcp = p.tiles_x > 1 ? ~0ull / (unsigned long long)p.tiles_x + ((~0ull % (unsigned long long)p.tiles_x) + 1 == (unsigned long long)p.tiles_x ? 2 : 1) : 0;
    p.tiles_per_chunk = 0;       // the work decomposition follows the kernel choice (it needs the resident warp count)
    p.pixel_area = (2.0f * kPi / (float)W) * (kPi / (float)H);            // sh.comp:32
    p.work_counter = ctx->d_work_counter.as<unsigned int>();
    p.stats = ctx->d_stats.as<unsigned long long>();
    p.ref_order = (s->flags & VLB_BAKE_REFERENCE_PROBE_ORDER) ? 1 : 0;
    p.world_frame = (s->flags & VLB_BAKE_SH_WORLD_FRAME) ? 1 : 0;
    p.node_min = std::max(1, std::min(32, env_flag("VLB_BAKE_NODE_MIN", 6)));
    p.leaf_min = std::max(1, std::min(32, env_flag("VLB_BAKE_LEAF_MIN", p.node_min)));
    p.refill_min = std::max(1, std::min(32, env_flag("VLB_BAKE_REFILL_MIN", 20)));
    p.vis_refill_min = std::max(1, std::min(32, env_flag("VLB_BAKE_VIS_REFILL_MIN", 20)));
    p.g.prev = d_prev_full; p.g.px = p.px; p.g.py = p.py; p.g.pz = p.pz; p.g.Nx = Nx; p.g.Ny = Ny; p.g.Nz = Nz;
    for (int k = 0; k < 3; ++k) { p.g.origin[k] = s->origin[k]; p.g.step[k] = s->step[k]; }
    p.g.gain = s->indirect_gain; p.g.world_frame = p.world_frame;
    const bool gather = d_prev_full != nullptr;
    p.out = d_out;

    const bool count = env_flag("VLB_BAKE_COUNTERS", 0) != 0;
    const int K = s->sh_order == 2 ? 9 : 16;
    // per-direction tables (ray direction, projection weights): for direction grids up to 64 Ki slots -- beyond that the
    // tables would be streamed from HBM once per probe, and the kernel computes the values in place as before
    if (d_prev_full == nullptr && (uint64_t)p.n_tiles * 32 <= (1u << 16) && env_flag("VLB_BAKE_DIR_TABLES", 1)) {   // direct passes only
        const int key[5] = {W, H, p.tile_lw, K, p.world_frame};
        const size_t n_slots = (size_t)p.n_tiles * 32;
        if (std::memcmp(key, ctx->dir_tab_key, sizeof key) != 0 || ctx->dir_tab_stream != st) {
            VLB_CUDA(ctx, ctx->d_dir_tab.reserve(n_slots * kDirTabQuads * sizeof(float4)));
            VLB_CUDA(ctx, ctx->d_proj_tab.reserve(n_slots * 16 * sizeof(float)));
            const unsigned blocks = (unsigned)((n_slots + 127) / 128);
            if (K == 9) k_dir_tables<9><<<blocks, 128, 0, st>>>(p, ctx->d_dir_tab.as<float4>(), ctx->d_proj_tab.as<float>());
            else        k_dir_tables<16><<<blocks, 128, 0, st>>>(p, ctx->d_dir_tab.as<float4>(), ctx->d_proj_tab.as<float>());
            VLB_LAUNCH_CHECK(ctx);
            std::memcpy(ctx->dir_tab_key, key, sizeof key);
            ctx->dir_tab_stream = st;
        }
        p.dir_tab = ctx->d_dir_tab.as<float4>();
        p.proj_tab = ctx->d_proj_tab.as<float>();
    }
    // TEX: only scenes with a textured material pay for the texture branch of the hit shading
    const bool tex = ctx->max_tex_index >= 0;
#define VLB_PICK(KK) (tex ? (count ? k_bake_stream<KK, true, false, true> : k_bake_stream<KK, false, false, true>)   \
                          : (count ? k_bake_stream<KK, true, false, false> : k_bake_stream<KK, false, false, false>))
    BakeKernel kern = gather ? bake_gather_kernel(K, count, tex) : (K == 9 ? VLB_PICK(9) : VLB_PICK(16));
#undef VLB_PICK
    // Shared memory holds the short stacks and the hit queues (kBakeSmemPerBlock per block); everything else of the
    // SM's 228 KB stays L1 (BVH nodes, triangles, shadow-ray queues, direction slots).
    const int min_blocks = gather ? 6 : VLB_BAKE_MIN_BLOCKS;
    int carve = (int)((min_blocks * ((gather ? kBakeSmemPerBlockGather : kBakeSmemPerBlock) + 1024) * 100 + 228 * 1024 - 1) / (228 * 1024));
    carve = env_flag("VLB_BAKE_CARVEOUT", std::min(100, carve));
    if (carve >= 0) VLB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
    int per_sm = 0;
    VLB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kBakeBlock, 0));
    per_sm = std::max(per_sm, 1);
    const uint32_t full_grid = (uint32_t)(ctx->sm_count * per_sm);

    // Work decomposition. A probe's directions are traced in chunks of kChunkTiles tiles = 512 directions whose SH sums are
    // added up in chunk order, so a probe can be ONE work item (a warp walks all its chunks and writes the 192-byte
    // result itself) or one item PER CHUNK (each writes a partial, k_sum_partials adds them in chunk order): both
    // give bit-identical coefficients, because 0 + a_0 + a_1 + ... is evaluated left to right either way. Whole-probe
    // items need no partials (round 1 wrote and re-read 403 MB of them at C3) but are long (16 chunks at 64x64
    // directions); so the first probes of the call are whole-probe items and the last ones, about `tail_waves`
    // rounds of all resident warps, are per-chunk items that fill the gaps while the long items retire. Which probes
    // fall into the tail depends on the call; the coefficients do not, so a probe's result is bit-identical however
    // the grid is sharded. Direction grids of less than one chunk per item (small grids: the chunk is halved down to
    // one tile until the whole grid has >= 2^18 items) are all partial items, as in round 1; that choice is a function
    // of the WHOLE grid and the direction grid only.
    const uint64_t total_probes = (uint64_t)Nx * Ny * Nz;
    int tiles_per_item = kChunkTiles;
    while (tiles_per_item > 1 && total_probes * (uint64_t)((p.n_tiles + tiles_per_item - 1) / tiles_per_item) < (1ull << 18))
        tiles_per_item >>= 1;
    int chunks = (p.n_tiles + tiles_per_item - 1) / tiles_per_item;
    chunks = std::max(1, std::min(chunks, p.n_tiles));
    p.tiles_per_chunk = (p.n_tiles + chunks - 1) / chunks;
    p.chunks = (p.n_tiles + p.tiles_per_chunk - 1) / p.tiles_per_chunk;
    uint64_t n_whole = 0;
    if (p.chunks == 1) {
        n_whole = n_probes;
    } else if (p.tiles_per_chunk == kChunkTiles) {
        const uint64_t tail_waves = (uint64_t)std::max(0, env_flag("VLB_BAKE_TAIL_WAVES", 8));
        const uint64_t warps = (uint64_t)full_grid * kStreamWarps;
        // Whole-probe items only pay when there are several probes per resident warp; a call with few probes and many
        // directions (the reference's own 7x7x7 x 3141x1000 bake: 343 probes of 6,141 chunks for 4,736 warps) is all
        // per-chunk items, or most of the GPU would idle behind 343 long items.
        const uint64_t tail_probes = (env_flag("VLB_BAKE_TAIL_WAVES", 8) < 0 || n_probes < 2 * warps) ? n_probes
                                     : std::min<uint64_t>(n_probes, (tail_waves * warps + p.chunks - 1) / p.chunks);
        n_whole = n_probes - tail_probes;
    }
    // Ray-slot policy: chunks of at least half the full size are traced in two phases (3: all closest-hit rays first, their
    // hits queued; then shading and any-hit batches) -- the drain between the phases only pays when a chunk is long
    // (measured on the atrium: 16 tiles -3.2 %, 8 tiles -2.8 %, 4 tiles -0.6 %, 2 tiles +8 %, 1 tile +32 %); short chunks
    // (small grids split into many items, C2: one tile per item) keep the interleaved order 1. Never changes a result.
    p.refill_order = env_flag("VLB_BAKE_REFILL_ORDER", 2 * p.tiles_per_chunk >= kChunkTiles ? 3 : 1);
    const uint64_t n_items = n_whole + (n_probes - n_whole) * (uint64_t)p.chunks;
    if (n_items >= (1ull << 32)) return ctx->fail(VLB_ERR_UNSUPPORTED, "bake: too many work items");
    p.n_items = (uint32_t)n_items; p.n_whole = (uint32_t)n_whole;
    if (n_whole < n_probes) {
        VLB_CUDA(ctx, ctx->d_partials.reserve((size_t)(n_items - n_whole) * VLB_SH_STRIDE * sizeof(float)));
        p.partials = ctx->d_partials.as<float>();
    }
    const uint32_t grid = std::max(1u, std::min(full_grid, (p.n_items + kStreamWarps - 1) / kStreamWarps));

    VLB_CUDA(ctx, ctx->d_stream_scratch.reserve((size_t)grid * kStreamWarps * kWarpQueuesBytes));
    p.stream_scratch = ctx->d_stream_scratch.p;
    VLB_CUDA(ctx, ctx->d_stream_spill.reserve((size_t)grid * kStreamWarps * sizeof(WarpSpill)));
    p.stream_spill = ctx->d_stream_spill.as<WarpSpill>();
    if (gather && Nx > 1 && Ny > 1 && Nz > 1 && env_flag("VLB_BAKE_CELL_ROOTS", 1)) {
        const int cx = Nx - 1, cy = Ny - 1, cz = Nz - 1;
        VLB_CUDA(ctx, ctx->d_cell_root.reserve((size_t)cx * cy * cz * sizeof(int)));
        // the biased ray origin lies within shadow_bias of the hit; a little more for rounding
        const float ext = std::max({fabsf(s->step[0]) * Nx, fabsf(s->step[1]) * Ny, fabsf(s->step[2]) * Nz});
        p.g.cell_margin = fabsf(s->shadow_bias) * 1.5f + 1e-5f * ext;
        k_cell_roots<<<(unsigned)((cx * cy * cz + 127) / 128), 128, 0, st>>>(p.bvh, p.px, p.py, p.pz, cx, cy, cz, p.g.cell_margin, ctx->d_cell_root.as<int>());
        VLB_LAUNCH_CHECK(ctx);
        p.g.cell_root = ctx->d_cell_root.as<int>();
    }
    if (gather) {
        VLB_CUDA(ctx, ctx->d_vis_ovf.reserve((size_t)grid * kStreamWarps * kOvfStack * kStackWords * 32 * sizeof(int)));
        p.vis_ovf = ctx->d_vis_ovf.as<int>();
    }
    VLB_CUDA(ctx, cudaEventRecord(ctx->ev[0], st));
    VLB_CUDA(ctx, cudaEventRecord(ctx->ev[2], st));
    {
        // The warps' direction slots and shadow-ray queues (the hot scratch: 10.75 KB per warp, 51 MB for 4,736 warps) are
        // written and re-read all through the launch, while scene + skybox + scratch together just overflow the L2: ncu
        // showed GBs of scratch lines written back to HBM and fetched again. The launch can carry an L2
        // access-policy window over the hot scratch (persisting lines, as far as the device's set-aside reaches), so the
        // capacity misses fall on the read-only scene instead. Measured (profiles/r02_bake_l2_persist_ab.log): no change
        // in time, MORE write-backs (the set-aside evicts dirty lines earlier) -> off; VLB_BAKE_L2_PERSIST=<MB> turns it on.
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kBakeBlock); cfg.dynamicSmemBytes = 0; cfg.stream = st;
        cudaLaunchAttribute attr[1];
        unsigned n_attr = 0;
        const size_t hot_bytes = (size_t)grid * kStreamWarps * kWarpQueuesBytes;
        const int persist_mb = env_flag("VLB_BAKE_L2_PERSIST", 0);
        if (persist_mb > 0 && ctx->l2_persist_max > 0 && ctx->l2_window_max > 0) {
            const size_t want = std::min<size_t>((size_t)persist_mb << 20, (size_t)ctx->l2_persist_max);
            if (ctx->l2_persist_set != want) {
                VLB_CUDA(ctx, cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want));
                ctx->l2_persist_set = want;
            }
            // VLB_BAKE_L2_WINDOW=1 (A/B): the window covers the BVH nodes instead of the scratch
            const bool on_nodes = env_flag("VLB_BAKE_L2_WINDOW", 0) == 1;
            const size_t win = std::min(on_nodes ? ctx->d_nodes.cap : hot_bytes, (size_t)ctx->l2_window_max);
            attr[0].id = cudaLaunchAttributeAccessPolicyWindow;
            attr[0].val.accessPolicyWindow.base_ptr = on_nodes ? ctx->d_nodes.p : (void*)p.stream_scratch;
            attr[0].val.accessPolicyWindow.num_bytes = win;
            attr[0].val.accessPolicyWindow.hitRatio = win <= want ? 1.0f : (float)((double)want / (double)win);
            attr[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            attr[0].val.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
            n_attr = 1;
            ctx->l2_persist_dirty = true;
        }
        cfg.attrs = attr; cfg.numAttrs = n_attr;
        VLB_CUDA(ctx, cudaLaunchKernelEx(&cfg, kern, p));
    }
    VLB_LAUNCH_CHECK(ctx);
    VLB_CUDA(ctx, cudaEventRecord(ctx->ev[3], st));
    if (p.n_whole < n_probes) {
        const uint32_t n = (uint32_t)(n_probes - p.n_whole) * VLB_SH_STRIDE;
        k_sum_partials<<<(n + 255) / 256, 256, 0, st>>>(p, (uint32_t)n_probes);
        VLB_LAUNCH_CHECK(ctx);
    }
    if (s->flags & VLB_BAKE_ACCUMULATE_ACROSS_PROBES) {
        k_running_total<<<1, 64, 0, st>>>(d_out, (uint32_t)n_probes);
        VLB_LAUNCH_CHECK(ctx);
    }
    VLB_CUDA(ctx, cudaEventRecord(ctx->ev[1], st));
    // statistics + overflow flag travel to pinned host memory asynchronously; bake_collect_stats() waits for them.
    // (The axis table above was copied from pageable memory, which is staged before cudaMemcpyAsync returns.)
    if (!ctx->h_bake_stats) VLB_CUDA(ctx, cudaHostAlloc(reinterpret_cast<void**>(&ctx->h_bake_stats), 16 * sizeof(unsigned long long), cudaHostAllocDefault));
    ctx->h_bake_stats[3] = 0;
    VLB_CUDA(ctx, cudaMemcpyAsync(ctx->h_bake_stats, ctx->d_stats.p, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    VLB_CUDA(ctx, cudaMemcpyAsync(ctx->h_bake_stats + 3, ctx->d_scratch.as<float>() + 13, 4, cudaMemcpyDeviceToHost, st));
    VLB_CUDA(ctx, cudaMemcpyAsync(ctx->h_bake_stats + 4, ctx->d_stats.as<unsigned long long>() + 4, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
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
    if (ctx->l2_persist_dirty) {        // hand the persisting lines of the bake's scratch back to the kernels that follow
        ctx->l2_persist_dirty = false;
        VLB_CUDA(ctx, cudaCtxResetPersistingL2Cache());
    }
    vlb_bake_stats& b = ctx->last_bake;
    b.n_shadow_rays = ctx->h_bake_stats[0]; b.n_nodes_visited = ctx->h_bake_stats[1]; b.n_tris_tested = ctx->h_bake_stats[2];
    VLB_CUDA(ctx, cudaEventElapsedTime(&b.kernel_ms, ctx->ev[2], ctx->ev[3]));
    VLB_CUDA(ctx, cudaEventElapsedTime(&b.total_ms, ctx->ev[0], ctx->ev[1]));
    if (env_flag("VLB_BAKE_COUNTERS", 0) >= 2) {     // phase utilisation of the instrumented kernel (diagnostics)
        const unsigned long long* h = ctx->h_bake_stats;
        fprintf(stderr, "[vlb bake counters] node steps: %llu warp iterations, %.2f lanes active | leaf steps: %llu, %.2f lanes | "
                        "shade: %llu, %.2f lanes | outer iterations %llu | node steps with the stack beyond %d entries: %.4f %%\n",
                h[4], h[4] ? (double)h[5] / h[4] : 0.0, h[6], h[6] ? (double)h[7] / h[6] : 0.0, h[8], h[8] ? (double)h[9] / h[8] : 0.0,
                h[10], kSmemStack, h[1] ? 100.0 * h[11] / h[1] : 0.0);
    }
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

#endif   // !VLB_BAKE_GATHER_TU

}  // namespace vlb

// vlb_context.h — internal C++ side of libvlb_bake.so: the context object behind the C ABI
// (include/vlb_bake.h), device buffers, error plumbing and launch accounting.
#pragma once

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/vlb_bake.h"

#ifndef VLB_MAX_LANES
#define VLB_MAX_LANES 4
#endif

namespace vlb {

void set_thread_error(const char* msg);

// Owning device allocation. Grows on demand, never shrinks (buffers are reused across calls).
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes ? bytes : 16);
        if (e == cudaSuccess) cap = bytes ? bytes : 16;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct InstanceDev {       // device-side instance record used by the flatten kernel
    float m[12];
    float minv[9];
    uint32_t first_index, first_vertex, material, tri_offset;
    uint32_t index_count, vertex_count;
};

}  // namespace vlb

struct vlb_ctx {
    int device = 0;
    int sm_count = 0;
    int l2_persist_max = 0, l2_window_max = 0;   // device limits of the L2 set-aside and of one access-policy window (bytes)
    size_t l2_persist_set = 0;                   // cudaLimitPersistingL2CacheSize as last set by the bake
    bool l2_persist_dirty = false;               // a bake left persisting lines behind (reset when its stats are collected)
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    std::string err;
    uint64_t launches = 0;

    // ---- scene ----
    bool have_scene = false, have_bvh = false;
    uint64_t n_tris = 0, n_verts = 0, n_indices = 0;
    uint32_t n_insts = 0, n_mats = 0;
    float ref_bounds[6] = {0, 0, 0, 0, 0, 0};
    float tight_bounds[6] = {0, 0, 0, 0, 0, 0};
    bool have_tight_bounds = false;
    vlb::DevBuf d_verts, d_indices, d_insts_in, d_tri_offsets, d_inst_check;   // d_inst_check: per-instance local AABB + bad-index flag (k_instance_checks)
    vlb::DevBuf d_tri_flat;   // 3 float4 per triangle, flat order (before Morton sort)
    vlb::DevBuf d_tri_shade;  // 3 float4 per triangle, flat order
    vlb::DevBuf d_inst;       // 3 float4 per instance
    vlb::DevBuf d_base_color; // float4 per material (w = bits of the baseColor texture index, -1 = none)
    vlb::DevBuf d_tri_uv;     // 2 float4 per triangle, flat order: Vertex::uv0 of the three corners
    vlb::DevBuf d_tex_desc, d_tex_texels;   // vlb_scene_set_textures: int4 per texture, RGBA8 atlas
    uint32_t n_textures = 0;
    int max_tex_index = -1;   // largest baseColor texture index any material names
    // ---- BVH ----
    vlb::DevBuf d_tris;       // 3 float4 per triangle, Morton order
    vlb::DevBuf d_nodes;      // 4 float4 per node
    vlb::DevBuf d_keys, d_keys_sorted, d_vals, d_vals_sorted, d_sort_tmp;
    vlb::DevBuf d_frontier[2], d_frontier_n;   // wide-node roots of the current / next level of the collapse + level counters
    vlb::DevBuf d_left, d_right, d_first, d_last, d_parent_i, d_parent_l, d_flags, d_ibox, d_lbox, d_scratch;
    uint64_t n_nodes = 0;
    int max_leaf = 3;
    int bvh_builder = 0, ploc_radius = 16;      // vlb_bvh_set_builder
    vlb::DevBuf d_ploc, d_tris_alt, d_lbox_alt; // PLOC scratch; the triangles / leaf boxes in the tree's leaf order
    // ---- skybox ----
    vlb::DevBuf d_sky;        // RGBA32F
    int sky_w = 0, sky_h = 0;
    // vlb_skybox_set_async: the upload runs on its own stream, the next bake waits for it
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_sky_free = nullptr, ev_sky_ready = nullptr;
    bool sky_upload_pending = false;
    vlb::DevBuf d_sky_stage;  // RGBA8 staging of an asynchronous upload
    // ---- projection (skybox / envmap) ----
    vlb::DevBuf d_proj_in, d_proj_out, d_proj_partials, d_proj_counters, d_row_tab, d_col_tab;
    int tab_w = 0, tab_h = 0, tab_variant = -1;
    // Back-to-back projections of device-resident maps on the ctx's OWN stream are chained with programmatic dependent
    // launch (skybox_sh.cu): `proj_chain` says that the last thing this ctx enqueued on its stream was such a launch.
    // Every other API entry point clears it (check_device), so a chained launch never overtakes foreign work.
    bool proj_chain = false;
    struct TmapEntry { const void* ptr = nullptr; uint64_t stride = 0; uint32_t n_maps = 0; int W = 0, H = 0, fmt = -1; alignas(64) unsigned char map[128]; };
    TmapEntry tmap_cache[8];
    unsigned tmap_next = 0;
    // vlb_skybox_project_sh_device_ptrs: a call repeated with the same arguments is replayed from a CUDA graph of its
    // launches (one kernel node per map, lanes = parallel branches): issuing a launch costs the host ~6 us, a
    // 32 MiB map streams in 5
    struct ProjGraph {
        std::vector<const void*> maps; int fmt = 0, W = 0, H = 0, order = 0; float* out = nullptr;
        const void* partials = nullptr; const void* row_tab = nullptr;   // device buffers baked into the nodes
        int seen = 0; cudaGraphExec_t exec = nullptr;
    };
    std::vector<ProjGraph> proj_graphs;
    cudaStream_t cap_stream = nullptr;
    // ---- bake ----
    vlb::DevBuf d_bake_out, d_bake_prev, d_partials, d_work_counter, d_axis, d_row_sc, d_col_sc, d_stats, d_stream_scratch, d_stream_spill, d_vis_ovf, d_cell_root;
    int dir_w = 0, dir_h = 0;
    vlb::DevBuf d_dir_tab, d_proj_tab;            // per-direction tables of the bake (k_dir_tables)
    int dir_tab_key[5] = {0, 0, 0, 0, -1};        // (W, H, tile_lw, K, world_frame) the tables were built for
    cudaStream_t dir_tab_stream = nullptr;        // ... and the stream that built them (another stream rebuilds: no cross-stream order is assumed)
    vlb_bake_stats last_bake{};
    bool bake_pending = false;             // a device bake was enqueued and its statistics not yet collected
    unsigned long long* h_bake_stats = nullptr;   // pinned: [0] shadow rays, [1] nodes, [2] triangles, [3] stack overflow
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_done = nullptr;         // recorded after the last bake's statistics copies
    // ---- auxiliary lanes (independent skybox maps in flight side by side) ----
    cudaStream_t lane_stream[VLB_MAX_LANES] = {};
    cudaEvent_t lane_fork = nullptr, lane_join[VLB_MAX_LANES] = {};
    // ---- trace ----
    vlb::DevBuf d_ray_o, d_ray_d, d_hit_id, d_hit_tuv, d_hit_key;
    // ---- multi-GPU (comm.cu): one NCCL rank per ctx ----
    void* comm = nullptr;                  // ncclComm_t
    int comm_rank = 0, comm_world = 1;
    bool comm_sharded_uploads = false;     // scene / skybox uploads: every rank copies 1/world over PCIe, NVLink all-gather replicates
    vlb::DevBuf d_share, d_gather_stage;   // this rank's padded share of a sharded bake, all-gather staging [world][max share]

    int fail(int code, const char* fmt, ...) {
        char buf[512];
        va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
        err = buf;
        vlb::set_thread_error(buf);
        return code;
    }
};

#define VLB_CUDA(ctx, expr)                                                                      \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess)                                                                   \
            return (ctx)->fail(VLB_ERR_CUDA, "%s:%d: %s -> %s", __FILE__, __LINE__, #expr,        \
                               cudaGetErrorString(_e));                                          \
    } while (0)

#define VLB_LAUNCH_CHECK(ctx)                                                                    \
    do {                                                                                         \
        (ctx)->launches++;                                                                       \
        cudaError_t _e = cudaGetLastError();                                                     \
        if (_e != cudaSuccess)                                                                   \
            return (ctx)->fail(VLB_ERR_CUDA, "%s:%d: kernel launch -> %s", __FILE__, __LINE__,   \
                               cudaGetErrorString(_e));                                          \
    } while (0)

namespace vlb {
// host_tables.cpp
void host_axis_coords(float origin, float step, int n, float* out);
void host_dir_tables(int W, int H, float phi_shift, float* row_sc /*H x 2: sin,cos theta*/,
                     float* col_cs /*W x 2: cos,sin phi*/);
void host_proj_row_table(int W, int H, float* row_tab /*H x 8*/);
void host_inverse3x3(const float* m12, float* out9);
size_t ref_order_index(int i, int j, int k, int Nx, int Ny, int Nz);

// implemented in the .cu files
int scene_flatten(vlb_ctx* ctx);
int bvh_build(vlb_ctx* ctx, vlb_bvh_stats* stats);
int project_sh_device(vlb_ctx* ctx, const void* d_texels, uint64_t map_stride, uint32_t n_maps, int fmt,
                      int W, int H, int order, int variant, float* d_out, int lane = -1, bool chain_in_lane = false);
int bake_device(vlb_ctx* ctx, const vlb_bake_settings* s, const float* d_prev_full, float* d_out);
int bake_collect_stats(vlb_ctx* ctx);
int sky_upload_join(vlb_ctx* ctx);     // context.cu: orders a pending vlb_skybox_set_async before the ctx stream's next work
// comm.cu
int upload_replicated(vlb_ctx* ctx, void* d_dst, const void* h_src, size_t bytes, cudaStream_t st);
size_t comm_padded_bytes(const vlb_ctx* ctx, size_t bytes);   // capacity d_dst needs for upload_replicated
int comm_destroy(vlb_ctx* ctx);
int trace_rays(vlb_ctx* ctx, const float* o, const float* d, uint64_t n, float tmin, float tmax, int accel,
               int kind, int32_t* ids, float* tuv);
}  // namespace vlb

// context.cu — C ABI glue of libvlb_bake.so (include/vlb_bake.h): context life cycle, scene and
// skybox upload, host-pointer wrappers around the device paths. No compute happens on the host.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <new>
#include <thread>
#include <vector>

#include "vlb_context.h"
#include "vlb_math.cuh"

namespace vlb {

static thread_local std::string g_thread_error;
void set_thread_error(const char* msg) { g_thread_error = msg ? msg : ""; }

// ---- flatten: instances x primitives -> world-space triangle soup -------------
// One thread per flat triangle. Positions are read at the reference's 44-byte vertex stride
// (shader::Vertex, structures.h:20-26) and moved to world space by the instance transform
// (Node_t::getMatrix, src/scene_manager.cpp:445-461); normals stay in object space as
// env_map.rchit:59-64 reads them.
__global__ void k_flatten(const float* __restrict__ verts, const uint32_t* __restrict__ indices,
                          const InstanceDev* __restrict__ insts, const uint32_t* __restrict__ tri_offsets,
                          uint32_t n_insts, uint32_t n_tris, float4* __restrict__ tri_flat,
                          float4* __restrict__ tri_shade, float4* __restrict__ tri_uv) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tris) return;
    // binary search: last instance with tri_offset <= t
    uint32_t lo = 0, hi = n_insts;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (tri_offsets[mid] <= t) lo = mid; else hi = mid;
    }
    const InstanceDev& in = insts[lo];
    const uint32_t local = t - tri_offsets[lo];
    const uint32_t* ix = indices + in.first_index + 3ull * local;
    Vec3 p[3], n[3];
    float uv[6];
    for (int k = 0; k < 3; ++k) {
        // an out-of-range index is reported by k_instance_checks (the call then fails); here it must just not read outside the array
        const float* v = verts + 11ull * (in.first_vertex + min(ix[k], in.vertex_count ? in.vertex_count - 1u : 0u));
        p[k] = xform_point(in.m, mk3(v[0], v[1], v[2]));
        n[k] = mk3(v[4], v[5], v[6]);
        uv[2 * k] = v[7]; uv[2 * k + 1] = v[8];                          // Vertex::uv0 (structures.h:24)
    }
    tri_uv[2ull * t + 0] = make_float4(uv[0], uv[1], uv[2], uv[3]);
    tri_uv[2ull * t + 1] = make_float4(uv[4], uv[5], 0.f, 0.f);
    tri_flat[3ull * t + 0] = make_float4(p[0].x, p[0].y, p[0].z, __int_as_float((int)t));
    tri_flat[3ull * t + 1] = make_float4(f_sub(p[1].x, p[0].x), f_sub(p[1].y, p[0].y), f_sub(p[1].z, p[0].z), 0.f);
    tri_flat[3ull * t + 2] = make_float4(f_sub(p[2].x, p[0].x), f_sub(p[2].y, p[0].y), f_sub(p[2].z, p[0].z), 0.f);
    tri_shade[3ull * t + 0] = make_float4(n[0].x, n[0].y, n[0].z, __int_as_float((int)lo));
    tri_shade[3ull * t + 1] = make_float4(n[1].x, n[1].y, n[1].z, 0.f);
    tri_shade[3ull * t + 2] = make_float4(n[2].x, n[2].y, n[2].z, 0.f);
}

// Per instance (one block each): the local AABB of its vertices (Scene_t::loadNode takes its two corners through the node
// matrix, src/scene_manager.cpp:497-507) and whether any of its indices points past its vertex range. Replaces two host
// loops over every vertex and index of the scene that vlb_scene_set_triangles ran before it could upload anything.
// out[i] = {lo.xyz, hi.xyz, bad-index flag, 0} as 8 floats.
__global__ void __launch_bounds__(256) k_instance_checks(const float* __restrict__ verts, const uint32_t* __restrict__ indices,
                                                         const InstanceDev* __restrict__ insts, float* __restrict__ out) {
    const InstanceDev& in = insts[blockIdx.x];
    float lo[3] = {3.4e38f, 3.4e38f, 3.4e38f}, hi[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
    for (uint32_t v = threadIdx.x; v < in.vertex_count; v += blockDim.x) {
        const float* p = verts + 11ull * (in.first_vertex + v);
        for (int k = 0; k < 3; ++k) { lo[k] = fminf(lo[k], p[k]); hi[k] = fmaxf(hi[k], p[k]); }
    }
    int bad = 0;
    for (uint32_t k = threadIdx.x; k < in.index_count; k += blockDim.x) bad |= indices[in.first_index + k] >= in.vertex_count;
    __shared__ float s_lo[3][8], s_hi[3][8];
    __shared__ int s_bad[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int off = 16; off > 0; off >>= 1) {
        for (int k = 0; k < 3; ++k) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], off));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], off));
        }
        bad |= __shfl_xor_sync(0xffffffffu, bad, off);
    }
    if (lane == 0) { for (int k = 0; k < 3; ++k) { s_lo[k][warp] = lo[k]; s_hi[k][warp] = hi[k]; } s_bad[warp] = bad; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) {
            for (int k = 0; k < 3; ++k) { lo[k] = fminf(lo[k], s_lo[k][w]); hi[k] = fmaxf(hi[k], s_hi[k][w]); }
            bad |= s_bad[w];
        }
        float* o = out + 8ull * blockIdx.x;
        for (int k = 0; k < 3; ++k) { o[k] = lo[k]; o[3 + k] = hi[k]; }
        o[6] = bad ? 1.0f : 0.0f; o[7] = 0.0f;
    }
}

__global__ void k_rgba8_to_f32(const uchar4* __restrict__ in, float4* __restrict__ out, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uchar4 c = in[i];
    out[i] = make_float4((float)c.x / 255.0f, (float)c.y / 255.0f, (float)c.z / 255.0f, (float)c.w / 255.0f);
}

static int env_lanes() {
    const char* v = getenv("VLB_PROJ_LANES");
    return v && *v ? atoi(v) : VLB_MAX_LANES;
}

static int check_device(vlb_ctx* ctx) {
    ctx->proj_chain = false;       // whatever this call enqueues, the next projection launch must not overtake it (skybox_sh.cu)
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return ctx->fail(VLB_ERR_NO_DEVICE, "cudaSetDevice(%d): %s", ctx->device, cudaGetErrorString(e));
    return VLB_OK;
}

}  // namespace vlb

using namespace vlb;
#ifdef VLB_PROJ_TIMING
namespace vlb { int proj_timing_dump(vlb_ctx* ctx, int n_launches); }
#endif

extern "C" {

int vlb_abi_version(void) { return VLB_ABI_VERSION; }

const char* vlb_last_error(const vlb_ctx* ctx) {
    if (ctx) return ctx->err.c_str();
    return g_thread_error.c_str();
}

int vlb_ctx_create(int device_id, vlb_ctx** out) {
    if (!out) { set_thread_error("vlb_ctx_create: out is NULL"); return VLB_ERR_INVALID; }
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        char buf[256];
        snprintf(buf, sizeof buf, "vlb_ctx_create: no CUDA device (%s); this library has no CPU path",
                 e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        set_thread_error(buf);
        cudaGetLastError();
        return VLB_ERR_NO_DEVICE;
    }
    if (device_id < 0 || device_id >= count) { set_thread_error("vlb_ctx_create: device id out of range"); return VLB_ERR_INVALID; }
    vlb_ctx* ctx = new (std::nothrow) vlb_ctx();
    if (!ctx) { set_thread_error("vlb_ctx_create: out of host memory"); return VLB_ERR_NOMEM; }
    ctx->device = device_id;
    e = cudaSetDevice(device_id);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking);
    cudaDeviceProp prop;
    if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, device_id);
    for (int i = 0; i < 4 && e == cudaSuccess; ++i) e = cudaEventCreate(&ctx->ev[i]);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_done, cudaEventDisableTiming);
    if (e != cudaSuccess) {
        set_thread_error(cudaGetErrorString(e));
        delete ctx;
        return VLB_ERR_CUDA;
    }
    ctx->sm_count = prop.multiProcessorCount;
    ctx->l2_persist_max = prop.persistingL2CacheMaxSize;
    ctx->l2_window_max = prop.accessPolicyMaxWindowSize;
    ctx->stream = ctx->own_stream;
    *out = ctx;
    return VLB_OK;
}

void vlb_ctx_destroy(vlb_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    comm_destroy(ctx);
    DevBuf* bufs[] = {&ctx->d_inst_check, &ctx->d_share, &ctx->d_gather_stage, &ctx->d_vis_ovf, &ctx->d_cell_root, &ctx->d_ploc, &ctx->d_tris_alt, &ctx->d_lbox_alt, &ctx->d_verts, &ctx->d_indices, &ctx->d_insts_in, &ctx->d_tri_offsets, &ctx->d_tri_flat,
                      &ctx->d_frontier[0], &ctx->d_frontier[1], &ctx->d_frontier_n, &ctx->d_tri_shade, &ctx->d_tri_uv, &ctx->d_tex_desc, &ctx->d_tex_texels, &ctx->d_inst, &ctx->d_base_color, &ctx->d_tris, &ctx->d_nodes, &ctx->d_keys,
                      &ctx->d_keys_sorted, &ctx->d_vals, &ctx->d_vals_sorted, &ctx->d_sort_tmp, &ctx->d_left,
                      &ctx->d_right, &ctx->d_first, &ctx->d_last, &ctx->d_parent_i, &ctx->d_parent_l, &ctx->d_flags,
                      &ctx->d_ibox, &ctx->d_lbox, &ctx->d_scratch, &ctx->d_sky, &ctx->d_proj_in, &ctx->d_proj_out,
                      &ctx->d_proj_partials, &ctx->d_proj_counters, &ctx->d_row_tab, &ctx->d_col_tab, &ctx->d_bake_out, &ctx->d_bake_prev,
                      &ctx->d_partials, &ctx->d_work_counter, &ctx->d_axis, &ctx->d_row_sc, &ctx->d_col_sc,
                      &ctx->d_stats, &ctx->d_stream_scratch, &ctx->d_ray_o, &ctx->d_ray_d, &ctx->d_hit_id, &ctx->d_hit_tuv, &ctx->d_hit_key};
    if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); }
    for (vlb_ctx::ProjGraph& g : ctx->proj_graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
    if (ctx->cap_stream) cudaStreamDestroy(ctx->cap_stream);
    if (ctx->ev_sky_free) cudaEventDestroy(ctx->ev_sky_free);
    if (ctx->ev_sky_ready) cudaEventDestroy(ctx->ev_sky_ready);
    ctx->d_sky_stage.release();
    for (DevBuf* b : bufs) b->release();
    for (int i = 0; i < 4; ++i) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    if (ctx->ev_done) cudaEventDestroy(ctx->ev_done);
    for (int l = 0; l < VLB_MAX_LANES; ++l) {
        if (ctx->lane_stream[l]) { cudaStreamSynchronize(ctx->lane_stream[l]); cudaStreamDestroy(ctx->lane_stream[l]); }
        if (ctx->lane_join[l]) cudaEventDestroy(ctx->lane_join[l]);
    }
    if (ctx->lane_fork) cudaEventDestroy(ctx->lane_fork);
    if (ctx->h_bake_stats) cudaFreeHost(ctx->h_bake_stats);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

int vlb_ctx_set_stream(vlb_ctx* ctx, uint64_t handle) {
    if (!ctx) return VLB_ERR_INVALID;
    ctx->stream = handle == VLB_STREAM_OWN ? ctx->own_stream : reinterpret_cast<cudaStream_t>(handle);
    ctx->proj_chain = false;
    return VLB_OK;
}

uint64_t vlb_ctx_stream(const vlb_ctx* ctx) { return ctx ? (uint64_t)reinterpret_cast<uintptr_t>(ctx->stream) : 0; }

static int join_sky_upload(vlb_ctx* ctx, cudaStream_t st);

int vlb_ctx_synchronize(vlb_ctx* ctx) {
    if (!ctx) return VLB_ERR_INVALID;
    if (int r = check_device(ctx)) return r;
    if (int r = join_sky_upload(ctx, ctx->stream)) return r;
    VLB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return VLB_OK;
}

uint64_t vlb_ctx_launch_count(const vlb_ctx* ctx) { return ctx ? ctx->launches : 0; }

int vlb_scene_set_triangles(vlb_ctx* ctx, const vlb_vertex* vertices, uint64_t n_vertices,
                            const uint32_t* indices, uint64_t n_indices, const vlb_instance* instances,
                            uint32_t n_instances, const vlb_material* materials, uint32_t n_materials) {
    if (!ctx) return VLB_ERR_INVALID;
    if (int r = check_device(ctx)) return r;
    static_assert(sizeof(vlb_vertex) == 44, "shader::Vertex is 44 bytes (structures.h:20-26)");
    static_assert(sizeof(vlb_material) == 144, "shader::Material is 144 bytes (structures.h:28-71)");
    if ((n_vertices && !vertices) || (n_indices && !indices) || (n_instances && !instances) ||
        (n_materials && !materials))
        return ctx->fail(VLB_ERR_INVALID, "vlb_scene_set_triangles: NULL array with non-zero count");
    ctx->have_scene = false; ctx->have_bvh = false; ctx->have_tight_bounds = false;

    // per-instance records + reference bounds (Scene_t::loadNode, src/scene_manager.cpp:497-507:
    // bounds start at the origin and take the two transformed corners of the local AABB)
    std::vector<InstanceDev> insts(n_instances);
    std::vector<uint32_t> offsets(n_instances + 1, 0);
    std::vector<float4> inst_rec(3 * (size_t)n_instances);
    for (int k = 0; k < 6; ++k) ctx->ref_bounds[k] = 0.f;
    uint64_t tri_total = 0;
    for (uint32_t i = 0; i < n_instances; ++i) {
        const vlb_instance& vi = instances[i];
        if ((uint64_t)vi.first_index + vi.index_count > n_indices || (uint64_t)vi.first_vertex + vi.vertex_count > n_vertices)
            return ctx->fail(VLB_ERR_INVALID, "vlb_scene_set_triangles: instance %u exceeds the vertex/index arrays", i);
        InstanceDev& d = insts[i];
        std::memcpy(d.m, vi.transform, sizeof d.m);
        host_inverse3x3(d.m, d.minv);
        d.first_index = vi.first_index; d.first_vertex = vi.first_vertex;
        // default material = last entry (src/scene_manager.cpp:510, 851)
        d.material = vi.material_index < n_materials ? vi.material_index : (n_materials ? n_materials - 1 : 0);
        d.tri_offset = (uint32_t)tri_total;
        d.index_count = vi.index_count; d.vertex_count = vi.vertex_count;
        offsets[i] = (uint32_t)tri_total;
        tri_total += vi.index_count / 3;
        inst_rec[3 * i + 0] = make_float4(d.minv[0], d.minv[1], d.minv[2], 0.f);
        std::memcpy(&inst_rec[3 * i + 0].w, &d.material, 4);
        inst_rec[3 * i + 1] = make_float4(d.minv[3], d.minv[4], d.minv[5], 0.f);
        inst_rec[3 * i + 2] = make_float4(d.minv[6], d.minv[7], d.minv[8], 0.f);
        if (vi.index_count && vi.vertex_count == 0)
            return ctx->fail(VLB_ERR_INVALID, "vlb_scene_set_triangles: index out of range in instance %u", i);
    }
    offsets[n_instances] = (uint32_t)tri_total;
    if (tri_total >= (1ull << 28)) return ctx->fail(VLB_ERR_UNSUPPORTED, "more than 2^28 triangles");

    // resolved baseColor per material (env_map.rchit:36-49): rgb = factor (or 1), w = bits(texture index or -1)
    const int no_tex = -1;
    float no_tex_f; std::memcpy(&no_tex_f, &no_tex, 4);
    std::vector<float4> base(std::max<uint32_t>(n_materials, 1), make_float4(1.f, 1.f, 1.f, no_tex_f));
    ctx->max_tex_index = -1;
    for (uint32_t m = 0; m < n_materials; ++m) {
        const float* f = materials[m].base_color_factor;
        if (f[0] != 0.f || f[1] != 0.f || f[2] != 0.f || f[3] != 0.f) base[m] = make_float4(f[0], f[1], f[2], no_tex_f);
        const int32_t ti = materials[m].base_color.index;
        if (ti < -1) return ctx->fail(VLB_ERR_INVALID, "vlb_scene_set_triangles: material %u has baseColor texture index %d", m, ti);
        if (ti >= 0) { std::memcpy(&base[m].w, &ti, 4); ctx->max_tex_index = std::max(ctx->max_tex_index, (int)ti); }
    }

    ctx->n_tris = tri_total; ctx->n_verts = n_vertices; ctx->n_indices = n_indices;
    ctx->n_insts = n_instances; ctx->n_mats = n_materials;
    VLB_CUDA(ctx, ctx->d_verts.reserve(comm_padded_bytes(ctx, n_vertices * sizeof(vlb_vertex))));
    VLB_CUDA(ctx, ctx->d_indices.reserve(comm_padded_bytes(ctx, n_indices * sizeof(uint32_t))));
    VLB_CUDA(ctx, ctx->d_insts_in.reserve(insts.size() * sizeof(InstanceDev)));
    VLB_CUDA(ctx, ctx->d_tri_offsets.reserve(offsets.size() * sizeof(uint32_t)));
    VLB_CUDA(ctx, ctx->d_inst.reserve(inst_rec.size() * sizeof(float4)));
    VLB_CUDA(ctx, ctx->d_base_color.reserve(base.size() * sizeof(float4)));
    VLB_CUDA(ctx, ctx->d_tri_flat.reserve(3 * tri_total * sizeof(float4)));
    VLB_CUDA(ctx, ctx->d_tri_shade.reserve(3 * tri_total * sizeof(float4)));
    VLB_CUDA(ctx, ctx->d_tri_uv.reserve(2 * tri_total * sizeof(float4)));
    cudaStream_t st = ctx->stream;
    // the two big arrays: plain H2D copies, or (vlb_comm_sharded_uploads) 1/world over PCIe + an NVLink all-gather
    if (int r = upload_replicated(ctx, ctx->d_verts.p, vertices, n_vertices * sizeof(vlb_vertex), st)) return r;
    if (int r = upload_replicated(ctx, ctx->d_indices.p, indices, n_indices * sizeof(uint32_t), st)) return r;
    if (n_instances) {
        VLB_CUDA(ctx, cudaMemcpyAsync(ctx->d_insts_in.p, insts.data(), insts.size() * sizeof(InstanceDev), cudaMemcpyHostToDevice, st));
        VLB_CUDA(ctx, cudaMemcpyAsync(ctx->d_inst.p, inst_rec.data(), inst_rec.size() * sizeof(float4), cudaMemcpyHostToDevice, st));
    }
    VLB_CUDA(ctx, cudaMemcpyAsync(ctx->d_tri_offsets.p, offsets.data(), offsets.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    VLB_CUDA(ctx, cudaMemcpyAsync(ctx->d_base_color.p, base.data(), base.size() * sizeof(float4), cudaMemcpyHostToDevice, st));
    if (tri_total) {
        const uint32_t n = (uint32_t)tri_total;
        k_flatten<<<(n + 255) / 256, 256, 0, st>>>(ctx->d_verts.as<float>(), ctx->d_indices.as<uint32_t>(),
                                                   ctx->d_insts_in.as<InstanceDev>(), ctx->d_tri_offsets.as<uint32_t>(),
                                                   n_instances, n, ctx->d_tri_flat.as<float4>(), ctx->d_tri_shade.as<float4>(),
                                                   ctx->d_tri_uv.as<float4>());
        VLB_LAUNCH_CHECK(ctx);
    }
    // per-instance local AABBs + index validation on the device (k_instance_checks), read back with the sync below
    std::vector<float> chk(8 * (size_t)n_instances);
    if (n_instances) {
        VLB_CUDA(ctx, ctx->d_inst_check.reserve(chk.size() * sizeof(float)));
        k_instance_checks<<<n_instances, 256, 0, st>>>(ctx->d_verts.as<float>(), ctx->d_indices.as<uint32_t>(), ctx->d_insts_in.as<InstanceDev>(),
                                                       ctx->d_inst_check.as<float>());
        VLB_LAUNCH_CHECK(ctx);
        VLB_CUDA(ctx, cudaMemcpyAsync(chk.data(), ctx->d_inst_check.p, chk.size() * sizeof(float), cudaMemcpyDeviceToHost, st));
    }
    // host staging vectors die at scope exit: the copies above must have completed
    VLB_CUDA(ctx, cudaStreamSynchronize(st));
    for (uint32_t i = 0; i < n_instances; ++i) {
        const float* c = &chk[8 * (size_t)i];
        if (c[6] != 0.0f) return ctx->fail(VLB_ERR_INVALID, "vlb_scene_set_triangles: index out of range in instance %u", i);
        if (instances[i].vertex_count == 0) continue;
        // reference bounds (Scene_t::loadNode, src/scene_manager.cpp:497-507): the two corners of the local AABB through the node matrix
        const float* m = insts[i].m;
        for (int r = 0; r < 3; ++r) {
            const float a = fmaf(m[4 * r + 2], c[2], fmaf(m[4 * r + 1], c[1], fmaf(m[4 * r + 0], c[0], m[4 * r + 3])));
            const float b = fmaf(m[4 * r + 2], c[5], fmaf(m[4 * r + 1], c[4], fmaf(m[4 * r + 0], c[3], m[4 * r + 3])));
            ctx->ref_bounds[r] = std::min(ctx->ref_bounds[r], a);
            ctx->ref_bounds[3 + r] = std::max(ctx->ref_bounds[3 + r], b);
        }
    }
    ctx->have_scene = true;
    return VLB_OK;
}

int vlb_scene_set_textures(vlb_ctx* ctx, const vlb_texture* textures, uint32_t n_textures) {
    if (!ctx) return VLB_ERR_INVALID;
    if (int r = check_device(ctx)) return r;
    if (n_textures && !textures) return ctx->fail(VLB_ERR_INVALID, "vlb_scene_set_textures: null textures");
    std::vector<int4> desc(n_textures);
    size_t total = 0;
    for (uint32_t i = 0; i < n_textures; ++i) {
        const vlb_texture& t = textures[i];
        if (!t.texels || t.width <= 0 || t.height <= 0 || t.width > 32768 || t.height > 32768)
            return ctx->fail(VLB_ERR_INVALID, "vlb_scene_set_textures: texture %u has no texels or a bad size", i);
        if (t.wrap_u < 0 || t.wrap_u > 2 || t.wrap_v < 0 || t.wrap_v > 2 || t.filter < 0 || t.filter > 1)
            return ctx->fail(VLB_ERR_INVALID, "vlb_scene_set_textures: texture %u has an unknown wrap mode or filter", i);
        if (total + (size_t)t.width * t.height >= (1ull << 31))
            return ctx->fail(VLB_ERR_UNSUPPORTED, "vlb_scene_set_textures: more than 2^31 texels in total");
        desc[i] = make_int4((int)total, t.width, t.height, t.wrap_u | (t.wrap_v << 2) | (t.filter << 4));
        total += (size_t)t.width * t.height;
    }
    cudaStream_t st = ctx->stream;
    VLB_CUDA(ctx, ctx->d_tex_desc.reserve(std::max<size_t>(n_textures, 1) * sizeof(int4)));
    VLB_CUDA(ctx, ctx->d_tex_texels.reserve(std::max<size_t>(total, 1) * 4));
    if (n_textures) VLB_CUDA(ctx, cudaMemcpyAsync(ctx->d_tex_desc.p, desc.data(), n_textures * sizeof(int4), cudaMemcpyHostToDevice, st));
    for (uint32_t i = 0; i < n_textures; ++i)
        VLB_CUDA(ctx, cudaMemcpyAsync(static_cast<unsigned char*>(ctx->d_tex_texels.p) + 4 * (size_t)desc[i].x, textures[i].texels,
                                      4 * (size_t)textures[i].width * textures[i].height, cudaMemcpyHostToDevice, st));
    VLB_CUDA(ctx, cudaStreamSynchronize(st));     // caller-owned host buffers may go away after the call
    ctx->n_textures = n_textures;
    return VLB_OK;
}

int vlb_scene_bounds(vlb_ctx* ctx, int tight, float out[6]) {
    if (!ctx || !out) return VLB_ERR_INVALID;
    if (!ctx->have_scene) return ctx->fail(VLB_ERR_STATE, "vlb_scene_bounds: no scene set");
    if (!tight) { std::memcpy(out, ctx->ref_bounds, sizeof ctx->ref_bounds); return VLB_OK; }
    if (!ctx->have_tight_bounds) {
        int r = vlb_bvh_build(ctx, nullptr);   // the build computes the tight AABB on the device
        if (r) return r;
    }
    std::memcpy(out, ctx->tight_bounds, sizeof ctx->tight_bounds);
    return VLB_OK;
}

int vlb_bvh_set_builder(vlb_ctx* ctx, int builder, int ploc_radius) {
    if (!ctx) return VLB_ERR_INVALID;
    if (builder != VLB_BVH_BUILDER_LBVH && builder != VLB_BVH_BUILDER_PLOC) return ctx->fail(VLB_ERR_INVALID, "vlb_bvh_set_builder: unknown builder %d", builder);
    if (builder == VLB_BVH_BUILDER_PLOC && (ploc_radius < 0 || ploc_radius > 64)) return ctx->fail(VLB_ERR_INVALID, "vlb_bvh_set_builder: PLOC radius must be 1..64 (0 = default)");
    ctx->bvh_builder = builder;
    if (ploc_radius > 0) ctx->ploc_radius = ploc_radius;
    ctx->have_bvh = false;
    return VLB_OK;
}

int vlb_bvh_recommend_builder(uint64_t n_triangles, uint64_t n_primary_rays) {
    // PLOC costs ~2.3 ns per triangle more than the LBVH (1.06 vs 0.46 ms at 262,144 triangles) and takes ~4.2 % off
    // ~0.26 ns per primary ray (C3 at 138 ms, shadow rays included): worth it above ~215 primary rays per triangle.
    // Measured without gain at 1 M and 3 M triangles.
    if (n_triangles < 4096 || n_triangles > (1u << 20)) return VLB_BVH_BUILDER_LBVH;    // a tiny tree has nothing to gain
    return n_primary_rays > 215ull * n_triangles ? VLB_BVH_BUILDER_PLOC : VLB_BVH_BUILDER_LBVH;
}

int vlb_bvh_build(vlb_ctx* ctx, vlb_bvh_stats* stats) {
    if (!ctx) return VLB_ERR_INVALID;
    if (int r = check_device(ctx)) return r;
    if (!ctx->have_scene) return ctx->fail(VLB_ERR_STATE, "vlb_bvh_build: no scene set");
    return bvh_build(ctx, stats);
}

// An asynchronous skybox upload still in flight is ordered before whatever `st` does next.
static int join_sky_upload(vlb_ctx* ctx, cudaStream_t st) {
    if (ctx->sky_upload_pending) {
        VLB_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_sky_ready, 0));
        ctx->sky_upload_pending = false;
    }
    return VLB_OK;
}

int vlb_skybox_set_async(vlb_ctx* ctx, const void* texels, int format, int width, int height) {
    if (!ctx) return VLB_ERR_INVALID;
    if (int r = check_device(ctx)) return r;
    if (!texels || width <= 0 || height <= 0 || (format != VLB_FMT_RGBA8 && format != VLB_FMT_RGBA32F))
        return ctx->fail(VLB_ERR_INVALID, "vlb_skybox_set_async: bad arguments");
    const size_t n = (size_t)width * height;
    if (!ctx->copy_stream) {
        VLB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        VLB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_sky_free, cudaEventDisableTiming));
        VLB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_sky_ready, cudaEventDisableTiming));
    }
    VLB_CUDA(ctx, ctx->d_sky.reserve(comm_padded_bytes(ctx, n * sizeof(float4))));
    // whatever the ctx stream has queued (a bake sampling the old skybox) finishes before the texels are replaced
    VLB_CUDA(ctx, cudaEventRecord(ctx->ev_sky_free, ctx->stream));
    VLB_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_sky_free, 0));
    if (format == VLB_FMT_RGBA32F) {
        if (int r = upload_replicated(ctx, ctx->d_sky.p, texels, n * sizeof(float4), ctx->copy_stream)) return r;
    } else {
        VLB_CUDA(ctx, ctx->d_sky_stage.reserve(comm_padded_bytes(ctx, n * 4)));
        if (int r = upload_replicated(ctx, ctx->d_sky_stage.p, texels, n * 4, ctx->copy_stream)) return r;
        k_rgba8_to_f32<<<(unsigned)((n + 255) / 256), 256, 0, ctx->copy_stream>>>(ctx->d_sky_stage.as<uchar4>(), ctx->d_sky.as<float4>(), n);
        VLB_LAUNCH_CHECK(ctx);
    }
    VLB_CUDA(ctx, cudaEventRecord(ctx->ev_sky_ready, ctx->copy_stream));
    ctx->sky_upload_pending = true;
    ctx->sky_w = width; ctx->sky_h = height;
    return VLB_OK;
}

int vlb_skybox_set(vlb_ctx* ctx, const void* texels, int format, int width, int height) {
    if (!ctx) return VLB_ERR_INVALID;
    if (int r = check_device(ctx)) return r;
    if (!texels || width <= 0 || height <= 0 || (format != VLB_FMT_RGBA8 && format != VLB_FMT_RGBA32F))
        return ctx->fail(VLB_ERR_INVALID, "vlb_skybox_set: bad arguments");
    if (int r = join_sky_upload(ctx, ctx->stream)) return r;
    const size_t n = (size_t)width * height;
    VLB_CUDA(ctx, ctx->d_sky.reserve(comm_padded_bytes(ctx, n * sizeof(float4))));
    if (format == VLB_FMT_RGBA32F) {
        if (int r = upload_replicated(ctx, ctx->d_sky.p, texels, n * sizeof(float4), ctx->stream)) return r;
    } else {
        VLB_CUDA(ctx, ctx->d_proj_in.reserve(comm_padded_bytes(ctx, n * 4)));
        if (int r = upload_replicated(ctx, ctx->d_proj_in.p, texels, n * 4, ctx->stream)) return r;
        k_rgba8_to_f32<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_proj_in.as<uchar4>(), ctx->d_sky.as<float4>(), n);
        VLB_LAUNCH_CHECK(ctx);
    }
    VLB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->sky_w = width; ctx->sky_h = height;
    return VLB_OK;
}

static int project_host(vlb_ctx* ctx, const void* const* maps, uint32_t n_maps, int format, int W, int H, int order,
                        int variant, float* out) {
    if (!ctx) return VLB_ERR_INVALID;
    if (int r = check_device(ctx)) return r;
    if (!maps || !out || n_maps == 0 || W <= 0 || H <= 0 || (order != 2 && order != 3) ||
        (format != VLB_FMT_RGBA8 && format != VLB_FMT_RGBA32F))
        return ctx->fail(VLB_ERR_INVALID, "project_sh: bad arguments");
    const size_t bpt = format == VLB_FMT_RGBA32F ? 16 : 4;
    const size_t map_bytes = (size_t)W * H * bpt;
    const size_t stride = (map_bytes + 255) & ~size_t(255);
    VLB_CUDA(ctx, ctx->d_proj_in.reserve(stride * n_maps));
    VLB_CUDA(ctx, ctx->d_proj_out.reserve((size_t)n_maps * VLB_SH_STRIDE * sizeof(float)));
    for (uint32_t i = 0; i < n_maps; ++i) {
        if (!maps[i]) return ctx->fail(VLB_ERR_INVALID, "project_sh: maps[%u] is NULL", i);
        VLB_CUDA(ctx, cudaMemcpyAsync((char*)ctx->d_proj_in.p + stride * i, maps[i], map_bytes, cudaMemcpyHostToDevice, ctx->stream));
    }
    int r = project_sh_device(ctx, ctx->d_proj_in.p, stride, n_maps, format, W, H, order, variant, ctx->d_proj_out.as<float>());
    if (r) return r;
    VLB_CUDA(ctx, cudaMemcpyAsync(out, ctx->d_proj_out.p, (size_t)n_maps * VLB_SH_STRIDE * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    VLB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return VLB_OK;
}

int vlb_skybox_project_sh(vlb_ctx* ctx, const void* texels, int format, int width, int height, int sh_order, float* out48) {
    const void* maps[1] = {texels};
    return project_host(ctx, texels ? maps : nullptr, 1, format, width, height, sh_order, 0, out48);
}

int vlb_skybox_project_sh_batched(vlb_ctx* ctx, const void* const* maps, uint32_t n_maps, int format, int width,
                                  int height, int sh_order, float* out) {
    return project_host(ctx, maps, n_maps, format, width, height, sh_order, 0, out);
}

int vlb_envmap_project_sh(vlb_ctx* ctx, const void* texels, int format, int width, int height, int sh_order, float* out48) {
    const void* maps[1] = {texels};
    return project_host(ctx, texels ? maps : nullptr, 1, format, width, height, sh_order, 1, out48);
}

int vlb_skybox_project_sh_device(vlb_ctx* ctx, const void* d_texels, uint64_t map_stride_bytes, uint32_t n_maps,
                                 int format, int width, int height, int sh_order, float* d_out) {
    if (!ctx) return VLB_ERR_INVALID;
    const bool chain_was_open = ctx->proj_chain;
    if (int r = check_device(ctx)) return r;
    if (!d_texels || !d_out || n_maps == 0 || width <= 0 || height <= 0 || (sh_order != 2 && sh_order != 3) ||
        (format != VLB_FMT_RGBA8 && format != VLB_FMT_RGBA32F))
        return ctx->fail(VLB_ERR_INVALID, "vlb_skybox_project_sh_device: bad arguments");
    ctx->proj_chain = chain_was_open;      // check_device closed it; the launch below may chain onto the previous projection
    return project_sh_device(ctx, d_texels, map_stride_bytes, n_maps, format, width, height, sh_order, 0, d_out);
}

// The launches of one vlb_skybox_project_sh_device_ptrs call: one per map, alternating over auxiliary streams forked from /
// joined to `origin`; the second and later launches of a lane are chained to their predecessor (programmatic dependent launch).
static int project_ptrs_lanes(vlb_ctx* ctx, cudaStream_t origin, const void* const* d_maps, uint32_t n_maps, int format, int width,
                              int height, int sh_order, float* d_out) {
    const int lanes = std::max(1, std::min<int>({env_lanes(), VLB_MAX_LANES, (int)n_maps}));
    if (!ctx->lane_fork) VLB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->lane_fork, cudaEventDisableTiming));
    for (int l = 0; l < lanes; ++l) {
        if (!ctx->lane_stream[l]) VLB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->lane_stream[l], cudaStreamNonBlocking));
        if (!ctx->lane_join[l]) VLB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->lane_join[l], cudaEventDisableTiming));
    }
    VLB_CUDA(ctx, cudaEventRecord(ctx->lane_fork, origin));
    for (int l = 0; l < lanes; ++l) VLB_CUDA(ctx, cudaStreamWaitEvent(ctx->lane_stream[l], ctx->lane_fork, 0));
    int rc = VLB_OK;
    for (uint32_t i = 0; i < n_maps && rc == VLB_OK; ++i)
        rc = project_sh_device(ctx, d_maps[i], 0, 1, format, width, height, sh_order, 0,
                               d_out + (size_t)i * VLB_SH_STRIDE, (int)(i % lanes), /*chain_in_lane=*/i >= (uint32_t)lanes);
    for (int l = 0; l < lanes; ++l) {      // always join, also after an error, so the streams stay ordered
        VLB_CUDA(ctx, cudaEventRecord(ctx->lane_join[l], ctx->lane_stream[l]));
        VLB_CUDA(ctx, cudaStreamWaitEvent(origin, ctx->lane_join[l], 0));
    }
    return rc;
}

static int env_proj_graph() {
    static const int v = [] { const char* e = getenv("VLB_PROJ_GRAPH"); return e && *e ? atoi(e) : 1; }();
    return v;
}

int vlb_skybox_project_sh_device_ptrs(vlb_ctx* ctx, const void* const* d_maps, uint32_t n_maps, int format, int width,
                                      int height, int sh_order, float* d_out) {
    if (!ctx) return VLB_ERR_INVALID;
    if (int r = check_device(ctx)) return r;
    if (!d_maps || !d_out || n_maps == 0 || width <= 0 || height <= 0 || (sh_order != 2 && sh_order != 3) ||
        (format != VLB_FMT_RGBA8 && format != VLB_FMT_RGBA32F))
        return ctx->fail(VLB_ERR_INVALID, "vlb_skybox_project_sh_device_ptrs: bad arguments");
    for (uint32_t i = 0; i < n_maps; ++i)
        if (!d_maps[i]) return ctx->fail(VLB_ERR_INVALID, "vlb_skybox_project_sh_device_ptrs: d_maps[%u] is NULL", i);
#ifndef VLB_PROJ_TIMING
    // A call repeated with the same arguments (re-projecting the pushed skyboxes, src/skybox_manager.cpp:284-298) is
    // replayed from a CUDA graph: first sighting runs the launches directly, the second captures them, later ones replay.
    if (env_proj_graph() && n_maps >= 2 && n_maps <= 64) {
        vlb_ctx::ProjGraph* g = nullptr;
        for (vlb_ctx::ProjGraph& e : ctx->proj_graphs)
            if (e.fmt == format && e.W == width && e.H == height && e.order == sh_order && e.out == d_out && e.maps.size() == n_maps &&
                std::equal(e.maps.begin(), e.maps.end(), d_maps)) { g = &e; break; }
        if (g && g->exec && (g->partials != ctx->d_proj_partials.p || g->row_tab != ctx->d_row_tab.p || ctx->tab_w != width || ctx->tab_h != height || ctx->tab_variant != 0)) {
            cudaGraphExecDestroy(g->exec);      // a scratch buffer or the tables moved since the capture
            g->exec = nullptr; g->seen = 0;
        }
        if (g && g->exec) {
            VLB_CUDA(ctx, cudaGraphLaunch(g->exec, ctx->stream));
            ctx->launches += n_maps;
            return VLB_OK;
        }
        if (!g) {
            if (ctx->proj_graphs.size() >= 4) {
                if (ctx->proj_graphs.front().exec) cudaGraphExecDestroy(ctx->proj_graphs.front().exec);
                ctx->proj_graphs.erase(ctx->proj_graphs.begin());
            }
            ctx->proj_graphs.emplace_back();
            g = &ctx->proj_graphs.back();
            g->maps.assign(d_maps, d_maps + n_maps); g->fmt = format; g->W = width; g->H = height; g->order = sh_order; g->out = d_out;
        }
        if (++g->seen >= 2 && ctx->tab_w == width && ctx->tab_h == height && ctx->tab_variant == 0) {
            // tables, scratch and kernel attributes exist (the first sighting ran the same launches): capture
            if (!ctx->cap_stream) VLB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->cap_stream, cudaStreamNonBlocking));
            const uint64_t launches0 = ctx->launches;
            VLB_CUDA(ctx, cudaStreamBeginCapture(ctx->cap_stream, cudaStreamCaptureModeThreadLocal));
            const int rc = project_ptrs_lanes(ctx, ctx->cap_stream, d_maps, n_maps, format, width, height, sh_order, d_out);
            cudaGraph_t graph = nullptr;
            const cudaError_t ce = cudaStreamEndCapture(ctx->cap_stream, &graph);
            ctx->launches = launches0;
            if (rc == VLB_OK && ce == cudaSuccess && graph) {
                cudaGraphExec_t exec = nullptr;
                const cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
                cudaGraphDestroy(graph);
                if (ie == cudaSuccess) {
                    g->exec = exec; g->partials = ctx->d_proj_partials.p; g->row_tab = ctx->d_row_tab.p;
                    VLB_CUDA(ctx, cudaGraphLaunch(g->exec, ctx->stream));
                    ctx->launches += n_maps;
                    return VLB_OK;
                }
            } else if (graph) {
                cudaGraphDestroy(graph);
            }
            cudaGetLastError();            // capture not possible here: fall through to the direct launches
            g->seen = -(1 << 30);          // and do not try again for this argument list
        }
    }
#endif
    const int rc = project_ptrs_lanes(ctx, ctx->stream, d_maps, n_maps, format, width, height, sh_order, d_out);
#ifdef VLB_PROJ_TIMING
    if (rc == VLB_OK) return proj_timing_dump(ctx, (int)n_maps);
#endif
    return rc;
}

// ---- bake -------------------------------------------------------------------------------
void vlb_bake_settings_default(vlb_bake_settings* s) {
    if (!s) return;
    std::memset(s, 0, sizeof *s);
    s->probes[0] = s->probes[1] = s->probes[2] = 7;        // light_baker.cpp:38
    s->step[0] = s->step[1] = s->step[2] = 1.f;
    s->dir_w = 3141; s->dir_h = 1000;                       // light_baker.cpp:65, env_map_generator.cpp:24-28
    s->sh_order = 3;                                        // 16 coefficients, light_baker.cpp:294
    s->light_pos[0] = 1.f; s->light_pos[1] = 10.f; s->light_pos[2] = 1.f;   // env_map.rchit:25
    s->shadow_bias = 0.005f; s->c_diffuse = 0.5f; s->c_specular = 0.5f; s->gloss = 16.f; s->ambient = 0.f;
    s->tmin = 0.001f; s->tmax = 10000.f;                    // env_map.rgen:22-23
    s->flags = VLB_BAKE_SHADOW_RAYS | VLB_BAKE_SKYBOX_ON_MISS | VLB_BAKE_SRGB_ENCODE | VLB_BAKE_QUANTIZE_RGBA8;
    s->slab_k0 = 0; s->slab_k1 = -1;
    s->bounces = 0; s->indirect_gain = 1.f;                 // 0 bounces = the reference bake (direct light only)
}

int vlb_bake_settings_from_bounds(vlb_bake_settings* s, const float b[6]) {
    if (!s || !b) return VLB_ERR_INVALID;
    for (int d = 0; d < 3; ++d) {
        if (s->probes[d] < 1) return VLB_ERR_INVALID;
        s->origin[d] = b[d];
        // gridStep = (max - min) / (count - 1) (light_baker.cpp:85); a single probe divides by zero
        // in the reference, here the step is 0.
        s->step[d] = s->probes[d] > 1 ? (b[3 + d] - b[d]) / ((float)s->probes[d] - 1.f) : 0.f;
    }
    return VLB_OK;
}

int vlb_probe_positions(const vlb_bake_settings* s, float* out) {
    if (!s || !out || s->probes[0] < 1 || s->probes[1] < 1 || s->probes[2] < 1) return VLB_ERR_INVALID;
    const int Nx = s->probes[0], Ny = s->probes[1], Nz = s->probes[2];
    std::vector<float> px(Nx), py(Ny), pz(Nz);
    host_axis_coords(s->origin[0], s->step[0], Nx, px.data());
    host_axis_coords(s->origin[1], s->step[1], Ny, py.data());
    host_axis_coords(s->origin[2], s->step[2], Nz, pz.data());
    for (int k = 0; k < Nz; ++k) for (int j = 0; j < Ny; ++j) for (int i = 0; i < Nx; ++i) {
        const size_t idx = (s->flags & VLB_BAKE_REFERENCE_PROBE_ORDER) ? ref_order_index(i, j, k, Nx, Ny, Nz)
                                                                     : (size_t)i + (size_t)j * Nx + (size_t)k * Nx * Ny;
        out[3 * idx] = px[i]; out[3 * idx + 1] = py[j]; out[3 * idx + 2] = pz[k];
    }
    return VLB_OK;
}

// Validates the settings and returns the z-slices the call bakes: k0, k0 + stride, ... < k1.
static int slab_range(vlb_ctx* ctx, const vlb_bake_settings* s, int* k0, int* k1, int* stride) {
    if (!s) return ctx->fail(VLB_ERR_INVALID, "bake: settings is NULL");
    if (s->probes[0] < 1 || s->probes[1] < 1 || s->probes[2] < 1 || s->dir_w < 1 || s->dir_h < 1 ||
        (s->sh_order != 2 && s->sh_order != 3))
        return ctx->fail(VLB_ERR_INVALID, "bake: bad probe grid / direction grid / sh_order");
    *k0 = s->slab_k1 < 0 ? 0 : s->slab_k0;
    *k1 = s->slab_k1 < 0 ? s->probes[2] : s->slab_k1;
    *stride = s->slab_stride > 1 ? s->slab_stride : 1;
    if (*k0 < 0 || *k1 > s->probes[2] || *k0 > *k1) return ctx->fail(VLB_ERR_INVALID, "bake: bad slab range");
    if ((s->flags & (VLB_BAKE_REFERENCE_PROBE_ORDER | VLB_BAKE_ACCUMULATE_ACROSS_PROBES)) &&
        (*k0 != 0 || *k1 != s->probes[2] || *stride != 1))
        return ctx->fail(VLB_ERR_INVALID, "bake: reference probe order / accumulation need the whole grid");
    return VLB_OK;
}

int vlb_bake_gather_device(vlb_ctx* ctx, const vlb_bake_settings* s, const float* d_prev_full, float* d_out) {
    if (!ctx) return VLB_ERR_INVALID;
    if (int r = check_device(ctx)) return r;
    int k0, k1, stride;
    if (int r = slab_range(ctx, s, &k0, &k1, &stride)) return r;
    if (!d_out) return ctx->fail(VLB_ERR_INVALID, "bake: output is NULL");
    if (!ctx->have_scene) return ctx->fail(VLB_ERR_STATE, "bake: no scene set (vlb_scene_set_triangles)");
    if (!ctx->have_bvh) { if (int r = bvh_build(ctx, nullptr)) return r; }
    if ((s->flags & VLB_BAKE_SKYBOX_ON_MISS) && ctx->sky_w == 0)
        return ctx->fail(VLB_ERR_STATE, "bake: VLB_BAKE_SKYBOX_ON_MISS without a skybox (vlb_skybox_set)");
    if (d_prev_full && (s->flags & (VLB_BAKE_REFERENCE_PROBE_ORDER | VLB_BAKE_ACCUMULATE_ACROSS_PROBES)))
        return ctx->fail(VLB_ERR_INVALID, "bake: a gather pass reads the previous pass in x-fastest order; "
                                          "REFERENCE_PROBE_ORDER / ACCUMULATE_ACROSS_PROBES cannot be combined with it");
    if (d_prev_full && (reinterpret_cast<uintptr_t>(d_prev_full) & 15u))
        return ctx->fail(VLB_ERR_INVALID, "vlb_bake_gather_device: d_prev_full must be 16-byte aligned (probe records are read as float4)");
    return bake_device(ctx, s, d_prev_full, d_out);
}

int vlb_bake_probes_device(vlb_ctx* ctx, const vlb_bake_settings* s, float* d_out) {
    return vlb_bake_gather_device(ctx, s, nullptr, d_out);
}

int vlb_bake_probes(vlb_ctx* ctx, const vlb_bake_settings* s, float* out) {
    if (!ctx) return VLB_ERR_INVALID;
    if (int r = check_device(ctx)) return r;
    int k0, k1, stride;
    if (int r = slab_range(ctx, s, &k0, &k1, &stride)) return r;
    if (!out) return ctx->fail(VLB_ERR_INVALID, "bake: output is NULL");
    const size_t n = (size_t)s->probes[0] * s->probes[1] * (size_t)((k1 - k0 + stride - 1) / stride);
    if (n == 0) return VLB_OK;
    VLB_CUDA(ctx, ctx->d_bake_out.reserve(n * VLB_SH_STRIDE * sizeof(float)));
    if (s->bounces <= 0) {
        if (int r = vlb_bake_probes_device(ctx, s, ctx->d_bake_out.as<float>())) return r;
        VLB_CUDA(ctx, cudaMemcpyAsync(out, ctx->d_bake_out.p, n * VLB_SH_STRIDE * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
        // the copy was enqueued AFTER the bake recorded its completion event: wait for the stream itself, so that a
        // pinned `out` is complete when this blocking call returns (a pageable one is staged synchronously anyway)
        VLB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        vlb_bake_stats st;
        return vlb_bake_last_stats(ctx, &st);      // surfaces a traversal stack overflow
    }
    // Multi-bounce: pass 0 is the direct bake, pass b gathers from pass b-1 (two device buffers,
    // ping-pong). Every pass needs the previous one over the WHOLE grid, so a sharded grid has to be
    // driven by the caller: vlb_bake_gather_device per pass + an all-gather between passes.
    if (k0 != 0 || k1 != s->probes[2] || stride != 1)
        return ctx->fail(VLB_ERR_INVALID, "bake: bounces > 0 needs the whole grid in one call; for a sharded grid "
                                          "iterate vlb_bake_gather_device and all-gather the slabs between passes");
    VLB_CUDA(ctx, ctx->d_bake_prev.reserve(n * VLB_SH_STRIDE * sizeof(float)));
    float* buf[2] = {ctx->d_bake_out.as<float>(), ctx->d_bake_prev.as<float>()};
    vlb_bake_stats total{};
    int cur = 0;
    for (int pass = 0; pass <= s->bounces; ++pass) {
        if (int r = vlb_bake_gather_device(ctx, s, pass ? buf[cur ^ 1] : nullptr, buf[cur])) return r;
        vlb_bake_stats st;
        if (int r = vlb_bake_last_stats(ctx, &st)) return r;
        total.n_probes = st.n_probes; total.n_primary_rays += st.n_primary_rays; total.n_shadow_rays += st.n_shadow_rays;
        total.n_nodes_visited += st.n_nodes_visited; total.n_tris_tested += st.n_tris_tested;
        total.kernel_ms += st.kernel_ms; total.total_ms += st.total_ms;
        cur ^= 1;
    }
    ctx->last_bake = total;
    VLB_CUDA(ctx, cudaMemcpyAsync(out, buf[cur ^ 1], n * VLB_SH_STRIDE * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    VLB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return VLB_OK;
}

// One rank of vlb_bake_probes_multi: bakes the cyclic share of ctx `r` for one pass and copies it into its rows of
// the host grid (one strided copy: slice k = r + i * n goes to out + k * nxy * 192 bytes).
static int multi_rank_pass(vlb_ctx* ctx, const vlb_bake_settings* s, uint32_t r, uint32_t n, const float* prev_host, float* out) {
    if (int rc = check_device(ctx)) return rc;
    const int Nz = s->probes[2];
    const size_t nxy = (size_t)s->probes[0] * s->probes[1];
    const size_t n_slices = (size_t)(Nz > (int)r ? (Nz - (int)r + (int)n - 1) / (int)n : 0);
    if (n_slices == 0) return VLB_OK;
    vlb_bake_settings mine = *s;
    mine.slab_k0 = (int32_t)r; mine.slab_k1 = Nz; mine.slab_stride = (int32_t)n; mine.bounces = 0;
    VLB_CUDA(ctx, ctx->d_bake_out.reserve(n_slices * nxy * VLB_SH_STRIDE * sizeof(float)));
    const float* d_prev = nullptr;
    if (prev_host) {       // gather pass: every device needs the whole previous grid
        const size_t bytes = nxy * (size_t)Nz * VLB_SH_STRIDE * sizeof(float);
        VLB_CUDA(ctx, ctx->d_bake_prev.reserve(bytes));
        VLB_CUDA(ctx, cudaMemcpyAsync(ctx->d_bake_prev.p, prev_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
        d_prev = ctx->d_bake_prev.as<float>();
    }
    if (int rc = vlb_bake_gather_device(ctx, &mine, d_prev, ctx->d_bake_out.as<float>())) return rc;
    const size_t row = nxy * VLB_SH_STRIDE * sizeof(float);
    VLB_CUDA(ctx, cudaMemcpy2DAsync(reinterpret_cast<char*>(out) + (size_t)r * row, (size_t)n * row, ctx->d_bake_out.p, row, row, n_slices,
                                    cudaMemcpyDeviceToHost, ctx->stream));
    VLB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));    // the strided copy follows the bake's completion event: wait for it too
    vlb_bake_stats st;
    return vlb_bake_last_stats(ctx, &st);      // surfaces a traversal stack overflow
}

int vlb_bake_probes_multi(vlb_ctx* const* ctxs, uint32_t n_ctx, const vlb_bake_settings* s, float* out) {
    if (!ctxs || n_ctx == 0 || !ctxs[0]) return VLB_ERR_INVALID;
    vlb_ctx* c0 = ctxs[0];
    if (!s || !out) return c0->fail(VLB_ERR_INVALID, "vlb_bake_probes_multi: NULL settings or output");
    for (uint32_t r = 0; r < n_ctx; ++r) {
        if (!ctxs[r]) return c0->fail(VLB_ERR_INVALID, "vlb_bake_probes_multi: ctx %u is NULL", r);
        for (uint32_t q = 0; q < r; ++q) if (ctxs[q] == ctxs[r]) return c0->fail(VLB_ERR_INVALID, "vlb_bake_probes_multi: ctx %u given twice", r);
    }
    if (s->slab_k1 >= 0) return c0->fail(VLB_ERR_INVALID, "vlb_bake_probes_multi: takes the whole grid (slab_k1 < 0) and shards it itself");
    if (s->probes[0] <= 0 || s->probes[1] <= 0 || s->probes[2] <= 0) return c0->fail(VLB_ERR_INVALID, "bake: probe counts must be positive");
    // Distinct devices: one NCCL rank per ctx (vlb_comm_init_all, done here on first use), the shares are all-gathered
    // on the devices (NVLink) and between gather passes nothing touches the host; ctx 0 copies the grid out once.
    // Several ctxs on ONE device (NCCL allows one rank per GPU), or no NCCL in the process: the host-side path below.
    bool distinct = n_ctx > 1;
    for (uint32_t r = 0; r < n_ctx; ++r) for (uint32_t q = 0; q < r; ++q) if (ctxs[q]->device == ctxs[r]->device) distinct = false;
    bool have_comm = distinct;
    for (uint32_t r = 0; r < n_ctx && have_comm; ++r) have_comm = ctxs[r]->comm && ctxs[r]->comm_world == (int)n_ctx && ctxs[r]->comm_rank == (int)r;
    if (distinct && !have_comm) {
        bool none = true;
        for (uint32_t r = 0; r < n_ctx; ++r) none = none && !ctxs[r]->comm;
        have_comm = none && vlb_comm_init_all(ctxs, n_ctx) == VLB_OK;
    }
    if (have_comm) {
        std::vector<int> rc(n_ctx, VLB_OK);
        std::vector<std::thread> workers;
        // every ctx copies the slices it baked into its rows of `out`, each over its own PCIe link
        for (uint32_t r = 1; r < n_ctx; ++r) workers.emplace_back([&, r] { rc[r] = vlb_bake_probes_sharded_rows(ctxs[r], s, out); });
        rc[0] = vlb_bake_probes_sharded_rows(c0, s, out);
        for (std::thread& t : workers) t.join();
        for (uint32_t r = 0; r < n_ctx; ++r)
            if (rc[r] != VLB_OK) {
                if (r != 0) c0->fail(rc[r], "vlb_bake_probes_multi: ctx %u: %s", r, vlb_last_error(ctxs[r]));
                return rc[r];
            }
        return VLB_OK;
    }
    const size_t grid_floats = (size_t)s->probes[0] * s->probes[1] * (size_t)s->probes[2] * VLB_SH_STRIDE;
    std::vector<float> prev;                                   // previous pass over the whole grid (host)
    for (int pass = 0; pass <= std::max(0, (int)s->bounces); ++pass) {
        if (pass > 0) prev.assign(out, out + grid_floats);
        std::vector<int> rc(n_ctx, VLB_OK);
        std::vector<std::thread> workers;
        for (uint32_t r = 1; r < n_ctx; ++r)
            workers.emplace_back([&, r] { rc[r] = multi_rank_pass(ctxs[r], s, r, n_ctx, pass ? prev.data() : nullptr, out); });
        rc[0] = multi_rank_pass(c0, s, 0, n_ctx, pass ? prev.data() : nullptr, out);
        for (std::thread& t : workers) t.join();
        for (uint32_t r = 0; r < n_ctx; ++r)
            if (rc[r] != VLB_OK) {
                if (r != 0) c0->fail(rc[r], "vlb_bake_probes_multi: ctx %u: %s", r, vlb_last_error(ctxs[r]));
                return rc[r];
            }
    }
    return VLB_OK;
}

int vlb_bake_last_stats(vlb_ctx* ctx, vlb_bake_stats* out) {
    if (!ctx || !out) return VLB_ERR_INVALID;
    if (int r = check_device(ctx)) return r;
    if (int r = bake_collect_stats(ctx)) return r;
    *out = ctx->last_bake;
    return VLB_OK;
}

int vlb_trace_rays(vlb_ctx* ctx, const float* origins, const float* dirs, uint64_t n, float tmin, float tmax,
                   int accel, int kind, int32_t* hit_ids, float* hit_tuv) {
    if (!ctx) return VLB_ERR_INVALID;
    if (int r = check_device(ctx)) return r;
    if (!origins || !dirs || !hit_ids) return ctx->fail(VLB_ERR_INVALID, "vlb_trace_rays: NULL argument");
    if (!ctx->have_scene) return ctx->fail(VLB_ERR_STATE, "vlb_trace_rays: no scene set");
    if (accel == VLB_TRACE_BVH && !ctx->have_bvh) { if (int r = bvh_build(ctx, nullptr)) return r; }
    if (n == 0) return VLB_OK;
    return trace_rays(ctx, origins, dirs, n, tmin, tmax, accel, kind, hit_ids, hit_tuv);
}

}  // extern "C"

namespace vlb {
int sky_upload_join(vlb_ctx* ctx) { return join_sky_upload(ctx, ctx->stream); }
}  // namespace vlb

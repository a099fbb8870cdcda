// comm.cu — the multi-GPU side of the C ABI (include/vlb_bake.h, "multi-GPU" section).
//
// The reference is single-device (src/application.cpp:90-136 always picks physical device 0); BASELINE.json's
// north star shards the probe grid over the GPUs of one box: "probes shard across the 8 GPUs ... with the scene
// and BVH replicated, and the per-GPU SH slabs are gathered with one NCCL allgather over NVLink". This file is
// that exchange step behind the ABI, so that a C++ host (the reference's language, src/baker/main.cpp) reaches
// the same path bench.py measures:
//   * a communicator per ctx (one rank per GPU: one process per GPU, or several ctxs of one process),
//   * vlb_bake_probes_sharded[_device]: bake of this rank's cyclic z-slices, ONE ncclAllGather of the padded
//     shares, one un-interleave kernel into the consumer layout (x-fastest full grid on every rank); the
//     multi-bounce passes iterate on the device with that all-gather between passes,
//   * replicated uploads: every rank holds the same host arrays (scene, skybox) but copies only its 1/N-th over
//     PCIe; the rest arrives over NVLink with an in-place all-gather. Eight ranks of one box then move the scene
//     once over the host's PCIe root instead of eight times.
// NCCL is resolved at run time (dlopen of libnccl.so.2: the copy already loaded into the process, e.g. by torch,
// else the system one), so the library still loads on a machine without NCCL; every entry point below then fails
// with VLB_ERR_UNSUPPORTED and says why.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <mutex>

#include "vlb_context.h"

namespace vlb {

struct NcclApi {
    void* handle = nullptr;
    std::string why;
    decltype(&ncclGetVersion) GetVersion = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    bool ok() const { return handle != nullptr; }
};

static const NcclApi& nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        void* h = nullptr;
        for (const char* n : names) if (!h) h = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);   // the copy the process already has
        for (const char* n : names) if (!h) h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (!h) { const char* e = dlerror(); api.why = std::string("NCCL not found (dlopen libnccl.so.2: ") + (e ? e : "?") + ")"; return; }
#define VLB_SYM(field, name)                                                         \
        api.field = reinterpret_cast<decltype(api.field)>(dlsym(h, name));             \
        if (!api.field) { api.why = std::string("NCCL symbol missing: ") + name; return; }
        VLB_SYM(GetVersion, "ncclGetVersion") VLB_SYM(GetUniqueId, "ncclGetUniqueId") VLB_SYM(CommInitRank, "ncclCommInitRank")
        VLB_SYM(CommDestroy, "ncclCommDestroy") VLB_SYM(AllGather, "ncclAllGather") VLB_SYM(GroupStart, "ncclGroupStart")
        VLB_SYM(GroupEnd, "ncclGroupEnd") VLB_SYM(GetErrorString, "ncclGetErrorString")
#undef VLB_SYM
        api.handle = h;
    });
    return api;
}

#define VLB_NCCL(ctx, expr)                                                                                    \
    do {                                                                                                       \
        ncclResult_t _r = (expr);                                                                              \
        if (_r != ncclSuccess)                                                                                 \
            return (ctx)->fail(VLB_ERR_CUDA, "%s:%d: %s -> NCCL: %s", __FILE__, __LINE__, #expr, nccl_api().GetErrorString(_r)); \
    } while (0)

static inline ncclComm_t comm_of(const vlb_ctx* ctx) { return static_cast<ncclComm_t>(ctx->comm); }

// bytes of one rank's part when `bytes` are dealt to comm_world ranks in equal 256-byte-aligned parts
size_t comm_part_bytes(const vlb_ctx* ctx, size_t bytes) {
    const size_t w = (size_t)std::max(1, ctx->comm_world);
    return ((bytes + w - 1) / w + 255) & ~size_t(255);
}
size_t comm_padded_bytes(const vlb_ctx* ctx, size_t bytes) {
    return ctx->comm && ctx->comm_sharded_uploads ? comm_part_bytes(ctx, bytes) * (size_t)ctx->comm_world : bytes;
}

// Host array -> device buffer, identical on every rank afterwards. Plain copy without a communicator (or with
// sharded uploads off); otherwise rank r copies only part r and ONE in-place all-gather replicates the parts.
// d_dst must have room for comm_padded_bytes(bytes).
int upload_replicated(vlb_ctx* ctx, void* d_dst, const void* h_src, size_t bytes, cudaStream_t st) {
    if (bytes == 0) return VLB_OK;
    if (!ctx->comm || !ctx->comm_sharded_uploads || ctx->comm_world == 1) {
        VLB_CUDA(ctx, cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, st));
        return VLB_OK;
    }
    const size_t part = comm_part_bytes(ctx, bytes), off = part * (size_t)ctx->comm_rank;
    if (off < bytes)
        VLB_CUDA(ctx, cudaMemcpyAsync(static_cast<char*>(d_dst) + off, static_cast<const char*>(h_src) + off, std::min(part, bytes - off),
                                      cudaMemcpyHostToDevice, st));
    VLB_NCCL(ctx, nccl_api().AllGather(static_cast<char*>(d_dst) + off, d_dst, part, ncclChar, comm_of(ctx), st));
    return VLB_OK;
}

// stage[r][i][.] (rank r's i-th slice, shares padded to max_slices) -> full[k = r + i * world][.]; float4 granularity
__global__ void k_uninterleave(const float4* __restrict__ stage, float4* __restrict__ full, uint32_t slice_quads, uint32_t nz,
                               uint32_t world, uint32_t max_slices) {
    const size_t n = (size_t)slice_quads * nz;
    for (size_t f = (size_t)blockIdx.x * blockDim.x + threadIdx.x; f < n; f += (size_t)gridDim.x * blockDim.x) {
        const uint32_t k = (uint32_t)(f / slice_quads), within = (uint32_t)(f % slice_quads);
        const uint32_t r = k % world, i = k / world;
        full[f] = stage[((size_t)r * max_slices + i) * slice_quads + within];
    }
}

static int check_whole_grid(vlb_ctx* ctx, const vlb_bake_settings* s, const char* who) {
    if (!s) return ctx->fail(VLB_ERR_INVALID, "%s: settings is NULL", who);
    if (s->slab_k1 >= 0) return ctx->fail(VLB_ERR_INVALID, "%s: takes the whole grid (slab_k1 < 0) and shards it itself", who);
    if (s->probes[0] < 1 || s->probes[1] < 1 || s->probes[2] < 1) return ctx->fail(VLB_ERR_INVALID, "%s: probe counts must be positive", who);
    if (s->flags & (VLB_BAKE_REFERENCE_PROBE_ORDER | VLB_BAKE_ACCUMULATE_ACROSS_PROBES))
        return ctx->fail(VLB_ERR_INVALID, "%s: reference probe order / accumulation need the whole grid on one device", who);
    return VLB_OK;
}

// One pass of the sharded bake: this rank's cyclic share, all-gather, un-interleave into d_full (whole grid).
static int sharded_pass(vlb_ctx* ctx, const vlb_bake_settings* s, const float* d_prev_full, float* d_full) {
    const int world = ctx->comm ? ctx->comm_world : 1, rank = ctx->comm ? ctx->comm_rank : 0;
    vlb_bake_settings mine = *s;
    mine.bounces = 0;
    if (world == 1) return vlb_bake_gather_device(ctx, &mine, d_prev_full, d_full);
    const int Nz = s->probes[2];
    const size_t slice_floats = (size_t)s->probes[0] * s->probes[1] * VLB_SH_STRIDE;
    const size_t max_slices = ((size_t)Nz + world - 1) / world;
    mine.slab_k0 = rank; mine.slab_k1 = Nz; mine.slab_stride = world;
    const size_t share_bytes = max_slices * slice_floats * sizeof(float);
    VLB_CUDA(ctx, ctx->d_share.reserve(share_bytes));
    VLB_CUDA(ctx, ctx->d_gather_stage.reserve(share_bytes * world));
    if (rank < Nz) {
        if (int r = vlb_bake_gather_device(ctx, &mine, d_prev_full, ctx->d_share.as<float>())) return r;
    }
    cudaStream_t st = ctx->stream;
    VLB_NCCL(ctx, nccl_api().AllGather(ctx->d_share.p, ctx->d_gather_stage.p, max_slices * slice_floats, ncclFloat, comm_of(ctx), st));
    const size_t quads = slice_floats / 4 * (size_t)Nz;
    const unsigned grid = (unsigned)std::min<size_t>((quads + 255) / 256, (size_t)ctx->sm_count * 8);
    k_uninterleave<<<grid, 256, 0, st>>>(ctx->d_gather_stage.as<float4>(), reinterpret_cast<float4*>(d_full), (uint32_t)(slice_floats / 4),
                                         (uint32_t)Nz, (uint32_t)world, (uint32_t)max_slices);
    VLB_LAUNCH_CHECK(ctx);
    return VLB_OK;
}

int comm_destroy(vlb_ctx* ctx) {
    if (ctx->comm) {
        nccl_api().CommDestroy(comm_of(ctx));
        ctx->comm = nullptr;
    }
    ctx->comm_rank = 0; ctx->comm_world = 1; ctx->comm_sharded_uploads = false;
    return VLB_OK;
}

}  // namespace vlb

using namespace vlb;

extern "C" {

int vlb_comm_get_unique_id(void* id_out, uint64_t capacity_bytes) {
    const NcclApi& a = nccl_api();
    if (!a.ok()) { set_thread_error(a.why.c_str()); return VLB_ERR_UNSUPPORTED; }
    if (!id_out || capacity_bytes < VLB_COMM_ID_BYTES) { set_thread_error("vlb_comm_get_unique_id: buffer too small"); return VLB_ERR_INVALID; }
    static_assert(VLB_COMM_ID_BYTES == sizeof(ncclUniqueId), "VLB_COMM_ID_BYTES must match ncclUniqueId");
    ncclUniqueId id;
    const ncclResult_t r = a.GetUniqueId(&id);
    if (r != ncclSuccess) { set_thread_error(a.GetErrorString(r)); return VLB_ERR_CUDA; }
    memcpy(id_out, &id, sizeof id);
    return VLB_OK;
}

int vlb_comm_init_rank(vlb_ctx* ctx, const void* id, int rank, int world) {
    if (!ctx) return VLB_ERR_INVALID;
    const NcclApi& a = nccl_api();
    if (!a.ok()) return ctx->fail(VLB_ERR_UNSUPPORTED, "vlb_comm_init_rank: %s", a.why.c_str());
    if (!id || world < 1 || rank < 0 || rank >= world) return ctx->fail(VLB_ERR_INVALID, "vlb_comm_init_rank: bad rank / world / id");
    if (ctx->comm) return ctx->fail(VLB_ERR_STATE, "vlb_comm_init_rank: this ctx already has a communicator");
    VLB_CUDA(ctx, cudaSetDevice(ctx->device));
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof uid);
    ncclComm_t c = nullptr;
    VLB_NCCL(ctx, a.CommInitRank(&c, world, uid, rank));
    ctx->comm = c; ctx->comm_rank = rank; ctx->comm_world = world;
    return VLB_OK;
}

int vlb_comm_init_all(vlb_ctx* const* ctxs, uint32_t n) {
    if (!ctxs || n == 0 || !ctxs[0]) return VLB_ERR_INVALID;
    vlb_ctx* c0 = ctxs[0];
    const NcclApi& a = nccl_api();
    if (!a.ok()) return c0->fail(VLB_ERR_UNSUPPORTED, "vlb_comm_init_all: %s", a.why.c_str());
    for (uint32_t r = 0; r < n; ++r) {
        if (!ctxs[r]) return c0->fail(VLB_ERR_INVALID, "vlb_comm_init_all: ctx %u is NULL", r);
        if (ctxs[r]->comm) return c0->fail(VLB_ERR_STATE, "vlb_comm_init_all: ctx %u already has a communicator", r);
        for (uint32_t q = 0; q < r; ++q)
            if (ctxs[q]->device == ctxs[r]->device) return c0->fail(VLB_ERR_INVALID, "vlb_comm_init_all: ctx %u and %u share device %d (one rank per GPU)", q, r, ctxs[r]->device);
    }
    ncclUniqueId uid;
    VLB_NCCL(c0, a.GetUniqueId(&uid));
    std::vector<ncclComm_t> comms(n, nullptr);
    VLB_NCCL(c0, a.GroupStart());
    ncclResult_t first_bad = ncclSuccess;
    cudaError_t dev_bad = cudaSuccess;
    for (uint32_t r = 0; r < n && first_bad == ncclSuccess && dev_bad == cudaSuccess; ++r) {
        dev_bad = cudaSetDevice(ctxs[r]->device);
        if (dev_bad == cudaSuccess) first_bad = a.CommInitRank(&comms[r], (int)n, uid, (int)r);
    }
    const ncclResult_t end = a.GroupEnd();       // always closed, also after a failed rank, so no group is left open
    if (dev_bad != cudaSuccess) return c0->fail(VLB_ERR_CUDA, "vlb_comm_init_all: cudaSetDevice -> %s", cudaGetErrorString(dev_bad));
    if (first_bad != ncclSuccess || end != ncclSuccess) {
        for (uint32_t r = 0; r < n; ++r) if (comms[r]) a.CommDestroy(comms[r]);
        return c0->fail(VLB_ERR_CUDA, "vlb_comm_init_all: NCCL: %s", a.GetErrorString(first_bad != ncclSuccess ? first_bad : end));
    }
    for (uint32_t r = 0; r < n; ++r) { ctxs[r]->comm = comms[r]; ctxs[r]->comm_rank = (int)r; ctxs[r]->comm_world = (int)n; }
    return VLB_OK;
}

int vlb_comm_destroy(vlb_ctx* ctx) {
    if (!ctx) return VLB_ERR_INVALID;
    cudaSetDevice(ctx->device);
    return comm_destroy(ctx);
}

int vlb_comm_info(const vlb_ctx* ctx, int32_t* rank, int32_t* world, int32_t* nccl_version) {
    if (!ctx) return VLB_ERR_INVALID;
    if (rank) *rank = ctx->comm ? ctx->comm_rank : 0;
    if (world) *world = ctx->comm ? ctx->comm_world : 1;
    if (nccl_version) {
        int v = 0;
        if (nccl_api().ok()) nccl_api().GetVersion(&v);
        *nccl_version = v;
    }
    return VLB_OK;
}

int vlb_comm_sharded_uploads(vlb_ctx* ctx, int enable) {
    if (!ctx) return VLB_ERR_INVALID;
    if (enable && !ctx->comm) return ctx->fail(VLB_ERR_STATE, "vlb_comm_sharded_uploads: no communicator (vlb_comm_init_rank)");
    ctx->comm_sharded_uploads = enable != 0;
    return VLB_OK;
}

int vlb_bake_probes_sharded_device(vlb_ctx* ctx, const vlb_bake_settings* s, const float* d_prev_full, float* d_full_out) {
    if (!ctx) return VLB_ERR_INVALID;
    VLB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (int r = check_whole_grid(ctx, s, "vlb_bake_probes_sharded_device")) return r;
    if (!d_full_out) return ctx->fail(VLB_ERR_INVALID, "vlb_bake_probes_sharded_device: output is NULL");
    if (reinterpret_cast<uintptr_t>(d_full_out) & 15u) return ctx->fail(VLB_ERR_INVALID, "vlb_bake_probes_sharded_device: d_full_out must be 16-byte aligned");
    return sharded_pass(ctx, s, d_prev_full, d_full_out);
}

static int bake_sharded_host(vlb_ctx* ctx, const vlb_bake_settings* s, float* out_or_null, bool own_rows_only);

int vlb_bake_probes_sharded(vlb_ctx* ctx, const vlb_bake_settings* s, float* out_or_null) {
    return bake_sharded_host(ctx, s, out_or_null, false);
}

int vlb_bake_probes_sharded_rows(vlb_ctx* ctx, const vlb_bake_settings* s, float* grid) {
    if (ctx && !grid) return ctx->fail(VLB_ERR_INVALID, "vlb_bake_probes_sharded_rows: grid is NULL");
    return bake_sharded_host(ctx, s, grid, true);
}

static int bake_sharded_host(vlb_ctx* ctx, const vlb_bake_settings* s, float* out_or_null, bool own_rows_only) {
    if (!ctx) return VLB_ERR_INVALID;
    VLB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (int r = check_whole_grid(ctx, s, "vlb_bake_probes_sharded")) return r;
    const size_t bytes = (size_t)s->probes[0] * s->probes[1] * (size_t)s->probes[2] * VLB_SH_STRIDE * sizeof(float);
    VLB_CUDA(ctx, ctx->d_bake_out.reserve(bytes));
    if (s->bounces > 0) VLB_CUDA(ctx, ctx->d_bake_prev.reserve(bytes));
    float* buf[2] = {ctx->d_bake_out.as<float>(), ctx->d_bake_prev.as<float>()};
    vlb_bake_stats total{};
    int cur = 0;
    for (int pass = 0; pass <= std::max(0, (int)s->bounces); ++pass) {
        if (int r = sharded_pass(ctx, s, pass ? buf[cur ^ 1] : nullptr, buf[cur])) return r;
        if (s->bounces > 0) {          // per-pass statistics add up (this synchronises with the pass, as vlb_bake_probes does)
            vlb_bake_stats st;
            if (int r = vlb_bake_last_stats(ctx, &st)) return r;
            total.n_probes = st.n_probes; total.n_primary_rays += st.n_primary_rays; total.n_shadow_rays += st.n_shadow_rays;
            total.n_nodes_visited += st.n_nodes_visited; total.n_tris_tested += st.n_tris_tested;
            total.kernel_ms += st.kernel_ms; total.total_ms += st.total_ms;
        }
        cur ^= 1;
    }
    const int world = ctx->comm ? ctx->comm_world : 1, rank = ctx->comm ? ctx->comm_rank : 0;
    if (out_or_null && own_rows_only && world > 1) {
        // only the z-slices this rank baked (k = rank, rank + world, ...): one strided copy into the caller's grid
        const int Nz = s->probes[2];
        const size_t row = (size_t)s->probes[0] * s->probes[1] * VLB_SH_STRIDE * sizeof(float);
        const size_t n_slices = Nz > rank ? (size_t)(Nz - rank + world - 1) / world : 0;
        if (n_slices)
            VLB_CUDA(ctx, cudaMemcpy2DAsync(reinterpret_cast<char*>(out_or_null) + (size_t)rank * row, (size_t)world * row,
                                            reinterpret_cast<const char*>(buf[cur ^ 1]) + (size_t)rank * row, (size_t)world * row, row, n_slices,
                                            cudaMemcpyDeviceToHost, ctx->stream));
    } else if (out_or_null) {
        VLB_CUDA(ctx, cudaMemcpyAsync(out_or_null, buf[cur ^ 1], bytes, cudaMemcpyDeviceToHost, ctx->stream));
    }
    VLB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (s->bounces > 0) { ctx->last_bake = total; return VLB_OK; }
    vlb_bake_stats st;
    return vlb_bake_last_stats(ctx, &st);      // surfaces a traversal stack overflow
}

}  // extern "C"

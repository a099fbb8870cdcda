/*
 * vlb_bake.h — C ABI of the B200-native light-probe baker (libvlb_bake.so).
 *
 * This is the drop-in boundary for the bake path of Reefufui/vulkan-light-bakery.
 * The reference has no FFI layer of its own (its entry points are C++ classes), so every
 * function below cites the reference interface it replaces (paths relative to the reference
 * repository root). Plain pointers and sizes only; no C++/torch/CUDA types cross this boundary.
 *
 * Conventions
 *   - every call returns a vlb_status (0 = ok, <0 = error); the text of the last error is
 *     available from vlb_last_error(). Nothing throws or aborts across the ABI.
 *   - the caller owns every host buffer passed in or out; the library owns all device memory
 *     inside a vlb_ctx. A ctx is bound to ONE CUDA device; it is not thread-safe, distinct
 *     contexts are independent (one ctx per GPU / per rank).
 *   - functions without a `_device` suffix take HOST pointers and are synchronous (they return
 *     after the ctx stream has drained), mirroring the reference's blocking submits
 *     (src/application.cpp:255-277). `_device` variants take DEVICE pointers, enqueue on the
 *     ctx stream and return without synchronising.
 *   - there is no CPU fallback: every compute entry point fails with VLB_ERR_NO_DEVICE when no
 *     CUDA device is usable.
 */
#ifndef VLB_BAKE_H
#define VLB_BAKE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VLB_ABI_VERSION 1

typedef struct vlb_ctx vlb_ctx;

typedef enum vlb_status {
    VLB_OK              =  0,
    VLB_ERR_INVALID     = -1,  /* bad argument                                   */
    VLB_ERR_CUDA        = -2,  /* CUDA runtime error (text in vlb_last_error)     */
    VLB_ERR_NO_DEVICE   = -3,  /* no usable CUDA device                          */
    VLB_ERR_STATE       = -4,  /* call order violated (e.g. bake before scene)   */
    VLB_ERR_IO          = -5,  /* file could not be read / written               */
    VLB_ERR_UNSUPPORTED = -6,  /* valid glTF / format feature that is not handled*/
    VLB_ERR_NOMEM       = -7
} vlb_status;

/* ------------------------------------------------------------------------------------------
 * Layouts shared with the reference's shaders (shaders/structures.h:13-71, scalar layout).
 * They are accepted verbatim, so a maintainer can pass the very std::vectors that
 * Scene_t::fetchVertices / loadMaterials build (src/scene_manager.cpp:257-290, 695-857).
 * ---------------------------------------------------------------------------------------- */
typedef struct vlb_vertex {      /* shader::Vertex, 44 bytes (structures.h:20-26) */
    float position[4];
    float normal[3];
    float uv0[2];
    float uv1[2];
} vlb_vertex;

typedef struct vlb_texture_ref { /* shader::Texture (structures.h:46-50) */
    int32_t index;               /* -1 = none */
    int32_t coord_set;
} vlb_texture_ref;

typedef struct vlb_material {    /* shader::Material, 144 bytes (structures.h:28-71) */
    /* Textures (64 B) */
    vlb_texture_ref normal, occlusion, base_color, metallic_roughness, emissive,
                    diffuse_ext, specular_ext, dummy;
    /* Factors (80 B) */
    float base_color_factor[4];
    float emissive_factor[4];
    float diffuse_factor[4];
    float specular_factor[3];
    float metallic;
    float roughness;
    float alpha_cutoff;
    float pad[2];
} vlb_material;

/* One reference "instance" = one glTF primitive placed by its node's world matrix
 * (src/scene_manager.cpp:385-443: one BLAS per primitive, instance transform =
 * Node_t::getMatrix() :445-461, customIndex = running instance id). Indices are relative to
 * `first_vertex`, as each reference primitive owns its own vertex buffer. */
typedef struct vlb_instance {
    uint32_t first_index;        /* offset into the index array                              */
    uint32_t index_count;        /* triangles = index_count / 3 (scene_manager.cpp:236)      */
    uint32_t first_vertex;       /* offset into the vertex array                             */
    uint32_t vertex_count;
    uint32_t material_index;
    float    transform[12];      /* object->world, 3x4 row-major (VkTransformMatrixKHR)      */
} vlb_instance;

/* One glTF texture = image + sampler, as Scene_t::loadTextures / loadSamplers create them
 * (src/scene_manager.cpp:941-973, 650-690): RGBA8 texels (tinygltf always expands to 4 components; unorm,
 * no sRGB decode: src/application.cpp:779-780), wrap mode per axis, filter. The hit shader samples
 * `texture(textures[material.textures.baseColor.index], uv0)` (shaders/env_map.rchit:36-49): in a ray-tracing
 * stage that is the base mip level, so only level 0 is kept here. Bilinear weights are exact fp32 (the
 * Vulkan driver's fixed-point weights are "parity unpinned"). */
enum { VLB_WRAP_REPEAT = 0, VLB_WRAP_CLAMP_TO_EDGE = 1, VLB_WRAP_MIRRORED_REPEAT = 2 };
enum { VLB_FILTER_LINEAR = 0, VLB_FILTER_NEAREST = 1 };
typedef struct vlb_texture {
    const void* texels;          /* width * height * 4 bytes, row 0 first                      */
    int32_t width, height;
    int32_t wrap_u, wrap_v;      /* VLB_WRAP_*  (default sampler: repeat, application.hpp:45-52) */
    int32_t filter;              /* VLB_FILTER_* (default sampler: linear)                      */
    int32_t reserved;
} vlb_texture;

typedef enum vlb_texel_format {
    VLB_FMT_RGBA8   = 0,         /* reference skyboxes: stb RGBA8, value/255, no sRGB decode
                                    (src/skybox_manager.cpp:18, src/application.cpp:690,779)  */
    VLB_FMT_RGBA32F = 1          /* BASELINE configs: float4 texels                           */
} vlb_texel_format;

/* ------------------------------------------------------------------------------------------
 * Bake settings. Defaults (vlb_bake_settings_default) are the reference's hard-coded
 * constants: 7x7x7 probes (src/baker/light_baker.cpp:38), 3141x1000 directions
 * (light_baker.cpp:65, src/baker/env_map_generator.cpp:24-28), 16 coefficients
 * (light_baker.cpp:294), light (1,10,1), bias .005, Cdiffuse/Cspecular .5, gloss 16,
 * ambient 0 (shaders/env_map.rchit:25,75-96), tmin 1e-3 / tmax 1e4 (shaders/env_map.rgen:22-23).
 * ---------------------------------------------------------------------------------------- */
enum {
    VLB_BAKE_SHADOW_RAYS             = 1u << 0, /* env_map.rchit:83-88                        */
    VLB_BAKE_SKYBOX_ON_MISS          = 1u << 1, /* main.rmiss:18-40 (intent)  */
    VLB_BAKE_SRGB_ENCODE             = 1u << 2, /* env_map.rchit:101, main.rmiss:40           */
    VLB_BAKE_QUANTIZE_RGBA8          = 1u << 3, /* rgba8 storage image, env_map_generator.hpp:39 */
    VLB_BAKE_REFERENCE_PROBE_ORDER   = 1u << 4, /* writer order of light_baker.cpp:80-101     */
    VLB_BAKE_ACCUMULATE_ACROSS_PROBES= 1u << 5, /* literal light_baker.cpp:110-121 behaviour  */
    VLB_BAKE_SH_WORLD_FRAME          = 1u << 6  /* evaluate SH on the world ray dir (App. B-6) */
};

typedef struct vlb_bake_settings {
    int32_t  probes[3];          /* probe grid counts Nx,Ny,Nz                                */
    float    origin[3];          /* position of probe (0,0,0) = bounds min                    */
    float    step[3];            /* gridStep = (max-min)/(N-1) (light_baker.cpp:85)           */
    int32_t  dir_w, dir_h;       /* equirect direction grid (env_map.rgen:20-21)              */
    int32_t  sh_order;           /* 2 -> 9 coefficients written, 3 -> 16                      */
    float    light_pos[3];
    float    shadow_bias;
    float    c_diffuse;
    float    c_specular;
    float    gloss;
    float    ambient;
    float    tmin, tmax;
    uint32_t flags;
    int32_t  slab_k0, slab_k1;   /* bake only z-slices k0 <= k < k1; k1 < 0 = whole grid      */
    int32_t  slab_stride;        /* > 1: only every slab_stride-th slice of [k0, k1), i.e.
                                    k = k0, k0 + stride, ... (cyclic sharding over GPUs: rank r of
                                    N bakes k0 = r, k1 = Nz, stride = N); 0 or 1 = contiguous     */
    /* Multi-bounce gather (BASELINE configs[3]). The reference bakes direct light only;
     * its run-time shader gathers indirect light from the baked probes (shaders/main.rchit:124-167,
     * shaders/sh.rmiss:27-36). bounces = B > 0 re-bakes the grid B more times, every hit adding
     * indirect_gain * (that gather operator applied to the previous pass). 0 = the reference bake. */
    int32_t  bounces;
    float    indirect_gain;      /* scale of the gathered term (main.rchit:165 uses ambient * 1250) */
    int32_t  reserved;
} vlb_bake_settings;

#define VLB_SH_STRIDE 48         /* floats per probe: vec3 coeffs[16] (shaders/sh.comp:21)     */

typedef struct vlb_bvh_stats {
    uint64_t n_triangles;
    uint64_t n_nodes;            /* traversal nodes emitted                                   */
    uint32_t max_leaf_size;
    uint32_t reserved;
    float    bounds[6];          /* tight world AABB min xyz, max xyz                         */
    float    build_ms;           /* device time of the whole build                            */
    float    sort_ms;
} vlb_bvh_stats;

typedef struct vlb_bake_stats {
    uint64_t n_probes;           /* probes baked by this call (slab)                          */
    uint64_t n_primary_rays;
    uint64_t n_shadow_rays;      /* shadow rays actually traced (sDotN != 0)                  */
    uint64_t n_nodes_visited;    /* filled only by the instrumented build      */
    uint64_t n_tris_tested;
    float    kernel_ms;          /* device time of the bake kernel(s) of the last call        */
    float    total_ms;           /* device time of the whole call                             */
} vlb_bake_stats;

/* --- context ---------------------------------------------------------------------------- */
/* Replaces vlb::Application's device bring-up (src/application.cpp:540-556). */
int  vlb_ctx_create(int device_id, vlb_ctx** out);
void vlb_ctx_destroy(vlb_ctx* ctx);
/* Use an existing CUDA stream (a cudaStream_t passed as an integer handle; 0 is CUDA's legacy
 * default stream, as everywhere in CUDA); VLB_STREAM_OWN restores the ctx's own stream. */
#define VLB_STREAM_OWN (~(uint64_t)0)
int  vlb_ctx_set_stream(vlb_ctx* ctx, uint64_t cuda_stream_handle);
/* The cudaStream_t the ctx currently enqueues on, as an integer handle (its own non-blocking stream unless
 * vlb_ctx_set_stream changed it): lets a caller record events on it or make other streams wait for it. */
uint64_t vlb_ctx_stream(const vlb_ctx* ctx);
int  vlb_ctx_synchronize(vlb_ctx* ctx);
/* Last error text of this ctx (or of the calling thread when ctx == NULL). Never NULL. */
const char* vlb_last_error(const vlb_ctx* ctx);
int  vlb_abi_version(void);
/* Names of the CUDA kernels this library launched since the ctx was created, and how many
 * launches in total (bench.py's gpu_launches). */
uint64_t vlb_ctx_launch_count(const vlb_ctx* ctx);

/* --- scene (replaces Scene_t ingest output + buildAccelerationStructures) --------------- */
/* SceneManager::pushScene -> Scene_t::loadNode output (src/scene_manager.cpp:479-538). */
int vlb_scene_set_triangles(vlb_ctx* ctx,
                            const vlb_vertex* vertices, uint64_t n_vertices,
                            const uint32_t* indices, uint64_t n_indices,
                            const vlb_instance* instances, uint32_t n_instances,
                            const vlb_material* materials, uint32_t n_materials);
/* Scene_t::loadTextures (src/scene_manager.cpp:941-973): the textures that vlb_material.base_color.index
 * refers to (index -1 = none: the factor path of env_map.rchit:43-47). May be called before or after
 * vlb_scene_set_triangles; replaces any previous set; n = 0 removes them. A bake fails with VLB_ERR_STATE
 * if a material names a texture that has not been set. */
int vlb_scene_set_textures(vlb_ctx* ctx, const vlb_texture* textures, uint32_t n_textures);
/* SceneManager::pushScene(std::string&) (src/scene_manager.cpp:1013-1034): tinygltf load of a .gltf /
 * .glb file (:32-67), materials (+ trailing default, :837-857), node hierarchy in loadNode's
 * pre-order with world matrices (:445-538), vertices as shader::Vertex and u32 indices
 * (:257-337); then the same upload as vlb_scene_set_triangles. vlb_scene_bounds(tight=0) afterwards
 * returns the reference's bounds including its local-matrix quirk (:497-507). Textures referenced by
 * baseColorTexture are decoded (PNG: 1-8 bit grey / palette, 8-bit RGB / RGBA, non-interlaced; JPEG:
 * 8-bit baseline or progressive, grey or YCbCr 4:4:4 / 4:2:2 / 4:4:0 / 4:2:0; anything else is
 * VLB_ERR_UNSUPPORTED) and set
 * with their samplers (:650-690) as by vlb_scene_set_textures. */
int vlb_scene_load_gltf(vlb_ctx* ctx, const char* gltf_path);
/* Host-only: parse a glTF and report counts = {vertices, indices, instances (node x primitive),
 * materials incl. the default, triangles} and the reference-mode bounds. No CUDA device needed. */
int vlb_gltf_probe(const char* gltf_path, uint64_t counts[5], float ref_bounds_min_max[6]);
/* Host-only: texture `index` of a glTF as vlb_scene_load_gltf would set it. info = {width, height, wrap_u,
 * wrap_v, filter, 1 if the texture is used as a baseColor texture (else it is a 1x1 white placeholder)};
 * texels (may be NULL to query the size first) receives width*height*4 RGBA8 bytes if capacity allows. */
int vlb_gltf_texture(const char* gltf_path, uint32_t index, void* texels, uint64_t capacity_bytes, int32_t info[6]);
/* Scene_t::getBounds (src/scene_manager.cpp:214-217). mode 0: the reference's semantics
 * (bounds start at the origin, only the two local AABB corners are transformed,
 * scene_manager.cpp:497-507); mode 1: tight world-space AABB of all triangles. */
int vlb_scene_bounds(vlb_ctx* ctx, int tight, float out_min_max[6]);
/* Scene_t::buildAccelerationStructures (src/scene_manager.cpp:385-443) -> software LBVH. */
int vlb_bvh_build(vlb_ctx* ctx, vlb_bvh_stats* stats_or_null);
/* The build preference the reference passes to its driver (VK_BUILD_ACCELERATION_STRUCTURE_PREFER_FAST_TRACE_BIT,
 * src/scene_manager.cpp:346-347): which builder makes the binary hierarchy under the wide nodes. LBVH (default):
 * Karras 2012, fastest build. PLOC: agglomerative clustering within `ploc_radius` (1..64, 0 = keep / default 16) places
 * of the Morton order, a few per cent fewer node visits per ray on irregular scenes for about twice the build time.
 * Results never depend on the builder (hit ids are bit-exact with any tree). Takes effect at the next vlb_bvh_build. */
enum { VLB_BVH_BUILDER_LBVH = 0, VLB_BVH_BUILDER_PLOC = 1 };
int vlb_bvh_set_builder(vlb_ctx* ctx, int builder, int ploc_radius);
/* Host-only: the builder that minimises build + trace time for a bake of n_primary_rays rays (on this GPU / rank) through
 * a scene of n_triangles, from the measured costs on B200 (PLOC: +2.5 ns per triangle of build time, -4 % of trace time at
 * ~0.3 ns per primary ray; no gain below a few thousand or beyond ~1 M triangles on the scenes measured): PLOC when the
 * trace is long enough to pay for it. The caller knows the probe grid before it builds, as the reference knows its build flags. */
int vlb_bvh_recommend_builder(uint64_t n_triangles, uint64_t n_primary_rays);

/* --- skybox (replaces Skybox_t) ---------------------------------------------------------- */
/* Host-only: decode an image file into RGBA8 texels, as the reference's stbi_load(file, &w, &h, &ch, STBI_rgb_alpha)
 * does for skyboxes (Skybox_t ctor, src/skybox_manager.cpp:14-20). PNG and baseline JPEG (see vlb_scene_load_gltf for
 * the exact subsets). size[2] = {width, height}; texels may be NULL to query the size first; it receives
 * width*height*4 bytes if capacity allows. Feed the result to vlb_skybox_set / vlb_skybox_project_sh (VLB_FMT_RGBA8). */
int vlb_image_load_rgba8(const char* path, void* texels, uint64_t capacity_bytes, int32_t size[2]);
/* Host-only: decode a Radiance RGBE (.hdr) file into linear RGBA32F texels (alpha 1): the usual container of the
 * RGBA32F equirect skyboxes the BASELINE configs use. Same calling convention as vlb_image_load_rgba8; texels
 * receives width*height*16 bytes. Feed the result to vlb_skybox_set / vlb_skybox_project_sh (VLB_FMT_RGBA32F). */
int vlb_image_load_rgba32f(const char* path, void* texels, uint64_t capacity_bytes, int32_t size[2]);
/* Skybox_t::createTexture (src/skybox_manager.cpp:49-63): the map sampled on ray miss. */
int vlb_skybox_set(vlb_ctx* ctx, const void* texels, int format, int width, int height);
/* Same, without blocking: the upload is enqueued on a copy stream of the ctx and the call returns at once; the
 * next bake (or vlb_ctx_synchronize / vlb_skybox_set) waits for it on the device. `texels` must stay valid and
 * unchanged until then, and should be pinned host memory for the copy to overlap the calls in between
 * (scene upload, LBVH build). */
int vlb_skybox_set_async(vlb_ctx* ctx, const void* texels, int format, int width, int height);
/* Skybox_t::createSHBuffer + computeSH (src/skybox_manager.cpp:65-76,107-130), i.e. one
 * dispatch of shaders/skybox_sh.comp. out = vec3 coeffs[16] (48 floats); order 2 fills
 * entries 0..8 and zero-fills 9..15. */
int vlb_skybox_project_sh(vlb_ctx* ctx, const void* texels, int format, int width, int height,
                          int sh_order, float* out48);
/* n maps of identical size; maps[i] are host pointers; out = n x 48 floats. */
int vlb_skybox_project_sh_batched(vlb_ctx* ctx, const void* const* maps, uint32_t n_maps,
                                  int format, int width, int height, int sh_order, float* out);
/* Device-resident variant: n maps at d_texels + i*map_stride_bytes, d_out = n x 48 floats. */
int vlb_skybox_project_sh_device(vlb_ctx* ctx, const void* d_texels, uint64_t map_stride_bytes,
                                 uint32_t n_maps, int format, int width, int height,
                                 int sh_order, float* d_out);
/* n independent maps of identical size at n DEVICE addresses (host array of device pointers),
 * d_out = n x 48 floats. One launch per map; the launches alternate over internal streams forked
 * from and joined to the ctx stream, so that one map streams from HBM while another is being
 * reduced; for the caller the call is ordered on the ctx stream like any other. Replaces a loop of Skybox_t::computeSH over pushed skyboxes
 * (src/skybox_manager.cpp:284-298 -> 107-130). */
int vlb_skybox_project_sh_device_ptrs(vlb_ctx* ctx, const void* const* d_maps, uint32_t n_maps,
                                      int format, int width, int height, int sh_order, float* d_out);
/* shaders/sh.comp applied to a caller-supplied environment image (the reference's image-input
 * debug path, src/baker/light_baker.cpp:68-73): SH argument is the un-swizzled toVector. */
int vlb_envmap_project_sh(vlb_ctx* ctx, const void* texels, int format, int width, int height,
                          int sh_order, float* out48);

/* --- bake (replaces LightBaker::bake's per-probe getMap + dispatchBakingKernel loop) ----- */
void vlb_bake_settings_default(vlb_bake_settings* s);
/* LightBaker::probePositionsFromBoudingBox (src/baker/light_baker.cpp:80-101): fills
 * origin/step from bounds and probes[]. */
int vlb_bake_settings_from_bounds(vlb_bake_settings* s, const float bounds_min_max[6]);
/* Probe positions in OUTPUT order (x-fastest, or the reference writer's order under
 * VLB_BAKE_REFERENCE_PROBE_ORDER); out = Nx*Ny*Nz*3 floats. Host-only helper. */
int vlb_probe_positions(const vlb_bake_settings* s, float* out_xyz);
/* LightBaker::bake (src/baker/light_baker.cpp:287-328). out = n_slab_probes x 48 floats
 * (float[probe][16][3], light_baker.cpp:294-322). With s->bounces > 0 the call must cover the whole
 * grid (no slab) and runs 1 + bounces passes on the device, returning the last one. */
int vlb_bake_probes(vlb_ctx* ctx, const vlb_bake_settings* s, float* out);
/* The same bake spread over several GPUs from ONE host process (how the reference's single-process `baker` would
 * use a multi-GPU box): ctxs[r] are n distinct contexts (normally one per device) that already hold the same
 * scene, LBVH and skybox; ctx r bakes z-slices r, r + n, ... on its own device and host thread, and every share
 * lands in the one host grid `out` (Nx*Ny*Nz x 48 floats). Contexts on distinct devices become the ranks of an NCCL
 * communicator (vlb_comm_init_all, done here on first use): the shares are all-gathered on the devices over NVLink,
 * also between the gather passes of s->bounces > 0, and ctx 0 copies the grid out (vlb_bake_probes_sharded per
 * ctx). Contexts that share a device, or a process without NCCL, take the host route instead: every share is copied
 * straight into its rows of `out` and the previous pass is redistributed from the host between passes. Either way
 * bit-identical to vlb_bake_probes on one ctx. Errors: the code of the first failing ctx, text in
 * vlb_last_error(ctxs[0]). */
int vlb_bake_probes_multi(vlb_ctx* const* ctxs, uint32_t n_ctx, const vlb_bake_settings* s, float* out);
/* --- multi-GPU: the exchange step of the sharded bake, behind the ABI ----------------------------------------
 * The reference is single-device (src/application.cpp:90-136 always takes physical device 0); BASELINE.json shards
 * the probe grid over the GPUs of one box (scene + LBVH + skybox replicated, z-slices dealt cyclically: rank r of N
 * bakes k = r, r + N, ...) and gathers the per-GPU SH shares with ONE NCCL all-gather over NVLink. One communicator
 * rank per ctx: one process per GPU (vlb_comm_get_unique_id on rank 0, the 128 bytes carried to the other ranks by
 * whatever the host program uses -- MPI, a file, torch.distributed --, then vlb_comm_init_rank everywhere) or several
 * ctxs of one process (vlb_comm_init_all). NCCL is resolved at run time (libnccl.so.2); without it these calls
 * return VLB_ERR_UNSUPPORTED. Collective calls (init, sharded bakes, and every upload while sharded uploads are on)
 * must be made by all ranks in the same order. */
#define VLB_COMM_ID_BYTES 128
int vlb_comm_get_unique_id(void* id_out, uint64_t capacity_bytes);
int vlb_comm_init_rank(vlb_ctx* ctx, const void* id, int rank, int world);
int vlb_comm_init_all(vlb_ctx* const* ctxs, uint32_t n_ctx);      /* ctx r becomes rank r of n_ctx; distinct devices */
int vlb_comm_destroy(vlb_ctx* ctx);                                /* also done by vlb_ctx_destroy                    */
int vlb_comm_info(const vlb_ctx* ctx, int32_t* rank, int32_t* world, int32_t* nccl_version);   /* any pointer may be NULL */
/* Replicated uploads: with this on, every rank must pass IDENTICAL host arrays to vlb_scene_set_triangles /
 * vlb_scene_load_gltf / vlb_skybox_set[_async]; each rank then copies only its 1/N-th of the vertex, index and texel
 * arrays over PCIe and one in-place all-gather over NVLink completes them on every GPU (the box's host memory and
 * PCIe root see the scene once instead of N times). Off by default. */
int vlb_comm_sharded_uploads(vlb_ctx* ctx, int enable);
/* LightBaker::bake over the communicator: `s` describes the WHOLE grid (slab_k1 < 0); this rank bakes its cyclic
 * z-slices, the shares are all-gathered and un-interleaved, and d_full_out ([Nx*Ny*Nz][48] floats, x-fastest, 16-byte
 * aligned) holds the whole grid on EVERY rank when the ctx stream reaches this point. d_prev_full as in
 * vlb_bake_gather_device (NULL = direct pass). Enqueues on the ctx stream, does not synchronise. Without a
 * communicator it is vlb_bake_gather_device over the whole grid. Bit-identical to the one-GPU bake. */
int vlb_bake_probes_sharded_device(vlb_ctx* ctx, const vlb_bake_settings* s, const float* d_prev_full, float* d_full_out);
/* Host-pointer variant: 1 + s->bounces passes on the devices with the all-gather between passes; the last pass is
 * copied to `out` (Nx*Ny*Nz x 48 floats) on the ranks that pass a non-NULL pointer. Synchronous. */
int vlb_bake_probes_sharded(vlb_ctx* ctx, const vlb_bake_settings* s, float* out_or_null);
/* The same, but this rank copies only the z-slices IT baked into `grid` (the whole grid's layout, Nx*Ny*Nz x 48 floats):
 * the ranks of one host all pass the SAME buffer -- shared memory between processes (pin it with cudaHostRegister), the
 * same pointer between the contexts of one process -- and fill it together, every rank over its own PCIe link (1/N of
 * the grid each, instead of one rank reading all of it). The grid is complete once every rank has returned. The device
 * side is unchanged: the all-gather still leaves the whole grid on every GPU (the multi-bounce passes read it). */
int vlb_bake_probes_sharded_rows(vlb_ctx* ctx, const vlb_bake_settings* s, float* grid);

/* Device-resident output; enqueues on the ctx stream and returns WITHOUT synchronising. */
int vlb_bake_probes_device(vlb_ctx* ctx, const vlb_bake_settings* s, float* d_out);
/* One gather pass with a device-resident source: bakes the slab exactly like
 * vlb_bake_probes_device, but every hit adds indirect_gain * gather(d_prev_full) where d_prev_full
 * is the PREVIOUS pass over the WHOLE grid ([Nx*Ny*Nz][48] floats, x-fastest probe order; the
 * reference's consumer layout, shaders/sh.rmiss:25-28). s->bounces is ignored here: the caller
 * iterates (and, when the grid is sharded over GPUs, all-gathers the slabs between passes).
 * d_prev_full == NULL is the direct pass; otherwise it must be 16-byte aligned. The gather operator (shaders/main.rchit:124-163): cell
 * ijk = floor((P - origin) / step) clamped into the grid; for its 8 corners in the reference's order
 * a visibility ray from P + bias*N to the corner probe, weight = |step| - distance (clamped at 0),
 * value = sum_i prev[corner][i] * SH_i(N); result = sum(w * value) / sum(w) over visible corners. */
int vlb_bake_gather_device(vlb_ctx* ctx, const vlb_bake_settings* s, const float* d_prev_full, float* d_out);
/* Statistics of the last bake call. Synchronises with that call; returns VLB_ERR_UNSUPPORTED if
 * its traversal overflowed the per-ray stack (vlb_bake_probes reports that itself). */
int vlb_bake_last_stats(vlb_ctx* ctx, vlb_bake_stats* out);

/* --- diagnostics: measured ceilings for the traversal kernel's roofline ------------------
 * The bake kernel is bound by the L1 data pipe (scattered 16-byte loads of L1/L2-resident BVH nodes), so bench.py divides
 * its load rates by what THIS device delivers to three micro-kernels (csrc/diag.cu): an L2-resident stream (L1 bypassed),
 * an L1-resident stream, and the traversal's own access shape (8 distinct 128-byte lines per warp-level load, 4 lanes per
 * address). No reference counterpart (SURVEY 8(d) asks for the roofline the numbers are held against). */
typedef struct vlb_cache_peaks {
    double l2_read_gbs;                   /* GB/s, 48 MB buffer, ld.global.cg */
    double l1_read_gbs;                   /* GB/s, 32 KB per block, coalesced */
    double l1_scatter_lines_per_request;  /* 8 */
    double l1_scatter_requests_per_s;     /* warp-level 16-byte load instructions per second, whole device */
    double l1_scatter_wavefronts_per_s;   /* requests x lines */
    double l1_scatter_gbs;                /* bytes delivered to the 32 lanes of those requests, GB/s */
} vlb_cache_peaks;
int vlb_diag_cache_peaks(vlb_ctx* ctx, vlb_cache_peaks* out);

/* --- validation entry points (BVH hit IDs bit-exact vs brute force) --------------------- */
enum { VLB_TRACE_BVH = 0, VLB_TRACE_BRUTE_FORCE = 1 };
enum { VLB_TRACE_CLOSEST = 0, VLB_TRACE_ANY = 1 };
/* origins/dirs: n x 3 floats (host). hit_ids: flat triangle id (instance prefix sum +
 * primitive id) or -1; hit_tuv: n x 3 floats (t, u, v) or NULL. */
int vlb_trace_rays(vlb_ctx* ctx, const float* origins, const float* dirs, uint64_t n,
                   float tmin, float tmax, int accel, int kind,
                   int32_t* hit_ids, float* hit_tuv);

/* --- on-disk format (LightBaker::serialize, src/baker/light_baker.cpp:375-402; reader
 * Scene_t::loadBakedLight, src/scene_manager.cpp:613-648) -------------------------------- */
int vlb_bake_serialize_gltf(const char* in_gltf_path, const char* out_gltf_path,
                            const float* coeffs, uint64_t n_probes,
                            const vlb_bake_settings* s);
/* Reads back what vlb_bake_serialize_gltf (or the reference) wrote. coeffs may be NULL to
 * query n_probes_out first. */
int vlb_bake_deserialize_gltf(const char* gltf_path, float* coeffs, uint64_t capacity_floats,
                              uint64_t* n_floats_out, float grid_step_out[3]);

#ifdef __cplusplus
}
#endif
#endif /* VLB_BAKE_H */

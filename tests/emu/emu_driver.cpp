// emu_driver.cpp — TEST-ONLY host emulation of the device code.
//
// The per-thread bodies of the CUDA kernels live in `__host__ __device__` headers
// (csrc/vlb_math.cuh, vlb_bvh.cuh, vlb_shade.cuh). This driver runs those same bodies serially
// on the CPU so the LBVH build, traversal and shading LOGIC can be checked against the oracle
// in the GPU-less container before GPU time is spent. It is never linked into libvlb_bake.so
// and nothing in the product path can reach it; the GPU parity tests (-m gpu) remain the gate.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <vector>

#include "../../vulkan-light-bakery_b200/csrc/vlb_context.h"
#include "../../vulkan-light-bakery_b200/csrc/vlb_shade.cuh"
#include "../../vulkan-light-bakery_b200/csrc/vlb_ploc.cuh"

using namespace vlb;

struct EmuScene {
    std::vector<float4> tri_flat, tri_shade, tri_uv, inst, base_color, tris, nodes, sky;
    std::vector<int4> tex_desc;
    std::vector<uchar4> tex_texels;
    std::vector<int> left, right, first, last, parent_i, parent_l;
    int sky_w = 0, sky_h = 0;
    uint32_t n = 0;
    int max_depth = 0;
};

static float4 mk4(float x, float y, float z, float w) { float4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }

static int g_builder = 0, g_ploc_radius = 16;    // 0: Karras LBVH, 1: PLOC (vlb_ploc.cuh)

// PLOC (experiment, see vlb_ploc.cuh), round by round; fills the arrays emit_node4 consumes and reorders the
// triangles / leaf boxes into depth-first order.
static void emu_ploc(EmuScene* s, std::vector<float4>& lbox, std::vector<float4>& ibox, int radius) {
    const int n = (int)s->n;
    std::vector<float4> box(2 * (size_t)(2 * n - 1));
    std::copy(lbox.begin(), lbox.end(), box.begin());
    std::vector<int> C(n), Cn(n), nn(n), left(n - 1), right(n - 1), parent(2 * n - 1, -1), count(2 * n - 1, 1), leftmost(2 * n - 1);
    std::iota(C.begin(), C.end(), 0);
    std::iota(leftmost.begin(), leftmost.begin() + n, 0);
    int m = n, created = 0;
    const bool trace = getenv("VLB_EMU_PLOC_TRACE") != nullptr;
    int round = 0;
    while (m > 1) {
        if (trace) fprintf(stderr, "ploc round %d: m = %d\n", round++, m);
        for (int i = 0; i < m; ++i) nn[i] = ploc_nearest(C.data(), m, i, radius, box.data());
        int out = 0;
        for (int i = 0; i < m; ++i) {
            const int role = ploc_role(nn.data(), i);
            if (role == 2) continue;
            Cn[out++] = role == 1 ? ploc_merge(C.data(), nn.data(), i, created++, n, box.data(), left.data(), right.data(), parent.data(),
                                               count.data(), leftmost.data())
                                  : C[i];
        }
        C.swap(Cn);
        m = out;
    }
    std::vector<int> pos(n);
    for (int i = 0; i < n; ++i) pos[i] = ploc_leaf_position(i, n, left.data(), parent.data(), count.data());
    std::vector<float4> tris(s->tris.size()), lb(lbox.size());
    for (int i = 0; i < n; ++i) {
        for (int k = 0; k < 3; ++k) tris[3 * (size_t)pos[i] + k] = s->tris[3 * (size_t)i + k];
        lb[2 * (size_t)pos[i]] = lbox[2 * (size_t)i]; lb[2 * (size_t)pos[i] + 1] = lbox[2 * (size_t)i + 1];
    }
    s->tris.swap(tris); lbox.swap(lb);
    for (int k = 0; k < n - 1; ++k)
        ploc_finish_node(k, n, left.data(), right.data(), count.data(), leftmost.data(), pos.data(), box.data(), s->left.data(), s->right.data(),
                         s->first.data(), s->last.data(), ibox.data());
}

extern "C" {

void* emu_scene_create(const vlb_vertex* verts, const uint32_t* indices, const vlb_instance* insts, uint32_t n_insts,
                       const vlb_material* mats, uint32_t n_mats, int max_leaf) {
    EmuScene* s = new EmuScene();
    for (uint32_t m = 0; m < std::max(n_mats, 1u); ++m) {
        float4 b = mk4(1, 1, 1, i2f(-1));
        if (m < n_mats) {
            const float* f = mats[m].base_color_factor;
            if (f[0] != 0.f || f[1] != 0.f || f[2] != 0.f || f[3] != 0.f) b = mk4(f[0], f[1], f[2], i2f(-1));
            if (mats[m].base_color.index >= 0) b.w = i2f(mats[m].base_color.index);
        }
        s->base_color.push_back(b);
    }
    for (uint32_t i = 0; i < n_insts; ++i) {   // k_flatten
        const vlb_instance& vi = insts[i];
        float minv[9];
        host_inverse3x3(vi.transform, minv);
        uint32_t mat = vi.material_index < n_mats ? vi.material_index : (n_mats ? n_mats - 1 : 0);
        s->inst.push_back(mk4(minv[0], minv[1], minv[2], i2f((int)mat)));
        s->inst.push_back(mk4(minv[3], minv[4], minv[5], 0));
        s->inst.push_back(mk4(minv[6], minv[7], minv[8], 0));
        for (uint32_t t = 0; t < vi.index_count / 3; ++t) {
            Vec3 p[3], nn[3];
            float uv[6];
            for (int k = 0; k < 3; ++k) {
                const vlb_vertex& v = verts[vi.first_vertex + indices[vi.first_index + 3 * t + k]];
                p[k] = xform_point(vi.transform, mk3(v.position[0], v.position[1], v.position[2]));
                nn[k] = mk3(v.normal[0], v.normal[1], v.normal[2]);
                uv[2 * k] = v.uv0[0]; uv[2 * k + 1] = v.uv0[1];
            }
            s->tri_uv.push_back(mk4(uv[0], uv[1], uv[2], uv[3]));
            s->tri_uv.push_back(mk4(uv[4], uv[5], 0, 0));
            const int id = (int)(s->tri_flat.size() / 3);
            s->tri_flat.push_back(mk4(p[0].x, p[0].y, p[0].z, i2f(id)));
            s->tri_flat.push_back(mk4(f_sub(p[1].x, p[0].x), f_sub(p[1].y, p[0].y), f_sub(p[1].z, p[0].z), 0));
            s->tri_flat.push_back(mk4(f_sub(p[2].x, p[0].x), f_sub(p[2].y, p[0].y), f_sub(p[2].z, p[0].z), 0));
            s->tri_shade.push_back(mk4(nn[0].x, nn[0].y, nn[0].z, i2f((int)i)));
            s->tri_shade.push_back(mk4(nn[1].x, nn[1].y, nn[1].z, 0));
            s->tri_shade.push_back(mk4(nn[2].x, nn[2].y, nn[2].z, 0));
        }
    }
    const uint32_t n = (uint32_t)(s->tri_flat.size() / 3);
    s->n = n;
    if (n == 0) return s;
    // k_scene_bounds
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    float clo[3] = {INFINITY, INFINITY, INFINITY}, chi[3] = {-INFINITY, -INFINITY, -INFINITY};
    std::vector<float> cen(3 * (size_t)n);
    for (uint32_t t = 0; t < n; ++t) {
        float4 a, b;
        tri_aabb(s->tri_flat[3 * t], s->tri_flat[3 * t + 1], s->tri_flat[3 * t + 2], &a, &b);
        const float al[3] = {a.x, a.y, a.z}, bh[3] = {b.x, b.y, b.z};
        for (int k = 0; k < 3; ++k) {
            lo[k] = std::min(lo[k], al[k]); hi[k] = std::max(hi[k], bh[k]);
            cen[3 * t + k] = 0.5f * (al[k] + bh[k]);
            clo[k] = std::min(clo[k], cen[3 * t + k]); chi[k] = std::max(chi[k], cen[3 * t + k]);
        }
    }
    // k_morton + sort
    std::vector<uint64_t> keys(n);
    std::vector<uint32_t> order(n);
    for (uint32_t t = 0; t < n; ++t) {
        float nr[3];
        for (int k = 0; k < 3; ++k) { float e = chi[k] - clo[k]; nr[k] = e > 0 ? (cen[3 * t + k] - clo[k]) / e : 0.f; }
        keys[t] = morton63(nr[0], nr[1], nr[2]);
    }
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return keys[a] < keys[b]; });
    std::vector<uint64_t> ks(n);
    for (uint32_t j = 0; j < n; ++j) ks[j] = keys[order[j]];
    // k_gather
    s->tris.resize(3 * (size_t)n);
    std::vector<float4> lbox(2 * (size_t)n), ibox(2 * (size_t)n);
    for (uint32_t j = 0; j < n; ++j) {
        for (int k = 0; k < 3; ++k) s->tris[3 * j + k] = s->tri_flat[3 * order[j] + k];
        tri_aabb(s->tris[3 * j], s->tris[3 * j + 1], s->tris[3 * j + 2], &lbox[2 * j], &lbox[2 * j + 1]);
    }
    const float ext = std::max(hi[0] - lo[0], std::max(hi[1] - lo[1], hi[2] - lo[2]));
    s->nodes.resize((size_t)kNodeQuads * std::max(n - 1, 1u));
    if (n == 1) {
        emit_wide_single(lbox.data(), ext * 1e-6f, s->nodes.data());
        return s;
    }
    s->left.assign(n, 0); s->right.assign(n, 0); s->first.assign(n, 0); s->last.assign(n, 0);
    s->parent_i.assign(n, -1); s->parent_l.assign(n, -1);
    if (g_builder == 1) {
        emu_ploc(s, lbox, ibox, g_ploc_radius);
    } else {
        for (int i = 0; i < (int)n - 1; ++i)   // k_karras
            karras_node(ks.data(), (int)n, i, s->left.data(), s->right.data(), s->first.data(), s->last.data(),
                        s->parent_i.data(), s->parent_l.data());
        s->parent_i[0] = -1;
        std::vector<int> flags(n, 0);
        for (int j = 0; j < (int)n; ++j) {     // k_refit
            int cur = s->parent_l[j];
            while (cur >= 0) {
                if (flags[cur]++ == 0) break;
                float4 l[2], h[2];
                const int ch[2] = {s->left[cur], s->right[cur]};
                for (int c = 0; c < 2; ++c) {
                    if (ch[c] < 0) { l[c] = lbox[2 * (~ch[c])]; h[c] = lbox[2 * (~ch[c]) + 1]; }
                    else { l[c] = ibox[2 * ch[c]]; h[c] = ibox[2 * ch[c] + 1]; }
                }
                ibox[2 * cur] = mk4(fminf(l[0].x, l[1].x), fminf(l[0].y, l[1].y), fminf(l[0].z, l[1].z), 0);
                ibox[2 * cur + 1] = mk4(fmaxf(h[0].x, h[1].x), fmaxf(h[0].y, h[1].y), fmaxf(h[0].z, h[1].z), 0);
                cur = s->parent_i[cur];
            }
        }
    }
    {   // k_emit_level, level by level as bvh_build.cu does
        std::vector<int> frontier{0}, next(n);
        while (!frontier.empty()) {
            unsigned int n_next = 0;
            for (int root : frontier)
                emit_wide_node(root, s->left.data(), s->right.data(), s->first.data(), s->last.data(), ibox.data(), lbox.data(), max_leaf,
                           ext * 1e-6f, s->nodes.data(), next.data(), &n_next);
            frontier.assign(next.begin(), next.begin() + n_next);
        }
    }
    // worst-case number of pending stack entries, by DFS over the emitted nodes (3 pushes per level)
    std::vector<std::pair<int, int>> st; st.push_back({0, 1});
    while (!st.empty()) {
        auto [nd, dp] = st.back(); st.pop_back();
        s->max_depth = std::max(s->max_depth, dp);
        const int* r = reinterpret_cast<const int*>(&s->nodes[(size_t)kNodeQuads * nd + kRefQuad]);
        for (int k = 0; k < kWide; ++k) if (r[k] >= 0) st.push_back({r[k], dp + 1});
    }
    return s;
}

// {wide nodes reachable from the root, child slots in use, leaf children} of the emitted tree
void emu_scene_node_stats(void* h, uint64_t out[3]) {
    EmuScene* s = (EmuScene*)h;
    out[0] = out[1] = out[2] = 0;
    if (s->n == 0) return;
    std::vector<int> st{0};
    while (!st.empty()) {
        const int nd = st.back(); st.pop_back();
        ++out[0];
        const int* r = reinterpret_cast<const int*>(&s->nodes[(size_t)kNodeQuads * nd + kRefQuad]);
        for (int k = 0; k < kWide; ++k) {
            if (r[k] == kNoChild) continue;
            ++out[1];
            if (r[k] < 0) ++out[2]; else st.push_back(r[k]);
        }
    }
}
void emu_scene_destroy(void* h) { delete (EmuScene*)h; }
void emu_set_builder(int builder, int ploc_radius) { g_builder = builder; g_ploc_radius = ploc_radius > 0 ? ploc_radius : 16; }
int emu_scene_max_depth(void* h) { return ((EmuScene*)h)->max_depth; }
// vlb_scene_set_textures, as context.cu lays the atlas out
void emu_scene_set_textures(void* h, const vlb_texture* tex, uint32_t n) {
    EmuScene* s = (EmuScene*)h;
    s->tex_desc.clear(); s->tex_texels.clear();
    for (uint32_t i = 0; i < n; ++i) {
        int4 d; d.x = (int)s->tex_texels.size(); d.y = tex[i].width; d.z = tex[i].height;
        d.w = tex[i].wrap_u | (tex[i].wrap_v << 2) | (tex[i].filter << 4);
        s->tex_desc.push_back(d);
        const uchar4* p = static_cast<const uchar4*>(tex[i].texels);
        s->tex_texels.insert(s->tex_texels.end(), p, p + (size_t)tex[i].width * tex[i].height);
    }
}
void emu_tex_sample(void* h, int tex, const float* uv, uint64_t n, float* rgb) {
    EmuScene* s = (EmuScene*)h;
    ShadeView sv{}; sv.tex_desc = s->tex_desc.data(); sv.tex_texels = s->tex_texels.data();
    for (uint64_t i = 0; i < n; ++i) tex_sample(sv, tex, uv[2 * i], uv[2 * i + 1], rgb + 3 * i);
}
void emu_scene_set_skybox(void* h, const float* rgba32f, int W, int H) {
    EmuScene* s = (EmuScene*)h;
    s->sky.resize((size_t)W * H);
    memcpy(s->sky.data(), rgba32f, (size_t)W * H * 16);
    s->sky_w = W; s->sky_h = H;
}

static BvhView view(EmuScene* s) { BvhView b; b.nodes = s->nodes.data(); b.tris = s->tris.data(); b.n_tris = s->n; b.overflow = nullptr; return b; }

void emu_trace_rays(void* h, const float* o, const float* d, uint64_t n, float tmin, float tmax, int kind, int32_t* ids,
                    float* tuv, uint64_t* counters) {
    EmuScene* s = (EmuScene*)h;
    BvhView b = view(s);
    TraceCounters cnt; cnt.nodes = 0; cnt.tris = 0;
    uint64_t nn = 0, nt = 0;
    for (uint64_t i = 0; i < n; ++i) {
        const Vec3 ro = mk3(o[3 * i], o[3 * i + 1], o[3 * i + 2]), rd = mk3(d[3 * i], d[3 * i + 1], d[3 * i + 2]);
        HitRec r; r.id = -1; r.t = tmax; r.u = r.v = 0;
        cnt.nodes = cnt.tris = 0;
        if (kind == VLB_TRACE_ANY) { HitRec a; if (trace_any<true>(b, ro, rd, tmin, tmax, &cnt, &a)) r = a; }
        else r = trace_closest<true>(b, ro, rd, tmin, tmax, &cnt);
        nn += cnt.nodes; nt += cnt.tris;
        ids[i] = r.id;
        if (tuv) { tuv[3 * i] = r.t; tuv[3 * i + 1] = r.u; tuv[3 * i + 2] = r.v; }
    }
    if (counters) { counters[0] = nn; counters[1] = nt; }
}

// k_bake, serially: probes of the slab in x-fastest order, out = n x 48 floats. prev_full != NULL:
// a gather pass over the previous pass of the whole grid (vlb_bake_gather_device).
void emu_bake_gather(void* h, const vlb_bake_settings* st, const float* prev_full, float* out) {
    EmuScene* s = (EmuScene*)h;
    BvhView b = view(s);
    ShadeView sv; sv.tri_shade = s->tri_shade.data(); sv.inst = s->inst.data(); sv.base_color = s->base_color.data();
    sv.tri_uv = s->tri_uv.data(); sv.tex_desc = s->tex_desc.data(); sv.tex_texels = s->tex_texels.data();
    sv.sky = s->sky_w ? s->sky.data() : nullptr; sv.sky_w = s->sky_w; sv.sky_h = s->sky_h;
    BakeConsts c;
    for (int k = 0; k < 3; ++k) c.light[k] = st->light_pos[k];
    c.shadow_bias = st->shadow_bias; c.c_diffuse = st->c_diffuse; c.c_specular = st->c_specular; c.gloss = st->gloss;
    c.ambient = st->ambient; c.tmin = st->tmin; c.tmax = st->tmax; c.flags = st->flags;
    const int Nx = st->probes[0], Ny = st->probes[1], Nz = st->probes[2], W = st->dir_w, H = st->dir_h;
    std::vector<float> px(Nx), py(Ny), pz(Nz), row(2 * (size_t)H), col(2 * (size_t)W);
    host_axis_coords(st->origin[0], st->step[0], Nx, px.data());
    host_axis_coords(st->origin[1], st->step[1], Ny, py.data());
    host_axis_coords(st->origin[2], st->step[2], Nz, pz.data());
    host_dir_tables(W, H, 0.f, row.data(), col.data());
    const float pixel_area = (2.0f * kPi / (float)W) * (kPi / (float)H);
    const int k0 = st->slab_k1 < 0 ? 0 : st->slab_k0, k1 = st->slab_k1 < 0 ? Nz : st->slab_k1;
    const int K = st->sh_order == 2 ? 9 : 16;
    GatherView g; g.prev = prev_full; g.px = px.data(); g.py = py.data(); g.pz = pz.data(); g.Nx = Nx; g.Ny = Ny; g.Nz = Nz;
    for (int k = 0; k < 3; ++k) { g.origin[k] = st->origin[k]; g.step[k] = st->step[k]; }
    g.gain = st->indirect_gain; g.world_frame = (st->flags & VLB_BAKE_SH_WORLD_FRAME) ? 1 : 0;
    size_t q = 0;
    for (int k = k0; k < k1; ++k) for (int j = 0; j < Ny; ++j) for (int i = 0; i < Nx; ++i, ++q) {
        double acc[48] = {0};
        const Vec3 o = mk3(px[i], py[j], pz[k]);
        for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) {
            const Vec3 t = to_vector_sc(row[2 * y], row[2 * y + 1], col[2 * x], col[2 * x + 1]);
            const Vec3 r = mk3(t.x, t.z, t.y);
            float rgb[3];
            if (K == 9) probe_ray_radiance<false, 9>(b, sv, c, o, r, rgb, nullptr, nullptr, &g);
            else        probe_ray_radiance<false, 16>(b, sv, c, o, r, rgb, nullptr, nullptr, &g);
            const float w = pixel_area * row[2 * y];
            float bs[16];
            sh_basis<16>((st->flags & VLB_BAKE_SH_WORLD_FRAME) ? r : t, bs);
            for (int m = 0; m < K; ++m) for (int ch = 0; ch < 3; ++ch) acc[3 * m + ch] += (double)(bs[m] * w) * rgb[ch];
        }
        for (int m = 0; m < 48; ++m) out[q * 48 + m] = (float)acc[m];
    }
}

void emu_bake(void* h, const vlb_bake_settings* st, float* out) { emu_bake_gather(h, st, nullptr, out); }

}  // extern "C"

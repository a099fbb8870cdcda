// vlb_baker — the `baker` executable of the reference (src/baker/main.cpp:8-44) on top of the C ABI:
//
//     vlb_baker <scene.gltf|.glb> [options]       ->  baked_<scene>.gltf
//
// Same life cycle as `vlb::LightBaker baker{file}; baker.bake(); baker.serialize();`
// (main.cpp:22-24): load the scene (SceneManager::pushScene, scene_manager.cpp:1013-1034), lay the
// probe grid over Scene_t::getBounds() (light_baker.cpp:50-53,80-101), bake every probe
// (light_baker.cpp:287-328) and write the glTF with the coefficient buffer appended
// (light_baker.cpp:375-402). With no options every constant is the reference's: 7x7x7 probes,
// 3141x1000 directions, 16 coefficients, light (1,10,1) (light_baker.cpp:38,65,294;
// env_map.rchit:25). Errors print "std::exception: <text>" and exit with EXIT_FAILURE, success prints
// "exiting..." (main.cpp:26-43). Only vlb_bake.h is used: this file is also the worked example of
// the boundary.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "vlb_bake.h"

namespace {

struct Options {
    std::string scene, out, skybox;
    int probes[3] = {0, 0, 0};      // 0 = reference default
    int dirs[2] = {0, 0};
    int order = 0;
    int device = 0;
    std::vector<int> devices;       // --devices a,b,...: one context per GPU, vlb_bake_probes_multi
    int builder = -1;               // --builder lbvh|ploc|auto: -1 = auto (vlb_bvh_recommend_builder)
    int bounces = 0;
    float gain = -1.f;
    bool tight = false, have_light = false, dry = false;
    float light[3] = {0, 0, 0};
    unsigned flags_set = 0, flags_clear = 0;
};

[[noreturn]] void die(const std::string& what) {
    fprintf(stderr, "std::exception: %s\n", what.c_str());     // main.cpp:33
    exit(EXIT_FAILURE);
}

bool parse_ints(const char* s, int* v, int n) {
    for (int i = 0; i < n; ++i) {
        char* end = nullptr;
        const long x = strtol(s, &end, 10);
        if (end == s || x <= 0) return false;
        v[i] = (int)x;
        s = end;
        if (i + 1 < n) {
            if (*s != 'x' && *s != 'X' && *s != ',') return false;
            ++s;
        }
    }
    return *s == 0;
}

bool parse_floats(const char* s, float* v, int n) {
    for (int i = 0; i < n; ++i) {
        char* end = nullptr;
        v[i] = strtof(s, &end);
        if (end == s) return false;
        s = end;
        if (i + 1 < n) {
            if (*s != ',') return false;
            ++s;
        }
    }
    return *s == 0;
}

void usage() {
    puts("usage: vlb_baker <scene.gltf|scene.glb> [options]\n"
         "  --probes NxNyxNz     probe grid (default 7x7x7, light_baker.cpp:38)\n"
         "  --builder B          lbvh | ploc | auto (default): hierarchy builder, auto = by job size (vlb_bvh_recommend_builder)\n"
         "  --dirs WxH           equirect direction grid per probe (default 3141x1000, light_baker.cpp:65)\n"
         "  --order 2|3          SH order written: 9 or 16 coefficients (default 3)\n"
         "  --light x,y,z        point light position (default 1,10,1, env_map.rchit:25)\n"
         "  --bounces B          gather passes after the direct pass (default 0 = the reference bake)\n"
         "  --gain g             indirect gain of the gather passes\n"
         "  --tight-bounds       lay the grid over the true world AABB instead of Scene_t::getBounds()\n"
         "  --no-shadows --no-srgb --no-quantize8 --reference-order --world-frame   behaviour flags (vlb_bake.h)\n"
         "  --device N           CUDA device (default 0)\n"
         "  --devices a,b,...    bake on several GPUs from this one process (probe z-slices dealt cyclically)\n"
         "  --out path           output file (default baked_<scene>)\n"
         "  --skybox image       equirect PNG / JPEG / Radiance .hdr sampled where a ray leaves the scene (main.rmiss:18-40); default: none\n"
         "  --dry-run            parse the scene and print what would be baked; needs no GPU");
}

std::string default_out(const std::string& in) {
    // light_baker.cpp:399 writes "baked_" + fileName; for a bare file name that is what this returns,
    // for a path the prefix goes on the file name so the result lands beside the input.
    const size_t slash = in.find_last_of('/');
    if (slash == std::string::npos) return "baked_" + in;
    return in.substr(0, slash + 1) + "baked_" + in.substr(slash + 1);
}

}  // namespace

int main(int argc, char** argv) {
    Options o;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto need = [&](const char* name) -> const char* {
            if (i + 1 >= argc) die(std::string("option ") + name + " needs a value");
            return argv[++i];
        };
        if (a == "-h" || a == "--help") { usage(); return EXIT_SUCCESS; }
        else if (a == "--builder") {
            const std::string b = need("--builder");
            if (b == "lbvh") o.builder = VLB_BVH_BUILDER_LBVH; else if (b == "ploc") o.builder = VLB_BVH_BUILDER_PLOC;
            else if (b == "auto") o.builder = -1; else die("--builder expects lbvh, ploc or auto");
        }
        else if (a == "--probes") { if (!parse_ints(need("--probes"), o.probes, 3)) die("--probes expects NxNyxNz"); }
        else if (a == "--dirs") { if (!parse_ints(need("--dirs"), o.dirs, 2)) die("--dirs expects WxH"); }
        else if (a == "--order") { o.order = atoi(need("--order")); if (o.order != 2 && o.order != 3) die("--order expects 2 or 3"); }
        else if (a == "--light") { if (!parse_floats(need("--light"), o.light, 3)) die("--light expects x,y,z"); o.have_light = true; }
        else if (a == "--bounces") { o.bounces = atoi(need("--bounces")); if (o.bounces < 0) die("--bounces expects >= 0"); }
        else if (a == "--gain") { o.gain = strtof(need("--gain"), nullptr); }
        else if (a == "--device") { o.device = atoi(need("--device")); }
        else if (a == "--devices") {
            const char* v = need("--devices");
            while (*v) {
                char* end = nullptr;
                const long d = strtol(v, &end, 10);
                if (end == v || d < 0) die("--devices expects a comma separated list of CUDA device ids");
                o.devices.push_back((int)d);
                v = *end == ',' ? end + 1 : end;
                if (*end && *end != ',') die("--devices expects a comma separated list of CUDA device ids");
            }
            if (o.devices.empty()) die("--devices expects at least one device id");
        }
        else if (a == "--out") { o.out = need("--out"); }
        else if (a == "--skybox") { o.skybox = need("--skybox"); }
        else if (a == "--tight-bounds") o.tight = true;
        else if (a == "--no-shadows") o.flags_clear |= VLB_BAKE_SHADOW_RAYS;
        else if (a == "--no-srgb") o.flags_clear |= VLB_BAKE_SRGB_ENCODE;
        else if (a == "--no-quantize8") o.flags_clear |= VLB_BAKE_QUANTIZE_RGBA8;   // the RGBA8 image store (env_map_generator.hpp:39) is in the defaults
        else if (a == "--quantize8") {}                                              // accepted for compatibility: already the default
        else if (a == "--reference-order") o.flags_set |= VLB_BAKE_REFERENCE_PROBE_ORDER;
        else if (a == "--world-frame") o.flags_set |= VLB_BAKE_SH_WORLD_FRAME;
        else if (a == "--dry-run") o.dry = true;
        else if (!a.empty() && a[0] == '-') die("unknown option " + a);
        else if (o.scene.empty()) o.scene = a;
        else die("more than one scene given");
    }
    if (o.scene.empty()) die("Select scene to bake.");                       // main.cpp:15
    if (o.scene.find(".gltf") == std::string::npos && o.scene.find(".glb") == std::string::npos)
        // light_baker.cpp:68-73: any other asset is an environment IMAGE projected by sh.comp; decoding
        // image files is outside the bake path here (vlb_envmap_project_sh takes raw texels).
        die("image input is not handled by vlb_baker; pass raw texels to vlb_envmap_project_sh");
    if (o.out.empty()) o.out = default_out(o.scene);

    vlb_bake_settings s;
    vlb_bake_settings_default(&s);
    for (int k = 0; k < 3; ++k) if (o.probes[k]) s.probes[k] = o.probes[k];
    if (o.dirs[0]) { s.dir_w = o.dirs[0]; s.dir_h = o.dirs[1]; }
    if (o.order) s.sh_order = o.order;
    if (o.have_light) for (int k = 0; k < 3; ++k) s.light_pos[k] = o.light[k];
    // without --skybox misses contribute 0, as in the reference's own bake pipeline, which binds no skybox
    // (SURVEY App. B-5); with it the evident intent of main.rmiss is what runs
    s.flags = (s.flags | o.flags_set) & ~o.flags_clear;
    if (o.skybox.empty()) s.flags &= ~(unsigned)VLB_BAKE_SKYBOX_ON_MISS; else s.flags |= VLB_BAKE_SKYBOX_ON_MISS;
    s.bounces = o.bounces;
    if (o.gain >= 0.f) s.indirect_gain = o.gain;

    uint64_t counts[5] = {0, 0, 0, 0, 0};
    float ref_bounds[6];
    if (vlb_gltf_probe(o.scene.c_str(), counts, ref_bounds) != VLB_OK) die(vlb_last_error(nullptr));
    const uint64_t n_probes = (uint64_t)s.probes[0] * s.probes[1] * s.probes[2];
    printf("scene %s: %llu triangles, %llu instances, %llu materials\n", o.scene.c_str(), (unsigned long long)counts[4],
           (unsigned long long)counts[2], (unsigned long long)counts[3]);
    if (o.dry) {
        if (vlb_bake_settings_from_bounds(&s, ref_bounds) != VLB_OK) die(vlb_last_error(nullptr));
        printf("dry run: %dx%dx%d probes x %dx%d rays, order %d, origin (%g %g %g), step (%g %g %g) -> %s\n", s.probes[0],
               s.probes[1], s.probes[2], s.dir_w, s.dir_h, s.sh_order, s.origin[0], s.origin[1], s.origin[2], s.step[0],
               s.step[1], s.step[2], o.out.c_str());
        puts("exiting...");
        return EXIT_SUCCESS;
    }

    if (o.devices.empty()) o.devices.push_back(o.device);
    std::vector<vlb_ctx*> ctxs;
    auto destroy_all = [&] { for (vlb_ctx* c : ctxs) vlb_ctx_destroy(c); ctxs.clear(); };
    for (int d : o.devices) {
        vlb_ctx* c = nullptr;
        if (vlb_ctx_create(d, &c) != VLB_OK) { const std::string m = vlb_last_error(nullptr); destroy_all(); die(m); }
        ctxs.push_back(c);
    }
    vlb_ctx* ctx = ctxs[0];
    vlb_ctx* cur = ctx;                    // the context whose error text `check` reports
    auto check = [&](int r) {
        if (r != VLB_OK) {
            const std::string msg = vlb_last_error(cur);
            destroy_all();
            die(msg);
        }
    };
    std::vector<unsigned char> sky_px;
    int32_t sky_wh[2] = {0, 0};
    bool sky_hdr = false;
    if (!o.skybox.empty()) {                                                  // Skybox_t ctor: stbi_load -> RGBA8 (skybox_manager.cpp:14-20)
        sky_hdr = o.skybox.size() >= 4 && o.skybox.compare(o.skybox.size() - 4, 4, ".hdr") == 0;   // Radiance RGBE -> RGBA32F
        auto load = sky_hdr ? vlb_image_load_rgba32f : vlb_image_load_rgba8;
        if (load(o.skybox.c_str(), nullptr, 0, sky_wh) != VLB_OK) { const std::string m = vlb_last_error(nullptr); destroy_all(); die(m); }
        sky_px.resize((size_t)sky_wh[0] * sky_wh[1] * (sky_hdr ? 16 : 4));
        if (load(o.skybox.c_str(), sky_px.data(), sky_px.size(), sky_wh) != VLB_OK) { const std::string m = vlb_last_error(nullptr); destroy_all(); die(m); }
        printf("skybox %s: %dx%d\n", o.skybox.c_str(), sky_wh[0], sky_wh[1]);
    }
    // the build preference (the reference passes ePreferFastTrace to its driver, scene_manager.cpp:346-347): chosen from
    // the size of the job unless --builder says otherwise
    int builder = o.builder;
    if (builder < 0) {
        uint64_t counts[5] = {0, 0, 0, 0, 0};
        float ref_bounds[6];
        if (vlb_gltf_probe(o.scene.c_str(), counts, ref_bounds) != VLB_OK) { const std::string m = vlb_last_error(nullptr); destroy_all(); die(m); }
        const uint64_t rays = n_probes * (uint64_t)s.dir_w * (uint64_t)s.dir_h * (uint64_t)(1 + (s.bounces > 0 ? s.bounces : 0)) / ctxs.size();
        builder = vlb_bvh_recommend_builder(counts[4], rays);
    }
    vlb_bvh_stats bs;
    for (vlb_ctx* c : ctxs) {               // scene, LBVH and skybox are replicated on every GPU
        cur = c;
        check(vlb_scene_load_gltf(c, o.scene.c_str()));                      // LightBaker ctor, light_baker.cpp:40-53
        check(vlb_bvh_set_builder(c, builder, 0));
        check(vlb_bvh_build(c, &bs));                                        // Scene_t::buildAccelerationStructures
        if (!sky_px.empty()) check(vlb_skybox_set(c, sky_px.data(), sky_hdr ? VLB_FMT_RGBA32F : VLB_FMT_RGBA8, sky_wh[0], sky_wh[1]));
    }
    cur = ctx;
    float bounds[6];
    check(vlb_scene_bounds(ctx, o.tight ? 1 : 0, bounds));                   // Scene_t::getBounds
    check(vlb_bake_settings_from_bounds(&s, bounds));                        // probePositionsFromBoudingBox
    printf("%s: %llu nodes in %.2f ms; grid %dx%dx%d over (%g %g %g)-(%g %g %g); %zu GPU(s)\n", builder == VLB_BVH_BUILDER_PLOC ? "BVH (PLOC)" : "LBVH", (unsigned long long)bs.n_nodes, bs.build_ms,
           s.probes[0], s.probes[1], s.probes[2], bounds[0], bounds[1], bounds[2], bounds[3], bounds[4], bounds[5], ctxs.size());

    std::vector<float> coeffs((size_t)n_probes * VLB_SH_STRIDE);
    if (ctxs.size() == 1) check(vlb_bake_probes(ctx, &s, coeffs.data()));    // LightBaker::bake
    else check(vlb_bake_probes_multi(ctxs.data(), (uint32_t)ctxs.size(), &s, coeffs.data()));
    double rays_p = 0, rays_s = 0, ms = 0;
    for (vlb_ctx* c : ctxs) {               // every GPU's share of the last pass; the slowest one is the bake time
        cur = c;
        vlb_bake_stats st;
        check(vlb_bake_last_stats(c, &st));
        rays_p += (double)st.n_primary_rays; rays_s += (double)st.n_shadow_rays;
        if (st.total_ms > ms) ms = st.total_ms;
    }
    cur = ctx;
    printf("baked %llu probes: %.0f primary + %.0f shadow rays in %.2f ms (%.3f Grays/s)\n", (unsigned long long)n_probes, rays_p, rays_s, ms,
           ms > 0 ? (rays_p + rays_s) / (ms * 1e-3) / 1e9 : 0.0);
    check(vlb_bake_serialize_gltf(o.scene.c_str(), o.out.c_str(), coeffs.data(), n_probes, &s));   // LightBaker::serialize
    printf("wrote %s\n", o.out.c_str());
    destroy_all();
    puts("exiting...");                                                      // main.cpp:42
    return EXIT_SUCCESS;
}

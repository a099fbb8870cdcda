// shaders/main.rmiss compiled from the reference's text (oracle/make_ref_shaders.py). TEST INFRASTRUCTURE.
#include "glsl_shim.h"
#define GLUE_DECLS "main_rmiss_decls.inc"
namespace glsl { namespace ref_rmiss {
#include "main.rmiss.inc"
}}
using namespace glsl;
extern "C" {
void ref_rmiss_srgb(const float in[4], float out[4]) {
    const vec4 r = ref_rmiss::sRGB(vec4(in[0], in[1], in[2], in[3]));
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}
void ref_rmiss_dir2uv(const float dir[3], float uv[2]) {
    const vec2 r = ref_rmiss::dir2SkyboxUV(vec3(dir[0], dir[1], dir[2]));
    uv[0] = r.x; uv[1] = r.y;
}
void ref_rmiss_run(const float dir[3], const float* sky_texels, int w, int h, float rgb_out[3]) {
    ref_rmiss::skybox.texels = sky_texels; ref_rmiss::skybox.w = w; ref_rmiss::skybox.h = h;
    ref_rmiss::gl_WorldRayDirectionEXT = vec3(dir[0], dir[1], dir[2]);
    ref_rmiss::shader_main();
    rgb_out[0] = ref_rmiss::payLoad.x; rgb_out[1] = ref_rmiss::payLoad.y; rgb_out[2] = ref_rmiss::payLoad.z;
}
}

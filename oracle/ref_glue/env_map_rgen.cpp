// shaders/env_map.rgen compiled from the reference's text (oracle/make_ref_shaders.py). TEST INFRASTRUCTURE.
#include "glsl_shim.h"
#define GLUE_DECLS "env_map_rgen_decls.inc"
namespace glsl { namespace ref_rgen {
#include "env_map.rgen.inc"
}}
using namespace glsl;
extern "C" {
void ref_rgen_set_color(const float rgb[3]) { ref_rgen::color = vec3(rgb[0], rgb[1], rgb[2]); }
// One ray-generation invocation (x, y) of a W x H launch; the result lands in image[y][x] (float RGBA, rgba8: quantised).
void ref_rgen_run(int x, int y, int W, int H, const float origin[3], float* image_texels, int rgba8) {
    ref_rgen::gl_LaunchIDEXT.x = (uint)x; ref_rgen::gl_LaunchIDEXT.y = (uint)y; ref_rgen::gl_LaunchIDEXT.z = 0;
    ref_rgen::gl_LaunchSizeEXT.x = (uint)W; ref_rgen::gl_LaunchSizeEXT.y = (uint)H; ref_rgen::gl_LaunchSizeEXT.z = 1;
    ref_rgen::envConst.origin = vec3(origin[0], origin[1], origin[2]);
    ref_rgen::image.texels = image_texels; ref_rgen::image.w = W; ref_rgen::image.h = H; ref_rgen::image.rgba8 = rgba8 != 0;
    ref_rgen::shader_main();
}
}

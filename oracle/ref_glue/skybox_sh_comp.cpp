// shaders/skybox_sh.comp compiled from the reference's text (oracle/make_ref_shaders.py). TEST INFRASTRUCTURE.
#include "glsl_shim.h"
#define GLUE_DECLS "skybox_sh_comp_decls.inc"
namespace glsl { namespace ref_skybox_sh_comp {
#include "skybox_sh.comp.inc"
}}
using namespace glsl;
// The whole dispatch, serially (skybox_manager.cpp:107-130). texels: float RGBA (an RGBA8 skybox is passed as n/255).
extern "C" void ref_skybox_sh_comp_dispatch(const float* texels, int W, int H, double* out48) {
    using namespace ref_skybox_sh_comp;
    for (int i = 0; i < 16; ++i) coeffs[i] = dvec3acc();
    constants.width = W; constants.height = H;
    skybox.texels = texels; skybox.w = W; skybox.h = H;
    const int gx = (W + WORKGROUP_SIZE - 1) / WORKGROUP_SIZE * WORKGROUP_SIZE, gy = (H + WORKGROUP_SIZE - 1) / WORKGROUP_SIZE * WORKGROUP_SIZE;
    for (int y = 0; y < gy; ++y)
        for (int x = 0; x < gx; ++x) {
            gl_GlobalInvocationID.x = (uint)x; gl_GlobalInvocationID.y = (uint)y; gl_GlobalInvocationID.z = 0;
            shader_main();
        }
    for (int i = 0; i < 16; ++i) { out48[3 * i] = coeffs[i].x; out48[3 * i + 1] = coeffs[i].y; out48[3 * i + 2] = coeffs[i].z; }
}

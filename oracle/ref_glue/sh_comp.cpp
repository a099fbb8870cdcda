// shaders/sh.comp compiled from the reference's text (oracle/make_ref_shaders.py). TEST INFRASTRUCTURE.
#include "glsl_shim.h"
#define GLUE_DECLS "sh_comp_decls.inc"
namespace glsl { namespace ref_sh_comp {
#include "sh.comp.inc"
}}
using namespace glsl;
// The whole dispatch, serially: ceil(W/16) x ceil(H/16) workgroups of 16 x 16 invocations (light_baker.cpp:269-285),
// including the out-of-range invocations the shader's first line rejects. out[16][3] in double.
extern "C" void ref_sh_comp_dispatch(float* texels, int W, int H, double* out48) {
    using namespace ref_sh_comp;
    for (int i = 0; i < 16; ++i) coeffs[i] = dvec3acc();
    constants.width = W; constants.height = H;
    environmentMap.texels = texels; environmentMap.w = W; environmentMap.h = H;
    const int gx = (W + WORKGROUP_SIZE - 1) / WORKGROUP_SIZE * WORKGROUP_SIZE, gy = (H + WORKGROUP_SIZE - 1) / WORKGROUP_SIZE * WORKGROUP_SIZE;
    for (int y = 0; y < gy; ++y)
        for (int x = 0; x < gx; ++x) {
            gl_GlobalInvocationID.x = (uint)x; gl_GlobalInvocationID.y = (uint)y; gl_GlobalInvocationID.z = 0;
            shader_main();
        }
    for (int i = 0; i < 16; ++i) { out48[3 * i] = coeffs[i].x; out48[3 * i + 1] = coeffs[i].y; out48[3 * i + 2] = coeffs[i].z; }
}

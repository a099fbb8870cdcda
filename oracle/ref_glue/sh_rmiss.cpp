// shaders/sh.rmiss compiled from the reference's text (oracle/make_ref_shaders.py). TEST INFRASTRUCTURE.
// The viewer's miss shader of a probe-visibility ray: evaluates the probe's SH on the shading normal (sh.rmiss:20-36).
#include "glsl_shim.h"
#define GLUE_DECLS "sh_rmiss_decls.inc"
namespace glsl { namespace ref_sh_rmiss {
#include "sh.rmiss.inc"
}}
using namespace glsl;
// One invocation on a copy of the caller's payload (sum, normal, ijk, lmax, occluded); sh: the SHCoeffs buffer.
extern "C" void ref_sh_rmiss_run(float sum[3], const float normal[3], const int ijk[3], unsigned lmax, int* occluded, const float* sh) {
    using namespace ref_sh_rmiss;
    pl.sum = vec3(sum[0], sum[1], sum[2]);
    pl.normal = vec3(normal[0], normal[1], normal[2]);
    pl.ijk = ivec3(ijk[0], ijk[1], ijk[2]);
    pl.lmax = lmax;
    pl.occluded = *occluded != 0;
    shCoeffs.sh = reinterpret_cast<const vec3*>(sh);
    shader_main();
    sum[0] = pl.sum.x; sum[1] = pl.sum.y; sum[2] = pl.sum.z;
    *occluded = pl.occluded ? 1 : 0;
}

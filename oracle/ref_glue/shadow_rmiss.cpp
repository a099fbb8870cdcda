// shaders/shadow.rmiss compiled from the reference's text (oracle/make_ref_shaders.py). TEST INFRASTRUCTURE.
#include "glsl_shim.h"
#define GLUE_DECLS "shadow_rmiss_decls.inc"
namespace glsl { namespace ref_shadow_rmiss {
#include "shadow.rmiss.inc"
}}
extern "C" int ref_shadow_rmiss_run(void) {
    glsl::ref_shadow_rmiss::inShadow = true;
    glsl::ref_shadow_rmiss::shader_main();
    return glsl::ref_shadow_rmiss::inShadow ? 1 : 0;
}

// shaders/env_map.rchit compiled from the reference's text (oracle/make_ref_shaders.py). TEST INFRASTRUCTURE.
#include "glsl_shim.h"
#define GLUE_DECLS "env_map_rchit_decls.inc"
namespace glsl { namespace ref_rchit {
#include "env_map.rchit.inc"
static_assert(sizeof(InstanceInfo) == 24 && sizeof(Vertex) == 44 && sizeof(Material) == 144, "structures.h scalar layout");
}}
using namespace glsl;
extern "C" {
void ref_rchit_srgb(const float in[4], float out[4]) {
    const vec4 r = ref_rchit::sRGB(vec4(in[0], in[1], in[2], in[3]));
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}
void ref_rchit_get_base_color(const void* material, const float uv[2], const void* samplers, float out[4]) {
    ref_rchit::textures = static_cast<const sampler2D*>(samplers);
    const vec4 r = ref_rchit::getBaseColor(*static_cast<const ref_rchit::Material*>(material), vec2(uv[0], uv[1]));
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}
void ref_rchit_set_light(const float p[3]) { ref_rchit::lightPos = vec3(p[0], p[1], p[2]); }
void ref_rchit_get_light(float p[3]) { p[0] = ref_rchit::lightPos.x; p[1] = ref_rchit::lightPos.y; p[2] = ref_rchit::lightPos.z; }
void ref_rchit_set_in_shadow(int v) { ref_rchit::inShadow = v != 0; }
// One closest-hit invocation. w2o: the 4 columns of gl_WorldToObjectEXT (3 floats each).
void ref_rchit_run(const void* instance_infos, const void* materials, const void* samplers, int instance, int primitive,
                   const float bary_uv[2], const float origin[3], const float dir[3], float t, const float w2o[12], float color_out[3]) {
    using namespace ref_rchit;
    instanceInfo.i = static_cast<const InstanceInfo*>(instance_infos);
    ref_rchit::materials.m = static_cast<const Material*>(materials);
    textures = static_cast<const sampler2D*>(samplers);
    gl_InstanceCustomIndexEXT = instance; gl_PrimitiveID = primitive;
    attribs = vec3(bary_uv[0], bary_uv[1], 0.0f);
    gl_WorldRayOriginEXT = vec3(origin[0], origin[1], origin[2]);
    gl_WorldRayDirectionEXT = vec3(dir[0], dir[1], dir[2]);
    gl_HitTEXT = t;
    for (int c = 0; c < 4; ++c) gl_WorldToObjectEXT.c[c] = vec3(w2o[3 * c], w2o[3 * c + 1], w2o[3 * c + 2]);
    shader_main();
    color_out[0] = color.x; color_out[1] = color.y; color_out[2] = color.z;
}
}

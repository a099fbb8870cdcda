// shaders/main.rchit compiled from the reference's text (oracle/make_ref_shaders.py). TEST INFRASTRUCTURE.
// The viewer's closest-hit shader: direct lighting as env_map.rchit plus the run-time gather over the baked probes
// (main.rchit:124-167) -- the operator the multi-bounce passes iterate.
#include "glsl_shim.h"
#define GLUE_DECLS "main_rchit_decls.inc"
namespace glsl { namespace ref_main_rchit {
#include "main.rchit.inc"
static_assert(sizeof(InstanceInfo) == 24 && sizeof(Vertex) == 44 && sizeof(Material) == 144, "structures.h scalar layout");
}}
using namespace glsl;
extern "C" {
void ref_main_rchit_set_constants(const float grid_step[3], unsigned lmax, const float light[3], float shadow_bias, float ambient,
                                  float c_diffuse, float c_specular, float c_gloss) {
    using namespace ref_main_rchit;
    constants.gridStep = vec3(grid_step[0], grid_step[1], grid_step[2]); constants.lmax = lmax;
    constants.lightPosition = vec3(light[0], light[1], light[2]); constants.shadowBias = shadow_bias; constants.ambient = ambient;
    constants.Cdiffuse = c_diffuse; constants.Cspecular = c_specular; constants.Cglossyness = c_gloss;
}
void ref_main_rchit_set_in_shadow(int v) { ref_main_rchit::inShadow = v != 0; }
// payload 2 (the probe-visibility ray): read for / written back by the miss shader the pipeline runs (sh.rmiss)
void ref_main_rchit_get_sh_payload(float sum[3], float normal[3], int ijk[3], unsigned* lmax, int* occluded) {
    const ref_main_rchit::SHPayload& p = ref_main_rchit::shPayload;
    sum[0] = p.sum.x; sum[1] = p.sum.y; sum[2] = p.sum.z; normal[0] = p.normal.x; normal[1] = p.normal.y; normal[2] = p.normal.z;
    ijk[0] = p.ijk.x; ijk[1] = p.ijk.y; ijk[2] = p.ijk.z; *lmax = p.lmax; *occluded = p.occluded ? 1 : 0;
}
void ref_main_rchit_set_sh_payload(const float sum[3], int occluded) {
    ref_main_rchit::shPayload.sum = vec3(sum[0], sum[1], sum[2]);
    ref_main_rchit::shPayload.occluded = occluded != 0;
}
// One closest-hit invocation. w2o: the 4 columns of gl_WorldToObjectEXT (3 floats each).
void ref_main_rchit_run(const void* instance_infos, const void* materials, const void* samplers, int instance, int primitive,
                        const float bary_uv[2], const float origin[3], const float dir[3], float t, const float w2o[12], float color_out[3]) {
    using namespace ref_main_rchit;
    instanceInfo.i = static_cast<const InstanceInfo*>(instance_infos);
    ref_main_rchit::materials.m = static_cast<const Material*>(materials);
    textures = static_cast<const sampler2D*>(samplers);
    gl_InstanceCustomIndexEXT = instance; gl_PrimitiveID = primitive;
    attribs = vec3(bary_uv[0], bary_uv[1], 0.0f);
    gl_WorldRayOriginEXT = vec3(origin[0], origin[1], origin[2]);
    gl_WorldRayDirectionEXT = vec3(dir[0], dir[1], dir[2]);
    gl_HitTEXT = t;
    for (int c = 0; c < 4; ++c) gl_WorldToObjectEXT.c[c] = vec3(w2o[3 * c], w2o[3 * c + 1], w2o[3 * c + 2]);
    skyboxRadiance = vec3(0.0f);
    shader_main();
    color_out[0] = color.x; color_out[1] = color.y; color_out[2] = color.z;
}
}

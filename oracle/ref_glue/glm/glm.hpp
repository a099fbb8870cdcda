// Stand-in for <glm/glm.hpp>, which the reference's shaders/structures.h includes on its C++ side (inside
// `namespace shader`): its vec types are the GLSL shim's. TEST INFRASTRUCTURE (oracle/glsl_shim.h).
namespace glm {
using vec4 = ::glsl::vec4;
using vec3 = ::glsl::vec3;
using vec2 = ::glsl::vec2;
}

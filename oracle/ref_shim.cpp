// ref_shim.cpp — compiles the REFERENCE's own shaders/sh_common.h (unmodified, included from
// where it lies under /root/reference) as C++ behind a minimal GLSL shim, and exports its
// functions with C linkage. Output goes to oracle/_ref/ only (git-ignored). TEST INFRASTRUCTURE:
// used to pin oracle/vlb_oracle.cpp's restated SH basis / equirect maps and to generate
// tests/golden/sh_common_golden.npz (tests/golden/make_golden.py).
// Build flags (oracle/build_oracle.py): -fsingle-precision-constant so that unsuffixed literals
// are fp32 as in GLSL, -ffp-contract=off so no FMA contraction changes the arithmetic.
#include <cmath>

struct vec3 {
    float x, y, z;
    vec3(float a, float b, float c) : x(a), y(b), z(c) {}
};
static inline vec3 normalize(const vec3 v) {
    const float l = std::sqrt(v.x * v.x + v.y * v.y + v.z * v.z);
    return vec3(v.x / l, v.y / l, v.z / l);
}
using std::cos;
using std::sin;

#include "sh_common.h"  // -I /root/reference/shaders

extern "C" {
float ref_SH(int l, int m, float x, float y, float z) { return SH(l, m, vec3(x, y, z)); }
float ref_x2phi(int x, int w) { return x2phi(x, w); }
float ref_y2theta(int y, int h) { return y2theta(y, h); }
void ref_toVector(float phi, float theta, float* out3) {
    const vec3 v = toVector(phi, theta);
    out3[0] = v.x; out3[1] = v.y; out3[2] = v.z;
}
float ref_PI(void) { return PI; }
float ref_calcNormalizationConst(float h, float w) { return calcNormalizationConst(h, w); }
}

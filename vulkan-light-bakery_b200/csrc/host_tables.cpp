// host_tables.cpp — small host-side tables uploaded once per call: probe axis coordinates,
// equirect sin/cos tables and the per-row quadrature weights. O(W + H + N) work; everything
// that scales with texels, rays or triangles runs on the GPU.
#include <cmath>
#include <cstddef>

#include "vlb_context.h"

namespace vlb {

static const float kPiF = 3.1415926538f;  // shaders/sh_common.h:1

// shaders/sh_common.h:14-22 (fp32, same operation order as GLSL evaluates it)
static inline float x2phi(int x, int w) { return 2.f * kPiF * ((float)x + 0.5f) / (float)w; }
static inline float y2theta(int y, int h) { return kPiF * ((float)y + 0.5f) / (float)h; }
// sin/cos: GLSL precision is implementation-defined; evaluate in double and round once.
static inline float sin_f(float x) { return (float)std::sin((double)x); }
static inline float cos_f(float x) { return (float)std::cos((double)x); }

// LightBaker::probePositionsFromBoudingBox (src/baker/light_baker.cpp:92-96): coordinate i of an
// axis is the bounds minimum with gridStep added i times in fp32.
void host_axis_coords(float origin, float step, int n, float* out) {
    volatile float p = origin;
    for (int i = 0; i < n; ++i) { out[i] = p; p = p + step; }
}

void host_dir_tables(int W, int H, float phi_shift, float* row_sc, float* col_cs) {
    for (int y = 0; y < H; ++y) {
        const float th = y2theta(y, H);
        row_sc[2 * y + 0] = sin_f(th);
        row_sc[2 * y + 1] = cos_f(th);
    }
    for (int x = 0; x < W; ++x) {
        volatile float ph = x2phi(x, W);
        if (phi_shift != 0.f) ph = ph - phi_shift;   // shaders/skybox_sh.comp:28: - PI / 2.0f
        col_cs[2 * x + 0] = cos_f(ph);
        col_cs[2 * x + 1] = sin_f(ph);
    }
}

// Per-row factors of the separable projection (skybox_sh.cu): with S = sin(theta), C = cos(theta)
// and w = (2 PI / W)(PI / H) S (shaders/sh.comp:32-33): {w, wS, wS^2, wC, wSC, wS^3, wS^2C, 0}.
void host_proj_row_table(int W, int H, float* row_tab) {
    const float pixel_area = (2.0f * kPiF / (float)W) * (kPiF / (float)H);
    for (int y = 0; y < H; ++y) {
        const float th = y2theta(y, H);
        const double S = (double)sin_f(th), C = (double)cos_f(th);
        const double w = (double)(pixel_area * sin_f(th));
        float* r = row_tab + 8 * (size_t)y;
        r[0] = (float)w;
        r[1] = (float)(w * S);
        r[2] = (float)(w * S * S);
        r[3] = (float)(w * C);
        r[4] = (float)(w * S * C);
        r[5] = (float)(w * S * S * S);
        r[6] = (float)(w * S * S * C);
        r[7] = 0.f;
    }
}

void host_inverse3x3(const float* m, float* nm) {
    const double a = m[0], b = m[1], c = m[2], d = m[4], e = m[5], f = m[6], g = m[8], h = m[9], i = m[10];
    const double A = e * i - f * h, B = -(d * i - f * g), C = d * h - e * g;
    const double det = a * A + b * B + c * C;
    const double id = det != 0.0 ? 1.0 / det : 0.0;
    nm[0] = (float)(A * id); nm[1] = (float)(-(b * i - c * h) * id); nm[2] = (float)((b * f - c * e) * id);
    nm[3] = (float)(B * id); nm[4] = (float)((a * i - c * g) * id);  nm[5] = (float)(-(a * f - c * d) * id);
    nm[6] = (float)(C * id); nm[7] = (float)(-(a * h - b * g) * id); nm[8] = (float)((a * e - b * d) * id);
}

// Output slot of grid cell (i,j,k) in the order LightBaker::probePositionsFromBoudingBox pushes
// positions (src/baker/light_baker.cpp:87-98).
size_t ref_order_index(int i, int j, int k, int Nx, int Ny, int Nz) {
    const size_t q = (j == 0) ? (size_t)i : (size_t)Nx + (size_t)i * (size_t)(Ny - 1) + (size_t)(j - 1);
    return (k == 0) ? q : (size_t)Nx * Ny + q * (size_t)(Nz - 1) + (size_t)(k - 1);
}

}  // namespace vlb

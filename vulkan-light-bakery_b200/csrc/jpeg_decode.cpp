// jpeg_decode.cpp — baseline JPEG reader for glTF baseColor textures; with png_decode.cpp it replaces the
// stb_image decode tinygltf performs for the reference (Scene_t::loadTextures, src/scene_manager.cpp:941-973).
// Host-side ingest, off the bake path. Handles what glTF exporters write: 8-bit baseline / extended-sequential
// Huffman JPEG (SOF0 / SOF1), greyscale or YCbCr, sampling 4:4:4 / 4:2:2 / 4:4:0 / 4:2:0 (chroma upsampled with
// the triangle filter libjpeg calls "fancy upsampling"), restart intervals. Progressive, arithmetic-coded,
// lossless, 12-bit and CMYK files are reported as unsupported, never guessed.
// "Parity unpinned": JPEG decoders legitimately differ by a level or two (IDCT and upsampling arithmetic are not
// normative); this one uses an exact double-precision IDCT, and tests compare it with libjpeg within that band.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace vlb {

namespace {

struct Unsup : std::runtime_error { using std::runtime_error::runtime_error; };

struct HuffTable {
    bool present = false;
    uint8_t counts[17] = {0};
    uint8_t symbols[256] = {0};
    int mincode[17], maxcode[18], valptr[17];
    void build() {
        int code = 0, k = 0;
        for (int len = 1; len <= 16; ++len) {
            valptr[len] = k;
            mincode[len] = code;
            code += counts[len];
            k += counts[len];
            maxcode[len] = counts[len] ? code - 1 : -1;
            code <<= 1;
        }
        maxcode[17] = 0x7fffffff;
    }
};

struct Component {
    int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0;
    int bw = 0, bh = 0;              // blocks per row / column of this component's (padded) plane
    std::vector<uint8_t> plane;      // bw*8 x bh*8 samples
    int pred = 0;
};

struct BitReader {
    const uint8_t* p; const uint8_t* end;
    uint32_t acc = 0; int n = 0;
    bool hit_marker = false;
    int padded = 0;                  // zero bytes fed past a marker / the end of the data since the last reset
    void fill() {
        while (n <= 24) {
            int byte = 0;
            if (!hit_marker && p < end) {
                byte = *p;
                if (byte == 0xFF) {
                    if (p + 1 < end && p[1] == 0x00) p += 2;            // stuffed zero
                    else { hit_marker = true; byte = 0; }               // a marker: feed zeros from here on
                } else ++p;
            } else {
                hit_marker = true;
            }
            if (hit_marker && ++padded > 16) throw std::runtime_error("JPEG: entropy-coded data ends early");
            acc |= (uint32_t)byte << (24 - n);
            n += 8;
        }
    }
    int bit() { if (n == 0) fill(); const int b = (acc >> 31) & 1; acc <<= 1; --n; return b; }
    int bits(int k) { int v = 0; for (int i = 0; i < k; ++i) v = (v << 1) | bit(); return v; }
    void reset() { acc = 0; n = 0; hit_marker = false; padded = 0; }
};

int decode_symbol(BitReader& br, const HuffTable& t) {
    int code = 0;
    for (int len = 1; len <= 16; ++len) {
        code = (code << 1) | br.bit();
        if (t.maxcode[len] >= 0 && code <= t.maxcode[len] && code >= t.mincode[len]) return t.symbols[t.valptr[len] + code - t.mincode[len]];
    }
    throw std::runtime_error("JPEG: bad Huffman code");
}

int extend(int v, int bits) { return bits == 0 ? 0 : (v < (1 << (bits - 1)) ? v - (1 << bits) + 1 : v); }

const int kZigzag[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
                         35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

void idct8x8(const int* coef, uint8_t* out, int stride) {
    static double c[8][8];
    static bool init = false;
    if (!init) {
        for (int x = 0; x < 8; ++x)
            for (int u = 0; u < 8; ++u) c[x][u] = (u == 0 ? std::sqrt(0.125) : 0.5) * std::cos((2 * x + 1) * u * M_PI / 16.0);
        init = true;
    }
    double tmp[64];
    for (int v = 0; v < 8; ++v)
        for (int x = 0; x < 8; ++x) {
            double s = 0;
            for (int u = 0; u < 8; ++u) s += c[x][u] * coef[v * 8 + u];
            tmp[v * 8 + x] = s;
        }
    for (int x = 0; x < 8; ++x)
        for (int y = 0; y < 8; ++y) {
            double s = 0;
            for (int v = 0; v < 8; ++v) s += c[y][v] * tmp[v * 8 + x];
            const long r = std::lround(s + 128.0);
            out[y * stride + x] = (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r));
        }
}

// libjpeg's "fancy" 2x upsampling along one axis of a plane region: out[2i] = (3 in[i] + in[i-1] + 1) / 4,
// out[2i+1] = (3 in[i] + in[i+1] + 2) / 4, the outermost samples copied.
void upsample2_h(const std::vector<uint8_t>& in, int w, int h, int in_stride, std::vector<uint8_t>& out) {
    out.assign((size_t)2 * w * h, 0);
    for (int y = 0; y < h; ++y) {
        const uint8_t* r = &in[(size_t)y * in_stride];
        uint8_t* o = &out[(size_t)y * 2 * w];
        for (int x = 0; x < w; ++x) {
            const int c3 = 3 * r[x];
            o[2 * x] = x == 0 ? r[0] : (uint8_t)((c3 + r[x - 1] + 1) >> 2);
            o[2 * x + 1] = x == w - 1 ? r[w - 1] : (uint8_t)((c3 + r[x + 1] + 2) >> 2);
        }
    }
}
void upsample2_v(const std::vector<uint8_t>& in, int w, int h, int in_stride, std::vector<uint8_t>& out) {
    out.assign((size_t)w * 2 * h, 0);
    for (int y = 0; y < h; ++y) {
        const uint8_t* r = &in[(size_t)y * in_stride];
        const uint8_t* up = &in[(size_t)(y == 0 ? 0 : y - 1) * in_stride];
        const uint8_t* dn = &in[(size_t)(y == h - 1 ? h - 1 : y + 1) * in_stride];
        for (int x = 0; x < w; ++x) {
            out[(size_t)(2 * y) * w + x] = (uint8_t)((3 * r[x] + up[x] + 1) >> 2);
            out[(size_t)(2 * y + 1) * w + x] = (uint8_t)((3 * r[x] + dn[x] + 2) >> 2);
        }
    }
}

uint8_t clamp8(double v) { const long r = std::lround(v); return (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r)); }

}  // namespace

// Decodes `data` into RGBA8 (row 0 first, alpha 255). Throws std::runtime_error; `unsupported` is set when the
// file is a JPEG variant this reader does not handle.
void jpeg_decode_rgba8(const uint8_t* data, size_t size, std::vector<uint8_t>& rgba, int& width, int& height, bool& unsupported) {
    unsupported = false;
    try {
        if (size < 4 || data[0] != 0xFF || data[1] != 0xD8) throw std::runtime_error("JPEG: missing SOI");
        uint16_t qt[4][64];
        bool have_qt[4] = {false, false, false, false};
        HuffTable dc[4], ac[4];
        std::vector<Component> comps;
        int W = 0, H = 0, hmax = 1, vmax = 1, restart = 0;
        bool have_sof = false, adobe = false;
        int adobe_transform = -1;
        size_t pos = 2;
        bool done = false;
        while (!done) {
            if (pos + 4 > size) throw std::runtime_error("JPEG: truncated before the scan");
            if (data[pos] != 0xFF) throw std::runtime_error("JPEG: marker expected");
            while (pos < size && data[pos] == 0xFF) ++pos;           // fill bytes
            const int m = data[pos++];
            if (m == 0xD8 || (m >= 0xD0 && m <= 0xD7) || m == 0x01) continue;
            if (m == 0xD9) throw std::runtime_error("JPEG: EOI before any scan");
            if (pos + 2 > size) throw std::runtime_error("JPEG: truncated segment");
            const size_t len = ((size_t)data[pos] << 8) | data[pos + 1];
            if (len < 2 || pos + len > size) throw std::runtime_error("JPEG: truncated segment");
            const uint8_t* s = data + pos + 2;
            const size_t n = len - 2;
            if (m == 0xDB) {                                          // DQT
                size_t i = 0;
                while (i < n) {
                    const int pq = s[i] >> 4, tq = s[i] & 15;
                    ++i;
                    if (tq > 3 || i + (pq ? 128 : 64) > n) throw std::runtime_error("JPEG: bad DQT");
                    for (int k = 0; k < 64; ++k) {
                        qt[tq][kZigzag[k]] = pq ? (uint16_t)((s[i] << 8) | s[i + 1]) : s[i];
                        i += pq ? 2 : 1;
                    }
                    have_qt[tq] = true;
                }
            } else if (m == 0xC4) {                                   // DHT
                size_t i = 0;
                while (i < n) {
                    if (i + 17 > n) throw std::runtime_error("JPEG: bad DHT");
                    const int tc = s[i] >> 4, th = s[i] & 15;
                    if (tc > 1 || th > 3) throw std::runtime_error("JPEG: bad DHT id");
                    HuffTable& t = tc ? ac[th] : dc[th];
                    int total = 0;
                    for (int k = 1; k <= 16; ++k) { t.counts[k] = s[i + k]; total += t.counts[k]; }
                    i += 17;
                    if (total > 256 || i + total > n) throw std::runtime_error("JPEG: bad DHT");
                    std::memcpy(t.symbols, s + i, total);
                    i += total;
                    t.present = true;
                    t.build();
                }
            } else if (m == 0xC0 || m == 0xC1) {                      // SOF0 / SOF1
                if (n < 6) throw std::runtime_error("JPEG: bad SOF");
                if (s[0] != 8) throw Unsup("JPEG: only 8-bit samples are decoded");
                H = (s[1] << 8) | s[2]; W = (s[3] << 8) | s[4];
                const int nc = s[5];
                if (W <= 0 || H <= 0 || W > 32768 || H > 32768) throw std::runtime_error("JPEG: bad size");
                if (nc != 1 && nc != 3) throw Unsup("JPEG: only greyscale and YCbCr images are decoded (" + std::to_string(nc) + " components)");
                if (n < (size_t)(6 + 3 * nc)) throw std::runtime_error("JPEG: bad SOF");
                comps.assign(nc, Component());
                for (int c = 0; c < nc; ++c) {
                    comps[c].id = s[6 + 3 * c]; comps[c].h = s[7 + 3 * c] >> 4; comps[c].v = s[7 + 3 * c] & 15; comps[c].tq = s[8 + 3 * c];
                    if (comps[c].h < 1 || comps[c].h > 2 || comps[c].v < 1 || comps[c].v > 2 || comps[c].tq > 3)
                        throw Unsup("JPEG: sampling factors other than 1 and 2 are not decoded");
                    hmax = std::max(hmax, comps[c].h); vmax = std::max(vmax, comps[c].v);
                }
                if (nc == 3 && (comps[1].h != 1 || comps[1].v != 1 || comps[2].h != 1 || comps[2].v != 1))
                    throw Unsup("JPEG: subsampled luma / oversampled chroma layouts are not decoded");
                have_sof = true;
            } else if (m == 0xC2 || m == 0xC3 || (m >= 0xC5 && m <= 0xCF && m != 0xC8)) {
                throw Unsup("JPEG: progressive / lossless / arithmetic-coded files are not decoded (SOF marker 0x" +
                            std::string(1, "0123456789ABCDEF"[m >> 4]) + std::string(1, "0123456789ABCDEF"[m & 15]) + ")");
            } else if (m == 0xDD) {                                   // DRI
                if (n < 2) throw std::runtime_error("JPEG: bad DRI");
                restart = (s[0] << 8) | s[1];
            } else if (m == 0xEE && n >= 12 && !std::memcmp(s, "Adobe", 5)) {
                adobe = true; adobe_transform = s[11];
            } else if (m == 0xDA) {                                   // SOS
                if (!have_sof) throw std::runtime_error("JPEG: SOS before SOF");
                const int ns = s[0];
                if (ns != (int)comps.size() || n < (size_t)(4 + 2 * ns)) throw Unsup("JPEG: non-interleaved multi-scan files are not decoded");
                for (int k = 0; k < ns; ++k) {
                    const int cid = s[1 + 2 * k];
                    Component* c = nullptr;
                    for (Component& cc : comps) if (cc.id == cid) c = &cc;
                    if (!c) throw std::runtime_error("JPEG: scan names an unknown component");
                    c->td = s[2 + 2 * k] >> 4; c->ta = s[2 + 2 * k] & 15;
                    if (c->td > 3 || c->ta > 3 || !dc[c->td].present || !ac[c->ta].present || !have_qt[c->tq]) throw std::runtime_error("JPEG: scan uses a missing table");
                }
                pos += len;
                done = true;
                continue;
            }
            pos += len;
        }
        if (adobe && comps.size() == 3 && adobe_transform == 0) throw Unsup("JPEG: Adobe RGB (untransformed) files are not decoded");

        // ---- entropy-coded data: interleaved MCUs ----
        const int mcu_w = 8 * hmax, mcu_h = 8 * vmax;
        const int mcus_x = (W + mcu_w - 1) / mcu_w, mcus_y = (H + mcu_h - 1) / mcu_h;
        for (Component& c : comps) {
            c.bw = mcus_x * c.h; c.bh = mcus_y * c.v;
            c.plane.assign((size_t)c.bw * 8 * c.bh * 8, 0);
        }
        BitReader br{data + pos, data + size};
        int coef[64];
        int until_restart = restart;
        for (int my = 0; my < mcus_y; ++my)
            for (int mx = 0; mx < mcus_x; ++mx) {
                if (restart && until_restart == 0) {                 // RSTn: byte-align, skip the marker, reset predictors
                    const uint8_t* q = br.p;
                    while (q + 1 < br.end && !(q[0] == 0xFF && q[1] >= 0xD0 && q[1] <= 0xD7)) ++q;
                    if (q + 1 >= br.end) throw std::runtime_error("JPEG: restart marker missing");
                    br.p = q + 2;
                    br.reset();
                    for (Component& c : comps) c.pred = 0;
                    until_restart = restart;
                }
                for (Component& c : comps)
                    for (int by = 0; by < c.v; ++by)
                        for (int bx = 0; bx < c.h; ++bx) {
                            std::memset(coef, 0, sizeof coef);
                            const int t = decode_symbol(br, dc[c.td]);
                            if (t > 11) throw std::runtime_error("JPEG: bad DC size");
                            c.pred += extend(br.bits(t), t);
                            coef[0] = c.pred * qt[c.tq][0];
                            for (int k = 1; k < 64;) {
                                const int rs = decode_symbol(br, ac[c.ta]);
                                const int r = rs >> 4, sz = rs & 15;
                                if (sz == 0) {
                                    if (r != 15) break;              // EOB
                                    k += 16;
                                    continue;
                                }
                                k += r;
                                if (k > 63) throw std::runtime_error("JPEG: AC run past the block");
                                coef[kZigzag[k]] = extend(br.bits(sz), sz) * qt[c.tq][kZigzag[k]];
                                ++k;
                            }
                            const int px = (mx * c.h + bx) * 8, py = (my * c.v + by) * 8;
                            idct8x8(coef, &c.plane[(size_t)py * c.bw * 8 + px], c.bw * 8);
                        }
                --until_restart;
            }

        // ---- chroma upsampling + colour conversion ----
        width = W; height = H;
        rgba.assign((size_t)W * H * 4, 255);
        if (comps.size() == 1) {
            const Component& y = comps[0];
            for (int j = 0; j < H; ++j)
                for (int i = 0; i < W; ++i) {
                    const uint8_t v = y.plane[(size_t)j * y.bw * 8 + i];
                    uint8_t* o = &rgba[((size_t)j * W + i) * 4];
                    o[0] = o[1] = o[2] = v;
                }
            return;
        }
        const Component& Y = comps[0];
        std::vector<uint8_t> up[2];
        int up_stride[2];
        for (int k = 0; k < 2; ++k) {
            const Component& c = comps[1 + k];
            // the part of the chroma plane that covers the image: ceil(W / hmax) x ceil(H / vmax) samples
            const int cw = (W + hmax - 1) / hmax, chh = (H + vmax - 1) / vmax;
            std::vector<uint8_t> cur(c.plane);
            int w = cw, h = chh, stride = c.bw * 8;
            if (hmax == 2) { std::vector<uint8_t> t; upsample2_h(cur, w, h, stride, t); cur.swap(t); w *= 2; stride = w; }
            if (vmax == 2) { std::vector<uint8_t> t; upsample2_v(cur, w, h, stride, t); cur.swap(t); h *= 2; stride = w; }
            up[k].swap(cur); up_stride[k] = stride;
        }
        for (int j = 0; j < H; ++j)
            for (int i = 0; i < W; ++i) {
                const double y = Y.plane[(size_t)j * Y.bw * 8 + i];
                const double cb = up[0][(size_t)j * up_stride[0] + i] - 128.0, cr = up[1][(size_t)j * up_stride[1] + i] - 128.0;
                uint8_t* o = &rgba[((size_t)j * W + i) * 4];
                o[0] = clamp8(y + 1.402 * cr);
                o[1] = clamp8(y - 0.344136 * cb - 0.714136 * cr);
                o[2] = clamp8(y + 1.772 * cb);
            }
    } catch (const Unsup&) {
        unsupported = true;
        throw;
    }
}

}  // namespace vlb

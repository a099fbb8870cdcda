// jpeg_decode.cpp — baseline + progressive JPEG reader for glTF baseColor textures; with png_decode.cpp it replaces the
// stb_image decode tinygltf performs for the reference (Scene_t::loadTextures, src/scene_manager.cpp:941-973).
// Host-side ingest, off the bake path. Handles what glTF exporters write: 8-bit baseline / extended-sequential
// Huffman JPEG (SOF0 / SOF1), greyscale or YCbCr, sampling 4:4:4 / 4:2:2 / 4:4:0 / 4:2:0 (chroma upsampled with
// the triangle filter libjpeg calls "fancy upsampling"), restart intervals, and progressive files (SOF2: spectral
// selection + successive approximation, any scan script). Arithmetic-coded, lossless, hierarchical, 12-bit and
// CMYK files are reported as unsupported, never guessed.
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

// One scan of a (possibly progressive) JPEG: which components, which band of coefficients, which bit.
struct Scan {
    std::vector<Component*> comps;
    int ss = 0, se = 63, ah = 0, al = 0;
};

struct Decoder {
    const uint8_t* data; size_t size;
    uint16_t qt[4][64];
    bool have_qt[4] = {false, false, false, false};
    HuffTable dc[4], ac[4];
    std::vector<Component> comps;
    std::vector<std::vector<int16_t>> coefs;      // per component: bw * bh blocks x 64 coefficients, natural order
    int W = 0, H = 0, hmax = 1, vmax = 1, restart = 0, mcus_x = 0, mcus_y = 0;
    bool progressive = false;
    int eobrun = 0;

    int16_t* block(const Component& c, int bx, int by) { return &coefs[&c - comps.data()][((size_t)by * c.bw + bx) * 64]; }

    void decode_block(BitReader& br, Component& c, int16_t* coef, const Scan& sc) {
        if (!progressive) {                                              // sequential: the whole block at once
            const int t = decode_symbol(br, dc[c.td]);
            if (t > 11) throw std::runtime_error("JPEG: bad DC size");
            c.pred += extend(br.bits(t), t);
            coef[0] = (int16_t)c.pred;
            for (int k = 1; k < 64;) {
                const int rs = decode_symbol(br, ac[c.ta]);
                const int r = rs >> 4, sz = rs & 15;
                if (sz == 0) {
                    if (r != 15) break;                                  // EOB
                    k += 16;
                    continue;
                }
                k += r;
                if (k > 63) throw std::runtime_error("JPEG: AC run past the block");
                coef[kZigzag[k]] = (int16_t)extend(br.bits(sz), sz);
                ++k;
            }
            return;
        }
        if (sc.ss == 0) {                                                // progressive DC scan
            if (sc.ah == 0) {
                const int t = decode_symbol(br, dc[c.td]);
                if (t > 11) throw std::runtime_error("JPEG: bad DC size");
                c.pred += extend(br.bits(t), t);
                coef[0] = (int16_t)(c.pred * (1 << sc.al));
            } else if (br.bit()) {
                coef[0] = (int16_t)(coef[0] | (1 << sc.al));
            }
            return;
        }
        const int p1 = 1 << sc.al, m1 = -(1 << sc.al);
        if (sc.ah == 0) {                                                // progressive AC, first pass over this band
            if (eobrun > 0) { --eobrun; return; }
            for (int k = sc.ss; k <= sc.se;) {
                const int rs = decode_symbol(br, ac[c.ta]);
                const int r = rs >> 4, sz = rs & 15;
                if (sz == 0) {
                    if (r < 15) {
                        eobrun = (1 << r) - 1;
                        if (r) eobrun += br.bits(r);
                        break;
                    }
                    k += 16;
                } else {
                    k += r;
                    if (k > sc.se) throw std::runtime_error("JPEG: AC run past the band");
                    coef[kZigzag[k]] = (int16_t)(extend(br.bits(sz), sz) * p1);
                    ++k;
                }
            }
            return;
        }
        // progressive AC refinement: one more bit for the coefficients that are already non-zero, new +-1 ones
        int k = sc.ss;
        if (eobrun <= 0) {
            for (; k <= sc.se; ++k) {
                const int rs = decode_symbol(br, ac[c.ta]);
                int r = rs >> 4;
                const int sz = rs & 15;
                int val = 0;
                if (sz == 0) {
                    if (r < 15) {
                        eobrun = 1 << r;
                        if (r) eobrun += br.bits(r);
                        break;
                    }
                } else {
                    if (sz != 1) throw std::runtime_error("JPEG: bad refinement symbol");
                    val = br.bit() ? p1 : m1;
                }
                while (k <= sc.se) {
                    int16_t& cf = coef[kZigzag[k]];
                    if (cf != 0) {
                        if (br.bit() && (cf & p1) == 0) cf = (int16_t)(cf + (cf > 0 ? p1 : m1));
                    } else {
                        if (r == 0) { if (val) cf = (int16_t)val; break; }
                        --r;
                    }
                    ++k;
                }
            }
        }
        if (eobrun > 0) {
            for (; k <= sc.se; ++k) {
                int16_t& cf = coef[kZigzag[k]];
                if (cf != 0 && br.bit() && (cf & p1) == 0) cf = (int16_t)(cf + (cf > 0 ? p1 : m1));
            }
            --eobrun;
        }
    }

    // Entropy-coded segment of one scan, starting at `pos`; returns the position of the marker that ends it.
    size_t decode_scan(size_t pos, const Scan& sc) {
        BitReader br{data + pos, data + size};
        for (Component* c : sc.comps) c->pred = 0;
        eobrun = 0;
        const bool interleaved = sc.comps.size() > 1;
        int units_x = mcus_x, units_y = mcus_y;
        if (!interleaved) {                                              // one block per "MCU", over the component's own size
            const Component& c = *sc.comps[0];
            units_x = ((W * c.h + hmax - 1) / hmax + 7) / 8;
            units_y = ((H * c.v + vmax - 1) / vmax + 7) / 8;
        }
        int until_restart = restart;
        for (int uy = 0; uy < units_y; ++uy)
            for (int ux = 0; ux < units_x; ++ux) {
                if (restart && until_restart == 0) {                     // RSTn: byte-align, skip the marker, reset the predictors
                    const uint8_t* q = br.p;
                    while (q + 1 < br.end && !(q[0] == 0xFF && q[1] >= 0xD0 && q[1] <= 0xD7)) ++q;
                    if (q + 1 >= br.end) throw std::runtime_error("JPEG: restart marker missing");
                    br.p = q + 2;
                    br.reset();
                    for (Component* c : sc.comps) c->pred = 0;
                    eobrun = 0;
                    until_restart = restart;
                }
                if (interleaved) {
                    for (Component* c : sc.comps)
                        for (int by = 0; by < c->v; ++by)
                            for (int bx = 0; bx < c->h; ++bx) decode_block(br, *c, block(*c, ux * c->h + bx, uy * c->v + by), sc);
                } else {
                    decode_block(br, *sc.comps[0], block(*sc.comps[0], ux, uy), sc);
                }
                --until_restart;
            }
        const uint8_t* q = br.p;                                          // the next marker (not a stuffed zero, not RSTn)
        while (q + 1 < data + size && !(q[0] == 0xFF && q[1] != 0x00 && q[1] != 0xFF && !(q[1] >= 0xD0 && q[1] <= 0xD7))) ++q;
        return (size_t)(q - data);
    }
};

// Decodes `data` into RGBA8 (row 0 first, alpha 255). Throws std::runtime_error; `unsupported` is set when the
// file is a JPEG variant this reader does not handle.
void jpeg_decode_rgba8(const uint8_t* data, size_t size, std::vector<uint8_t>& rgba, int& width, int& height, bool& unsupported) {
    unsupported = false;
    try {
        if (size < 4 || data[0] != 0xFF || data[1] != 0xD8) throw std::runtime_error("JPEG: missing SOI");
        Decoder d;
        d.data = data; d.size = size;
        std::vector<Component>& comps = d.comps;
        bool have_sof = false, adobe = false;
        int adobe_transform = -1, n_scans = 0;
        size_t pos = 2;
        bool done = false;
        while (!done) {
            if (pos + 2 > size) { if (n_scans) break; throw std::runtime_error("JPEG: truncated before the scan"); }
            if (data[pos] != 0xFF) throw std::runtime_error("JPEG: marker expected");
            while (pos < size && data[pos] == 0xFF) ++pos;           // fill bytes
            if (pos >= size) { if (n_scans) break; throw std::runtime_error("JPEG: truncated before the scan"); }
            const int m = data[pos++];
            if (m == 0xD8 || (m >= 0xD0 && m <= 0xD7) || m == 0x01) continue;
            if (m == 0xD9) { if (n_scans) break; throw std::runtime_error("JPEG: EOI before any scan"); }
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
                        d.qt[tq][kZigzag[k]] = pq ? (uint16_t)((s[i] << 8) | s[i + 1]) : s[i];
                        i += pq ? 2 : 1;
                    }
                    d.have_qt[tq] = true;
                }
            } else if (m == 0xC4) {                                   // DHT
                size_t i = 0;
                while (i < n) {
                    if (i + 17 > n) throw std::runtime_error("JPEG: bad DHT");
                    const int tc = s[i] >> 4, th = s[i] & 15;
                    if (tc > 1 || th > 3) throw std::runtime_error("JPEG: bad DHT id");
                    HuffTable& t = tc ? d.ac[th] : d.dc[th];
                    int total = 0;
                    for (int k = 1; k <= 16; ++k) { t.counts[k] = s[i + k]; total += t.counts[k]; }
                    i += 17;
                    if (total > 256 || i + total > n) throw std::runtime_error("JPEG: bad DHT");
                    std::memcpy(t.symbols, s + i, total);
                    i += total;
                    t.present = true;
                    t.build();
                }
            } else if (m == 0xC0 || m == 0xC1 || m == 0xC2) {         // SOF0 / SOF1 / SOF2 (progressive)
                if (have_sof) throw std::runtime_error("JPEG: more than one frame");
                if (n < 6) throw std::runtime_error("JPEG: bad SOF");
                if (s[0] != 8) throw Unsup("JPEG: only 8-bit samples are decoded");
                d.progressive = m == 0xC2;
                d.H = (s[1] << 8) | s[2]; d.W = (s[3] << 8) | s[4];
                const int nc = s[5];
                if (d.W <= 0 || d.H <= 0 || d.W > 32768 || d.H > 32768) throw std::runtime_error("JPEG: bad size");
                if (nc != 1 && nc != 3) throw Unsup("JPEG: only greyscale and YCbCr images are decoded (" + std::to_string(nc) + " components)");
                if (n < (size_t)(6 + 3 * nc)) throw std::runtime_error("JPEG: bad SOF");
                comps.assign(nc, Component());
                for (int c = 0; c < nc; ++c) {
                    comps[c].id = s[6 + 3 * c]; comps[c].h = s[7 + 3 * c] >> 4; comps[c].v = s[7 + 3 * c] & 15; comps[c].tq = s[8 + 3 * c];
                    if (comps[c].h < 1 || comps[c].h > 2 || comps[c].v < 1 || comps[c].v > 2 || comps[c].tq > 3)
                        throw Unsup("JPEG: sampling factors other than 1 and 2 are not decoded");
                    d.hmax = std::max(d.hmax, comps[c].h); d.vmax = std::max(d.vmax, comps[c].v);
                }
                if (nc == 3 && (comps[1].h != 1 || comps[1].v != 1 || comps[2].h != 1 || comps[2].v != 1))
                    throw Unsup("JPEG: subsampled luma / oversampled chroma layouts are not decoded");
                if (nc == 1) { comps[0].h = comps[0].v = 1; d.hmax = d.vmax = 1; }   // a single component is never subsampled
                d.mcus_x = (d.W + 8 * d.hmax - 1) / (8 * d.hmax); d.mcus_y = (d.H + 8 * d.vmax - 1) / (8 * d.vmax);
                d.coefs.resize(nc);
                for (int c = 0; c < nc; ++c) {
                    comps[c].bw = d.mcus_x * comps[c].h; comps[c].bh = d.mcus_y * comps[c].v;
                    d.coefs[c].assign((size_t)comps[c].bw * comps[c].bh * 64, 0);
                }
                have_sof = true;
            } else if (m == 0xC3 || (m >= 0xC5 && m <= 0xCF && m != 0xC8)) {
                throw Unsup("JPEG: lossless / hierarchical / arithmetic-coded files are not decoded (SOF marker 0x" +
                            std::string(1, "0123456789ABCDEF"[m >> 4]) + std::string(1, "0123456789ABCDEF"[m & 15]) + ")");
            } else if (m == 0xDD) {                                   // DRI
                if (n < 2) throw std::runtime_error("JPEG: bad DRI");
                d.restart = (s[0] << 8) | s[1];
            } else if (m == 0xEE && n >= 12 && !std::memcmp(s, "Adobe", 5)) {
                adobe = true; adobe_transform = s[11];
            } else if (m == 0xDA) {                                   // SOS
                if (!have_sof) throw std::runtime_error("JPEG: SOS before SOF");
                const int ns = n ? s[0] : 0;
                if (ns < 1 || ns > (int)comps.size() || n < (size_t)(4 + 2 * ns)) throw std::runtime_error("JPEG: bad SOS");
                Scan sc;
                for (int k = 0; k < ns; ++k) {
                    const int cid = s[1 + 2 * k];
                    Component* c = nullptr;
                    for (Component& cc : comps) if (cc.id == cid) c = &cc;
                    if (!c) throw std::runtime_error("JPEG: scan names an unknown component");
                    c->td = s[2 + 2 * k] >> 4; c->ta = s[2 + 2 * k] & 15;
                    if (c->td > 3 || c->ta > 3) throw std::runtime_error("JPEG: bad table id");
                    sc.comps.push_back(c);
                }
                sc.ss = s[1 + 2 * ns]; sc.se = s[2 + 2 * ns]; sc.ah = s[3 + 2 * ns] >> 4; sc.al = s[3 + 2 * ns] & 15;
                if (!d.progressive) { sc.ss = 0; sc.se = 63; sc.ah = sc.al = 0; }
                if (sc.ss > sc.se || sc.se > 63 || sc.al > 13 || (d.progressive && sc.ss == 0 && sc.se != 0) || (sc.ss > 0 && ns != 1))
                    throw std::runtime_error("JPEG: bad progressive scan parameters");
                for (Component* c : sc.comps) {
                    const bool need_dc = !d.progressive || (sc.ss == 0 && sc.ah == 0), need_ac = !d.progressive || sc.ss > 0;
                    if ((need_dc && !d.dc[c->td].present) || (need_ac && !d.ac[c->ta].present)) throw std::runtime_error("JPEG: scan uses a missing Huffman table");
                }
                pos = d.decode_scan(pos + len, sc);
                ++n_scans;
                continue;
            }
            pos += len;
        }
        if (adobe && comps.size() == 3 && adobe_transform == 0) throw Unsup("JPEG: Adobe RGB (untransformed) files are not decoded");
        const int W = d.W, H = d.H, hmax = d.hmax, vmax = d.vmax;

        // ---- dequantise + inverse DCT of every block ----
        for (Component& c : comps) {
            if (!d.have_qt[c.tq]) throw std::runtime_error("JPEG: missing quantisation table");
            c.plane.assign((size_t)c.bw * 8 * c.bh * 8, 0);
            int coef[64];
            for (int by = 0; by < c.bh; ++by)
                for (int bx = 0; bx < c.bw; ++bx) {
                    const int16_t* src = d.block(c, bx, by);
                    for (int k = 0; k < 64; ++k) coef[k] = (int)src[k] * d.qt[c.tq][k];
                    idct8x8(coef, &c.plane[(size_t)by * 8 * c.bw * 8 + (size_t)bx * 8], c.bw * 8);
                }
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

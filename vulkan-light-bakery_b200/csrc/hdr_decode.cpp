// hdr_decode.cpp — Radiance RGBE (.hdr) reader: the usual container of RGBA32F equirect skyboxes (BASELINE
// configs 0 and 4 take RGBA32F maps). Host-side ingest, off the bake path. Flat and new-style run-length encoded
// scanlines, -Y +X orientation (the only one in common use; others are reported as unsupported). Values are linear
// radiance: rgb = mantissa * 2^(e - 136), as written by the Radiance tools (no exposure or gamma applied).
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace vlb {

// Decodes `data` into RGBA32F (row 0 = top, alpha 1). Throws std::runtime_error; `unsupported` marks valid files
// outside the handled subset.
void hdr_decode_rgba32f(const uint8_t* data, size_t size, std::vector<float>& rgba, int& width, int& height, bool& unsupported) {
    unsupported = false;
    size_t pos = 0;
    auto line = [&](std::string& out) {
        out.clear();
        while (pos < size && data[pos] != '\n') out.push_back((char)data[pos++]);
        if (pos >= size) return false;
        ++pos;
        return true;
    };
    std::string l;
    if (!line(l) || (l != "#?RADIANCE" && l != "#?RGBE")) { unsupported = true; throw std::runtime_error("HDR: not a Radiance RGBE file"); }
    bool have_format = false;
    for (;;) {
        if (!line(l)) throw std::runtime_error("HDR: truncated header");
        if (l.empty()) break;
        if (l.rfind("FORMAT=", 0) == 0) {
            if (l != "FORMAT=32-bit_rle_rgbe") { unsupported = true; throw std::runtime_error("HDR: only 32-bit_rle_rgbe is decoded"); }
            have_format = true;
        }
    }
    (void)have_format;
    if (!line(l)) throw std::runtime_error("HDR: missing resolution line");
    int H = 0, W = 0;
    if (std::sscanf(l.c_str(), "-Y %d +X %d", &H, &W) != 2) { unsupported = true; throw std::runtime_error("HDR: only the -Y +X orientation is decoded"); }
    if (W <= 0 || H <= 0 || W > 65536 || H > 65536) throw std::runtime_error("HDR: bad size");
    width = W; height = H;
    rgba.assign((size_t)W * H * 4, 1.0f);
    std::vector<uint8_t> scan((size_t)W * 4);
    for (int y = 0; y < H; ++y) {
        if (pos + 4 > size) throw std::runtime_error("HDR: pixel data ends early");
        const bool rle = W >= 8 && W < 32768 && data[pos] == 2 && data[pos + 1] == 2 && (((int)data[pos + 2] << 8) | data[pos + 3]) == W;
        if (rle) {
            pos += 4;
            for (int c = 0; c < 4; ++c) {                           // the four channels are coded one after the other
                int x = 0;
                while (x < W) {
                    if (pos >= size) throw std::runtime_error("HDR: pixel data ends early");
                    int n = data[pos++];
                    if (n > 128) {                                    // run
                        n -= 128;
                        if (pos >= size || x + n > W) throw std::runtime_error("HDR: bad run");
                        const uint8_t v = data[pos++];
                        for (int k = 0; k < n; ++k) scan[(size_t)(x++) * 4 + c] = v;
                    } else {                                          // literal
                        if (n == 0 || pos + n > size || x + n > W) throw std::runtime_error("HDR: bad literal");
                        for (int k = 0; k < n; ++k) scan[(size_t)(x++) * 4 + c] = data[pos++];
                    }
                }
            }
        } else {
            if (pos + (size_t)W * 4 > size) throw std::runtime_error("HDR: pixel data ends early");
            std::memcpy(scan.data(), data + pos, (size_t)W * 4);
            pos += (size_t)W * 4;
        }
        float* o = &rgba[(size_t)y * W * 4];
        for (int x = 0; x < W; ++x, o += 4) {
            const uint8_t* p = &scan[(size_t)x * 4];
            if (p[3] == 0) { o[0] = o[1] = o[2] = 0.f; continue; }
            const float f = std::ldexp(1.0f, (int)p[3] - 136);
            o[0] = p[0] * f; o[1] = p[1] * f; o[2] = p[2] * f;
        }
    }
}

}  // namespace vlb

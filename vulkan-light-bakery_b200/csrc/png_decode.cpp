// png_decode.cpp — minimal PNG reader for glTF baseColor textures; replaces the stb_image decode that
// tinygltf performs for the reference (Scene_t::loadTextures, src/scene_manager.cpp:941-973, reads
// gltfImage.image as RGBA8). Host-side ingest, off the bake path. Handles what glTF exporters write for
// colour textures: 8-bit greyscale / grey+alpha / RGB / RGBA, 1/2/4/8-bit palette and greyscale,
// non-interlaced; the deflate stream is inflated by zlib. Everything else (16-bit, Adam7) is reported as
// unsupported, never guessed.
#include <zlib.h>

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace vlb {

static uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

static int paeth(int a, int b, int c) {
    const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

// Decodes `data` into RGBA8 (row 0 first). Throws std::runtime_error; `unsupported` is set when the file is a
// valid PNG (or another image format) that this reader does not handle.
void png_decode_rgba8(const uint8_t* data, size_t size, std::vector<uint8_t>& rgba, int& width, int& height, bool& unsupported) {
    unsupported = false;
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    if (size < 8 || std::memcmp(data, sig, 8) != 0) {
        unsupported = true;
        throw std::runtime_error("image is neither PNG nor JPEG (only those are decoded)");
    }
    uint32_t W = 0, H = 0;
    int depth = 0, ctype = -1, interlace = 0;
    std::vector<uint8_t> idat, plte, trns;
    size_t pos = 8;
    bool end = false;
    while (!end && pos + 12 <= size) {
        const uint32_t len = be32(data + pos);
        const uint8_t* type = data + pos + 4;
        const uint8_t* body = data + pos + 8;
        if (pos + 12 + (size_t)len > size) throw std::runtime_error("PNG: truncated chunk");
        if (!std::memcmp(type, "IHDR", 4)) {
            if (len < 13) throw std::runtime_error("PNG: bad IHDR");
            W = be32(body); H = be32(body + 4); depth = body[8]; ctype = body[9]; interlace = body[12];
        } else if (!std::memcmp(type, "PLTE", 4)) plte.assign(body, body + len);
        else if (!std::memcmp(type, "tRNS", 4)) trns.assign(body, body + len);
        else if (!std::memcmp(type, "IDAT", 4)) idat.insert(idat.end(), body, body + len);
        else if (!std::memcmp(type, "IEND", 4)) end = true;
        pos += 12 + (size_t)len;
    }
    if (ctype < 0 || W == 0 || H == 0 || W > 32768 || H > 32768) throw std::runtime_error("PNG: missing or bad IHDR");
    const bool sub_byte = (depth == 1 || depth == 2 || depth == 4) && (ctype == 0 || ctype == 3);
    if ((depth != 8 && !sub_byte) || interlace != 0) {
        unsupported = true;
        throw std::runtime_error("PNG: only 8-bit non-interlaced images are decoded (bit depth " + std::to_string(depth) + ")");
    }
    const int ch = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
    if (!ch) throw std::runtime_error("PNG: unknown colour type");
    if (ctype == 3 && plte.empty()) throw std::runtime_error("PNG: palette image without PLTE");
    const size_t stride = ((size_t)W * ch * depth + 7) / 8;      // bytes per row; filters work on whole bytes (bpp >= 1)
    std::vector<uint8_t> raw((stride + 1) * H);
    uLongf out_len = (uLongf)raw.size();
    if (uncompress(raw.data(), &out_len, idat.data(), (uLong)idat.size()) != Z_OK || out_len != raw.size())
        throw std::runtime_error("PNG: inflate failed");
    // unfilter in place (filter types 0..4, PNG spec 9.2)
    std::vector<uint8_t> prev(stride, 0);
    for (uint32_t y = 0; y < H; ++y) {
        uint8_t* row = raw.data() + (stride + 1) * y;
        const int f = row[0];
        uint8_t* px = row + 1;
        for (size_t i = 0; i < stride; ++i) {
            const int a = i >= (size_t)ch ? px[i - ch] : 0, b = prev[i], c = i >= (size_t)ch ? prev[i - ch] : 0;
            int v = px[i];
            switch (f) {
                case 0: break;
                case 1: v += a; break;
                case 2: v += b; break;
                case 3: v += (a + b) >> 1; break;
                case 4: v += paeth(a, b, c); break;
                default: throw std::runtime_error("PNG: bad filter type");
            }
            px[i] = (uint8_t)v;
        }
        std::memcpy(prev.data(), px, stride);
    }
    width = (int)W; height = (int)H;
    rgba.resize((size_t)W * H * 4);
    std::vector<uint8_t> unpacked(sub_byte ? W : 0);
    for (uint32_t y = 0; y < H; ++y) {
        const uint8_t* px = raw.data() + (stride + 1) * y + 1;
        if (sub_byte) {                                          // 1/2/4-bit samples, most significant bits first
            const int maxv = (1 << depth) - 1;
            for (uint32_t x = 0; x < W; ++x) {
                const int bit = (int)x * depth;
                const int v = (px[bit >> 3] >> (8 - depth - (bit & 7))) & maxv;
                unpacked[x] = (uint8_t)(ctype == 0 ? v * 255 / maxv : v);
            }
            px = unpacked.data();
        }
        uint8_t* o = rgba.data() + (size_t)y * W * 4;
        for (uint32_t x = 0; x < W; ++x, o += 4) {
            switch (ctype) {
                case 0: o[0] = o[1] = o[2] = px[x]; o[3] = 255; break;
                case 2: o[0] = px[3 * x]; o[1] = px[3 * x + 1]; o[2] = px[3 * x + 2]; o[3] = 255; break;
                case 3: {
                    const size_t k = px[x];
                    if (3 * k + 2 >= plte.size()) throw std::runtime_error("PNG: palette index out of range");
                    o[0] = plte[3 * k]; o[1] = plte[3 * k + 1]; o[2] = plte[3 * k + 2]; o[3] = k < trns.size() ? trns[k] : 255;
                    break;
                }
                case 4: o[0] = o[1] = o[2] = px[2 * x]; o[3] = px[2 * x + 1]; break;
                default: o[0] = px[4 * x]; o[1] = px[4 * x + 1]; o[2] = px[4 * x + 2]; o[3] = px[4 * x + 3]; break;
            }
        }
    }
}

}  // namespace vlb

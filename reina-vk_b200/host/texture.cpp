// See texture.h.
#include "texture.h"

#include <zlib.h>

#include <cstdlib>
#include <cstring>
#include <fstream>
#include <stdexcept>
#include <vector>

namespace rbhost {

namespace {

[[noreturn]] void fail(const std::string& name, const std::string& why) {
    throw std::runtime_error("Could not load image at path: " + name + ": " + why);
}

uint32_t be32(const uint8_t* p) { return (uint32_t(p[0]) << 24) | (uint32_t(p[1]) << 16) | (uint32_t(p[2]) << 8) | uint32_t(p[3]); }

int paeth(int a, int b, int c) {
    const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
    if (pa <= pb && pa <= pc) return a;
    return pb <= pc ? b : c;
}

}  // namespace

Image8 decode_png_rgba8(const uint8_t* data, size_t size, bool flip, const std::string& name) {
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1A, '\n'};
    if (size < 8 || std::memcmp(data, sig, 8) != 0) fail(name, "not a PNG file");
    uint32_t width = 0, height = 0;
    int depth = 0, ctype = -1, interlace = 0;
    std::vector<uint8_t> idat, palette, trns;
    bool sawHeader = false, sawEnd = false;
    size_t pos = 8;
    while (pos + 12 <= size && !sawEnd) {
        const uint32_t len = be32(data + pos);
        const uint8_t* type = data + pos + 4;
        const uint8_t* body = data + pos + 8;
        if (len > size - pos - 12) fail(name, "truncated chunk");
        const uint32_t crc = be32(body + len);
        if (uint32_t(crc32(0L, type, uInt(len + 4))) != crc) fail(name, "chunk checksum mismatch");
        if (!std::memcmp(type, "IHDR", 4)) {
            if (len != 13) fail(name, "bad IHDR");
            width = be32(body); height = be32(body + 4);
            depth = body[8]; ctype = body[9]; interlace = body[12];
            if (body[10] != 0 || body[11] != 0) fail(name, "unknown compression or filter method");
            sawHeader = true;
        } else if (!std::memcmp(type, "PLTE", 4)) palette.assign(body, body + len);
        else if (!std::memcmp(type, "tRNS", 4)) trns.assign(body, body + len);
        else if (!std::memcmp(type, "IDAT", 4)) idat.insert(idat.end(), body, body + len);
        else if (!std::memcmp(type, "IEND", 4)) sawEnd = true;
        pos += size_t(len) + 12;
    }
    if (!sawHeader || idat.empty()) fail(name, "missing IHDR or IDAT");
    if (width == 0 || height == 0 || width > 65536 || height > 65536) fail(name, "unsupported dimensions");
    int channels;
    switch (ctype) {
        case 0: channels = 1; break;
        case 2: channels = 3; break;
        case 3: channels = 1; break;
        case 4: channels = 2; break;
        case 6: channels = 4; break;
        default: fail(name, "unknown colour type");
    }
    const bool depthOk = (ctype == 0 && (depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16)) ||
                         (ctype == 3 && (depth == 1 || depth == 2 || depth == 4 || depth == 8)) ||
                         ((ctype == 2 || ctype == 4 || ctype == 6) && (depth == 8 || depth == 16));
    if (!depthOk) fail(name, "bit depth not allowed for this colour type");
    if (ctype == 3 && palette.size() < 3) fail(name, "palette image without PLTE");

    const size_t bitsPerPixel = size_t(channels) * depth;
    const size_t bpp = bitsPerPixel < 8 ? 1 : bitsPerPixel / 8;       // filter distance in bytes
    // one pass for a sequential image, the seven Adam7 passes for an interlaced one: {first column, first row, column
    // step, row step}; every pass is a small image of its own (own scanline length, own filter history)
    struct Pass { uint32_t x0, y0, dx, dy, w, h; size_t stride, offset; };
    static const uint32_t kAdam7[7][4] = {{0, 0, 8, 8}, {4, 0, 8, 8}, {0, 4, 4, 8}, {2, 0, 4, 4}, {0, 2, 2, 4}, {1, 0, 2, 2}, {0, 1, 1, 2}};
    if (interlace > 1) fail(name, "unknown interlace method");
    std::vector<Pass> passes;
    size_t rawSize = 0;
    for (int k = 0; k < (interlace ? 7 : 1); k++) {
        Pass ps;
        ps.x0 = interlace ? kAdam7[k][0] : 0; ps.y0 = interlace ? kAdam7[k][1] : 0;
        ps.dx = interlace ? kAdam7[k][2] : 1; ps.dy = interlace ? kAdam7[k][3] : 1;
        ps.w = width > ps.x0 ? (width - ps.x0 + ps.dx - 1) / ps.dx : 0;
        ps.h = height > ps.y0 ? (height - ps.y0 + ps.dy - 1) / ps.dy : 0;
        if (ps.w == 0 || ps.h == 0) continue;                       // an empty pass has no scanlines at all
        ps.stride = (size_t(ps.w) * bitsPerPixel + 7) / 8;
        ps.offset = rawSize;
        rawSize += (ps.stride + 1) * ps.h;
        passes.push_back(ps);
    }
    std::vector<uint8_t> raw(rawSize);
    uLongf rawLen = uLongf(raw.size());
    const int zrc = uncompress(raw.data(), &rawLen, idat.data(), uLong(idat.size()));
    if (zrc != Z_OK || rawLen != raw.size()) fail(name, "corrupt image data");

    // defilter in place (scanlines keep their leading filter byte)
    std::vector<uint8_t> zero((size_t(width) * bitsPerPixel + 7) / 8, 0);
    for (const Pass& ps : passes)
        for (uint32_t y = 0; y < ps.h; y++) {
            uint8_t* cur = &raw[ps.offset + (ps.stride + 1) * y + 1];
            const uint8_t* up = y ? &raw[ps.offset + (ps.stride + 1) * (y - 1) + 1] : zero.data();
            const int filter = cur[-1];
            for (size_t x = 0; x < ps.stride; x++) {
                const int a = x >= bpp ? cur[x - bpp] : 0, b = up[x], c = x >= bpp ? up[x - bpp] : 0;
                int v = cur[x];
                switch (filter) {
                    case 0: break;
                    case 1: v += a; break;
                    case 2: v += b; break;
                    case 3: v += (a + b) >> 1; break;
                    case 4: v += paeth(a, b, c); break;
                    default: fail(name, "unknown scanline filter");
                }
                cur[x] = uint8_t(v);
            }
        }

    Image8 img;
    img.width = int(width);
    img.height = int(height);
    img.rgba.resize(size_t(width) * height * 4);
    const int maxv = (1 << (depth < 8 ? depth : 8)) - 1;
    auto sample = [&](const uint8_t* row, size_t index) -> uint32_t {   // index-th sample of the row, full depth
        if (depth == 8) return row[index];
        if (depth == 16) return (uint32_t(row[2 * index]) << 8) | row[2 * index + 1];
        const size_t bit = index * depth;
        return (row[bit >> 3] >> (8 - depth - (bit & 7))) & uint32_t(maxv);
    };
    auto to8 = [&](uint32_t v) -> uint8_t {
        if (depth == 16) return uint8_t(v >> 8);
        if (depth == 8) return uint8_t(v);
        return uint8_t(v * 255u / uint32_t(maxv));      // 1 -> x255, 2 -> x85, 4 -> x17
    };
    for (const Pass& ps : passes)
        for (uint32_t py = 0; py < ps.h; py++) {
            const uint8_t* row = &raw[ps.offset + (ps.stride + 1) * py + 1];
            const uint32_t y = ps.y0 + py * ps.dy;
            uint8_t* line = &img.rgba[size_t(flip ? height - 1 - y : y) * width * 4];
            for (uint32_t x = 0; x < ps.w; x++) {
                uint8_t* out = line + size_t(ps.x0 + x * ps.dx) * 4;
                switch (ctype) {
                    case 0: {
                        const uint32_t g = sample(row, x);
                        out[0] = out[1] = out[2] = to8(g);
                        out[3] = 255;
                        if (trns.size() >= 2 && g == ((uint32_t(trns[0]) << 8) | trns[1])) out[3] = 0;
                        break;
                    }
                    case 2: {
                        const uint32_t r = sample(row, 3 * x), g = sample(row, 3 * x + 1), b = sample(row, 3 * x + 2);
                        out[0] = to8(r); out[1] = to8(g); out[2] = to8(b); out[3] = 255;
                        if (trns.size() >= 6 && r == ((uint32_t(trns[0]) << 8) | trns[1]) && g == ((uint32_t(trns[2]) << 8) | trns[3]) &&
                            b == ((uint32_t(trns[4]) << 8) | trns[5])) out[3] = 0;
                        break;
                    }
                    case 3: {
                        const uint32_t i = sample(row, x);
                        if (size_t(i) * 3 + 2 >= palette.size()) fail(name, "palette index out of range");
                        out[0] = palette[3 * i]; out[1] = palette[3 * i + 1]; out[2] = palette[3 * i + 2];
                        out[3] = i < trns.size() ? trns[i] : 255;
                        break;
                    }
                    case 4: {
                        out[0] = out[1] = out[2] = to8(sample(row, 2 * x));
                        out[3] = to8(sample(row, 2 * x + 1));
                        break;
                    }
                    default: {
                        for (int k = 0; k < 4; k++) out[k] = to8(sample(row, 4 * x + k));
                        break;
                    }
                }
            }
        }
    return img;
}

Image8 load_png_rgba8(const std::string& path, bool flip) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("Could not load image at path: " + path);
    std::vector<uint8_t> bytes((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    return decode_png_rgba8(bytes.data(), bytes.size(), flip, path);
}

Image8 decode_image_rgba8(const uint8_t* data, size_t size, bool flip, const std::string& name) {
    if (looks_like_jpeg(data, size)) return decode_jpeg_rgba8(data, size, flip, name);
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1A, '\n'};
    if (size >= 8 && std::memcmp(data, sig, 8) == 0) return decode_png_rgba8(data, size, flip, name);
    fail(name, "neither a PNG nor a JPEG file");
}

Image8 load_image_rgba8(const std::string& path, bool flip) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("Could not load image at path: " + path);
    std::vector<uint8_t> bytes((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    return decode_image_rgba8(bytes.data(), bytes.size(), flip, path);
}

}  // namespace rbhost

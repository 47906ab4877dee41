// See png.h.
#include "png.h"

#include <zlib.h>

#include <cstdio>
#include <stdexcept>

namespace rbhost {

namespace {

void put_u32(std::vector<uint8_t>& out, uint32_t v) {
    out.push_back(uint8_t(v >> 24));
    out.push_back(uint8_t(v >> 16));
    out.push_back(uint8_t(v >> 8));
    out.push_back(uint8_t(v));
}

void put_chunk(std::vector<uint8_t>& out, const char type[4], const uint8_t* data, size_t n) {
    put_u32(out, uint32_t(n));
    const size_t start = out.size();
    out.insert(out.end(), type, type + 4);
    if (n) out.insert(out.end(), data, data + n);
    const uint32_t crc = uint32_t(crc32(0L, out.data() + start, uInt(out.size() - start)));
    put_u32(out, crc);
}

}  // namespace

std::vector<uint8_t> encode_png_rgba8(const uint8_t* rgba, uint32_t width, uint32_t height) {
    if (!rgba || width == 0 || height == 0) throw std::runtime_error("Could not save PNG");
    const size_t stride = size_t(width) * 4;
    std::vector<uint8_t> raw((stride + 1) * height);
    for (uint32_t y = 0; y < height; y++) {
        raw[(stride + 1) * y] = 0;   // filter type None
        std::copy(rgba + stride * y, rgba + stride * (y + 1), raw.begin() + (stride + 1) * y + 1);
    }
    uLongf bound = compressBound(uLong(raw.size()));
    std::vector<uint8_t> z(bound);
    if (compress2(z.data(), &bound, raw.data(), uLong(raw.size()), 6) != Z_OK) throw std::runtime_error("Could not save PNG");
    z.resize(bound);

    std::vector<uint8_t> out = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1A, '\n'};
    std::vector<uint8_t> ihdr;
    put_u32(ihdr, width);
    put_u32(ihdr, height);
    const uint8_t tail[5] = {8, 6, 0, 0, 0};   // 8 bits, colour type 6 (RGBA), deflate, adaptive filtering, no interlace
    ihdr.insert(ihdr.end(), tail, tail + 5);
    put_chunk(out, "IHDR", ihdr.data(), ihdr.size());
    put_chunk(out, "IDAT", z.data(), z.size());
    put_chunk(out, "IEND", nullptr, 0);
    return out;
}

void write_png_rgba8(const std::string& filename, const uint8_t* rgba, uint32_t width, uint32_t height) {
    std::vector<uint8_t> bytes = encode_png_rgba8(rgba, width, height);
    FILE* f = std::fopen(filename.c_str(), "wb");
    if (!f) throw std::runtime_error("Could not save PNG");
    const size_t n = std::fwrite(bytes.data(), 1, bytes.size(), f);
    const int rc = std::fclose(f);
    if (n != bytes.size() || rc != 0) throw std::runtime_error("Could not save PNG");
}

}  // namespace rbhost

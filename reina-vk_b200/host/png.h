// PNG writer for the RGBA8 frame (the reference saves through stb_image_write: 8-bit RGBA, rows top-down,
// stride 4*width, src/Reina.cpp:28-35). Encoding: one IDAT, filter 0 on every row, zlib deflate.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace rbhost {

std::vector<uint8_t> encode_png_rgba8(const uint8_t* rgba, uint32_t width, uint32_t height);

// throws std::runtime_error("Could not save PNG") on failure, as the reference (src/Reina.cpp:37-40)
void write_png_rgba8(const std::string& filename, const uint8_t* rgba, uint32_t width, uint32_t height);

}  // namespace rbhost

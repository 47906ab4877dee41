// Texture ingest for the headless host: what reina::graphics::Image's file constructor does with
// stb_image (src/graphics/Image.cpp:10-23) — decode to 8-bit RGBA, rows flipped vertically for file textures (:14;
// in-memory glTF images are not flipped, :28), bytes used as UNORM without sRGB decoding (:56).
// PNG: sequential or Adam7-interlaced, colour types 0/2/3/4/6, bit depths 1-16 (16-bit samples keep their high byte, low-depth grey is
// scaled to 0..255, palette + tRNS and colour-key tRNS honoured), inflate through zlib.
// JPEG (jpeg.cpp): 8-bit baseline / extended sequential / progressive Huffman files, grayscale or YCbCr with 4:4:4 /
// 4:2:2 / 4:4:0 / 4:2:0 sampling, decoded with the IJG integer pipeline (what PIL yields, byte for byte); arithmetic-coded
// and CMYK files are refused with the reference's message.
#pragma once
#include <string>

#include "model.h"

namespace rbhost {

// throws std::runtime_error("Could not load image at path: <path>[: reason]")
Image8 load_png_rgba8(const std::string& path, bool flipVertically = true);
Image8 decode_png_rgba8(const uint8_t* data, size_t size, bool flipVertically, const std::string& nameForErrors);
bool looks_like_jpeg(const uint8_t* data, size_t size);
Image8 decode_jpeg_rgba8(const uint8_t* data, size_t size, bool flipVertically, const std::string& nameForErrors);
// PNG or JPEG by signature (what stbi_load / stbi_load_from_memory accept of the reference's assets)
Image8 decode_image_rgba8(const uint8_t* data, size_t size, bool flipVertically, const std::string& nameForErrors);
Image8 load_image_rgba8(const std::string& path, bool flipVertically = true);

}  // namespace rbhost

// Texture ingest for the headless host: what reina::graphics::Image's file constructor does with
// stb_image (src/graphics/Image.cpp:10-23) — decode to 8-bit RGBA, rows flipped vertically for file textures (:14;
// in-memory glTF images are not flipped, :28), bytes used as UNORM without sRGB decoding (:56).
// PNG only: non-interlaced, colour types 0/2/3/4/6, bit depths 1-16 (16-bit samples keep their high byte, low-depth
// grey is scaled to 0..255, palette + tRNS and colour-key tRNS honoured), inflate through zlib. JPEG and Adam7 PNGs
// are refused with the reference's message; none of the reference's shipped textures needs them.
#pragma once
#include <string>

#include "model.h"

namespace rbhost {

// throws std::runtime_error("Could not load image at path: <path>[: reason]")
Image8 load_png_rgba8(const std::string& path, bool flipVertically = true);
Image8 decode_png_rgba8(const uint8_t* data, size_t size, bool flipVertically, const std::string& nameForErrors);

}  // namespace rbhost

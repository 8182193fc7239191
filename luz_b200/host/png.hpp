// png.hpp -- minimal PNG decoder for the scene importers (textures embedded in .glb files and referenced by .mtl).
// The reference decodes images with stb_image through tiny_gltf / AssetIO::ImportTexture (AssetIO.cpp:86-103,
// :147-169), always asking for 4 channels; decode_png returns the same RGBA8 bytes for non-interlaced PNGs of bit
// depth <= 8 (every colour type, tRNS included) and reports anything else as an error instead of guessing.
#pragma once

#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace luzhost {

bool decode_png(const uint8_t* data, size_t size, std::vector<uint8_t>& rgba, int& width, int& height, std::string& err);
// zlib stream (RFC 1950 / 1951) -> bytes; false on malformed input
bool inflate_zlib(const uint8_t* data, size_t size, std::vector<uint8_t>& out, std::string& err);

} // namespace luzhost

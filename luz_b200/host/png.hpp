// png.hpp -- minimal PNG decoder for the scene importers (textures embedded in .glb files and referenced by .mtl).
// The reference decodes images with stb_image through tiny_gltf / AssetIO::ImportTexture (AssetIO.cpp:86-103,
// :147-169), always asking for 4 channels of 8 bits; decode_png returns the same RGBA8 bytes for every colour type and
// bit depth (16-bit samples keep their high byte, like stb's 8-bit API), tRNS and Adam7 interlacing included.
#pragma once

#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace luzhost {

// `deep`, if given, receives the full 16-bit RGBA samples (host byte order) of a 16-bit PNG and is left empty otherwise.
bool decode_png(const uint8_t* data, size_t size, std::vector<uint8_t>& rgba, int& width, int& height, std::string& err,
                std::vector<uint16_t>* deep = nullptr);
// zlib stream (RFC 1950 / 1951) -> bytes; false on malformed input
bool inflate_zlib(const uint8_t* data, size_t size, std::vector<uint8_t>& out, std::string& err);

} // namespace luzhost

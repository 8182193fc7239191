// import.hpp -- scene import (glTF 2.0 .gltf / .glb, Wavefront .obj + .mtl, PNG textures) into the host mirror's
// AssetManager.  Mirrors the reference's AssetIO (source/Resources/AssetIO.hpp:5-22, AssetIO.cpp:32-43, :73-122,
// :124-403 glTF, :405-528 OBJ): same function names, same asset / node naming, same vertex order, tangent generation,
// OBJ v-flip, vertex de-duplication and per-material splitting -- written from the file format specifications, without
// tiny_gltf / tiny_obj_loader / stb_image.  Checked bit for bit against the reference's own compiled importers
// (oracle/_ref/ref_import -> tests/golden/import/*.json, tests/test_host_import.py).
#pragma once

#include <string>

#include "scene.hpp"

namespace luzhost {
namespace AssetIO {

bool IsTexture(const std::string& path); // .jpg .png .jpeg .tga .bmp (AssetIO.cpp:32-35); only .png can be decoded here
bool IsScene(const std::string& path);   // .obj .gltf .glb (AssetIO.cpp:37-40)
// Return the uuid of the imported scene / texture, or 0 with manager.error set (the reference logs and returns 0).
UUID Import(const std::string& path, AssetManager& manager);
UUID ImportScene(const std::string& path, AssetManager& manager);
UUID ImportSceneGLTF(const std::string& path, AssetManager& manager);
UUID ImportSceneOBJ(const std::string& path, AssetManager& manager);
UUID ImportTexture(const std::string& path, AssetManager& manager);

// Test hook: the imported scene in the JSON layout of oracle/ref_import.cpp (floats as bit patterns).
std::string DumpImportedScene(AssetManager& manager, UUID scene);

} // namespace AssetIO
} // namespace luzhost

// glsl_bindings.hpp -- the descriptor arrays the reference's own macros index (LuzCommon.h: `#define scene
// sceneBuffers[ctx.sceneBufferIndex].block`, `#define tlas tlasBuffer[scene.tlasRid]`).  Included by the generated
// translation unit right after the LuzCommon.h structs, inside the shader's namespace.  No include guard: each shader
// namespace gets its own set.
struct SceneBufferSlot {
    SceneBlock block;
};
GLSL_GLOBAL SceneBufferSlot sceneBuffers[1];
GLSL_GLOBAL sampler2D textures[8];
GLSL_GLOBAL samplerCube cubeTextures[1];
GLSL_GLOBAL accelerationStructureEXT tlasBuffer[1];
GLSL_GLOBAL image2D images[1];

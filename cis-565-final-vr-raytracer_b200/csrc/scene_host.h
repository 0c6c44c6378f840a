// scene_host.h — host half of the Scene object: the tables the kernels read, built exactly the way the
// reference's Scene::load does (src/scene.cpp:57-125), plus the camera state of Scene::updateCamera.
#pragma once
#include <vector>
#include "gltf_import.h"
#include "host_device.h"

namespace eid {

// per TLAS instance (== glTF node, accelstruct.cpp:132-162)
struct InstanceXform {
  float objectToWorld[12];   // 4 columns x 3 rows, column-major (GLSL mat4x3)
  float worldToObject[12];
  int32_t primMesh;          // instanceCustomIndex
  uint32_t flags;            // bit0: triangle facing cull disabled (double sided), bit1: mirroring transform,
                             // bit2: force opaque
  uint32_t firstTriangle;    // offset of this instance's triangles in the flattened world-space list
  uint32_t triangleCount;
};
enum { INST_CULL_DISABLE = 1u, INST_MIRROR = 2u, INST_FORCE_OPAQUE = 4u };

// one entry of texturesMap[] (layouts.glsl:51): 8-bit UNORM texels + sampler state (Scene::createTextureImages, scene.cpp:513-646)
struct TextureHost {
  uint32_t width = 1, height = 1;
  uint64_t texelOffset = 0;      // into SceneHost::texels (uint32 RGBA8 each)
  int32_t linear = 1;            // magnification filter; every tap is textureLod(..., 0)
  int32_t wrapS = 0, wrapT = 0;  // 0 REPEAT, 1 MIRRORED_REPEAT, 2 CLAMP_TO_EDGE
};

struct SceneHost {
  HostGltf gltf;
  std::vector<TextureHost> textures;
  std::vector<uint32_t> texels;
  // concatenated vertex / index storage; prim mesh p owns vertices [vtxBase[p], +vertexCount) and
  // indices [idxBase[p], +indexCount).  Prim meshes sharing an accessor set share the vertex range
  // (the reference's m_cachePrimitive, scene.cpp:223-234).
  std::vector<VertexAttributes> vertices;
  std::vector<uint32_t> indices;
  std::vector<uint64_t> vtxBase, idxBase;
  std::vector<GltfShadeMaterial> materials;
  std::vector<PuncLight> puncLights;
  std::vector<TrigLight> trigLights;
  LightBufInfo lightInfo{};
  float trigLightWeight = 0.f, puncLightWeight = 0.f;
  std::vector<InstanceXform> instances;
  uint64_t triangleInstances = 0;
  bool hasNonOpaque = false;
  bool hasTextures = false;      // any material with a texture index > -1 (selects the texture-aware kernel variants)

  // camera (CameraManip state + Scene::m_camera)
  SceneCamera camera{};
  float eye[3] = {2.f, 2.f, -5.f}, center[3] = {-1.f, 2.f, -1.f}, up[3] = {0.f, 1.f, 0.f};   // main.cpp:68
  float fovDeg = 60.f;
  float prevEye[3] = {0.f, 0.f, 0.f};

  void build();                                  // all tables from `gltf`
  void updateCamera(uint32_t w, uint32_t h);     // scene.cpp:777-826
};

}  // namespace eid

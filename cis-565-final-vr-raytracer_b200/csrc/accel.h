// accel.h — device-resident scene + acceleration structure shared by accel.cu and render.cu.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "scene_host.h"

#ifndef EID_BVH_WIDTH
#define EID_BVH_WIDTH 4
#endif
// EID_NODE_Q8 = 1 (BVH4 only): 64-byte node with the child boxes quantised to 8 bits per plane inside the node's own box
#ifndef EID_NODE_Q8
#define EID_NODE_Q8 0
#endif
#define EID_NODE_BYTES ((EID_BVH_WIDTH == 4 && !EID_NODE_Q8) ? 128 : 64)

namespace eid {

// What a kernel needs to shade: the reference's S_SCENE descriptor set (layouts.glsl:48-54).
struct TextureDev {                     // device twin of TextureHost with a resolved texel pointer
  const uint32_t* texels;
  int32_t width, height, linear, wrapS, wrapT, pad;
};

struct DeviceSceneView {
  const TextureDev* textures;           // texturesMap[]
  const InstanceData* geoInfo;          // per prim mesh: vertex/index device addresses + material
  const GltfShadeMaterial* materials;
  const PuncLight* puncLights;
  const TrigLight* trigLights;
  const InstanceXform* instances;       // per TLAS instance
  LightBufInfo lightBufInfo;
};

// EID_BVH_WIDTH == 4 (default): 128 B node = 8 x float4: lo.x[4], lo.y[4], lo.z[4], hi.x[4], hi.y[4], hi.z[4], child refs[4], -
// EID_NODE_Q8: 64 B = 4 x uint4: (origin.xyz, scale.x) (scale.y, scale.z, lo.x[4 x u8], lo.y[4 x u8]) (lo.z, hi.x, hi.y, hi.z [4 x u8 each])
//   (child refs[4]); plane = origin + q * scale, lo rounded down / hi rounded up so the decoded box contains the exact one
// EID_BVH_WIDTH == 2: BVH2 node, 64 B = 4 x float4:
//   q0 = lo0.xyz, hi0.x   q1 = hi0.yz, lo1.xy   q2 = lo1.z, hi1.xyz   q3 = child0, child1, -, -  (as int bits)
// child >= 0: inner node index; child < 0: leaf, ~child = (firstTriangle << 3) | count  (count 0..4)
// Triangle record, 48 B = 3 x float4 (world space, Moller-Trumbore ready):
//   t0 = v0.xyz, e1.x   t1 = e1.yz, e2.xy   t2 = e2.z, primitiveID, instanceID, flags   (as int bits)
struct AccelView {
  const float4* nodes;
  const float4* tris;
  uint32_t triCount;
  int32_t rootRef;      // child-style reference of the root
  cudaTextureObject_t nodeTex, triTex;   // the same two arrays as linear float4 textures (TEX-pipe fetch experiments, EID_FETCH_TEX)
  // two-level form (trace.cuh: traverse2): `nodes` / `tris` then hold the bottom-level trees (object-space triangles) of all prim meshes
  int32_t twoLevel;
  int32_t tlasRootRef;
  uint32_t tlasPrimCount;
  const float4* tlasNodes;
  const float4* tlasPrims;               // per TLAS primitive: (instance index, BLAS root reference, -, -) x 3 float4 (48-byte records like triangles)
  const InstanceXform* instances;
};

struct SceneDevice {
  int device = 0;
  VertexAttributes* vertices = nullptr;
  uint32_t* indices = nullptr;
  InstanceData* geoInfo = nullptr;
  GltfShadeMaterial* materials = nullptr;
  PuncLight* puncLights = nullptr;
  TrigLight* trigLights = nullptr;
  InstanceXform* instances = nullptr;
  TextureDev* textures = nullptr;
  uint32_t* texels = nullptr;
  uint32_t* instFirstTri = nullptr;     // exclusive prefix of triangle counts per instance (+ total)
  void upload(const SceneHost& h);
  void release();
  DeviceSceneView view(const SceneHost& h) const;
};

}  // namespace eid

struct eid_scene {
  std::vector<eid::HostGltf::Image> providedImages;   // eid_scene_provide_image: decoded by the host before eid_scene_load_gltf
  eid::SceneHost host;
  eid::SceneDevice dev;
  bool loaded = false;
};

struct eid_accel {
  eid_scene* scene = nullptr;
  float4* nodes = nullptr;
  float4* tris = nullptr;
  uint32_t triCount = 0;
  uint32_t nodeCount = 0;
  uint32_t nodeAlloc = 0;
  uint32_t maxDepth = 0;
  int32_t rootRef = -1;
  float buildMs = 0.f;
  cudaTextureObject_t nodeTex = 0, triTex = 0;
  // two-level form: BLAS per prim mesh (nodes / tris above hold all of them) + TLAS over the instances
  bool twoLevel = false;
  float4* tlasNodes = nullptr;
  float4* tlasPrims = nullptr;
  uint32_t tlasPrimCount = 0, tlasNodeCount = 0, blasCount = 0;
  int32_t tlasRootRef = -1;
  uint64_t uniqueTriangles = 0;
  bool sahBuild = false;                 // tree topology from the host binned-SAH builder (EID_ACCEL_FAST_TRACE) instead of the Morton build
  eid::AccelView view() const {
    return eid::AccelView{nodes, tris, triCount, rootRef, nodeTex, triTex, twoLevel ? 1 : 0, tlasRootRef, tlasPrimCount, tlasNodes, tlasPrims,
                          scene ? scene->dev.instances : nullptr};
  }
};

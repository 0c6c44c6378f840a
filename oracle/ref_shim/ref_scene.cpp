/*
 * oracle/ref_shim/ref_scene.cpp — TEST INFRASTRUCTURE.
 * Runs the reference's OWN src/scene.cpp — compiled where it lies against the stand-ins of scene/scene_shim.h — on a scene the harness
 * describes (eid_scene_desc = what nvh::GltfScene holds after the un-vendored nvpro_core import), and hands back the tables it uploads:
 * Scene::load -> createMaterialBuffer, createPuncLightBuffer (+ createPuncLightImptSampAccel), createVertexBuffer, createInstanceDataBuffer,
 * createTrigLightBuffer (+ createTrigLightImptSampAccel, alias_table.hpp), the LightBufInfo block, m_trigLightWeight / m_puncLightWeight;
 * Scene::updateCamera -> SceneCamera (history roll, jitter) with the contract's nvmath stand-ins; AccelStructure::create (src/accelstruct.cpp,
 * called like SampleExample::loadScene does, sample_example.cpp:85) -> what it asks the builder to build: per node the instance record
 * (transform, instanceCustomIndex, mask, flags = FORCE_OPAQUE / TRIANGLE_FACING_CULL_DISABLE rule, BLAS reference), per prim mesh the geometry.  The table numbering is
 * eid_scene_table's (include/eidola.h), so the tests compare the three sides — reference code, oracle, product — table by table.
 */
#include "scene_shim.h"
#include "shaders/host_device.h"      // /root/reference/shaders/host_device.h
#define private public                 // the tables are private members of Scene
#include "scene.hpp"                   // /root/reference/src/scene.hpp
#include "accelstruct.hpp"             // /root/reference/src/accelstruct.hpp
#undef private

static const eidc::eid_scene_desc* g_desc;

// ---- the stand-ins that have something to do ---------------------------------------------------------------------------------------
void vkCmdUpdateBuffer(VkCommandBuffer, VkBuffer dst, VkDeviceSize offset, VkDeviceSize size, const void* data) {
  if (dst && offset + size <= dst->bytes.size()) memcpy(dst->bytes.data() + offset, data, (size_t)size);
}
static bool fillModel(tinygltf::Model* m) {       // tinygltf::Model: images (decoded RGBA8), textures, samplers of the described scene
  const eidc::eid_scene_desc& d = *g_desc;
  m->images.clear(); m->textures.clear(); m->samplers.clear();
  for (uint32_t i = 0; i < d.imageCount; ++i) {
    tinygltf::Image im;
    if (d.images[i].rgba8 && d.images[i].width && d.images[i].height) {
      im.width = (int)d.images[i].width; im.height = (int)d.images[i].height;
      im.image.assign(d.images[i].rgba8, d.images[i].rgba8 + 4 * (size_t)im.width * im.height);
    }
    m->images.push_back(im);
  }
  for (uint32_t i = 0; i < d.textureCount; ++i) {
    tinygltf::Texture t; t.source = d.textures[i].image;
    if (d.textures[i].hasSampler) {
      tinygltf::Sampler s; s.magFilter = d.textures[i].magFilter; s.minFilter = d.textures[i].minFilter; s.wrapS = d.textures[i].wrapS; s.wrapT = d.textures[i].wrapT;
      t.sampler = (int)m->samplers.size(); m->samplers.push_back(s);
    }
    m->textures.push_back(t);
  }
  return true;
}
bool tinygltf::TinyGLTF::LoadASCIIFromFile(Model* m, std::string*, std::string*, const std::string&) { return fillModel(m); }
bool tinygltf::TinyGLTF::LoadBinaryFromFile(Model* m, std::string*, std::string*, const std::string&) { return fillModel(m); }

void nvh::GltfScene::importMaterials(const tinygltf::Model&) {
  const eidc::eid_scene_desc& d = *g_desc;
  m_materials.clear();
  for (uint32_t i = 0; i < d.materialCount; ++i) {
    const eidc::eid_material_desc& s = d.materials[i];
    GltfMaterial m;
    m.baseColorFactor = nvmath::vec4f(s.baseColorFactor[0], s.baseColorFactor[1], s.baseColorFactor[2], s.baseColorFactor[3]);
    m.baseColorTexture = s.baseColorTexture; m.metallicFactor = s.metallicFactor; m.roughnessFactor = s.roughnessFactor;
    m.metallicRoughnessTexture = s.metallicRoughnessTexture; m.emissiveTexture = s.emissiveTexture;
    m.emissiveFactor = nvmath::vec3f(s.emissiveFactor[0], s.emissiveFactor[1], s.emissiveFactor[2]);
    m.alphaMode = s.alphaMode; m.alphaCutoff = s.alphaCutoff; m.doubleSided = s.doubleSided; m.normalTexture = s.normalTexture;
    m.normalTextureScale = s.normalTextureScale; m.transmission.factor = s.transmissionFactor; m.transmission.texture = s.transmissionTexture;
    m.ior.ior = s.ior;
    m_materials.push_back(m);
  }
}
void nvh::GltfScene::importDrawableNodes(const tinygltf::Model&, GltfAttributes) {
  const eidc::eid_scene_desc& d = *g_desc;
  m_positions.clear(); m_normals.clear(); m_tangents.clear(); m_texcoords0.clear(); m_colors0.clear();
  for (uint32_t i = 0; i < d.vertexCount; ++i) {
    m_positions.emplace_back(d.positions[3 * i], d.positions[3 * i + 1], d.positions[3 * i + 2]);
    m_normals.emplace_back(d.normals[3 * i], d.normals[3 * i + 1], d.normals[3 * i + 2]);
    m_tangents.emplace_back(d.tangents[4 * i], d.tangents[4 * i + 1], d.tangents[4 * i + 2], d.tangents[4 * i + 3]);
    m_texcoords0.emplace_back(d.texcoords0[2 * i], d.texcoords0[2 * i + 1]);
    m_colors0.emplace_back(d.colors0[4 * i], d.colors0[4 * i + 1], d.colors0[4 * i + 2], d.colors0[4 * i + 3]);
  }
  m_indices.assign(d.indices, d.indices + d.indexCount);
  m_primMeshes.clear();
  for (uint32_t i = 0; i < d.primMeshCount; ++i) {
    GltfPrimMesh p; p.firstIndex = d.primMeshes[i].firstIndex; p.indexCount = d.primMeshes[i].indexCount; p.vertexOffset = d.primMeshes[i].vertexOffset;
    p.vertexCount = d.primMeshes[i].vertexCount; p.materialIndex = d.primMeshes[i].materialIndex;
    m_primMeshes.push_back(p);
  }
  m_nodes.clear();
  for (uint32_t i = 0; i < d.nodeCount; ++i) { GltfNode n; memcpy(n.worldMatrix.m, d.nodes[i].worldMatrix, 64); n.primMesh = d.nodes[i].primMesh; m_nodes.push_back(n); }
  m_lights.clear();
  for (uint32_t i = 0; i < d.lightCount; ++i) {
    GltfLight l; memcpy(l.worldMatrix.m, d.lights[i].worldMatrix, 64);
    l.light.type = d.lights[i].type == LightType_Point ? "point" : d.lights[i].type == LightType_Directional ? "directional" : "spot";
    l.light.color = {d.lights[i].color[0], d.lights[i].color[1], d.lights[i].color[2]};
    l.light.intensity = d.lights[i].intensity; l.light.range = d.lights[i].range;
    l.light.spot.innerConeAngle = d.lights[i].innerConeAngle; l.light.spot.outerConeAngle = d.lights[i].outerConeAngle;
    m_lights.push_back(l);
  }
  m_cameras.clear();
  if (d.hasCamera) {
    GltfCamera c; c.eye = nvmath::vec3f(d.camEye[0], d.camEye[1], d.camEye[2]); c.center = nvmath::vec3f(d.camCenter[0], d.camCenter[1], d.camCenter[2]);
    c.up = nvmath::vec3f(d.camUp[0], d.camUp[1], d.camUp[2]); c.cam.perspective.yfov = d.camYfovRad;
    m_cameras.push_back(c);
  }
}

// ---- C entry points ----------------------------------------------------------------------------------------------------------------
#define REF_API extern "C" __attribute__((visibility("default")))
struct RefScene { nvvk::ResourceAllocator alloc; Scene scene; };

REF_API void* ref_scene_load(const void* desc) {
  g_desc = (const eidc::eid_scene_desc*)desc;
  RefScene* r = new RefScene;
  r->scene.setup(nullptr, nullptr, nvvk::Queue{}, &r->alloc);
  const bool ok = r->scene.load("injected.gltf");
  g_desc = nullptr;
  if (!ok) { delete r; return nullptr; }
  return r;
}
REF_API void ref_scene_destroy(void* h) { delete (RefScene*)h; }
// bytes of table `which` (eid_scene_table numbering); dst == nullptr: size query.  InstanceData is returned with its addresses as uploaded
// (host pointers of the stand-in buffers): only materialIndex is comparable.
REF_API long ref_scene_table(void* h, int which, unsigned index, void* dst, long cap) {
  RefScene* r = (RefScene*)h;
  VkBuffer b = nullptr;
  switch (which) {
    case 0: b = r->scene.m_buffer[Scene::eMaterial].buffer; break;
    case 1: b = r->scene.m_buffer[Scene::ePuncLights].buffer; break;
    case 2: b = r->scene.m_buffer[Scene::eTrigLights].buffer; break;
    case 3: b = r->scene.m_buffer[Scene::eLightBufInfo].buffer; break;
    case 4: b = r->scene.m_buffer[Scene::eInstData].buffer; break;
    case 5: if (index < r->scene.m_buffers[Scene::eVertex].size()) b = r->scene.m_buffers[Scene::eVertex][index].buffer; break;
    case 6: if (index < r->scene.m_buffers[Scene::eIndex].size()) b = r->scene.m_buffers[Scene::eIndex][index].buffer; break;
    case 7: b = r->scene.m_buffer[Scene::eCameraMat].buffer; break;
  }
  if (!b) return -1;
  const long n = (long)b->bytes.size();
  if (dst) memcpy(dst, b->bytes.data(), (size_t)(n < cap ? n : cap));
  return n;
}
REF_API void ref_scene_weights(void* h, float* trig, float* punc) { RefScene* r = (RefScene*)h; *trig = r->scene.m_trigLightWeight; *punc = r->scene.m_puncLightWeight; }
REF_API void ref_scene_set_lookat(const float* eye, const float* center, const float* up, float fovDeg) {   // CameraManip.setLookat / setFov (main.cpp:67-68)
  CameraManip.setLookat(nvmath::vec3f(eye[0], eye[1], eye[2]), nvmath::vec3f(center[0], center[1], center[2]), nvmath::vec3f(up[0], up[1], up[2]));
  CameraManip.setFov(fovDeg);
}
REF_API void ref_scene_update_camera(void* h, unsigned w, unsigned hgt) { ((RefScene*)h)->scene.updateCamera(nullptr, VkExtent2D{w, hgt}); }

// AccelStructure::create on the loaded scene.  inst: rows of (instanceCustomIndex, mask, sbtRecordOffset, flags, blas index); xforms: 12 floats
// per instance, row-major 3x4 (VkTransformMatrixKHR); blas: rows of (primitiveCount, maxVertex, vertexStride, geometry flags, vertexFormat,
// indexType); build[0..1] = BLAS / TLAS build flags.  Returns the instance count, *nBlas the BLAS count.
REF_API int ref_scene_accel(void* h, int* inst, float* xforms, int capInst, int* blas, int capBlas, int* nBlas, int* build) {
  RefScene* r = (RefScene*)h;
  AccelStructure as;
  as.setup(nullptr, nullptr, 0u, &r->alloc);
  as.create(r->scene.getScene(), r->scene.getBuffers(Scene::eVertex), r->scene.getBuffers(Scene::eIndex));
  const auto& b = as.m_rtBuilder;
  for (size_t i = 0; i < b.tlas.size() && (int)i < capInst; ++i) {
    const VkAccelerationStructureInstanceKHR& t = b.tlas[i];
    inst[5 * i] = (int)t.instanceCustomIndex; inst[5 * i + 1] = (int)t.mask; inst[5 * i + 2] = (int)t.instanceShaderBindingTableRecordOffset;
    inst[5 * i + 3] = (int)t.flags; inst[5 * i + 4] = (int)(t.accelerationStructureReference - 0x1000u);
    memcpy(xforms + 12 * i, t.transform.matrix, 48);
  }
  for (size_t i = 0; i < b.blas.size() && (int)i < capBlas; ++i) {
    const auto& g = b.blas[i].asGeometry[0]; const auto& o = b.blas[i].asBuildOffsetInfo[0];
    blas[6 * i] = (int)o.primitiveCount; blas[6 * i + 1] = (int)g.geometry.triangles.maxVertex; blas[6 * i + 2] = (int)g.geometry.triangles.vertexStride;
    blas[6 * i + 3] = (int)g.flags; blas[6 * i + 4] = g.geometry.triangles.vertexFormat; blas[6 * i + 5] = g.geometry.triangles.indexType;
  }
  *nBlas = (int)b.blas.size(); build[0] = (int)b.blasFlags; build[1] = (int)b.tlasFlags;
  return (int)b.tlas.size();
}

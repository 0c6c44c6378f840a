/* oracle/ref_shim/scene/scene_shim.h — TEST INFRASTRUCTURE.
 * Everything src/scene.{hpp,cpp} of the reference names but does not define — Vulkan, nvpro_core (nvvk, nvh, nvmath, ImGuiH), tinygltf,
 * FreeImage — as stand-ins, so that scene.cpp COMPILES WHERE IT LIES and its table builders run: createMaterialBuffer,
 * createPuncLightBuffer (+ createPuncLightImptSampAccel), createVertexBuffer, createInstanceDataBuffer, createTrigLightBuffer
 * (+ createTrigLightImptSampAccel), the LightBufInfo block of Scene::load, updateCamera.  Inert: every Vulkan call.  Functional:
 *   - nvvk::ResourceAllocator::createBuffer keeps a copy of the bytes it is given (that is how the tests read the tables back);
 *   - nvh::GltfScene is the plain data its importer would have produced; importMaterials / importDrawableNodes (nvpro_core, un-vendored)
 *     copy an injected scene (ref_scene.cpp fills it from the harness's eid_scene_desc);
 *   - nvmath / CameraManip: vector and matrix operations with the evaluation order of the numerical contract (include/eid_vecmath.h) —
 *     nvmath is un-vendored, so these are contract, not reference arithmetic.
 * The header guards of the real headers are irrelevant: the forwarding headers next to this file all include it. */
#pragma once
#include <array>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <filesystem>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>
#include <cassert>
#include <algorithm>
#include <math.h>
#include <stdint.h>
#include <stddef.h>
#include <string.h>
namespace eidc {              // the product's public headers, kept apart from the reference's host_device.h (same struct names)
#include "eidola.h"           // include/: eid_scene_desc (what ref_scene.cpp injects)
#include "eid_vecmath.h"      // include/: eidc::eid_mat4 + the contract's look-at / perspectiveVK / invert / mul
}
using std::abs; using std::isinf;

// ---- Vulkan --------------------------------------------------------------------------------------------------------------------
typedef struct VkBuffer_T* VkBuffer; typedef void* VkDevice; typedef void* VkPhysicalDevice; typedef void* VkQueue; typedef void* VkCommandBuffer;
typedef void* VkImage; typedef void* VkImageView; typedef void* VkSampler; typedef void* VkDescriptorPool; typedef void* VkDescriptorSetLayout;
typedef void* VkDescriptorSet; typedef uint64_t VkDeviceSize; typedef uint32_t VkFlags; typedef VkFlags VkShaderStageFlags; typedef uint64_t VkDeviceAddress;
#define VK_NULL_HANDLE nullptr
#define VK_WHOLE_SIZE (~0ULL)
struct VkExtent2D { uint32_t width, height; };
typedef struct VkPipeline_T* VkPipeline; typedef void* VkPipelineLayout; typedef void* VkShaderModule;
struct VkPipeline_T { int tag; };                                       // a pipeline is the tag of the shader it was created from
enum VkPipelineBindPoint { VK_PIPELINE_BIND_POINT_GRAPHICS = 0, VK_PIPELINE_BIND_POINT_COMPUTE = 1 };
typedef void* VkRenderPass;
enum { VK_SHADER_STAGE_VERTEX_BIT = 1, VK_CULL_MODE_NONE = 0 };
#define LABEL_SCOPE_VK(cmd) do { } while (0)
enum VkImageLayout { VK_IMAGE_LAYOUT_UNDEFINED = 0, VK_IMAGE_LAYOUT_GENERAL = 1 };
enum { VK_IMAGE_USAGE_STORAGE_BIT = 8, VK_IMAGE_USAGE_COLOR_ATTACHMENT_BIT = 0x10 };
struct VkPushConstantRange { VkFlags stageFlags; uint32_t offset, size; };
struct VkPipelineLayoutCreateInfo { int sType; const void* pNext; VkFlags flags; uint32_t setLayoutCount; void* const* pSetLayouts; uint32_t pushConstantRangeCount; const VkPushConstantRange* pPushConstantRanges; };
struct VkPipelineShaderStageCreateInfo { int sType; const void* pNext; VkFlags flags; int stage; VkShaderModule module; const char* pName; const void* pSpecializationInfo; };
struct VkComputePipelineCreateInfo { int sType; const void* pNext; VkFlags flags; VkPipelineShaderStageCreateInfo stage; VkPipelineLayout layout; VkPipeline basePipelineHandle; int basePipelineIndex; };
enum { VK_STRUCTURE_TYPE_PIPELINE_SHADER_STAGE_CREATE_INFO = 18, VK_STRUCTURE_TYPE_COMPUTE_PIPELINE_CREATE_INFO = 29, VK_STRUCTURE_TYPE_PIPELINE_LAYOUT_CREATE_INFO = 30 };
// ---- the recording "device": what a command buffer executes and which resource each descriptor names (renderer.cpp) -------------
struct ShimResource { int kind; int id; uint64_t bytes; uint32_t width, height; int format; int mips; };      // kind 0 = buffer, 1 = image
struct ShimEvent { int what; int a, b, c; std::vector<unsigned char> data; };                        // see ref_renderer.cpp
struct ShimDevice {
  std::vector<std::unique_ptr<ShimResource>> resources;
  std::vector<ShimEvent> log;
  struct Write { int set, binding, resource; uint64_t range; };
  std::vector<Write> writes;
  int nextSet = 0;
  ShimResource* add(int kind, uint64_t bytes, uint32_t w, uint32_t h, int format, int mips = 1) { resources.emplace_back(new ShimResource{kind, (int)resources.size(), bytes, w, h, format, mips}); return resources.back().get(); }
  static ShimDevice& get() { static ShimDevice d; return d; }
};
enum VkStructureType { VK_STRUCTURE_TYPE_SAMPLER_CREATE_INFO = 31, VK_STRUCTURE_TYPE_BUFFER_MEMORY_BARRIER = 44 };
enum VkFilter { VK_FILTER_NEAREST = 0, VK_FILTER_LINEAR = 1 };
enum VkSamplerMipmapMode { VK_SAMPLER_MIPMAP_MODE_NEAREST = 0, VK_SAMPLER_MIPMAP_MODE_LINEAR = 1 };
enum VkSamplerAddressMode { VK_SAMPLER_ADDRESS_MODE_REPEAT = 0, VK_SAMPLER_ADDRESS_MODE_MIRRORED_REPEAT = 1, VK_SAMPLER_ADDRESS_MODE_CLAMP_TO_EDGE = 2 };
enum VkFormat { VK_FORMAT_X8_D24_UNORM_PACK32 = 125, VK_FORMAT_R8G8B8A8_UNORM = 37, VK_FORMAT_B8G8R8A8_UNORM = 44, VK_FORMAT_R16G16_SINT = 82, VK_FORMAT_R16G16_SFLOAT = 83, VK_FORMAT_R32G32B32A32_UINT = 107, VK_FORMAT_R32G32B32A32_SFLOAT = 109 };
enum VkDescriptorType { VK_DESCRIPTOR_TYPE_COMBINED_IMAGE_SAMPLER = 1, VK_DESCRIPTOR_TYPE_STORAGE_IMAGE = 3, VK_DESCRIPTOR_TYPE_UNIFORM_BUFFER = 6, VK_DESCRIPTOR_TYPE_STORAGE_BUFFER = 7 };
enum { VK_BUFFER_USAGE_TRANSFER_DST_BIT = 2, VK_BUFFER_USAGE_UNIFORM_BUFFER_BIT = 0x10, VK_BUFFER_USAGE_STORAGE_BUFFER_BIT = 0x20,
       VK_BUFFER_USAGE_SHADER_DEVICE_ADDRESS_BIT = 0x20000, VK_BUFFER_USAGE_ACCELERATION_STRUCTURE_BUILD_INPUT_READ_ONLY_BIT_KHR = 0x80000,
       VK_MEMORY_PROPERTY_DEVICE_LOCAL_BIT = 1, VK_IMAGE_USAGE_SAMPLED_BIT = 4, VK_COMMAND_POOL_CREATE_TRANSIENT_BIT = 1,
       VK_SHADER_STAGE_FRAGMENT_BIT = 0x10, VK_SHADER_STAGE_COMPUTE_BIT = 0x20, VK_SHADER_STAGE_RAYGEN_BIT_KHR = 0x100, VK_SHADER_STAGE_ANY_HIT_BIT_KHR = 0x200,
       VK_SHADER_STAGE_CLOSEST_HIT_BIT_KHR = 0x400, VK_ACCESS_SHADER_READ_BIT = 0x20, VK_ACCESS_TRANSFER_WRITE_BIT = 0x1000,
       VK_PIPELINE_STAGE_VERTEX_SHADER_BIT = 8, VK_PIPELINE_STAGE_TRANSFER_BIT = 0x1000, VK_PIPELINE_STAGE_RAY_TRACING_SHADER_BIT_KHR = 0x200000,
       VK_DEPENDENCY_DEVICE_GROUP_BIT = 4 };
struct VkSamplerCreateInfo { VkStructureType sType; const void* pNext; VkFlags flags; VkFilter magFilter, minFilter; VkSamplerMipmapMode mipmapMode;
                             VkSamplerAddressMode addressModeU, addressModeV, addressModeW; float mipLodBias, maxAnisotropy, minLod, maxLod; };
struct VkImageCreateInfo { VkExtent2D extent; VkFormat format; uint32_t mipLevels; };
struct VkImageViewCreateInfo { VkImage image; };                          // image = the ShimResource of the image
struct VkDescriptorBufferInfo { VkBuffer buffer; VkDeviceSize offset, range; };
struct VkDescriptorImageInfo { VkSampler sampler; VkImageView imageView; int imageLayout; };     // imageView = the ShimResource of the image
struct VkWriteDescriptorSet { int unused; };
struct VkBufferMemoryBarrier { VkStructureType sType; const void* pNext; VkFlags srcAccessMask, dstAccessMask; uint32_t srcQueueFamilyIndex, dstQueueFamilyIndex;
                               VkBuffer buffer; VkDeviceSize offset, size; };
inline void vkDestroyDescriptorPool(VkDevice, VkDescriptorPool, const void*) {}
inline void vkDestroyDescriptorSetLayout(VkDevice, VkDescriptorSetLayout, const void*) {}
inline void vkDestroyImageView(VkDevice, VkImageView, const void*) {}
inline void vkUpdateDescriptorSets(VkDevice, uint32_t, const VkWriteDescriptorSet*, uint32_t, const void*) {}
inline void vkCmdPipelineBarrier(VkCommandBuffer, VkFlags, VkFlags, VkFlags, uint32_t, const void*, uint32_t, const VkBufferMemoryBarrier*, uint32_t, const void*) {}
// vkCmdUpdateBuffer(cmdBuf, deviceUBO, 0, sizeof(SceneCamera), &m_camera): the one Vulkan call with an effect the tests read — defined in ref_scene.cpp
void vkCmdUpdateBuffer(VkCommandBuffer, VkBuffer dst, VkDeviceSize offset, VkDeviceSize size, const void* data);
inline void vkCmdBlitImage(...) {}
inline void vkDestroyPipeline(VkDevice, VkPipeline, const void*) {}
inline void vkDestroyPipelineLayout(VkDevice, VkPipelineLayout, const void*) {}
inline void vkDestroyShaderModule(VkDevice, VkShaderModule, const void*) {}
inline void vkCreatePipelineLayout(VkDevice, const VkPipelineLayoutCreateInfo* ci, const void*, VkPipelineLayout* out) {
  ShimEvent e{0, (int)ci->setLayoutCount, (int)ci->pushConstantRangeCount, ci->pushConstantRangeCount ? (int)ci->pPushConstantRanges[0].size : 0, {}};
  ShimDevice::get().log.push_back(e); *out = (VkPipelineLayout)1;
}
inline void vkCreateComputePipelines(VkDevice, void*, uint32_t, const VkComputePipelineCreateInfo* ci, const void*, VkPipeline* out) { *out = new VkPipeline_T{(int)(intptr_t)ci->stage.module}; }
inline void vkCmdBindDescriptorSets(VkCommandBuffer, VkPipelineBindPoint, VkPipelineLayout, uint32_t first, uint32_t n, const VkDescriptorSet* sets, uint32_t, const void*) {
  ShimEvent e{1, (int)first, (int)n, n ? (int)(intptr_t)sets[n - 1] : -1, {}}; ShimDevice::get().log.push_back(e);   // c = the LAST set bound = the renderer's own (S_RAYQ)
}
inline void vkCmdPushConstants(VkCommandBuffer, VkPipelineLayout, VkFlags, uint32_t offset, uint32_t size, const void* data) {
  ShimEvent e{2, (int)offset, (int)size, 0, std::vector<unsigned char>((const unsigned char*)data, (const unsigned char*)data + size)}; ShimDevice::get().log.push_back(e);
}
inline void vkCmdBindPipeline(VkCommandBuffer, VkPipelineBindPoint, VkPipeline p) { ShimEvent e{3, p ? p->tag : -1, 0, 0, {}}; ShimDevice::get().log.push_back(e); }
inline void vkCmdDraw(VkCommandBuffer, uint32_t vertices, uint32_t instances, uint32_t firstVertex, uint32_t firstInstance) { ShimEvent e{6, (int)vertices, (int)instances, (int)(firstVertex + firstInstance), {}}; ShimDevice::get().log.push_back(e); }
inline void vkCmdDispatch(VkCommandBuffer, uint32_t x, uint32_t y, uint32_t z) { ShimEvent e{4, (int)x, (int)y, (int)z, {}}; ShimDevice::get().log.push_back(e); }

// ---- VK_KHR_acceleration_structure names (accelstruct.cpp) -----------------------------------------------------------------------
typedef void* VkAccelerationStructureKHR; typedef VkFlags VkGeometryInstanceFlagsKHR; typedef VkFlags VkBuildAccelerationStructureFlagsKHR;
enum { VK_STRUCTURE_TYPE_BUFFER_DEVICE_ADDRESS_INFO = 1000244001, VK_STRUCTURE_TYPE_ACCELERATION_STRUCTURE_GEOMETRY_TRIANGLES_DATA_KHR = 1000150005,
       VK_STRUCTURE_TYPE_ACCELERATION_STRUCTURE_GEOMETRY_KHR = 1000150006, VK_STRUCTURE_TYPE_WRITE_DESCRIPTOR_SET_ACCELERATION_STRUCTURE_KHR = 1000150007,
       VK_FORMAT_R32G32B32_SFLOAT = 106, VK_INDEX_TYPE_UINT32 = 1, VK_GEOMETRY_TYPE_TRIANGLES_KHR = 0,
       VK_GEOMETRY_OPAQUE_BIT_KHR = 1, VK_GEOMETRY_NO_DUPLICATE_ANY_HIT_INVOCATION_BIT_KHR = 2,
       VK_GEOMETRY_INSTANCE_TRIANGLE_FACING_CULL_DISABLE_BIT_KHR = 1, VK_GEOMETRY_INSTANCE_FORCE_OPAQUE_BIT_KHR = 4,
       VK_BUILD_ACCELERATION_STRUCTURE_ALLOW_COMPACTION_BIT_KHR = 2, VK_BUILD_ACCELERATION_STRUCTURE_PREFER_FAST_TRACE_BIT_KHR = 4,
       VK_DESCRIPTOR_TYPE_ACCELERATION_STRUCTURE_KHR = 1000150000 };
struct VkBufferDeviceAddressInfo { int sType; const void* pNext; VkBuffer buffer; };
struct VkDeviceOrHostAddressConstKHR { VkDeviceAddress deviceAddress; };
struct VkAccelerationStructureGeometryTrianglesDataKHR { int sType; const void* pNext; int vertexFormat; VkDeviceOrHostAddressConstKHR vertexData; VkDeviceSize vertexStride;
                                                         uint32_t maxVertex; int indexType; VkDeviceOrHostAddressConstKHR indexData, transformData; };
struct VkAccelerationStructureGeometryDataKHR { VkAccelerationStructureGeometryTrianglesDataKHR triangles; };
struct VkAccelerationStructureGeometryKHR { int sType; const void* pNext; int geometryType; VkAccelerationStructureGeometryDataKHR geometry; VkFlags flags; };
struct VkAccelerationStructureBuildRangeInfoKHR { uint32_t primitiveCount, primitiveOffset, firstVertex, transformOffset; };
struct VkTransformMatrixKHR { float matrix[3][4]; };
struct VkAccelerationStructureInstanceKHR { VkTransformMatrixKHR transform; uint32_t instanceCustomIndex : 24; uint32_t mask : 8;
                                            uint32_t instanceShaderBindingTableRecordOffset : 24; uint32_t flags : 8; uint64_t accelerationStructureReference; };
struct VkWriteDescriptorSetAccelerationStructureKHR { int sType; const void* pNext; uint32_t accelerationStructureCount; const VkAccelerationStructureKHR* pAccelerationStructures; };
VkDeviceAddress vkGetBufferDeviceAddress(VkDevice, const VkBufferDeviceAddressInfo* info);   // defined after VkBuffer_T

// ---- nvmath (un-vendored): contract arithmetic ---------------------------------------------------------------------------------
namespace nvmath {
template <class T> struct vector4;
template <class T> struct vector2 {
  T x, y;
  vector2() : x(0), y(0) {}
  vector2(T a, T b) : x(a), y(b) {}
  T& operator[](int i) { return (&x)[i]; }
  const T& operator[](int i) const { return (&x)[i]; }
};
template <class T> vector2<T> operator/(const vector2<T>& a, T s) { return vector2<T>(a.x / s, a.y / s); }   // ivec2 / 2: C integer division, like nvmath
template <class T> struct vector3 {
  T x, y, z;
  vector3() : x(0), y(0), z(0) {}
  vector3(T a, T b, T c) : x(a), y(b), z(c) {}
  explicit vector3(T a) : x(a), y(a), z(a) {}
  vector3(const vector4<T>& v);
  T& operator[](int i) { return (&x)[i]; }
  const T& operator[](int i) const { return (&x)[i]; }
};
template <class T> struct vector4 {
  T x, y, z, w;
  vector4() : x(0), y(0), z(0), w(0) {}
  vector4(T a, T b, T c, T d) : x(a), y(b), z(c), w(d) {}
  vector4(const vector3<T>& v, T d) : x(v.x), y(v.y), z(v.z), w(d) {}
  T& operator[](int i) { return (&x)[i]; }
  const T& operator[](int i) const { return (&x)[i]; }
};
template <class T> vector3<T>::vector3(const vector4<T>& v) : x(v.x), y(v.y), z(v.z) {}
template <class T> struct matrix4 {   // column-major storage, aRC = row R, column C (nvmath's naming)
  union { T m[16]; struct { T a00, a10, a20, a30, a01, a11, a21, a31, a02, a12, a22, a32, a03, a13, a23, a33; }; };
  matrix4() { for (int i = 0; i < 16; ++i) m[i] = T(0); }
};
typedef vector2<int> vec2i; typedef vector2<float> vec2f; typedef vector2<unsigned int> vec2ui;
typedef vector3<float> vec3f; typedef vector4<float> vec4f; typedef vector4<unsigned int> vec4ui;
typedef matrix4<float> mat4f;
inline eidc::eid_mat4 toEid(const mat4f& a) { eidc::eid_mat4 r; memcpy(r.m, a.m, 64); return r; }
inline mat4f fromEid(const eidc::eid_mat4& a) { mat4f r; memcpy(r.m, a.m, 64); return r; }
inline vec4f operator*(const mat4f& M, const vec4f& v) { eidc::eid_vec4 r = eidc::eid_mat4_mulv(toEid(M), eidc::eid_vec4{v.x, v.y, v.z, v.w}); return vec4f(r.x, r.y, r.z, r.w); }
inline mat4f operator*(const mat4f& a, const mat4f& b) { return fromEid(eidc::eid_mat4_mul(toEid(a), toEid(b))); }
inline mat4f invert(const mat4f& a) { return fromEid(eidc::eid_mat4_invert(toEid(a))); }
inline mat4f perspectiveVK(float fovDeg, float aspect, float n, float f) { return fromEid(eidc::eid_perspectiveVK(fovDeg, aspect, n, f)); }
inline vec3f operator-(const vec3f& a, const vec3f& b) { return vec3f(a.x - b.x, a.y - b.y, a.z - b.z); }
inline float length(const vec3f& a) { return sqrtf(a.x * a.x + a.y * a.y + a.z * a.z); }
template <class T> T nv_clamp(T v, T lo, T hi) { return v < lo ? lo : (v > hi ? hi : v); }
template <class T> T nv_random() { return T(0); }
inline vec3f normalize(const vec3f& u) {
  float norm = sqrtf(u.x * u.x + u.y * u.y + u.z * u.z);
  norm = (norm > 1e-6f) ? 1.0f / norm : 0.0f;
  return vec3f(u.x * norm, u.y * norm, u.z * norm);
}
}  // namespace nvmath
using nvmath::normalize;
inline double rad2deg(double a) { return (double)((float)a * 57.29577951308232f); }   // contract: fp32 product, like the product's host side

// ---- nvh -----------------------------------------------------------------------------------------------------------------------
#define LOGI(...) do { } while (0)
#define LOGW(...) do { } while (0)
#define LOGE(...) do { } while (0)
namespace tinygltf {
struct Image { int width = -1, height = -1; std::vector<unsigned char> image; };
struct Texture { int sampler = -1, source = -1; };
struct Sampler { int minFilter = -1, magFilter = -1, wrapS = 10497, wrapT = 10497; };
struct SpotLight { double innerConeAngle = 0.0, outerConeAngle = 0.7853981634; };
struct Light { std::string type; std::vector<double> color; double intensity = 1.0, range = 0.0; SpotLight spot; };
struct PerspectiveCamera { double yfov = 0.0; };
struct Camera { PerspectiveCamera perspective; };
struct Model { std::vector<Image> images; std::vector<Texture> textures; std::vector<Sampler> samplers; };
// the file loader: Scene::loadGltfScene only needs it to report success; the model arrives through ref_scene_inject()
struct TinyGLTF {
  void RemoveImageLoader() {}
  template <class F> void SetImageLoader(F, void*) {}
  bool LoadASCIIFromFile(Model* m, std::string*, std::string*, const std::string&);
  bool LoadBinaryFromFile(Model* m, std::string*, std::string*, const std::string&);
};
inline void loadExternalImages(Model*, const std::string&) {}
inline bool LoadFreeImageData(...) { return true; }
}  // namespace tinygltf
namespace nvh {
struct Stopwatch { double elapsed() { return 0.0; } };
struct GltfMaterial {
  int shadingModel = 0;
  nvmath::vec4f baseColorFactor{1, 1, 1, 1}; int baseColorTexture = -1; float metallicFactor = 1.f, roughnessFactor = 1.f; int metallicRoughnessTexture = -1;
  int emissiveTexture = -1; nvmath::vec3f emissiveFactor{0, 0, 0}; int alphaMode = 0; float alphaCutoff = 0.5f; int doubleSided = 0;
  int normalTexture = -1; float normalTextureScale = 1.f;
  struct { float factor = 0.f; int texture = -1; } transmission;
  struct { float ior = 1.5f; } ior;
};
#define MATERIAL_SPECULARGLOSSINESS 1
struct GltfPrimMesh { uint32_t firstIndex = 0, indexCount = 0, vertexOffset = 0, vertexCount = 0; int materialIndex = 0; std::string name; };
struct GltfNode { nvmath::mat4f worldMatrix; int primMesh = 0; };
struct GltfLight { nvmath::mat4f worldMatrix; tinygltf::Light light; };
struct GltfCamera { nvmath::mat4f worldMatrix; nvmath::vec3f eye, center, up; tinygltf::Camera cam; };
struct GltfStats { uint32_t nbCameras = 0, nbImages = 0, nbTextures = 0, nbMaterials = 0, nbSamplers = 0, nbNodes = 0, nbMeshes = 0, nbLights = 0; };
enum class GltfAttributes : uint8_t { Position = 0, Normal = 1, Texcoord_0 = 2, Texcoord_1 = 4, Tangent = 8, Color_0 = 16 };
inline GltfAttributes operator|(GltfAttributes a, GltfAttributes b) { return (GltfAttributes)((uint8_t)a | (uint8_t)b); }
struct GltfScene {
  std::vector<GltfMaterial> m_materials; std::vector<GltfNode> m_nodes; std::vector<GltfPrimMesh> m_primMeshes; std::vector<GltfCamera> m_cameras;
  std::vector<GltfLight> m_lights;
  std::vector<nvmath::vec3f> m_positions; std::vector<uint32_t> m_indices; std::vector<nvmath::vec3f> m_normals; std::vector<nvmath::vec4f> m_tangents;
  std::vector<nvmath::vec2f> m_texcoords0; std::vector<nvmath::vec4f> m_colors0;
  struct { nvmath::vec3f min, max; } m_dimensions;
  GltfStats getStatistics(const tinygltf::Model&) { return GltfStats(); }
  void importMaterials(const tinygltf::Model&);                          // nvpro_core: here, copies the injected scene (ref_scene.cpp)
  void importDrawableNodes(const tinygltf::Model&, GltfAttributes);
};
// CameraManip: the global camera manipulator (nvh/cameramanipulator.hpp)
struct CameraManipulator {
  struct Camera { nvmath::vec3f eye{10, 10, 10}, ctr{0, 0, 0}, up{0, 1, 0}; float fov = 60.0f; };
  Camera cam;
  void setCamera(Camera c) { cam = c; }
  void setLookat(const nvmath::vec3f& e, const nvmath::vec3f& c, const nvmath::vec3f& u) { cam.eye = e; cam.ctr = c; cam.up = u; }
  void setFov(float f) { cam.fov = f; }
  nvmath::mat4f getMatrix() const { return nvmath::fromEid(eidc::eid_look_at(eidc::eid_vec3{cam.eye.x, cam.eye.y, cam.eye.z}, eidc::eid_vec3{cam.ctr.x, cam.ctr.y, cam.ctr.z}, eidc::eid_vec3{cam.up.x, cam.up.y, cam.up.z})); }
  float getFov() const { return cam.fov; }
  void getLookat(nvmath::vec3f& e, nvmath::vec3f& c, nvmath::vec3f& u) const { e = cam.eye; c = cam.ctr; u = cam.up; }
  void fit(const nvmath::vec3f&, const nvmath::vec3f&, bool) {}         // no glTF camera: the harness sets the look-at explicitly
  static CameraManipulator& Singleton() { static CameraManipulator s; return s; }
};
}  // namespace nvh
#define CameraManip nvh::CameraManipulator::Singleton()
namespace ImGuiH {
inline void SetCameraJsonFile(const std::string&) {}
inline void SetHomeCamera(const nvh::CameraManipulator::Camera&) {}
inline void AddCamera(const nvh::CameraManipulator::Camera&) {}
}

// ---- nvvk ----------------------------------------------------------------------------------------------------------------------
struct VkBuffer_T { std::vector<unsigned char> bytes; ShimResource* res = nullptr; };   // a "buffer" is the host copy of what was uploaded into it
inline VkDeviceAddress vkGetBufferDeviceAddress(VkDevice, const VkBufferDeviceAddressInfo* info) { return (VkDeviceAddress)(uintptr_t)(info->buffer ? info->buffer->bytes.data() : nullptr); }
namespace nvvk {
struct Image { VkImage image = nullptr; };                                // image = ShimResource* when created through createImage(info)
struct Texture { VkImage image = nullptr; VkDescriptorImageInfo descriptor{}; };
struct Buffer { VkBuffer buffer = nullptr; };
inline VkDeviceAddress getBufferDeviceAddress(VkDevice, VkBuffer b) { return (VkDeviceAddress)(uintptr_t)(b ? b->bytes.data() : nullptr); }   // the shaders' buffer_reference
// nvvk::mipLevels (nvpro_core): floor(log2(max(w, h))) + 1 — contract
inline uint32_t mipLevels(VkExtent2D e) { uint32_t m = e.width > e.height ? e.width : e.height, n = 1; while (m > 1) { m >>= 1; ++n; } return n; }
inline VkImageCreateInfo makeImage2DCreateInfo(VkExtent2D e, VkFormat f = VK_FORMAT_R8G8B8A8_UNORM, VkFlags = 0, bool mipmaps = false) { return VkImageCreateInfo{e, f, mipmaps ? mipLevels(e) : 1u}; }
inline VkFormat findDepthFormat(VkPhysicalDevice) { return VK_FORMAT_X8_D24_UNORM_PACK32; }
inline void cmdBarrierImageLayout(VkCommandBuffer, VkImage, VkImageLayout, VkImageLayout) {}
inline VkShaderModule createShaderModule(VkDevice, const uint32_t* code, size_t) { return (VkShaderModule)(intptr_t)code[0]; }   // the stand-in "SPIR-V" is one word: the shader's tag
struct ProfilerVK {};
struct GraphicsPipelineGeneratorCombined {            // nvvk/pipeline_vk.hpp: the pipeline is the tag of its fragment shader
  struct { int cullMode = 0; } rasterizationState;
  int fragTag = -1;
  GraphicsPipelineGeneratorCombined(VkDevice, VkPipelineLayout, VkRenderPass) {}
  void addShader(const std::vector<uint32_t>& code, int stage) { if (stage == VK_SHADER_STAGE_FRAGMENT_BIT && !code.empty()) fragTag = (int)code[0]; }
  VkPipeline createPipeline() { return new VkPipeline_T{fragTag}; }
};
inline VkImageViewCreateInfo makeImageViewCreateInfo(VkImage i, const VkImageCreateInfo&) { return VkImageViewCreateInfo{i}; }
inline void cmdGenerateMipmaps(VkCommandBuffer, VkImage image, VkFormat, VkExtent2D size, uint32_t levels, uint32_t layers = 1, VkImageLayout = VK_IMAGE_LAYOUT_GENERAL) {
  ShimEvent e{5, image ? ((ShimResource*)image)->id : -1, (int)levels, (int)layers, {}}; (void)size; ShimDevice::get().log.push_back(e);
}
class ResourceAllocator {
public:
  std::vector<std::unique_ptr<VkBuffer_T>> owned;
  Buffer make(const void* p, size_t n) { owned.emplace_back(new VkBuffer_T); if (p) owned.back()->bytes.assign((const unsigned char*)p, (const unsigned char*)p + n); else owned.back()->bytes.assign(n, 0); return Buffer{owned.back().get()}; }
  void destroy(Texture&) {}
  void destroy(Image&) {}
  void destroy(Buffer&) {}
  Image createImage(VkCommandBuffer, VkDeviceSize, const void*, const VkImageCreateInfo&) { return Image(); }
  Image createImage(const VkImageCreateInfo& ci) { const uint64_t texel = ci.format == VK_FORMAT_R16G16_SINT || ci.format == VK_FORMAT_R16G16_SFLOAT ? 4 : 16; return Image{(VkImage)ShimDevice::get().add(1, texel * ci.extent.width * ci.extent.height, ci.extent.width, ci.extent.height, (int)ci.format, (int)ci.mipLevels)}; }
  Texture createTexture(const Image& im, const VkImageViewCreateInfo& iv) { Texture t; t.image = im.image; t.descriptor.imageView = (VkImageView)iv.image; return t; }
  Texture createTexture(const Image& im, const VkImageViewCreateInfo& iv, const VkSamplerCreateInfo&) { Texture t; t.image = im.image; t.descriptor.imageView = (VkImageView)iv.image; return t; }
  Texture createTexture(VkCommandBuffer, VkDeviceSize, const void*, const VkImageCreateInfo&, const VkSamplerCreateInfo&) { return Texture(); }
  template <class T> Buffer createBuffer(VkCommandBuffer, const std::vector<T>& v, VkFlags) { return make(v.data(), v.size() * sizeof(T)); }
  Buffer createBuffer(VkCommandBuffer, VkDeviceSize n, const void* p, VkFlags) { return make(p, (size_t)n); }
  Buffer createBuffer(VkDeviceSize n, VkFlags, VkFlags = 0) { Buffer b = make(nullptr, (size_t)n); b.buffer->res = ShimDevice::get().add(0, n, 0, 0, 0); return b; }
  void finalizeAndReleaseStaging() {}
};
struct DebugUtil { void setup(VkDevice) {} template <class T> void setObjectName(T, const std::string&) {} template <class T> void setObjectName(T, const char*) {} };
struct CommandPool {
  CommandPool(VkDevice, uint32_t, VkFlags = 0, VkQueue = nullptr) {}
  VkCommandBuffer createCommandBuffer() { return nullptr; }
  void submitAndWait(VkCommandBuffer) {}
};
struct DescriptorSetBindings {
  struct Binding { int binding; int type; uint32_t count; VkShaderStageFlags flags; };
  void addBinding(Binding) {}
  VkDescriptorPool createPool(VkDevice, uint32_t = 1) { return nullptr; }
  VkWriteDescriptorSet makeWrite(VkDescriptorSet, int, const VkWriteDescriptorSetAccelerationStructureKHR*) { return VkWriteDescriptorSet(); }
  VkDescriptorSetLayout createLayout(VkDevice) { return nullptr; }
  VkWriteDescriptorSet makeWrite(VkDescriptorSet s, int binding, const VkDescriptorBufferInfo* b) {
    ShimDevice::get().writes.push_back({(int)(intptr_t)s, binding, b && b->buffer && b->buffer->res ? b->buffer->res->id : -1, b ? b->range : 0}); return VkWriteDescriptorSet();
  }
  VkWriteDescriptorSet makeWrite(VkDescriptorSet s, int binding, const VkDescriptorImageInfo* im) {
    ShimDevice::get().writes.push_back({(int)(intptr_t)s, binding, im && im->imageView ? ((ShimResource*)im->imageView)->id : -1, 0}); return VkWriteDescriptorSet();
  }
  VkWriteDescriptorSet makeWriteArray(VkDescriptorSet, int, const void*) { return VkWriteDescriptorSet(); }
};
// nvvk::RaytracingBuilderKHR (nvpro_core): here it only keeps what it is asked to build
inline VkTransformMatrixKHR toTransformMatrixKHR(const nvmath::mat4f& m) {   // column-major 4x4 -> row-major 3x4
  VkTransformMatrixKHR t;
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) t.matrix[r][c] = m.m[c * 4 + r];
  return t;
}
class RaytracingBuilderKHR {
public:
  struct BlasInput { std::vector<VkAccelerationStructureGeometryKHR> asGeometry; std::vector<VkAccelerationStructureBuildRangeInfoKHR> asBuildOffsetInfo; VkFlags flags = 0; };
  std::vector<BlasInput> blas; VkFlags blasFlags = 0; std::vector<VkAccelerationStructureInstanceKHR> tlas; VkFlags tlasFlags = 0;
  void setup(VkDevice, ResourceAllocator*, uint32_t) {}
  void destroy() { blas.clear(); tlas.clear(); }
  void buildBlas(const std::vector<BlasInput>& in, VkFlags f) { blas = in; blasFlags = f; }
  void buildTlas(const std::vector<VkAccelerationStructureInstanceKHR>& in, VkFlags f) { tlas = in; tlasFlags = f; }
  VkDeviceAddress getBlasDeviceAddress(uint32_t i) { return 0x1000u + i; }
  VkAccelerationStructureKHR getAccelerationStructure() { return (VkAccelerationStructureKHR)this; }
};
inline VkDescriptorSet allocateDescriptorSet(VkDevice, VkDescriptorPool, VkDescriptorSetLayout) { return (VkDescriptorSet)(intptr_t)(++ShimDevice::get().nextSet); }   // sets are numbered 1, 2, ...
}  // namespace nvvk
#define NAME_VK(x) do { } while (0)
#define NAME_IDX_VK(x, i) do { } while (0)
#define NAMED_VK(x) x
#define CREATE_NAMED_VK(dst, expr) dst = (expr)

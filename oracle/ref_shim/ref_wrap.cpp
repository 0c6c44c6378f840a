/*
 * oracle/ref_shim/ref_wrap.cpp — TEST INFRASTRUCTURE.
 * Compiles the three stand-alone-buildable files of the reference *where they lie* under
 * /root/reference (include paths only; nothing is copied) and exposes them to ctypes, so the
 * oracle's restatement can be checked against the reference's own code:
 *   shaders/host_device.h  -> sizeof of every shared struct
 *   shaders/compress.glsl  -> compress_unit_vec / decompress_unit_vec / packUnorm4x8 (C++ branch)
 *   src/alias_table.hpp    -> DiscreteSampler1D<float>
 *   src/hdr_sampling.cpp   -> HdrSampling::createEnvironmentAccel / buildAliasmap (the environment alias map, integral, average);
 *                             the file is compiled whole against inert Vulkan / nvvk / stb stand-ins (vk_shim.h, nvvk/, nvh/, stb_image.h)
 */
#include "nvmath/nvmath.h"
#include <string>
#include "host_device.h"   // /root/reference/shaders/host_device.h
#include "compress.glsl"   // /root/reference/shaders/compress.glsl
#include "alias_table.hpp" // /root/reference/src/alias_table.hpp

#include <array>
#include <chrono>
#include <cstdio>
#include <ios>
#include <sstream>
#include <vector>
#define private public     // createEnvironmentAccel / buildAliasmap are private members (std headers are already in, above)
#include "hdr_sampling.hpp" // /root/reference/src/hdr_sampling.hpp
#undef private

#define REF_API extern "C" __attribute__((visibility("default")))

REF_API void ref_env_accel(const float* rgba, uint32_t w, uint32_t h, ImptSampData* accel, float* integral, float* average) {
  HdrSampling hs;
  VkExtent2D size{w, h};
  std::vector<ImptSampData> a = hs.createEnvironmentAccel(rgba, size);
  for (size_t i = 0; i < a.size(); ++i) accel[i] = a[i];
  *integral = hs.getIntegral(); *average = hs.getAverage();
}

REF_API uint32_t ref_compress_unit_vec(float x, float y, float z) { return compress_unit_vec(vec3(x, y, z)); }
REF_API void ref_decompress_unit_vec(uint32_t p, float* out) { vec3 v = decompress_unit_vec(p); out[0] = v.x; out[1] = v.y; out[2] = v.z; }
REF_API uint32_t ref_pack_unorm4x8(const float* v) { vec4 q; q.x = v[0]; q.y = v[1]; q.z = v[2]; q.w = v[3]; return packUnorm4x8(q); }
REF_API void ref_alias_table(const float* values, int n, float* prob, int* failId) {
  DiscreteSampler1D<float> t(std::vector<float>(values, values + n));
  for (int i = 0; i < n; ++i) { prob[i] = t.binomDistribs[i].prob; failId[i] = t.binomDistribs[i].failId; }
}
REF_API int ref_sizeof(const char* name) {
  std::string s(name);
#define SZ(T) if (s == #T) return (int)sizeof(T);
  SZ(SceneCamera) SZ(VertexAttributes) SZ(GltfShadeMaterial) SZ(RtxState) SZ(InstanceData) SZ(LightSample) SZ(GISample)
  SZ(DirectReservoir) SZ(IndirectReservoir) SZ(ImptSampData) SZ(PuncLight) SZ(TrigLight) SZ(LightBufInfo) SZ(Tonemapper) SZ(SunAndSky)
#undef SZ
  return -1;
}

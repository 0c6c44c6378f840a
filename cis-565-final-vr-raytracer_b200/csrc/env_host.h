// env_host.h — host half of HdrSampling (reference src/hdr_sampling.{hpp,cpp}): the RGBA32F lat-long environment map and the
// per-texel alias table used to importance-sample it (https://arxiv.org/pdf/1901.05423.pdf).
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include "host_device.h"

namespace eid {

struct EnvHost {
  uint32_t width = 0, height = 0;
  std::vector<float> pixels;             // rgba32f, row 0 = +Y pole
  std::vector<ImptSampData> accel;       // one cell per texel
  float integral = 1.f;                  // HdrSampling::getIntegral(): sum of solid-angle-weighted max(r,g,b)
  float average = 1.f;                   // HdrSampling::getAverage(): mean Rec.709 luminance
  void build(const float* rgba, uint32_t w, uint32_t h);   // createEnvironmentAccel (hdr_sampling.cpp:181-242)
  void loadRadianceHdr(const std::string& path);           // stbi_loadf replacement (.hdr / RGBE), then build()
};

}  // namespace eid

struct eid_env {
  eid::EnvHost host;
  int device = 0;
  struct float4* tex = nullptr;          // device copy of pixels (RGBA32F)
  ImptSampData* accel = nullptr;
  // Renderers that were handed this environment (eid_renderer_set_env) keep it alive: eid_env_destroy on an environment still
  // in use only marks it, and the last renderer to let go frees it (HdrSampling::loadEnvironment is called again on a live object
  // by the reference's hosts, sample_example.cpp:97-106, while the renderer still holds the previous map).
  int users = 0;
  bool destroyRequested = false;
};

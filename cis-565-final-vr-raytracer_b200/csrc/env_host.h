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
};

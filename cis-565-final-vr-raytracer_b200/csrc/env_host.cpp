// env_host.cpp — environment importance-sampling tables (hdr_sampling.cpp:107-242) and a Radiance RGBE reader that stands in
// for stbi_loadf (hdr_sampling.cpp:64; stb is un-vendored third-party code of the reference).
#include "env_host.h"
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include "common.h"
#include "pack.h"

namespace eid {

// buildAliasmap (hdr_sampling.cpp:107-176): texels below the average are paired with texels above it; a bright texel keeps
// absorbing dim ones until its own ratio drops below 1.  Same scan order as the reference, so (q, alias) match bit for bit.
static float aliasMap(const std::vector<float>& w, std::vector<ImptSampData>& cells) {
  const uint32_t n = (uint32_t)w.size();
  float sum = 0.f;
  for (float v : w) sum += v;
  const float invAvg = static_cast<float>(n) / sum;
  for (uint32_t i = 0; i < n; ++i) { cells[i].q = w[i] * invAvg; cells[i].alias = (int32_t)i; }
  std::vector<uint32_t> part(n);
  uint32_t lo = 0, hi = n;
  for (uint32_t i = 0; i < n; ++i) { if (cells[i].q < 1.f) part[lo++] = i; else part[--hi] = i; }
  for (uint32_t s = 0; s < hi && hi < n; ++s) {
    const uint32_t dim = part[s], bright = part[hi];
    cells[dim].alias = (int32_t)bright;
    cells[bright].q -= 1.f - cells[dim].q;
    if (cells[bright].q < 1.0f) ++hi;
  }
  return sum;
}

void EnvHost::build(const float* rgba, uint32_t w, uint32_t h) {
  if (!rgba || !w || !h || (uint64_t)w * h > (1u << 28)) raise(EID_ERR_INVALID, "bad environment map %ux%u", w, h);
  width = w; height = h;
  pixels.assign(rgba, rgba + 4 * (size_t)w * h);
  const size_t n = (size_t)w * h;
  accel.assign(n, ImptSampData{});
  std::vector<float> importance(n);
  const float stepPhi = float(2.0 * M_PI) / float(w), stepTheta = float(M_PI) / float(h);
  float cosPrev = 1.0f;
  double lumSum = 0;
  for (uint32_t y = 0; y < h; ++y) {
    const float cosNext = std::cos(float(y + 1) * stepTheta);
    const float solidAngle = (cosPrev - cosNext) * stepPhi;
    cosPrev = cosNext;
    for (uint32_t x = 0; x < w; ++x) {
      const float* p = &pixels[4 * ((size_t)y * w + x)];
      importance[(size_t)y * w + x] = solidAngle * std::max(p[0], std::max(p[1], p[2]));
      lumSum += lum709(p[0], p[1], p[2]);
    }
  }
  average = static_cast<float>(lumSum) / static_cast<float>(w * h);
  integral = aliasMap(importance, accel);
  const float inv = 1.0f / integral;
  for (size_t i = 0; i < n; ++i) { const float* p = &pixels[4 * i]; accel[i].pdf = std::max(p[0], std::max(p[1], p[2])) * inv; }
  for (size_t i = 0; i < n; ++i) accel[i].aliasPdf = accel[accel[i].alias].pdf;
}

// Radiance .hdr (RGBE, "-Y h +X w", flat or new-style RLE scanlines) -> RGBA32F with alpha 1, like stbi_loadf(..., STBI_rgb_alpha)
void EnvHost::loadRadianceHdr(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) raise(EID_ERR_IO, "cannot open '%s'", path.c_str());
  std::string line;
  std::getline(f, line);
  if (line.rfind("#?", 0) != 0) raise(EID_ERR_PARSE, "'%s' is not a Radiance .hdr file", path.c_str());
  bool rgbe = false;
  while (std::getline(f, line)) {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    if (line.empty()) break;
    if (line.find("FORMAT=32-bit_rle_rgbe") != std::string::npos) rgbe = true;
  }
  if (!rgbe) raise(EID_ERR_UNSUPPORTED, "only FORMAT=32-bit_rle_rgbe .hdr files are supported");
  std::getline(f, line);
  int h = 0, w = 0;
  if (sscanf(line.c_str(), "-Y %d +X %d", &h, &w) != 2 || w <= 0 || h <= 0) raise(EID_ERR_UNSUPPORTED, "unsupported .hdr orientation '%s'", line.c_str());
  std::vector<float> out(4 * (size_t)w * h);
  std::vector<uint8_t> scan(4 * (size_t)w);
  auto decode = [](const uint8_t* p, float* o) {
    if (p[3]) { const float s = std::ldexp(1.0f, (int)p[3] - (128 + 8)); o[0] = p[0] * s; o[1] = p[1] * s; o[2] = p[2] * s; }
    else o[0] = o[1] = o[2] = 0.f;
    o[3] = 1.0f;
  };
  for (int y = 0; y < h; ++y) {
    uint8_t hd[4];
    f.read((char*)hd, 4);
    if (!f) raise(EID_ERR_PARSE, "truncated .hdr");
    if (w >= 8 && w < 32768 && hd[0] == 2 && hd[1] == 2 && !(hd[2] & 0x80) && ((hd[2] << 8) | hd[3]) == w) {
      for (int c = 0; c < 4; ++c) {                       // new RLE: each channel separately
        int x = 0;
        while (x < w) {
          int count = f.get();
          if (count < 0) raise(EID_ERR_PARSE, "truncated .hdr");
          if (count > 128) { count -= 128; int v = f.get(); if (x + count > w) raise(EID_ERR_PARSE, "bad .hdr run"); for (int k = 0; k < count; ++k) scan[4 * (size_t)(x++) + c] = (uint8_t)v; }
          else { if (!count || x + count > w) raise(EID_ERR_PARSE, "bad .hdr run"); for (int k = 0; k < count; ++k) scan[4 * (size_t)(x++) + c] = (uint8_t)f.get(); }
        }
      }
    } else {                                              // flat scanline: the 4 bytes read were the first pixel
      memcpy(scan.data(), hd, 4);
      f.read((char*)scan.data() + 4, 4 * (size_t)(w - 1));
      if (!f) raise(EID_ERR_PARSE, "truncated .hdr");
    }
    for (int x = 0; x < w; ++x) decode(&scan[4 * (size_t)x], &out[4 * ((size_t)y * w + x)]);
  }
  build(out.data(), (uint32_t)w, (uint32_t)h);
}

}  // namespace eid

/*
 * oracle/ref_shim/ref_display.cpp — TEST INFRASTRUCTURE.
 * The reference's display pass — shaders/post.frag with tonemapping.glsl and pcg3d of random.glsl — compiled WHOLE, main() included, as
 * C++ from the transliterations of glsl_prep.py and run once per pixel of the presented image (RenderOutput::run draws one full-screen
 * triangle, render_output.cpp:224-240).  What the fragment stage binds is provided here as plain globals: uvCoords = (pixel + 0.5) / size
 * (the interpolated attribute of passthrough.vert at the pixel centre), gl_FragCoord, fragColor, the push constant (Tonemapper +
 * debugging_mode) and the two samplers.  The samplers are the reference's: NEAREST / REPEAT (zero-initialised VkSamplerCreateInfo,
 * render_output.cpp:123-128), so texture() is a texel fetch; textureLod(img, vec2(0.5), 20) clamps to the last mip level, the 1x1 texel of
 * the chain RenderOutput::genMipmap blits — that chain is made by the Vulkan driver (vkCmdBlitImage), so its single texel is an INPUT
 * here (the contract's value).  toneLocalExposure (autoExposure bit 1, never selected by the reference's GUI, and not
 * valid C++ as written: an array constructor) is replaced by a trap.  Every other expression evaluated is the reference's own text.
 */
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include "glsl/glsl_builtins.h"
#include "host_device.h"          // /root/reference/shaders/host_device.h (C++ branch): Tonemapper, DebugMode

// integer / boolean vectors post.frag computes with (global scope, next to the GLSL_ built-ins they overload)
struct uvec3 {
  unsigned int x, y, z;
  uvec3() : x(0), y(0), z(0) {}
  explicit uvec3(unsigned int a) : x(a), y(a), z(a) {}
  uvec3(unsigned int a, unsigned int b, unsigned int c) : x(a), y(b), z(c) {}
  uvec3(orc::vec2 v, int c) : x((unsigned int)v.x), y((unsigned int)v.y), z((unsigned int)c) {}   // uvec3(gl_FragCoord.xy, 0): float -> uint truncates
};
struct bvec3 { bool x, y, z; };
inline uvec3 operator*(uvec3 a, unsigned int s) { return uvec3(a.x * s, a.y * s, a.z * s); }
inline uvec3 operator+(uvec3 a, uvec3 b) { return uvec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline uvec3 operator>>(uvec3 a, uvec3 b) { return uvec3(a.x >> b.x, a.y >> b.y, a.z >> b.z); }
inline uvec3 operator>>(uvec3 a, int s) { return uvec3(a.x >> s, a.y >> s, a.z >> s); }
inline uvec3 operator|(int s, uvec3 a) { return uvec3((unsigned int)s | a.x, (unsigned int)s | a.y, (unsigned int)s | a.z); }
inline uvec3& operator^=(uvec3& a, uvec3 b) { a.x ^= b.x; a.y ^= b.y; a.z ^= b.z; return a; }
inline orc::vec3 GLSL_uintBitsToFloat(uvec3 u) { return orc::vec3(orc::uintBitsToFloat(u.x), orc::uintBitsToFloat(u.y), orc::uintBitsToFloat(u.z)); }
inline bvec3 lessThan(orc::vec3 a, orc::vec3 b) { return bvec3{a.x < b.x, a.y < b.y, a.z < b.z}; }
inline orc::vec3 GLSL_mix(orc::vec3 x, orc::vec3 y, bvec3 a) { return orc::vec3(a.x ? y.x : x.x, a.y ? y.y : x.y, a.z ? y.z : x.z); }   // mix(x, y, bvec): selects
inline orc::vec3 GLSL_floor(orc::vec3 a) { return orc::vec3(eid_floorf(a.x), eid_floorf(a.y), eid_floorf(a.z)); }
inline orc::vec3 operator-(orc::vec3 a, float s) { return orc::vec3(a.x - s, a.y - s, a.z - s); }
inline orc::vec3 operator/(float s, orc::vec3 a) { return orc::vec3(s / a.x, s / a.y, s / a.z); }

namespace refdisplay {
using orc::vec2; using orc::vec3; using orc::vec4;

struct sampler2D { const vec4* data; int w, h, pitch; vec4 lastMip; };
static vec4 texture(const sampler2D& s, vec2 uv) {        // NEAREST, REPEAT: texel floor(uv * size) modulo size
  int x = (int)eid_floorf(uv.x * (float)s.w) % s.w, y = (int)eid_floorf(uv.y * (float)s.h) % s.h;
  if (x < 0) x += s.w;
  if (y < 0) y += s.h;
  return s.data[(size_t)y * s.pitch + x];
}
static vec4 textureLod(const sampler2D& s, vec2, float) { return s.lastMip; }   // only called as textureLod(img, vec2(0.5), 20): the 1x1 level

// the fragment stage's interface
static vec2 uvCoords;
static vec4 fragColor, gl_FragCoord;
static sampler2D inDirectImage, inIndirectImage;
static Tonemapper tm;
static int debugging_mode;
static vec3 toneLocalExposure(vec3, float) { abort(); }    // see the header

#include "../_ref/gen/random_pcg3d.hpp"
#define TONEMAP_UNCHARTED                  // post.frag:28 defines it before its #include "tonemapping.glsl" (include lines are not transliterated)
#include "../_ref/gen/tonemapping.hpp"
#include "../_ref/gen/post_frag.hpp"
}  // namespace refdisplay

using namespace refdisplay;

// RenderOutput::run over a width x height image (allocation pitch = width); out = RGBA32F per pixel
extern "C" __attribute__((visibility("default")))
void ref_display_run(const Tonemapper* t, int mode, int width, int height, const float* direct, const float* indirect,
                     const float* lastMipDirect, const float* lastMipIndirect, float* out) {
  tm = *t; debugging_mode = mode;
  inDirectImage = sampler2D{(const vec4*)direct, width, height, width, vec4(lastMipDirect[0], lastMipDirect[1], lastMipDirect[2], lastMipDirect[3])};
  inIndirectImage = sampler2D{(const vec4*)indirect, width, height, width, vec4(lastMipIndirect[0], lastMipIndirect[1], lastMipIndirect[2], lastMipIndirect[3])};
  for (int y = 0; y < height; ++y)
    for (int x = 0; x < width; ++x) {
      uvCoords = vec2(((float)x + 0.5f) / (float)width, ((float)y + 0.5f) / (float)height);
      gl_FragCoord = vec4((float)x + 0.5f, (float)y + 0.5f, 0.0f, 1.0f);
      fragColor = vec4();
      refdisplay::main();
      float* o = out + 4 * ((size_t)y * width + x);
      o[0] = fragColor.x; o[1] = fragColor.y; o[2] = fragColor.z; o[3] = fragColor.w;
    }
}

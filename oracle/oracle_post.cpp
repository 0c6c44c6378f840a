/*
 * oracle/oracle_post.cpp — TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * Restatement of the reference's display pass shaders/post.frag (tonemapper + dither; RenderOutput::run,
 * src/render_output.cpp:224-240) with shaders/tonemapping.glsl and pcg3d of shaders/random.glsl:81-92, evaluated once per
 * rendered pixel: uvCoords = (pixel + 0.5) / size, tm.zoom = 1, tm.renderingRatio = (1, 1) — the 1:1 presentation.  The
 * reference's sampler is NEAREST / REPEAT (zero-initialised VkSamplerCreateInfo, render_output.cpp:123-128), so
 * texture(img, uvCoords) is texel (x, y).  autoExposure bit 0 (the GUI's check box, sample_gui.cpp:238-259): the average colour is the
 * 1x1 level of the mip chain RenderOutput::genMipmap blits from the result images (mip_chain_average below); bit 1 (toneLocalExposure —
 * never set by the reference's GUI, reads an uninitialised variable in the default view) is outside the contract.
 * Numerics: DESIGN.md §3 (fp32, one rounding per operation, pow from eid_detmath.h).  Parity: pinned to post.frag itself — main() and
 * every function it calls, transliterated and compiled as C++ (oracle/ref_shim/ref_display.cpp) — bit for bit in every view and with auto
 * exposure (tests/golden/ref_display.npz); what stays contract is the driver's part: the blit chain behind textureLod(img, vec2(0.5), 20).
 */
#include <vector>
#include "oracle.h"

namespace orc {

static inline vec3 vpow(vec3 c, float e) { return vec3(eid_powf(c.x, e), eid_powf(c.y, e), eid_powf(c.z, e)); }
static inline vec3 vfloor(vec3 c) { return vec3(eid_floorf(c.x), eid_floorf(c.y), eid_floorf(c.z)); }
static inline vec3 vclamp01(vec3 c) { return vec3(gclamp(c.x, 0.0f, 1.0f), gclamp(c.y, 0.0f, 1.0f), gclamp(c.z, 0.0f, 1.0f)); }

// tonemapping.glsl:20-32
static const float GAMMA = 2.2f;
static const float INV_GAMMA = 1.0f / GAMMA;
static vec3 linearTosRGB(vec3 color) { return vpow(color, INV_GAMMA); }
static vec3 sRGBToLinear(vec3 srgbIn) { return vpow(srgbIn, GAMMA); }

// tonemapping.glsl:39-58
static vec3 toneMapUncharted2Impl(vec3 color) {
  const float A = 0.15f, B = 0.50f, C = 0.10f, D = 0.20f, E = 0.02f, F = 0.30f;
  return ((color * (A * color + C * B) + D * E) / (color * (A * color + B) + D * F)) - vec3(E / F);
}
static vec3 toneMapUncharted(vec3 color) {
  const float W = 11.2f;
  color = toneMapUncharted2Impl(color * 2.0f);
  vec3 whiteScale = vec3(1.0f) / toneMapUncharted2Impl(vec3(W));
  return linearTosRGB(color * whiteScale);
}
// tonemapping.glsl:78-95 with TONEMAP_UNCHARTED (post.frag:30)
static vec3 toneMap(vec3 color, float u_Exposure) {
  color *= u_Exposure;
  return toneMapUncharted(color);
}

vec3 post_toneMap(vec3 color, float exposure) { return toneMap(color, exposure); }

// post.frag:50-57
static vec3 dither(vec3 linear_color, vec3 noise, float quant) {
  vec3 c0 = vfloor(linearTosRGB(linear_color) / quant) * quant;
  vec3 c1 = c0 + quant;
  vec3 discr = mix(sRGBToLinear(c0), sRGBToLinear(c1), noise);
  return vec3(discr.x < linear_color.x ? c1.x : c0.x, discr.y < linear_color.y ? c1.y : c0.y, discr.z < linear_color.z ? c1.z : c0.z);
}

// random.glsl:81-92
static void pcg3d(uint& x, uint& y, uint& z) {
  x = x * 1664525u + 1013904223u; y = y * 1664525u + 1013904223u; z = z * 1664525u + 1013904223u;
  x += y * z; y += z * x; z += x * y;
  x ^= x >> 16u; y ^= y >> 16u; z ^= z >> 16u;
  x += y * z; y += z * x; z += x * y;
}

// RenderOutput::genMipmap (render_output.cpp:243-253) -> nvvk::cmdGenerateMipmaps (nvpro_core, un-vendored): level i is blitted from level
// i-1 with VK_FILTER_LINEAR, extent max(1, previous / 2) per axis, floor(log2(max(w, h))) + 1 levels, i.e. down to 1 x 1.  A blit with
// a linear filter (Vulkan spec, "Image Blits"): destination texel (i, j) samples the source at u = (i + 0.5) * (srcW / dstW), v likewise,
// bilinear between the texels around (u - 0.5, v - 0.5), clamped to the edge.  Contract: fp32, full-float weights,
// mix(mix(t00, t10, a), mix(t01, t11, a), b).  Returns the single texel of the last level = what textureLod(img, vec2(0.5), 20) reads.
vec4 mip_chain_average(const vec4* img, int width, int height, int pitch) {
  std::vector<vec4> src((size_t)width * height), dst;
  for (int y = 0; y < height; ++y)
    for (int x = 0; x < width; ++x) src[(size_t)y * width + x] = img[(size_t)y * pitch + x];
  int sw = width, sh = height;
  auto mix4 = [](vec4 a, vec4 b, float t) { return vec4(mix(a.x, b.x, t), mix(a.y, b.y, t), mix(a.z, b.z, t), mix(a.w, b.w, t)); };
  while (sw > 1 || sh > 1) {
    const int dw = sw > 1 ? sw / 2 : 1, dh = sh > 1 ? sh / 2 : 1;
    const float scaleU = float(sw) / float(dw), scaleV = float(sh) / float(dh);
    dst.assign((size_t)dw * dh, vec4());
    for (int j = 0; j < dh; ++j)
      for (int i = 0; i < dw; ++i) {
        const float a = (float(i) + 0.5f) * scaleU - 0.5f, b = (float(j) + 0.5f) * scaleV - 0.5f;
        const float af = eid_floorf(a), bf = eid_floorf(b);
        const float fa = a - af, fb = b - bf;
        const int x0 = imax(0, imin(sw - 1, f2i(af))), x1 = imax(0, imin(sw - 1, f2i(af) + 1));
        const int y0 = imax(0, imin(sh - 1, f2i(bf))), y1 = imax(0, imin(sh - 1, f2i(bf) + 1));
        dst[(size_t)j * dw + i] = mix4(mix4(src[(size_t)y0 * sw + x0], src[(size_t)y0 * sw + x1], fa),
                                       mix4(src[(size_t)y1 * sw + x0], src[(size_t)y1 * sw + x1], fa), fb);
      }
    src.swap(dst); sw = dw; sh = dh;
  }
  return src[0];
}

// post.frag:60-62, 65-70
static float luminancePost(vec3 color) { return dot(color, vec3(0.2126f, 0.7152f, 0.0722f)); }
static vec3 toneExposure(const Tonemapper& tm, vec3 RGB, float logAvgLum) {
  // RGB2XYZ = mat3(0.4124564, 0.3575761, 0.1804375, 0.2126729, ...) is filled column by column, so XYZ.y = dot of the SECOND ROW of
  // that storage = (0.3575761, 0.7151522, 0.1191920) with RGB: kept as written
  const float XYZy = (0.3575761f * RGB.x + 0.7151522f * RGB.y) + 0.1191920f * RGB.z;
  const float Y = (tm.key / logAvgLum) * XYZy;
  const float Yd = (Y * (1.0f + Y / (tm.Ywhite * tm.Ywhite))) / (1.0f + Y);
  return RGB / XYZy * Yd;
}
vec3 post_toneExposure(const Tonemapper& tm, vec3 RGB, float logAvgLum) { return toneExposure(tm, RGB, logAvgLum); }

// post.frag main :107-178 for one pixel; direct / indirect = RGBA of texel (px, py); avgDirect / avgIndirect = the 1x1 mip level of the
// two result images (only read when tm.autoExposure bit 0 is set)
vec4 post_frag(const Tonemapper& tm, int debugging_mode, vec4 direct, vec4 indirect, int px, int py, int width, int height, vec4 avgDirect, vec4 avgIndirect) {
  const vec2 uvCoords((float(px) + 0.5f) / float(width), (float(py) + 0.5f) / float(height));
  if (debugging_mode == eDepth) {
    float depth = direct.w;
    depth *= eid_powf(2.0f, tm.brightness);
    depth += tm.saturation;
    depth = gclamp(eid_powf(depth, 1.0f / tm.contrast), 0.f, 1.f);
    return vec4(depth, depth, depth, 1.0f);
  } else if (debugging_mode > eIndirectStage) {
    vec3 color = direct.xyz();
    if (debugging_mode == eBaseColor) color = vclamp01(vpow(color, 0.45454545454545f));
    return vec4(color, 1.0f);
  }
  vec4 hdr;
  if (debugging_mode == eDirectStage) hdr = direct;
  else if (debugging_mode == eIndirectStage) hdr = indirect;
  else hdr = vec4(direct.x + indirect.x, direct.y + indirect.y, direct.z + indirect.z, direct.w + indirect.w);
  hdr.w = 1.0f;
  if (((tm.autoExposure >> 0) & 1) == 1) {              // :133-152
    vec4 avg;
    if (debugging_mode == eDirectStage) avg = avgDirect;
    else if (debugging_mode == eIndirectStage) avg = avgIndirect;
    else avg = vec4(avgDirect.x + avgIndirect.x, avgDirect.y + avgIndirect.y, avgDirect.z + avgIndirect.z, avgDirect.w + avgIndirect.w);
    const float avgLum2 = luminancePost(avg.xyz());
    const vec3 e = toneExposure(tm, hdr.xyz(), avgLum2);
    hdr = vec4(e, 1.0f);
  }

  vec3 color = toneMap(hdr.xyz(), tm.avgLum);          // tonemap + linear to sRGB

  uint rx = (uint)px, ry = (uint)py, rz = 0u;           // uvec3(gl_FragCoord.xy, 0): the pixel centre truncates to the pixel
  pcg3d(rx, ry, rz);
  vec3 noise = vec3(uintBitsToFloat(0x3f800000u | (rx >> 9)), uintBitsToFloat(0x3f800000u | (ry >> 9)), uintBitsToFloat(0x3f800000u | (rz >> 9))) + (-1.0f);
  color = dither(sRGBToLinear(color), noise, 1.f / 255.f);

  color = vclamp01(mix(vec3(0.5f), color, tm.contrast));                 // contrast
  color = vpow(color, 1.0f / tm.brightness);                            // brightness
  vec3 i = vec3(dot(color, vec3(0.299f, 0.587f, 0.114f)));              // saturation
  color = mix(i, color, tm.saturation);
  vec2 uv = ((uvCoords * vec2(tm.renderingRatio.x, tm.renderingRatio.y)) - 0.5f) * 2.0f;   // vignette
  color *= 1.0f - dot(uv, uv) * tm.vignette;
  return vec4(color, 1.0f);
}

}  // namespace orc

// pack.h — bit-level packing shared by the host table builders and the kernels:
// octahedral unit-vector codec (reference shaders/compress.glsl:111-180), unorm4x8 pack/unpack,
// the 8-bit material hash (common.glsl:141-143) and Rec.709 luminance (tools.hpp:57-61).
#pragma once
#include <math.h>
#include <stdint.h>
#include "eid_detmath.h"

namespace eid {

struct f3 { float x, y, z; };

EID_HD float lum709(float r, float g, float b) { return EID_ADD(EID_ADD(EID_MUL(0.2126f, r), EID_MUL(0.7152f, g)), EID_MUL(0.0722f, b)); }
EID_HD uint32_t hash8(uint32_t a) { return (a ^ (a >> 8)) << 24; }

// float -> int with a defined result for NaN / out of range
EID_HD int f2i_sat(float f) {
  if (f != f) return 0;
  if (f >= 2147483520.0f) return 2147483520;
  if (f <= -2147483648.0f) return (int)0x80000000;
  return (int)f;
}
EID_HD uint32_t f2u_sat(float f) {
  if (f != f || f <= 0.0f) return 0u;
  if (f >= 4294967040.0f) return 4294967040u;
  return (uint32_t)f;
}

// round-half-to-even (GLSL roundEven; identical to the reference's C++ twin for every input)
EID_HD float roundEvenF(float x) { return rintf(x); }

EID_HD uint32_t octEncode(float nx, float ny, float nz) {
  const float BIG = 3.402823466e+38f;
  if (!(nx < BIG) || isinf(nx)) return ~0u;
  const float d = EID_DIV(32767.0f, EID_ADD(EID_ADD(fabsf(nx), fabsf(ny)), fabsf(nz)));
  int x = f2i_sat(roundEvenF(EID_MUL(nx, d)));
  int y = f2i_sat(roundEvenF(EID_MUL(ny, d)));
  if (nz < 0.0f) {
    const int mx = x >> 31, my = y >> 31;
    const int t = 32767 + mx + my;
    const int ox = x;
    x = (t - (y ^ my)) ^ mx;
    y = (t - (ox ^ mx)) ^ my;
  }
  uint32_t packed = ((uint32_t)(y + 32767) << 16) | (uint32_t)(x + 32767);
  return (packed == ~0u) ? ~0x1u : packed;
}

EID_HD float snorm15ToFloat(int v) {   // short_to_floatm11
  return (v >= 0) ? EID_SUB(eid_u2f(0x3F800000u | ((uint32_t)v << 8)), 1.0f)
                  : EID_ADD(eid_u2f(0xBF800000u | ((uint32_t)(-v) << 8)), 1.0f);
}

EID_HD f3 octDecode(uint32_t packed) {
  if (packed == ~0u) { f3 r = {3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f}; return r; }
  int x = (int)(packed & 0xFFFFu) - 32767;
  int y = (int)(packed >> 16) - 32767;
  const int mx = x >> 31, my = y >> 31;
  const int t0 = 32767 + mx + my;
  const int ym = y ^ my;
  const int t1 = t0 - (x ^ mx);
  const int z = t1 - ym;
  float zf;
  if (z < 0) {
    x = (t0 - ym) ^ mx;
    y = t1 ^ my;
    zf = EID_ADD(eid_u2f(0xBF800000u | ((uint32_t)(-z) << 8)), 1.0f);
  } else {
    zf = EID_SUB(eid_u2f(0x3F800000u | ((uint32_t)z << 8)), 1.0f);
  }
  float fx = snorm15ToFloat(x), fy = snorm15ToFloat(y);
  float inv = EID_DIV(1.0f, eid_sqrtf(EID_ADD(EID_ADD(EID_MUL(fx, fx), EID_MUL(fy, fy)), EID_MUL(zf, zf))));
  f3 r = {EID_MUL(fx, inv), EID_MUL(fy, inv), EID_MUL(zf, inv)};
  return r;
}

EID_HD uint32_t unormByte(float c) {
  float v = (c < 0.0f) ? 0.0f : c;      // max(c, 0): NaN stays NaN ...
  v = (1.0f < v) ? 1.0f : v;            // min(v, 1)
  return f2u_sat(roundf(EID_MUL(v, 255.0f))) & 0xffu;   // ... and NaN packs to 0
}
EID_HD uint32_t packUnorm4(float a, float b, float c, float d) {
  return unormByte(a) | (unormByte(b) << 8) | (unormByte(c) << 16) | (unormByte(d) << 24);
}
EID_HD float unormToFloat(uint32_t byte) { return EID_DIV((float)byte, 255.0f); }

}  // namespace eid

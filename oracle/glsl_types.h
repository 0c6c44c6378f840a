/*
 * oracle/glsl_types.h — TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * Minimal GLSL-like vector/matrix types so the restatement of the reference shaders in
 * oracle_shaders.cpp can follow the GLSL text line by line.  Every operator is a plain IEEE-754
 * single-precision operation evaluated in the written order; the oracle is compiled with
 * -ffp-contract=off so no FMA is ever formed.  The definitions of the GLSL built-ins whose
 * rounding GLSL leaves open (normalize, dot, mix, reflect, inverse, pack/unpack) are the
 * "numerical contract" of DESIGN.md §3 and are mirrored by the CUDA kernels.
 */
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include "eid_detmath.h"

namespace orc {

typedef unsigned int uint;

struct vec2;
struct ivec2 { int x, y; ivec2() : x(0), y(0) {} ivec2(int a, int b) : x(a), y(b) {} explicit ivec2(const vec2& v); ivec2 xy() const { return *this; } };
struct vec2 {
  float x, y;
  vec2() : x(0), y(0) {}
  vec2(float a) : x(a), y(a) {}
  vec2(float a, float b) : x(a), y(b) {}
  explicit vec2(ivec2 v) : x((float)v.x), y((float)v.y) {}   // GLSL vec2(ivec2)
  vec2 xy() const { return *this; }
};
struct vec4;
struct vec3 {
  float x, y, z;
  vec3() : x(0), y(0), z(0) {}
  explicit vec3(const vec4& v);        // GLSL vec3(vec4): xyz
  vec3(float a) : x(a), y(a), z(a) {}
  vec3(float a, float b, float c) : x(a), y(b), z(c) {}
  vec3(vec2 v, float c) : x(v.x), y(v.y), z(c) {}
  float& operator[](int i) { return (&x)[i]; }
  float operator[](int i) const { return (&x)[i]; }
  vec2 xy() const { return vec2(x, y); }
  vec3 xyz() const { return *this; }
  vec3& xyz() { return *this; }          // lvalue swizzle `v.xyz = ...`
  vec3 rgb() const { return *this; }
};
struct vec4 {
  float x, y, z, w;
  vec4() : x(0), y(0), z(0), w(0) {}
  vec4(float a) : x(a), y(a), z(a), w(a) {}
  vec4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
  vec4(vec3 v, float d) : x(v.x), y(v.y), z(v.z), w(d) {}
  vec3 xyz() const { return vec3(x, y, z); }
  vec3 rgb() const { return vec3(x, y, z); }
  vec2 xy() const { return vec2(x, y); }
};
inline vec3::vec3(const vec4& v) : x(v.x), y(v.y), z(v.z) {}
inline vec4 operator+(vec4 a, vec4 b) { return vec4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
inline vec4 operator*(vec4 a, float s) { return vec4(a.x * s, a.y * s, a.z * s, a.w * s); }
inline vec4& operator*=(vec4& a, vec4 b) { a = vec4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); return a; }
struct uvec2 { uint x, y; uvec2() : x(0), y(0) {} uvec2(uint a, uint b) : x(a), y(b) {} };
struct uvec4 { uint x, y, z, w; uvec4() : x(0), y(0), z(0), w(0) {} uvec4(uint a, uint b, uint c, uint d) : x(a), y(b), z(c), w(d) {} uvec2 xy() const { return uvec2(x, y); } };

inline vec2 operator+(vec2 a, vec2 b) { return vec2(a.x + b.x, a.y + b.y); }
inline vec2 operator-(vec2 a, vec2 b) { return vec2(a.x - b.x, a.y - b.y); }
inline vec2 operator*(vec2 a, vec2 b) { return vec2(a.x * b.x, a.y * b.y); }
inline vec2 operator*(vec2 a, float s) { return vec2(a.x * s, a.y * s); }
inline vec2 operator*(float s, vec2 a) { return vec2(s * a.x, s * a.y); }
inline vec2 operator/(vec2 a, vec2 b) { return vec2(a.x / b.x, a.y / b.y); }
inline vec2 operator+(vec2 a, float s) { return vec2(a.x + s, a.y + s); }
inline vec2 operator-(vec2 a, float s) { return vec2(a.x - s, a.y - s); }

inline vec3 operator+(vec3 a, vec3 b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator-(vec3 a, vec3 b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator*(vec3 a, vec3 b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline vec3 operator/(vec3 a, vec3 b) { return vec3(a.x / b.x, a.y / b.y, a.z / b.z); }
inline vec3 operator*(vec3 a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
inline vec3 operator*(float s, vec3 a) { return vec3(s * a.x, s * a.y, s * a.z); }
inline vec3 operator/(vec3 a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }
inline vec3 operator+(vec3 a, float s) { return vec3(a.x + s, a.y + s, a.z + s); }
inline vec3 operator-(float s, vec3 a) { return vec3(s - a.x, s - a.y, s - a.z); }
inline vec3 operator-(vec3 a) { return vec3(-a.x, -a.y, -a.z); }
inline vec3& operator+=(vec3& a, vec3 b) { a = a + b; return a; }
inline vec3& operator*=(vec3& a, vec3 b) { a = a * b; return a; }
inline vec3& operator*=(vec3& a, float s) { a = a * s; return a; }
inline vec3& operator/=(vec3& a, float s) { a = a / s; return a; }
inline ivec2 operator+(ivec2 a, ivec2 b) { return ivec2(a.x + b.x, a.y + b.y); }
inline ivec2 operator*(ivec2 a, int s) { return ivec2(a.x * s, a.y * s); }
inline ivec2 operator/(ivec2 a, int s) { return ivec2(a.x / s, a.y / s); }   // C truncation == GLSL

// --- built-ins with the rounding order fixed by the contract --------------------------------
inline float dot(vec2 a, vec2 b) { return a.x * b.x + a.y * b.y; }
inline float dot(vec3 a, vec3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline vec3 cross(vec3 a, vec3 b) { return vec3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline float length(vec3 a) { return sqrtf(dot(a, a)); }
inline vec3 normalize(vec3 a) { float inv = 1.0f / sqrtf(dot(a, a)); return a * inv; }
inline float gmin(float a, float b) { return (b < a) ? b : a; }   // GLSL min
inline float gmax(float a, float b) { return (a < b) ? b : a; }   // GLSL max
inline int imin(int a, int b) { return (b < a) ? b : a; }
inline int imax(int a, int b) { return (a < b) ? b : a; }
inline float gclamp(float x, float lo, float hi) { return gmin(gmax(x, lo), hi); }
inline float gabs(float a) { return fabsf(a); }
inline float mix(float x, float y, float a) { return x * (1.0f - a) + y * a; }
inline vec3 mix(vec3 x, vec3 y, float a) { return x * (1.0f - a) + y * a; }
inline vec3 mix(vec3 x, vec3 y, vec3 a) { return x * (1.0f - a) + y * a; }
inline vec3 reflect(vec3 I, vec3 N) { return I - N * (2.0f * dot(N, I)); }
inline bool gisnan(float a) { return a != a; }
inline bool gisinf(float a) { return std::isinf(a); }
inline float uintBitsToFloat(uint u) { return eid_u2f(u); }
inline uint floatBitsToUint(float f) { return eid_f2u(f); }
inline int floatBitsToInt(float f) { return (int)eid_f2u(f); }
inline float intBitsToFloat(int i) { return eid_u2f((uint)i); }
// float -> int with a defined result for NaN / out-of-range (GLSL leaves it undefined)
inline int f2i(float f) {
  if (f != f) return 0;
  if (f >= 2147483520.0f) return 2147483520;
  if (f <= -2147483648.0f) return (int)0x80000000;
  return (int)f;
}
inline ivec2::ivec2(const vec2& v) : x(f2i(v.x)), y(f2i(v.y)) {}   // GLSL ivec2(vec2): truncation (saturating / NaN -> 0 by the contract)
inline uint f2u(float f) {
  if (f != f || f <= 0.0f) return 0u;
  if (f >= 4294967040.0f) return 4294967040u;
  return (uint)f;
}

// column-major matrices (GLSL): m[c] is column c
struct mat3 {
  vec3 c[3];
  mat3() {}
  mat3(vec3 a, vec3 b, vec3 d) { c[0] = a; c[1] = b; c[2] = d; }
  mat3(float a0, float a1, float a2, float b0, float b1, float b2, float d0, float d1, float d2) {   // GLSL fills column by column
    c[0] = vec3(a0, a1, a2); c[1] = vec3(b0, b1, b2); c[2] = vec3(d0, d1, d2);
  }
};
inline vec3 operator*(const mat3& m, vec3 v) { return (m.c[0] * v.x + m.c[1] * v.y) + m.c[2] * v.z; }
// cofactor inverse, fixed evaluation order (contract)
inline mat3 inverse(const mat3& m) {
  float a00 = m.c[0].x, a01 = m.c[0].y, a02 = m.c[0].z;
  float a10 = m.c[1].x, a11 = m.c[1].y, a12 = m.c[1].z;
  float a20 = m.c[2].x, a21 = m.c[2].y, a22 = m.c[2].z;
  float b01 = a22 * a11 - a12 * a21;
  float b11 = a12 * a20 - a22 * a10;
  float b21 = a21 * a10 - a11 * a20;
  float det = (a00 * b01 + a01 * b11) + a02 * b21;
  float id = 1.0f / det;
  mat3 r;
  r.c[0] = vec3(b01 * id, (a02 * a21 - a22 * a01) * id, (a12 * a01 - a02 * a11) * id);
  r.c[1] = vec3(b11 * id, (a22 * a00 - a02 * a20) * id, (a02 * a10 - a12 * a00) * id);
  r.c[2] = vec3(b21 * id, (a01 * a20 - a21 * a00) * id, (a11 * a00 - a01 * a10) * id);
  return r;
}
// mat4x3: 4 columns of vec3 (objectToWorld / worldToObject of the ray query)
struct mat4x3 { vec3 c[4]; };
// M * vec4(v, 1)
inline vec3 mulPoint(const mat4x3& m, vec3 v) { return ((m.c[0] * v.x + m.c[1] * v.y) + m.c[2] * v.z) + m.c[3]; }
// mat4(M) * vec4(v, 0)  (xyz)
inline vec3 mulVector(const mat4x3& m, vec3 v) { return (m.c[0] * v.x + m.c[1] * v.y) + m.c[2] * v.z; }
// vec3(v * M): row-vector times matrix = dot with each column (first three)
inline vec3 mulTransposed(vec3 v, const mat4x3& m) { return vec3(dot(v, m.c[0]), dot(v, m.c[1]), dot(v, m.c[2])); }

struct mat4 { float m[16]; };   // column-major, m[c*4+r]
// M * vec4(x,y,z,w) with the contract's summation order ((c0*x + c1*y) + c2*z) + c3*w
inline vec4 mul(const mat4& M, vec4 v) {
  vec4 r;
  float* o = &r.x;
  for (int i = 0; i < 4; ++i) o[i] = ((M.m[i] * v.x + M.m[4 + i] * v.y) + M.m[8 + i] * v.z) + M.m[12 + i] * v.w;
  return r;
}
// xyz of M * vec4(v, 0): the w term is dropped (contract)
inline vec3 mulDir(const mat4& M, vec3 v) {
  vec3 r;
  for (int i = 0; i < 3; ++i) r[i] = (M.m[i] * v.x + M.m[4 + i] * v.y) + M.m[8 + i] * v.z;
  return r;
}

// pack/unpack (GLSL packUnorm4x8: round(clamp(c,0,1)*255), ties away from zero like the C++ twin
// in the reference's compress.glsl:60-74)
inline uint packUnorm4x8(vec4 v) {
  uint r = 0;
  const float* p = &v.x;
  for (int i = 0; i < 4; ++i) {
    float c = gmin(gmax(p[i], 0.0f), 1.0f) * 255.0f;
    uint b = f2u(roundf(c));
    r |= (b & 0xffu) << (8 * i);
  }
  return r;
}
inline vec4 unpackUnorm4x8(uint p) {
  return vec4(float(p & 0xffu) / 255.0f, float((p >> 8) & 0xffu) / 255.0f, float((p >> 16) & 0xffu) / 255.0f,
              float((p >> 24) & 0xffu) / 255.0f);
}

}  // namespace orc

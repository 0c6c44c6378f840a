/*
 * oracle/ref_shim/nvmath/nvmath.h — TEST INFRASTRUCTURE.
 * Minimal stand-in for nvpro_core's nvmath (un-vendored) — just enough for the
 * reference's shaders/host_device.h, shaders/compress.glsl (C++ branch) and src/alias_table.hpp to
 * compile unmodified from /root/reference into oracle/_ref/libref.so.  No reference code is copied.
 */
#pragma once
#include <math.h>
#include <stdlib.h>
#include <cmath>
#include <cstdlib>
#include <algorithm>
#include <vector>
using std::abs;
using std::isinf;
namespace nvmath {
template <class T> struct vector2 { T x, y; };
template <class T> struct vector3 { T x, y, z; vector3() = default; vector3(T a, T b, T c) : x(a), y(b), z(c) {} explicit vector3(T a) : x(a), y(a), z(a) {} };
template <class T> struct vector4 { T x, y, z, w; };
template <class T> struct matrix4 { T m[16]; };
typedef vector2<int> vec2i; typedef vector2<float> vec2f; typedef vector2<unsigned int> vec2ui;
typedef vector3<float> vec3f; typedef vector4<float> vec4f; typedef vector4<unsigned int> vec4ui;
typedef matrix4<float> mat4f;
// nvmath::normalize: norm = sqrt(x^2+y^2+z^2); scale by 1/norm (0 when norm <= eps)
inline vec3f normalize(const vec3f& u) {
  float norm = sqrtf(u.x * u.x + u.y * u.y + u.z * u.z);
  norm = (norm > 1e-6f) ? 1.0f / norm : 0.0f;
  return vec3f(u.x * norm, u.y * norm, u.z * norm);
}
}  // namespace nvmath
using nvmath::normalize;

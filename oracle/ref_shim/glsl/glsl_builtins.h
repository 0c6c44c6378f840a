/* oracle/ref_shim/glsl/glsl_builtins.h — TEST INFRASTRUCTURE.  The GLSL built-ins the transliterated reference shaders call
 * (glsl_prep.py renames `pow(` -> `GLSL_pow(` ...), bound to the numerical contract of DESIGN.md §3: IEEE fp32 +,-,*,/,sqrt; the
 * deterministic transcendental functions of include/eid_detmath.h; dot / normalize / mix / reflect / inverse with the evaluation order
 * of oracle/glsl_types.h.  The EXPRESSIONS evaluated with them are the reference's own text. */
#pragma once
#include "../../glsl_types.h"
#include "nvmath/nvmath.h"
using mat3 = orc::mat3;
struct ivec3 { int x, y, z; ivec3(int a, int b, int c) : x(a), y(b), z(c) {} };
using mat4x3 = orc::mat4x3;
inline float GLSL_abs(float a) { return fabsf(a); }
inline float GLSL_max(float a, float b) { return orc::gmax(a, b); }
inline float GLSL_min(float a, float b) { return orc::gmin(a, b); }
inline float GLSL_clamp(float x, float lo, float hi) { return orc::gclamp(x, lo, hi); }
inline float GLSL_mix(float x, float y, float a) { return orc::mix(x, y, a); }
inline float GLSL_sqrt(float a) { return sqrtf(a); }
inline float GLSL_pow(float a, float b) { return eid_powf(a, b); }
inline float GLSL_exp(float a) { return eid_expf(a); }
inline float GLSL_sin(float a) { return eid_sinf(a); }
inline float GLSL_cos(float a) { return eid_cosf(a); }
inline float GLSL_tan(float a) { float s, c; eid_sincosf(a, &s, &c); return s / c; }
inline float GLSL_acos(float a) { return eid_acosf(a); }
inline float GLSL_asin(float a) { return eid_asinf(a); }
inline float GLSL_atan(float y, float x) { return eid_atan2f(y, x); }
inline float GLSL_atan(float x) { return eid_atanf(x); }
inline float GLSL_floor(float a) { return eid_floorf(a); }
inline float GLSL_smoothstep(float e0, float e1, float x) { float t = orc::gclamp((x - e0) / (e1 - e0), 0.0f, 1.0f); return t * t * (3.0f - 2.0f * t); }
inline bool GLSL_isnan(float a) { return a != a; }
inline bool GLSL_isinf(float a) { return std::isinf(a); }
inline float GLSL_intBitsToFloat(int i) { return orc::intBitsToFloat(i); }
inline int GLSL_floatBitsToInt(float f) { return orc::floatBitsToInt(f); }
inline float GLSL_uintBitsToFloat(unsigned int u) { return orc::uintBitsToFloat(u); }
inline unsigned int GLSL_floatBitsToUint(float f) { return orc::floatBitsToUint(f); }
inline float GLSL_dot(orc::vec2 a, orc::vec2 b) { return orc::dot(a, b); }
inline float GLSL_dot(orc::vec3 a, orc::vec3 b) { return orc::dot(a, b); }
inline orc::vec3 GLSL_cross(orc::vec3 a, orc::vec3 b) { return orc::cross(a, b); }
inline orc::vec3 GLSL_normalize(orc::vec3 a) { return orc::normalize(a); }
inline float GLSL_length(orc::vec3 a) { return orc::length(a); }
inline orc::vec3 GLSL_reflect(orc::vec3 i, orc::vec3 n) { return orc::reflect(i, n); }
inline orc::mat3 GLSL_inverse(const orc::mat3& m) { return orc::inverse(m); }
inline orc::vec3 GLSL_mix(orc::vec3 x, orc::vec3 y, float a) { return orc::mix(x, y, a); }
inline orc::vec3 GLSL_mix(orc::vec3 x, orc::vec3 y, orc::vec3 a) { return orc::mix(x, y, a); }
inline orc::vec3 GLSL_pow(orc::vec3 a, orc::vec3 b) { return orc::vec3(eid_powf(a.x, b.x), eid_powf(a.y, b.y), eid_powf(a.z, b.z)); }
inline orc::vec3 GLSL_exp(orc::vec3 a) { return orc::vec3(eid_expf(a.x), eid_expf(a.y), eid_expf(a.z)); }
inline orc::vec3 GLSL_max(orc::vec3 a, orc::vec3 b) { return orc::vec3(orc::gmax(a.x, b.x), orc::gmax(a.y, b.y), orc::gmax(a.z, b.z)); }
inline orc::vec3 GLSL_clamp(orc::vec3 a, float lo, float hi) { return orc::vec3(orc::gclamp(a.x, lo, hi), orc::gclamp(a.y, lo, hi), orc::gclamp(a.z, lo, hi)); }
inline int GLSL_min(int a, int b) { return b < a ? b : a; }
inline unsigned int GLSL_min(unsigned int a, unsigned int b) { return b < a ? b : a; }
inline orc::vec4 operator*(const nvmath::mat4f& M, orc::vec4 v) { return orc::mul(*reinterpret_cast<const orc::mat4*>(&M), v); }   // ((c0 x + c1 y) + c2 z) + c3 w
inline orc::vec4 GLSL_unpackUnorm4x8(unsigned int p) { return orc::unpackUnorm4x8(p); }
inline int GLSL_max(int a, int b) { return a < b ? b : a; }
inline bool GLSL_isnan(orc::vec3 v) { return v.x != v.x || v.y != v.y || v.z != v.z; }
// matrix products of shade_state.glsl: mat4x3 * vec4 -> vec3, vec3 * mat4x3 -> vec4 (one dot per column), mat4(mat4x3) * vec4
inline orc::vec3 operator*(const orc::mat4x3& m, orc::vec4 v) { return ((m.c[0] * v.x + m.c[1] * v.y) + m.c[2] * v.z) + m.c[3] * v.w; }
inline orc::vec4 operator*(orc::vec3 v, const orc::mat4x3& m) { return orc::vec4(orc::dot(v, m.c[0]), orc::dot(v, m.c[1]), orc::dot(v, m.c[2]), orc::dot(v, m.c[3])); }
inline float GLSL_max(float a, int b) { return orc::gmax(a, (float)b); }
inline float GLSL_max(int a, float b) { return orc::gmax((float)a, b); }
inline float GLSL_min(float a, int b) { return orc::gmin(a, (float)b); }

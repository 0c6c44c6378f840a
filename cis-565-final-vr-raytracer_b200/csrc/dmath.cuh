// dmath.cuh — device vector helpers with the rounding order fixed by the numerical contract
// (DESIGN.md §3): every GLSL built-in the reference leaves open (dot, cross, normalize, mix, reflect,
// mat*vec) is spelled out as individually rounded fp32 operations so the CUDA kernels and the CPU oracle
// agree bit for bit.  The translation unit is compiled with --fmad=false; the explicit __f*_rn
// intrinsics below keep the contract even if that flag is ever dropped.  Traversal box tests (which do
// not influence results, only which triangles get tested) use explicit fmaf for speed instead.
#pragma once
#include <cuda_runtime.h>
#include "pack.h"

namespace eid {

#define DEV __device__ __forceinline__

DEV f3 mk3(float x, float y, float z) { f3 r = {x, y, z}; return r; }
DEV f3 mk3(float s) { f3 r = {s, s, s}; return r; }
DEV f3 ld3(const eid_vec3& v) { f3 r = {v.x, v.y, v.z}; return r; }
DEV eid_vec3 st3(f3 v) { eid_vec3 r = {v.x, v.y, v.z}; return r; }

DEV f3 operator+(f3 a, f3 b) { return mk3(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z)); }
DEV f3 operator-(f3 a, f3 b) { return mk3(__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y), __fsub_rn(a.z, b.z)); }
DEV f3 operator*(f3 a, f3 b) { return mk3(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y), __fmul_rn(a.z, b.z)); }
DEV f3 operator/(f3 a, f3 b) { return mk3(__fdiv_rn(a.x, b.x), __fdiv_rn(a.y, b.y), __fdiv_rn(a.z, b.z)); }
DEV f3 operator*(f3 a, float s) { return mk3(__fmul_rn(a.x, s), __fmul_rn(a.y, s), __fmul_rn(a.z, s)); }
DEV f3 operator*(float s, f3 a) { return mk3(__fmul_rn(s, a.x), __fmul_rn(s, a.y), __fmul_rn(s, a.z)); }
DEV f3 operator/(f3 a, float s) { return mk3(__fdiv_rn(a.x, s), __fdiv_rn(a.y, s), __fdiv_rn(a.z, s)); }
DEV f3 operator+(f3 a, float s) { return mk3(__fadd_rn(a.x, s), __fadd_rn(a.y, s), __fadd_rn(a.z, s)); }
DEV f3 operator-(float s, f3 a) { return mk3(__fsub_rn(s, a.x), __fsub_rn(s, a.y), __fsub_rn(s, a.z)); }
DEV f3 operator-(f3 a) { return mk3(-a.x, -a.y, -a.z); }

DEV float dot3(f3 a, f3 b) { return __fadd_rn(__fadd_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)), __fmul_rn(a.z, b.z)); }
DEV f3 cross3(f3 a, f3 b) {
  return mk3(__fsub_rn(__fmul_rn(a.y, b.z), __fmul_rn(a.z, b.y)), __fsub_rn(__fmul_rn(a.z, b.x), __fmul_rn(a.x, b.z)),
             __fsub_rn(__fmul_rn(a.x, b.y), __fmul_rn(a.y, b.x)));
}
DEV float len3(f3 a) { return __fsqrt_rn(dot3(a, a)); }
DEV f3 norm3(f3 a) { float inv = __fdiv_rn(1.0f, __fsqrt_rn(dot3(a, a))); return a * inv; }
DEV float gmin(float a, float b) { return (b < a) ? b : a; }   // GLSL min / max definitions
DEV float gmax(float a, float b) { return (a < b) ? b : a; }
DEV float mixf(float x, float y, float a) { return __fadd_rn(__fmul_rn(x, __fsub_rn(1.0f, a)), __fmul_rn(y, a)); }
DEV f3 mix3(f3 x, f3 y, float a) { return x * __fsub_rn(1.0f, a) + y * a; }
DEV f3 mix3(f3 x, f3 y, f3 a) { return x * (1.0f - a) + y * a; }
DEV f3 reflect3(f3 I, f3 N) { return I - N * __fmul_rn(2.0f, dot3(N, I)); }
DEV float lum3(f3 c) { return lum709(c.x, c.y, c.z); }
DEV bool nan3(f3 v) { return v.x != v.x || v.y != v.y || v.z != v.z; }

// GLSL mat4x3 (4 columns of 3 floats) helpers
DEV f3 col(const float* m, int c) { return mk3(m[3 * c], m[3 * c + 1], m[3 * c + 2]); }
DEV f3 xfPoint(const float* m, f3 v) { return ((col(m, 0) * v.x + col(m, 1) * v.y) + col(m, 2) * v.z) + col(m, 3); }   // M * vec4(v,1)
DEV f3 xfVector(const float* m, f3 v) { return (col(m, 0) * v.x + col(m, 1) * v.y) + col(m, 2) * v.z; }                 // mat4(M) * vec4(v,0)
DEV f3 xfTransposed(f3 v, const float* m) { return mk3(dot3(v, col(m, 0)), dot3(v, col(m, 1)), dot3(v, col(m, 2))); }  // vec3(v * M)

// column-major mat4 * vec4 with the order ((c0*x + c1*y) + c2*z) + c3*w
DEV void mat4MulV(const eid_mat4& M, float x, float y, float z, float w, float o[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
    o[i] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(M.m[i], x), __fmul_rn(M.m[4 + i], y)), __fmul_rn(M.m[8 + i], z)), __fmul_rn(M.m[12 + i], w));
}
DEV f3 mat4MulDir(const eid_mat4& M, f3 v) {   // xyz of M * vec4(v, 0); the w term is dropped
  f3 r;
  r.x = __fadd_rn(__fadd_rn(__fmul_rn(M.m[0], v.x), __fmul_rn(M.m[4], v.y)), __fmul_rn(M.m[8], v.z));
  r.y = __fadd_rn(__fadd_rn(__fmul_rn(M.m[1], v.x), __fmul_rn(M.m[5], v.y)), __fmul_rn(M.m[9], v.z));
  r.z = __fadd_rn(__fadd_rn(__fmul_rn(M.m[2], v.x), __fmul_rn(M.m[6], v.y)), __fmul_rn(M.m[10], v.z));
  return r;
}

}  // namespace eid

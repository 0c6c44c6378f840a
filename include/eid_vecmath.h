/*
 * eid_vecmath.h — the handful of nvmath (nvpro_core, un-vendored) entry
 * points the hot path's host side calls: column-major mat4f, perspectiveVK, look-at, invert.
 *
 * Call sites being served (reference): scene.cpp:783-795 (Scene::updateCamera),
 * scene.cpp:331-332,388-390 (light/world transforms), accelstruct.cpp:152 (toTransformMatrixKHR).
 *
 * Semantics restated from the nvmath documentation/behaviour (from memory, un-vendored):
 *   mat4f is column-major, element aRC = row R / column C lives at m[C*4+R];
 *   perspectiveVK: Vulkan clip space (y flipped, depth 0..1), right-handed view space;
 *   look_at: right-handed, camera looks down -Z.
 * Shared by the product and by the CPU oracle in the same way both would link nvmath.
 */
#ifndef EIDOLA_VECMATH_H
#define EIDOLA_VECMATH_H

#include "host_device.h"
#include "eid_detmath.h"

EID_HD eid_mat4 eid_mat4_identity() {
  eid_mat4 r;
  for (int i = 0; i < 16; ++i) r.m[i] = (i % 5 == 0) ? 1.0f : 0.0f;
  return r;
}

/* r = a * b (column vectors: (a*b)*v == a*(b*v)) */
EID_HD eid_mat4 eid_mat4_mul(const eid_mat4& a, const eid_mat4& b) {
  eid_mat4 r;
  for (int c = 0; c < 4; ++c)
    for (int row = 0; row < 4; ++row) {
      float s = EID_MUL(a.m[0 * 4 + row], b.m[c * 4 + 0]);
      s = EID_ADD(s, EID_MUL(a.m[1 * 4 + row], b.m[c * 4 + 1]));
      s = EID_ADD(s, EID_MUL(a.m[2 * 4 + row], b.m[c * 4 + 2]));
      s = EID_ADD(s, EID_MUL(a.m[3 * 4 + row], b.m[c * 4 + 3]));
      r.m[c * 4 + row] = s;
    }
  return r;
}

EID_HD eid_vec4 eid_mat4_mulv(const eid_mat4& a, eid_vec4 v) {
  eid_vec4 r;
  float* o = &r.x;
  for (int row = 0; row < 4; ++row) {
    float s = EID_MUL(a.m[0 + row], v.x);
    s = EID_ADD(s, EID_MUL(a.m[4 + row], v.y));
    s = EID_ADD(s, EID_MUL(a.m[8 + row], v.z));
    s = EID_ADD(s, EID_MUL(a.m[12 + row], v.w));
    o[row] = s;
  }
  return r;
}

/* General 4x4 inverse by cofactors (nvmath::invert). Returns identity-like garbage (division by
 * zero -> inf/nan) for singular input exactly like a plain cofactor inverse would. */
EID_HD eid_mat4 eid_mat4_invert(const eid_mat4& M) {
  const float* m = M.m;
  float inv[16];
  inv[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] + m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
  inv[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] - m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
  inv[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] + m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
  inv[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] - m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
  inv[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] - m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
  inv[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] + m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
  inv[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] - m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
  inv[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] + m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
  inv[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] + m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
  inv[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] - m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
  inv[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] + m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
  inv[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] - m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
  inv[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] - m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
  inv[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] + m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
  inv[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] - m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
  inv[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] + m[8] * m[1] * m[6] - m[8] * m[2] * m[5];
  float det = m[0] * inv[0] + m[1] * inv[4] + m[2] * inv[8] + m[3] * inv[12];
  float idet = 1.0f / det;
  eid_mat4 r;
  for (int i = 0; i < 16; ++i) r.m[i] = inv[i] * idet;
  return r;
}

/* tan via the deterministic sin/cos so both build hosts agree bit-for-bit */
EID_HD float eid_tanf(float x) { float s, c; eid_sincosf(x, &s, &c); return s / c; }

/* nvmath::perspectiveVK(fovDeg, aspect, n, f) — see SURVEY.md §8c */
EID_HD eid_mat4 eid_perspectiveVK(float fovDeg, float aspect, float n, float f) {
  const float DEG2RAD = 0.01745329251994329577f;
  float t = eid_tanf(fovDeg * DEG2RAD * 0.5f);
  eid_mat4 r;
  for (int i = 0; i < 16; ++i) r.m[i] = 0.0f;
  r.m[0 * 4 + 0] = 1.0f / (aspect * t);   /* a00 */
  r.m[1 * 4 + 1] = -1.0f / t;             /* a11 (Vulkan y flip) */
  r.m[2 * 4 + 2] = -f / (f - n);          /* a22 */
  r.m[2 * 4 + 3] = -1.0f;                 /* a32 */
  r.m[3 * 4 + 2] = f * n / (n - f);       /* a23 */
  return r;
}

/* right-handed look-at (CameraManip.getMatrix()) */
EID_HD eid_mat4 eid_look_at(eid_vec3 eye, eid_vec3 center, eid_vec3 up) {
  float fx = center.x - eye.x, fy = center.y - eye.y, fz = center.z - eye.z;
  float fl = 1.0f / eid_sqrtf(fx * fx + fy * fy + fz * fz);
  fx *= fl; fy *= fl; fz *= fl;
  /* s = f x up */
  float sx = fy * up.z - fz * up.y, sy = fz * up.x - fx * up.z, sz = fx * up.y - fy * up.x;
  float sl = 1.0f / eid_sqrtf(sx * sx + sy * sy + sz * sz);
  sx *= sl; sy *= sl; sz *= sl;
  /* u = s x f */
  float ux = sy * fz - sz * fy, uy = sz * fx - sx * fz, uz = sx * fy - sy * fx;
  eid_mat4 r;
  r.m[0] = sx;  r.m[4] = sy;  r.m[8]  = sz;  r.m[12] = -(sx * eye.x + sy * eye.y + sz * eye.z);
  r.m[1] = ux;  r.m[5] = uy;  r.m[9]  = uz;  r.m[13] = -(ux * eye.x + uy * eye.y + uz * eye.z);
  r.m[2] = -fx; r.m[6] = -fy; r.m[10] = -fz; r.m[14] = (fx * eye.x + fy * eye.y + fz * eye.z);
  r.m[3] = 0.f; r.m[7] = 0.f; r.m[11] = 0.f; r.m[15] = 1.0f;
  return r;
}

#endif /* EIDOLA_VECMATH_H */

// trace.cuh — software replacement for the reference's ray-query traversal (shaders/traceray_rq.glsl:
// ClosestHit :108-147, AnyHit :153-185), which runs on RT cores inside the Vulkan driver.
//
// Semantics: t in (0, tmax) exclusive; back faces culled unless the
// instance disables culling (accelstruct.cpp:148-149), facing decided in object space; barycentrics (u,v)
// weight vertices 1 and 2; closest hit = smallest t, ties broken by lowest (instanceID, primitiveID).
// The triangle test is Moller-Trumbore in individually rounded fp32 operations — the one piece of
// arithmetic that decides results, identical to the oracle's.  Box tests are conservative (boxes padded at
// build time) and use fmaf; they only decide which triangles get tested, never the outcome.
#pragma once
#include "accel.h"
#include "dmath.cuh"

namespace eid {

struct RayHit {
  float t, u, v;
  int tri;          // index into AccelView::tris, -1 = miss
  int prim, inst;
  uint32_t flags;   // triangle flags of the hit (INST_FORCE_OPAQUE ...)
};
// Candidates are totally ordered by (t, instanceID, primitiveID).  LOWER traversals only accept candidates strictly after `low`:
// the stochastic-alpha loop visits candidates front to back by re-querying with the last rejected candidate as the bound.
struct HitKey { float t; int inst, prim; };

#define EID_STACK_SIZE 128

// returns true and fills t,u,v when the ray hits triangle (v0,e1,e2) inside (0, tmax)
DEV bool triangleTest(f3 v0, f3 e1, f3 e2, uint32_t flags, f3 o, f3 d, float tmax, float& t, float& u, float& v) {
#ifndef EID_TRI_EARLYOUT
  // every quantity evaluated, one combined accept predicate: the same booleans as the early-out form below (a rejected
  // candidate's u, v, t are never used), without the four divergence points inside the leaf loop
  const f3 pvec = cross3(d, e2);
  const float det = dot3(e1, pvec);
  const float fd = (flags & INST_MIRROR) ? -det : det;
  const bool okDet = (flags & INST_CULL_DISABLE) ? (det != 0.0f) : (fd > 0.0f);
  const float inv = __fdiv_rn(1.0f, det);
  const f3 tvec = o - v0;
  u = __fmul_rn(dot3(tvec, pvec), inv);
  const f3 qvec = cross3(tvec, e1);
  v = __fmul_rn(dot3(d, qvec), inv);
  t = __fmul_rn(dot3(e2, qvec), inv);
  return okDet && (u >= 0.0f && u <= 1.0f) && (v >= 0.0f && __fadd_rn(u, v) <= 1.0f) && (t > 0.0f && t < tmax);
#else
  f3 pvec = cross3(d, e2);
  float det = dot3(e1, pvec);
  if (flags & INST_CULL_DISABLE) { if (det == 0.0f) return false; }
  else { float fd = (flags & INST_MIRROR) ? -det : det; if (!(fd > 0.0f)) return false; }
  float inv = __fdiv_rn(1.0f, det);
  f3 tvec = o - v0;
  u = __fmul_rn(dot3(tvec, pvec), inv);
  if (!(u >= 0.0f && u <= 1.0f)) return false;
  f3 qvec = cross3(tvec, e1);
  v = __fmul_rn(dot3(d, qvec), inv);
  if (!(v >= 0.0f && __fadd_rn(u, v) <= 1.0f)) return false;
  t = __fmul_rn(dot3(e2, qvec), inv);
  return t > 0.0f && t < tmax;
#endif
}

struct RayBox {   // per-ray constants of the slab test
  float ix, iy, iz, ox, oy, oz;   // 1/d and o/d
  int nx, ny, nz;                 // BVH4 node: float4 index of the NEAR plane per axis (0..2 = lo x/y/z, 3..5 = hi x/y/z); far = near ^ 3 ...
  int fx, fy, fz;                 // ... kept explicit: {0,3} {1,4} {2,5}
};
DEV RayBox makeRayBox(f3 o, f3 d) {
  const float tiny = 1e-20f;
  float dx = fabsf(d.x) > tiny ? d.x : copysignf(tiny, d.x);
  float dy = fabsf(d.y) > tiny ? d.y : copysignf(tiny, d.y);
  float dz = fabsf(d.z) > tiny ? d.z : copysignf(tiny, d.z);
  RayBox r;
  r.ix = 1.0f / dx; r.iy = 1.0f / dy; r.iz = 1.0f / dz;
  r.ox = o.x * r.ix; r.oy = o.y * r.iy; r.oz = o.z * r.iz;
  r.nx = r.ix < 0.0f ? 3 : 0; r.fx = 3 - r.nx;
  r.ny = r.iy < 0.0f ? 4 : 1; r.fy = 5 - r.ny;
  r.nz = r.iz < 0.0f ? 5 : 2; r.fz = 7 - r.nz;
  return r;
}
// entry distance of the slab intersection, or +inf when the box is missed within [0, tbest]
DEV float boxEntry(const RayBox& rb, float lx, float ly, float lz, float hx, float hy, float hz, float tbest) {
  float x0 = fmaf(lx, rb.ix, -rb.ox), x1 = fmaf(hx, rb.ix, -rb.ox);
  float y0 = fmaf(ly, rb.iy, -rb.oy), y1 = fmaf(hy, rb.iy, -rb.oy);
  float z0 = fmaf(lz, rb.iz, -rb.oz), z1 = fmaf(hz, rb.iz, -rb.oz);
  float tn = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), 0.0f));
  float tf = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), tbest));
  return (tn <= tf * 1.0000004f) ? tn : __int_as_float(0x7f800000);
}
// the same test with the near / far plane of each axis already selected by the ray's direction signs (the BVH4 walk picks the
// float4 it loads by address, so the six min/max of the generic form disappear)
DEV float boxEntryNF(const RayBox& rb, float nx, float ny, float nz, float fx, float fy, float fz, float tbest) {
  const float tn = fmaxf(fmaxf(fmaf(nx, rb.ix, -rb.ox), fmaf(ny, rb.iy, -rb.oy)), fmaxf(fmaf(nz, rb.iz, -rb.oz), 0.0f));
  const float tf = fminf(fminf(fmaf(fx, rb.ix, -rb.ox), fmaf(fy, rb.iy, -rb.oy)), fminf(fmaf(fz, rb.iz, -rb.oz), tbest));
  return (tn <= tf * 1.0000004f) ? tn : __int_as_float(0x7f800000);
}

// ANY = true: terminate on the first accepted triangle (AnyHit); false: closest hit with tie-break.
// STATS = true additionally counts inner-node visits and triangle tests (profiling builds of the kernels).
template <bool ANY, bool STATS = false, bool LOWER = false>
DEV bool traverse(const AccelView& A, f3 o, f3 d, float tmax, RayHit& hit, unsigned int* nodeVisits = nullptr, unsigned int* triTests = nullptr,
                  HitKey low = HitKey{0.f, 0, 0}) {
  hit.t = tmax; hit.tri = -1; hit.prim = 0x7fffffff; hit.inst = 0x7fffffff; hit.u = hit.v = 0.f; hit.flags = 0;
  if (A.triCount == 0) return false;
  // a direction with NaN/zero length can never produce det != 0; skip the walk
  if (!(fabsf(d.x) + fabsf(d.y) + fabsf(d.z) > 0.0f)) return false;
  const RayBox rb = makeRayBox(o, d);
  int stack[EID_STACK_SIZE];
  int sp = 0;
  int cur = A.rootRef;
  const float INF = __int_as_float(0x7f800000);
  const int DONE = (int)0x80000000;            // never a valid reference (it would be a leaf starting at triangle 2^28 - 1)
#define EID_POP() (sp ? stack[--sp] : DONE)
  // "while-while" walk: every lane first descends inner nodes until it holds a leaf (or is done); the warp then reconverges
  // and intersects leaves together.  In the interleaved form the triangle tests ran with ~5 of 32 lanes active (ncu source page).
  for (;;) {
    while (cur >= 0) {
#if EID_BVH_WIDTH == 2
      const float4* n = A.nodes + 4 * (size_t)cur;
      if (STATS) ++*nodeVisits;
      const float4 q0 = __ldg(n), q1 = __ldg(n + 1), q2 = __ldg(n + 2), q3 = __ldg(n + 3);
      float e0 = boxEntry(rb, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, hit.t);
      float e1 = boxEntry(rb, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, hit.t);
      int c0 = __float_as_int(q3.x), c1 = __float_as_int(q3.y);
      bool h0 = e0 < INF, h1 = e1 < INF;
      if (h0 && h1) {
        if (e1 < e0) { int t = c0; c0 = c1; c1 = t; }
        if (sp < EID_STACK_SIZE) stack[sp++] = c1;
        cur = c0;
      } else if (h0) cur = c0;
      else if (h1) cur = c1;
      else cur = EID_POP();
#else
      const float4* n = A.nodes + 8 * (size_t)cur;
      if (STATS) ++*nodeVisits;
      const float4 lx = __ldg(n + rb.nx), ly = __ldg(n + rb.ny), lz = __ldg(n + rb.nz);   // near planes of the 4 children
      const float4 hx = __ldg(n + rb.fx), hy = __ldg(n + rb.fy), hz = __ldg(n + rb.fz);   // far planes
      const float4 rf = __ldg(n + 6);
      float e0 = boxEntryNF(rb, lx.x, ly.x, lz.x, hx.x, hy.x, hz.x, hit.t);
      float e1 = boxEntryNF(rb, lx.y, ly.y, lz.y, hx.y, hy.y, hz.y, hit.t);
      float e2 = boxEntryNF(rb, lx.z, ly.z, lz.z, hx.z, hy.z, hz.z, hit.t);
      float e3 = boxEntryNF(rb, lx.w, ly.w, lz.w, hx.w, hy.w, hz.w, hit.t);
      int c0 = __float_as_int(rf.x), c1 = __float_as_int(rf.y), c2 = __float_as_int(rf.z), c3 = __float_as_int(rf.w);
      if (ANY) {
        // occlusion rays: order is irrelevant, just visit every child the ray enters
        int next = 0; bool have = false;
        if (e0 < INF) { next = c0; have = true; }
        if (e1 < INF) { if (have) { if (sp < EID_STACK_SIZE) stack[sp++] = c1; } else { next = c1; have = true; } }
        if (e2 < INF) { if (have) { if (sp < EID_STACK_SIZE) stack[sp++] = c2; } else { next = c2; have = true; } }
        if (e3 < INF) { if (have) { if (sp < EID_STACK_SIZE) stack[sp++] = c3; } else { next = c3; have = true; } }
        cur = have ? next : EID_POP();
      } else {
        // sort the four (entry distance, ref) pairs ascending: 5-comparator network
#define EID_CSWAP(ea, ca, eb, cb) { if (eb < ea) { float te = ea; ea = eb; eb = te; int tc = ca; ca = cb; cb = tc; } }
        EID_CSWAP(e0, c0, e1, c1) EID_CSWAP(e2, c2, e3, c3) EID_CSWAP(e0, c0, e2, c2) EID_CSWAP(e1, c1, e3, c3) EID_CSWAP(e1, c1, e2, c2)
#undef EID_CSWAP
        if (e0 < INF) {
          if (e3 < INF && sp < EID_STACK_SIZE) stack[sp++] = c3;
          if (e2 < INF && sp < EID_STACK_SIZE) stack[sp++] = c2;
          if (e1 < INF && sp < EID_STACK_SIZE) stack[sp++] = c1;
          cur = c0;
        } else cur = EID_POP();
      }
#endif
    }
    if (cur == DONE) break;
    {
      const uint32_t ref = ~(uint32_t)cur;
      const uint32_t first = ref >> 3, count = ref & 7u;
      for (uint32_t k = 0; k < count; ++k) {
        const float4* tp = A.tris + 3 * (size_t)(first + k);
        if (STATS) ++*triTests;
        const float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
        float t, u, v;
        const uint32_t flags = __float_as_uint(c.w);
        if (triangleTest(mk3(a.x, a.y, a.z), mk3(a.w, b.x, b.y), mk3(b.z, b.w, c.x), flags, o, d, tmax, t, u, v)) {
          if (ANY) { hit.t = t; hit.tri = (int)(first + k); return true; }
          const int prim = __float_as_int(c.y), inst = __float_as_int(c.z);
          if (LOWER && (t < low.t || (t == low.t && (inst < low.inst || (inst == low.inst && prim <= low.prim))))) continue;
          bool better = t < hit.t || (t == hit.t && (inst < hit.inst || (inst == hit.inst && prim < hit.prim)));
          if (hit.tri < 0 || better) { hit.t = t; hit.u = u; hit.v = v; hit.tri = (int)(first + k); hit.prim = prim; hit.inst = inst; hit.flags = flags; }
        }
      }
    }
    cur = EID_POP();
  }
#undef EID_POP
  return hit.tri >= 0;
}

}  // namespace eid

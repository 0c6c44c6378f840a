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

// Blackwell packed fp32 (FFMA2 / FMUL2: two IEEE fp32 operations per issue slot; a scalar operand is broadcast for free):
// (a.x, a.y) * s + t and (a.x, a.y) * s
DEV float2 ffma2s(float2 a, float s, float t) {
  unsigned long long ra, rs, rt, rd;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %1};" : "=l"(rs) : "f"(s));
  asm("mov.b64 %0, {%1, %1};" : "=l"(rt) : "f"(t));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rs), "l"(rt));
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
  return d;
}
DEV float2 fmul2s(float2 a, float s) {
  unsigned long long ra, rs, rd;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %1};" : "=l"(rs) : "f"(s));
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rs));
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
  return d;
}

// 256-bit global load (sm_100a LDG.E.256): one L1 wavefront per lane for 32 B instead of two
DEV void ldg256(const float4* p, float4& a, float4& b) {
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
}

#ifndef EID_Q8_TEX
#define EID_Q8_TEX 0
#endif
#ifndef EID_SORT_NEAREST
#define EID_SORT_NEAREST 0
#endif
#ifndef EID_Q8_MAGIC
#define EID_Q8_MAGIC 0
#endif
#ifndef EID_K1_SORT
#define EID_K1_SORT 0          // one-ray-per-thread walks: 1 = all entered children front to back, 0 = only the nearest selected
#endif
#ifndef EID_TQ_SORT
#define EID_TQ_SORT 1          // the same choice for the ray-queue kernel
#endif
#ifndef EID_SPECULATIVE
#define EID_SPECULATIVE 0     // one-ray-per-thread walks: postponed leaves (see traverse)
#endif
#ifndef EID_TRAV_V1
#define EID_TRAV_V1 0       // 1: the round-1 node step (scalar FFMA slab tests, 5-comparator child sort) for A/B measurements
#endif
#ifndef EID_FETCH_TEX
#define EID_FETCH_TEX 2     // 0: every BVH fetch is an LDG; 2: far planes of a node come through tex1Dfetch (default); 1/3/4: experiments
#endif

// eid_accel_build refuses a BVH whose depth could overflow the traversal stack (3 * levels + 1 < EID_STACK_SIZE), so the pushes
// need no bound check; -DEID_STACK_CHECK=1 puts the checks back (debug builds)
#ifndef EID_STACK_CHECK
#define EID_STACK_CHECK 0
#endif
#if EID_STACK_CHECK
#define EID_SP_OK(sp) ((sp) < EID_STACK_SIZE)
#else
#define EID_SP_OK(sp) true
#endif

#define EID_TRAV_DONE ((int)0x80000000)   // never a valid reference (it would be a leaf starting at triangle 2^28 - 1)

// One inner-node visit: box tests of the children, `cur` becomes the next reference to look at (nearest entered child, or the
// top of the stack, or EID_TRAV_DONE), the other entered children go onto the stack.  ANY = occlusion ray: child order is irrelevant.
// SORT (closest-hit only): all entered children in front-to-back order (incoherent rays of the ray queues); otherwise only the nearest one is
// selected (coherent rays of the one-thread-per-pixel kernels, where the full order did not lower the node count).
// pf (ray-queue kernel only): prefetch the 128-byte line of every entered child node (and the first triangle of an entered leaf) into L1, so that
// the next visit of this ray is an L1 hit instead of an L2 round trip — for rays walked with few lanes active (the drained tail of a queue, where
// the dependent-visit latency is all that is left) the extra LSU traffic is free.
DEV void prefetchL1(const void* p) { asm volatile("prefetch.global.L1 [%0];" :: "l"(p)); }
template <bool ANY, bool SORT = false>
DEV void nodeStep(const AccelView& A, const RayBox& rb, float tbest, int& cur, int* stack, int& sp, const bool pf = false) {
  const float INF = __int_as_float(0x7f800000);
#define EID_POP() (sp ? stack[--sp] : EID_TRAV_DONE)
#if EID_BVH_WIDTH == 4 && !EID_NODE_Q8 && !EID_TRAV_V1 && EID_FETCH_TEX == 2 && !defined(EID_NODE_LDG256)
  // Round-2 node step.  Same slab arithmetic as boxEntryNF (bit-identical entry distances, hence the same visited set), issued as
  // packed FFMA2 on the child pairs the SoA node layout already provides (12 instead of 24 FMA issue slots), the 1 + 4e-7 slack
  // as two packed multiplies, and instead of a 5-comparator sort of (distance, ref) pairs only the NEAREST entered child is
  // selected (min / max / select, 14 instructions); the other entered children are pushed in pair order.  The hit does not
  // depend on the visiting order (total order on (t, instance, primitive)), and on the C3 scene the full sort did not lower
  // the node count (13.5 per ray either way, profiles/README.md).
  const float4* n = A.nodes + 8 * (size_t)cur;
  const int nb = 8 * cur;
  const float4 lx = __ldg(n + rb.nx), ly = __ldg(n + rb.ny), lz = __ldg(n + rb.nz);
  const float4 hx = tex1Dfetch<float4>(A.nodeTex, nb + rb.fx), hy = tex1Dfetch<float4>(A.nodeTex, nb + rb.fy), hz = tex1Dfetch<float4>(A.nodeTex, nb + rb.fz);
  const float4 rf = __ldg(n + 6);
  const float2 ax = ffma2s(make_float2(lx.x, lx.y), rb.ix, -rb.ox), bx = ffma2s(make_float2(lx.z, lx.w), rb.ix, -rb.ox);
  const float2 ay = ffma2s(make_float2(ly.x, ly.y), rb.iy, -rb.oy), by = ffma2s(make_float2(ly.z, ly.w), rb.iy, -rb.oy);
  const float2 az = ffma2s(make_float2(lz.x, lz.y), rb.iz, -rb.oz), bz = ffma2s(make_float2(lz.z, lz.w), rb.iz, -rb.oz);
  const float2 fax = ffma2s(make_float2(hx.x, hx.y), rb.ix, -rb.ox), fbx = ffma2s(make_float2(hx.z, hx.w), rb.ix, -rb.ox);
  const float2 fay = ffma2s(make_float2(hy.x, hy.y), rb.iy, -rb.oy), fby = ffma2s(make_float2(hy.z, hy.w), rb.iy, -rb.oy);
  const float2 faz = ffma2s(make_float2(hz.x, hz.y), rb.iz, -rb.oz), fbz = ffma2s(make_float2(hz.z, hz.w), rb.iz, -rb.oz);
  const float tn0 = fmaxf(fmaxf(ax.x, ay.x), fmaxf(az.x, 0.0f)), tn1 = fmaxf(fmaxf(ax.y, ay.y), fmaxf(az.y, 0.0f));
  const float tn2 = fmaxf(fmaxf(bx.x, by.x), fmaxf(bz.x, 0.0f)), tn3 = fmaxf(fmaxf(bx.y, by.y), fmaxf(bz.y, 0.0f));
  const float2 tfa = fmul2s(make_float2(fminf(fminf(fax.x, fay.x), fminf(faz.x, tbest)), fminf(fminf(fax.y, fay.y), fminf(faz.y, tbest))), 1.0000004f);
  const float2 tfb = fmul2s(make_float2(fminf(fminf(fbx.x, fby.x), fminf(fbz.x, tbest)), fminf(fminf(fbx.y, fby.y), fminf(fbz.y, tbest))), 1.0000004f);
  float e0 = (tn0 <= tfa.x) ? tn0 : INF, e1 = (tn1 <= tfa.y) ? tn1 : INF, e2 = (tn2 <= tfb.x) ? tn2 : INF, e3 = (tn3 <= tfb.y) ? tn3 : INF;
  const int c0 = __float_as_int(rf.x), c1 = __float_as_int(rf.y), c2 = __float_as_int(rf.z), c3 = __float_as_int(rf.w);
  if (pf) {
#define EID_PF(e, c) if (e < INF) { if (c >= 0) prefetchL1(A.nodes + 8 * (size_t)c); else prefetchL1(A.tris + 3 * (size_t)((~(uint32_t)c) >> 3)); }
    EID_PF(e0, c0) EID_PF(e1, c1) EID_PF(e2, c2) EID_PF(e3, c3)
#undef EID_PF
  }
  if (ANY) {
    int next = 0; bool have = false;
    if (e0 < INF) { next = c0; have = true; }
    if (e1 < INF) { if (have) { if (EID_SP_OK(sp)) stack[sp++] = c1; } else { next = c1; have = true; } }
    if (e2 < INF) { if (have) { if (EID_SP_OK(sp)) stack[sp++] = c2; } else { next = c2; have = true; } }
    if (e3 < INF) { if (have) { if (EID_SP_OK(sp)) stack[sp++] = c3; } else { next = c3; have = true; } }
    cur = have ? next : EID_POP();
  } else if (SORT) {
    // 5-comparator network on (distance, ref) pairs, each comparator = compare + min + max + 2 selects
    int d0 = c0, d1 = c1, d2 = c2, d3 = c3;
#define EID_CSWAP2(ea, ca, eb, cb) { const bool s_ = eb < ea; const float lo_ = fminf(ea, eb), hi_ = fmaxf(ea, eb); const int cl_ = s_ ? cb : ca, ch_ = s_ ? ca : cb; ea = lo_; eb = hi_; ca = cl_; cb = ch_; }
    EID_CSWAP2(e0, d0, e1, d1) EID_CSWAP2(e2, d2, e3, d3) EID_CSWAP2(e0, d0, e2, d2) EID_CSWAP2(e1, d1, e3, d3) EID_CSWAP2(e1, d1, e2, d2)
#undef EID_CSWAP2
    if (e0 < INF) {
      if (e3 < INF && EID_SP_OK(sp)) stack[sp++] = d3;
      if (e2 < INF && EID_SP_OK(sp)) stack[sp++] = d2;
      if (e1 < INF && EID_SP_OK(sp)) stack[sp++] = d1;
      cur = d0;
    } else cur = EID_POP();
  } else {
    // nearest of each pair, then the nearer of the two winners; the three children not chosen are pushed when entered
    const bool s01 = e1 < e0, s23 = e3 < e2;
    const float w01 = fminf(e0, e1), l01 = fmaxf(e0, e1), w23 = fminf(e2, e3), l23 = fmaxf(e2, e3);
    const int cw01 = s01 ? c1 : c0, cl01 = s01 ? c0 : c1, cw23 = s23 ? c3 : c2, cl23 = s23 ? c2 : c3;
    const bool sw = w23 < w01;
    const float wn = fminf(w01, w23), wl = fmaxf(w01, w23);
    const int cn = sw ? cw23 : cw01, cl = sw ? cw01 : cw23;
    if (wn < INF) {
      if (l01 < INF && EID_SP_OK(sp)) stack[sp++] = cl01;
      if (l23 < INF && EID_SP_OK(sp)) stack[sp++] = cl23;
      if (wl < INF && EID_SP_OK(sp)) stack[sp++] = cl;          // the second winner is popped first
      cur = cn;
    } else cur = EID_POP();
  }
#elif EID_BVH_WIDTH == 2
  const float4* n = A.nodes + 4 * (size_t)cur;
  const float4 q0 = __ldg(n), q1 = __ldg(n + 1), q2 = __ldg(n + 2), q3 = __ldg(n + 3);
  float e0 = boxEntry(rb, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, tbest);
  float e1 = boxEntry(rb, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, tbest);
  int c0 = __float_as_int(q3.x), c1 = __float_as_int(q3.y);
  bool h0 = e0 < INF, h1 = e1 < INF;
  if (h0 && h1) {
    if (e1 < e0) { int t = c0; c0 = c1; c1 = t; }
    if (EID_SP_OK(sp)) stack[sp++] = c1;
    cur = c0;
  } else if (h0) cur = c0;
  else if (h1) cur = c1;
  else cur = EID_POP();
#else
#if EID_NODE_Q8
  // 64-byte node, child planes quantised to 8 bits (accel.h): 4 loads per visit instead of 7; plane distance t = q * (scale / d) +
  // (origin - o) / d, near / far byte planes selected by the ray's direction signs.  EID_Q8_TEX: the second half through the TEX pipe.
  const uint4* n = (const uint4*)A.nodes + 4 * (size_t)cur;
  const uint4 a0 = __ldg(n), a1 = __ldg(n + 1);
#if EID_Q8_TEX
  const uint4 a2 = tex1Dfetch<uint4>(A.nodeTex, 4 * cur + 2), a3 = tex1Dfetch<uint4>(A.nodeTex, 4 * cur + 3);   // (the node texture is uint4 in this build)
#else
  const uint4 a2 = __ldg(n + 2), a3 = __ldg(n + 3);
#endif
  const float ax = __uint_as_float(a0.w) * rb.ix, ay = __uint_as_float(a1.x) * rb.iy, az = __uint_as_float(a1.y) * rb.iz;
  const float bx = fmaf(__uint_as_float(a0.x), rb.ix, -rb.ox), by = fmaf(__uint_as_float(a0.y), rb.iy, -rb.oy), bz = fmaf(__uint_as_float(a0.z), rb.iz, -rb.oz);
  const bool sx = rb.nx != 0, sy = rb.ny != 1, sz = rb.nz != 2;      // direction component negative: the upper plane is the near one
  const uint32_t nqx = sx ? a2.y : a1.z, fqx = sx ? a1.z : a2.y;
  const uint32_t nqy = sy ? a2.z : a1.w, fqy = sy ? a1.w : a2.z;
  const uint32_t nqz = sz ? a2.w : a2.x, fqz = sz ? a2.x : a2.w;
  float e0, e1, e2, e3;
#if EID_Q8_MAGIC
  // byte -> float without the conversion unit: the byte becomes the low mantissa byte of 2^23 (one PRMT), minus 2^23 (exact)
#define EID_Q8_F(w, sh) (__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7540u | ((sh) >> 3))) - 8388608.0f)
#else
#define EID_Q8_F(w, sh) ((float)(((w) >> (sh)) & 255u))
#endif
#define EID_Q8_CHILD(e, sh, ref) { \
    const float tn = fmaxf(fmaxf(fmaf(EID_Q8_F(nqx, sh), ax, bx), fmaf(EID_Q8_F(nqy, sh), ay, by)), fmaxf(fmaf(EID_Q8_F(nqz, sh), az, bz), 0.0f)); \
    const float tf = fminf(fminf(fmaf(EID_Q8_F(fqx, sh), ax, bx), fmaf(EID_Q8_F(fqy, sh), ay, by)), fminf(fmaf(EID_Q8_F(fqz, sh), az, bz), tbest)); \
    e = (tn <= tf * 1.0000004f && (ref) != 0xffffffffu) ? tn : INF; }
  EID_Q8_CHILD(e0, 0, a3.x) EID_Q8_CHILD(e1, 8, a3.y) EID_Q8_CHILD(e2, 16, a3.z) EID_Q8_CHILD(e3, 24, a3.w)
#undef EID_Q8_CHILD
#undef EID_Q8_F
  const float4 rf = make_float4(__uint_as_float(a3.x), __uint_as_float(a3.y), __uint_as_float(a3.z), __uint_as_float(a3.w));
#else
  const float4* n = A.nodes + 8 * (size_t)cur;
#ifdef EID_NODE_LDG256
  // the whole 128-byte node in four 256-bit loads (4 L1 wavefronts per lane instead of 7); near/far by min/max
  float4 lx, ly, lz, hx, hy, hz, rf, pad_;
  ldg256(n, lx, ly); ldg256(n + 2, lz, hx); ldg256(n + 4, hy, hz); ldg256(n + 6, rf, pad_);
  float e0 = boxEntry(rb, lx.x, ly.x, lz.x, hx.x, hy.x, hz.x, tbest);
  float e1 = boxEntry(rb, lx.y, ly.y, lz.y, hx.y, hy.y, hz.y, tbest);
  float e2 = boxEntry(rb, lx.z, ly.z, lz.z, hx.z, hy.z, hz.z, tbest);
  float e3 = boxEntry(rb, lx.w, ly.w, lz.w, hx.w, hy.w, hz.w, tbest);
#elif EID_FETCH_TEX
  // Far planes through the TEX pipe, near planes + references through the LSU pipe: an incoherent node visit is one L1 data
  // wavefront per lane per 16 bytes, and with all seven loads on the LSU pipe that pipe ran at 72-81 % of its peak in the trace
  // kernels (ncu l1tex__data_pipe_lsu_wavefronts); split over both pipes K1 went 0.995 -> 0.926 ms (profiles/README.md).
  const int nb = 8 * cur;
#if EID_FETCH_TEX >= 2
  const float4 lx = __ldg(n + rb.nx), ly = __ldg(n + rb.ny), lz = __ldg(n + rb.nz);
#else
  const float4 lx = tex1Dfetch<float4>(A.nodeTex, nb + rb.nx), ly = tex1Dfetch<float4>(A.nodeTex, nb + rb.ny), lz = tex1Dfetch<float4>(A.nodeTex, nb + rb.nz);
#endif
  const float4 hx = tex1Dfetch<float4>(A.nodeTex, nb + rb.fx), hy = tex1Dfetch<float4>(A.nodeTex, nb + rb.fy), hz = tex1Dfetch<float4>(A.nodeTex, nb + rb.fz);
  const float4 rf = __ldg(n + 6);
  float e0 = boxEntryNF(rb, lx.x, ly.x, lz.x, hx.x, hy.x, hz.x, tbest);
  float e1 = boxEntryNF(rb, lx.y, ly.y, lz.y, hx.y, hy.y, hz.y, tbest);
  float e2 = boxEntryNF(rb, lx.z, ly.z, lz.z, hx.z, hy.z, hz.z, tbest);
  float e3 = boxEntryNF(rb, lx.w, ly.w, lz.w, hx.w, hy.w, hz.w, tbest);
#else
  const float4 lx = __ldg(n + rb.nx), ly = __ldg(n + rb.ny), lz = __ldg(n + rb.nz);   // near planes of the 4 children
  const float4 hx = __ldg(n + rb.fx), hy = __ldg(n + rb.fy), hz = __ldg(n + rb.fz);   // far planes
  const float4 rf = __ldg(n + 6);
  float e0 = boxEntryNF(rb, lx.x, ly.x, lz.x, hx.x, hy.x, hz.x, tbest);
  float e1 = boxEntryNF(rb, lx.y, ly.y, lz.y, hx.y, hy.y, hz.y, tbest);
  float e2 = boxEntryNF(rb, lx.z, ly.z, lz.z, hx.z, hy.z, hz.z, tbest);
  float e3 = boxEntryNF(rb, lx.w, ly.w, lz.w, hx.w, hy.w, hz.w, tbest);
#endif
#endif   // EID_NODE_Q8
  int c0 = __float_as_int(rf.x), c1 = __float_as_int(rf.y), c2 = __float_as_int(rf.z), c3 = __float_as_int(rf.w);
  if (ANY) {
    // occlusion rays: order is irrelevant, just visit every child the ray enters
    int next = 0; bool have = false;
    if (e0 < INF) { next = c0; have = true; }
    if (e1 < INF) { if (have) { if (EID_SP_OK(sp)) stack[sp++] = c1; } else { next = c1; have = true; } }
    if (e2 < INF) { if (have) { if (EID_SP_OK(sp)) stack[sp++] = c2; } else { next = c2; have = true; } }
    if (e3 < INF) { if (have) { if (EID_SP_OK(sp)) stack[sp++] = c3; } else { next = c3; have = true; } }
    cur = have ? next : EID_POP();
  } else {
#define EID_CSWAP(ea, ca, eb, cb) { if (eb < ea) { float te = ea; ea = eb; eb = te; int tc = ca; ca = cb; cb = tc; } }
#if EID_SORT_NEAREST
    // only the nearest child is brought to the front (3 comparators); the other entered children are pushed unordered
    EID_CSWAP(e0, c0, e1, c1) EID_CSWAP(e2, c2, e3, c3) EID_CSWAP(e0, c0, e2, c2)
#else
    // sort the four (entry distance, ref) pairs ascending: 5-comparator network
    EID_CSWAP(e0, c0, e1, c1) EID_CSWAP(e2, c2, e3, c3) EID_CSWAP(e0, c0, e2, c2) EID_CSWAP(e1, c1, e3, c3) EID_CSWAP(e1, c1, e2, c2)
#endif
#undef EID_CSWAP
    if (e0 < INF) {
      if (e3 < INF && EID_SP_OK(sp)) stack[sp++] = c3;
      if (e2 < INF && EID_SP_OK(sp)) stack[sp++] = c2;
      if (e1 < INF && EID_SP_OK(sp)) stack[sp++] = c1;
      cur = c0;
    } else cur = EID_POP();
  }
#endif
}

// All triangles of leaf reference `cur` (< 0, != EID_TRAV_DONE).  Returns true when an ANY ray is finished (first accepted triangle).
template <bool ANY, bool STATS, bool LOWER>
DEV bool leafStep(const AccelView& A, int cur, f3 o, f3 d, float tmax, RayHit& hit, unsigned int* triTests, const HitKey& low) {
  const uint32_t ref = ~(uint32_t)cur;
  const uint32_t first = ref >> 3, count = ref & 7u;
  for (uint32_t k = 0; k < count; ++k) {
    const float4* tp = A.tris + 3 * (size_t)(first + k);
    if (STATS) ++*triTests;
#if EID_FETCH_TEX >= 3
    const int tb = 3 * (int)(first + k);
#if EID_FETCH_TEX == 4
    const float4 a = __ldg(tp), b = tex1Dfetch<float4>(A.triTex, tb + 1), c = __ldg(tp + 2);
#else
    const float4 a = tex1Dfetch<float4>(A.triTex, tb), b = tex1Dfetch<float4>(A.triTex, tb + 1), c = tex1Dfetch<float4>(A.triTex, tb + 2);
#endif
#else
    const float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
#endif
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
  return false;
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
  // "while-while" walk: every lane first descends inner nodes until it holds a leaf (or is done); the warp then reconverges
  // and intersects leaves together.  In the interleaved form the triangle tests ran with ~5 of 32 lanes active (ncu source page).
#if EID_SPECULATIVE
  // Speculative while-while (Aila & Laine 2009): a lane that reaches a leaf postpones it and keeps descending, so it does useful node work
  // while the other lanes of the warp are still looking for theirs; the warp switches to the leaf phase when every lane holds a leaf
  // (EID_SPECULATIVE & 1: decided by a vote) or when the lane meets its second leaf.  & 2: closest-hit rays only (an occlusion ray that
  // postpones a leaf may walk on past its terminating hit).  The visiting order never changes a result.
  const bool spec = !(ANY && (EID_SPECULATIVE & 2));
  for (;;) {
    int leaf = EID_TRAV_DONE;                                  // the postponed leaf (EID_TRAV_DONE = none)
    while (cur >= 0) {
      if (STATS) ++*nodeVisits;
      nodeStep<ANY, EID_K1_SORT != 0>(A, rb, hit.t, cur, stack, sp);
      if (spec && cur < 0 && cur != EID_TRAV_DONE && leaf == EID_TRAV_DONE) { leaf = cur; cur = EID_POP(); }
      if ((EID_SPECULATIVE & 1) && spec && !__any_sync(__activemask(), leaf == EID_TRAV_DONE)) break;
    }
    if (leaf != EID_TRAV_DONE && leafStep<ANY, STATS, LOWER>(A, leaf, o, d, tmax, hit, triTests, low)) return true;
    if (cur < 0) {
      if (cur == EID_TRAV_DONE) break;
      if (leafStep<ANY, STATS, LOWER>(A, cur, o, d, tmax, hit, triTests, low)) return true;
      cur = EID_POP();
    }
  }
  return hit.tri >= 0;
#else
  for (;;) {
    while (cur >= 0) {
      if (STATS) ++*nodeVisits;
      nodeStep<ANY, EID_K1_SORT != 0>(A, rb, hit.t, cur, stack, sp);
    }
    if (cur == EID_TRAV_DONE) break;
    if (leafStep<ANY, STATS, LOWER>(A, cur, o, d, tmax, hit, triTests, low)) return true;
    cur = EID_POP();
  }
  return hit.tri >= 0;
#endif
}

// ------------------------------------------------------------------------------------------------------------------------------
// Two-level traversal (AccelStructure::create builds one BLAS per prim mesh and one TLAS instance per glTF node,
// accelstruct.cpp:55-162): a top-level BVH4 over the instances' world boxes, and per prim mesh ONE bottom-level BVH4 over its
// object-space triangles, shared by all its instances — memory no longer grows with the instance count.  The walk enters an instance
// by taking the ray to object space (box tests only: conservative, with a slack per axis that covers the rounding of that transform) and
// leaves it through a sentinel on the stack.  What DECIDES a hit is unchanged: the triangle's object-space vertices are taken to world
// space with the instance matrix in the contract arithmetic (exactly what the flat build bakes, accel.cu: k_emit_triangles) and
// Moller-Trumbore runs on the WORLD ray — so every hit, barycentric and tie-break is bit-identical to the flat BVH's.
// Object triangle record (48 B): t0 = p0.xyz, p1.x   t1 = p1.yz, p2.xy   t2 = p2.z, primitiveID, -, -
// TLAS primitive record (48 B):  t0 = instance index, BLAS root reference (as int bits), -, -
// ------------------------------------------------------------------------------------------------------------------------------
#define EID_TRAV_LEAVE_INSTANCE ((int)0x80000001)      // stack sentinel (like EID_TRAV_DONE never a valid reference)

DEV float boxEntrySlack(const RayBox& rb, f3 sl, float nx, float ny, float nz, float fx, float fy, float fz, float tbest) {
  const float tn = fmaxf(fmaxf(fmaf(nx, rb.ix, -rb.ox) - sl.x, fmaf(ny, rb.iy, -rb.oy) - sl.y), fmaxf(fmaf(nz, rb.iz, -rb.oz) - sl.z, 0.0f));
  const float tf = fminf(fminf(fmaf(fx, rb.ix, -rb.ox) + sl.x, fmaf(fy, rb.iy, -rb.oy) + sl.y), fminf(fmaf(fz, rb.iz, -rb.oz) + sl.z, tbest));
  return (tn <= tf * 1.000002f) ? tn : __int_as_float(0x7f800000);
}
// one inner-node visit of either level: the entered children are pushed far to near (closest hit) / in slot order (occlusion)
template <bool ANY>
DEV void nodeStep2(const float4* __restrict__ nodes, const RayBox& rb, f3 sl, float tbest, int& cur, int* stack, int& sp) {
  const float INF = __int_as_float(0x7f800000);
  const float4* n = nodes + 8 * (size_t)cur;
  const float4 lx = __ldg(n + rb.nx), ly = __ldg(n + rb.ny), lz = __ldg(n + rb.nz);
  const float4 hx = __ldg(n + rb.fx), hy = __ldg(n + rb.fy), hz = __ldg(n + rb.fz);
  const float4 rf = __ldg(n + 6);
  float e[4] = {boxEntrySlack(rb, sl, lx.x, ly.x, lz.x, hx.x, hy.x, hz.x, tbest), boxEntrySlack(rb, sl, lx.y, ly.y, lz.y, hx.y, hy.y, hz.y, tbest),
                boxEntrySlack(rb, sl, lx.z, ly.z, lz.z, hx.z, hy.z, hz.z, tbest), boxEntrySlack(rb, sl, lx.w, ly.w, lz.w, hx.w, hy.w, hz.w, tbest)};
  int c[4] = {__float_as_int(rf.x), __float_as_int(rf.y), __float_as_int(rf.z), __float_as_int(rf.w)};
  if (!ANY) {
#define EID_CSWAP(a, b) { if (e[b] < e[a]) { const float te = e[a]; e[a] = e[b]; e[b] = te; const int tc = c[a]; c[a] = c[b]; c[b] = tc; } }
    EID_CSWAP(0, 1) EID_CSWAP(2, 3) EID_CSWAP(0, 2) EID_CSWAP(1, 3) EID_CSWAP(1, 2)
#undef EID_CSWAP
  }
  int next = EID_TRAV_DONE; bool have = false;
#pragma unroll
  for (int k = 3; k >= 0; --k) {
    if (e[k] < INF) {
      if (have) stack[sp++] = next;               // what was found so far is farther (or, for occlusion rays, just another child)
      next = c[k]; have = true;
    }
  }
  cur = have ? next : stack[--sp];                // the stack always holds at least the bottom sentinel
}

// (not inlined: the callers are the already register-starved full variants of the stage kernels, whose flat path must not pay for this one)
template <bool ANY, bool STATS = false, bool LOWER = false>
__device__ __noinline__ bool traverse2(const AccelView& A, f3 o, f3 d, float tmax, RayHit& hit, unsigned int* nodeVisits = nullptr, unsigned int* triTests = nullptr,
                   HitKey low = HitKey{0.f, 0, 0}) {
  hit.t = tmax; hit.tri = -1; hit.prim = 0x7fffffff; hit.inst = 0x7fffffff; hit.u = hit.v = 0.f; hit.flags = 0;
  if (A.tlasPrimCount == 0) return false;
  if (!(fabsf(d.x) + fabsf(d.y) + fabsf(d.z) > 0.0f)) return false;
  int stack[EID_STACK_SIZE];
  int sp = 0;
  stack[sp++] = EID_TRAV_DONE;                    // bottom sentinel: popping it ends the walk
  RayBox rb = makeRayBox(o, d);
  f3 sl = mk3(0.f);                               // box-test slack of the current level (0 in world space)
  const float4* nodes = A.tlasNodes;
  int inst = -1;                                  // instance the walk is inside of (-1: top level)
  uint32_t iflags = 0;
  int cur = A.tlasRootRef;
  for (;;) {
    while (cur >= 0) {
      if (STATS) ++*nodeVisits;
      nodeStep2<ANY>(nodes, rb, sl, hit.t, cur, stack, sp);
    }
    if (cur == EID_TRAV_DONE) break;
    if (cur == EID_TRAV_LEAVE_INSTANCE) {         // back to the top level
      inst = -1; nodes = A.tlasNodes; rb = makeRayBox(o, d); sl = mk3(0.f);
      cur = stack[--sp];
      continue;
    }
    const uint32_t ref = ~(uint32_t)cur;
    const uint32_t first = ref >> 3, count = ref & 7u;
    if (inst < 0) {
      // top-level leaf: `count` instances.  Enter the first one now, leave the others on the stack as single-instance leaves.
      if (count == 0) { cur = stack[--sp]; continue; }
      for (uint32_t k = count - 1; k >= 1; --k) stack[sp++] = ~(int)(((first + k) << 3) | 1u);
      const float4 rec = __ldg(A.tlasPrims + 3 * (size_t)first);
      inst = __float_as_int(rec.x);
      const InstanceXform& X = A.instances[inst];
      iflags = X.flags & (INST_CULL_DISABLE | INST_MIRROR | INST_FORCE_OPAQUE);
      // the ray in object space, for box tests only: plain fmaf arithmetic + a slack per axis that bounds its rounding error
      // (|error of a transformed point| <= 4 ulp * sum |W_kj| |o_j| + |W_k3|; times 1 / |d_k| in the slab parameter; 16x safety)
      const float* W = X.worldToObject;
      const f3 oo = mk3(fmaf(W[0], o.x, fmaf(W[3], o.y, fmaf(W[6], o.z, W[9]))), fmaf(W[1], o.x, fmaf(W[4], o.y, fmaf(W[7], o.z, W[10]))),
                        fmaf(W[2], o.x, fmaf(W[5], o.y, fmaf(W[8], o.z, W[11]))));
      const f3 dd = mk3(fmaf(W[0], d.x, fmaf(W[3], d.y, W[6] * d.z)), fmaf(W[1], d.x, fmaf(W[4], d.y, W[7] * d.z)), fmaf(W[2], d.x, fmaf(W[5], d.y, W[8] * d.z)));
      const f3 mag = mk3(fabsf(W[0] * o.x) + fabsf(W[3] * o.y) + fabsf(W[6] * o.z) + fabsf(W[9]), fabsf(W[1] * o.x) + fabsf(W[4] * o.y) + fabsf(W[7] * o.z) + fabsf(W[10]),
                         fabsf(W[2] * o.x) + fabsf(W[5] * o.y) + fabsf(W[8] * o.z) + fabsf(W[11]));
      rb = makeRayBox(oo, dd);
      sl = mk3(4e-6f * mag.x * fabsf(rb.ix), 4e-6f * mag.y * fabsf(rb.iy), 4e-6f * mag.z * fabsf(rb.iz));
      nodes = A.nodes;
      stack[sp++] = EID_TRAV_LEAVE_INSTANCE;
      cur = __float_as_int(rec.y);                // root of the prim mesh's bottom-level tree (an inner node or a leaf)
      continue;
    }
    // bottom-level leaf: object-space triangles -> world space (contract arithmetic) -> Moller-Trumbore on the world ray
    const InstanceXform& X = A.instances[inst];
    bool done = false;
    for (uint32_t k = 0; k < count; ++k) {
      const float4* tp = A.tris + 3 * (size_t)(first + k);
      if (STATS) ++*triTests;
      const float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
      const f3 p0 = xfPoint(X.objectToWorld, mk3(a.x, a.y, a.z)), p1 = xfPoint(X.objectToWorld, mk3(a.w, b.x, b.y)), p2 = xfPoint(X.objectToWorld, mk3(b.z, b.w, c.x));
      float t, u, v;
      if (triangleTest(p0, p1 - p0, p2 - p0, iflags, o, d, tmax, t, u, v)) {
        if (ANY) { hit.t = t; hit.tri = (int)(first + k); done = true; break; }
        const int prim = __float_as_int(c.y);
        if (LOWER && (t < low.t || (t == low.t && (inst < low.inst || (inst == low.inst && prim <= low.prim))))) continue;
        const bool better = t < hit.t || (t == hit.t && (inst < hit.inst || (inst == hit.inst && prim < hit.prim)));
        if (hit.tri < 0 || better) { hit.t = t; hit.u = u; hit.v = v; hit.tri = (int)(first + k); hit.prim = prim; hit.inst = inst; hit.flags = iflags; }
      }
    }
    if (done) return true;
    cur = stack[--sp];
  }
  return hit.tri >= 0;
}

// ------------------------------------------------------------------------------------------------------------------------------
// Ray-queue traversal with dynamic fetch (wavefront stages): a persistent grid; every lane walks one ray of the queue at a
// time and, when its ray ends, takes the next queue entry, so a warp keeps (nearly) all lanes busy although the rays' visit
// counts differ by 10x (mean 13.5 nodes, worst ~150 on the C3 scene; in the one-ray-per-thread kernels a warp runs as long as
// its slowest lane and traversal executes with ~10 of 32 lanes active — ncu source page, profiles/README.md).
// Queue entry = 2 x float4: (origin.xyz, w0), (direction.xyz, id bits).  ANY: w0 = tmax, result occl[id] = 1/0.
// Closest: tmax = 1e28 (w0 is the producer's payload), result hits[entry] = (t, u, v, triangle index bits or -1).
// ------------------------------------------------------------------------------------------------------------------------------
#ifndef EID_TQ_MIN_BLOCKS
#define EID_TQ_MIN_BLOCKS 6
#endif
// Inner-node visits per round before the warp reconverges for the leaf phase.  The one-ray-per-thread kernels use the
// "while-while" form (descend until a leaf); with the lanes refilled from the queue that form wastes ~70 % of the issue
// slots on lanes waiting at a leaf (a warp-level model of this loop with the measured 13.5 node / 2.2 leaf visits per ray gives
// 24 % lane efficiency, ncu measured 11-12 of 32), while a short bounded node phase keeps ~70 % of the lanes busy.
#ifndef EID_TQ_NODE_STEPS
#define EID_TQ_NODE_STEPS 4
#endif
#ifndef EID_TQ_REFILL
#define EID_TQ_REFILL 1          // idle lanes that trigger a queue fetch
#endif
#ifndef EID_TQ_STEAL
#define EID_TQ_STEAL 0           // 1: warp-level work stealing once the queue has drained (see k_trace_queue).  Built, bit-identical, and
                                 // measured SLOWER (indirect_stage 0.602 -> 0.633-0.658 ms for every threshold tried, profiles/README.md): the
                                 // queue launches are bound by the latency of TYPICAL rays on under-filled SMs, not by a few very long ones
#endif
#ifndef EID_TQ_STATIC_FIRST
#define EID_TQ_STATIC_FIRST 1    // first batch of every warp assigned statically, no atomic (see k_trace_queue)
#endif
#ifndef EID_TQ_SPREAD
#define EID_TQ_SPREAD 0          // n > 0: small queues are spread over all warps, at least n rays per warp (see k_trace_queue)
#endif
#ifndef EID_TQ_PREFETCH
#define EID_TQ_PREFETCH 0        // 1: L1 prefetch of the entered children once the warp can fetch no more rays (tail), 2: always (see nodeStep)
#endif
#ifndef EID_TQ_STEALS_PER_ROUND
#define EID_TQ_STEALS_PER_ROUND 4
#endif
#ifndef EID_TQ_STEAL_MIN_IDLE
#define EID_TQ_STEAL_MIN_IDLE 16      // lanes of the warp that must be idle before anything is stolen (the tail, not the bulk)
#endif
#ifndef EID_TQ_STEAL_MIN_AGE
#define EID_TQ_STEAL_MIN_AGE 24       // node visits a piece must have behind it before it can be robbed (only long rays are worth splitting)
#endif
// Active-lane re-balancing for the tail.  A small queue is latency-bound: the longest ray of the C3 scene needs ~150 dependent node
// visits while the mean is 13.5, so once the queue has drained a warp sits on one or two long rays with 30 idle lanes, and every queue
// launch of the wavefront stages pays that tail.  From the moment a warp can fetch no more rays, an idle lane STEALS the oldest stack
// entry (the largest pending subtree) of the lane with the most pending entries, together with the ray and its best hit so far, and
// walks that subtree itself; thieves can be robbed in turn.  Every finished piece is sent to the lane that OWNS the ray (warp
// shuffles), which keeps the best hit under the total order (t, instance, primitive) — the result is the one an undivided walk finds,
// whatever the split — and writes it when its last piece has arrived.  Occlusion rays stop all their pieces at the first hit.
struct StealSlot { float t, u, v; int tri, prim, inst; int outstanding; int pad; };
template <bool ANY, bool STATS>
__global__ void __launch_bounds__(128, EID_TQ_MIN_BLOCKS) k_trace_queue(const AccelView A, const float4* __restrict__ rays, const uint32_t* __restrict__ countPtr,
                                                                        uint32_t* __restrict__ cursor, float4* __restrict__ hits, uint32_t* __restrict__ occl,
                                                                        unsigned long long* __restrict__ counters, unsigned long long* __restrict__ totals) {
  const uint32_t n = *countPtr;
  const unsigned lane = threadIdx.x & 31u;
  if (blockIdx.x == 0 && threadIdx.x == 0 && n) {   // ray counters of the frame ([0] closest, [1] any) and since creation ([5], [6])
    atomicAdd(&counters[ANY ? 1 : 0], (unsigned long long)n); atomicAdd(&totals[ANY ? 6 : 5], (unsigned long long)n);
  }
#if EID_TQ_STEAL
  __shared__ StealSlot s_slot[128];
  StealSlot& mySlot = s_slot[threadIdx.x];
  StealSlot* warpSlots = s_slot + (threadIdx.x & ~31u);
  bool finalPhase = false;
  int owner = (int)lane;           // lane of this warp that owns the ray whose piece this lane is walking
  uint32_t ownDst = 0;             // where the ray this lane OWNS reports to: hits[entry] / occl[id]
#endif
  // rays a warp holds at a time: 32, or (EID_TQ_SPREAD = s > 0) a queue smaller than the grid spread over ALL warps, at least s rays each
  const uint32_t warps = gridDim.x * (blockDim.x >> 5);
  const int per = EID_TQ_SPREAD ? (int)min(32u, max((uint32_t)EID_TQ_SPREAD, (n + warps - 1u) / warps)) : 32;
  const uint32_t firstBatch = EID_TQ_STATIC_FIRST ? warps * (uint32_t)per : 0u;
  bool first = true;
  int stack[EID_STACK_SIZE];
  int sp = 0, sb = 0, cur = EID_TRAV_DONE;   // valid stack entries: stack[sb .. sb + sp)
  int age = 0;                               // node visits of the piece this lane is walking
  bool active = false, more = true;
  uint32_t entry = 0, id = 0;
  f3 o = mk3(0.f), d = mk3(0.f);
  float tmax = 0.f;
  RayBox rb = makeRayBox(o, mk3(1.f));
  RayHit hit; hit.t = 0.f; hit.tri = -1; hit.prim = hit.inst = 0x7fffffff; hit.u = hit.v = 0.f; hit.flags = 0;
  unsigned int nodeVisits = 0, triTests = 0, rayVisits = 0;
  const HitKey low = {0.f, 0, 0};
  for (;;) {
    const unsigned idle = __ballot_sync(0xffffffffu, !active);
    uint32_t my = n;                                   // the queue entry this lane takes now (n = none)
    if (EID_TQ_STATIC_FIRST && first) {
      // The first batch is assigned statically (warp w takes the entries [w * per, (w + 1) * per)): 3552 warps hitting ONE cursor word with
      // an atomicAdd at the same moment serialise in the L2 — a floor of every queue launch, however few rays it holds.  The cursor only
      // counts the entries fetched dynamically after that grid-wide batch.
      first = false;
      const uint32_t wid = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
      more = firstBatch < n;
      if (lane < (unsigned)per) my = wid * (uint32_t)per + lane;
    } else if (more && 32 - __popc(idle) < per) {
      // dynamic fetch: one atomicAdd per warp and refill round tops the warp up to `per` rays
      const int nIdle = per - (32 - __popc(idle)), leader = __ffs(idle) - 1;
      uint32_t base = 0;
      if ((int)lane == leader) base = atomicAdd(cursor, (uint32_t)nIdle);
      base = __shfl_sync(0xffffffffu, base, leader) + firstBatch;
      if (base + (uint32_t)nIdle >= n) more = false;
      const uint32_t rank = (uint32_t)__popc(idle & ((1u << lane) - 1u));
      if (!active && rank < (uint32_t)nIdle) my = base + rank;
    }
    if (my < n) {
      const float4 r0 = __ldg(rays + 2 * (size_t)my), r1 = __ldg(rays + 2 * (size_t)my + 1);
      entry = my; id = __float_as_uint(r1.w);
      o = mk3(r0.x, r0.y, r0.z); d = mk3(r1.x, r1.y, r1.z);
      tmax = ANY ? r0.w : 1e28f;
      hit.t = tmax; hit.tri = -1; hit.prim = 0x7fffffff; hit.inst = 0x7fffffff; hit.u = hit.v = 0.f; hit.flags = 0;
      rb = makeRayBox(o, d);
      sp = 0; sb = 0; active = true; age = 0; rayVisits = 0;
      // no triangles, or a direction with NaN / zero length (det can never be != 0): finished at once, as in traverse()
      cur = (A.triCount == 0 || !(fabsf(d.x) + fabsf(d.y) + fabsf(d.z) > 0.0f)) ? EID_TRAV_DONE : A.rootRef;
    }
    if (!__ballot_sync(0xffffffffu, active)) break;
#if EID_TQ_STEAL
    if (!more) {                                       // warp-uniform: this warp fetches no more rays
      if (!finalPhase) {                               // every ray in flight is owned by the lane that fetched it
        finalPhase = true;
        owner = (int)lane; ownDst = ANY ? id : entry;
        mySlot.t = 0.f; mySlot.u = mySlot.v = 0.f; mySlot.tri = -1; mySlot.prim = mySlot.inst = 0x7fffffff; mySlot.outstanding = active ? 1 : 0;
        __syncwarp();
      }
      if (ANY && active && warpSlots[owner].tri >= 0) cur = EID_TRAV_DONE;      // another piece of this occlusion ray already hit something
#pragma unroll 1
      for (int s = 0; s < EID_TQ_STEALS_PER_ROUND; ++s) {
        const unsigned idleM = __ballot_sync(0xffffffffu, !active);
        if (__popc(idleM) < EID_TQ_STEAL_MIN_IDLE) break;
        const int avail = (active && cur != EID_TRAV_DONE && age >= EID_TQ_STEAL_MIN_AGE) ? sp : 0;
        const int best = __reduce_max_sync(0xffffffffu, (avail << 5) | (int)lane);
        if ((best >> 5) < 1) break;
        const int donor = best & 31, thief = __ffs(idleM) - 1;
        int ref = 0;
        if ((int)lane == donor) { ref = stack[sb]; ++sb; --sp; }                // the oldest entry = the largest pending subtree
        ref = __shfl_sync(0xffffffffu, ref, donor);
        const int ow = __shfl_sync(0xffffffffu, owner, donor);
        const float ox = __shfl_sync(0xffffffffu, o.x, donor), oy = __shfl_sync(0xffffffffu, o.y, donor), oz = __shfl_sync(0xffffffffu, o.z, donor);
        const float dx = __shfl_sync(0xffffffffu, d.x, donor), dy = __shfl_sync(0xffffffffu, d.y, donor), dz = __shfl_sync(0xffffffffu, d.z, donor);
        const float tm = __shfl_sync(0xffffffffu, tmax, donor);
        const float ht = __shfl_sync(0xffffffffu, hit.t, donor), hu = __shfl_sync(0xffffffffu, hit.u, donor), hv = __shfl_sync(0xffffffffu, hit.v, donor);
        const int htri = __shfl_sync(0xffffffffu, hit.tri, donor), hprim = __shfl_sync(0xffffffffu, hit.prim, donor), hinst = __shfl_sync(0xffffffffu, hit.inst, donor);
        if ((int)lane == thief) {
          o = mk3(ox, oy, oz); d = mk3(dx, dy, dz); tmax = tm;
          hit.t = ht; hit.u = hu; hit.v = hv; hit.tri = htri; hit.prim = hprim; hit.inst = hinst; hit.flags = 0;   // only better candidates are taken
          rb = makeRayBox(o, d);
          owner = ow; cur = ref; sp = 0; sb = 0; active = true; age = EID_TQ_STEAL_MIN_AGE;   // a piece of a long ray may be split again at once
        }
        if ((int)lane == ow) mySlot.outstanding++;
        __syncwarp();
      }
    }
#endif
#pragma unroll 1
    for (int it = 0; it < EID_TQ_NODE_STEPS; ++it) {
      const bool inner = active && cur >= 0;
      if (!__any_sync(0xffffffffu, inner)) break;
      if (inner) {
        if (STATS) { ++nodeVisits; ++rayVisits; }
        ++age;
        nodeStep<ANY, EID_TQ_SORT != 0>(A, rb, hit.t, cur, stack + sb, sp, EID_TQ_PREFETCH == 2 || (EID_TQ_PREFETCH == 1 && !more));
      }
    }
    if (active && cur < 0) {
      if (cur != EID_TRAV_DONE) {
        if (leafStep<ANY, STATS, false>(A, cur, o, d, tmax, hit, &triTests, low)) cur = EID_TRAV_DONE;
        else cur = (sp ? stack[sb + --sp] : EID_TRAV_DONE);
      }
#if EID_TQ_STEAL
      if (cur == EID_TRAV_DONE && !finalPhase) {
#else
      if (cur == EID_TRAV_DONE) {
#endif
        if (ANY) occl[id] = hit.tri >= 0 ? 1u : 0u;
        else hits[entry] = make_float4(hit.t, hit.u, hit.v, __int_as_float(hit.tri));
        if (STATS) atomicMax(&totals[ANY ? 9 : 8], (unsigned long long)rayVisits);   // longest queued ray since creation (the tail every queue launch waits for)
        active = false;
      }
    }
#if EID_TQ_STEAL
    if (finalPhase) {
      // finished pieces report to the lane that owns their ray; the owner writes once its last piece is in
      bool fin = active && cur == EID_TRAV_DONE;
      if (fin && owner == (int)lane && mySlot.outstanding == 1 && mySlot.tri < 0) {
        // the common case — a ray that was never robbed (and is no thief's piece): written at once, like before the queue drained
        if (ANY) occl[ownDst] = hit.tri >= 0 ? 1u : 0u;
        else hits[ownDst] = make_float4(hit.t, hit.u, hit.v, __int_as_float(hit.tri));
        mySlot.outstanding = 0;
        active = false; fin = false;
      }
      unsigned finM = __ballot_sync(0xffffffffu, fin);
      while (finM) {
        const int src = __ffs(finM) - 1;
        finM &= finM - 1;
        const int ow = __shfl_sync(0xffffffffu, owner, src);
        const float ht = __shfl_sync(0xffffffffu, hit.t, src), hu = __shfl_sync(0xffffffffu, hit.u, src), hv = __shfl_sync(0xffffffffu, hit.v, src);
        const int htri = __shfl_sync(0xffffffffu, hit.tri, src), hprim = __shfl_sync(0xffffffffu, hit.prim, src), hinst = __shfl_sync(0xffffffffu, hit.inst, src);
        if ((int)lane == ow) {
          const bool better = htri >= 0 && (mySlot.tri < 0 || ht < mySlot.t || (ht == mySlot.t && (hinst < mySlot.inst || (hinst == mySlot.inst && hprim < mySlot.prim))));
          if (better) { mySlot.t = ht; mySlot.u = hu; mySlot.v = hv; mySlot.tri = htri; mySlot.prim = hprim; mySlot.inst = hinst; }
          if (--mySlot.outstanding == 0) {
            if (ANY) occl[ownDst] = mySlot.tri >= 0 ? 1u : 0u;
            else hits[ownDst] = make_float4(mySlot.tri >= 0 ? mySlot.t : 1e28f, mySlot.u, mySlot.v, __int_as_float(mySlot.tri));
          }
        }
        __syncwarp();
      }
      if (fin) active = false;
    }
#endif
  }
  if (STATS) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) { nodeVisits += __shfl_xor_sync(0xffffffffu, nodeVisits, s); triTests += __shfl_xor_sync(0xffffffffu, triTests, s); }
    if (lane == 0) { if (nodeVisits) atomicAdd(&counters[3], (unsigned long long)nodeVisits); if (triTests) atomicAdd(&counters[4], (unsigned long long)triTests); }
  }
}
#undef EID_POP

}  // namespace eid

// stage_denoise.cuh — K3 / K4 (denoise_direct.comp / denoise_indirect.comp) as a shared-memory tile kernel fed by TMA.
//
// An A-Trous level with step s = 1 << level only ever combines pixels of one phase (x mod s, y mod s) of the dilated lattice.
// A block therefore owns a 32 x (4 R) tile of lattice points of ONE phase: its 5 x 5 taps are then unit-stride neighbours in
// lattice space, whatever the level, and the tile plus a 2-point halo (36 x (4 R + 4) texels of the three input planes: position +
// material hash, normal, colour) is fetched by three `cp.async.bulk.tensor.5d` (TMA) loads.  The lattice view of a pitch-linear
// image is a 5-D tensor (component, phase x, lattice x, phase y, lattice y) with strides (4, 16, 16 s, 16 P, 16 P s) bytes, built
// once per (buffer, level) by the host (render.cu: tensorMapFor).  Coordinates left of / above the image are zero-filled by the
// TMA unit; texels right of / below the rendered size are real memory of the (padded) allocation.  Both are invalidated by a
// block-uniform fix-up pass that only border tiles run, so the tap loop has no bounds tests and no address arithmetic: every
// tap is three LDS.128 at immediate offsets.
//
// STRICT: the reference's arithmetic, tap order and skips (bit-identical to the oracle) on the raw planes.
// Fast (default): the same weights, w = (exp(-dl/sl) + .01) min(1, exp(-|dn|^2/sn)) (exp(-|dp|^2/sd) + .01) G, evaluated on planes
// that k_denoise_prep pre-scaled by sqrt(log2(e) / sigma) (so each exponent is a plain squared distance fed to MUFU ex2) and with
// |dn|^2 expanded to nn + qq - 2 n.q (|q|^2 precomputed per texel): 25 instead of 43 instructions per tap.
#pragma once
#include <cuda.h>
#include "stages.h"
#include "stage_post.cuh"

namespace eid {

// ---- mbarrier / TMA primitives (inline PTX, sm_90+ forms that sm_100a keeps) ------------------------------------------------------
DEV uint32_t smemAddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
DEV void mbarInit(uint64_t* bar, uint32_t arrivals) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(arrivals) : "memory"); }
DEV void fenceBarrierInit() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
DEV void fenceProxyAsync() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
DEV void mbarArriveExpectTx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes) : "memory");
}
DEV void mbarWait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "EID_MBAR_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra EID_MBAR_DONE;\n"
      "bra EID_MBAR_WAIT;\n"
      "EID_MBAR_DONE:\n"
      "}\n" ::"r"(smemAddr(bar)), "r"(parity) : "memory");
}
DEV void tmaLoad5D(void* smemDst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(smemAddr(smemDst)),
               "l"(map), "r"(smemAddr(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}
DEV void cpAsync16(void* smemDst, const void* gsrc, bool valid) {   // 16-byte LDGSTS, zero-filled when !valid
  const uint32_t n = valid ? 16u : 0u;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smemAddr(smemDst)), "l"(gsrc), "r"(n) : "memory");
}
DEV void cpAsyncWaitAll() { asm volatile("cp.async.wait_all;" ::: "memory"); }


#ifndef EID_DENOISE_PACKED
#define EID_DENOISE_PACKED 0     // 1: fast path on pixel pairs with FADD2 / FFMA2 (measured: see profiles/README.md)
#endif
// ---- Blackwell packed fp32 pairs (FADD2 / FMUL2 / FFMA2: two IEEE fp32 operations per issue slot) ---------------------------------
typedef unsigned long long p2;                         // (lo, hi) in one 64-bit register pair
DEV p2 pk(float lo, float hi) { p2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
DEV p2 pkv(float lo, float hi) { p2 r; asm volatile("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }   // never rematerialised
DEV void upk(p2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
DEV p2 padd(p2 a, p2 b) { p2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
DEV p2 psub(p2 a, p2 b) { p2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
DEV p2 pmul(p2 a, p2 b) { p2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
DEV p2 pfma(p2 a, p2 b, p2 c) { p2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
DEV float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// weights of the fast path on pre-scaled planes; returns w * (hash match)
DEV float fastWeightDirect(float cLum, float4 cN2, float cNN, f3 cPos, uint32_t cHash, const float4& qp, const float4& qn, const float4& qc, float g) {
  float wc, wn, wd;
  const float dl = cLum - qc.w;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(wc) : "f"(-fabsf(dl)));
  const float en = fmaf(cN2.x, qn.x, fmaf(cN2.y, qn.y, fmaf(cN2.z, qn.z, cNN + qn.w)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(wn) : "f"(en));
  const float dx = cPos.x - qp.x, dy = cPos.y - qp.y, dz = cPos.z - qp.z;
  const float ep = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(wd) : "f"(-ep));
  const float w = ((wc + 1e-2f) * wn) * fmaf(wd, g, 1e-2f * g);
  return (__float_as_uint(qp.w) == cHash) ? w : 0.0f;
}
DEV float fastWeightIndirect(f3 cCol, float4 cN2, float cNN, f3 cPos, uint32_t cHash, const float4& qp, const float4& qn, const float4& qc, float g) {
  float wc, wn, wd;
  const float cx = cCol.x - qc.x, cy = cCol.y - qc.y, cz = cCol.z - qc.z;
  const float ec = fmaf(cz, cz, fmaf(cy, cy, fmaf(cx, cx, qc.w)));      // qc.w = 0, or NaN for a non-finite tap colour
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(wc) : "f"(-ec));
  const float en = fmaf(cN2.x, qn.x, fmaf(cN2.y, qn.y, fmaf(cN2.z, qn.z, cNN + qn.w)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(wn) : "f"(en));
  const float dx = cPos.x - qp.x, dy = cPos.y - qp.y, dz = cPos.z - qp.z;
  const float ep = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(wd) : "f"(-ep));
  const float w = ((wc + 1e-2f) * wn) * fmaf(wd, g, 1e-2f * g);
  return (__float_as_uint(qp.w) == cHash) ? w : 0.0f;
}

// `R` pixels per thread (consecutive lattice rows of one column), 4 warps per block: tile = 32 x 4R lattice points.
template <bool INDIRECT, bool STRICT, int R>
__global__ void __launch_bounds__(128) k_atrous_tile(const FrameParams P, const __grid_constant__ CUtensorMap mapPos,
                                                     const __grid_constant__ CUtensorMap mapNrm, const __grid_constant__ CUtensorMap mapCol,
                                                     const AtrousArgs A) {
  constexpr int TR = 4 * R, TH = TR + 4, PW = EID_TILE_PW;
  __shared__ __align__(128) float4 sPos[TH][PW];
  __shared__ __align__(128) float4 sNrm[TH][PW];
  __shared__ __align__(128) float4 sCol[TH][PW];
  __shared__ __align__(8) uint64_t bar;

  const int level = A.level, s = 1 << level;
  const int bw = INDIRECT ? P.st.size.x / 2 : P.st.size.x, bh = INDIRECT ? P.st.size.y / 2 : P.st.size.y;
  // block -> (tile column, x phase), (stripe, tile row, y phase)
  const int px = blockIdx.x & (s - 1), tX = blockIdx.x >> level;
  const int perStripe = A.nTy << level;
  const int ks = blockIdx.y / perStripe, rem = blockIdx.y - ks * perStripe;
  const int py = rem & (s - 1), tYrel = rem >> level;
  const int base = A.first + ks * A.stride;
  const int ylo = max(base, 0), yhi = min(base + A.rows, bh);           // output rows of this stripe
  if (yhi <= ylo) return;
  const int tY = (ylo >> level) / TR + tYrel;
  if (tY > ((yhi - 1) >> level) / TR) return;                           // surplus tile row of the host's upper bound
  const int X0 = tX * EID_TILE_W - 2, Y0 = tY * TR - 2;                 // lattice origin of the tile incl. halo
  const int tid = threadIdx.y * 32 + threadIdx.x;

  // ---- tile load ----
  if (A.useTma) {
    if (tid == 0) { mbarInit(&bar, 1); fenceBarrierInit(); }
    __syncthreads();
    if (tid == 0) {
      mbarArriveExpectTx(&bar, 3u * TH * PW * 16u);
      tmaLoad5D(&sPos[0][0], &mapPos, &bar, 0, px, X0, py, Y0);
      tmaLoad5D(&sNrm[0][0], &mapNrm, &bar, 0, px, X0, py, Y0);
      tmaLoad5D(&sCol[0][0], &mapCol, &bar, 0, px, X0, py, Y0);
    }
    mbarWait(&bar, 0);
  } else {
    for (int t = tid; t < TH * PW; t += 128) {
      const int j = t / PW, i = t - j * PW;
      const int x = px + (X0 + i) * s, y = py + (Y0 + j) * s;
      const bool okG = x >= 0 && y >= 0 && x < A.gPitch && y < A.allocRows, okI = x >= 0 && y >= 0 && x < A.iPitch && y < A.allocRows;
      const size_t gi = okG ? (size_t)y * A.gPitch + x : 0, ii = okI ? (size_t)y * A.iPitch + x : 0;
      cpAsync16(&sPos[j][i], A.gPos + gi, okG);
      cpAsync16(&sNrm[j][i], A.gNrm + gi, okG);
      cpAsync16(&sCol[j][i], A.inImg + ii, okI);
    }
    cpAsyncWaitAll();
    __syncthreads();
  }

  // ---- fix-up (border tiles only): texels outside the rendered image never match a centre; fast path: colour -> weight inputs ----
  const bool interior = px + X0 * s >= 0 && px + (X0 + PW - 1) * s < bw && py + Y0 * s >= 0 && py + (Y0 + TH - 1) * s < bh;
  const float LOG2E = 1.44269504088896341f;
  const float sigL = INDIRECT ? P.st.sigLuminIndirect : P.st.sigLuminDirect;
  const float sigN = INDIRECT ? P.st.sigNormalIndirect : P.st.sigNormalDirect;
  const float sigD = INDIRECT ? P.st.sigDepthIndirect : P.st.sigDepthDirect;
  if (!interior || !STRICT) {
    const float kL = INDIRECT ? sqrtf(LOG2E / sigL) : LOG2E / sigL;
    for (int t = tid; t < TH * PW; t += 128) {
      const int j = t / PW, i = t - j * PW;
      if (!interior) {
        const int x = px + (X0 + i) * s, y = py + (Y0 + j) * s;
        if (x < 0 || y < 0 || x >= bw || y >= bh) {
          sPos[j][i] = make_float4(0.f, 0.f, 0.f, __uint_as_float(EID_INVALID_MAT));
          sNrm[j][i] = make_float4(0.f, 0.f, 0.f, 0.f);
          sCol[j][i] = make_float4(0.f, 0.f, 0.f, 0.f);
          continue;
        }
      }
      if (!STRICT) {
        float4 c = sCol[j][i];
        const bool bad = !(fabsf(c.x) <= 3.0e38f && fabsf(c.y) <= 3.0e38f && fabsf(c.z) <= 3.0e38f);   // inf / NaN
        if (INDIRECT) c = bad ? make_float4(0.f, 0.f, 0.f, __int_as_float(0x7fffffff)) : make_float4(c.x * kL, c.y * kL, c.z * kL, 0.f);
        else c = bad ? make_float4(0.f, 0.f, 0.f, __int_as_float(0x7fffffff)) : make_float4(c.x, c.y, c.z, fmaf(0.0722f, c.z, fmaf(0.7152f, c.y, 0.2126f * c.x)) * kL);
        sCol[j][i] = c;
      }
    }
    __syncthreads();
  }

  // ---- centres ----
  const int lx = threadIdx.x + 2, ly0 = threadIdx.y * R + 2;            // tile coordinates of pixel 0
  const int x = px + (X0 + lx) * s;
  const int y0 = py + (Y0 + ly0) * s;
  bool inside[R];
  uint32_t hash[R];
  f3 sum[R];
  float sumW[R];
#pragma unroll
  for (int k = 0; k < R; ++k) {
    const int y = y0 + k * s;
    inside[k] = x < bw && y >= ylo && y < yhi;
    hash[k] = __float_as_uint(sPos[ly0 + k][lx].w);
    sum[k] = mk3(0.0f); sumW[k] = 0.0f;
  }

  if (STRICT) {
    f3 pos[R], norm[R], color[R];
    float lumC[R];
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const float4 cp = sPos[ly0 + k][lx], cn = sNrm[ly0 + k][lx], c4 = sCol[ly0 + k][lx];
      pos[k] = mk3(cp.x, cp.y, cp.z); norm[k] = mk3(cn.x, cn.y, cn.z); color[k] = mk3(c4.x, c4.y, c4.z);
      lumC[k] = lum3(color[k]);
    }
#pragma unroll
    for (int rr = 0; rr < R + 4; ++rr) {                  // tap row rr serves pixel k as j = rr - 2 - k (reference order: j outer, i inner)
#pragma unroll
      for (int i = -2; i <= 2; i++) {
        const float4 qp = sPos[ly0 - 2 + rr][lx + i];
        const uint32_t hq = __float_as_uint(qp.w);
        bool any = false;
#pragma unroll
        for (int k = 0; k < R; ++k)
          if (rr - 2 - k >= -2 && rr - 2 - k <= 2) any = any || (hash[k] == hq);
        if (!any || hq == EID_INVALID_MAT) continue;
        const float4 qn = sNrm[ly0 - 2 + rr][lx + i], q4 = sCol[ly0 - 2 + rr][lx + i];
        const f3 cq = mk3(q4.x, q4.y, q4.z);
#pragma unroll
        for (int k = 0; k < R; ++k) {
          const int j = rr - 2 - k;
          if (j < -2 || j > 2) continue;
          if (hash[k] != hq) continue;
          const float w = tapWeight<INDIRECT, true>(color[k], lumC[k], norm[k], pos[k], qp, qn, cq, sigL, sigN, sigD, 0.f, 0.f, 0.f, c_gauss5x5[(i + 2) * 5 + (j + 2)]);
          sum[k] = sum[k] + cq * w;
          sumW[k] = __fadd_rn(sumW[k], w);
        }
      }
    }
#if EID_DENOISE_PACKED
  } else {
    // Fast path on pixel PAIRS (k, k + 1): every tap that serves both pixels of a pair is evaluated once with packed fp32 (the tap's
    // value is the broadcast scalar operand of FADD2 / FFMA2), so the 19 fp32 operations of a tap weight cost 10 issue slots per pixel;
    // the three exponentials stay scalar MUFU ex2.  A tap row that serves only one pixel of the pair runs the same code with the
    // other half's Gaussian weight = 0.
    constexpr float G[25] = {.0030f, .0133f, .0219f, .0133f, .0030f, .0133f, .0596f, .0983f, .0596f, .0133f, .0219f, .0983f, .1621f,
                             .0983f, .0219f, .0133f, .0596f, .0983f, .0596f, .0133f, .0030f, .0133f, .0219f, .0133f, .0030f};
    static_assert(R % 2 == 0, "the fast path works on pixel pairs");
    constexpr int NP = R / 2;
    p2 cPx[NP], cPy[NP], cPz[NP], cNx[NP], cNy[NP], cNz[NP], cNN[NP], cC0[NP], cC1[NP], cC2[NP];   // centres; cC0 = lum (direct) or colour.x
    p2 sx[NP], sy[NP], sz[NP], sw[NP];
#pragma unroll
    for (int pr = 0; pr < NP; ++pr) {
      const float4 p0 = sPos[ly0 + 2 * pr][lx], p1 = sPos[ly0 + 2 * pr + 1][lx];
      const float4 n0 = sNrm[ly0 + 2 * pr][lx], n1 = sNrm[ly0 + 2 * pr + 1][lx];
      const float4 c0 = sCol[ly0 + 2 * pr][lx], c1 = sCol[ly0 + 2 * pr + 1][lx];
      cPx[pr] = pkv(p0.x, p1.x); cPy[pr] = pkv(p0.y, p1.y); cPz[pr] = pkv(p0.z, p1.z);
      cNx[pr] = pkv(2.0f * n0.x, 2.0f * n1.x); cNy[pr] = pkv(2.0f * n0.y, 2.0f * n1.y); cNz[pr] = pkv(2.0f * n0.z, 2.0f * n1.z);
      cNN[pr] = pkv(n0.w, n1.w);
      if (INDIRECT) { cC0[pr] = pkv(c0.x, c1.x); cC1[pr] = pkv(c0.y, c1.y); cC2[pr] = pkv(c0.z, c1.z); }
      else { cC0[pr] = pkv(c0.w, c1.w); cC1[pr] = cC2[pr] = 0; }
      sx[pr] = sy[pr] = sz[pr] = sw[pr] = pk(0.f, 0.f);
    }
    const p2 c001 = pk(1e-2f, 1e-2f);
#pragma unroll
    for (int rr = 0; rr < R + 4; ++rr) {
#pragma unroll
      for (int i = -2; i <= 2; i++) {
        const float4 qp = sPos[ly0 - 2 + rr][lx + i], qn = sNrm[ly0 - 2 + rr][lx + i], qc = sCol[ly0 - 2 + rr][lx + i];
        const uint32_t hq = __float_as_uint(qp.w);
#pragma unroll
        for (int pr = 0; pr < NP; ++pr) {
          const int j0 = rr - 2 - 2 * pr, j1 = j0 - 1;
          const bool v0 = j0 >= -2 && j0 <= 2, v1 = j1 >= -2 && j1 <= 2;
          if (!v0 && !v1) continue;
          if (v0 != v1) {                                  // the tap serves one pixel of the pair: scalar weight on that half
            const int k = v0 ? 2 * pr : 2 * pr + 1, j = v0 ? j0 : j1;
            float a0, a1, ax, ay, az, aw, b0, b1;
            f3 cp, cc; float4 cn2; float cnn, cl;
            upk(cPx[pr], a0, a1); cp.x = v0 ? a0 : a1; upk(cPy[pr], a0, a1); cp.y = v0 ? a0 : a1; upk(cPz[pr], a0, a1); cp.z = v0 ? a0 : a1;
            upk(cNx[pr], a0, a1); ax = v0 ? a0 : a1; upk(cNy[pr], a0, a1); ay = v0 ? a0 : a1; upk(cNz[pr], a0, a1); az = v0 ? a0 : a1;
            upk(cNN[pr], a0, a1); cnn = v0 ? a0 : a1; aw = 0.f;
            cn2 = make_float4(ax, ay, az, aw);
            upk(cC0[pr], a0, a1); cl = v0 ? a0 : a1; cc.x = cl;
            upk(cC1[pr], a0, a1); cc.y = v0 ? a0 : a1; upk(cC2[pr], a0, a1); cc.z = v0 ? a0 : a1;
            const float g = G[(i + 2) * 5 + (j + 2)];
            const float w = INDIRECT ? fastWeightIndirect(cc, cn2, cnn, cp, hash[k], qp, qn, qc, g) : fastWeightDirect(cl, cn2, cnn, cp, hash[k], qp, qn, qc, g);
            upk(sx[pr], b0, b1); if (v0) b0 = fmaf(qc.x, w, b0); else b1 = fmaf(qc.x, w, b1); sx[pr] = pk(b0, b1);
            upk(sy[pr], b0, b1); if (v0) b0 = fmaf(qc.y, w, b0); else b1 = fmaf(qc.y, w, b1); sy[pr] = pk(b0, b1);
            upk(sz[pr], b0, b1); if (v0) b0 = fmaf(qc.z, w, b0); else b1 = fmaf(qc.z, w, b1); sz[pr] = pk(b0, b1);
            upk(sw[pr], b0, b1); if (v0) b0 += w; else b1 += w; sw[pr] = pk(b0, b1);
            continue;
          }
          const float g0 = G[(i + 2) * 5 + (j0 + 2)], g1 = G[(i + 2) * 5 + (j1 + 2)];
          // colour term
          float wc0, wc1;
          if (INDIRECT) {
            const p2 dx = psub(cC0[pr], pk(qc.x, qc.x)), dy = psub(cC1[pr], pk(qc.y, qc.y)), dz = psub(cC2[pr], pk(qc.z, qc.z));
            float e0, e1;
            upk(pfma(dz, dz, pfma(dy, dy, pfma(dx, dx, pk(qc.w, qc.w)))), e0, e1);   // qc.w = 0, or NaN for a non-finite tap colour
            wc0 = ex2f(-e0); wc1 = ex2f(-e1);
          } else {
            float d0, d1;
            upk(psub(cC0[pr], pk(qc.w, qc.w)), d0, d1);
            wc0 = ex2f(-fabsf(d0)); wc1 = ex2f(-fabsf(d1));
          }
          // normal term: -(|n|^2 + |q|^2 - 2 n.q) on the pre-scaled normals
          float en0, en1;
          upk(pfma(cNx[pr], pk(qn.x, qn.x), pfma(cNy[pr], pk(qn.y, qn.y), pfma(cNz[pr], pk(qn.z, qn.z), padd(cNN[pr], pk(qn.w, qn.w))))), en0, en1);
          const float wn0 = ex2f(en0), wn1 = ex2f(en1);
          // depth term
          const p2 px_ = psub(cPx[pr], pk(qp.x, qp.x)), py_ = psub(cPy[pr], pk(qp.y, qp.y)), pz_ = psub(cPz[pr], pk(qp.z, qp.z));
          float ep0, ep1;
          upk(pfma(pz_, pz_, pfma(py_, py_, pmul(px_, px_))), ep0, ep1);
          const float wd0 = ex2f(-ep0), wd1 = ex2f(-ep1);
          float w0, w1;
          upk(pmul(pmul(padd(pk(wc0, wc1), c001), pk(wn0, wn1)), padd(pk(wd0, wd1), c001)), w0, w1);
          w0 = (hq == hash[2 * pr]) ? w0 * g0 : 0.0f;
          w1 = (hq == hash[2 * pr + 1]) ? w1 * g1 : 0.0f;
          const p2 w = pk(w0, w1);
          sx[pr] = pfma(w, pk(qc.x, qc.x), sx[pr]); sy[pr] = pfma(w, pk(qc.y, qc.y), sy[pr]); sz[pr] = pfma(w, pk(qc.z, qc.z), sz[pr]);
          sw[pr] = padd(sw[pr], w);
        }
      }
    }
    const float inv = INDIRECT ? 1.0f / sqrtf(LOG2E / sigL) : 1.0f;   // the indirect colour plane was scaled by sqrt(log2e / sigL) for the distance
#pragma unroll
    for (int pr = 0; pr < NP; ++pr) {
      float a0, a1, b0, b1, c0, c1, d0, d1;
      upk(sx[pr], a0, a1); upk(sy[pr], b0, b1); upk(sz[pr], c0, c1); upk(sw[pr], d0, d1);
      sum[2 * pr] = mk3(a0 * inv, b0 * inv, c0 * inv); sum[2 * pr + 1] = mk3(a1 * inv, b1 * inv, c1 * inv);
      sumW[2 * pr] = d0; sumW[2 * pr + 1] = d1;
    }
  }

#else
  } else {
    constexpr float G[25] = {.0030f, .0133f, .0219f, .0133f, .0030f, .0133f, .0596f, .0983f, .0596f, .0133f, .0219f, .0983f, .1621f,
                             .0983f, .0219f, .0133f, .0596f, .0983f, .0596f, .0133f, .0030f, .0133f, .0219f, .0133f, .0030f};
    f3 cPos[R], cCol[R];
    float4 cN2[R];
    float cNN[R], cLum[R];
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const float4 cp = sPos[ly0 + k][lx], cn = sNrm[ly0 + k][lx], c4 = sCol[ly0 + k][lx];
      cPos[k] = mk3(cp.x, cp.y, cp.z);
      cN2[k] = make_float4(2.0f * cn.x, 2.0f * cn.y, 2.0f * cn.z, 0.f); cNN[k] = cn.w;
      cCol[k] = mk3(c4.x, c4.y, c4.z); cLum[k] = c4.w;
    }
#pragma unroll
    for (int rr = 0; rr < R + 4; ++rr) {
#pragma unroll
      for (int i = -2; i <= 2; i++) {
        const float4 qp = sPos[ly0 - 2 + rr][lx + i], qn = sNrm[ly0 - 2 + rr][lx + i], qc = sCol[ly0 - 2 + rr][lx + i];
#pragma unroll
        for (int k = 0; k < R; ++k) {
          const int j = rr - 2 - k;
          if (j < -2 || j > 2) continue;
          const float g = G[(i + 2) * 5 + (j + 2)];
          const float w = INDIRECT ? fastWeightIndirect(cCol[k], cN2[k], cNN[k], cPos[k], hash[k], qp, qn, qc, g)
                                   : fastWeightDirect(cLum[k], cN2[k], cNN[k], cPos[k], hash[k], qp, qn, qc, g);
          sum[k] = mk3(fmaf(qc.x, w, sum[k].x), fmaf(qc.y, w, sum[k].y), fmaf(qc.z, w, sum[k].z));
          sumW[k] += w;
        }
      }
    }
    if (INDIRECT) {                                       // the colour plane was scaled by sqrt(log2e / sigL) for the distance
      const float inv = 1.0f / sqrtf(LOG2E / sigL);
#pragma unroll
      for (int k = 0; k < R; ++k) sum[k] = sum[k] * inv;
    }
  }

#endif
#pragma unroll
  for (int k = 0; k < R; ++k) {
    if (!inside[k]) continue;
    f3 res = mk3(0.0f);
    if (hash[k] != EID_INVALID_MAT) {                     // waveletFilter (denoise_direct.comp:19-71 / denoise_indirect.comp:23-75)
      if (STRICT) res = (sumW[k] < 1e-5f) ? mk3(0.0f) : sum[k] / sumW[k];
      else { const float inv = __fdividef(1.0f, sumW[k]); res = (sumW[k] < 1e-5f) ? mk3(0.0f) : mk3(sum[k].x * inv, sum[k].y * inv, sum[k].z * inv); }
      if (nan3(res) || res.x < 0 || res.y < 0 || res.z < 0 || res.x > 1e8f || res.y > 1e8f || res.z > 1e8f) res = mk3(0.0f);
    }
    if (level == A.lastLevel) {                           // denoise_direct.comp:168 / denoise_indirect.comp:169
      if (STRICT) res = ldrToHdr(res);
      else res = mk3(__fdividef(res.x, 1.01f - res.x), __fdividef(res.y, 1.01f - res.y), __fdividef(res.z, 1.01f - res.z));
    }
    A.outImg[(size_t)(y0 + k * s) * A.iPitch + x] = make_float4(res.x, res.y, res.z, 1.0f);
  }
}

}  // namespace eid

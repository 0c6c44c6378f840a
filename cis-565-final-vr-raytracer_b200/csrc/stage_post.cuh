// stage_post.cuh — K3 / K4 (A-Trous denoisers), K5 (compose.comp) and the display pass (post.frag, mip chain).
#pragma once
#include "frame.cuh"

namespace eid {

// =================================================================================================
// K3 / K4 — denoise_direct.comp / denoise_indirect.comp (edge-avoiding A-Trous, one level per launch)
// =================================================================================================
__constant__ float c_gauss5x5[25] = {.0030f, .0133f, .0219f, .0133f, .0030f, .0133f, .0596f, .0983f, .0596f, .0133f, .0219f, .0983f, .1621f,
                                     .0983f, .0219f, .0133f, .0596f, .0983f, .0596f, .0133f, .0030f, .0133f, .0219f, .0133f, .0030f};

// loadThisGeometry (denoise_common.glsl:42-47) evaluates, for every one of the 25 taps of every pass, the octahedral normal
// decode and a camera-ray spawn (two 4x4 products, a normalize) — ~250 instructions that depend only on the G-buffer texel.
// k_denoise_prep evaluates it ONCE per texel per frame with the identical arithmetic and stores the result in two float4
// planes (pos.xyz + material hash bits, normal.xyz); the nine filter passes then only load.  The indirect passes use their
// own quarter-res planes because the reference spawns that ray with full-res coordinates against the half-res image size
// (uv runs to ~2 — reference quirk, kept).
DEV void thisGeometry(const FrameParams& P, int gx, int gy, int sw, int sh, float4& posHash, float4& nrm) {
  const uint4 g = loadG(P.thisG, P, gx, gy);
  const f3 n = octDecode(g.y);
  f3 o, d;
  raySpawn<false>(P.cam, gx, gy, sw, sh, o, d);
  const f3 pos = o + d * __uint_as_float(g.x);
  posHash = make_float4(pos.x, pos.y, pos.z, __uint_as_float(g.w & 0xFF000000u));
  nrm = make_float4(n.x, n.y, n.z, 0.f);
}

// FAST (tile kernel, default numerics): the planes are pre-scaled for k_atrous_tile's fast path — position * sqrt(log2e / sigDepth),
// normal * sqrt(log2e / sigNormal) with w = -|scaled normal|^2 — so that every exponent of a tap weight is a plain squared distance.
template <bool FAST, bool FASTQ>
__global__ void __launch_bounds__(256) k_denoise_prep(const FrameParams P, int first, int stride, int rows) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = stripeRow(first, stride, rows, 8);
  const int W = P.st.size.x, H = P.st.size.y, Wi = W / 2, Hi = H / 2;
  if (x >= W || y >= H || y < 0) return;
  const float LOG2E = 1.44269504088896341f;
  float4 a, b;
  thisGeometry(P, x, y, W, H, a, b);
  if (FAST) {
    const float sd = sqrtf(LOG2E / P.st.sigDepthDirect), sn = sqrtf(LOG2E / P.st.sigNormalDirect);
    a = make_float4(a.x * sd, a.y * sd, a.z * sd, a.w);
    b = make_float4(b.x * sn, b.y * sn, b.z * sn, 0.f);
    b.w = -fmaf(b.z, b.z, fmaf(b.y, b.y, b.x * b.x));
  }
  const size_t pix = (size_t)y * P.pitch + x;
  P.geomPos[pix] = a; P.geomNrm[pix] = b;
  if (!(x & 1) && !(y & 1) && (x >> 1) < Wi && (y >> 1) < Hi) {
    thisGeometry(P, x, y, Wi, Hi, a, b);
    if (FASTQ) {
      const float sd = sqrtf(LOG2E / P.st.sigDepthIndirect), sn = sqrtf(LOG2E / P.st.sigNormalIndirect);
      a = make_float4(a.x * sd, a.y * sd, a.z * sd, a.w);
      b = make_float4(b.x * sn, b.y * sn, b.z * sn, 0.f);
      b.w = -fmaf(b.z, b.z, fmaf(b.y, b.y, b.x * b.x));
    }
    const size_t hp = (size_t)(y >> 1) * (P.pitch / 2) + (x >> 1);
    P.geomPosH[hp] = a; P.geomNrmH[hp] = b;
  }
}

// exp of the three edge-stopping weights.  STRICT: the bit-reproducible polynomial shared with the oracle (parity runs).
// Fast (default): one MUFU ex2 on a pre-scaled exponent — relative error ~2^-21, far inside the 1e-3 radiance tolerance.
template <bool STRICT> DEV float edgeExp(float num, float sigma, float negLog2eOverSigma) {
  if (STRICT) return eid_expf(__fdiv_rn(-num, sigma));
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(num * negLog2eOverSigma));   // exponent <= 0: no range fix-up needed
  return y;
}

// weight of one tap (denoise_direct.comp:40-62 / denoise_indirect.comp:44-66)
template <bool INDIRECT, bool STRICT>
DEV float tapWeight(const f3& color, float lumC, const f3& norm, const f3& pos, const float4& qp, const float4& qn, const f3& cq,
                    float sigL, float sigN, float sigD, float nL, float nN, float nD, float gauss) {
  if (STRICT) {
    float distColor;
    if (INDIRECT) { f3 dc = color - cq; distColor = dot3(dc, dc); }
    else distColor = fabsf(__fsub_rn(lumC, lum3(cq)));
    const float wColor = __fadd_rn(edgeExp<true>(distColor, sigL, nL), 1e-2f);
    const f3 dn = norm - mk3(qn.x, qn.y, qn.z);
    const float wNorm = gmin(1.0f, edgeExp<true>(dot3(dn, dn), sigN, nN));
    const f3 dp = pos - mk3(qp.x, qp.y, qp.z);
    const float wDepth = __fadd_rn(edgeExp<true>(dot3(dp, dp), sigD, nD), 1e-2f);
    return __fmul_rn(__fmul_rn(__fmul_rn(wColor, wNorm), wDepth), gauss);
  } else {
    // fast path (default): same formula with fused multiply-adds; deviates from the strict path by ~1e-6 relative
    float distColor;
    if (INDIRECT) { const float dx = color.x - cq.x, dy = color.y - cq.y, dz = color.z - cq.z; distColor = fmaf(dz, dz, fmaf(dy, dy, dx * dx)); }
    else distColor = fabsf(lumC - fmaf(0.0722f, cq.z, fmaf(0.7152f, cq.y, 0.2126f * cq.x)));
    const float wColor = edgeExp<false>(distColor, sigL, nL) + 1e-2f;
    const float nx = norm.x - qn.x, ny = norm.y - qn.y, nz = norm.z - qn.z;
    const float wNorm = edgeExp<false>(fmaf(nz, nz, fmaf(ny, ny, nx * nx)), sigN, nN);   // <= 1 by construction: min(1, .) is the identity
    const float px = pos.x - qp.x, py = pos.y - qp.y, pz = pos.z - qp.z;
    const float wDepth = edgeExp<false>(fmaf(pz, pz, fmaf(py, py, px * px)), sigD, nD) + 1e-2f;
    return (wColor * wNorm) * (wDepth * gauss);
  }
}

// One A-Trous level.  A thread filters R pixels of one column that are `step` rows apart (the same phase of the dilated
// lattice), so the 5 tap rows of neighbouring pixels overlap: R+4 tap rows are loaded for R pixels instead of 5R, every load
// still a fully coalesced 16-B access along x.  Each pixel receives its taps in the reference's j-major / i-minor order, so
// the sums are bit-identical for every R.  Virtual row v of a stripe of `rows` rows: phase p = v % step, chunk c = v / step
// -> stripe rows p + (R c + k) step, k < R.
// CHECK = false is the interior variant (block-uniform choice): every tap of every pixel of the block is inside the image,
// so no bounds tests are emitted.  The fast path accumulates branch-free (mismatching taps get weight 0).
template <bool INDIRECT, bool STRICT, int R, bool CHECK>
DEV void atrousBody(const FrameParams& P, const float4* __restrict__ inImg, float4* __restrict__ outImg, int level, int lastLevel, int x, int y0,
                    int lr0, int rows, int bw, int bh) {
  const int step = 1 << level;
  const float sigL = INDIRECT ? P.st.sigLuminIndirect : P.st.sigLuminDirect;
  const float sigN = INDIRECT ? P.st.sigNormalIndirect : P.st.sigNormalDirect;
  const float sigD = INDIRECT ? P.st.sigDepthIndirect : P.st.sigDepthDirect;
  const float LOG2E = 1.44269504088896341f;
  const float nL = -LOG2E / sigL, nN = -LOG2E / sigN, nD = -LOG2E / sigD;   // fast path: exp(-d/sigma) = exp2(d * nX)
  const float4* __restrict__ gPos = INDIRECT ? P.geomPosH : P.geomPos;
  const float4* __restrict__ gNrm = INDIRECT ? P.geomNrmH : P.geomNrm;
  const unsigned gp = INDIRECT ? P.pitch / 2 : P.pitch, ip = P.pitch;

  f3 pos[R], norm[R], color[R], sum[R];
  float lumC[R], sumW[R];
  uint32_t hash[R];
  bool inside[R];
#pragma unroll
  for (int k = 0; k < R; ++k) {
    const int y = y0 + k * step;
    inside[k] = !CHECK || ((lr0 + k * step < rows) && y >= 0 && y < bh);
    hash[k] = EID_INVALID_MAT;
    sum[k] = mk3(0.0f); sumW[k] = 0.0f;
    pos[k] = norm[k] = color[k] = mk3(0.0f); lumC[k] = 0.0f;
    if (inside[k]) {
      const float4 cp = __ldg(gPos + ((unsigned)y * gp + (unsigned)x));
      hash[k] = __float_as_uint(cp.w);
      if (!STRICT || hash[k] != EID_INVALID_MAT) {
        const float4 cn = __ldg(gNrm + ((unsigned)y * gp + (unsigned)x));
        const float4 c4 = inImg[(unsigned)y * ip + (unsigned)x];
        pos[k] = mk3(cp.x, cp.y, cp.z); norm[k] = mk3(cn.x, cn.y, cn.z); color[k] = mk3(c4.x, c4.y, c4.z);
        lumC[k] = lum3(color[k]);
      }
    }
  }
#pragma unroll
  for (int rr = 0; rr < R + 4; ++rr) {                    // tap row rr serves pixel k as j = rr - 2 - k
    const int qy = y0 + (rr - 2) * step;
    if (CHECK && (qy >= bh || qy < 0)) continue;
#pragma unroll
    for (int i = -2; i <= 2; i++) {
      const int qx = x + i * step;
      if (CHECK && (qx >= bw || qx < 0)) continue;
      const unsigned gi = (unsigned)qy * gp + (unsigned)qx, ii = (unsigned)qy * ip + (unsigned)qx;
      const float4 qp = __ldg(gPos + gi);
      const uint32_t hq = __float_as_uint(qp.w);
      if (STRICT) {
        bool any = false;
#pragma unroll
        for (int k = 0; k < R; ++k)
          if (rr - 2 - k >= -2 && rr - 2 - k <= 2) any = any || (hash[k] == hq);
        if (!any || hq == EID_INVALID_MAT) continue;
      }
      const float4 qn = __ldg(gNrm + gi);
      const float4 q4 = inImg[ii];
      const f3 cq = mk3(q4.x, q4.y, q4.z);
#pragma unroll
      for (int k = 0; k < R; ++k) {
        const int j = rr - 2 - k;
        if (j < -2 || j > 2) continue;
        if (STRICT) {
          if (hash[k] != hq) continue;
          const float w = tapWeight<INDIRECT, true>(color[k], lumC[k], norm[k], pos[k], qp, qn, cq, sigL, sigN, sigD, nL, nN, nD,
                                                    c_gauss5x5[(i + 2) * 5 + (j + 2)]);
          sum[k] = sum[k] + cq * w;
          sumW[k] = __fadd_rn(sumW[k], w);
        } else {
          float w = tapWeight<INDIRECT, false>(color[k], lumC[k], norm[k], pos[k], qp, qn, cq, sigL, sigN, sigD, nL, nN, nD,
                                               c_gauss5x5[(i + 2) * 5 + (j + 2)]);
          w = (hash[k] == hq) ? w : 0.0f;                 // (an invalid centre is zeroed below, whatever it accumulated)
          sum[k] = mk3(fmaf(cq.x, w, sum[k].x), fmaf(cq.y, w, sum[k].y), fmaf(cq.z, w, sum[k].z));
          sumW[k] += w;
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < R; ++k) {
    if (!inside[k]) continue;
    f3 res = mk3(0.0f);
    if (hash[k] != EID_INVALID_MAT) {                      // waveletFilter (denoise_direct.comp:19-71 / denoise_indirect.comp:23-75)
      res = (sumW[k] < 1e-5f) ? mk3(0.0f) : sum[k] / sumW[k];
      if (nan3(res) || res.x < 0 || res.y < 0 || res.z < 0 || res.x > 1e8f || res.y > 1e8f || res.z > 1e8f) res = mk3(0.0f);
    }
    if (level == lastLevel) res = ldrToHdr(res);           // denoise_direct.comp:168 / denoise_indirect.comp:169
    outImg[(unsigned)(y0 + k * step) * ip + (unsigned)x] = make_float4(res.x, res.y, res.z, 1.0f);
  }
}

template <bool INDIRECT, bool STRICT, int R>
__global__ void __launch_bounds__(128) k_denoise(const FrameParams P, const float4* __restrict__ inImg, float4* __restrict__ outImg, int level,
                                                 int lastLevel, int first, int stride, int rows) {
  const int step = 1 << level;
  const int vrows = step * ((((rows + step - 1) >> level) + R - 1) / R);
  const int bps = (vrows + 3) / 4;
  const int ks = blockIdx.y / bps, v0 = (blockIdx.y - ks * bps) * 4;
  const int bw = INDIRECT ? P.st.size.x / 2 : P.st.size.x, bh = INDIRECT ? P.st.size.y / 2 : P.st.size.y;
  const int x0 = blockIdx.x * 32, base = first + ks * stride;
  // interior test over the whole block (4 virtual rows v0..v0+3, 32 columns): block-uniform
  bool interior = x0 - 2 * step >= 0 && x0 + 31 + 2 * step < bw && v0 + 3 < vrows;
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int v = v0 + t, lr = (v & (step - 1)) + ((v >> level) * R) * step;
    interior = interior && lr + (R - 1) * step < rows && base + lr - 2 * step >= 0 && base + lr + (R + 1) * step < bh;
  }
  const int x = x0 + threadIdx.x, v = v0 + threadIdx.y;
  const int lr0 = (v & (step - 1)) + ((v >> level) * R) * step;     // row of pixel 0 inside the stripe
  if (interior) atrousBody<INDIRECT, STRICT, R, false>(P, inImg, outImg, level, lastLevel, x, base + lr0, lr0, rows, bw, bh);
  else if (x < bw && v < vrows) atrousBody<INDIRECT, STRICT, R, true>(P, inImg, outImg, level, lastLevel, x, base + lr0, lr0, rows, bw, bh);
}

// =================================================================================================
// DENOISER_DIRECT_BILATERAL / DENOISER_INDIRECT_BILATERAL (host_device.h:28-29; EID_VARIANT_*): ONE cross-bilateral pass instead of the
// A-Trous levels — denoise_direct.comp:73-137 (9 x 9, spatial term exp(-(i^2 + j^2) / 10) + .01) and denoise_indirect.comp:77-130
// (11 x 11, no spatial term).  Colour distance is the squared RGB difference in both.  Reads the RAW geometry planes of k_denoise_prep.
// STRICT: the reference's arithmetic and tap order; otherwise MUFU ex2 on pre-scaled exponents and fused multiply-adds.
// =================================================================================================
template <bool INDIRECT, bool STRICT>
__global__ void __launch_bounds__(128) k_bilateral(const FrameParams P, const float4* __restrict__ inImg, float4* __restrict__ outImg, int first, int stride, int rows) {
  constexpr int RADIUS = INDIRECT ? 5 : 4;
  const int x = blockIdx.x * 32 + threadIdx.x;
  const int y = stripeRow(first, stride, rows, 4);
  const int bw = INDIRECT ? P.st.size.x / 2 : P.st.size.x, bh = INDIRECT ? P.st.size.y / 2 : P.st.size.y;
  if (x >= bw || y >= bh || y < 0) return;
  const float sigL = INDIRECT ? P.st.sigLuminIndirect : P.st.sigLuminDirect;
  const float sigN = INDIRECT ? P.st.sigNormalIndirect : P.st.sigNormalDirect;
  const float sigD = INDIRECT ? P.st.sigDepthIndirect : P.st.sigDepthDirect;
  const float LOG2E = 1.44269504088896341f;
  const float nL = -LOG2E / sigL, nN = -LOG2E / sigN, nD = -LOG2E / sigD, nS = -LOG2E / 10.0f;
  const float4* __restrict__ gPos = INDIRECT ? P.geomPosH : P.geomPos;
  const float4* __restrict__ gNrm = INDIRECT ? P.geomNrmH : P.geomNrm;
  const unsigned gp = INDIRECT ? P.pitch / 2 : P.pitch, ip = P.pitch;
  const float4 cp = __ldg(gPos + ((unsigned)y * gp + (unsigned)x)), cn = __ldg(gNrm + ((unsigned)y * gp + (unsigned)x));
  const float4 c4 = inImg[(unsigned)y * ip + (unsigned)x];
  const uint32_t hash = __float_as_uint(cp.w);
  const f3 pos = mk3(cp.x, cp.y, cp.z), norm = mk3(cn.x, cn.y, cn.z), color = mk3(c4.x, c4.y, c4.z);
  f3 sum = mk3(0.0f);
  float sumW = 0.0f;
  if (!(INDIRECT && hash == EID_INVALID_MAT)) {           // the indirect variant returns 0 for sky pixels at once; the direct one runs (and finds no matching tap)
    for (int j = -RADIUS; j <= RADIUS; j++) {
      const int qy = y + j;
      if (qy >= bh || qy < 0) continue;
      for (int i = -RADIUS; i <= RADIUS; i++) {
        const int qx = x + i;
        if (qx >= bw || qx < 0) continue;
        const float4 qp = __ldg(gPos + ((unsigned)qy * gp + (unsigned)qx));
        const uint32_t hq = __float_as_uint(qp.w);
        if (hash != hq || hq == EID_INVALID_MAT) continue;
        const float4 qn = __ldg(gNrm + ((unsigned)qy * gp + (unsigned)qx)), q4 = inImg[(unsigned)qy * ip + (unsigned)qx];
        const f3 cq = mk3(q4.x, q4.y, q4.z);
        float w;
        if (STRICT) {
          const f3 dc = color - cq;
          const float wColor = __fadd_rn(eid_expf(__fdiv_rn(-dot3(dc, dc), sigL)), 1e-2f);
          const f3 dn = norm - mk3(qn.x, qn.y, qn.z);
          const float wNorm = gmin(1.0f, eid_expf(__fdiv_rn(-dot3(dn, dn), sigN)));
          const f3 dp = pos - mk3(qp.x, qp.y, qp.z);
          const float wDepth = __fadd_rn(eid_expf(__fdiv_rn(-dot3(dp, dp), sigD)), 1e-2f);
          w = __fmul_rn(__fmul_rn(wColor, wNorm), wDepth);
          if (!INDIRECT) {
            const float dist2 = (float)(i * i + j * j);
            w = __fmul_rn(w, __fadd_rn(eid_expf(__fdiv_rn(-dist2, 10.0f)), 1e-2f));
          }
          sum = sum + cq * w;
          sumW = __fadd_rn(sumW, w);
        } else {
          const float dx = color.x - cq.x, dy = color.y - cq.y, dz = color.z - cq.z;
          const float nx = norm.x - qn.x, ny = norm.y - qn.y, nz = norm.z - qn.z;
          const float px = pos.x - qp.x, py = pos.y - qp.y, pz = pos.z - qp.z;
          const float wColor = edgeExp<false>(fmaf(dz, dz, fmaf(dy, dy, dx * dx)), sigL, nL) + 1e-2f;
          const float wNorm = edgeExp<false>(fmaf(nz, nz, fmaf(ny, ny, nx * nx)), sigN, nN);
          const float wDepth = edgeExp<false>(fmaf(pz, pz, fmaf(py, py, px * px)), sigD, nD) + 1e-2f;
          w = (wColor * wNorm) * wDepth;
          if (!INDIRECT) w *= edgeExp<false>((float)(i * i + j * j), 10.0f, nS) + 1e-2f;
          sum = mk3(fmaf(cq.x, w, sum.x), fmaf(cq.y, w, sum.y), fmaf(cq.z, w, sum.z));
          sumW += w;
        }
      }
    }
  }
  f3 res = (sumW < 1e-5f) ? mk3(0.0f) : sum / sumW;
  if (nan3(res) || res.x < 0 || res.y < 0 || res.z < 0 || res.x > 1e8f || res.y > 1e8f || res.z > 1e8f) res = mk3(0.0f);
  res = ldrToHdr(res);                                    // denoise_direct.comp:168 / denoise_indirect.comp:169
  outImg[(unsigned)y * ip + (unsigned)x] = make_float4(res.x, res.y, res.z, 1.0f);
}

// =================================================================================================
// K5 — compose.comp:23-42
// =================================================================================================
__global__ void __launch_bounds__(256) k_compose(const FrameParams P, const float4* __restrict__ indSrc, int first, int stride, int rows) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = stripeRow(first, stride, rows, 8);
  if (x >= P.st.size.x || y >= P.st.size.y || y < 0) return;
  const size_t pix = (size_t)y * P.pitch + x;
  const float4 ind = loadImg(indSrc, P, x / 2, y / 2);
  if (P.st.modulate == 0) {
    P.indirectImg[pix] = ind;
  } else {
    const uint32_t gw = loadG(P.thisG, P, x, y).w;
    const f3 albedo = mk3(unormToFloat(gw & 0xffu), unormToFloat((gw >> 8) & 0xffu), unormToFloat((gw >> 16) & 0xffu));
    const float4 d4 = P.directImg[pix];
    const f3 d = mk3(d4.x, d4.y, d4.z) * albedo, i = mk3(ind.x, ind.y, ind.z) * albedo;
    P.directImg[pix] = make_float4(d.x, d.y, d.z, 1.0f);
    P.indirectImg[pix] = make_float4(i.x, i.y, i.z, 1.0f);
  }
}

// =================================================================================================
// Display pass — shaders/post.frag (RenderOutput::run, render_output.cpp:224-240) as a compute kernel: one thread per rendered
// pixel (uvCoords = (pixel + 0.5) / size, tm.zoom = 1, tm.renderingRatio = (1, 1); the reference's sampler is NEAREST, so
// texture(img, uvCoords) is texel (x, y)), direct + indirect, Uncharted-2 tonemap (tonemapping.glsl:39-95), pcg3d-noise dither at
// 1/255 (post.frag:50-57, random.glsl:81-92), contrast / brightness / saturation / vignette.  Writes the float colour and its
// RGBA8 packing (what a UNORM swapchain stores).  tm.autoExposure bit 0: the average colour is the 1x1 level of the mip chain that
// RenderOutput::genMipmap blits from the result images (k_mip_blit, level by level), then toneExposure (post.frag:65-70).
// =================================================================================================
// One level of nvvk::cmdGenerateMipmaps: vkCmdBlitImage with VK_FILTER_LINEAR from (sw x sh) to (dw x dh) = max(1, previous / 2);
// destination texel (i, j) samples the source at (i + 0.5) * sw / dw - 0.5, bilinear, clamped to the edge
__global__ void __launch_bounds__(256) k_mip_blit(const float4* __restrict__ src, int sw, int sh, int spitch, float4* __restrict__ dst, int dw, int dh) {
  const int i = blockIdx.x * 32 + threadIdx.x, j = blockIdx.y * 8 + threadIdx.y;
  if (i >= dw || j >= dh) return;
  const float scaleU = (float)sw / (float)dw, scaleV = (float)sh / (float)dh;
  const float a = ((float)i + 0.5f) * scaleU - 0.5f, b = ((float)j + 0.5f) * scaleV - 0.5f;
  const float af = eid_floorf(a), bf = eid_floorf(b);
  const float fa = a - af, fb = b - bf;
  const int x0 = max(0, min(sw - 1, f2i_sat(af))), x1 = max(0, min(sw - 1, f2i_sat(af) + 1));
  const int y0 = max(0, min(sh - 1, f2i_sat(bf))), y1 = max(0, min(sh - 1, f2i_sat(bf) + 1));
  const float4 t00 = src[(size_t)y0 * spitch + x0], t10 = src[(size_t)y0 * spitch + x1], t01 = src[(size_t)y1 * spitch + x0], t11 = src[(size_t)y1 * spitch + x1];
  float4 o;
  o.x = mixf(mixf(t00.x, t10.x, fa), mixf(t01.x, t11.x, fa), fb); o.y = mixf(mixf(t00.y, t10.y, fa), mixf(t01.y, t11.y, fa), fb);
  o.z = mixf(mixf(t00.z, t10.z, fa), mixf(t01.z, t11.z, fa), fb); o.w = mixf(mixf(t00.w, t10.w, fa), mixf(t01.w, t11.w, fa), fb);
  dst[(size_t)j * dw + i] = o;
}
DEV f3 pPow3(f3 c, float e) { return mk3(eid_powf(c.x, e), eid_powf(c.y, e), eid_powf(c.z, e)); }
DEV f3 pUncharted2(f3 c) {
  const float A = 0.15f, B = 0.50f, C = 0.10f, D = 0.20f, E = 0.02f, F = 0.30f;
  return ((c * ((A * c) + C * B)) + D * E) / ((c * ((A * c) + B)) + D * F) + (-(E / F));
}
// toneMap (tonemapping.glsl:78-95, TONEMAP_UNCHARTED): exposure, Uncharted 2 with white scale, linear -> sRGB
DEV f3 pToneMap(f3 hdr, float exposure) {
  f3 c = hdr * exposure;
  c = pUncharted2(c * 2.0f);
  const f3 whiteScale = mk3(1.0f) / pUncharted2(mk3(11.2f));
  return pPow3(c * whiteScale, 1.0f / 2.2f);
}
DEV f3 pClamp01(f3 c) { return mk3(gmin(gmax(c.x, 0.0f), 1.0f), gmin(gmax(c.y, 0.0f), 1.0f), gmin(gmax(c.z, 0.0f), 1.0f)); }

__global__ void __launch_bounds__(256) k_post(const FrameParams P, const Tonemapper tm, float4* __restrict__ outF, uchar4* __restrict__ out8,
                                              const float4* __restrict__ avg) {   // avg[0] / avg[1]: 1x1 mip level of the direct / indirect image
  const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
  const int W = P.st.size.x, H = P.st.size.y;
  if (x >= W || y >= H) return;
  const size_t pix = (size_t)y * P.pitch + x;
  const float4 d4 = P.directImg[pix], i4 = P.indirectImg[pix];
  const int mode = P.st.debugging_mode;
  f3 color;
  if (mode == eDepth) {
    float depth = d4.w;
    depth = depth * eid_powf(2.0f, tm.brightness);
    depth = depth + tm.saturation;
    depth = gmin(gmax(eid_powf(depth, 1.0f / tm.contrast), 0.0f), 1.0f);
    color = mk3(depth);
  } else if (mode > eIndirectStage) {
    color = mk3(d4.x, d4.y, d4.z);
    if (mode == eBaseColor) color = pClamp01(pPow3(color, 0.45454545454545f));
  } else {
    f3 hdr;
    if (mode == eDirectStage) hdr = mk3(d4.x, d4.y, d4.z);
    else if (mode == eIndirectStage) hdr = mk3(i4.x, i4.y, i4.z);
    else hdr = mk3(d4.x, d4.y, d4.z) + mk3(i4.x, i4.y, i4.z);
    if (tm.autoExposure & 1) {                                                    // post.frag:133-152, toneExposure :65-70
      const float4 aD = avg[0], aI = avg[1];
      f3 av;
      if (mode == eDirectStage) av = mk3(aD.x, aD.y, aD.z);
      else if (mode == eIndirectStage) av = mk3(aI.x, aI.y, aI.z);
      else av = mk3(aD.x, aD.y, aD.z) + mk3(aI.x, aI.y, aI.z);
      const float avgLum2 = dot3(av, mk3(0.2126f, 0.7152f, 0.0722f));
      const float XYZy = (0.3575761f * hdr.x + 0.7151522f * hdr.y) + 0.1191920f * hdr.z;   // second row of the column-filled RGB2XYZ, as written
      const float Y = (tm.key / avgLum2) * XYZy;
      const float Yd = (Y * (1.0f + Y / (tm.Ywhite * tm.Ywhite))) / (1.0f + Y);
      hdr = (hdr / XYZy) * Yd;
    }
    // toneMap (TONEMAP_UNCHARTED): exposure, Uncharted 2 with white scale, linear -> sRGB
    const float GAMMA = 2.2f, INV_GAMMA = 1.0f / 2.2f;
    color = pToneMap(hdr, tm.avgLum);
    // dither (post.frag:50-57) with pcg3d noise of the pixel
    uint32_t rx = (uint32_t)x, ry = (uint32_t)y, rz = 0u;
    rx = rx * 1664525u + 1013904223u; ry = ry * 1664525u + 1013904223u; rz = rz * 1664525u + 1013904223u;
    rx += ry * rz; ry += rz * rx; rz += rx * ry;
    rx ^= rx >> 16; ry ^= ry >> 16; rz ^= rz >> 16;
    rx += ry * rz; ry += rz * rx; rz += rx * ry;
    const f3 noise = mk3(__uint_as_float(0x3f800000u | (rx >> 9)), __uint_as_float(0x3f800000u | (ry >> 9)), __uint_as_float(0x3f800000u | (rz >> 9))) + (-1.0f);
    const f3 lin = pPow3(color, GAMMA);
    const float quant = 1.0f / 255.0f;
    const f3 q = pPow3(lin, INV_GAMMA) / quant;
    const f3 c0 = mk3(eid_floorf(q.x), eid_floorf(q.y), eid_floorf(q.z)) * quant;
    const f3 c1 = c0 + quant;
    const f3 discr = mix3(pPow3(c0, GAMMA), pPow3(c1, GAMMA), noise);
    color = mk3(discr.x < lin.x ? c1.x : c0.x, discr.y < lin.y ? c1.y : c0.y, discr.z < lin.z ? c1.z : c0.z);
    color = pClamp01(mix3(mk3(0.5f), color, tm.contrast));                       // contrast
    color = pPow3(color, 1.0f / tm.brightness);                                  // brightness
    const float lumI = dot3(color, mk3(0.299f, 0.587f, 0.114f));                 // saturation
    color = mix3(mk3(lumI), color, tm.saturation);
    const float ux = ((((float)x + 0.5f) / (float)W) * tm.renderingRatio.x - 0.5f) * 2.0f;   // vignette
    const float uy = ((((float)y + 0.5f) / (float)H) * tm.renderingRatio.y - 0.5f) * 2.0f;
    color = color * (1.0f - (ux * ux + uy * uy) * tm.vignette);
  }
  outF[pix] = make_float4(color.x, color.y, color.z, 1.0f);
  const uint32_t p8 = packUnorm4(color.x, color.y, color.z, 1.0f);
  out8[pix] = make_uchar4((unsigned char)(p8 & 0xffu), (unsigned char)((p8 >> 8) & 0xffu), (unsigned char)((p8 >> 16) & 0xffu), (unsigned char)(p8 >> 24));
}

}  // namespace eid

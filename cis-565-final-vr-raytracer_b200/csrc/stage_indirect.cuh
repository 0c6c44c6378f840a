// stage_indirect.cuh — K2, shaders/indirect_stage.comp: shared pieces + the one-thread-per-pixel kernel (the wavefront form is stage_wave.cuh).
#pragma once
#include "frame.cuh"

namespace eid {

// =================================================================================================
// K2 — indirect_stage.comp
// =================================================================================================
struct GISampleD { f3 L, xv, nv, xs, ns; float pHat; };

DEV float misWeight(const FrameParams& P, float f, float g) { return (P.st.MIS > 0) ? powerHeuristic(f, g) : 1.0f; }   // :59-61
DEV bool giSampleValid(const GISampleD& g) { return g.nv.x < 1.1f && !nan3(g.L); }                                       // :117-119

DEV void loadIResv(const float* base, size_t i, GISampleD& g, uint32_t& num, float& weight, float& bigW) {
  const float* p = base + 19 * i;
  g.L = mk3(__ldg(p), __ldg(p + 1), __ldg(p + 2)); g.xv = mk3(__ldg(p + 3), __ldg(p + 4), __ldg(p + 5)); g.nv = mk3(__ldg(p + 6), __ldg(p + 7), __ldg(p + 8));
  g.xs = mk3(__ldg(p + 9), __ldg(p + 10), __ldg(p + 11)); g.ns = mk3(__ldg(p + 12), __ldg(p + 13), __ldg(p + 14)); g.pHat = __ldg(p + 15);
  num = __float_as_uint(__ldg(p + 16)); weight = __ldg(p + 17); bigW = __ldg(p + 18);
}
DEV void storeIResv(float* base, size_t i, const GISampleD& g, uint32_t num, float weight, float bigW) {
  float* p = base + 19 * i;
  p[0] = g.L.x; p[1] = g.L.y; p[2] = g.L.z; p[3] = g.xv.x; p[4] = g.xv.y; p[5] = g.xv.z; p[6] = g.nv.x; p[7] = g.nv.y; p[8] = g.nv.z;
  p[9] = g.xs.x; p[10] = g.xs.y; p[11] = g.xs.z; p[12] = g.ns.x; p[13] = g.ns.y; p[14] = g.ns.z; p[15] = g.pHat;
  p[16] = __uint_as_float(num); p[17] = weight; p[18] = bigW;
}

// State of the primary surface rebuilt from the G-buffer texel of full-res pixel 2*coord (getIndirectStateFromGBuffer,
// pathtrace.glsl:296-313) with the +2e-2 push along ffnormal (indirect_stage.comp:299).  false = sky pixel.
struct GIPrimary { f3 ro, rd; State st; };
DEV f3 unorm3(uint32_t w) { return mk3(unormToFloat(w & 0xffu), unormToFloat((w >> 8) & 0xffu), unormToFloat((w >> 16) & 0xffu)); }
// seed != nullptr: the caller is at the point of main() where getIndirectStateFromGBuffer runs — the FETCH_GEOM_CHECK_4_SUBPIXELS variant
// draws the material id there (pathtrace.glsl:346-357).  seed == nullptr (k_gi_finish rebuilding the state): no draw, st.matID is left to the caller.
DEV bool giPrimary(const FrameParams& P, int x, int y, int Wi, int Hi, GIPrimary& pr, uint32_t* seed) {
  raySpawn<true>(P.cam, x, y, Wi, Hi, pr.ro, pr.rd);
  State& st = pr.st;
  st.mat.emission = mk3(0.f);
  st.isEmitter = false; st.area = 0.f; st.eta = 0.f; st.u = st.v = 0.f;
  st.tangent = mk3(0.f); st.bitangent = mk3(0.f);
  if (P.variant & EID_VARIANT_FETCH_4_SUBPIXELS) {                  // pathtrace.glsl:314-358
    const uint4 g00 = loadG(P.thisG, P, 2 * x, 2 * y), g10 = loadG(P.thisG, P, 2 * x + 1, 2 * y);
    const uint4 g11 = loadG(P.thisG, P, 2 * x + 1, 2 * y + 1), g01 = loadG(P.thisG, P, 2 * x, 2 * y + 1);
    const float depth = __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(__uint_as_float(g00.x), __uint_as_float(g10.x)), __uint_as_float(g11.x)), __uint_as_float(g01.x)), 0.25f);
    if (depth >= __fsub_rn(EID_INFINITY, __fmul_rn(0.0001f, 10.0f))) return false;
    st.position = pr.ro + pr.rd * depth;
    st.normal = (((octDecode(g00.y) + octDecode(g10.y)) + octDecode(g11.y)) + octDecode(g01.y)) * 0.25f;
    st.ffnormal = dot3(st.normal, pr.rd) <= 0.0f ? st.normal : -st.normal;
    st.mat.albedo = (((unorm3(g00.w) + unorm3(g10.w)) + unorm3(g11.w)) + unorm3(g01.w)) * 0.25f;
    auto ch = [](uint32_t w, int k) { return unormToFloat((w >> (8 * k)) & 0xffu); };
    st.mat.metallic = __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(ch(g00.z, 0), ch(g10.z, 0)), ch(g11.z, 0)), ch(g01.z, 0)), 0.25f);
    st.mat.roughness = __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(ch(g00.z, 1), ch(g10.z, 1)), ch(g11.z, 1)), ch(g01.z, 1)), 0.25f);
    st.mat.ior = __fadd_rn(__fmul_rn(__fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(ch(g00.z, 2), ch(g10.z, 2)), ch(g11.z, 2)), ch(g01.z, 2)), 0.25f), MAX_IOR_MINUS_ONE), 1.f);
    st.mat.transmission = __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(ch(g00.z, 3), ch(g01.z, 3)), ch(g11.z, 3)), ch(g10.z, 3)), 0.25f);
    st.matID = 0;
    if (seed) {
      const float r = rnd(*seed);
      st.matID = (r < 0.25f ? g00.w : r < 0.5f ? g10.w : r < 0.75f ? g11.w : g01.w) >> 24;
    }
  } else {
    const uint4 gi = loadG(P.thisG, P, 2 * x, 2 * y);
    const float depth = __uint_as_float(gi.x);
    if (depth >= __fmul_rn(EID_INFINITY, 0.8f)) return false;
    st.position = pr.ro + pr.rd * depth;
    st.normal = octDecode(gi.y);
    st.ffnormal = dot3(st.normal, pr.rd) <= 0.0f ? st.normal : -st.normal;
    st.mat.albedo = unorm3(gi.w);
    st.mat.metallic = unormToFloat(gi.z & 0xffu);
    st.mat.roughness = unormToFloat((gi.z >> 8) & 0xffu);
    st.mat.ior = __fadd_rn(__fmul_rn(unormToFloat((gi.z >> 16) & 0xffu), MAX_IOR_MINUS_ONE), 1.f);
    st.mat.transmission = unormToFloat(gi.z >> 24);
    st.matID = gi.w >> 24;                                // hashed material id
  }
  st.position = st.position + st.ffnormal * 2e-2f;      // :299
  return true;
}

// ReSTIRIndirect (indirect_stage.comp:228-268) + the tail of main (:296-309): temporal reuse, reservoir update with the new
// sample `gs`, validity check, clamp, store, shade, tone-compress, write the pre-denoise indirect image.
DEV void giFinish(const FrameParams& P, int x, int y, int Wi, int Hi, uint32_t& seed, GISampleD gs, float primSamplePdf,
                  f3 primPos, f3 primFfn, float primRough, float primMetal, uint32_t primMatHash, f3 primWo) {
  GISampleD rs; rs.L = mk3(0.f); rs.xv = mk3(0.f); rs.nv = mk3(0.f); rs.xs = mk3(0.f); rs.ns = mk3(0.f); rs.pHat = 0.f;
  uint32_t rnum = 0; float rweight = 0.f, rbigW = 0.f;
  if (P.st.ReSTIRState == eTemporal || P.st.ReSTIRState == eSpatiotemporal) {   // findTemporalNeighbor :74-108
    const float reprojDepth = len3(ld3(P.cam.lastPosition) - primPos);
    // motionVector[2 * coord] and lastGbuffer[that motion index], as direct_stage gathered them for this pixel (FrameParams::k2G)
    const size_t q = (size_t)y * (P.pitch >> 1) + x;
    const short2 mv = P.k2Mv[q];
    const uint4 gl = P.k2G[q];
    const f3 pnorm = octDecode(gl.y);
    const float pdepth = __uint_as_float(gl.x);
    const int cx = mv.x / 2, cy = mv.y / 2;
    if (cx >= 0 && cx < Wi && cy >= 0 && cy < Hi && hash8(primMatHash) == (gl.w & 0xFF000000u) && dot3(primFfn, pnorm) > 0.5f &&
        reprojDepth < __fmul_rn(pdepth, 1.1f))
      loadIResv(P.lastIR, (size_t)cy * Wi + cx, rs, rnum, rweight, rbigW);
  }
  float sampleWeight = 0.0f;
  if (giSampleValid(gs)) {
    gs.pHat = lum3(gs.L);                               // pHatIndirect :63-64
    sampleWeight = __fdiv_rn(gs.pHat, primSamplePdf);
    if (sampleWeight != sampleWeight || sampleWeight < 0.0f) sampleWeight = 0.0f;
  }
  {                                                     // resvUpdate (reservoir.glsl:55-61)
    const float rv = rnd(seed);
    rweight = __fadd_rn(rweight, sampleWeight);
    rnum += 1;
    if (__fmul_rn(rv, rweight) < sampleWeight) rs = gs;
  }
  if (resvInvalidW(rweight)) { rnum = 0; rweight = 0.f; rbigW = 0.f; }
  const int clampN = P.st.reservoirClamp * 2;
  if (rnum > (uint32_t)clampN) { rweight = __fmul_rn(rweight, __fdiv_rn((float)clampN, (float)rnum)); rnum = (uint32_t)clampN; }
  storeIResv(P.thisIR, (size_t)y * Wi + x, rs, rnum, rweight, rbigW);

  f3 indirect = mk3(0.0f);
  if (!resvInvalidW(rweight) && giSampleValid(rs)) {
    const f3 primWi = norm3(rs.xs - rs.xv);
    const float bigW = __fdiv_rn(rweight, __fmul_rn(lum3(rs.L), (float)rnum));   // bigWIndirect :70-72
    indirect = ((rs.L * bsdfEval(mk3(1.0f), primRough, primMetal, rs.nv, primWo, primWi)) * satDot(rs.nv, primWi)) * bigW;
  }
  f3 res = hdrToLdr(clampRadiance(indirect, P.st.fireflyClampThreshold));
  res = clampRadiance(res, P.st.fireflyClampThreshold);
  P.indA[(size_t)y * P.pitch + x] = make_float4(res.x, res.y, res.z, 1.0f);
}

template <bool STATS, bool TEX>
__global__ void __launch_bounds__(64, EID_K2_MIN_BLOCKS) k_indirect_stage(const FrameParams P) {
  const int x = blockIdx.x * 8 + threadIdx.x;
  const int y = tileAlignedRow(P.sFirst / 2, P.sStride / 2, P.sRows / 2);
  RayCounters rc = {0, 0, 0, 0, 0};
  const int Wi = P.st.size.x / 2, Hi = P.st.size.y / 2;
  if (x < Wi && y < Hi) {
    uint32_t seed = tea((uint32_t)Wi * (uint32_t)y + (uint32_t)x, P.st.time);   // :280
    // TILED_MULTIBOUNCE (:283-288): invocation (0,0) of each 8x8 group draws once, the flag is group-wide.
    // Every thread re-derives that draw from the tile origin's seed instead of a shared variable + barrier.
    bool multiBounce;
    if (threadIdx.x == 0 && threadIdx.y == 0) multiBounce = rnd(seed) < 0.25f;
    else {
      uint32_t s0 = tea((uint32_t)Wi * (uint32_t)(y - (int)threadIdx.y) + (uint32_t)(x - (int)threadIdx.x), P.st.time);
      multiBounce = rnd(s0) < 0.25f;
    }
    GIPrimary pr;
    if (!giPrimary(P, x, y, Wi, Hi, pr, &seed)) {
      P.indA[(size_t)y * P.pitch + x] = make_float4(0.f, 0.f, 0.f, 0.f);      // :292-295
    } else {
      State st = pr.st;
      const f3 ro = pr.ro, rd = pr.rd;
      // ---- pathTraceIndirect (:129-226)
      const f3 primWo = -rd;
      const f3 primPos = st.position, primFfn = st.ffnormal;
      const float primRough = st.mat.roughness, primMetal = st.mat.metallic;
      const uint32_t primMatHash = st.matID;
      float primSamplePdf = 0.f;
      GISampleD gs; gs.L = mk3(0.f); gs.nv = mk3(100.0f); gs.xv = mk3(0.f); gs.xs = mk3(0.f); gs.ns = mk3(0.f); gs.pHat = 0.f;   // newGISample :110-115
      f3 throughput = mk3(multiBounce ? 4.0f : 1.0f);
      st.mat.albedo = mk3(1.0f);
      f3 rayO = ro, rayD = rd;
      for (int d = 1; d <= P.st.maxDepth; d++) {
        const f3 wo = -rayD;
        if (d > 1 && P.st.MIS > 0) {                        // SampleDirectLight (pathtrace.glsl:185-202)
          LightSampleD ls; ls.Li = mk3(0.f); ls.wi = mk3(0.f); ls.dist = 0.f;
          float lightPdf = sampleDirectLightNoVisibility<TEX>(P.sc, P.env, P.st, st.position, seed, ls);
          if (!isPdfInvalid(lightPdf)) {
            if (occlusion<STATS, TEX>(P, offsetRay(st.position, st.ffnormal), ls.wi, st.position, ls.dist, seed, rc)) lightPdf = EID_INVALID_PDF;
          } else lightPdf = EID_INVALID_PDF;
          if (!isPdfInvalid(lightPdf)) {
            float bp = bsdfPdf(st.mat.roughness, st.mat.metallic, st.ffnormal, wo, ls.wi);
            float w = misWeight(P, lightPdf, bp);
            gs.L = gs.L + ((((ls.Li * bsdfEval(st.mat.albedo, st.mat.roughness, st.mat.metallic, st.ffnormal, wo, ls.wi)) * absDot(st.ffnormal, ls.wi)) * throughput) / lightPdf) * w;
          }
        }
        f3 sampleWi, sampleBSDF;
        const float samplePdf = bsdfSample(st, st.ffnormal, wo, seed, sampleBSDF, sampleWi);
        if (isPdfInvalid(samplePdf)) break;
        if (d > 1) {
          if (!multiBounce) break;                          // `return` at :164-166 — nothing follows the loop
          throughput = throughput * ((sampleBSDF / samplePdf) * absDot(st.ffnormal, sampleWi));
        } else {
          primSamplePdf = samplePdf;
          gs.xv = st.position;
          gs.nv = st.ffnormal;
        }
        rayO = offsetRay(st.position, st.ffnormal);
        rayD = sampleWi;
        Payload prd;
        closestHit<STATS, TEX>(P, rayO, rayD, prd, seed, rc);
        if (prd.hitT >= __fsub_rn(EID_INFINITY, 1e-4f)) {   // miss (:183-198)
          if (d > 1) {
            float lightPdf;
            const f3 env = envEvalOf<TEX>(P.env, P.st, sampleWi, lightPdf);         // EnvEval (pathtrace.glsl:60-72)
            gs.L = gs.L + (env * throughput) * misWeight(P, samplePdf, lightPdf);
          } else {
            gs.xs = st.position + (sampleWi * EID_INFINITY) * 0.8f;
            gs.ns = -sampleWi;
          }
          break;
        }
        st = getState<TEX>(P.sc, prd, rayD);
        getMaterials<TEX>(P.sc, st, rayD);
        if (st.isEmitter) {                                 // :203-215, LightEval (pathtrace.glsl:74-88)
          if (d > 1) {
            const float lightProb = __fsub_rn(1.0f, P.st.environmentProb);
            const float4 em = __ldg((const float4*)(P.sc.materials + st.matID) + 2);   // emissiveFactor (untextured) drives the pdf
            float lightPdf = __fmul_rn(__fmul_rn(lum709(em.y, em.z, em.w), P.st.lightLuminIntegInv), lightProb);
            lightPdf = __fmul_rn(lightPdf, __fdiv_rn(__fmul_rn(prd.hitT, prd.hitT), absDot(st.ffnormal, sampleWi)));
            // LightEval (pathtrace.glsl:74-88): pdf from the emissive FACTOR, radiance from factor x texture (st.mat.emission has both)
            const f3 Li = st.mat.emission / st.area;
            gs.L = gs.L + (Li * throughput) * misWeight(P, samplePdf, lightPdf);
          } else {
            gs.xs = st.position;
            gs.ns = st.ffnormal;
          }
          break;
        }
        if (d == 1) { gs.xs = st.position; gs.ns = st.ffnormal; }
      }
      giFinish(P, x, y, Wi, Hi, seed, gs, primSamplePdf, primPos, primFfn, primRough, primMetal, primMatHash, primWo);
    }
  }
  flushCounters<STATS>(P, rc);
}

}  // namespace eid

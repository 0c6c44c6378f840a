// stage_indirect.cuh — K2, shaders/indirect_stage.comp: the one-thread-per-pixel kernel and the wavefront form (ray queues + k_trace_queue).
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
DEV bool giPrimary(const FrameParams& P, int x, int y, int Wi, int Hi, GIPrimary& pr) {
  raySpawn<true>(P.cam, x, y, Wi, Hi, pr.ro, pr.rd);
  const uint4 gi = loadG(P.thisG, P, 2 * x, 2 * y);
  const float depth = __uint_as_float(gi.x);
  if (depth >= __fmul_rn(EID_INFINITY, 0.8f)) return false;
  State& st = pr.st;
  st.position = pr.ro + pr.rd * depth;
  st.normal = octDecode(gi.y);
  st.ffnormal = dot3(st.normal, pr.rd) <= 0.0f ? st.normal : -st.normal;
  st.mat.albedo = mk3(unormToFloat(gi.w & 0xffu), unormToFloat((gi.w >> 8) & 0xffu), unormToFloat((gi.w >> 16) & 0xffu));
  st.mat.metallic = unormToFloat(gi.z & 0xffu);
  st.mat.roughness = unormToFloat((gi.z >> 8) & 0xffu);
  st.mat.ior = __fadd_rn(__fmul_rn(unormToFloat((gi.z >> 16) & 0xffu), MAX_IOR_MINUS_ONE), 1.f);
  st.mat.transmission = unormToFloat(gi.z >> 24);
  st.mat.emission = mk3(0.f);
  st.matID = gi.w >> 24;                                // hashed material id
  st.isEmitter = false; st.area = 0.f; st.eta = 0.f; st.u = st.v = 0.f;
  st.tangent = mk3(0.f); st.bitangent = mk3(0.f);
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
    short2 mv = make_short2(0, 0);
    if (2 * x < P.pitch && 2 * y < P.allocH) mv = P.motion[(size_t)(2 * y) * P.pitch + 2 * x];
    const uint4 gl = loadG(P.lastG, P, mv.x, mv.y);
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
  const int y = stripeRow(P.sFirst / 2, P.sStride / 2, P.sRows / 2, 8);
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
    if (!giPrimary(P, x, y, Wi, Hi, pr)) {
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

// =================================================================================================
// K2, wavefront form (scenes without stochastic alpha).  The same per-path arithmetic and RNG draw order as k_indirect_stage,
// cut at the ray queries:
//   k_gi_begin              primary state, multibounce lottery, BSDF sample of depth 1 -> closest-hit queue 1
//   k_trace_queue<false>    closest hits of queue d                                     (dynamic fetch, trace.cuh)
//   k_gi_bounce(d)          miss / emitter / surface of the depth-d hit; for depth d+1: light sample -> shadow queue + its
//                           MIS-weighted term, BSDF sample, throughput, next ray -> closest-hit queue d+1
//   k_trace_queue<true>     the shadow rays of depth d+1, on a second stream beside the closest-hit chain of the deeper bounces
//                           (a shadow result only gates one addition in k_gi_finish)
//   k_gi_finish             L = ordered sum of the unoccluded NEE terms (+ the terminal emitter/environment term), ReSTIR GI
// A shadow ray of an opaque scene consumes no RNG draw, so deferring it does not change any other value; the radiance terms
// are added in the mega-kernel's order (NEE of depth 2, 3, ..., then the terminal term, which always comes last).
// =================================================================================================
// one queue slot per lane that wants one: a single atomicAdd per warp; must be called by all 32 lanes
DEV uint32_t warpEnqueue(uint32_t* counter, bool want) {
  const unsigned m = __ballot_sync(0xffffffffu, want);
  if (!m) return 0;
  const unsigned lane = (threadIdx.y * blockDim.x + threadIdx.x) & 31u;
  const int leader = __ffs(m) - 1;
  uint32_t base = 0;
  if ((int)lane == leader) base = atomicAdd(counter, (uint32_t)__popc(m));
  base = __shfl_sync(0xffffffffu, base, leader);
  return base + (uint32_t)__popc(m & ((1u << lane) - 1u));
}

template <bool TEX>
__global__ void __launch_bounds__(64) k_gi_begin(const FrameParams P) {
  const int x = blockIdx.x * 8 + threadIdx.x;
  const int y = stripeRow(P.sFirst / 2, P.sStride / 2, P.sRows / 2, 8);
  const int Wi = P.st.size.x / 2, Hi = P.st.size.y / 2;
  const uint32_t slot = (blockIdx.y * gridDim.x + blockIdx.x) * 64u + threadIdx.y * 8u + threadIdx.x;
  const WaveView& V = P.wv;
  bool wantRay = false;
  f3 rayO = mk3(0.f), rayD = mk3(0.f);
  float samplePdf = 0.f;
  if (x < Wi && y < Hi) {
    uint32_t seed = tea((uint32_t)Wi * (uint32_t)y + (uint32_t)x, P.st.time);   // :280
    bool multiBounce;                                                          // TILED_MULTIBOUNCE, see k_indirect_stage
    if (threadIdx.x == 0 && threadIdx.y == 0) multiBounce = rnd(seed) < 0.25f;
    else {
      uint32_t s0 = tea((uint32_t)Wi * (uint32_t)(y - (int)threadIdx.y) + (uint32_t)(x - (int)threadIdx.x), P.st.time);
      multiBounce = rnd(s0) < 0.25f;
    }
    GIPrimary pr;
    if (!giPrimary(P, x, y, Wi, Hi, pr)) {
      P.indA[(size_t)y * P.pitch + x] = make_float4(0.f, 0.f, 0.f, 0.f);      // :292-295
    } else {
      State& st = pr.st;
      st.mat.albedo = mk3(1.0f);
      f3 xv = mk3(0.f), nv = mk3(100.0f);                  // newGISample :110-115
      float primSamplePdf = 0.f;
      if (P.st.maxDepth >= 1) {
        f3 sampleWi, sampleBSDF;
        samplePdf = bsdfSample(st, st.ffnormal, -pr.rd, seed, sampleBSDF, sampleWi);
        if (!isPdfInvalid(samplePdf)) {
          primSamplePdf = samplePdf; xv = st.position; nv = st.ffnormal;
          rayO = offsetRay(st.position, st.ffnormal); rayD = sampleWi;
          wantRay = true;
        }
      }
      const float t0 = multiBounce ? 4.0f : 1.0f;
      V.misc[slot] = make_uint4(seed, multiBounce ? GI_MULTIBOUNCE : 0u, 0u, 0u);
      V.thr[slot] = make_float4(t0, t0, t0, 0.f);
      V.gsXv[slot] = make_float4(xv.x, xv.y, xv.z, primSamplePdf);
      V.gsNv[slot] = make_float4(nv.x, nv.y, nv.z, 0.f);
      V.gsXs[slot] = make_float4(0.f, 0.f, 0.f, 0.f);
      V.gsNs[slot] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  const uint32_t j = warpEnqueue(&V.ctr[1], wantRay);
  if (wantRay) {
    V.rayQ[1][2 * (size_t)j] = make_float4(rayO.x, rayO.y, rayO.z, samplePdf);
    V.rayQ[1][2 * (size_t)j + 1] = make_float4(rayD.x, rayD.y, rayD.z, __uint_as_float(slot));
  }
}

template <bool TEX>
__global__ void __launch_bounds__(128) k_gi_bounce(const FrameParams P, int d) {
  const WaveView& V = P.wv;
  const uint32_t n = V.ctr[d];
  const float4* __restrict__ inQ = V.rayQ[d & 1];
  float4* __restrict__ outQ = V.rayQ[(d + 1) & 1];
  const uint32_t nRound = (n + 31u) & ~31u;
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < nRound; j += gridDim.x * blockDim.x) {
    bool wantRay = false, wantShadow = false;
    f3 rayO = mk3(0.f), rayD2 = mk3(0.f), shO = mk3(0.f), shD = mk3(0.f);
    float nextPdf = 0.f, shTmax = 0.f;
    uint32_t slot = 0;
    if (j < n) {
      const float4 r0 = __ldg(inQ + 2 * (size_t)j), r1 = __ldg(inQ + 2 * (size_t)j + 1), h = __ldg(V.hitQ + j);
      const f3 rayD = mk3(r1.x, r1.y, r1.z);               // = sampleWi of depth d
      const float samplePdf = r0.w;
      slot = __float_as_uint(r1.w);
      uint4 misc = V.misc[slot];
      uint32_t seed = misc.x;
      const bool multiBounce = (misc.y & GI_MULTIBOUNCE) != 0u;
      const float4 t4 = V.thr[slot];
      f3 throughput = mk3(t4.x, t4.y, t4.z);
      const int tri = __float_as_int(h.w);
      if (tri < 0) {                                        // miss (:183-198)
        if (d > 1) {
          float lightPdf;
          const f3 env = envEvalOf<TEX>(P.env, P.st, rayD, lightPdf);              // EnvEval (pathtrace.glsl:60-72)
          const f3 add = (env * throughput) * misWeight(P, samplePdf, lightPdf);
          V.hitL[slot] = make_float4(add.x, add.y, add.z, 0.f);
          misc.y |= GI_HITL;
        } else {
          const float4 xv = V.gsXv[slot];                   // = the primary position (the depth-1 sample was valid)
          const f3 xs = mk3(xv.x, xv.y, xv.z) + (rayD * EID_INFINITY) * 0.8f, ns = -rayD;
          V.gsXs[slot] = make_float4(xs.x, xs.y, xs.z, 0.f);
          V.gsNs[slot] = make_float4(ns.x, ns.y, ns.z, 0.f);
        }
      } else {
        const float4 tc = __ldg(P.accel.tris + 3 * (size_t)tri + 2);          // primitiveID, instanceID of the hit triangle
        Payload prd;
        prd.hitT = h.x; prd.baryU = h.y; prd.baryV = h.z; prd.primitiveID = __float_as_int(tc.y); prd.instanceID = __float_as_int(tc.z);
        prd.instanceCustomIndex = P.sc.instances[prd.instanceID].primMesh;
        State st = getState<TEX>(P.sc, prd, rayD);
        getMaterials<TEX>(P.sc, st, rayD);
        if (st.isEmitter) {                                 // :203-215, LightEval (pathtrace.glsl:74-88)
          if (d > 1) {
            const float lightProb = __fsub_rn(1.0f, P.st.environmentProb);
            const float4 em = __ldg((const float4*)(P.sc.materials + st.matID) + 2);
            float lightPdf = __fmul_rn(__fmul_rn(lum709(em.y, em.z, em.w), P.st.lightLuminIntegInv), lightProb);
            lightPdf = __fmul_rn(lightPdf, __fdiv_rn(__fmul_rn(prd.hitT, prd.hitT), absDot(st.ffnormal, rayD)));
            const f3 Li = st.mat.emission / st.area;
            const f3 add = (Li * throughput) * misWeight(P, samplePdf, lightPdf);
            V.hitL[slot] = make_float4(add.x, add.y, add.z, 0.f);
            misc.y |= GI_HITL;
          } else {
            V.gsXs[slot] = make_float4(st.position.x, st.position.y, st.position.z, 0.f);
            V.gsNs[slot] = make_float4(st.ffnormal.x, st.ffnormal.y, st.ffnormal.z, 0.f);
          }
        } else {
          if (d == 1) {
            V.gsXs[slot] = make_float4(st.position.x, st.position.y, st.position.z, 0.f);
            V.gsNs[slot] = make_float4(st.ffnormal.x, st.ffnormal.y, st.ffnormal.z, 0.f);
          }
          if (d + 1 <= P.st.maxDepth) {                     // ---- loop iteration d + 1 up to its ray query
            const f3 wo = -rayD;
            if (P.st.MIS > 0) {                             // SampleDirectLight (pathtrace.glsl:185-202), visibility deferred
              LightSampleD ls; ls.Li = mk3(0.f); ls.wi = mk3(0.f); ls.dist = 0.f;
              const float lightPdf = sampleDirectLightNoVisibility<TEX>(P.sc, P.env, P.st, st.position, seed, ls);
              if (!isPdfInvalid(lightPdf)) {
                shO = offsetRay(st.position, st.ffnormal); shD = ls.wi;
                shTmax = __fsub_rn(__fsub_rn(__fsub_rn(ls.dist, fabsf(__fsub_rn(shO.x, st.position.x))), fabsf(__fsub_rn(shO.y, st.position.y))),
                                   fabsf(__fsub_rn(shO.z, st.position.z)));                       // Occlusion (pathtrace.glsl:18-22)
                wantShadow = true;
                const float bp = bsdfPdf(st.mat.roughness, st.mat.metallic, st.ffnormal, wo, ls.wi);
                const float w = misWeight(P, lightPdf, bp);
                const f3 term = ((((ls.Li * bsdfEval(st.mat.albedo, st.mat.roughness, st.mat.metallic, st.ffnormal, wo, ls.wi)) * absDot(st.ffnormal, ls.wi)) * throughput) / lightPdf) * w;
                V.neeTerm[(size_t)(d - 1) * V.slots + slot] = make_float4(term.x, term.y, term.z, 0.f);
                misc.y |= 1u << (GI_NEE_SHIFT + d - 1);
              }
            }
            f3 sampleWi, sampleBSDF;
            nextPdf = bsdfSample(st, st.ffnormal, wo, seed, sampleBSDF, sampleWi);
            if (!isPdfInvalid(nextPdf) && multiBounce) {    // ordinary tiles `return` here (:164-166)
              throughput = throughput * ((sampleBSDF / nextPdf) * absDot(st.ffnormal, sampleWi));
              rayO = offsetRay(st.position, st.ffnormal); rayD2 = sampleWi;
              wantRay = true;
              V.thr[slot] = make_float4(throughput.x, throughput.y, throughput.z, 0.f);
            }
          }
        }
      }
      misc.x = seed;
      V.misc[slot] = misc;
    }
    const uint32_t js = warpEnqueue(&V.ctr[32 + d - 1], wantShadow);
    if (wantShadow) {
      float4* q = V.shadowQ + 2 * (size_t)(d - 1) * V.slots;
      q[2 * (size_t)js] = make_float4(shO.x, shO.y, shO.z, shTmax);
      q[2 * (size_t)js + 1] = make_float4(shD.x, shD.y, shD.z, __uint_as_float((uint32_t)(d - 1) * V.slots + slot));
    }
    const uint32_t jr = warpEnqueue(&V.ctr[d + 1], wantRay);
    if (wantRay) {
      outQ[2 * (size_t)jr] = make_float4(rayO.x, rayO.y, rayO.z, nextPdf);
      outQ[2 * (size_t)jr + 1] = make_float4(rayD2.x, rayD2.y, rayD2.z, __uint_as_float(slot));
    }
  }
}

__global__ void __launch_bounds__(64) k_gi_finish(const FrameParams P) {
  const int x = blockIdx.x * 8 + threadIdx.x;
  const int y = stripeRow(P.sFirst / 2, P.sStride / 2, P.sRows / 2, 8);
  const int Wi = P.st.size.x / 2, Hi = P.st.size.y / 2;
  if (x >= Wi || y >= Hi) return;
  const uint32_t slot = (blockIdx.y * gridDim.x + blockIdx.x) * 64u + threadIdx.y * 8u + threadIdx.x;
  const WaveView& V = P.wv;
  GIPrimary pr;
  if (!giPrimary(P, x, y, Wi, Hi, pr)) return;              // sky: k_gi_begin wrote the pixel
  const uint4 misc = V.misc[slot];
  uint32_t seed = misc.x;
  const float4 xv = V.gsXv[slot], nv = V.gsNv[slot], xs = V.gsXs[slot], ns = V.gsNs[slot];
  GISampleD gs;
  gs.xv = mk3(xv.x, xv.y, xv.z); gs.nv = mk3(nv.x, nv.y, nv.z); gs.xs = mk3(xs.x, xs.y, xs.z); gs.ns = mk3(ns.x, ns.y, ns.z); gs.pHat = 0.f;
  gs.L = mk3(0.f);
  uint32_t nee = misc.y >> GI_NEE_SHIFT;
  for (int k = 0; nee; ++k, nee >>= 1) {
    if ((nee & 1u) && V.occl[(size_t)k * V.slots + slot] == 0u) {
      const float4 t = V.neeTerm[(size_t)k * V.slots + slot];
      gs.L = gs.L + mk3(t.x, t.y, t.z);
    }
  }
  if (misc.y & GI_HITL) { const float4 t = V.hitL[slot]; gs.L = gs.L + mk3(t.x, t.y, t.z); }
  giFinish(P, x, y, Wi, Hi, seed, gs, xv.w, pr.st.position, pr.st.ffnormal, pr.st.mat.roughness, pr.st.mat.metallic, pr.st.matID, -pr.rd);
}

}  // namespace eid

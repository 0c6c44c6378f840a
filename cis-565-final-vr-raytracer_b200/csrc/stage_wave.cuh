// stage_wave.cuh — K2 (indirect_stage.comp) in its wavefront form: ray queues + k_trace_queue.
#pragma once
#include "stage_indirect.cuh"

namespace eid {

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
  const int y = tileAlignedRow(P.sFirst / 2, P.sStride / 2, P.sRows / 2);
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
    if (!giPrimary(P, x, y, Wi, Hi, pr, &seed)) {
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
      V.misc[slot] = make_uint4(seed, multiBounce ? GI_MULTIBOUNCE : 0u, st.matID, 0u);   // .z: the primary material id (drawn, with the 4-subpixel fetch)
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
  const int y = tileAlignedRow(P.sFirst / 2, P.sStride / 2, P.sRows / 2);
  const int Wi = P.st.size.x / 2, Hi = P.st.size.y / 2;
  if (x >= Wi || y >= Hi) return;
  const uint32_t slot = (blockIdx.y * gridDim.x + blockIdx.x) * 64u + threadIdx.y * 8u + threadIdx.x;
  const WaveView& V = P.wv;
  GIPrimary pr;
  if (!giPrimary(P, x, y, Wi, Hi, pr, nullptr)) return;     // sky: k_gi_begin wrote the pixel
  const uint4 misc = V.misc[slot];
  pr.st.matID = misc.z;
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

// stage_direct.cuh — K1, shaders/direct_stage.comp: k_direct_stage (+ k_direct_spatial for eSpatial / eSpatiotemporal).
#pragma once
#include "frame.cuh"

namespace eid {

// =================================================================================================
// K1 — direct_stage.comp
// =================================================================================================
// SPATIAL (eSpatial / eSpatiotemporal, :224-255): the pixel stops where the reference has its first barrier() — it writes
// tempDirectResv and its continuation record — and k_direct_spatial finishes it once every pixel's entry is written (the race-free
// reading of the reference).  halo = 1: the launch covers the row above and the row below each owned stripe (multi-GPU),
// 64 pixels of one row per block; such pixels write tempDirectResv, which the stripe's edge rows read, and keep their own G-buffer /
// motion / reservoir history (the values the owning rank computes), so that their temporal reuse matches the owner's next frame; they
// write no image and their rays are not counted.
DEV f3 debugInfo(const State& st, int mode) {          // DebugInfo (pathtrace.glsl:362-380)
  switch (mode) {
    case eMetallic: return mk3(st.mat.metallic);
    case eNormal: return (st.normal + mk3(1.0f)) * .5f;
    case eDepth: return mk3(0.0f);
    case eBaseColor: return st.mat.albedo;
    case eEmissive: return st.mat.emission;
    case eRoughness: return mk3(st.mat.roughness);
    case eTexcoord: return mk3(st.u, st.v, 0.f);
    default: return mk3(1000.f, 0.f, 0.f);
  }
}

template <bool STATS, bool TEX, bool SPATIAL>
__global__ void __launch_bounds__(64, EID_K1_MIN_BLOCKS) k_direct_stage(const FrameParams P, const int halo) {
  int x = blockIdx.x * 8 + threadIdx.x, y;
  if (SPATIAL && halo) {
    x = blockIdx.x * 64 + threadIdx.y * 8 + threadIdx.x;
    const int k = blockIdx.y >> 1;
    y = (blockIdx.y & 1) ? P.sFirst + k * P.sStride + P.sRows : P.sFirst + k * P.sStride - 1;
    if (y < 0) y = 0x3fffffff;
  } else {
    y = stripeRow(P.sFirst, P.sStride, P.sRows, 8);
  }
  const bool own = !(SPATIAL && halo);
  RayCounters rc = {0, 0, 0, 0, 0};
  const int W = P.st.size.x, H = P.st.size.y;
  if (x < W && y < H) {
    uint32_t seed = tea((uint32_t)W * (uint32_t)y + (uint32_t)x, P.st.time);   // :279
    f3 ro, rd;
    raySpawn<true>(P.cam, x, y, W, H, ro, rd);
    const size_t pix = (size_t)y * P.pitch + x;
    f3 radiance;
    Payload prd;
    bool finished = true;
    if (!closestHit<STATS, TEX>(P, ro, rd, prd, seed, rc)) {                 // :154-158
      P.thisG[pix] = make_uint4(__float_as_uint(EID_INFINITY), 0u, 0u, EID_INVALID_MAT);
      P.motion[pix] = make_short2(0, 0);
      if (own) radiance = envRadiance<TEX>(P, rd);
    } else {
      rc.primary++;
      State st = getState<TEX>(P.sc, prd, rd);
      getMaterials<TEX>(P.sc, st, rd);
      // createMotionIndex (:125-139)
      float pr[4];
      mat4MulV(P.cam.lastProjView, st.position.x, st.position.y, st.position.z, 1.0f, pr);
      const float mvx = __fadd_rn(__fmul_rn(__fdiv_rn(pr[0], pr[3]), 0.5f), 0.5f), mvy = __fadd_rn(__fmul_rn(__fdiv_rn(pr[1], pr[3]), 0.5f), 0.5f);
      const int mix_ = f2i_sat(__fmul_rn(mvx, (float)W)), miy = f2i_sat(__fmul_rn(mvy, (float)H));
      const short2 mvs = make_short2((short)max(-32768, min(32767, mix_)), (short)max(-32768, min(32767, miy)));   // RG16_SINT store saturates
      P.motion[pix] = mvs;
      if (!((x | y) & 1) && (x >> 1) < (W >> 1) && (y >> 1) < (H >> 1)) {     // pixel 2 * coord of the quarter-res stage: its temporal lookup, gathered here
        const size_t q = (size_t)(y >> 1) * (P.pitch >> 1) + (x >> 1);
        P.k2Mv[q] = mvs;
        P.k2G[q] = loadG(P.lastG, P, mvs.x, mvs.y);
      }
      P.thisG[pix] = encodeGeometryInfo(st, prd.hitT);

      if (P.st.debugging_mode > eIndirectStage) {
        radiance = debugInfo(st, P.st.debugging_mode);
      } else if (st.isEmitter) {
        radiance = st.mat.emission;                        // :172-174
      } else {
        const f3 wo = -rd;
        f3 direct = mk3(0.0f);
        const f3 one = mk3(1.0f);                          // state.mat.albedo = vec3(1.0) (:178-179)
        const f3 shadowOrigin = offsetRay(st.position, st.ffnormal);
        if (P.st.ReSTIRState == eNone) {                   // DirectLight (pathtrace.glsl:204-220)
          LightSampleD ls; ls.Li = mk3(0.f); ls.wi = mk3(0.f); ls.dist = 0.f;
          float pdf = sampleDirectLightNoVisibility<TEX>(P.sc, P.env, P.st, st.position, seed, ls);
          if (!isPdfInvalid(pdf) && !occlusion<STATS, TEX>(P, shadowOrigin, ls.wi, st.position, ls.dist, seed, rc))
            direct = ((ls.Li * bsdfEval(one, st.mat.roughness, st.mat.metallic, st.ffnormal, wo, ls.wi)) * gmax(dot3(st.ffnormal, ls.wi), 0.0f)) / pdf;
        } else {
          DResv resv; resv.Li = mk3(0.f); resv.wi = mk3(0.f); resv.dist = 0.f; resv.num = 0; resv.weight = 0.f;
          for (int i = 0; i < P.st.RISSampleNum; i++) {    // :188-199
            LightSampleD ls; ls.Li = mk3(0.f); ls.wi = mk3(0.f); ls.dist = 0.f;
            float p = sampleDirectLightNoVisibility<TEX>(P.sc, P.env, P.st, st.position, seed, ls);
            f3 pHat = (ls.Li * bsdfEval(one, st.mat.roughness, st.mat.metallic, st.ffnormal, wo, ls.wi)) * fabsf(dot3(st.ffnormal, ls.wi));
            float weight = lum3(pHat / p);
            if (isPdfInvalid(p) || weight != weight) weight = 0.0f;
            resvUpdate(resv, ls.Li, ls.wi, ls.dist, weight, rnd(seed));
          }
          if (occlusion<STATS, TEX>(P, shadowOrigin, resv.wi, st.position, resv.dist, seed, rc)) resv.weight = 0.0f;   // :200-207

          if (P.st.ReSTIRState == eTemporal || P.st.ReSTIRState == eSpatiotemporal) {   // :209-217, findTemporalNeighbor :47-84
            const float reprojDepth = len3(ld3(P.cam.lastPosition) - st.position);
            if (mix_ >= 2 && mix_ < W && miy >= 0 && miy < H) {
              const uint4 gl = loadG(P.lastG, P, mix_, miy);
              const f3 pnorm = octDecode(gl.y);
              const float pdepth = __uint_as_float(gl.x);
              if (hash8(st.matID) == (gl.w & 0xFF000000u) && dot3(st.normal, pnorm) > 0.9f && reprojDepth < __fmul_rn(pdepth, 1.05f)) {
                DResv t;
                loadDResv(P.lastDR, (size_t)miy * W + mix_, t);
                if (!resvInvalidW(t.weight)) {             // resvMerge (reservoir.glsl:69-75)
                  const float rv = rnd(seed);
                  resv.weight = __fadd_rn(resv.weight, t.weight);
                  resv.num += t.num;
                  if (__fmul_rn(rv, resv.weight) < t.weight) { resv.Li = t.Li; resv.wi = t.wi; resv.dist = t.dist; }
                }
              }
            }
          }
          {                                                // :219-222 stored copy: validity check + clamp
            DResv tmp = resv;
            if (resvInvalidW(tmp.weight)) { tmp.num = 0; tmp.weight = 0.f; }
            const int clampN = P.st.RISSampleNum * P.st.reservoirClamp;
            if (tmp.num > (uint32_t)clampN) { tmp.weight = __fmul_rn(tmp.weight, __fdiv_rn((float)clampN, (float)tmp.num)); tmp.num = (uint32_t)clampN; }
            storeDResv(P.thisDR, (size_t)y * W + x, tmp);
          }
          if (SPATIAL && (P.st.ReSTIRState == eSpatial || P.st.ReSTIRState == eSpatiotemporal)) {   // :224-231, up to the barrier
            if (resvInvalidW(resv.weight)) { resv.num = 0; resv.weight = 0.f; }                     // resvCheckValidity
            storeDResv(P.tempDR, (size_t)y * W + x, resv);                                          // cacheTempReservoir
            if (own) {
              const size_t plane = (size_t)P.pitch * P.allocH;
              P.spCont[pix] = make_float4(__uint_as_float(seed), st.mat.roughness, st.mat.metallic, st.mat.emission.z);
              P.spCont[plane + pix] = make_float4(st.normal.x, st.normal.y, st.normal.z, st.ffnormal.x);
              P.spCont[2 * plane + pix] = make_float4(st.ffnormal.y, st.ffnormal.z, st.mat.emission.x, st.mat.emission.y);
            }
            finished = false;
          } else
          if (!resvInvalidW(resv.weight)) {                // :256-261 — shading uses the un-clamped reservoir
            f3 LiBsdf = resv.Li * bsdfEval(one, st.mat.roughness, st.mat.metallic, st.ffnormal, wo, resv.wi);
            direct = ((LiBsdf / lum3(LiBsdf)) * resv.weight) / (float)resv.num;
          }
        }
        if (nan3(direct)) direct = mk3(0.0f);
        radiance = hdrToLdr(clampRadiance(st.mat.emission + direct, P.st.fireflyClampThreshold));
      }
    }
    if (!finished) {
      if (own) P.directOut[pix] = make_float4(0.f, 0.f, 0.f, -1.0f);     // marker: k_direct_spatial completes this pixel
    } else if (own) {
      const f3 px = clampRadiance(radiance, P.st.fireflyClampThreshold);   // :283
      P.directOut[pix] = make_float4(px.x, px.y, px.z, 1.0f);   // thisDirectResultImage, or denoiseDirTempA with DENOISER_DIRECT_BILATERAL (:284-288)
    }
  }
  if (!own) rc = RayCounters{0, 0, 0, 0, 0};     // a halo pixel's rays are the owner's, traced twice: not counted
  flushCounters<STATS>(P, rc);
}

// Second half of direct_stage.comp for eSpatial / eSpatiotemporal (:232-270): mergeSpatialNeighbors twice (5 candidates each, at most one
// pixel away — toConcentricDisk is never scaled by `Radius`), the final merge and the shading.  Reads the neighbours' tempDirectResv
// entries, all written by the k_direct_stage launches before it.
DEV bool mergeSpatialNeighbors(const FrameParams& P, int x, int y, f3 norm, float depth, f3 pnorm, float pdepth, uint32_t& seed, DResv& agg) {   // :110-123
  const int W = P.st.size.x, H = P.st.size.y;
  bool valid = false;
  agg.num = 0; agg.weight = 0.f;                                           // resvReset keeps the light sample
  for (int i = 0; i < 5; i++) {
    const float r0 = rnd(seed), r1 = rnd(seed);                            // findSpatialNeighbor :86-108
    float dx, dy;
    toConcentricDisk(r0, r1, dx, dy);
    const int px = f2i_sat(__fadd_rn(__fadd_rn((float)x, dx), 0.5f)), py = f2i_sat(__fadd_rn(__fadd_rn((float)y, dy), 0.5f));
    if (!(px >= 0 && px < W && py >= 0 && py < H)) continue;
    if (dot3(norm, pnorm) < 0.5f || fabsf(__fsub_rn(depth, pdepth)) > __fmul_rn(depth, 0.1f)) continue;   // against the pixel's OWN G-buffer entry, as there
    DResv sp;
    loadDResvPlain(P.tempDR, (size_t)py * W + px, sp);
    if (!resvInvalidW(sp.weight)) {
      const float rv = rnd(seed);
      agg.weight = __fadd_rn(agg.weight, sp.weight);
      agg.num += sp.num;
      if (__fmul_rn(rv, agg.weight) < sp.weight) { agg.Li = sp.Li; agg.wi = sp.wi; agg.dist = sp.dist; }
      valid = true;
    }
  }
  return valid;
}
__global__ void __launch_bounds__(64) k_direct_spatial(const FrameParams P) {
  const int x = blockIdx.x * 8 + threadIdx.x;
  const int y = stripeRow(P.sFirst, P.sStride, P.sRows, 8);
  const int W = P.st.size.x, H = P.st.size.y;
  if (x >= W || y >= H) return;
  const size_t pix = (size_t)y * P.pitch + x;
  if (P.directOut[pix].w != -1.0f) return;                                 // sky, emitter, debug view: finished by k_direct_stage
  const size_t plane = (size_t)P.pitch * P.allocH;
  const float4 c0 = P.spCont[pix], c1 = P.spCont[plane + pix], c2 = P.spCont[2 * plane + pix];
  uint32_t seed = __float_as_uint(c0.x);
  const float roughness = c0.y, metallic = c0.z;
  const f3 emission = mk3(c2.z, c2.w, c0.w), normal = mk3(c1.x, c1.y, c1.z), ffnormal = mk3(c1.w, c2.x, c2.y);
  const uint4 g = P.thisG[pix];                                            // loadThisGeometryInfo(imageCoords): depth is prd.hitT bit for bit
  const f3 pnorm = octDecode(g.y);
  const float pdepth = __uint_as_float(g.x), depth = pdepth;
  f3 ro, rd;
  raySpawn<true>(P.cam, x, y, W, H, ro, rd);
  const f3 wo = -rd;
  DResv resv;
  loadDResvPlain(P.tempDR, (size_t)y * W + x, resv);                       // the pixel's own entry = its reservoir at the barrier
  DResv spatial; spatial.Li = mk3(0.f); spatial.wi = mk3(0.f); spatial.dist = 0.f; spatial.num = 0; spatial.weight = 0.f;
  DResv agg; agg.Li = mk3(0.f); agg.wi = mk3(0.f); agg.dist = 0.f; agg.num = 0; agg.weight = 0.f;
  for (int round = 0; round < 2; ++round) {                                // :236-252 (the second cacheTempReservoir rewrites the same entry)
    if (mergeSpatialNeighbors(P, x, y, normal, depth, pnorm, pdepth, seed, agg)) {
      if (!resvInvalidW(agg.weight)) {
        const float rv = rnd(seed);
        spatial.weight = __fadd_rn(spatial.weight, agg.weight);
        spatial.num += agg.num;
        if (__fmul_rn(rv, spatial.weight) < agg.weight) { spatial.Li = agg.Li; spatial.wi = agg.wi; spatial.dist = agg.dist; }
      }
    }
  }
  if (!resvInvalidW(spatial.weight)) {                                     // :253-256
    const float rv = rnd(seed);
    resv.weight = __fadd_rn(resv.weight, spatial.weight);
    resv.num += spatial.num;
    if (__fmul_rn(rv, resv.weight) < spatial.weight) { resv.Li = spatial.Li; resv.wi = spatial.wi; resv.dist = spatial.dist; }
  }
  f3 direct = mk3(0.0f);
  if (!resvInvalidW(resv.weight)) {                                        // :259-262
    f3 LiBsdf = resv.Li * bsdfEval(mk3(1.0f), roughness, metallic, ffnormal, wo, resv.wi);
    direct = ((LiBsdf / lum3(LiBsdf)) * resv.weight) / (float)resv.num;
  }
  if (nan3(direct)) direct = mk3(0.0f);
  const f3 radiance = hdrToLdr(clampRadiance(emission + direct, P.st.fireflyClampThreshold));
  const f3 px = clampRadiance(radiance, P.st.fireflyClampThreshold);
  P.directOut[pix] = make_float4(px.x, px.y, px.z, 1.0f);
}

// =================================================================================================
// EID_VARIANT_DIRECT_SPLIT — direct_gen.comp + direct_reuse.comp (the reference builds both pipelines, renderer.cpp:129-132, and never
// dispatches them): the direct stage cut in two at the reservoir, reproduced as written.
// =================================================================================================
// updateGeometryAlbedo (direct_gen.comp:62-65)
DEV void updateGeometryAlbedo(uint4& g, f3 albedo) { g.w = (packUnorm4(albedo.x, albedo.y, albedo.z, 1.0f) & 0x00ffffffu) | (g.w & 0xff000000u); }

template <bool STATS, bool TEX>
__global__ void __launch_bounds__(64, EID_K1_MIN_BLOCKS) k_direct_gen(const FrameParams P) {      // generateGeometryAndReservoir :77-137, main :139-149
  const int x = blockIdx.x * 8 + threadIdx.x;
  const int y = stripeRow(P.sFirst, P.sStride, P.sRows, 8);
  RayCounters rc = {0, 0, 0, 0, 0};
  const int W = P.st.size.x, H = P.st.size.y;
  if (x < W && y < H) {
    uint32_t seed = tea((uint32_t)W * (uint32_t)y + (uint32_t)x, P.st.time);
    f3 ro, rd;
    raySpawn<true>(P.cam, x, y, W, H, ro, rd);
    const size_t pix = (size_t)y * P.pitch + x, index = (size_t)y * W + x;
    DResv resv; resv.Li = mk3(0.f); resv.wi = mk3(0.f); resv.dist = 0.f; resv.num = 0; resv.weight = 0.f;
    Payload prd;
    const bool hit = closestHit<STATS, TEX>(P, ro, rd, prd, seed, rc);
    if (!hit || prd.hitT >= __fmul_rn(EID_INFINITY, 0.8f)) {
      uint4 g = make_uint4(__float_as_uint(EID_INFINITY), 0u, 0u, EID_INVALID_MAT);
      updateGeometryAlbedo(g, envRadiance<TEX>(P, rd));
      P.thisG[pix] = g;
      P.motion[pix] = make_short2(0, 0);
      storeDResv(P.thisDR, index, resv);
    } else {
      rc.primary++;
      State st = getState<TEX>(P.sc, prd, rd);
      getMaterials<TEX>(P.sc, st, rd);
      float pr[4];
      mat4MulV(P.cam.lastProjView, st.position.x, st.position.y, st.position.z, 1.0f, pr);
      const float mvx = __fadd_rn(__fmul_rn(__fdiv_rn(pr[0], pr[3]), 0.5f), 0.5f), mvy = __fadd_rn(__fmul_rn(__fdiv_rn(pr[1], pr[3]), 0.5f), 0.5f);
      const int mix_ = f2i_sat(__fmul_rn(mvx, (float)W)), miy = f2i_sat(__fmul_rn(mvy, (float)H));
      const short2 mvs = make_short2((short)max(-32768, min(32767, mix_)), (short)max(-32768, min(32767, miy)));
      P.motion[pix] = mvs;
      if (!((x | y) & 1) && (x >> 1) < (W >> 1) && (y >> 1) < (H >> 1)) {     // the quarter-res stage's temporal lookup, gathered here (FrameParams::k2G)
        const size_t q = (size_t)(y >> 1) * (P.pitch >> 1) + (x >> 1);
        P.k2Mv[q] = mvs;
        P.k2G[q] = loadG(P.lastG, P, mvs.x, mvs.y);
      }
      uint4 g = encodeGeometryInfo(st, prd.hitT);
      if (P.st.debugging_mode > eIndirectStage) updateGeometryAlbedo(g, debugInfo(st, P.st.debugging_mode));
      else if (st.isEmitter) updateGeometryAlbedo(g, st.mat.emission);
      else {
        const f3 wo = -rd, one = mk3(1.0f);
        for (int i = 0; i < P.st.RISSampleNum; i++) {
          LightSampleD ls; ls.Li = mk3(0.f); ls.wi = mk3(0.f); ls.dist = 0.f;
          float p = sampleDirectLightNoVisibility<TEX>(P.sc, P.env, P.st, st.position, seed, ls);
          f3 pHat = (ls.Li * bsdfEval(one, st.mat.roughness, st.mat.metallic, st.ffnormal, wo, ls.wi)) * fabsf(dot3(st.ffnormal, ls.wi));
          float weight = lum3(pHat / p);
          if (isPdfInvalid(p) || weight != weight) weight = 0.0f;
          resvUpdate(resv, ls.Li, ls.wi, ls.dist, weight, rnd(seed));
        }
        if (occlusion<STATS, TEX>(P, offsetRay(st.position, st.ffnormal), resv.wi, st.position, resv.dist, seed, rc)) resv.weight = 0.0f;
      }
      P.thisG[pix] = g;
      storeDResv(P.thisDR, index, resv);
    }
  }
  flushCounters<STATS>(P, rc);
}

__global__ void __launch_bounds__(64) k_direct_reuse(const FrameParams P) {                       // direct_reuse.comp:102-153
  const int x = blockIdx.x * 8 + threadIdx.x;
  const int y = stripeRow(P.sFirst, P.sStride, P.sRows, 8);
  const int W = P.st.size.x, H = P.st.size.y;
  if (x >= W || y >= H) return;
  const size_t pix = (size_t)y * P.pitch + x, index = (size_t)y * W + x;
  uint32_t seed = tea((uint32_t)((int)index + W * H), P.st.time);
  f3 ro, rd;
  raySpawn<true>(P.cam, x, y, W, H, ro, rd);
  const uint4 g = P.thisG[pix];                                        // getDirectStateFromGBuffer (pathtrace.glsl:277-294)
  const float depth = __uint_as_float(g.x);
  if (depth >= __fmul_rn(EID_INFINITY, 0.8f)) { P.directImg[pix] = make_float4(0.f, 0.f, 0.f, 0.f); return; }   // (always thisDirectResultImage: no bilateral switch in this shader)
  const f3 position = ro + rd * depth, normal = octDecode(g.y);
  const uint32_t matID = g.w >> 24;
  DResv resv;
  loadDResvPlain(P.thisDR, index, resv);
  const f3 Li0 = resv.Li;                                              // LightSample lsample = resv.lightSample: taken BEFORE the merge
  const short2 mv = P.motion[pix];
  if (P.st.ReSTIRState == eTemporal || P.st.ReSTIRState == eSpatiotemporal) {
    const float reprojDepth = len3(ld3(P.cam.lastPosition) - position);
    const int mix_ = mv.x, miy = mv.y;
    if (mix_ >= 2 && mix_ < W && miy >= 0 && miy < H) {                // findTemporalNeighbor :47-84 (same text as direct_stage.comp)
      const uint4 gl = loadG(P.lastG, P, mix_, miy);
      if (hash8(matID) == (gl.w & 0xFF000000u) && dot3(normal, octDecode(gl.y)) > 0.9f && reprojDepth < __fmul_rn(__uint_as_float(gl.x), 1.05f)) {
        DResv t;
        loadDResv(P.lastDR, (size_t)miy * W + mix_, t);
        if (!resvInvalidW(t.weight)) {
          const float rv = rnd(seed);
          resv.weight = __fadd_rn(resv.weight, t.weight);
          resv.num += t.num;
          if (__fmul_rn(rv, resv.weight) < t.weight) { resv.Li = t.Li; resv.wi = t.wi; resv.dist = t.dist; }
        }
      }
    }
  }
  f3 direct = mk3(0.0f);
  if (!resvInvalidW(resv.weight)) direct = Li0;                        // `direct = lsample.Li` (:138; the LiBSDF above it is unused)
  const int clampN = P.st.RISSampleNum * P.st.reservoirClamp;          // resvClamp, then resvCheckValidity
  if (resv.num > (uint32_t)clampN) { resv.weight = __fmul_rn(resv.weight, __fdiv_rn((float)clampN, (float)resv.num)); resv.num = (uint32_t)clampN; }
  if (resvInvalidW(resv.weight)) { resv.num = 0; resv.weight = 0.f; }
  if (nan3(direct)) direct = mk3(0.0f);
  storeDResv(P.thisDR, index, resv);
  const f3 px = hdrToLdr(clampRadiance(direct, P.st.fireflyClampThreshold));
  P.directImg[pix] = make_float4(px.x, px.y, px.z, 1.0f);
}

}  // namespace eid

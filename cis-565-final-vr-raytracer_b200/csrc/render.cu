// render.cu — the per-frame render loop: CUDA kernels for the five live compute shaders of the reference
// and the Renderer object that owns their buffers and launch schedule.
//
//   k_direct_stage      shaders/direct_stage.comp    primary ray, G-buffer, motion index, RIS over M light
//                                                    candidates, shadow ray, temporal reservoir merge, shade
//   k_indirect_stage    shaders/indirect_stage.comp  quarter-res ReSTIR GI: BSDF-sampled path (NEE+MIS from
//                                                    depth 2), per-8x8-tile multibounce lottery, temporal reuse
//   k_denoise<false>    shaders/denoise_direct.comp  A-Trous level 0..3 (one launch per level)
//   k_denoise<true>     shaders/denoise_indirect.comp A-Trous level 0..4 at quarter res
//   k_compose           shaders/compose.comp         re-modulate by albedo, 2x nearest upsample of indirect
//
// Renderer::run (src/renderer.cpp:154-206) = strict stream order of the above: 1 + 1 + 4 + 5 + 1 launches.
#include <algorithm>
#include <cstring>
#include <vector>
#include "accel.h"
#include "common.h"
#include "shade.cuh"
#include "env_host.h"

namespace eid {

// minimum resident 64-thread blocks per SM the compiler must allow for (register cap = 65536 / (64 * blocks))
#ifndef EID_K1_MIN_BLOCKS
#define EID_K1_MIN_BLOCKS 16
#endif
#ifndef EID_K2_MIN_BLOCKS
#define EID_K2_MIN_BLOCKS 16
#endif

// Scratch of the wavefront form of K2 (k_gi_begin / k_trace_queue / k_gi_bounce / k_gi_finish): one slot per thread of the K2
// launch grid (8x8 tiles of quarter-res pixels), planes of float4 indexed by slot; ray queues are compact (filled through counters).
struct WaveView {
  uint32_t slots;          // capacity of every per-slot plane and of every queue
  float4* rayQ[2];         // closest-hit queue, ping-pong by depth parity; entry = (origin.xyz, samplePdf), (direction.xyz, slot bits)
  float4* hitQ;            // result of queue entry j: (hitT, baryU, baryV, triangle index bits; -1 = miss)
  uint4* misc;             // per slot: RNG state, flags (GI_* bits), -, -
  float4* thr;             // per slot: path throughput
  float4* gsXv; float4* gsNv; float4* gsXs; float4* gsNs;   // GISample: (xv, primSamplePdf), nv, xs, ns
  float4* hitL;            // radiance added by the path's terminal emitter hit / environment miss (depth >= 2)
  float4* neeTerm;         // [k * slots + slot]: next-event-estimation term of depth k + 2 (added iff its shadow ray is unoccluded)
  float4* shadowQ;         // any-hit queues, one of `slots` entries per NEE depth k; entry = (origin.xyz, tmax), (direction.xyz, id = k * slots + slot)
  uint32_t* occl;          // [id]: 1 = shadow ray occluded
  uint32_t* ctr;           // [p] entries of the depth-p closest-hit queue (p >= 1), [32 + k] entries of the shadow queue of NEE depth k + 2,
                           // [64 + p] / [96 + k] the fetch cursors of those queues
};
#define GI_MULTIBOUNCE 1u
#define GI_HITL 2u
#define GI_NEE_SHIFT 8
#define GI_MAX_WAVE_DEPTH 25   // flag bits 8..31 hold the NEE terms of depths 2..25

struct FrameParams {
  RtxState st;
  SceneCamera cam;
  DeviceSceneView sc;
  AccelView accel;
  uint4* thisG; const uint4* lastG;
  short2* motion;
  float* thisDR; const float* lastDR;     // DirectReservoir records, 9 floats each, pitch st.size.x
  float* thisIR; const float* lastIR;     // IndirectReservoir records, 19 floats each, pitch st.size.x/2
  float4* directImg; float4* indirectImg;
  float* tempDR;                          // tempDirectResv (spatial reuse), pitch st.size.x; one buffer, persists across frames
  float4* spCont;                         // spatial reuse: what k_direct_spatial needs of a pixel's State, 3 planes of pitch*allocH
  float4* dirA; float4* dirB; float4* indA; float4* indB;
  float4* geomPos; float4* geomNrm;       // denoiser geometry planes (full res): pos.xyz + hash bits / normal.xyz
  float4* geomPosH; float4* geomNrmH;     // same at quarter res (pitch/2), see k_denoise_prep
  EnvView env;                            // HDR lat-long map + alias table, or the constant environment
  int hasNonOpaque;                       // scene has alpha MASK / BLEND instances: ray queries run the stochastic HitTest loop
  int pitch, allocH;                      // allocation size of the 2-D images
  // rows owned by this rank: stripes k = 0..sCount-1 of sRows full-res rows starting at sFirst + k*sStride (all multiples of 16,
  // so no 8x8 quarter-res tile straddles two ranks).  Single GPU: one stripe covering the frame.
  int sFirst, sStride, sRows, sCount;
  WaveView wv;
  unsigned long long* counters;           // per frame: [0] closest-hit rays, [1] any-hit rays, [2] primary hits, [3] inner-node
                                          // visits, [4] triangle tests (STATS kernels only); since creation: [5] closest, [6] any
};
#define EID_NUM_COUNTERS 8

struct RayCounters { unsigned int closest, any, primary, nodes, tris; };

// blockIdx.y (blocks of `bh` rows) -> image row for a stripe layout given in the kernel's own resolution; rows >= limit are culled by the caller
DEV int stripeRow(int first, int stride, int rows, int bh) {
  const int bps = (rows + bh - 1) / bh;                  // blocks per stripe
  const int k = blockIdx.y / bps, j = blockIdx.y - k * bps;
  const int r = j * bh + threadIdx.y;
  return (r < rows) ? first + k * stride + r : 0x3fffffff;
}

template <bool STATS>
DEV void flushCounters(const FrameParams& P, const RayCounters& c) {
  unsigned int a = c.closest, b = c.any, d = c.primary, n = c.nodes, t = c.tris;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); d += __shfl_xor_sync(0xffffffffu, d, o);
    if (STATS) { n += __shfl_xor_sync(0xffffffffu, n, o); t += __shfl_xor_sync(0xffffffffu, t, o); }
  }
  if (((threadIdx.y * blockDim.x + threadIdx.x) & 31) == 0) {
    if (a) { atomicAdd(&P.counters[0], (unsigned long long)a); atomicAdd(&P.counters[5], (unsigned long long)a); }
    if (b) { atomicAdd(&P.counters[1], (unsigned long long)b); atomicAdd(&P.counters[6], (unsigned long long)b); }
    if (d) atomicAdd(&P.counters[2], (unsigned long long)d);
    if (STATS) { if (n) atomicAdd(&P.counters[3], (unsigned long long)n); if (t) atomicAdd(&P.counters[4], (unsigned long long)t); }
  }
  if (STATS) atomicMax(&P.counters[7], (unsigned long long)c.nodes);   // worst thread (all its rays) since the renderer was created
}

// image access: out-of-bounds loads return 0 (Vulkan robust image access), stores are dropped
DEV uint4 loadG(const uint4* img, const FrameParams& P, int x, int y) {
  if (x < 0 || y < 0 || x >= P.pitch || y >= P.allocH) return make_uint4(0, 0, 0, 0);
  return __ldg(img + (size_t)y * P.pitch + x);
}
DEV float4 loadImg(const float4* img, const FrameParams& P, int x, int y) {
  if (x < 0 || y < 0 || x >= P.pitch || y >= P.allocH) return make_float4(0, 0, 0, 0);
  return img[(size_t)y * P.pitch + x];
}

// HitTest (traceray_rq.glsl:32-102): stochastic alpha for a candidate of a non-FORCE_OPAQUE instance; exactly one draw
DEV bool hitTest(const FrameParams& P, const RayHit& c, uint32_t& seed) {
  const int customIndex = P.sc.instances[c.inst].primMesh;
  const InstanceData gi = P.sc.geoInfo[customIndex];
  const int mi = gi.materialIndex < 0 ? 0 : gi.materialIndex;
  const float4* m = (const float4*)(P.sc.materials + mi);
  const float4 q0 = __ldg(m), q1 = __ldg(m + 1), q4 = __ldg(m + 4);
  float alpha = q0.w;
  const int baseTex = __float_as_int(q1.x);
  if (baseTex > -1) {
    const uint32_t* idx = (const uint32_t*)(uintptr_t)gi.indexAddress + 3 * (size_t)c.prim;
    const float4* vb = (const float4*)(uintptr_t)gi.vertexAddress;
    const float4 a1 = __ldg(vb + 2 * (size_t)__ldg(idx) + 1), b1 = __ldg(vb + 2 * (size_t)__ldg(idx + 1) + 1), c1 = __ldg(vb + 2 * (size_t)__ldg(idx + 2) + 1);
    const float bx = __fsub_rn(__fsub_rn(1.0f, c.u), c.v);
    // raw texcoords, handedness bit included, exactly like the reference (traceray_rq.glsl:76-79)
    const float tu = __fadd_rn(__fadd_rn(__fmul_rn(a1.x, bx), __fmul_rn(b1.x, c.u)), __fmul_rn(c1.x, c.v));
    const float tv = __fadd_rn(__fadd_rn(__fmul_rn(a1.y, bx), __fmul_rn(b1.y, c.u)), __fmul_rn(c1.y, c.v));
    alpha = __fmul_rn(alpha, textureLod0(P.sc, baseTex, tu, tv).w);
  }
  const float opacity = (__float_as_int(q4.y) == ALPHA_MASK) ? (alpha > q4.z ? 1.0f : 0.0f) : alpha;
  return !(rnd(seed) > opacity);
}

// First accepted hit in front-to-back candidate order (t, instanceID, primitiveID): opaque candidates are accepted at once,
// others go through HitTest; a rejected candidate becomes the exclusive lower bound of the next query.
template <bool STATS>
DEV bool firstAcceptedHit(const FrameParams& P, f3 o, f3 d, float tmax, uint32_t& seed, RayHit& h, RayCounters& rc) {
  if (!traverse<false, STATS>(P.accel, o, d, tmax, h, &rc.nodes, &rc.tris)) return false;
  while (!(h.flags & INST_FORCE_OPAQUE)) {
    if (hitTest(P, h, seed)) return true;
    const HitKey low = {h.t, h.inst, h.prim};
    if (!traverse<false, STATS, true>(P.accel, o, d, tmax, h, &rc.nodes, &rc.tris, low)) return false;
  }
  return true;
}

// ClosestHit (traceray_rq.glsl:108-147).  FULL = the scene has non-opaque instances (alpha MASK / BLEND)
template <bool STATS, bool FULL>
DEV bool closestHit(const FrameParams& P, f3 o, f3 d, Payload& prd, uint32_t& seed, RayCounters& rc) {
  rc.closest++;
  RayHit h;
  const bool hit = (FULL && P.hasNonOpaque) ? firstAcceptedHit<STATS>(P, o, d, EID_INFINITY, seed, h, rc)
                                            : traverse<false, STATS>(P.accel, o, d, EID_INFINITY, h, &rc.nodes, &rc.tris);
  if (!hit) { prd.hitT = EID_INFINITY; return false; }
  prd.hitT = h.t; prd.baryU = h.u; prd.baryV = h.v; prd.primitiveID = h.prim; prd.instanceID = h.inst;
  prd.instanceCustomIndex = P.sc.instances[h.inst].primMesh;
  return true;
}
// Occlusion (pathtrace.glsl:18-22) -> AnyHit (traceray_rq.glsl:153-185)
template <bool STATS, bool FULL>
DEV bool occlusion(const FrameParams& P, f3 origin, f3 dir, f3 surfacePos, float dist, uint32_t& seed, RayCounters& rc) {
  rc.any++;
  float tmax = __fsub_rn(__fsub_rn(__fsub_rn(dist, fabsf(__fsub_rn(origin.x, surfacePos.x))), fabsf(__fsub_rn(origin.y, surfacePos.y))),
                         fabsf(__fsub_rn(origin.z, surfacePos.z)));
  RayHit h;
  if (FULL && P.hasNonOpaque) return firstAcceptedHit<STATS>(P, origin, dir, tmax, seed, h, rc);
  return traverse<true, STATS>(P.accel, origin, dir, tmax, h, &rc.nodes, &rc.tris);
}

template <bool FULL>
DEV f3 envRadiance(const FrameParams& P, f3 dir) { return envRadianceOf<FULL>(P.env, P.st, dir); }   // EnvRadiance (pathtrace.glsl:40-47)

// encodeGeometryInfo (direct_stage.comp:37-45)
DEV uint4 encodeGeometryInfo(const State& s, float depth) {
  uint4 g;
  g.x = __float_as_uint(depth);
  g.y = octEncode(s.normal.x, s.normal.y, s.normal.z);
  g.z = packUnorm4(s.mat.metallic, s.mat.roughness, __fdiv_rn(__fsub_rn(s.mat.ior, 1.0f), MAX_IOR_MINUS_ONE), s.mat.transmission);
  g.w = (packUnorm4(s.mat.albedo.x, s.mat.albedo.y, s.mat.albedo.z, 1.0f) & 0xFFFFFFu) + hash8(s.matID);
  return g;
}

DEV void loadDResv(const float* base, size_t i, DResv& r) {
  const float* p = base + 9 * i;
  r.Li = mk3(__ldg(p), __ldg(p + 1), __ldg(p + 2)); r.wi = mk3(__ldg(p + 3), __ldg(p + 4), __ldg(p + 5));
  r.dist = __ldg(p + 6); r.num = __float_as_uint(__ldg(p + 7)); r.weight = __ldg(p + 8);
}
DEV void loadDResvPlain(const float* base, size_t i, DResv& r) {
  const float* p = base + 9 * i;
  r.Li = mk3(p[0], p[1], p[2]); r.wi = mk3(p[3], p[4], p[5]); r.dist = p[6]; r.num = __float_as_uint(p[7]); r.weight = p[8];
}
DEV void storeDResv(float* base, size_t i, const DResv& r) {
  float* p = base + 9 * i;
  p[0] = r.Li.x; p[1] = r.Li.y; p[2] = r.Li.z; p[3] = r.wi.x; p[4] = r.wi.y; p[5] = r.wi.z; p[6] = r.dist; p[7] = __uint_as_float(r.num); p[8] = r.weight;
}

// =================================================================================================
// K1 — direct_stage.comp
// =================================================================================================
// SPATIAL (eSpatial / eSpatiotemporal, :224-255): the pixel stops where the reference has its first barrier() — it writes
// tempDirectResv and its continuation record — and k_direct_spatial finishes it once every pixel's entry is written (the race-free
// reading of the reference).  halo = 1: the launch covers the row above and the row below each owned stripe (multi-GPU),
// 64 pixels of one row per block; such pixels write tempDirectResv, which the stripe's edge rows read, and keep their own G-buffer /
// motion / reservoir history (the values the owning rank computes), so that their temporal reuse matches the owner's next frame; they
// write no image and their rays are not counted.
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
      P.motion[pix] = make_short2((short)max(-32768, min(32767, mix_)), (short)max(-32768, min(32767, miy)));   // RG16_SINT store saturates
      P.thisG[pix] = encodeGeometryInfo(st, prd.hitT);

      if (P.st.debugging_mode > eIndirectStage) {          // DebugInfo (pathtrace.glsl:362-380)
        switch (P.st.debugging_mode) {
          case eMetallic: radiance = mk3(st.mat.metallic); break;
          case eNormal: radiance = (st.normal + mk3(1.0f)) * .5f; break;
          case eDepth: radiance = mk3(0.0f); break;
          case eBaseColor: radiance = st.mat.albedo; break;
          case eEmissive: radiance = st.mat.emission; break;
          case eRoughness: radiance = mk3(st.mat.roughness); break;
          case eTexcoord: radiance = mk3(st.u, st.v, 0.f); break;
          default: radiance = mk3(1000.f, 0.f, 0.f);
        }
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
      if (own) P.directImg[pix] = make_float4(0.f, 0.f, 0.f, -1.0f);     // marker: k_direct_spatial completes this pixel
    } else if (own) {
      const f3 px = clampRadiance(radiance, P.st.fireflyClampThreshold);   // :283
      P.directImg[pix] = make_float4(px.x, px.y, px.z, 1.0f);
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
  if (P.directImg[pix].w != -1.0f) return;                                 // sky, emitter, debug view: finished by k_direct_stage
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
  P.directImg[pix] = make_float4(px.x, px.y, px.z, 1.0f);
}

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

__global__ void __launch_bounds__(256) k_denoise_prep(const FrameParams P, int first, int stride, int rows) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = stripeRow(first, stride, rows, 8);
  const int W = P.st.size.x, H = P.st.size.y, Wi = W / 2, Hi = H / 2;
  if (x >= W || y >= H || y < 0) return;
  float4 a, b;
  thisGeometry(P, x, y, W, H, a, b);
  const size_t pix = (size_t)y * P.pitch + x;
  P.geomPos[pix] = a; P.geomNrm[pix] = b;
  if (!(x & 1) && !(y & 1) && (x >> 1) < Wi && (y >> 1) < Hi) {
    thisGeometry(P, x, y, Wi, Hi, a, b);
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

// parity taps of the device-side shader functions (same numbering and arity as the oracle's orc_fn / the reference-GLSL ref_fn of
// oracle/ref_shim): 0 toConcentricDisk, 1 powerHeuristic, 2 GetSphericalUv, 3 CreateCoordinateSystem, 4 HDRToLDR, 5 LDRToHDR,
// 6 metallicWorkflowBSDF, 7 metallicWorkflowPdf, 8 metallicWorkflowSample, 11 toneMap, 12 OffsetRay, 13 tea, 14 rand x2
// (9 / 10, the reservoir operations, are written inline in the stage kernels and are covered by the frame-level parity tests)
__global__ void k_fn_tap(int which, int ni, int no, const float* __restrict__ in, uint32_t n, float* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = in + (size_t)i * ni;
  float* o = out + (size_t)i * no;
  auto v3 = [](const float* q) { return mk3(q[0], q[1], q[2]); };
  auto put = [](float* q, f3 v) { q[0] = v.x; q[1] = v.y; q[2] = v.z; };
  State st;
  st.mat.albedo = v3(p); st.mat.roughness = ni >= 14 ? p[3] : 0.f; st.mat.metallic = ni >= 14 ? p[4] : 0.f;
  switch (which) {
    case 0: toConcentricDisk(p[0], p[1], o[0], o[1]); break;
    case 1: o[0] = powerHeuristic(p[0], p[1]); break;
    case 2: sphericalUv(v3(p), o[0], o[1]); break;
    case 3: { f3 t, b; createCoordinateSystem(v3(p), t, b); put(o, t); put(o + 3, b); break; }
    case 4: put(o, hdrToLdr(v3(p))); break;
    case 5: put(o, ldrToHdr(v3(p))); break;
    case 6: put(o, bsdfEval(st.mat.albedo, st.mat.roughness, st.mat.metallic, v3(p + 5), v3(p + 8), v3(p + 11))); break;
    case 7: o[0] = bsdfPdf(st.mat.roughness, st.mat.metallic, v3(p + 5), v3(p + 8), v3(p + 11)); break;
    case 8: { f3 bsdf = mk3(0.f), dir = mk3(0.f); o[0] = bsdfSampleR(st, v3(p + 5), v3(p + 8), p[11], p[12], p[13], bsdf, dir); put(o + 1, bsdf); put(o + 4, dir); break; }
    case 11: put(o, pToneMap(v3(p), p[3])); break;
    case 12: put(o, offsetRay(v3(p), v3(p + 3))); break;
    case 13: o[0] = __uint_as_float(tea(__float_as_uint(p[0]), __float_as_uint(p[1]))); break;
    case 14: { uint32_t s = __float_as_uint(p[0]); const float a = rnd(s), b = rnd(s); o[0] = a; o[1] = b; o[2] = __uint_as_float(s); break; }
    default: break;
  }
}

// scene-dependent parity taps (same numbering as orc_ctx_fn / ref_ctx_fn): 0 SampleDirectLightNoVisibility (seed, pos -> pdf, Li, wi, dist,
// seed'), 2 EnvEval, 3 EnvRadiance, 4 raySpawn, 5 clampRadiance, 6 Sample (seed, albedo, roughness, metallic, V, N -> bsdf, L, pdf, seed');
// 1 (LightEval) is written inline in the indirect stage
__global__ void k_ctx_tap(const FrameParams P, int which, int ni, int no, const float* __restrict__ in, uint32_t n, float* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = in + (size_t)i * ni;
  float* o = out + (size_t)i * no;
  auto v3 = [](const float* q) { return mk3(q[0], q[1], q[2]); };
  auto put = [](float* q, f3 v) { q[0] = v.x; q[1] = v.y; q[2] = v.z; };
  switch (which) {
    case 0: {
      uint32_t seed = __float_as_uint(p[0]);
      LightSampleD ls; ls.Li = mk3(0.f); ls.wi = mk3(0.f); ls.dist = 0.f;
      o[0] = sampleDirectLightNoVisibility<true>(P.sc, P.env, P.st, v3(p + 1), seed, ls);
      put(o + 1, ls.Li); put(o + 4, ls.wi); o[7] = ls.dist; o[8] = __uint_as_float(seed);
      break;
    }
    case 2: { float pdf = 0.f; put(o, envEvalOf<true>(P.env, P.st, v3(p), pdf)); o[3] = pdf; break; }
    case 3: put(o, envRadianceOf<true>(P.env, P.st, v3(p))); break;
    case 4: { f3 ro, rd; raySpawn<true>(P.cam, (int)p[0], (int)p[1], (int)p[2], (int)p[3], ro, rd); put(o, ro); put(o + 3, rd); break; }
    case 5: put(o, clampRadiance(v3(p), P.st.fireflyClampThreshold)); break;
    case 6: {
      uint32_t seed = __float_as_uint(p[0]);
      State st; st.mat.albedo = v3(p + 1); st.mat.roughness = p[4]; st.mat.metallic = p[5];
      f3 bsdf = mk3(0.f), dir = mk3(0.f);
      const float pdf = bsdfSample(st, v3(p + 9), v3(p + 6), seed, bsdf, dir);
      put(o, bsdf); put(o + 3, dir); o[6] = pdf; o[7] = __uint_as_float(seed);
      break;
    }
    default: break;
  }
}

// parity tap of sun_and_sky (sun_and_sky.glsl:453-601): one direction per thread
__global__ void k_sun_and_sky(const SunAndSky ss, const float* __restrict__ dirs, uint32_t n, float* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const f3 c = sunAndSky(ss, mk3(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]));
  out[3 * i] = c.x; out[3 * i + 1] = c.y; out[3 * i + 2] = c.z;
}

}  // namespace eid

using namespace eid;

// ------------------------------------------------------------------------------------------------
// Renderer object
// ------------------------------------------------------------------------------------------------
struct eid_renderer {
  eid_scene* scene = nullptr;
  eid_accel* accel = nullptr;
  int device = 0;
  uint32_t width = 0, height = 0;
  cudaStream_t stream = nullptr;
  bool ownStream = false;
  uint4* gbuffer[2] = {nullptr, nullptr};
  short2* motion = nullptr;
  float* directResv[2] = {nullptr, nullptr};
  float* indirectResv[2] = {nullptr, nullptr};
  float4* directImg = nullptr; float4* indirectImg = nullptr;
  float* tempDirectResv = nullptr; float4* spatialCont = nullptr;   // spatial reuse (eSpatial / eSpatiotemporal), allocated on first use
  float4* denoiseTemp[4] = {nullptr, nullptr, nullptr, nullptr};
  float4* geom[4] = {nullptr, nullptr, nullptr, nullptr};   // geomPos, geomNrm, geomPosH, geomNrmH
  float4* displayF = nullptr; uchar4* display8 = nullptr;   // output of the display pass (post.frag), allocated on first use
  float4* mipScratch = nullptr;                             // auto exposure: two ping-pong mip levels + the two 1x1 averages
  // wavefront K2 scratch (WaveView): sized for the allocation and for `waveTerms` NEE depths; (re)allocated on demand
  void* waveMem = nullptr; uint32_t waveSlots = 0; int waveTerms = 0; uint32_t* waveCtr = nullptr;
  cudaStream_t shadowStream = nullptr; cudaEvent_t evWave = nullptr, evWaveJoin = nullptr; bool waveOverlap = true;
  int wavefront = 1;          // 1 (default): K2 runs as ray queues + dynamic-fetch traversal when the scene allows it; 0: one mega-kernel
  int traceBlocks = 0;        // grid of k_trace_queue (blocks of 128 threads); 0 = EID_TQ_MIN_BLOCKS per SM
  int smCount = 0;
  int denoiseRowBlock = 2;    // pixels of one column filtered per thread in the A-Trous passes (1, 2 or 4; 2 measured fastest)
  bool strictMath = false;    // bit-reproducible exp in the denoiser (parity runs) instead of MUFU ex2
  unsigned long long* counters = nullptr;
  unsigned long long* countersHost = nullptr;   // pinned
  float env[3] = {0.f, 0.f, 0.f};
  eid_env* envMap = nullptr;
  SunAndSky sunSky{};         // SampleExample::m_sunAndSky (sample_example.hpp:186-203); in_use = 0 until the host sets it
  int lastSet = 0;
  RtxState lastState{};
  bool hasRun = false;
  uint32_t sFirst = 0, sStride = 0, sRows = 0; bool stripesSet = false;   // multi-GPU row ownership (see FrameParams)
  bool profiling = false;
  bool countVisits = false;   // profiling level 2: STATS kernels (node / triangle visit counters)
  cudaEvent_t ev[2 * EID_K_COUNT] = {};   // start/stop per stage
  cudaEvent_t evFork = nullptr, evJoin = nullptr, evPost = nullptr;
  cudaStream_t aux = nullptr;             // second stream: K3 runs beside K2/K4 (see launchFrame)
  bool overlap = true;
  bool postStarted = false;
  cudaStream_t copyStream = nullptr;      // eid_renderer_render_host_async: D2H of frame f overlaps the kernels of frame f+1
  cudaEvent_t evFrameDone = nullptr, evCopyDone = nullptr;
  float4* staging[2] = {nullptr, nullptr};
  bool copyPending = false;
  eid_frame_stats stats{};
  bool statsPending = false;

  void allocate();
  void release();
  void ensureWave(int terms);
  WaveView waveView() const;
};

void eid_renderer::release() {
  for (int i = 0; i < 2; ++i) { cudaFree(gbuffer[i]); cudaFree(directResv[i]); cudaFree(indirectResv[i]); gbuffer[i] = nullptr; directResv[i] = nullptr; indirectResv[i] = nullptr; }
  cudaFree(motion); motion = nullptr;
  cudaFree(tempDirectResv); cudaFree(spatialCont); tempDirectResv = nullptr; spatialCont = nullptr;
  cudaFree(displayF); cudaFree(display8); displayF = nullptr; display8 = nullptr;
  cudaFree(mipScratch); mipScratch = nullptr;
  cudaFree(waveMem); waveMem = nullptr; cudaFree(waveCtr); waveCtr = nullptr; waveSlots = 0; waveTerms = 0;
  cudaFree(directImg); cudaFree(indirectImg); directImg = indirectImg = nullptr;
  for (auto& t : denoiseTemp) { cudaFree(t); t = nullptr; }
  for (auto& t : geom) { cudaFree(t); t = nullptr; }
}

// Renderer::createBuffer / createImage (renderer.cpp:227-302): 2 G-buffers, 1 motion image, 2+2 reservoir buffers,
// 4 denoise temporaries, 2 result images — all zero-initialised (the reference leaves them undefined).
void eid_renderer::allocate() {
  const size_t n = (size_t)width * height, ni = (size_t)(width / 2) * (height / 2);
  auto zalloc = [&](void** p, size_t bytes) {
    bytes = std::max<size_t>(bytes, 16);
    CUDA_CHECK(cudaMalloc(p, bytes));
    CUDA_CHECK(cudaMemsetAsync(*p, 0, bytes, stream));
  };
  for (int i = 0; i < 2; ++i) {
    zalloc((void**)&gbuffer[i], n * 16);
    zalloc((void**)&directResv[i], n * sizeof(DirectReservoir));
    zalloc((void**)&indirectResv[i], ni * sizeof(IndirectReservoir));
  }
  zalloc((void**)&motion, n * 4);
  zalloc((void**)&directImg, n * 16); zalloc((void**)&indirectImg, n * 16);
  for (auto& t : denoiseTemp) zalloc((void**)&t, n * 16);
  zalloc((void**)&geom[0], n * 16); zalloc((void**)&geom[1], n * 16);
  zalloc((void**)&geom[2], ni * 16); zalloc((void**)&geom[3], ni * 16);
  CUDA_CHECK(cudaStreamSynchronize(stream));
  hasRun = false; lastSet = 0;
  if (!stripesSet) { sFirst = 0; sRows = (height + 15) / 16 * 16; sStride = 1u << 20; }
}

// Scratch of the wavefront K2: planes of `waveSlots` float4 (one slot per thread of the largest possible K2 grid of this
// allocation).  Layout of waveMem in float4 units: rayQ0 2S | rayQ1 2S | hitQ S | misc S | thr S | gsXv S | gsNv S | gsXs S |
// gsNs S | hitL S | neeTerm T*S | shadowQ 2*T*S | occl (T*S uint32).  1080p, maxDepth 3: ~125 MB.
void eid_renderer::ensureWave(int terms) {
  terms = std::max(terms, 1);
  const uint32_t tilesX = (width / 2 + 7) / 8, tilesY = ((height + 15) / 16 * 16 / 2 + 7) / 8 + 1;
  const uint32_t S = tilesX * tilesY * 64u;
  if (waveMem && waveSlots == S && waveTerms >= terms) return;
  CUDA_CHECK(cudaStreamSynchronize(stream));
  cudaFree(waveMem); waveMem = nullptr;
  const size_t f4 = (size_t)S * (12 + 3 * (size_t)terms);
  CUDA_CHECK(cudaMalloc(&waveMem, f4 * 16 + (size_t)S * terms * 4));
  if (!waveCtr) CUDA_CHECK(cudaMalloc(&waveCtr, 128 * sizeof(uint32_t)));
  waveSlots = S; waveTerms = terms;
}
WaveView eid_renderer::waveView() const {
  WaveView V;
  const size_t S = waveSlots, T = (size_t)waveTerms;
  float4* b = (float4*)waveMem;
  V.slots = waveSlots;
  V.rayQ[0] = b; V.rayQ[1] = b + 2 * S; V.hitQ = b + 4 * S; V.misc = (uint4*)(b + 5 * S); V.thr = b + 6 * S;
  V.gsXv = b + 7 * S; V.gsNv = b + 8 * S; V.gsXs = b + 9 * S; V.gsNs = b + 10 * S; V.hitL = b + 11 * S;
  V.neeTerm = b + 12 * S; V.shadowQ = b + (12 + T) * S; V.occl = (uint32_t*)(b + (12 + 3 * T) * S);
  V.ctr = waveCtr;
  return V;
}

static void fillParams(eid_renderer* r, const RtxState& st, int frames, FrameParams& P) {
  if (st.size.x <= 0 || st.size.y <= 0 || (uint32_t)st.size.x > r->width || (uint32_t)st.size.y > r->height)
    raise(EID_ERR_INVALID, "RtxState.size %dx%d outside the renderer allocation %ux%u", st.size.x, st.size.y, r->width, r->height);
  if (st.environmentProb > 0.0f && !r->envMap && r->sunSky.in_use != 1)
    raise(EID_ERR_UNSUPPORTED, "environmentProb > 0 needs an environment to sample: an HDR map (eid_env_create + eid_renderer_set_env) or sun & sky (eid_renderer_set_sun_and_sky with in_use = 1)");
  if (st.RISSampleNum < 0 || st.maxDepth < 0) raise(EID_ERR_INVALID, "negative RISSampleNum / maxDepth");
  const int set = (frames + 1) % 2;   // renderer.cpp:157; set i: last* = [i], this* = [!i] (renderer.cpp:341-375)
  P.st = st;
  P.cam = r->scene->host.camera;
  P.sc = r->scene->dev.view(r->scene->host);
  P.accel = r->accel->view();
  P.thisG = r->gbuffer[!set]; P.lastG = r->gbuffer[set];
  P.motion = r->motion;
  P.thisDR = r->directResv[!set]; P.lastDR = r->directResv[set];
  P.thisIR = r->indirectResv[!set]; P.lastIR = r->indirectResv[set];
  P.directImg = r->directImg; P.indirectImg = r->indirectImg;
  if ((st.ReSTIRState == eSpatial || st.ReSTIRState == eSpatiotemporal) && !r->tempDirectResv) {   // m_directTempResv (renderer.cpp:235), on first use
    const size_t n = (size_t)r->width * r->height;
    CUDA_CHECK(cudaMalloc((void**)&r->tempDirectResv, n * sizeof(DirectReservoir)));
    CUDA_CHECK(cudaMemsetAsync(r->tempDirectResv, 0, n * sizeof(DirectReservoir), r->stream));
    CUDA_CHECK(cudaMalloc((void**)&r->spatialCont, 3 * n * sizeof(float4)));
  }
  P.tempDR = r->tempDirectResv; P.spCont = r->spatialCont;
  P.dirA = r->denoiseTemp[0]; P.dirB = r->denoiseTemp[1]; P.indA = r->denoiseTemp[2]; P.indB = r->denoiseTemp[3];
  P.geomPos = r->geom[0]; P.geomNrm = r->geom[1]; P.geomPosH = r->geom[2]; P.geomNrmH = r->geom[3];
  for (int k = 0; k < 3; ++k) P.env.constant[k] = r->env[k];
  P.env.sunSky = r->sunSky;
  P.env.tex = r->envMap ? r->envMap->tex : nullptr; P.env.accel = r->envMap ? r->envMap->accel : nullptr;
  P.env.width = r->envMap ? (int)r->envMap->host.width : 0; P.env.height = r->envMap ? (int)r->envMap->host.height : 0;
  P.hasNonOpaque = r->scene->host.hasNonOpaque ? 1 : 0;
  P.pitch = (int)r->width; P.allocH = (int)r->height;
  P.sFirst = (int)r->sFirst; P.sStride = (int)r->sStride; P.sRows = (int)r->sRows;
  P.sCount = ((int)r->sFirst < st.size.y) ? (st.size.y - 1 - (int)r->sFirst) / (int)r->sStride + 1 : 0;   // stripes that start inside the frame
  P.counters = r->counters;
  memset(&P.wv, 0, sizeof(P.wv));
  if (r->wavefront && !P.hasNonOpaque && st.maxDepth <= GI_MAX_WAVE_DEPTH) { r->ensureWave(st.maxDepth - 1); P.wv = r->waveView(); }
  r->lastSet = set; r->lastState = st; r->hasRun = true;
}

// ---- stage launchers -----------------------------------------------------------------------------------------------------------
// Every stage records a start/stop CUDA-event pair on the stream it runs on when profiling is enabled (stage k: ev[2k], ev[2k+1]).
static inline void markStart(eid_renderer* r, int stage, cudaStream_t st) { if (r->profiling) CUDA_CHECK(cudaEventRecord(r->ev[2 * stage], st)); }
static inline void markStop(eid_renderer* r, int stage, cudaStream_t st) { if (r->profiling) CUDA_CHECK(cudaEventRecord(r->ev[2 * stage + 1], st)); }

static void beginFrame(eid_renderer* r) {
  CUDA_CHECK(cudaMemsetAsync(r->counters, 0, 5 * sizeof(unsigned long long), r->stream));   // per-frame counters only
  memset(&r->stats, 0, sizeof(r->stats));
  r->postStarted = false;
}

static void stageDirect(eid_renderer* r, const FrameParams& P, cudaStream_t st) {
  markStart(r, EID_K_DIRECT, st);
  if (P.sCount > 0) {
    dim3 b(8, 8), g((P.st.size.x + 7) / 8, P.sCount * (P.sRows / 8));
    // TEX = false: lean variant for scenes without a single textured material (no texture branches, no tangent frame)
    const bool tex = r->scene->host.hasTextures || r->scene->host.hasNonOpaque || P.env.sunSky.in_use == 1;
    const bool spatial = P.st.ReSTIRState == eSpatial || P.st.ReSTIRState == eSpatiotemporal;
    if (!spatial) {
      if (r->countVisits) { if (tex) k_direct_stage<true, true, false><<<g, b, 0, st>>>(P, 0); else k_direct_stage<true, false, false><<<g, b, 0, st>>>(P, 0); }
      else { if (tex) k_direct_stage<false, true, false><<<g, b, 0, st>>>(P, 0); else k_direct_stage<false, false, false><<<g, b, 0, st>>>(P, 0); }
      r->stats.kernelLaunches[EID_K_DIRECT]++;
    } else {
      // spatial reuse: every pixel up to its tempDirectResv write (owned stripes, then — when the stripes do not cover the frame — the
      // row above and the row below each stripe), then the neighbour merge and the shading
      auto first = [&](dim3 grid, int halo) {
        if (r->countVisits) { if (tex) k_direct_stage<true, true, true><<<grid, b, 0, st>>>(P, halo); else k_direct_stage<true, false, true><<<grid, b, 0, st>>>(P, halo); }
        else { if (tex) k_direct_stage<false, true, true><<<grid, b, 0, st>>>(P, halo); else k_direct_stage<false, false, true><<<grid, b, 0, st>>>(P, halo); }
        r->stats.kernelLaunches[EID_K_DIRECT]++;
      };
      first(g, 0);
      if (P.sFirst > 0 || P.sCount > 1 || P.sFirst + P.sRows < P.st.size.y) first(dim3((P.st.size.x + 63) / 64, 2 * P.sCount), 1);
      k_direct_spatial<<<g, b, 0, st>>>(P);
      r->stats.kernelLaunches[EID_K_DIRECT]++;
    }
  }
  markStop(r, EID_K_DIRECT, st);
}

template <bool ANY>
static void launchTraceQueue(eid_renderer* r, const FrameParams& P, const float4* rays, const uint32_t* count, uint32_t* cursor, cudaStream_t st) {
  const int g = r->traceBlocks > 0 ? r->traceBlocks : r->smCount * EID_TQ_MIN_BLOCKS;
  if (r->countVisits) k_trace_queue<ANY, true><<<g, 128, 0, st>>>(P.accel, rays, count, cursor, P.wv.hitQ, P.wv.occl, P.counters);
  else k_trace_queue<ANY, false><<<g, 128, 0, st>>>(P.accel, rays, count, cursor, P.wv.hitQ, P.wv.occl, P.counters);
  r->stats.kernelLaunches[EID_K_INDIRECT]++;
}

static void stageIndirect(eid_renderer* r, const FrameParams& P, cudaStream_t st) {
  markStart(r, EID_K_INDIRECT, st);
  if (P.sCount > 0 && P.st.size.x / 2 > 0 && P.st.size.y / 2 > 0) {
    dim3 b(8, 8), g((P.st.size.x / 2 + 7) / 8, P.sCount * (P.sRows / 16));
    const bool tex = r->scene->host.hasTextures || r->scene->host.hasNonOpaque || P.env.sunSky.in_use == 1;
    if (P.wv.slots && (size_t)g.x * g.y * 64 <= P.wv.slots) {
      // wavefront form: begin, then per depth (closest-hit queue, bounce); the shadow queue a bounce fills is traced on the
      // `shadow` stream while the main stream goes on with the next depth (small queues are latency-bound: the longest ray
      // of the C3 scene needs ~150 dependent node visits, ~80 us, however few rays there are); join before finish
      cudaStream_t sh = r->waveOverlap ? r->shadowStream : st;
      CUDA_CHECK(cudaMemsetAsync(P.wv.ctr, 0, 128 * sizeof(uint32_t), st));
      if (tex) k_gi_begin<true><<<g, b, 0, st>>>(P); else k_gi_begin<false><<<g, b, 0, st>>>(P);
      r->stats.kernelLaunches[EID_K_INDIRECT]++;
      const int gb = r->smCount * 8;
      bool forked = false;
      for (int d = 1; d <= P.st.maxDepth; ++d) {
        launchTraceQueue<false>(r, P, P.wv.rayQ[d & 1], P.wv.ctr + d, P.wv.ctr + 64 + d, st);
        if (tex) k_gi_bounce<true><<<gb, 128, 0, st>>>(P, d); else k_gi_bounce<false><<<gb, 128, 0, st>>>(P, d);
        r->stats.kernelLaunches[EID_K_INDIRECT]++;
        if (d + 1 <= P.st.maxDepth && P.st.MIS > 0) {
          if (sh != st) { CUDA_CHECK(cudaEventRecord(r->evWave, st)); CUDA_CHECK(cudaStreamWaitEvent(sh, r->evWave, 0)); forked = true; }
          launchTraceQueue<true>(r, P, P.wv.shadowQ + 2 * (size_t)(d - 1) * P.wv.slots, P.wv.ctr + 32 + d - 1, P.wv.ctr + 96 + d - 1, sh);
        }
      }
      if (forked) { CUDA_CHECK(cudaEventRecord(r->evWaveJoin, sh)); CUDA_CHECK(cudaStreamWaitEvent(st, r->evWaveJoin, 0)); }
      k_gi_finish<<<g, b, 0, st>>>(P);
      r->stats.kernelLaunches[EID_K_INDIRECT]++;
    } else {
      if (r->countVisits) { if (tex) k_indirect_stage<true, true><<<g, b, 0, st>>>(P); else k_indirect_stage<true, false><<<g, b, 0, st>>>(P); }
      else { if (tex) k_indirect_stage<false, true><<<g, b, 0, st>>>(P); else k_indirect_stage<false, false><<<g, b, 0, st>>>(P); }
      r->stats.kernelLaunches[EID_K_INDIRECT]++;
    }
  }
  markStop(r, EID_K_INDIRECT, st);
}

// Row layout of the post stages.  sharded = false: the whole frame (single GPU, or the replicated post of multi-GPU mode A).
// sharded = true (multi-GPU mode B): only what this rank's stripes of the FINAL images need.  An A-Trous level l reaches 2*2^l
// rows, so level l is evaluated on the stripe plus the summed reach of the levels after it (direct: 28/24/16/0 rows for levels
// 0..3; indirect: 60/56/48/32/0 quarter-res rows for levels 0..4); the inputs of level 0 (pre-denoise images, G-buffer) are
// complete on every rank after the first exchange step.  Values are identical to the full-frame evaluation, only the evaluated
// row ranges shrink (overlapping ranges of neighbouring stripes recompute identical values).
struct PostLayout { int first, stride, srows, count; bool sharded; };
static PostLayout postLayout(const FrameParams& P, bool sharded) {
  PostLayout L;
  L.sharded = sharded;
  L.first = sharded ? P.sFirst : 0; L.stride = sharded ? P.sStride : (1 << 20);
  L.srows = sharded ? P.sRows : (P.st.size.y + 15) / 16 * 16; L.count = sharded ? P.sCount : 1;
  return L;
}
static inline unsigned gridRows(const PostLayout& L, int rows, int bh) { return (unsigned)(L.count * ((rows + bh - 1) / bh)); }

// one A-Trous level; `rowBlock` (default 4) selects the row-blocked kernel, 1 the one-pixel-per-thread kernel
template <bool INDIRECT>
static void launchDenoise(eid_renderer* r, const FrameParams& P, const float4* src, float4* dst, int level, int lastLevel, int width,
                          const PostLayout& L, int first, int stride, int rows, cudaStream_t st);

// geometry planes: +-30 full-res rows for K3, +-62 quarter-res rows (= 124 full-res rows) for K4; accounted to the direct denoiser
static void stagePrep(eid_renderer* r, const FrameParams& P, const PostLayout& L, cudaStream_t st) {
  if (P.st.denoise <= 0 || L.count <= 0) return;
  const int rows = L.srows + 2 * 124, W = P.st.size.x;
  dim3 b(32, 8), g((W + 31) / 32, gridRows(L, rows, 8));
  k_denoise_prep<<<g, b, 0, st>>>(P, L.first - 124, L.stride, rows);
  r->stats.kernelLaunches[EID_K_DENOISE_DIRECT]++;
}

template <bool INDIRECT, int R>
static void launchDenoiseR(eid_renderer* r, const FrameParams& P, const float4* src, float4* dst, int level, int lastLevel, int width,
                           const PostLayout& L, int first, int stride, int rows, cudaStream_t st) {
  const int step = 1 << level, vrows = step * ((((rows + step - 1) >> level) + R - 1) / R);
  dim3 b(32, 4), g((width + 31) / 32, gridRows(L, vrows, 4));
  if (r->strictMath) k_denoise<INDIRECT, true, R><<<g, b, 0, st>>>(P, src, dst, level, lastLevel, first, stride, rows);
  else k_denoise<INDIRECT, false, R><<<g, b, 0, st>>>(P, src, dst, level, lastLevel, first, stride, rows);
}
template <bool INDIRECT>
static void launchDenoise(eid_renderer* r, const FrameParams& P, const float4* src, float4* dst, int level, int lastLevel, int width,
                          const PostLayout& L, int first, int stride, int rows, cudaStream_t st) {
  if (r->denoiseRowBlock == 4) launchDenoiseR<INDIRECT, 4>(r, P, src, dst, level, lastLevel, width, L, first, stride, rows, st);
  else if (r->denoiseRowBlock == 2) launchDenoiseR<INDIRECT, 2>(r, P, src, dst, level, lastLevel, width, L, first, stride, rows, st);
  else launchDenoiseR<INDIRECT, 1>(r, P, src, dst, level, lastLevel, width, L, first, stride, rows, st);
}

static void stageDenoiseDirect(eid_renderer* r, const FrameParams& P, const PostLayout& L, cudaStream_t st) {   // renderer.cpp:178-189
  if (P.st.denoise > 0 && L.count > 0) {   // thisDirect -> A -> B -> A -> thisDirect
    const int W = P.st.size.x;
    const float4* src[4] = {P.directImg, P.dirA, P.dirB, P.dirA};
    float4* dst[4] = {P.dirA, P.dirB, P.dirA, P.directImg};
    const int halo[4] = {28, 24, 16, 0};
    for (int i = 0; i < 4; ++i) {
      const int h = L.sharded ? halo[i] : 0, rows = L.srows + 2 * h;
      launchDenoise<false>(r, P, src[i], dst[i], i, 3, W, L, L.first - h, L.stride, rows, st);
      r->stats.kernelLaunches[EID_K_DENOISE_DIRECT]++;
    }
  }
}

static void stageDenoiseIndirect(eid_renderer* r, const FrameParams& P, const PostLayout& L, cudaStream_t st) {   // renderer.cpp:191-202
  const int Wi = P.st.size.x / 2, Hi = P.st.size.y / 2;
  if (P.st.denoise > 0 && Wi > 0 && Hi > 0 && L.count > 0) {   // IndA -> IndB -> IndA -> thisIndirect -> IndA -> IndB
    const float4* src[5] = {P.indA, P.indB, P.indA, P.indirectImg, P.indA};
    float4* dst[5] = {P.indB, P.indA, P.indirectImg, P.indA, P.indB};
    const int halo[5] = {60, 56, 48, 32, 0};
    for (int i = 0; i < 5; ++i) {
      const int h = L.sharded ? halo[i] : 0, rows = L.srows / 2 + 2 * h;
      launchDenoise<true>(r, P, src[i], dst[i], i, 4, Wi, L, L.first / 2 - h, L.stride / 2, rows, st);
      r->stats.kernelLaunches[EID_K_DENOISE_INDIRECT]++;
    }
  }
}

static void stageCompose(eid_renderer* r, const FrameParams& P, const PostLayout& L, cudaStream_t st) {
  markStart(r, EID_K_COMPOSE, st);
  if (L.count > 0) {
    dim3 b(32, 8), g((P.st.size.x + 31) / 32, gridRows(L, L.srows, 8));
    k_compose<<<g, b, 0, st>>>(P, P.st.denoise > 0 ? P.indB : P.indA, L.first, L.stride, L.srows);
    r->stats.kernelLaunches[EID_K_COMPOSE]++;
  }
  markStop(r, EID_K_COMPOSE, st);
}

static void endFrame(eid_renderer* r) {
  CUDA_CHECK(cudaMemcpyAsync(r->countersHost, r->counters, EID_NUM_COUNTERS * sizeof(unsigned long long), cudaMemcpyDeviceToHost, r->stream));
  r->statsPending = true;
  CUDA_CHECK(cudaGetLastError());
}

// run `work` on the auxiliary stream between a fork after everything enqueued so far on the main stream and a later join
static void forkAux(eid_renderer* r) { CUDA_CHECK(cudaEventRecord(r->evFork, r->stream)); CUDA_CHECK(cudaStreamWaitEvent(r->aux, r->evFork, 0)); }
static void joinAux(eid_renderer* r) { CUDA_CHECK(cudaEventRecord(r->evJoin, r->aux)); CUDA_CHECK(cudaStreamWaitEvent(r->stream, r->evJoin, 0)); }

// The reference's schedule is K1, K2, K3x4, K4x5, K5 in strict order (renderer.cpp:154-206), but its true dependencies are
// K1 -> {K2, K3}, K2 -> K4, {K3, K4} -> K5.  K2 (quarter-res, long dependent chains, few threads) leaves most issue slots
// idle, so with `overlap` the direct denoiser K3 runs on a second stream concurrently with K2 + K4.  Results are identical.
static void launchTrace(eid_renderer* r, const FrameParams& P) {
  beginFrame(r);
  stageDirect(r, P, r->stream);
  stageIndirect(r, P, r->stream);
}

static void launchPost(eid_renderer* r, const FrameParams& P, bool sharded) {
  const PostLayout L = postLayout(P, sharded);
  if (r->profiling) CUDA_CHECK(cudaEventRecord(r->evPost, r->stream));   // everything between the trace stages and here = exchange
  r->postStarted = true;
  cudaStream_t sd = r->overlap ? r->aux : r->stream;
  markStart(r, EID_K_DENOISE_DIRECT, r->stream);
  stagePrep(r, P, L, r->stream);
  if (r->overlap) forkAux(r);
  stageDenoiseDirect(r, P, L, sd);
  markStop(r, EID_K_DENOISE_DIRECT, sd);
  markStart(r, EID_K_DENOISE_INDIRECT, r->stream);
  stageDenoiseIndirect(r, P, L, r->stream);
  markStop(r, EID_K_DENOISE_INDIRECT, r->stream);
  if (r->overlap) joinAux(r);
  stageCompose(r, P, L, r->stream);
  endFrame(r);
}

static void launchFrame(eid_renderer* r, const FrameParams& P) {
  if (!r->overlap) { launchTrace(r, P); launchPost(r, P, false); return; }
  const PostLayout L = postLayout(P, false);
  beginFrame(r);
  stageDirect(r, P, r->stream);
  markStart(r, EID_K_DENOISE_DIRECT, r->stream);
  stagePrep(r, P, L, r->stream);
  forkAux(r);
  stageDenoiseDirect(r, P, L, r->aux);
  markStop(r, EID_K_DENOISE_DIRECT, r->aux);
  stageIndirect(r, P, r->stream);
  markStart(r, EID_K_DENOISE_INDIRECT, r->stream);
  stageDenoiseIndirect(r, P, L, r->stream);
  markStop(r, EID_K_DENOISE_INDIRECT, r->stream);
  joinAux(r);
  stageCompose(r, P, L, r->stream);
  endFrame(r);
}

static void* bufferPtr(eid_renderer* r, int which, size_t& bytes) {
  const size_t n = (size_t)r->width * r->height, ni = (size_t)(r->width / 2) * (r->height / 2);
  const int set = r->lastSet;
  switch (which) {
    case EID_BUF_THIS_GBUFFER: bytes = n * 16; return r->gbuffer[!set];
    case EID_BUF_LAST_GBUFFER: bytes = n * 16; return r->gbuffer[set];
    case EID_BUF_MOTION: bytes = n * 4; return r->motion;
    case EID_BUF_THIS_DIRECT_RESV: bytes = n * sizeof(DirectReservoir); return r->directResv[!set];
    case EID_BUF_LAST_DIRECT_RESV: bytes = n * sizeof(DirectReservoir); return r->directResv[set];
    case EID_BUF_THIS_INDIRECT_RESV: bytes = ni * sizeof(IndirectReservoir); return r->indirectResv[!set];
    case EID_BUF_LAST_INDIRECT_RESV: bytes = ni * sizeof(IndirectReservoir); return r->indirectResv[set];
    case EID_BUF_DIRECT: bytes = n * 16; return r->directImg;
    case EID_BUF_INDIRECT: bytes = n * 16; return r->indirectImg;
    case EID_BUF_DENOISE_DIR_A: case EID_BUF_DENOISE_DIR_B: case EID_BUF_DENOISE_IND_A: case EID_BUF_DENOISE_IND_B:
      bytes = n * 16; return r->denoiseTemp[which - EID_BUF_DENOISE_DIR_A];
    case EID_BUF_DISPLAY_F32: bytes = n * 16; return r->displayF;
    case EID_BUF_DISPLAY_RGBA8: bytes = n * 4; return r->display8;
    case EID_BUF_TEMP_DIRECT_RESV: bytes = n * sizeof(DirectReservoir); return r->tempDirectResv;
    default: return nullptr;
  }
}

extern "C" {

int eid_renderer_create(eid_renderer** out, eid_scene* s, eid_accel* a, uint32_t width, uint32_t height, void* cuda_stream) {
  EID_TRY
  if (!out || !s || !a) raise(EID_ERR_INVALID, "eid_renderer_create: null argument");
  if (!s->loaded || a->scene != s) raise(EID_ERR_STATE, "eid_renderer_create: scene not loaded or accel built for another scene");
  if (!width || !height || width > 32768 || height > 32768) raise(EID_ERR_INVALID, "bad render size %ux%u", width, height);
  eid_renderer* r = new eid_renderer();
  try {
    r->scene = s; r->accel = a; r->device = s->dev.device; r->width = width; r->height = height;
    CUDA_CHECK(cudaSetDevice(r->device));
    CUDA_CHECK(cudaDeviceGetAttribute(&r->smCount, cudaDevAttrMultiProcessorCount, r->device));
    if (cuda_stream) r->stream = (cudaStream_t)cuda_stream;
    else { CUDA_CHECK(cudaStreamCreateWithFlags(&r->stream, cudaStreamNonBlocking)); r->ownStream = true; }
    CUDA_CHECK(cudaMalloc(&r->counters, EID_NUM_COUNTERS * sizeof(unsigned long long)));
    CUDA_CHECK(cudaMemset(r->counters, 0, EID_NUM_COUNTERS * sizeof(unsigned long long)));
    CUDA_CHECK(cudaMallocHost(&r->countersHost, EID_NUM_COUNTERS * sizeof(unsigned long long)));
    memset(r->countersHost, 0, EID_NUM_COUNTERS * sizeof(unsigned long long));
    for (auto& e : r->ev) CUDA_CHECK(cudaEventCreate(&e));
    CUDA_CHECK(cudaEventCreateWithFlags(&r->evFork, cudaEventDisableTiming)); CUDA_CHECK(cudaEventCreateWithFlags(&r->evJoin, cudaEventDisableTiming));
    CUDA_CHECK(cudaEventCreate(&r->evPost));
    CUDA_CHECK(cudaStreamCreateWithFlags(&r->aux, cudaStreamNonBlocking));
    CUDA_CHECK(cudaStreamCreateWithFlags(&r->shadowStream, cudaStreamNonBlocking));
    CUDA_CHECK(cudaEventCreateWithFlags(&r->evWave, cudaEventDisableTiming)); CUDA_CHECK(cudaEventCreateWithFlags(&r->evWaveJoin, cudaEventDisableTiming));
    r->allocate();
  } catch (...) { eid_renderer_destroy(r); throw; }
  *out = r;
  return EID_OK;
  EID_CATCH
}

int eid_renderer_resize(eid_renderer* r, uint32_t width, uint32_t height) {
  EID_TRY
  if (!r) raise(EID_ERR_INVALID, "eid_renderer_resize: null renderer");
  if (!width || !height || width > 32768 || height > 32768) raise(EID_ERR_INVALID, "bad render size %ux%u", width, height);
  CUDA_CHECK(cudaSetDevice(r->device));
  CUDA_CHECK(cudaStreamSynchronize(r->stream));
  if (r->copyStream) CUDA_CHECK(cudaStreamSynchronize(r->copyStream));
  cudaFree(r->staging[0]); cudaFree(r->staging[1]); r->staging[0] = r->staging[1] = nullptr; r->copyPending = false;
  r->release();
  r->width = width; r->height = height; r->stripesSet = false;
  r->allocate();
  return EID_OK;
  EID_CATCH
}

void eid_renderer_destroy(eid_renderer* r) {
  if (!r) return;
  cudaSetDevice(r->device);
  if (r->stream) cudaStreamSynchronize(r->stream);
  r->release();
  cudaFree(r->counters);
  if (r->countersHost) cudaFreeHost(r->countersHost);
  for (auto& e : r->ev) if (e) cudaEventDestroy(e);
  if (r->evFork) cudaEventDestroy(r->evFork); if (r->evJoin) cudaEventDestroy(r->evJoin); if (r->evPost) cudaEventDestroy(r->evPost);
  if (r->aux) { cudaStreamSynchronize(r->aux); cudaStreamDestroy(r->aux); }
  if (r->shadowStream) { cudaStreamSynchronize(r->shadowStream); cudaStreamDestroy(r->shadowStream); }
  if (r->evWave) cudaEventDestroy(r->evWave); if (r->evWaveJoin) cudaEventDestroy(r->evWaveJoin);
  if (r->copyStream) { cudaStreamSynchronize(r->copyStream); cudaStreamDestroy(r->copyStream); }
  if (r->evFrameDone) cudaEventDestroy(r->evFrameDone); if (r->evCopyDone) cudaEventDestroy(r->evCopyDone);
  cudaFree(r->staging[0]); cudaFree(r->staging[1]);
  if (r->ownStream && r->stream) cudaStreamDestroy(r->stream);
  delete r;
}

// ---- HdrSampling (hdr_sampling.hpp:43-48) ---------------------------------------------------------------------------------
static int envUpload(eid_env* e) {
  CUDA_CHECK(cudaSetDevice(e->device));
  const size_t n = (size_t)e->host.width * e->host.height;
  CUDA_CHECK(cudaMalloc(&e->tex, n * 16));
  CUDA_CHECK(cudaMalloc(&e->accel, n * sizeof(ImptSampData)));
  CUDA_CHECK(cudaMemcpy(e->tex, e->host.pixels.data(), n * 16, cudaMemcpyHostToDevice));
  CUDA_CHECK(cudaMemcpy(e->accel, e->host.accel.data(), n * sizeof(ImptSampData), cudaMemcpyHostToDevice));
  return EID_OK;
}

int eid_env_create(eid_env** out, int device, const float* rgba, uint32_t width, uint32_t height) {
  EID_TRY
  if (!out) raise(EID_ERR_INVALID, "eid_env_create: out is null");
  eid_env* e = new eid_env();
  try {
    e->device = device;
    e->host.build(rgba, width, height);
    if (device != EID_DEVICE_NONE) envUpload(e);
  } catch (...) { eid_env_destroy(e); throw; }
  *out = e;
  return EID_OK;
  EID_CATCH
}

int eid_env_load_hdr(eid_env** out, int device, const char* path) {
  EID_TRY
  if (!out || !path) raise(EID_ERR_INVALID, "eid_env_load_hdr: null argument");
  eid_env* e = new eid_env();
  try {
    e->device = device;
    e->host.loadRadianceHdr(path);
    if (device != EID_DEVICE_NONE) envUpload(e);
  } catch (...) { eid_env_destroy(e); throw; }
  *out = e;
  return EID_OK;
  EID_CATCH
}

void eid_env_destroy(eid_env* e) {
  if (!e) return;
  if (e->tex || e->accel) { cudaSetDevice(e->device); cudaFree(e->tex); cudaFree(e->accel); }
  delete e;
}

float eid_env_integral(eid_env* e) { return e ? e->host.integral : 0.f; }
float eid_env_average(eid_env* e) { return e ? e->host.average : 0.f; }

int eid_env_get_size(eid_env* e, uint32_t* width, uint32_t* height) {
  EID_TRY
  if (!e || !width || !height) raise(EID_ERR_INVALID, "eid_env_get_size: null argument");
  *width = e->host.width; *height = e->host.height;
  return EID_OK;
  EID_CATCH
}

int eid_env_read(eid_env* e, int what, void* dst, size_t bytes) {
  EID_TRY
  if (!e || !dst) raise(EID_ERR_INVALID, "eid_env_read: null argument");
  const size_t n = (size_t)e->host.width * e->host.height;
  const void* src = what == 0 ? (const void*)e->host.accel.data() : (const void*)e->host.pixels.data();
  const size_t total = what == 0 ? n * sizeof(ImptSampData) : n * 16;
  if (what != 0 && what != 1) raise(EID_ERR_INVALID, "eid_env_read: what must be 0 (alias table) or 1 (pixels)");
  if (bytes > total) raise(EID_ERR_INVALID, "read of %zu bytes from %zu", bytes, total);
  memcpy(dst, src, bytes);
  return EID_OK;
  EID_CATCH
}

int eid_renderer_set_env(eid_renderer* r, eid_env* e) {
  EID_TRY
  if (!r) raise(EID_ERR_INVALID, "eid_renderer_set_env: null renderer");
  if (e && (e->device != r->device || !e->tex)) raise(EID_ERR_INVALID, "environment map lives on another device (or is host-only)");
  r->envMap = e;
  return EID_OK;
  EID_CATCH
}

int eid_renderer_set_sun_and_sky(eid_renderer* r, const SunAndSky* ss) {
  EID_TRY
  if (!r || !ss) raise(EID_ERR_INVALID, "eid_renderer_set_sun_and_sky: null argument");
  r->sunSky = *ss;
  return EID_OK;
  EID_CATCH
}

int eid_renderer_run_output(eid_renderer* r, const Tonemapper* tm) {
  EID_TRY
  if (!r || !tm) raise(EID_ERR_INVALID, "eid_renderer_run_output: null argument");
  if (!r->hasRun) raise(EID_ERR_STATE, "eid_renderer_run_output: no frame has been rendered");
  if (tm->autoExposure & 2) raise(EID_ERR_UNSUPPORTED, "post.frag toneLocalExposure (autoExposure bit 1) is never selected by the reference's GUI and reads an uninitialised variable there: not implemented");
  CUDA_CHECK(cudaSetDevice(r->device));
  const size_t n = (size_t)r->width * r->height;
  if (!r->displayF) {
    CUDA_CHECK(cudaMalloc(&r->displayF, n * 16)); CUDA_CHECK(cudaMalloc(&r->display8, n * 4));
    CUDA_CHECK(cudaMemsetAsync(r->displayF, 0, n * 16, r->stream)); CUDA_CHECK(cudaMemsetAsync(r->display8, 0, n * 4, r->stream));
  }
  FrameParams P;
  memset(&P, 0, sizeof(P));
  P.st = r->lastState; P.pitch = (int)r->width; P.allocH = (int)r->height;
  P.directImg = r->directImg; P.indirectImg = r->indirectImg;
  dim3 b(32, 8), g((P.st.size.x + 31) / 32, (P.st.size.y + 7) / 8);
  if (tm->autoExposure & 1) {   // RenderOutput::genMipmap (render_output.cpp:243-253): the chain of the whole images (m_size = the allocation) down to 1 x 1
    if (!r->mipScratch) CUDA_CHECK(cudaMalloc(&r->mipScratch, (2 * ((size_t)(r->width / 2 + 1) * (r->height / 2 + 1)) + 2) * 16));
    const size_t half = (size_t)(r->width / 2 + 1) * (r->height / 2 + 1);
    float4* pp[2] = {r->mipScratch, r->mipScratch + half};
    float4* avg = r->mipScratch + 2 * half;
    for (int img = 0; img < 2; ++img) {
      const float4* src = img ? r->indirectImg : r->directImg;
      int sw = (int)r->width, sh = (int)r->height, spitch = (int)r->width, k = 0;
      if (sw == 1 && sh == 1) CUDA_CHECK(cudaMemcpyAsync(avg + img, src, 16, cudaMemcpyDeviceToDevice, r->stream));
      while (sw > 1 || sh > 1) {
        const int dw = sw > 1 ? sw / 2 : 1, dh = sh > 1 ? sh / 2 : 1;
        float4* dst = (dw == 1 && dh == 1) ? avg + img : pp[k & 1];
        k_mip_blit<<<dim3((dw + 31) / 32, (dh + 7) / 8), b, 0, r->stream>>>(src, sw, sh, spitch, dst, dw, dh);
        src = dst; sw = dw; sh = dh; spitch = dw; ++k;
      }
    }
    k_post<<<g, b, 0, r->stream>>>(P, *tm, r->displayF, r->display8, avg);
  } else {
    k_post<<<g, b, 0, r->stream>>>(P, *tm, r->displayF, r->display8, nullptr);
  }
  CUDA_CHECK(cudaGetLastError());
  return EID_OK;
  EID_CATCH
}

int eid_renderer_fn_tap(eid_renderer* r, const RtxState* st, int which, const float* in, uint32_t n, float* out) {
  EID_TRY
  static const int A[][2] = {{4, 9}, {0, 0}, {3, 4}, {3, 3}, {4, 6}, {3, 3}, {12, 8}};
  if (!r || !st || !in || !out) raise(EID_ERR_INVALID, "eid_renderer_fn_tap: null argument");
  if (which < 0 || which >= (int)(sizeof(A) / sizeof(A[0])) || A[which][0] == 0) raise(EID_ERR_INVALID, "eid_renderer_fn_tap: no device tap %d", which);
  if (n == 0) return EID_OK;
  const int ni = A[which][0], no = A[which][1];
  CUDA_CHECK(cudaSetDevice(r->device));
  FrameParams P;
  memset(&P, 0, sizeof(P));
  P.st = *st; P.cam = r->scene->host.camera; P.sc = r->scene->dev.view(r->scene->host);
  for (int k = 0; k < 3; ++k) P.env.constant[k] = r->env[k];
  P.env.sunSky = r->sunSky;
  P.env.tex = r->envMap ? r->envMap->tex : nullptr; P.env.accel = r->envMap ? r->envMap->accel : nullptr;
  P.env.width = r->envMap ? (int)r->envMap->host.width : 0; P.env.height = r->envMap ? (int)r->envMap->host.height : 0;
  float *din = nullptr, *dout = nullptr;
  CUDA_CHECK(cudaMalloc(&din, (size_t)n * ni * 4));
  if (cudaMalloc(&dout, (size_t)n * no * 4) != cudaSuccess) { cudaFree(din); raise(EID_ERR_CUDA, "cudaMalloc failed"); }
  cudaError_t e = cudaMemcpy(din, in, (size_t)n * ni * 4, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemset(dout, 0, (size_t)n * no * 4);
  if (e == cudaSuccess) {
    k_ctx_tap<<<(n + 63) / 64, 64>>>(P, which, ni, no, din, n, dout);
    e = cudaMemcpy(out, dout, (size_t)n * no * 4, cudaMemcpyDeviceToHost);
  }
  cudaFree(din); cudaFree(dout);
  if (e != cudaSuccess) raise(EID_ERR_CUDA, "eid_renderer_fn_tap: %s", cudaGetErrorString(e));
  return EID_OK;
  EID_CATCH
}

int eid_fn_tap(int device, int which, const float* in, uint32_t n, float* out) {
  EID_TRY
  static const int A[][2] = {{2, 2}, {2, 1}, {3, 2}, {3, 6}, {3, 3}, {3, 3}, {14, 3}, {14, 1}, {14, 7}, {0, 0}, {0, 0}, {4, 3}, {6, 3}, {2, 1}, {2, 3}};
  if (!in || !out) raise(EID_ERR_INVALID, "eid_fn_tap: null argument");
  if (which < 0 || which >= (int)(sizeof(A) / sizeof(A[0])) || A[which][0] == 0) raise(EID_ERR_INVALID, "eid_fn_tap: no device tap %d", which);
  if (n == 0) return EID_OK;
  const int ni = A[which][0], no = A[which][1];
  CUDA_CHECK(cudaSetDevice(device));
  float *din = nullptr, *dout = nullptr;
  CUDA_CHECK(cudaMalloc(&din, (size_t)n * ni * 4));
  if (cudaMalloc(&dout, (size_t)n * no * 4) != cudaSuccess) { cudaFree(din); raise(EID_ERR_CUDA, "cudaMalloc failed"); }
  cudaError_t e = cudaMemcpy(din, in, (size_t)n * ni * 4, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemset(dout, 0, (size_t)n * no * 4);
  if (e == cudaSuccess) {
    k_fn_tap<<<(n + 63) / 64, 64>>>(which, ni, no, din, n, dout);
    e = cudaMemcpy(out, dout, (size_t)n * no * 4, cudaMemcpyDeviceToHost);
  }
  cudaFree(din); cudaFree(dout);
  if (e != cudaSuccess) raise(EID_ERR_CUDA, "eid_fn_tap: %s", cudaGetErrorString(e));
  return EID_OK;
  EID_CATCH
}

int eid_sun_and_sky_eval(int device, const SunAndSky* ss, const float* dirs, uint32_t n, float* rgb) {
  EID_TRY
  if (!ss || !dirs || !rgb) raise(EID_ERR_INVALID, "eid_sun_and_sky_eval: null argument");
  if (n == 0) return EID_OK;
  CUDA_CHECK(cudaSetDevice(device));
  float *dd = nullptr, *dout = nullptr;
  CUDA_CHECK(cudaMalloc(&dd, (size_t)n * 12));
  if (cudaMalloc(&dout, (size_t)n * 12) != cudaSuccess) { cudaFree(dd); raise(EID_ERR_CUDA, "cudaMalloc failed"); }
  cudaError_t e = cudaMemcpy(dd, dirs, (size_t)n * 12, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    k_sun_and_sky<<<(n + 63) / 64, 64>>>(*ss, dd, n, dout);
    e = cudaMemcpy(rgb, dout, (size_t)n * 12, cudaMemcpyDeviceToHost);
  }
  cudaFree(dd); cudaFree(dout);
  if (e != cudaSuccess) raise(EID_ERR_CUDA, "eid_sun_and_sky_eval: %s", cudaGetErrorString(e));
  return EID_OK;
  EID_CATCH
}

int eid_renderer_set_env_constant(eid_renderer* r, const float rgb[3]) {
  EID_TRY
  if (!r || !rgb) raise(EID_ERR_INVALID, "eid_renderer_set_env_constant: null argument");
  for (int k = 0; k < 3; ++k) r->env[k] = rgb[k];
  return EID_OK;
  EID_CATCH
}

int eid_renderer_run(eid_renderer* r, const RtxState* state, int frames) {
  EID_TRY
  if (!r || !state) raise(EID_ERR_INVALID, "eid_renderer_run: null argument");
  CUDA_CHECK(cudaSetDevice(r->device));
  FrameParams P;
  fillParams(r, *state, frames, P);
  launchFrame(r, P);
  return EID_OK;
  EID_CATCH
}

int eid_renderer_run_trace(eid_renderer* r, const RtxState* state, int frames) {
  EID_TRY
  if (!r || !state) raise(EID_ERR_INVALID, "eid_renderer_run_trace: null argument");
  CUDA_CHECK(cudaSetDevice(r->device));
  FrameParams P;
  fillParams(r, *state, frames, P);
  launchTrace(r, P);
  CUDA_CHECK(cudaGetLastError());
  return EID_OK;
  EID_CATCH
}

// run_trace split in two so a multi-GPU host can start exchanging the G-buffer + direct image while indirect_stage runs
int eid_renderer_run_direct(eid_renderer* r, const RtxState* state, int frames) {
  EID_TRY
  if (!r || !state) raise(EID_ERR_INVALID, "eid_renderer_run_direct: null argument");
  CUDA_CHECK(cudaSetDevice(r->device));
  FrameParams P;
  fillParams(r, *state, frames, P);
  beginFrame(r);
  stageDirect(r, P, r->stream);
  CUDA_CHECK(cudaGetLastError());
  return EID_OK;
  EID_CATCH
}

int eid_renderer_run_indirect(eid_renderer* r, const RtxState* state, int frames) {
  EID_TRY
  if (!r || !state) raise(EID_ERR_INVALID, "eid_renderer_run_indirect: null argument");
  CUDA_CHECK(cudaSetDevice(r->device));
  FrameParams P;
  fillParams(r, *state, frames, P);
  stageIndirect(r, P, r->stream);
  CUDA_CHECK(cudaGetLastError());
  return EID_OK;
  EID_CATCH
}

int eid_renderer_run_post(eid_renderer* r, const RtxState* state, int frames) {
  EID_TRY
  if (!r || !state) raise(EID_ERR_INVALID, "eid_renderer_run_post: null argument");
  CUDA_CHECK(cudaSetDevice(r->device));
  FrameParams P;
  fillParams(r, *state, frames, P);
  launchPost(r, P, false);
  return EID_OK;
  EID_CATCH
}

int eid_renderer_run_post_band(eid_renderer* r, const RtxState* state, int frames) {
  EID_TRY
  if (!r || !state) raise(EID_ERR_INVALID, "eid_renderer_run_post_band: null argument");
  CUDA_CHECK(cudaSetDevice(r->device));
  FrameParams P;
  fillParams(r, *state, frames, P);
  launchPost(r, P, true);
  return EID_OK;
  EID_CATCH
}

int eid_renderer_sync(eid_renderer* r) {
  EID_TRY
  if (!r) raise(EID_ERR_INVALID, "eid_renderer_sync: null renderer");
  CUDA_CHECK(cudaSetDevice(r->device));
  CUDA_CHECK(cudaStreamSynchronize(r->stream));   // the aux stream is always joined into the main stream before a frame ends
  return EID_OK;
  EID_CATCH
}

int eid_renderer_get_outputs(eid_renderer* r, const float** direct, const float** indirect) {
  EID_TRY
  if (!r) raise(EID_ERR_INVALID, "eid_renderer_get_outputs: null renderer");
  if (direct) *direct = (const float*)r->directImg;
  if (indirect) *indirect = (const float*)r->indirectImg;
  return EID_OK;
  EID_CATCH
}

int64_t eid_renderer_buffer_bytes(eid_renderer* r, int which) {
  if (!r) return -1;
  size_t b = 0;
  return bufferPtr(r, which, b) ? (int64_t)b : -1;
}

int eid_renderer_read(eid_renderer* r, int which, void* host_dst, size_t bytes) {
  EID_TRY
  if (!r || !host_dst) raise(EID_ERR_INVALID, "eid_renderer_read: null argument");
  size_t b = 0; void* p = bufferPtr(r, which, b);
  if (!p) raise(EID_ERR_INVALID, "no such buffer %d", which);
  if (bytes > b) raise(EID_ERR_INVALID, "read of %zu bytes from a %zu-byte buffer", bytes, b);
  CUDA_CHECK(cudaSetDevice(r->device));
  CUDA_CHECK(cudaMemcpyAsync(host_dst, p, bytes, cudaMemcpyDeviceToHost, r->stream));
  CUDA_CHECK(cudaStreamSynchronize(r->stream));
  return EID_OK;
  EID_CATCH
}

int eid_renderer_write(eid_renderer* r, int which, const void* host_src, size_t bytes) {
  EID_TRY
  if (!r || !host_src) raise(EID_ERR_INVALID, "eid_renderer_write: null argument");
  size_t b = 0; void* p = bufferPtr(r, which, b);
  if (!p) raise(EID_ERR_INVALID, "no such buffer %d", which);
  if (bytes > b) raise(EID_ERR_INVALID, "write of %zu bytes into a %zu-byte buffer", bytes, b);
  CUDA_CHECK(cudaSetDevice(r->device));
  CUDA_CHECK(cudaMemcpyAsync(p, host_src, bytes, cudaMemcpyHostToDevice, r->stream));
  CUDA_CHECK(cudaStreamSynchronize(r->stream));
  return EID_OK;
  EID_CATCH
}

int eid_renderer_render_host(eid_renderer* r, const SceneCamera* cam, const RtxState* state, int frames, float* direct_host, float* indirect_host) {
  EID_TRY
  if (!r || !state) raise(EID_ERR_INVALID, "eid_renderer_render_host: null argument");
  CUDA_CHECK(cudaSetDevice(r->device));
  if (cam) r->scene->host.camera = *cam;
  FrameParams P;
  fillParams(r, *state, frames, P);
  launchFrame(r, P);
  const size_t rowBytes = (size_t)state->size.x * 16;
  if (direct_host) CUDA_CHECK(cudaMemcpy2DAsync(direct_host, rowBytes, r->directImg, (size_t)r->width * 16, rowBytes, state->size.y, cudaMemcpyDeviceToHost, r->stream));
  if (indirect_host) CUDA_CHECK(cudaMemcpy2DAsync(indirect_host, rowBytes, r->indirectImg, (size_t)r->width * 16, rowBytes, state->size.y, cudaMemcpyDeviceToHost, r->stream));
  CUDA_CHECK(cudaStreamSynchronize(r->stream));
  return EID_OK;
  EID_CATCH
}

// Pipelined variant: the frame is enqueued, its two result images are snapshotted device-to-device into staging buffers (so the
// next frame may overwrite the live images at once) and the device-to-host copies run on a dedicated copy stream, overlapping the
// next frame's kernels.  The host buffers are complete after eid_renderer_wait_host (or the next *_async call for the SAME
// buffers, which waits first).  Use two host buffer pairs alternately for full overlap.
int eid_renderer_render_host_async(eid_renderer* r, const SceneCamera* cam, const RtxState* state, int frames, float* direct_host, float* indirect_host) {
  EID_TRY
  if (!r || !state) raise(EID_ERR_INVALID, "eid_renderer_render_host_async: null argument");
  CUDA_CHECK(cudaSetDevice(r->device));
  if (!r->copyStream) {
    CUDA_CHECK(cudaStreamCreateWithFlags(&r->copyStream, cudaStreamNonBlocking));
    CUDA_CHECK(cudaEventCreateWithFlags(&r->evFrameDone, cudaEventDisableTiming));
    CUDA_CHECK(cudaEventCreateWithFlags(&r->evCopyDone, cudaEventDisableTiming));
  }
  const size_t n = (size_t)r->width * r->height * 16;
  if (!r->staging[0]) { CUDA_CHECK(cudaMalloc(&r->staging[0], n)); CUDA_CHECK(cudaMalloc(&r->staging[1], n)); }
  if (cam) r->scene->host.camera = *cam;
  FrameParams P;
  fillParams(r, *state, frames, P);
  launchFrame(r, P);
  // the previous frame's D2H copies must have drained the staging buffers before they are overwritten
  if (r->copyPending) CUDA_CHECK(cudaStreamWaitEvent(r->stream, r->evCopyDone, 0));
  if (direct_host) CUDA_CHECK(cudaMemcpyAsync(r->staging[0], r->directImg, n, cudaMemcpyDeviceToDevice, r->stream));
  if (indirect_host) CUDA_CHECK(cudaMemcpyAsync(r->staging[1], r->indirectImg, n, cudaMemcpyDeviceToDevice, r->stream));
  CUDA_CHECK(cudaEventRecord(r->evFrameDone, r->stream));
  CUDA_CHECK(cudaStreamWaitEvent(r->copyStream, r->evFrameDone, 0));
  const size_t rowBytes = (size_t)state->size.x * 16;
  if (direct_host) CUDA_CHECK(cudaMemcpy2DAsync(direct_host, rowBytes, r->staging[0], (size_t)r->width * 16, rowBytes, state->size.y, cudaMemcpyDeviceToHost, r->copyStream));
  if (indirect_host) CUDA_CHECK(cudaMemcpy2DAsync(indirect_host, rowBytes, r->staging[1], (size_t)r->width * 16, rowBytes, state->size.y, cudaMemcpyDeviceToHost, r->copyStream));
  CUDA_CHECK(cudaEventRecord(r->evCopyDone, r->copyStream));
  r->copyPending = true;
  return EID_OK;
  EID_CATCH
}

int eid_renderer_wait_host(eid_renderer* r) {
  EID_TRY
  if (!r) raise(EID_ERR_INVALID, "eid_renderer_wait_host: null renderer");
  CUDA_CHECK(cudaSetDevice(r->device));
  if (r->copyStream) CUDA_CHECK(cudaStreamSynchronize(r->copyStream));
  r->copyPending = false;
  return EID_OK;
  EID_CATCH
}

int eid_renderer_set_strict_math(eid_renderer* r, int enabled) {
  EID_TRY
  if (!r) raise(EID_ERR_INVALID, "eid_renderer_set_strict_math: null renderer");
  r->strictMath = enabled != 0;
  return EID_OK;
  EID_CATCH
}

int eid_renderer_set_denoise_rows(eid_renderer* r, int rowsPerThread) {
  EID_TRY
  if (!r) raise(EID_ERR_INVALID, "eid_renderer_set_denoise_rows: null renderer");
  if (rowsPerThread != 1 && rowsPerThread != 2 && rowsPerThread != 4) raise(EID_ERR_INVALID, "eid_renderer_set_denoise_rows: 1, 2 or 4");
  r->denoiseRowBlock = rowsPerThread;
  return EID_OK;
  EID_CATCH
}

int eid_renderer_set_wavefront(eid_renderer* r, int enabled, int traceBlocks) {
  EID_TRY
  if (!r) raise(EID_ERR_INVALID, "eid_renderer_set_wavefront: null renderer");
  if (traceBlocks < 0 || traceBlocks > 65535) raise(EID_ERR_INVALID, "eid_renderer_set_wavefront: traceBlocks out of range");
  r->wavefront = enabled != 0;
  r->waveOverlap = enabled != 2;      // 2: wavefront with every queue on the main stream (strictly serial stages)
  r->traceBlocks = traceBlocks;
  return EID_OK;
  EID_CATCH
}

int eid_renderer_set_overlap(eid_renderer* r, int enabled) {
  EID_TRY
  if (!r) raise(EID_ERR_INVALID, "eid_renderer_set_overlap: null renderer");
  r->overlap = enabled != 0;
  return EID_OK;
  EID_CATCH
}

int eid_renderer_set_profiling(eid_renderer* r, int enabled) {
  EID_TRY
  if (!r) raise(EID_ERR_INVALID, "eid_renderer_set_profiling: null renderer");
  r->profiling = enabled != 0;
  r->countVisits = enabled >= 2;
  return EID_OK;
  EID_CATCH
}

int eid_renderer_get_stats(eid_renderer* r, eid_frame_stats* out) {
  EID_TRY
  if (!r || !out) raise(EID_ERR_INVALID, "eid_renderer_get_stats: null argument");
  CUDA_CHECK(cudaSetDevice(r->device));
  CUDA_CHECK(cudaStreamSynchronize(r->stream));
  if (r->statsPending) {
    r->stats.closestHitRays = r->countersHost[0]; r->stats.anyHitRays = r->countersHost[1]; r->stats.primaryHits = r->countersHost[2];
    r->stats.nodeVisits = r->countersHost[3]; r->stats.triangleTests = r->countersHost[4];
    r->stats.totalClosestHitRays = r->countersHost[5]; r->stats.totalAnyHitRays = r->countersHost[6];
    r->stats.maxNodeVisitsPerThread = r->countersHost[7];
    r->stats.launches = 0;
    for (int k = 0; k < EID_K_COUNT; ++k) r->stats.launches += r->stats.kernelLaunches[k];
    if (r->profiling) {
      CUDA_CHECK(cudaStreamSynchronize(r->aux));
      for (int k = 0; k < EID_K_COUNT; ++k) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r->ev[2 * k], r->ev[2 * k + 1]) == cudaSuccess) r->stats.kernelMs[k] = ms; else cudaGetLastError();
      }
      float ms = 0.f;   // multi-GPU: time between the end of the trace stages and the start of the post stages = exchange step 1
      if (r->postStarted && cudaEventElapsedTime(&ms, r->ev[2 * EID_K_INDIRECT + 1], r->evPost) == cudaSuccess) r->stats.exchangeMs = ms; else cudaGetLastError();
    }
    r->statsPending = false;
  }
  *out = r->stats;
  return EID_OK;
  EID_CATCH
}

int eid_renderer_set_band(eid_renderer* r, uint32_t y0, uint32_t y1) {
  EID_TRY
  if (!r) raise(EID_ERR_INVALID, "eid_renderer_set_band: null renderer");
  if (y0 > y1 || y1 > r->height) raise(EID_ERR_INVALID, "band [%u,%u) outside 0..%u", y0, y1, r->height);
  if ((y0 % 16) != 0 || (y1 % 16) != 0) raise(EID_ERR_INVALID, "band edges must be multiples of 16 rows (8x8 half-res tiles must not straddle ranks)");
  if (y1 == y0) raise(EID_ERR_INVALID, "empty band");
  r->sFirst = y0; r->sRows = y1 - y0; r->sStride = 1u << 20; r->stripesSet = true;
  return EID_OK;
  EID_CATCH
}

int eid_renderer_set_stripes(eid_renderer* r, uint32_t rank, uint32_t world, uint32_t stripeRows) {
  EID_TRY
  if (!r) raise(EID_ERR_INVALID, "eid_renderer_set_stripes: null renderer");
  if (!world || rank >= world) raise(EID_ERR_INVALID, "rank %u outside world %u", rank, world);
  if (!stripeRows || (stripeRows % 16) != 0) raise(EID_ERR_INVALID, "stripeRows must be a positive multiple of 16");
  if (r->height % (world * stripeRows) != 0) raise(EID_ERR_INVALID, "renderer height %u is not a multiple of world*stripeRows = %u (pad the allocation)", r->height, world * stripeRows);
  r->sFirst = rank * stripeRows; r->sRows = stripeRows; r->sStride = world * stripeRows; r->stripesSet = true;
  return EID_OK;
  EID_CATCH
}

int eid_renderer_exchange_groups(eid_renderer* r) {
  if (!r) return -1;
  if (!r->stripesSet || r->sStride >= (1u << 20)) return 1;
  return (int)(r->height / r->sStride);
}

// Byte range of this rank's stripe in exchange group `group` of buffer `which`: the group occupies
// [groupOffset, groupOffset + world*chunkBytes) and rank k's chunk starts at groupOffset + k*chunkBytes, so an in-place
// all-gather over that region (equal chunks) completes the group on every rank.
int eid_renderer_exchange_range(eid_renderer* r, int which, uint32_t group, void** dev_base, uint64_t* offset, uint64_t* bytes) {
  EID_TRY
  if (!r || !dev_base || !offset || !bytes) raise(EID_ERR_INVALID, "eid_renderer_exchange_range: null argument");
  size_t total = 0; void* p = bufferPtr(r, which, total);
  if (!p) raise(EID_ERR_INVALID, "no such buffer %d", which);
  if ((int)group >= eid_renderer_exchange_groups(r)) raise(EID_ERR_INVALID, "exchange group %u out of range", group);
  const bool single = r->sStride >= (1u << 20);
  const uint64_t y0 = (uint64_t)r->sFirst + (single ? 0 : (uint64_t)group * r->sStride), y1 = y0 + r->sRows;
  const uint64_t W = r->width;
  const uint64_t sw = r->hasRun ? (uint64_t)r->lastState.size.x : W;   // reservoir buffers are pitched by RtxState.size.x
  uint64_t rowBytes = 0, a = y0, b = y1;
  switch (which) {
    case EID_BUF_THIS_GBUFFER: case EID_BUF_LAST_GBUFFER: case EID_BUF_DIRECT: case EID_BUF_INDIRECT:
    case EID_BUF_DENOISE_DIR_A: case EID_BUF_DENOISE_DIR_B: rowBytes = W * 16; break;
    case EID_BUF_MOTION: rowBytes = W * 4; break;
    case EID_BUF_DENOISE_IND_A: case EID_BUF_DENOISE_IND_B: rowBytes = W * 16; a = y0 / 2; b = y1 / 2; break;   // half-res rows, full-res pitch
    case EID_BUF_THIS_DIRECT_RESV: case EID_BUF_LAST_DIRECT_RESV: rowBytes = sw * sizeof(DirectReservoir); break;
    case EID_BUF_THIS_INDIRECT_RESV: case EID_BUF_LAST_INDIRECT_RESV: rowBytes = (sw / 2) * sizeof(IndirectReservoir); a = y0 / 2; b = y1 / 2; break;
    default: raise(EID_ERR_INVALID, "no band layout for buffer %d", which);
  }
  if (b * rowBytes > total) raise(EID_ERR_INVALID, "stripe exceeds the buffer (allocation not padded to the stripe layout?)");
  *dev_base = p; *offset = a * rowBytes; *bytes = (b - a) * rowBytes;
  return EID_OK;
  EID_CATCH
}

int eid_renderer_band_range(eid_renderer* r, int which, void** dev_base, uint64_t* offset, uint64_t* bytes) {
  return eid_renderer_exchange_range(r, which, 0, dev_base, offset, bytes);
}

}  // extern "C"

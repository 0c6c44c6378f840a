// frame.cuh — what every stage kernel of the render loop shares: FrameParams (all device pointers of a frame), the wavefront scratch view, ray counters,
// robust image access, ClosestHit / AnyHit with the stochastic-alpha HitTest loop, G-buffer and reservoir record access.
#pragma once
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
  float4* directOut;                      // where direct_stage stores: directImg, or dirA with EID_VARIANT_DIRECT_BILATERAL (direct_stage.comp:284-288)
  int variant;                            // EID_VARIANT_* bits (the reference's compile-time shader switches)
  float* tempDR;                          // tempDirectResv (spatial reuse), pitch st.size.x; one buffer, persists across frames
  float4* spCont;                         // spatial reuse: what k_direct_spatial needs of a pixel's State, 3 planes of pitch*allocH
  float4* dirA; float4* dirB; float4* indA; float4* indB;
  const float4* indIn;                    // what the indirect denoiser / compose read as the pre-denoise indirect image: indA, or (stage pipeline, post ranks) the
                                          // per-parity buffer the indirect ranks write into over NVLink
  float4* geomPos; float4* geomNrm;       // denoiser geometry planes (full res): pos.xyz + hash bits / normal.xyz
  float4* geomPosH; float4* geomNrmH;     // same at quarter res (pitch/2), see k_denoise_prep
  EnvView env;                            // HDR lat-long map + alias table, or the constant environment
  int hasNonOpaque;                       // scene has alpha MASK / BLEND instances: ray queries run the stochastic HitTest loop
  int pitch, allocH;                      // allocation size of the 2-D images
  // rows owned by this rank: stripes k = 0..sCount-1 of sRows full-res rows starting at sFirst + k*sStride (all multiples of 16,
  // so no 8x8 quarter-res tile straddles two ranks).  Single GPU: one stripe covering the frame.
  int sFirst, sStride, sRows, sCount;
  WaveView wv;
  unsigned long long* counters;           // per frame (one set per ping-pong parity, so that frames in flight do not mix): [0] closest-hit
                                          // rays, [1] any-hit rays, [2] primary hits, [3] inner-node visits, [4] triangle tests (STATS kernels only)
  unsigned long long* totals;             // since creation: [5] closest, [6] any, [7] worst thread's node visits, [8] / [9] longest queued closest-hit / any-hit ray
  // What indirect_stage needs of LAST frame's G-buffer and of this frame's motion image (findTemporalNeighbor, indirect_stage.comp:74-108),
  // gathered by direct_stage for the pixels 2 * coord: the quarter-res stage then reads neither image, so the NEXT frame's direct_stage may
  // overwrite them while this frame's indirect_stage is still running (frames in flight, eid_renderer_set_pipeline)
  uint4* k2G; short2* k2Mv;
};
#define EID_NUM_COUNTERS 8            // per set; device layout: set 0 | set 1 | totals (EID_NUM_TOTALS entries)
#define EID_NUM_TOTALS 10

struct RayCounters { unsigned int closest, any, primary, nodes, tris; };

// blockIdx.y (blocks of `bh` rows) -> image row for a stripe layout given in the kernel's own resolution; rows >= limit are culled by the caller
DEV int stripeRow(int first, int stride, int rows, int bh) {
  const int bps = (rows + bh - 1) / bh;                  // blocks per stripe
  const int k = blockIdx.y / bps, j = blockIdx.y - k * bps;
  const int r = j * bh + threadIdx.y;
  return (r < rows) ? first + k * stride + r : 0x3fffffff;
}

// The quarter-res stage works in 8 x 8 tiles that share one multibounce lottery draw (indirect_stage.comp:283-288), keyed by the ABSOLUTE tile
// origin.  A band may start in the middle of a tile (bands are multiples of 8 full-res = 4 quarter-res rows): blocks are laid over absolute
// tile rows from the one containing `first`, and rows outside [first, first + rows) are masked (returned as 0x3fffffff).
DEV int tileAlignedRow(int first, int stride, int rows) {
  const int bps = ((first & 7) + rows + 7) >> 3;         // blocks per stripe (stride is a multiple of 8, so every stripe has the same phase)
  const int k = blockIdx.y / bps, j = blockIdx.y - k * bps;
  const int base = first + k * stride;
  const int y = (base & ~7) + j * 8 + (int)threadIdx.y;
  return (y >= base && y < base + rows) ? y : 0x3fffffff;
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
    if (a) { atomicAdd(&P.counters[0], (unsigned long long)a); atomicAdd(&P.totals[5], (unsigned long long)a); }
    if (b) { atomicAdd(&P.counters[1], (unsigned long long)b); atomicAdd(&P.totals[6], (unsigned long long)b); }
    if (d) atomicAdd(&P.counters[2], (unsigned long long)d);
    if (STATS) { if (n) atomicAdd(&P.counters[3], (unsigned long long)n); if (t) atomicAdd(&P.counters[4], (unsigned long long)t); }
  }
  if (STATS) atomicMax(&P.totals[7], (unsigned long long)c.nodes);   // worst thread (all its rays) since the renderer was created
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
  const bool two = P.accel.twoLevel != 0;          // instanced scenes: BLAS per prim mesh + TLAS (trace.cuh: traverse2), same hits
  if (!(two ? traverse2<false, STATS>(P.accel, o, d, tmax, h, &rc.nodes, &rc.tris) : traverse<false, STATS>(P.accel, o, d, tmax, h, &rc.nodes, &rc.tris))) return false;
  while (!(h.flags & INST_FORCE_OPAQUE)) {
    if (hitTest(P, h, seed)) return true;
    const HitKey low = {h.t, h.inst, h.prim};
    if (!(two ? traverse2<false, STATS, true>(P.accel, o, d, tmax, h, &rc.nodes, &rc.tris, low) : traverse<false, STATS, true>(P.accel, o, d, tmax, h, &rc.nodes, &rc.tris, low))) return false;
  }
  return true;
}

// ClosestHit (traceray_rq.glsl:108-147).  FULL = the scene has non-opaque instances (alpha MASK / BLEND)
template <bool STATS, bool FULL>
DEV bool closestHit(const FrameParams& P, f3 o, f3 d, Payload& prd, uint32_t& seed, RayCounters& rc) {
  rc.closest++;
  RayHit h;
  const bool hit = (FULL && P.hasNonOpaque) ? firstAcceptedHit<STATS>(P, o, d, EID_INFINITY, seed, h, rc)
                   : (FULL && P.accel.twoLevel) ? traverse2<false, STATS>(P.accel, o, d, EID_INFINITY, h, &rc.nodes, &rc.tris)
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
  if (FULL && P.accel.twoLevel) return traverse2<true, STATS>(P.accel, origin, dir, tmax, h, &rc.nodes, &rc.tris);
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

}  // namespace eid

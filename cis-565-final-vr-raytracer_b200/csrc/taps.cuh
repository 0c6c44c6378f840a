// taps.cuh — parity taps: device-side shader functions exposed one call per item (tests only reach them through the C-ABI).
#pragma once
#include "frame.cuh"

namespace eid {

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

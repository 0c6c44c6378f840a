/*
 * oracle/ref_shim/ref_glsl.cpp — TEST INFRASTRUCTURE.
 * Compiles the reference's pure-arithmetic GLSL include files — globals, random, common (selected functions), pbr_metallicworkflow,
 * sun_and_sky, reservoir, tonemapping — as C++, from the transliterations glsl_prep.py writes to oracle/_ref/gen/ at build time
 * (qualifiers, literal suffixes, swizzle calls, built-in names; the expressions are the reference's own text), together with
 * shaders/host_device.h where it lies.  Exposed to ctypes as ref_fn(which, in, n, out) so that the oracle's restatement of the same
 * functions can be compared bit for bit (tests/test_oracle_kat.py) and golden vectors can be generated (tests/golden/make_golden.py).
 */
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include "glsl/glsl_builtins.h"
#include "host_device.h"          // /root/reference/shaders/host_device.h, C++ branch (nvmath -> glsl/nvmath/nvmath.h)
#undef M_PI
#undef M_PI_2
#undef M_PI_4
#undef INFINITY
#undef PI

namespace refglsl {
using orc::vec2; using orc::vec3; using orc::vec4; using orc::ivec2;
#include "../_ref/gen/globals.hpp"
#include "../_ref/gen/random.hpp"
#include "../_ref/gen/common.hpp"
#include "../_ref/gen/pbr.hpp"
#include "../_ref/gen/sun_and_sky.hpp"   // from here on M_PI is sun_and_sky.glsl's macro (3.1415926535f), as in the shader build
#include "../_ref/gen/reservoir.hpp"
#define TONEMAP_UNCHARTED                   // post.frag:30 defines it before including tonemapping.glsl
#include "../_ref/gen/tonemapping.hpp"

// ---- what layouts.glsl binds (descriptor sets, push constant, UBOs): plain globals set through ref_scene_set ----------------------
RtxState rtxState; SceneCamera sceneCamera; SunAndSky _sunAndSky; LightBufInfo lightBufInfo;
const GltfShadeMaterial* materials; const TrigLight* trigLights; const PuncLight* puncLights; const ImptSampData* envSamplingData;
PtPayload prd;
// samplers are NOT the reference's arithmetic (fixed-function hardware): texture() of the environment map goes through a function the
// test installs (the contract's bilinear sampler); material textures are not bound in these tests
typedef void (*SamplerFn)(void* obj, int index, const float* uv, int n, float* rgba);   // index < 0: the environment map
struct sampler2D { SamplerFn fn; void* obj; int index; unsigned int width, height; };
struct uvec2 { unsigned int x, y; };
sampler2D environmentTexture; sampler2D texturesMap[1];
#define nonuniformEXT(x) (x)
static uvec2 textureSize(const sampler2D& s, int) { return uvec2{s.width, s.height}; }
static vec4 texture(const sampler2D& s, vec2 uv) { float in[2] = {uv.x, uv.y}, o[4] = {0, 0, 0, 1}; if (s.fn && s.obj) s.fn(s.obj, s.index, in, 1, o); return vec4(o[0], o[1], o[2], o[3]); }
static vec4 textureLod(const sampler2D& s, vec2 uv, float) { return texture(s, uv); }
#include "../_ref/gen/gltf_material.hpp"   // SRGBtoLINEAR only
#include "../_ref/gen/env_sampling.hpp"    // Environment_sample, EnvSample
#include "../_ref/gen/pathtrace.hpp"       // EnvRadiance, EnvPdf, EnvEval, LightEval, SampleTriangleLight, SamplePuncLight, SampleDirectLightNoVisibility, clampRadiance, raySpawn

static State mkState(const float* p) {   // albedo.xyz, roughness, metallic
  State s{};
  s.mat.albedo = vec3(p[0], p[1], p[2]); s.mat.roughness = p[3]; s.mat.metallic = p[4];
  return s;
}
static vec3 v3(const float* p) { return vec3(p[0], p[1], p[2]); }
static void put(float* o, vec3 v) { o[0] = v.x; o[1] = v.y; o[2] = v.z; }
static uint32_t bitsOf(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static float floatOf(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static DirectReservoir getD(const float* p) {
  DirectReservoir r{}; r.lightSample.Li = v3(p); r.lightSample.wi = v3(p + 3); r.lightSample.dist = p[6]; r.num = bitsOf(p[7]); r.weight = p[8]; return r;
}
static void putD(float* o, const DirectReservoir& r) {
  put(o, r.lightSample.Li); put(o + 3, r.lightSample.wi); o[6] = r.lightSample.dist; o[7] = floatOf(r.num); o[8] = r.weight;
}
static GISample getG(const float* p) { GISample g{}; g.L = v3(p); g.xv = v3(p + 3); g.nv = v3(p + 6); g.xs = v3(p + 9); g.ns = v3(p + 12); g.pHat = p[15]; return g; }
static IndirectReservoir getI(const float* p) { IndirectReservoir r{}; r.giSample = getG(p); r.num = bitsOf(p[16]); r.weight = p[17]; r.bigW = p[18]; return r; }
static void putI(float* o, const IndirectReservoir& r) {
  put(o, r.giSample.L); put(o + 3, r.giSample.xv); put(o + 6, r.giSample.nv); put(o + 9, r.giSample.xs); put(o + 12, r.giSample.ns); o[15] = r.giSample.pHat;
  o[16] = floatOf(r.num); o[17] = r.weight; o[18] = r.bigW;
}
}  // namespace refglsl

using namespace refglsl;
#define REF_API extern "C" __attribute__((visibility("default")))

// number of input / output floats per item of function `which` (same table as orc_fn_arity in oracle_shaders.cpp)
REF_API int ref_fn_arity(int which, int* nin, int* nout) {
  static const int A[][2] = {{2, 2}, {2, 1}, {3, 2}, {3, 6}, {3, 3}, {3, 3}, {14, 3}, {14, 1}, {14, 7}, {30, 27}, {39, 19}, {4, 3}, {6, 3}, {2, 1}, {2, 3}};
  if (which < 0 || which >= (int)(sizeof(A) / sizeof(A[0]))) return -1;
  *nin = A[which][0]; *nout = A[which][1];
  return 0;
}

REF_API int ref_fn(int which, const float* in, int n, float* out) {
  int ni, no;
  if (ref_fn_arity(which, &ni, &no)) return -1;
  for (int i = 0; i < n; ++i) {
    const float* p = in + (size_t)i * ni;
    float* o = out + (size_t)i * no;
    switch (which) {
      case 0: { vec2 d = toConcentricDisk(vec2(p[0], p[1])); o[0] = d.x; o[1] = d.y; break; }
      case 1: o[0] = powerHeuristic(p[0], p[1]); break;
      case 2: { vec2 uv = GetSphericalUv(v3(p)); o[0] = uv.x; o[1] = uv.y; break; }
      case 3: { vec3 t, b; CreateCoordinateSystem(v3(p), t, b); put(o, t); put(o + 3, b); break; }
      case 4: put(o, HDRToLDR(v3(p))); break;
      case 5: put(o, LDRToHDR(v3(p))); break;
      case 6: put(o, metallicWorkflowBSDF(mkState(p), v3(p + 5), v3(p + 8), v3(p + 11))); break;
      case 7: o[0] = metallicWorkflowPdf(mkState(p), v3(p + 5), v3(p + 8), v3(p + 11)); break;
      case 8: { vec3 bsdf(0.0f), dir(0.0f); o[0] = metallicWorkflowSample(mkState(p), v3(p + 5), v3(p + 8), v3(p + 11), bsdf, dir); put(o + 1, bsdf); put(o + 4, dir); break; }
      case 9: {   // DirectReservoir: update(sample, w, r) | merge(rhs, r2) | checkValidity + clamp(c)
        DirectReservoir r = getD(p);
        LightSample s{}; s.Li = v3(p + 9); s.wi = v3(p + 12); s.dist = p[15];
        resvUpdate(r, s, p[16], p[17]); putD(o, r);
        resvMerge(r, getD(p + 18), p[27]); putD(o + 9, r);
        resvCheckValidity(r); resvClamp(r, (int)p[28]); putD(o + 18, r);
        (void)p[29];
        break;
      }
      case 10: {  // IndirectReservoir: update(sample, w, r), checkValidity, clamp(c)
        IndirectReservoir r = getI(p);
        resvUpdate(r, getG(p + 19), p[35], p[36]); resvCheckValidity(r); resvClamp(r, (int)p[37]); putI(o, r);
        (void)p[38];
        break;
      }
      case 11: put(o, toneMap(v3(p), p[3])); break;
      case 12: put(o, OffsetRay(v3(p), v3(p + 3))); break;
      case 13: o[0] = floatOf(tea(bitsOf(p[0]), bitsOf(p[1]))); break;
      case 14: { uint s = bitsOf(p[0]); float a = rand(s); float b = rand(s); o[0] = a; o[1] = b; o[2] = floatOf(s); (void)p[1]; break; }
    }
  }
  return 0;
}

REF_API void ref_scene_set(const RtxState* st, const SceneCamera* cam, const SunAndSky* ss, const LightBufInfo* lbi, const GltfShadeMaterial* mats,
                           const TrigLight* trig, const PuncLight* punc, const ImptSampData* envAccel, void* envSamplerFn, void* env, uint32_t envW, uint32_t envH) {
  rtxState = *st; sceneCamera = *cam; _sunAndSky = *ss; lightBufInfo = *lbi; materials = mats; trigLights = trig; puncLights = punc; envSamplingData = envAccel;
  environmentTexture = sampler2D{(SamplerFn)envSamplerFn, env, -1, envW, envH};
}
// scene-dependent functions (same numbering as orc_ctx_fn): 0 SampleDirectLightNoVisibility (seed, pos -> pdf, Li, wi, dist, seed'),
// 1 LightEval (matID, dist, dir, ffnormal, area -> Li, pdf), 2 EnvEval (dir -> radiance, pdf), 3 EnvRadiance, 4 raySpawn (coord, size ->
// origin, direction), 5 clampRadiance, 6 Sample (seed, albedo, roughness, metallic, V, N -> bsdf, L, pdf, seed')
REF_API int ref_ctx_fn(int which, const float* in, int n, float* out) {
  static const int A[][2] = {{4, 9}, {9, 4}, {3, 4}, {3, 3}, {4, 6}, {3, 3}, {12, 8}};
  if (which < 0 || which >= 7) return -1;
  const int ni = A[which][0], no = A[which][1];
  for (int i = 0; i < n; ++i) {
    const float* p = in + (size_t)i * ni;
    float* o = out + (size_t)i * no;
    switch (which) {
      case 6: {   // Sample (pathtrace.glsl:36-38): seed, albedo, roughness, metallic, V, N -> bsdf, L, pdf, seed'
        uint seed = bitsOf(p[0]); State s = mkState(p + 1); vec3 L(0.0f); float pdf = 0.0f;
        put(o, Sample(s, v3(p + 6), v3(p + 9), L, pdf, seed)); put(o + 3, L); o[6] = pdf; o[7] = floatOf(seed);
        break;
      }
      case 0: { prd.seed = bitsOf(p[0]); LightSample ls{}; o[0] = SampleDirectLightNoVisibility(v3(p + 1), ls); put(o + 1, ls.Li); put(o + 4, ls.wi); o[7] = ls.dist; o[8] = floatOf(prd.seed); break; }
      case 1: { State s{}; s.matID = bitsOf(p[0]); s.ffnormal = v3(p + 5); s.area = p[8]; float pdf = 0.0f; put(o, LightEval(s, p[1], v3(p + 2), pdf)); o[3] = pdf; break; }
      case 2: { float pdf = 0.0f; put(o, EnvEval(v3(p), pdf)); o[3] = pdf; break; }
      case 3: put(o, EnvRadiance(v3(p))); break;
      case 4: { Ray r = raySpawn(ivec2((int)p[0], (int)p[1]), ivec2((int)p[2], (int)p[3])); put(o, r.origin); put(o + 3, r.direction); break; }
      case 5: put(o, clampRadiance(v3(p))); break;
    }
  }
  return 0;
}

REF_API void ref_sun_and_sky(const SunAndSky* ss, const float* dirs, int n, float* out) {
  for (int i = 0; i < n; ++i) put(out + 3 * i, sun_and_sky(*ss, v3(dirs + 3 * i)));
}

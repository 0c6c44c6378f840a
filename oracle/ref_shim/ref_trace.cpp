/*
 * oracle/ref_shim/ref_trace.cpp — TEST INFRASTRUCTURE.
 * The reference's two trace stages — shaders/direct_stage.comp and indirect_stage.comp with everything they include (globals, random,
 * common, pathtrace, pbr_metallicworkflow, gltf_material, env_sampling, sun_and_sky, shade_state, reservoir, compress) — compiled WHOLE,
 * main() included, as C++ from the transliterations of glsl_prep.py, and dispatched over a frame in 8x8 work groups like
 * Renderer::run (src/renderer.cpp:163-176).  What a shader build binds (layouts.glsl) is provided as plain globals.  traceray_rq.glsl
 * (ClosestHit, AnyHit, HitTest) is the reference's text too; what stands in for the Vulkan driver is only the rayQuery*EXT built-ins
 * (ref_trace_stage.inl): candidates come from an intersector the test installs (the oracle's) in the order the contract fixes
 * (front to back by (t, instanceID, primitiveID); opaque candidates commit at once, others go through HitTest).
 * texture() / textureLod() are fixed-function hardware in the reference: they call the contract's samplers through a function pointer.
 * objectToWorld / worldToObject of a hit are what the driver would report for the instance: the test passes the scene's per-node matrices.
 * Compiled with -ftrivial-auto-var-init=zero: a GLSL local that is read before it is written (`LightSample lsample;` of a rejected
 * candidate) is undefined in the shader; the contract defines it as zero.
 */
#include <cmath>
#include <cstdint>
#include <cstring>
#include "glsl/glsl_builtins.h"
#include "host_device.h"          // /root/reference/shaders/host_device.h (C++ branch)
// REF_VARIANT: a second build of this file with the reference's compile-time switches flipped (Makefile: ref_trace_v.o) — DENOISER_DIRECT_BILATERAL
// (direct_stage.comp:284-288 stores to denoiseDirTempA) and FETCH_GEOM_CHECK_4_SUBPIXELS (indirect_stage_4sp.hpp = indirect_stage.comp without its
// own `#define FETCH_GEOM_CHECK_4_SUBPIXELS 0`, the macro is 1 on the command line).  Own namespaces, exported as ref_trace_run_variant.
#ifdef REF_VARIANT
#undef DENOISER_DIRECT_BILATERAL
#define DENOISER_DIRECT_BILATERAL 1
#define reftrace reftrace_v
#define reftrace_compress reftrace_compress_v
#define ref_trace_run ref_trace_run_variant
#define RefTraceBind RefTraceBindV
#endif
namespace reftrace_compress {
#include "compress.glsl"          // /root/reference/shaders/compress.glsl
}
#undef M_PI
#undef M_PI_2
#undef M_PI_4
#undef INFINITY
#undef PI

namespace reftrace {
using orc::vec2; using orc::vec3; using orc::vec4; using orc::ivec2; using orc::uvec4;
using reftrace_compress::decompress_unit_vec; using reftrace_compress::compress_unit_vec;
static unsigned int GLSL_packUnorm4x8(vec4 v) { return reftrace_compress::packUnorm4x8(v); }   // the reference's C++ twin of the GLSL built-in (compress.glsl:60-74)

// ---- types of the shader interface -----------------------------------------------------------------------------------------------
struct uvec3 { unsigned int x, y, z; };
struct ivec4 { int x, y, z, w; ivec4() : x(0), y(0), z(0), w(0) {} ivec4(int a, int b, int c, int d) : x(a), y(b), z(c), w(d) {} ivec4(ivec2 v, int c, int d) : x(v.x), y(v.y), z(c), w(d) {} ivec2 xy() const { return ivec2(x, y); } };
struct image2D { vec4* data; int w, h, pitch; };
struct uimage2D { uvec4* data; int w, h, pitch; };
struct iimage2D { int16_t* data; int w, h, pitch; };     // RG16_SINT (renderer.hpp:94)
static bool inside(int w, int h, ivec2 c) { return c.x >= 0 && c.y >= 0 && c.x < w && c.y < h; }
static vec4 imageLoad(const image2D& im, ivec2 c) { return inside(im.w, im.h, c) ? im.data[(size_t)c.y * im.pitch + c.x] : vec4(); }
static uvec4 imageLoad(const uimage2D& im, ivec2 c) { return inside(im.w, im.h, c) ? im.data[(size_t)c.y * im.pitch + c.x] : uvec4(); }
static ivec4 imageLoad(const iimage2D& im, ivec2 c) {
  if (!inside(im.w, im.h, c)) return ivec4();
  const int16_t* p = im.data + 2 * ((size_t)c.y * im.pitch + c.x);
  return ivec4(p[0], p[1], 0, 1);
}
static void imageStore(const image2D& im, ivec2 c, vec4 v) { if (inside(im.w, im.h, c)) im.data[(size_t)c.y * im.pitch + c.x] = v; }
static void imageStore(const uimage2D& im, ivec2 c, uvec4 v) { if (inside(im.w, im.h, c)) im.data[(size_t)c.y * im.pitch + c.x] = v; }
static void imageStore(const iimage2D& im, ivec2 c, ivec4 v) {   // a 16-bit signed store saturates
  if (!inside(im.w, im.h, c)) return;
  int16_t* p = im.data + 2 * ((size_t)c.y * im.pitch + c.x);
  p[0] = (int16_t)(v.x < -32768 ? -32768 : v.x > 32767 ? 32767 : v.x); p[1] = (int16_t)(v.y < -32768 ? -32768 : v.y > 32767 ? 32767 : v.y);
}
struct Indices { const uvec3* i; explicit Indices(uint64_t a) : i(reinterpret_cast<const uvec3*>(a)) {} };
struct Vertices { const VertexAttributes* v; explicit Vertices(uint64_t a) : v(reinterpret_cast<const VertexAttributes*>(a)) {} };
typedef void (*SamplerFn)(void* obj, int index, const float* uv, int n, float* rgba);   // the contract's samplers (index < 0: environment map)
struct sampler2D { SamplerFn fn; void* obj; int index; unsigned int width, height; };
using orc::uvec2;
struct InvocationId { unsigned int x, y, z; ivec2 xy() const { return ivec2((int)x, (int)y); } };
static uvec2 textureSize(const sampler2D& s, int) { return uvec2{s.width, s.height}; }
static vec4 texture(const sampler2D& s, vec2 uv) { float in[2] = {uv.x, uv.y}, o[4] = {0, 0, 0, 1}; if (s.fn && s.obj) s.fn(s.obj, s.index, in, 1, o); return vec4(o[0], o[1], o[2], o[3]); }
static vec4 textureLod(const sampler2D& s, vec2 uv, float) { return texture(s, uv); }
#define nonuniformEXT(x) (x)
#define shared static
static void barrier() {}

// ---- layouts.glsl bindings, push constant, built-in variables ----------------------------------------------------------------------
static image2D thisDirectResultImage, thisIndirectResultImage, lastDirectResultImage, lastIndirectResultImage, denoiseDirTempA, denoiseDirTempB, denoiseIndTempA, denoiseIndTempB;
static uimage2D thisGbuffer, lastGbuffer;
static iimage2D motionVector;
static const InstanceData* geoInfo; static SceneCamera sceneCamera; static const GltfShadeMaterial* materials; static const PuncLight* puncLights;
static const TrigLight* trigLights; static LightBufInfo lightBufInfo; static sampler2D texturesMap[256]; static SunAndSky _sunAndSky;
static sampler2D environmentTexture; static const ImptSampData* envSamplingData;
static DirectReservoir *lastDirectResv, *thisDirectResv, *tempDirectResv;
static IndirectReservoir *lastIndirectResv, *thisIndirectResv, *tempIndirectResv;
static RtxState rtxState;
static InvocationId gl_GlobalInvocationID, gl_LocalInvocationID, gl_WorkGroupID;
static unsigned int gl_LocalInvocationIndex;

// ---- the intersector (stands in for the driver's ray queries of traceray_rq.glsl) -----------------------------------------------------
// next candidate strictly after (lowT, lowInst, lowPrim) in front-to-back order; rec = {t, prim, inst, customIndex, u, v, opaque} (ints as bits)
typedef int (*NextCandidateFn)(void* scene, const float* ray, int haveLow, float lowT, int lowInst, int lowPrim, float* rec);
static NextCandidateFn g_trace; static void* g_scene;
static const float* g_xforms;     // per instance: objectToWorld, worldToObject (12 + 12 floats) as the driver's ray query reports them; null = identity
static unsigned long long g_closest, g_any;
struct Ray; struct PtPayload;
}  // namespace reftrace

// the two stages are separate shader modules with the same function names: each gets its own namespace and its own copy of the includes
namespace reftrace { namespace k1 {
#include "../_ref/gen/globals.hpp"
#include "ref_trace_stage.inl"
#include "../_ref/gen/direct_stage.hpp"
} }
#undef GLOBALS_GLSL
#undef RANDOM_GLSL
#undef RAYCOMMON_GLSL
#undef PBR_METALLICWORKFLOW_GLSL
#undef GLTFMATERIAL_GLSL
#undef ENV_SAMPLING_GLSL
#undef SUN_AND_SKY_GLSL
#undef SHADE_STATE_GLSL
#undef RESERVOIR_GLSL
#undef M_PI
namespace reftrace { namespace k2 {
#include "../_ref/gen/globals.hpp"
#include "ref_trace_stage.inl"
#ifdef REF_VARIANT
#include "../_ref/gen/indirect_stage_4sp.hpp"
#else
#include "../_ref/gen/indirect_stage.hpp"
#endif
} }

#ifdef REF_VARIANT
// direct_gen.comp / direct_reuse.comp: two more shader modules of the variant build (EID_VARIANT_DIRECT_SPLIT)
#define REF_UNDEF_GUARDS
#undef GLOBALS_GLSL
#undef RANDOM_GLSL
#undef RAYCOMMON_GLSL
#undef PBR_METALLICWORKFLOW_GLSL
#undef GLTFMATERIAL_GLSL
#undef ENV_SAMPLING_GLSL
#undef SUN_AND_SKY_GLSL
#undef SHADE_STATE_GLSL
#undef RESERVOIR_GLSL
#undef M_PI
namespace reftrace { namespace kg {
#include "../_ref/gen/globals.hpp"
#include "ref_trace_stage.inl"
#include "../_ref/gen/pathtrace_gd.hpp"
#include "../_ref/gen/direct_gen.hpp"
} }
#undef GLOBALS_GLSL
#undef RANDOM_GLSL
#undef RAYCOMMON_GLSL
#undef PBR_METALLICWORKFLOW_GLSL
#undef GLTFMATERIAL_GLSL
#undef ENV_SAMPLING_GLSL
#undef SUN_AND_SKY_GLSL
#undef SHADE_STATE_GLSL
#undef RESERVOIR_GLSL
#undef M_PI
namespace reftrace { namespace kr {
#include "../_ref/gen/globals.hpp"
#include "ref_trace_stage.inl"
#include "../_ref/gen/pathtrace_gd.hpp"
#include "../_ref/gen/direct_reuse.hpp"
} }
#endif

using namespace reftrace;
struct RefTraceBind {   // everything ref_trace_bind needs, as one C struct (filled by tests/oracle_lib.py)
  const RtxState* state; const SceneCamera* camera; const SunAndSky* sunSky; const LightBufInfo* lightInfo;
  const InstanceData* geoInfo; const GltfShadeMaterial* materials; const TrigLight* trigLights; const PuncLight* puncLights;
  const ImptSampData* envAccel; void* envSamplerFn; void* env; uint32_t envW, envH;
  void* traceFn; void* scene;
  int32_t allocW, allocH;
  void *thisG, *lastG, *motion, *thisDR, *lastDR, *thisIR, *lastIR, *direct, *indirect, *indA;
  const float* instanceXforms;
  void* tempDR;      // tempDirectResv (one buffer for both descriptor sets, renderer.cpp:235, 349); persists across frames
  void* dirA;        // denoiseDirTempA: what direct_stage stores to with DENOISER_DIRECT_BILATERAL (null = not bound)
};

template <class F>
static void dispatchGroups(int w, int h, F&& mainFn) {   // vkCmdDispatch(CEIL_DIV(w, 8), CEIL_DIV(h, 8), 1): invocations of a group in order
  for (int gy = 0; gy < (h + 7) / 8; ++gy)
    for (int gx = 0; gx < (w + 7) / 8; ++gx)
      for (int ly = 0; ly < 8; ++ly)
        for (int lx = 0; lx < 8; ++lx) {
          gl_WorkGroupID = InvocationId{(unsigned)gx, (unsigned)gy, 0u}; gl_LocalInvocationID = InvocationId{(unsigned)lx, (unsigned)ly, 0u};
          gl_GlobalInvocationID = InvocationId{(unsigned)(gx * 8 + lx), (unsigned)(gy * 8 + ly), 0u}; gl_LocalInvocationIndex = (unsigned)(ly * 8 + lx);
          mainFn();
        }
}

// Renderer::run, renderer.cpp:163-176: direct_stage over the frame, indirect_stage over (W/2) x (H/2); rays[0] / rays[1] = ClosestHit / AnyHit calls
extern "C" __attribute__((visibility("default")))
void ref_trace_run(const RefTraceBind* b, int runDirect, int runIndirect, unsigned long long* rays) {
  rtxState = *b->state; sceneCamera = *b->camera; _sunAndSky = *b->sunSky; lightBufInfo = *b->lightInfo;
  geoInfo = b->geoInfo; materials = b->materials; trigLights = b->trigLights; puncLights = b->puncLights; envSamplingData = b->envAccel;
  environmentTexture = sampler2D{(SamplerFn)b->envSamplerFn, b->env, -1, b->envW, b->envH};
  for (int i = 0; i < 256; ++i) texturesMap[i] = sampler2D{(SamplerFn)b->envSamplerFn, b->scene, i, 1u, 1u};   // material textures: the scene's samplers
  g_trace = (NextCandidateFn)b->traceFn; g_scene = b->scene; g_xforms = b->instanceXforms; g_closest = g_any = 0;
  auto img = [&](void* p) { return image2D{(vec4*)p, b->allocW, b->allocH, b->allocW}; };
  thisGbuffer = uimage2D{(uvec4*)b->thisG, b->allocW, b->allocH, b->allocW}; lastGbuffer = uimage2D{(uvec4*)b->lastG, b->allocW, b->allocH, b->allocW};
  motionVector = iimage2D{(int16_t*)b->motion, b->allocW, b->allocH, b->allocW};
  thisDirectResultImage = img(b->direct); thisIndirectResultImage = img(b->indirect); denoiseIndTempA = img(b->indA);
  if (b->dirA) denoiseDirTempA = img(b->dirA);
  thisDirectResv = (DirectReservoir*)b->thisDR; lastDirectResv = (DirectReservoir*)b->lastDR;
  thisIndirectResv = (IndirectReservoir*)b->thisIR; lastIndirectResv = (IndirectReservoir*)b->lastIR;
  const int W = rtxState.size.x, H = rtxState.size.y;
  tempDirectResv = (DirectReservoir*)b->tempDR;
  if (runDirect && (rtxState.ReSTIRState == eSpatial || rtxState.ReSTIRState == eSpatiotemporal)) {
    // Spatial reuse (direct_stage.comp:224-255) reads the neighbours' tempDirectResv entries behind a barrier() that orders one work group
    // only.  The contract is its race-free reading — all writes before all reads — obtained from the unmodified text by dispatching the
    // stage twice: each invocation writes the same entry both times (nothing it writes before the barrier depends on what it reads after
    // it), so in the second dispatch every read sees the neighbour's completed entry; the first dispatch's image is overwritten.
    dispatchGroups(W, H, [] { k1::main(); });
    g_closest = g_any = 0;
  }
#ifdef REF_VARIANT
  if (runDirect == 2) { dispatchGroups(W, H, [] { kg::main(); }); dispatchGroups(W, H, [] { kr::main(); }); }   // direct_gen.comp, then direct_reuse.comp
  else
#endif
  if (runDirect) dispatchGroups(W, H, [] { k1::main(); });
  if (runIndirect) dispatchGroups(W / 2, H / 2, [] { k2::main(); });
  if (rays) { rays[0] = g_closest; rays[1] = g_any; }
}

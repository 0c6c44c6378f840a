/*
 * oracle/oracle_api.cpp — TEST INFRASTRUCTURE: C entry points of the CPU oracle for ctypes.
 * The signatures deliberately shadow include/eidola.h (orc_* instead of eid_*) so parity tests
 * drive both sides with the same code.
 */
#include "oracle.h"
#include <cstdio>
#include <cstring>
#include <omp.h>
#include <string>

using namespace orc;

#define ORC_API extern "C" __attribute__((visibility("default")))

struct orc_renderer_h { Renderer r; };

ORC_API int orc_fp_contract_selftest(void) {
  // a*b+c differs between fused and unfused evaluation for these inputs; the oracle must be unfused
  // a*a = 1 + 2^-11 + 2^-24 rounds (ties-to-even) to 1 + 2^-11, so unfused a*a + c == 0 but fma gives 2^-24
  volatile float a = 1.0f + 1.0f / 4096.0f, c = -(1.0f + 1.0f / 2048.0f);
  float r = a * a + c;
  return (r == 0.0f) ? 0 : 1;
}
ORC_API int orc_num_threads(void) { return omp_get_max_threads(); }
ORC_API void orc_set_num_threads(int n) { omp_set_num_threads(n); }

// ---- known-answer taps ----------------------------------------------------------------------------
ORC_API uint32_t orc_tea(uint32_t a, uint32_t b) { return tea(a, b); }
ORC_API void orc_rand_chain(uint32_t seed, int n, uint32_t* states, float* vals) {
  uint s = seed;
  for (int i = 0; i < n; ++i) { vals[i] = rnd(s); states[i] = s; }
}
ORC_API uint32_t orc_hash8bit(uint32_t a) { return hash8bit(a); }
ORC_API uint32_t orc_compress_unit_vec(float x, float y, float z) { return compress_unit_vec(vec3(x, y, z)); }
ORC_API void orc_decompress_unit_vec(uint32_t p, float* out) { vec3 v = decompress_unit_vec(p); out[0] = v.x; out[1] = v.y; out[2] = v.z; }
ORC_API void orc_offset_ray(const float* p, const float* n, float* out) {
  vec3 r = OffsetRay(vec3(p[0], p[1], p[2]), vec3(n[0], n[1], n[2])); out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
ORC_API int orc_fn_arity(int which, int* nin, int* nout) { return fn_arity(which, nin, nout); }
ORC_API int orc_fn(int which, const float* in, int n, float* out) { return fn(which, in, n, out); }
ORC_API int orc_ctx_fn(Renderer* r, const RtxState* st, int which, const float* in, int n, float* out) { return ctx_fn(*r, *st, which, in, n, out); }
ORC_API void orc_alias_table(const float* values, int n, float* prob, int* failId) {
  std::vector<float> v(values, values + n), p; std::vector<int> f;
  discreteSampler1D(v, p, f);
  for (int i = 0; i < n; ++i) { prob[i] = p[i]; failId[i] = f[i]; }
}
ORC_API uint32_t orc_pack_unorm4x8(const float* v) { return packUnorm4x8(vec4(v[0], v[1], v[2], v[3])); }
ORC_API int orc_sizeof(const char* name) {
  std::string s(name);
#define SZ(T) if (s == #T) return (int)sizeof(T);
  SZ(SceneCamera) SZ(VertexAttributes) SZ(GltfShadeMaterial) SZ(RtxState) SZ(InstanceData) SZ(LightSample) SZ(GISample)
  SZ(DirectReservoir) SZ(IndirectReservoir) SZ(ImptSampData) SZ(PuncLight) SZ(TrigLight) SZ(LightBufInfo) SZ(Tonemapper) SZ(SunAndSky)
#undef SZ
  return -1;
}
// deterministic-math taps: op 0 sin, 1 cos, 2 exp, 3 log, 4 pow(x, y), 5 asin, 6 acos, 7 atan2(x, y)
ORC_API void orc_detmath(int op, const float* x, const float* y, int n, float* out) {
  for (int i = 0; i < n; ++i) {
    switch (op) {
      case 0: out[i] = eid_sinf(x[i]); break;
      case 1: out[i] = eid_cosf(x[i]); break;
      case 2: out[i] = eid_expf(x[i]); break;
      case 3: out[i] = eid_logf(x[i]); break;
      case 4: out[i] = eid_powf(x[i], y[i]); break;
      case 5: out[i] = eid_asinf(x[i]); break;
      case 6: out[i] = eid_acosf(x[i]); break;
      default: out[i] = eid_atan2f(x[i], y[i]); break;
    }
  }
}

// ---- scene ----------------------------------------------------------------------------------------
ORC_API Scene* orc_scene_create(void) { return new Scene(); }
ORC_API void orc_scene_destroy(Scene* s) { delete s; }
ORC_API void orc_scene_set_use_bvh(Scene* s, int use) { s->useBvh = use != 0; }
ORC_API int orc_scene_load_desc(Scene* s, const eid_scene_desc* d) { s->load(*d); return 0; }
ORC_API int orc_scene_set_lookat(Scene* s, const float* eye, const float* center, const float* up, float fovDeg) {
  for (int i = 0; i < 3; ++i) { s->eye[i] = eye[i]; s->center[i] = center[i]; s->up[i] = up[i]; }
  s->fovDeg = fovDeg;
  return 0;
}
ORC_API int orc_scene_update_camera(Scene* s, uint32_t w, uint32_t h) { s->updateCamera(w, h); return 0; }
ORC_API int orc_scene_set_camera(Scene* s, const SceneCamera* c) { s->camera = *c; return 0; }
ORC_API int orc_scene_get_camera(Scene* s, SceneCamera* c) { *c = s->camera; return 0; }
ORC_API int orc_scene_get_info(Scene* s, eid_scene_info* o) {
  memset(o, 0, sizeof(*o));
  o->primMeshCount = (uint32_t)s->primMeshes.size(); o->nodeCount = (uint32_t)s->nodes.size();
  o->materialCount = (uint32_t)s->materials.size();
  o->puncLightCount = s->lightBufInfo.puncLightSize; o->trigLightCount = s->lightBufInfo.trigLightSize;
  o->vertexCount = (uint32_t)(s->positions.size() / 3); o->indexCount = (uint32_t)s->indices.size();
  o->triangleInstances = s->tris.size();
  o->trigLightWeight = s->trigLightWeight; o->puncLightWeight = s->puncLightWeight;
  for (int i = 0; i < 3; ++i) { o->bboxMin[i] = s->bboxMin[i]; o->bboxMax[i] = s->bboxMax[i]; }
  return 0;
}
static const void* tablePtr(Scene* s, int table, uint32_t index, size_t& bytes) {
  switch (table) {
    case EID_TABLE_MATERIALS: bytes = s->shadeMaterials.size() * sizeof(GltfShadeMaterial); return s->shadeMaterials.data();
    case EID_TABLE_PUNC_LIGHTS: bytes = s->puncLights.size() * sizeof(PuncLight); return s->puncLights.data();
    case EID_TABLE_TRIG_LIGHTS: bytes = s->trigLights.size() * sizeof(TrigLight); return s->trigLights.data();
    case EID_TABLE_LIGHT_INFO: bytes = sizeof(LightBufInfo); return &s->lightBufInfo;
    case EID_TABLE_VERTICES: if (index >= s->vertexBufs.size()) return nullptr; bytes = s->vertexBufs[index].size() * sizeof(VertexAttributes); return s->vertexBufs[index].data();
    case EID_TABLE_INDICES: if (index >= s->indexBufs.size()) return nullptr; bytes = s->indexBufs[index].size() * 4; return s->indexBufs[index].data();
    case EID_TABLE_CAMERA: bytes = sizeof(SceneCamera); return &s->camera;
    default: return nullptr;
  }
}
ORC_API int64_t orc_scene_table_bytes(Scene* s, int table, uint32_t index) {
  if (table == EID_TABLE_INSTANCE_DATA) return (int64_t)s->primMeshes.size() * sizeof(InstanceData);
  size_t b = 0; return tablePtr(s, table, index, b) ? (int64_t)b : -1;
}
ORC_API int orc_scene_read_table(Scene* s, int table, uint32_t index, void* dst, size_t bytes) {
  if (table == EID_TABLE_INSTANCE_DATA) {   // addresses are meaningless on the CPU: zero them, keep materialIndex
    std::vector<InstanceData> v(s->primMeshes.size());
    if (!v.empty()) memset(v.data(), 0, v.size() * sizeof(InstanceData));   // padding bytes too
    for (size_t i = 0; i < v.size(); ++i) { v[i].vertexAddress = 0; v[i].indexAddress = 0; v[i].materialIndex = s->instMaterial[i]; }
    if (bytes > v.size() * sizeof(InstanceData)) return -1;
    memcpy(dst, v.data(), bytes); return 0;
  }
  size_t b = 0; const void* p = tablePtr(s, table, index, b);
  if (!p || bytes > b) return -1;
  memcpy(dst, p, bytes); return 0;
}
ORC_API int orc_accel_trace(Scene* s, const float* rays, uint32_t n, int any_hit, eid_hit* hits) {
#pragma omp parallel for schedule(dynamic, 256)
  for (int64_t i = 0; i < (int64_t)n; ++i) {
    const float* r = rays + 8 * i;
    vec3 o(r[0], r[1], r[2]), d(r[4], r[5], r[6]);
    if (any_hit) {
      bool h = s->anyHit(o, d, r[3], nullptr);
      hits[i] = eid_hit{h ? 0.f : 1e28f, -1, -1, -1, 0.f, 0.f};
    } else {
      Hit h = s->closestHit(o, d, r[3], nullptr);
      hits[i] = eid_hit{h.hitT, h.primitiveID, h.instanceID, h.instanceCustomIndex, h.bary.x, h.bary.y};
    }
  }
  return 0;
}

// ---- renderer ---------------------------------------------------------------------------------------
ORC_API Renderer* orc_renderer_create(Scene* s, uint32_t w, uint32_t h) { Renderer* r = new Renderer(); r->create(s, w, h); return r; }
ORC_API void orc_renderer_destroy(Renderer* r) { delete r; }
ORC_API Environment* orc_env_create(const float* rgba, uint32_t w, uint32_t h) { Environment* e = new Environment(); e->create(rgba, w, h); return e; }
ORC_API void orc_env_destroy(Environment* e) { delete e; }
ORC_API float orc_env_integral(Environment* e) { return e->integral; }
ORC_API float orc_env_average(Environment* e) { return e->average; }
ORC_API int orc_env_read_accel(Environment* e, void* dst, size_t bytes) {
  if (bytes > e->accel.size() * sizeof(ImptSampData)) return -1;
  memcpy(dst, e->accel.data(), bytes); return 0;
}
ORC_API void orc_env_texture(Environment* e, const float* uv, int n, float* out) {
  for (int i = 0; i < n; ++i) { vec3 c = e->texture(vec2(uv[2 * i], uv[2 * i + 1])); out[3 * i] = c.x; out[3 * i + 1] = c.y; out[3 * i + 2] = c.z; }
}
// sampler tap used by oracle/ref_shim (the reference's shader text compiled as C++): index < 0: obj = Environment*, lat-long map; else
// obj = Scene*, texturesMap[index]; textureLod(..., 0) of the contract's samplers -> RGBA
ORC_API void orc_sample(void* obj, int index, const float* uv, int n, float* rgba) {
  for (int i = 0; i < n; ++i) {
    const vec2 t(uv[2 * i], uv[2 * i + 1]);
    vec4 c;
    if (index < 0) c = vec4(static_cast<Environment*>(obj)->texture(t), 1.0f);
    else {
      const Scene* s = static_cast<Scene*>(obj);
      c = (index < (int)s->textures.size()) ? s->textures[index].sample(t) : vec4(1.0f);
    }
    rgba[4 * i] = c.x; rgba[4 * i + 1] = c.y; rgba[4 * i + 2] = c.z; rgba[4 * i + 3] = c.w;
  }
}
// the driver's candidate enumeration as the contract fixes it: the first candidate strictly after `low` in
// (t, instanceID, primitiveID) order with 0 < t < tmax; rec = {hitT, primitiveID, instanceID, instanceCustomIndex, baryU, baryV, opaque}
ORC_API int orc_accel_next_candidate(Scene* s, const float* ray, int haveLow, float lowT, int lowInst, int lowPrim, float* rec) {
  const HitKey key{lowT, lowInst, lowPrim};
  const Hit h = s->closestHit(vec3(ray[0], ray[1], ray[2]), vec3(ray[4], ray[5], ray[6]), ray[3], nullptr, haveLow ? &key : nullptr);
  if (!(h.hitT < 1e28f) || h.primitiveID < 0) return 0;
  rec[0] = h.hitT; rec[1] = intBitsToFloat(h.primitiveID); rec[2] = intBitsToFloat(h.instanceID); rec[3] = intBitsToFloat(h.instanceCustomIndex);
  rec[4] = h.bary.x; rec[5] = h.bary.y; rec[6] = intBitsToFloat(h.opaque);
  return 1;
}
// per instance (= node): objectToWorld, worldToObject as the ray query reports them (mat4x3, 12 + 12 floats, column-major)
ORC_API int orc_scene_instance_xforms(Scene* s, float* out, int maxInstances) {
  const int n = (int)s->objectToWorld.size();
  for (int i = 0; i < n && i < maxInstances; ++i)
    for (int c = 0; c < 4; ++c) {
      const vec3 a = s->objectToWorld[i].c[c], b = s->worldToObject[i].c[c];
      float* o = out + 24 * i;
      o[3 * c] = a.x; o[3 * c + 1] = a.y; o[3 * c + 2] = a.z; o[12 + 3 * c] = b.x; o[12 + 3 * c + 1] = b.y; o[12 + 3 * c + 2] = b.z;
    }
  return n;
}
// per instance (= node) what accelstruct.cpp:132-162 decides: (VkGeometryInstanceFlagsKHR bits: 4 FORCE_OPAQUE, 1 TRIANGLE_FACING_CULL_DISABLE;
// instanceCustomIndex; triangles) — read back from the triangle soup the intersector actually walks
ORC_API int orc_scene_instance_flags(Scene* s, int* out, int maxInstances) {
  const int n = (int)s->objectToWorld.size();
  for (int i = 0; i < n && i < maxInstances; ++i) { out[3 * i] = -1; out[3 * i + 1] = -1; out[3 * i + 2] = 0; }
  for (const auto& t : s->tris) {
    if (t.inst < 0 || t.inst >= n || t.inst >= maxInstances) continue;
    out[3 * t.inst] = (t.opaque ? 4 : 0) | (t.cullDisable ? 1 : 0); out[3 * t.inst + 1] = t.customIndex; out[3 * t.inst + 2] += 1;
  }
  return n;
}
ORC_API int orc_scene_texture_count(Scene* s) { return (int)s->textures.size(); }
ORC_API int orc_renderer_set_env(Renderer* r, Environment* e) { r->env = e; return 0; }
// RenderOutput::run over the frame rendered last: out = width*height RGBA32F at the allocation pitch
ORC_API int orc_renderer_run_output(Renderer* r, const Tonemapper* tm, const RtxState* st, float* out) {
  const int W = st->size.x, H = st->size.y;
  vec4 avgD, avgI;
  if (tm->autoExposure & 1) {   // genMipmap runs over the whole images (m_size = the allocation), render_output.cpp:243-253
    avgD = mip_chain_average(r->directResult.data(), (int)r->width, (int)r->height, (int)r->width);
    avgI = mip_chain_average(r->indirectResult.data(), (int)r->width, (int)r->height, (int)r->width);
  }
#pragma omp parallel for schedule(static)
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      const size_t pix = (size_t)y * r->width + x;
      const vec4 c = post_frag(*tm, st->debugging_mode, r->directResult[pix], r->indirectResult[pix], x, y, W, H, avgD, avgI);
      out[4 * pix] = c.x; out[4 * pix + 1] = c.y; out[4 * pix + 2] = c.z; out[4 * pix + 3] = c.w;
    }
  return 0;
}
// known-answer taps of the auto-exposure path: the 1x1 mip level of an RGBA32F image, and toneExposure (post.frag:65-70) per item
ORC_API void orc_mip_chain_average(const float* rgba, int w, int h, float* out4) {
  const vec4 a = mip_chain_average(reinterpret_cast<const vec4*>(rgba), w, h, w);
  out4[0] = a.x; out4[1] = a.y; out4[2] = a.z; out4[3] = a.w;
}
ORC_API void orc_tone_exposure(const Tonemapper* tm, const float* rgbAndLum, int n, float* out) {
  for (int i = 0; i < n; ++i) { const vec3 c = post_toneExposure(*tm, vec3(rgbAndLum[4 * i], rgbAndLum[4 * i + 1], rgbAndLum[4 * i + 2]), rgbAndLum[4 * i + 3]); out[3 * i] = c.x; out[3 * i + 1] = c.y; out[3 * i + 2] = c.z; }
}
ORC_API int orc_renderer_set_sun_and_sky(Renderer* r, const SunAndSky* ss) { r->sunSky = *ss; return 0; }
ORC_API int orc_renderer_set_variant(Renderer* r, int flags) { r->variant = flags; return 0; }
ORC_API void orc_sun_and_sky(const SunAndSky* ss, const float* dirs, int n, float* out) {   // known-answer tap
  for (int i = 0; i < n; ++i) { vec3 c = sun_and_sky(*ss, vec3(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2])); out[3 * i] = c.x; out[3 * i + 1] = c.y; out[3 * i + 2] = c.z; }
}
ORC_API int orc_renderer_set_env_constant(Renderer* r, const float* rgb) { r->envConstant = vec3(rgb[0], rgb[1], rgb[2]); return 0; }
static RtxState g_lastState{};
ORC_API int orc_renderer_run(Renderer* r, const RtxState* st, int frames) {
  r->closestRays = 0; r->anyRays = 0; r->primaryHits = 0; for (auto& k : r->kernelMs) k = 0;
  g_lastState = *st; r->run(*st, frames); return 0;
}
ORC_API int orc_renderer_run_trace(Renderer* r, const RtxState* st, int frames, int y0, int y1) {
  r->closestRays = 0; r->anyRays = 0; r->primaryHits = 0; for (auto& k : r->kernelMs) k = 0;
  g_lastState = *st; r->lastSet = (frames + 1) % 2;
  r->runDirect(*st, frames, y0, y1); r->runIndirect(*st, frames, y0 / 2, y1 / 2); return 0;
}
ORC_API int orc_renderer_run_post(Renderer* r, const RtxState* st, int frames) { r->runPost(*st, frames); return 0; }
ORC_API int orc_renderer_get_stats(Renderer* r, eid_frame_stats* o) {
  memset(o, 0, sizeof(*o));
  o->closestHitRays = r->closestRays; o->anyHitRays = r->anyRays; o->primaryHits = r->primaryHits;
  for (int i = 0; i < EID_K_COUNT; ++i) o->kernelMs[i] = (float)r->kernelMs[i];
  return 0;
}
static void* bufPtr(Renderer* r, int which, size_t& bytes) {
  int set = r->lastSet;   // this* = [!set], last* = [set]
  switch (which) {
    case EID_BUF_THIS_GBUFFER: bytes = r->gbuffer[!set].size() * 16; return r->gbuffer[!set].data();
    case EID_BUF_LAST_GBUFFER: bytes = r->gbuffer[set].size() * 16; return r->gbuffer[set].data();
    case EID_BUF_MOTION: bytes = r->motion.size() * 2; return r->motion.data();
    case EID_BUF_THIS_DIRECT_RESV: bytes = r->directResv[!set].size() * sizeof(DirectReservoir); return r->directResv[!set].data();
    case EID_BUF_LAST_DIRECT_RESV: bytes = r->directResv[set].size() * sizeof(DirectReservoir); return r->directResv[set].data();
    case EID_BUF_THIS_INDIRECT_RESV: bytes = r->indirectResv[!set].size() * sizeof(IndirectReservoir); return r->indirectResv[!set].data();
    case EID_BUF_LAST_INDIRECT_RESV: bytes = r->indirectResv[set].size() * sizeof(IndirectReservoir); return r->indirectResv[set].data();
    case EID_BUF_DIRECT: bytes = r->directResult.size() * 16; return r->directResult.data();
    case EID_BUF_INDIRECT: bytes = r->indirectResult.size() * 16; return r->indirectResult.data();
    case EID_BUF_DENOISE_DIR_A: case EID_BUF_DENOISE_DIR_B: case EID_BUF_DENOISE_IND_A: case EID_BUF_DENOISE_IND_B:
      bytes = r->denoiseTemp[which - EID_BUF_DENOISE_DIR_A].size() * 16; return r->denoiseTemp[which - EID_BUF_DENOISE_DIR_A].data();
    case EID_BUF_TEMP_DIRECT_RESV: bytes = r->tempDirectResv.size() * sizeof(DirectReservoir); return r->tempDirectResv.data();
    default: return nullptr;
  }
}
ORC_API int64_t orc_renderer_buffer_bytes(Renderer* r, int which) { size_t b = 0; return bufPtr(r, which, b) ? (int64_t)b : -1; }
ORC_API int orc_renderer_read(Renderer* r, int which, void* dst, size_t bytes) {
  size_t b = 0; void* p = bufPtr(r, which, b);
  if (!p || bytes > b) return -1;
  memcpy(dst, p, bytes); return 0;
}
ORC_API int orc_renderer_write(Renderer* r, int which, const void* src, size_t bytes) {
  size_t b = 0; void* p = bufPtr(r, which, b);
  if (!p || bytes > b) return -1;
  memcpy(p, src, bytes); return 0;
}

/*
 * oracle/oracle_shaders.cpp — TEST INFRASTRUCTURE (CPU oracle).
 *
 * Restatement of the reference's live device code as ordinary C++, one function per GLSL
 * function, in GLSL evaluation order (arguments left to right), so RNG draw order and
 * floating-point rounding are pinned.  Files followed:
 *   random.glsl, common.glsl, compress.glsl, shade_state.glsl (GetState), gltf_material.glsl,
 *   pbr_metallicworkflow.glsl, reservoir.glsl, pathtrace.glsl, direct_stage.comp,
 *   indirect_stage.comp, denoise_common.glsl, denoise_direct.comp, denoise_indirect.comp,
 *   compose.comp, and Renderer::run (renderer.cpp:154-206, 341-375).
 *
 * Scope of round 1 (documented in DESIGN.md): texture-less materials, opaque geometry,
 * constant environment radiance, ReSTIRState in {eNone, eRIS, eTemporal}.
 */
#include "oracle.h"
#include <chrono>
#include <cstdio>

namespace orc {

static const float INFINITY_ = 1e28f;   // globals.glsl:29
static const float EPS = 0.0001f;       // globals.glsl:30
static const float M_PI_F = 3.14159265358979323846f;
static const float InvalidPdf = -1.0f;  // common.glsl:30
static const uint InvalidMatId = 0xff000000u;   // globals.glsl:106

// ---- random.glsl:34-48 -------------------------------------------------------------------------
uint tea(uint val0, uint val1) {
  uint v0 = val0, v1 = val1, s0 = 0;
  for (uint n = 0; n < 16; n++) {
    s0 += 0x9e3779b9;
    v0 += ((v1 << 4) + 0xa341316c) ^ (v1 + s0) ^ ((v1 >> 5) + 0xc8013ea4);
    v1 += ((v0 << 4) + 0xad90777d) ^ (v0 + s0) ^ ((v0 >> 5) + 0x7e95761e);
  }
  return v0;
}
// ---- random.glsl:59-65 -------------------------------------------------------------------------
uint pcg(uint& state) {
  uint prev = state * 747796405u + 2891336453u;
  uint word = ((prev >> ((prev >> 28u) + 4u)) ^ prev) * 277803737u;
  state = prev;
  return (word >> 22u) ^ word;
}
// ---- random.glsl:98-102 ------------------------------------------------------------------------
float rnd(uint& seed) {
  uint r = pcg(seed);
  return uintBitsToFloat(0x3f800000 | (r >> 9)) - 1.0f;
}
// ---- common.glsl:141-143 -----------------------------------------------------------------------
uint hash8bit(uint a) { return (a ^ (a >> 8)) << 24; }

// ---- sun_and_sky.glsl:31-34 / denoise_common.glsl:23-25 ----------------------------------------
static inline float luminance(vec3 rgb) { return (0.2126f * rgb.x + 0.7152f * rgb.y) + 0.0722f * rgb.z; }

// ---- common.glsl:98-113 ------------------------------------------------------------------------
vec3 OffsetRay(vec3 p, vec3 n) {
  const float intScale = 256.0f;
  const float floatScale = 1.0f / 65536.0f;
  const float origin = 1.0f / 32.0f;
  int ofx = f2i(intScale * n.x), ofy = f2i(intScale * n.y), ofz = f2i(intScale * n.z);
  vec3 p_i(intBitsToFloat(floatBitsToInt(p.x) + ((p.x < 0) ? -ofx : ofx)),
           intBitsToFloat(floatBitsToInt(p.y) + ((p.y < 0) ? -ofy : ofy)),
           intBitsToFloat(floatBitsToInt(p.z) + ((p.z < 0) ? -ofz : ofz)));
  return vec3(gabs(p.x) < origin ? p.x + floatScale * n.x : p_i.x,
              gabs(p.y) < origin ? p.y + floatScale * n.y : p_i.y,
              gabs(p.z) < origin ? p.z + floatScale * n.z : p_i.z);
}

// ---- common.glsl:171-200 -----------------------------------------------------------------------
static vec2 toConcentricDisk(vec2 r) {
  float rx = sqrtf(r.x);
  float theta = r.y * 2.0f * M_PI_F;
  float s, c;
  eid_sincosf(theta, &s, &c);
  return vec2(c, s) * rx;
}
static float powerHeuristic(float f, float g) { float f2 = f * f; return f2 / (f2 + g * g); }
static bool hasNan(vec3 v) { return gisnan(v.x) || gisnan(v.y) || gisnan(v.z); }
static bool inBound(ivec2 p, ivec2 pMin, ivec2 pMax) { return p.x >= pMin.x && p.x < pMax.x && p.y >= pMin.y && p.y < pMax.y; }
static bool inBound(ivec2 p, ivec2 bound) { return inBound(p, ivec2(0, 0), bound); }
static vec3 HDRToLDR(vec3 color) { return color / (color + 1.0f); }
static vec3 LDRToHDR(vec3 color) { return color / (1.01f - color); }

// ---- globals.glsl:48-104 -----------------------------------------------------------------------
struct Ray { vec3 origin, direction; };
struct PtPayload {
  uint seed; float hitT; int primitiveID, instanceID, instanceCustomIndex; vec2 baryCoord;
  mat4x3 objectToWorld, worldToObject;
};
struct Material { vec3 albedo, emission; float metallic = 0, ior = 0, roughness = 0, transmission = 0; };
struct State {
  int depth = 0; float eta = 0;
  vec3 position, normal, tangent, bitangent, ffnormal; vec2 texCoord;
  bool isEmitter = false; uint matID = 0; Material mat; float area = 0;
};

// ---- pbr_metallicworkflow.glsl -----------------------------------------------------------------
static const float Pi = M_PI_F;
static const float PiInv = 1.0f / M_PI_F;

static mat3 localRefMatrix(vec3 n) {                                   // :11-16
  vec3 t = (gabs(n.y) > 0.9999f) ? vec3(0.0f, 0.0f, 1.0f) : vec3(0.0f, 1.0f, 0.0f);
  vec3 b = normalize(cross(n, t));
  t = cross(b, n);
  return mat3(t, b, n);
}
static vec3 localToWorld(vec3 n, vec3 v) { return normalize(localRefMatrix(n) * v); }   // :18-20
static vec3 sampleHemisphereCosine(vec3 n, vec2 r) {                   // :22-26
  vec2 d = toConcentricDisk(r);
  float z = sqrtf(1.0f - dot(d, d));
  return localToWorld(n, vec3(d, z));
}
static float satDot(vec3 a, vec3 b) { return gmax(dot(a, b), 0.0f); }
static float absDot(vec3 a, vec3 b) { return gabs(dot(a, b)); }
static vec3 FresnelSchlick(float cosTheta, vec3 f0) {                  // :36-41
  float cos4 = 1.0f - cosTheta;
  cos4 *= cos4;
  cos4 *= cos4;
  return mix(f0, vec3(1.0f), cos4 * (1.0f - cosTheta));
}
static float SchlickG(float cosTheta, float alpha) { float a = alpha * 0.5f; return cosTheta / (cosTheta * (1.0f - a) + a); }
static float SmithG(float cosWo, float cosWi, float alpha) { return SchlickG(gabs(cosWo), alpha) * SchlickG(gabs(cosWi), alpha); }
static float GTR2Distrib(float cosTheta, float alpha) {                // :52-61
  if (cosTheta < 1e-6f) return 0.0f;
  float aa = alpha * alpha;
  float nom = aa;
  float denom = cosTheta * cosTheta * (aa - 1.0f) + 1.0f;
  denom = denom * denom * Pi;
  return nom / denom;
}
static float GTR2Pdf(vec3 n, vec3 m, vec3 wo, float alpha) {           // :63-65
  return GTR2Distrib(dot(n, m), alpha) * SchlickG(dot(n, wo), alpha) * absDot(m, wo) / absDot(n, wo);
}
static vec3 GTR2Sample(vec3 n, vec3 wo, float alpha, vec2 r) {         // :67-84
  mat3 transMat = localRefMatrix(n);
  mat3 transInv = inverse(transMat);
  vec3 vh = normalize((transInv * wo) * vec3(alpha, alpha, 1.0f));
  float lenSq = vh.x * vh.x + vh.y * vh.y;
  vec3 t = lenSq > 0.0f ? vec3(-vh.y, vh.x, 0.0f) / sqrtf(lenSq) : vec3(1.0f, 0.0f, 0.0f);
  vec3 b = cross(vh, t);
  vec2 p = toConcentricDisk(r);
  float s = 0.5f * (vh.z + 1.0f);
  p.y = (1.0f - s) * sqrtf(1.0f - p.x * p.x) + s * p.y;
  vec3 h = (t * p.x + b * p.y) + vh * sqrtf(gmax(0.0f, 1.0f - dot(p, p)));
  h = vec3(h.x * alpha, h.y * alpha, gmax(0.0f, h.z));
  return normalize(transMat * h);
}
static vec3 metallicWorkflowBSDF(const State& state, vec3 n, vec3 wo, vec3 wi) {   // :86-106
  vec3 baseColor = state.mat.albedo;
  float roughness = state.mat.roughness;
  float metallic = state.mat.metallic;
  float alpha = roughness;
  vec3 h = normalize(wo + wi);
  float cosO = dot(n, wo);
  float cosI = dot(n, wi);
  if (cosI * cosO < 1e-7f) return vec3(0.0f);
  vec3 f = FresnelSchlick(dot(h, wo), mix(vec3(.08f), baseColor, metallic));
  float g = SmithG(cosO, cosI, alpha);
  float d = GTR2Distrib(dot(n, h), alpha);
  return mix(baseColor * PiInv * (1.0f - metallic), vec3(g * d / (4.0f * cosI * cosO)), f);
}
static float metallicWorkflowPdf(const State& state, vec3 n, vec3 wo, vec3 wi) {   // :108-121
  float roughness = state.mat.roughness;
  float metallic = state.mat.metallic;
  float alpha = roughness;
  vec3 h = normalize(wo + wi);
  return mix(satDot(n, wi) * PiInv, GTR2Pdf(n, h, wo, alpha) / (4.0f * absDot(h, wo)), 1.0f / (2.0f - metallic));
}
// metallicWorkflowEval (:123-144): value == metallicWorkflowBSDF; its pdf out-param only ever
// feeds the global `dummyPdf` on the live path, so it is not computed here.
static vec3 metallicWorkflowEval(const State& state, vec3 n, vec3 wo, vec3 wi) { return metallicWorkflowBSDF(state, n, wo, wi); }
static float metallicWorkflowSample(const State& state, vec3 n, vec3 wo, vec3 r, vec3& bsdf, vec3& dir) {   // :146-166
  float roughness = state.mat.roughness;
  float metallic = state.mat.metallic;
  float alpha = roughness;
  if (r.z > (1.0f / (2.0f - metallic))) {
    dir = sampleHemisphereCosine(n, vec2(r.x, r.y));
  } else {
    vec3 h = GTR2Sample(n, wo, alpha, vec2(r.x, r.y));
    dir = -reflect(wo, h);
  }
  if (dot(n, dir) < 0.0f) return InvalidPdf;
  bsdf = metallicWorkflowBSDF(state, n, wo, dir);
  return metallicWorkflowPdf(state, n, wo, dir);
}

// ---- reservoir.glsl ----------------------------------------------------------------------------
static float resvToScalar(vec3 x) { return luminance(x); }
static void resvReset(DirectReservoir& r) { r.num = 0; r.weight = 0; }
static void resvReset(IndirectReservoir& r) { r.num = 0; r.weight = 0; r.bigW = 0; }
static bool resvInvalid(const DirectReservoir& r) { return gisnan(r.weight) || r.weight < 0.0f; }
static bool resvInvalid(const IndirectReservoir& r) { return gisnan(r.weight) || r.weight < 0.0f; }
static void resvCheckValidity(DirectReservoir& r) { if (resvInvalid(r)) resvReset(r); }
static void resvCheckValidity(IndirectReservoir& r) { if (resvInvalid(r)) resvReset(r); }
static void resvUpdate(DirectReservoir& resv, const LightSample& s, float newWeight, float r) {   // :47-53
  resv.weight += newWeight;
  resv.num += 1;
  if (r * resv.weight < newWeight) resv.lightSample = s;
}
static void resvUpdate(IndirectReservoir& resv, const GISample& s, float newWeight, float r) {    // :55-61
  resv.weight += newWeight;
  resv.num += 1;
  if (r * resv.weight < newWeight) resv.giSample = s;
}
static void resvMerge(DirectReservoir& resv, const DirectReservoir& rhs, float r) {               // :69-75
  resv.weight += rhs.weight;
  resv.num += rhs.num;
  if (r * resv.weight < rhs.weight) resv.lightSample = rhs.lightSample;
}
static void resvClamp(DirectReservoir& resv, int clamp) {                                         // :115-120
  if (resv.num > (uint)clamp) { resv.weight *= float(clamp) / float(resv.num); resv.num = clamp; }
}
static void resvClamp(IndirectReservoir& resv, int clamp) {                                       // :122-127
  if (resv.num > (uint)clamp) { resv.weight *= float(clamp) / float(resv.num); resv.num = clamp; }
}

static inline vec3 V(const eid_vec3& v) { return vec3(v.x, v.y, v.z); }
static inline eid_vec3 E(vec3 v) { return eid_vec3{v.x, v.y, v.z}; }

// =================================================================================================
// Per-invocation context: the GLSL globals (prd, imageCoords, rtxState) + resource bindings
// =================================================================================================
struct Ctx {
  const Scene& sc; Renderer& rr; const RtxState& rtxState;
  const SceneCamera& cam;
  std::vector<uvec4>& thisGbuffer; const std::vector<uvec4>& lastGbuffer;
  std::vector<DirectReservoir>& thisDirectResv; const std::vector<DirectReservoir>& lastDirectResv;
  std::vector<IndirectReservoir>& thisIndirectResv; const std::vector<IndirectReservoir>& lastIndirectResv;
  PtPayload prd; ivec2 imageCoords;
  uint32_t pitch;   // allocation width of the 2-D images
  bool primeOnly = false;   // spatial reuse, first dispatch: run up to cacheTempReservoir and stop (see Renderer::runDirect)

  Ctx(const Scene& s, Renderer& r, const RtxState& st, int set)
      : sc(s), rr(r), rtxState(st), cam(s.camera),
        // descriptor set i: last* = [i], this* = [!i] (renderer.cpp:341-375)
        thisGbuffer(r.gbuffer[!set]), lastGbuffer(r.gbuffer[set]),
        thisDirectResv(r.directResv[!set]), lastDirectResv(r.directResv[set]),
        thisIndirectResv(r.indirectResv[!set]), lastIndirectResv(r.indirectResv[set]), pitch(r.width) {}

  ivec2 size() const { return ivec2(rtxState.size.x, rtxState.size.y); }
  ivec2 indSize() const { return ivec2(rtxState.size.x / 2, rtxState.size.y / 2); }
  float rand() { return rnd(prd.seed); }

  // image loads: out-of-bounds reads return 0 (Vulkan robust image access)
  uvec4 loadG(const std::vector<uvec4>& img, ivec2 c) const {
    if (c.x < 0 || c.y < 0 || c.x >= (int)rr.width || c.y >= (int)rr.height) return uvec4();
    return img[(size_t)c.y * pitch + c.x];
  }
  vec4 loadImg(const std::vector<vec4>& img, ivec2 c) const {
    if (c.x < 0 || c.y < 0 || c.x >= (int)rr.width || c.y >= (int)rr.height) return vec4();
    return img[(size_t)c.y * pitch + c.x];
  }
  void storeImg(std::vector<vec4>& img, ivec2 c, vec4 v) const {
    if (c.x < 0 || c.y < 0 || c.x >= (int)rr.width || c.y >= (int)rr.height) return;
    img[(size_t)c.y * pitch + c.x] = v;
  }

  // ---- traceray_rq.glsl:32-102 HitTest: stochastic alpha for candidates of non-FORCE_OPAQUE instances; one draw each ------
  bool HitTest(const Hit& c) {
    const uint matIndex = (uint)imax(0, sc.instMaterial[c.instanceCustomIndex]);
    const GltfShadeMaterial& mat = sc.shadeMaterials[matIndex];
    float baseColorAlpha = mat.pbrBaseColorFactor.w;
    if (mat.pbrBaseColorTexture > -1) {
      const auto& indices = sc.indexBufs[c.instanceCustomIndex];
      const auto& vertices = sc.vertexBufs[c.instanceCustomIndex];
      const VertexAttributes& a0 = vertices[indices[3 * c.primitiveID]];
      const VertexAttributes& a1 = vertices[indices[3 * c.primitiveID + 1]];
      const VertexAttributes& a2 = vertices[indices[3 * c.primitiveID + 2]];
      const vec3 barycentrics = vec3(1.0f - c.bary.x - c.bary.y, c.bary.x, c.bary.y);
      // (the reference interpolates the RAW texcoord here, handedness bit included: traceray_rq.glsl:76-79)
      vec2 texcoord0 = (vec2(a0.texcoord.x, a0.texcoord.y) * barycentrics.x + vec2(a1.texcoord.x, a1.texcoord.y) * barycentrics.y) +
                       vec2(a2.texcoord.x, a2.texcoord.y) * barycentrics.z;
      baseColorAlpha *= textureLod(mat.pbrBaseColorTexture, texcoord0).w;
    }
    float opacity;
    if (mat.alphaMode == ALPHA_MASK) opacity = baseColorAlpha > mat.alphaCutoff ? 1.0f : 0.0f;
    else opacity = baseColorAlpha;
    if (rand() > opacity) return false;
    return true;
  }
  // Candidate order is implementation-defined in Vulkan; the contract fixes it to front-to-back:
  // candidates are visited in increasing (t, instanceID, primitiveID) until one is opaque or passes HitTest.
  bool firstAcceptedHit(const Ray& r, float tmax, Hit& out) {
    HitKey key{-1.0f, -1, -1};
    bool haveKey = false;
    for (;;) {
      Hit h = sc.closestHit(r.origin, r.direction, tmax, nullptr, haveKey ? &key : nullptr);
      if (!(h.hitT < INFINITY_) || h.primitiveID < 0) return false;
      if (h.opaque || HitTest(h)) { out = h; return true; }
      key = HitKey{h.hitT, h.instanceID, h.primitiveID};
      haveKey = true;
    }
  }
  // ---- traceray_rq.glsl:108-147 ------------------------------------------------------------------
  void ClosestHit(const Ray& r) {
    prd.hitT = INFINITY_;
    rr.closestRays.fetch_add(1, std::memory_order_relaxed);
    Hit h;
    if (firstAcceptedHit(r, INFINITY_, h)) {
      prd.hitT = h.hitT; prd.primitiveID = h.primitiveID; prd.instanceID = h.instanceID;
      prd.instanceCustomIndex = h.instanceCustomIndex; prd.baryCoord = h.bary;
      prd.objectToWorld = sc.objectToWorld[h.instanceID];
      prd.worldToObject = sc.worldToObject[h.instanceID];
    }
  }
  // ---- traceray_rq.glsl:153-185 ----------------------------------------------------------------
  bool AnyHit(const Ray& r, float maxDist) {
    if (!sc.hasNonOpaque) return sc.anyHit(r.origin, r.direction, maxDist, &rr.anyRays);
    rr.anyRays.fetch_add(1, std::memory_order_relaxed);
    Hit h;
    return firstAcceptedHit(r, maxDist, h);
  }

  // ---- shade_state.glsl:147-221 ----------------------------------------------------------------
  State GetState(const PtPayload& hstate, vec3 rayDir) {
    State state;
    const uint idGeo = hstate.instanceCustomIndex;
    const uint idPrim = hstate.primitiveID;
    const vec3 bary = vec3(1.0f - hstate.baryCoord.x - hstate.baryCoord.y, hstate.baryCoord.x, hstate.baryCoord.y);
    const auto& indices = sc.indexBufs[idGeo];
    const auto& vertices = sc.vertexBufs[idGeo];
    uint t0 = indices[3 * idPrim], t1 = indices[3 * idPrim + 1], t2 = indices[3 * idPrim + 2];
    const VertexAttributes& attr0 = vertices[t0];
    const VertexAttributes& attr1 = vertices[t1];
    const VertexAttributes& attr2 = vertices[t2];
    const uint matIndex = (uint)imax(0, sc.instMaterial[idGeo]);

    const vec3 pos0 = V(attr0.position), pos1 = V(attr1.position), pos2 = V(attr2.position);
    const vec3 position = (pos0 * bary.x + pos1 * bary.y) + pos2 * bary.z;
    const vec3 world_position = mulPoint(hstate.objectToWorld, position);
    vec3 wpos0 = mulPoint(hstate.objectToWorld, pos0);
    vec3 wpos1 = mulPoint(hstate.objectToWorld, pos1);
    vec3 wpos2 = mulPoint(hstate.objectToWorld, pos2);

    vec3 nrm0 = decompress_unit_vec(attr0.normal), nrm1 = decompress_unit_vec(attr1.normal), nrm2 = decompress_unit_vec(attr2.normal);
    vec3 normal = normalize((nrm0 * bary.x + nrm1 * bary.y) + nrm2 * bary.z);
    vec3 world_normal = normalize(mulTransposed(normal, hstate.worldToObject));
    vec3 geom_normal = normalize(cross(pos1 - pos0, pos2 - pos0));
    vec3 wgeom_normal = normalize(mulTransposed(geom_normal, hstate.worldToObject));

    float h0 = (floatBitsToInt(attr0.texcoord.y) & 1) == 1 ? 1.0f : -1.0f;
    vec3 tng0 = decompress_unit_vec(attr0.tangent), tng1 = decompress_unit_vec(attr1.tangent), tng2 = decompress_unit_vec(attr2.tangent);
    vec3 tangent = (tng0 * bary.x + tng1 * bary.y) + tng2 * bary.z;
    tangent = normalize(tangent);
    vec3 world_tangent = normalize(mulVector(hstate.objectToWorld, tangent));
    world_tangent = normalize(world_tangent - world_normal * dot(world_tangent, world_normal));
    vec3 world_binormal = cross(world_normal, world_tangent) * h0;

    auto decode_texture = [](eid_vec2 t) { return vec2(t.x, uintBitsToFloat(floatBitsToUint(t.y) & ~1u)); };
    const vec2 uv0 = decode_texture(attr0.texcoord), uv1 = decode_texture(attr1.texcoord), uv2 = decode_texture(attr2.texcoord);
    const vec2 texcoord0 = (uv0 * bary.x + uv1 * bary.y) + uv2 * bary.z;

    state.position = world_position;
    state.normal = (dot(world_normal, wgeom_normal) > 0.0f) ? world_normal : -world_normal;
    state.ffnormal = dot(state.normal, rayDir) <= 0.0f ? state.normal : -state.normal;
    state.texCoord = texcoord0;
    state.tangent = world_tangent;
    state.bitangent = world_binormal;
    state.matID = matIndex;
    state.area = length(cross(wpos1 - wpos0, wpos2 - wpos0)) * 0.5f;
    return state;
  }

  // textureLod(texturesMap[i], uv, 0)
  vec4 textureLod(int i, vec2 uv) const {
    if (i < 0 || (size_t)i >= sc.textures.size()) return vec4(1.0f);   // robustness: unbound slot reads the white default
    return sc.textures[i].sample(uv);
  }
  // gltf_material.glsl:36-46 (SRGB_FAST_APPROXIMATION): pow(rgb, 2.2), alpha untouched
  static vec4 SRGBtoLINEAR(vec4 c) { return vec4(eid_powf(c.x, 2.2f), eid_powf(c.y, 2.2f), eid_powf(c.z, 2.2f), c.w); }
  // common.glsl:81-93
  static void CreateCoordinateSystem(vec3 N, vec3& Nt, vec3& Nb) {
    Nt = normalize((gabs(N.z) > 0.99999f) ? vec3(-N.x * N.y, 1.0f - N.y * N.y, -N.y * N.z) : vec3(-N.x * N.z, -N.y * N.z, 1.0f - N.z * N.z));
    Nb = cross(Nt, N);
  }
  // ---- gltf_material.glsl:130-176 GetMaterials + GetMetallicRoughness :52-91 --------------------
  void GetMaterials(State& state, const Ray& r) {
    const GltfShadeMaterial& material = sc.shadeMaterials[state.matID];
    mat3 TBN(state.tangent, state.bitangent, state.normal);
    if (material.normalTexture > -1) {
      vec3 normalVector = textureLod(material.normalTexture, state.texCoord).xyz();
      normalVector = normalize(normalVector * 2.0f - 1.0f);
      normalVector *= vec3(material.normalTextureScale, material.normalTextureScale, 1.0f);
      state.normal = normalize(TBN * normalVector);
      state.ffnormal = dot(state.normal, r.direction) <= 0.0f ? state.normal : -state.normal;
      CreateCoordinateSystem(state.ffnormal, state.tangent, state.bitangent);
    }
    state.mat.emission = V(material.emissiveFactor);
    if (material.emissiveTexture > -1)
      state.mat.emission *= SRGBtoLINEAR(textureLod(material.emissiveTexture, state.texCoord)).xyz();
    if ((state.mat.emission.x + state.mat.emission.y + state.mat.emission.z) > 1e-3f) state.isEmitter = true;
    else state.isEmitter = false;
    // GetMetallicRoughness (:52-91)
    float perceptualRoughness = material.pbrRoughnessFactor;
    float metallic = material.pbrMetallicFactor;
    if (material.pbrMetallicRoughnessTexture > -1) {
      vec4 mrSample = textureLod(material.pbrMetallicRoughnessTexture, state.texCoord);
      perceptualRoughness = mrSample.y * perceptualRoughness;
      metallic = mrSample.z * metallic;
    }
    vec4 baseColor(material.pbrBaseColorFactor.x, material.pbrBaseColorFactor.y, material.pbrBaseColorFactor.z, material.pbrBaseColorFactor.w);
    if (material.pbrBaseColorTexture > -1) {
      vec4 t = SRGBtoLINEAR(textureLod(material.pbrBaseColorTexture, state.texCoord));
      baseColor = vec4(baseColor.x * t.x, baseColor.y * t.y, baseColor.z * t.z, baseColor.w * t.w);
    }
    state.mat.albedo = baseColor.xyz();
    state.mat.metallic = metallic;
    state.mat.roughness = perceptualRoughness;
    state.mat.roughness = gmax(state.mat.roughness, 0.001f);
    state.mat.transmission = material.transmissionFactor;
    if (material.transmissionTexture > -1) state.mat.transmission *= textureLod(material.transmissionTexture, state.texCoord).x;
    state.mat.ior = material.ior;
    state.eta = dot(state.normal, state.ffnormal) > 0.0f ? (1.0f / state.mat.ior) : state.mat.ior;
  }

  // ---- pathtrace.glsl ---------------------------------------------------------------------------
  static bool IsPdfInvalid(float p) { return p <= 1e-8f || gisnan(p); }                      // :14-16
  bool Occlusion(const Ray& ray, const State& state, float dist) {                          // :18-22
    return AnyHit(ray, dist - gabs(ray.origin.x - state.position.x) - gabs(ray.origin.y - state.position.y) -
                           gabs(ray.origin.z - state.position.z));
  }
  static vec3 BSDF(const State& s, vec3 Vv, vec3 N, vec3 L) { return metallicWorkflowBSDF(s, N, Vv, L); }
  static float Pdf(const State& s, vec3 Vv, vec3 N, vec3 L) { return metallicWorkflowPdf(s, N, Vv, L); }
  static vec3 Eval(const State& s, vec3 Vv, vec3 N, vec3 L) { return metallicWorkflowEval(s, N, Vv, L); }
  vec3 Sample(const State& s, vec3 Vv, vec3 N, vec3& L, float& pdf) {                         // :36-38
    float r0 = rand(); float r1 = rand(); float r2 = rand();
    vec3 bsdf;
    pdf = metallicWorkflowSample(s, N, Vv, vec3(r0, r1, r2), bsdf, L);
    return bsdf;
  }
  // common.glsl:69-76
  static vec2 GetSphericalUv(vec3 v) {
    float gamma = eid_asinf(-v.y);
    float theta = eid_atan2f(v.z, v.x);
    const float M_1_OVER_PI = 0.318309886183790671538f;
    return vec2(theta * M_1_OVER_PI * 0.5f, gamma * M_1_OVER_PI) + 0.5f;
  }
  // texture(environmentTexture, uv).rgb — or the constant environment when no HDR map is installed (sun & sky: later row)
  vec3 envTexture(vec3 dir) { return rr.env ? rr.env->texture(GetSphericalUv(dir)) : rr.envConstant; }
  vec3 EnvRadiance(vec3 dir) {                                                                       // pathtrace.glsl:40-47
    if (rr.sunSky.in_use == 1) return sun_and_sky(rr.sunSky, dir) * rtxState.hdrMultiplier;
    return envTexture(dir) * rtxState.hdrMultiplier;
  }
  vec3 EnvEval(vec3 dir, float& pdf) {                                                               // pathtrace.glsl:60-72
    if (rr.sunSky.in_use == 1) {
      pdf = 0.5f * rtxState.environmentProb;
      return sun_and_sky(rr.sunSky, dir) * rtxState.hdrMultiplier;
    }
    vec3 radiance = envTexture(dir);
    pdf = luminance(radiance) * rtxState.envMapLuminIntegInv * rtxState.environmentProb;
    return radiance;
  }
  // env_sampling.glsl:38-94 Environment_sample
  vec3 Environment_sample(vec3 randVal, vec3& to_light, float& pdf) {
    const Environment& E = *rr.env;
    vec3 xi = randVal;
    const uint width = E.width, height = E.height;
    const uint size = width * height;
    const uint idx = (uint)imin((int)f2u(xi.x * float(size)), (int)size - 1);
    const ImptSampData sample_data = E.accel[idx];
    uint env_idx;
    if (xi.y < sample_data.q) {
      env_idx = idx;
      xi.y /= sample_data.q;
      pdf = sample_data.pdf;
    } else {
      env_idx = (uint)sample_data.alias;
      xi.y = (xi.y - sample_data.q) / (1.0f - sample_data.q);
      pdf = sample_data.aliasPdf;
    }
    const uint px = env_idx % width;
    uint py = env_idx / width;
    const float u = (float(px) + xi.y) / float(width);
    const float phi = u * (2.0f * M_PI_F) - M_PI_F;
    float sin_phi, cos_phi;
    eid_sincosf(phi, &sin_phi, &cos_phi);
    const float step_theta = M_PI_F / float(height);
    const float theta0 = float(py) * step_theta;
    const float cos_theta = eid_cosf(theta0) * (1.0f - xi.z) + eid_cosf(theta0 + step_theta) * xi.z;
    const float theta = eid_acosf(cos_theta);
    const float sin_theta = eid_sinf(theta);
    const float v = theta * 0.318309886183790671538f;
    to_light = vec3(cos_phi * sin_theta, cos_theta, sin_phi * sin_theta);
    return E.texture(vec2(u, v));
  }
  // env_sampling.glsl:100-135 EnvSample
  vec4 EnvSample(vec3& radiance) {
    vec3 lightDir; float pdf;
    if (rr.sunSky.in_use == 1) {                                                                     // :111-125
      const SunAndSky& ss = rr.sunSky;
      float sun_radius = (0.00465f * 10.0f) * ss.sun_disk_scale;
      vec3 sunDirection = V(ss.sun_direction);
      vec3 T, B;
      CreateCoordinateSystem(sunDirection, T, B);
      vec3 dir;
      dir.x = rand() * sun_radius;
      dir.y = rand() * sun_radius;
      dir.z = sqrtf(gmax(0.0f, 1.0f - dir.x * dir.x - dir.y * dir.y));
      lightDir = normalize(T * dir.x + B * dir.y + sunDirection * dir.z);
      radiance = sun_and_sky(ss, lightDir);
      pdf = 0.5f;
      radiance *= rtxState.hdrMultiplier;
      return vec4(lightDir, pdf);
    }
    float r0 = rand(); float r1 = rand(); float r2 = rand();
    radiance = Environment_sample(vec3(r0, r1, r2), lightDir, pdf);
    radiance *= rtxState.hdrMultiplier;
    return vec4(lightDir, pdf);
  }
  vec3 LightEval(const State& state, float dist, vec3 dir, float& pdf) {                      // :74-88
    float lightProb = (1.0f - rtxState.environmentProb);
    const GltfShadeMaterial& mat = sc.shadeMaterials[state.matID];
    vec3 emission = V(mat.emissiveFactor);
    pdf = luminance(emission) * rtxState.lightLuminIntegInv * lightProb;
    pdf *= dist * dist / absDot(state.ffnormal, dir);
    if (mat.emissiveTexture > -1) emission *= SRGBtoLINEAR(textureLod(mat.emissiveTexture, state.texCoord)).xyz();
    return emission / state.area;
  }
  vec2 SampleTriangleUniform() {                                                              // :90-97
    float ru = rand();
    float rv = rand();
    float r = sqrtf(rv);
    float u = 1.0f - r;
    float v = ru * r;
    return vec2(u, v);
  }
  float SampleTriangleLight(vec3 x, LightSample& lightSample) {                               // :103-139
    if (sc.lightBufInfo.trigLightSize == 0) return InvalidPdf;
    int id = imin(f2i(float(sc.lightBufInfo.trigLightSize) * rand()), int(sc.lightBufInfo.trigLightSize) - 1);
    if (rand() > sc.trigLights[id].impSamp.q) id = sc.trigLights[id].impSamp.alias;
    const TrigLight& light = sc.trigLights[id];
    vec3 v0 = V(light.v0), v1 = V(light.v1), v2 = V(light.v2);
    vec3 normal = cross(v1 - v0, v2 - v0);
    float area = length(normal) * 0.5f;
    normal = normalize(normal);
    vec2 baryCoord = SampleTriangleUniform();
    vec3 y = (baryCoord.x * v0 + baryCoord.y * v1) + (1 - baryCoord.x - baryCoord.y) * v2;
    const GltfShadeMaterial& mat = sc.shadeMaterials[light.matIndex];
    vec3 emission = V(mat.emissiveFactor);
    if (mat.emissiveTexture > -1) {
      vec2 uv = (baryCoord.x * vec2(light.uv0.x, light.uv0.y) + baryCoord.y * vec2(light.uv1.x, light.uv1.y)) +
                (1 - baryCoord.x - baryCoord.y) * vec2(light.uv2.x, light.uv2.y);
      emission *= SRGBtoLINEAR(textureLod(mat.emissiveTexture, uv)).xyz();
    }
    vec3 dir = y - x;
    float dist = length(dir);
    lightSample.Li = E(emission / area);
    lightSample.wi = E(dir / dist);
    lightSample.dist = dist;
    return light.impSamp.pdf * (dist * dist) / (area * gabs(dot(V(lightSample.wi), normal)));
  }
  float SamplePuncLight(vec3 x, LightSample& lightSample) {                                   // :141-159
    if (sc.lightBufInfo.puncLightSize == 0) return InvalidPdf;
    int id = imin(f2i(float(sc.lightBufInfo.puncLightSize) * rand()), int(sc.lightBufInfo.puncLightSize) - 1);
    if (rand() > sc.puncLights[id].impSamp.q) id = sc.puncLights[id].impSamp.alias;
    const PuncLight& light = sc.puncLights[id];
    vec3 dir = V(light.position) - x;
    float dist = length(dir);
    lightSample.Li = E(V(light.color) * light.intensity / (dist * dist));
    lightSample.wi = E(dir / dist);
    lightSample.dist = dist;
    return light.impSamp.pdf;
  }
  float SampleDirectLightNoVisibility(vec3 pos, LightSample& lightSample) {                   // :161-183
    float r = rand();
    if (r < rtxState.environmentProb) {
      if (!rr.env && rr.sunSky.in_use != 1) return InvalidPdf;   // nothing to sample: the C-ABI refuses environmentProb > 0 in that case
      vec3 Li;
      vec4 dirAndPdf = EnvSample(Li);
      lightSample.Li = E(Li);
      if (IsPdfInvalid(dirAndPdf.w)) return InvalidPdf;
      lightSample.wi = E(dirAndPdf.xyz());
      lightSample.dist = INFINITY_;
      return dirAndPdf.w * rtxState.environmentProb;
    } else {
      if (r < rtxState.environmentProb + (1.0f - rtxState.environmentProb) * sc.lightBufInfo.trigSampProb)
        return (1.0f - rtxState.environmentProb) * SampleTriangleLight(pos, lightSample) * sc.lightBufInfo.trigSampProb;
      else
        return (1.0f - rtxState.environmentProb) * SamplePuncLight(pos, lightSample) * (1.0f - sc.lightBufInfo.trigSampProb);
    }
  }
  float SampleDirectLight(const State& state, vec3& radiance, vec3& dir) {                    // :185-202
    LightSample lsample{};
    float pdf = SampleDirectLightNoVisibility(state.position, lsample);
    if (IsPdfInvalid(pdf)) return InvalidPdf;
    Ray shadowRay;
    shadowRay.origin = OffsetRay(state.position, state.ffnormal);
    shadowRay.direction = V(lsample.wi);
    if (Occlusion(shadowRay, state, lsample.dist)) return InvalidPdf;
    radiance = V(lsample.Li);
    dir = V(lsample.wi);
    return pdf;
  }
  vec3 DirectLight(const State& state, vec3 wo) {                                             // :204-220
    LightSample lightSample{};
    float pdf = SampleDirectLightNoVisibility(state.position, lightSample);
    if (IsPdfInvalid(pdf)) return vec3(0.0f);
    Ray shadowRay;
    shadowRay.origin = OffsetRay(state.position, state.ffnormal);
    shadowRay.direction = V(lightSample.wi);
    if (Occlusion(shadowRay, state, lightSample.dist)) return vec3(0.0f);
    return V(lightSample.Li) * Eval(state, wo, state.ffnormal, V(lightSample.wi)) *
           gmax(dot(state.ffnormal, V(lightSample.wi)), 0.0f) / pdf;
  }
  vec3 clampRadiance(vec3 radiance) {                                                          // :222-232
    if (gisnan(radiance.x) || gisnan(radiance.y) || gisnan(radiance.z)) return vec3(0.0f);
    float lum = luminance(radiance);
    if (lum > rtxState.fireflyClampThreshold) radiance *= rtxState.fireflyClampThreshold / lum;
    return radiance;
  }
  void loadLastGeometryInfo(ivec2 c, vec3& normal, float& depth, uint& matHash) {             // :240-245
    uvec4 gInfo = loadG(lastGbuffer, c);
    normal = decompress_unit_vec(gInfo.y);
    depth = uintBitsToFloat(gInfo.x);
    matHash = gInfo.w & 0xFF000000;
  }
  Ray raySpawn(ivec2 coord, ivec2 sizeImage) {                                                 // :260-270
    const vec2 pixelCenter = vec2((float)coord.x, (float)coord.y) + 0.5f;
    const vec2 inUV = pixelCenter / vec2((float)sizeImage.x, (float)sizeImage.y);
    vec2 d = inUV * 2.0f - 1.0f;
    const mat4& VI = *reinterpret_cast<const mat4*>(&cam.viewInverse);
    const mat4& PI = *reinterpret_cast<const mat4*>(&cam.projInverse);
    vec3 origin(VI.m[12], VI.m[13], VI.m[14]);            // viewInverse * (0,0,0,1)
    vec4 target = mul(PI, vec4(d.x, d.y, 1, 1));
    vec3 direction = mulDir(VI, normalize(target.xyz()));
    Ray r; r.origin = origin; r.direction = normalize(direction);
    return r;
  }
  bool getIndirectStateFromGBuffer(const std::vector<uvec4>& gBuffer, const Ray& ray, State& state, float& depth) {   // :296-313
    if (rr.variant & EID_VARIANT_FETCH_4_SUBPIXELS) {                                            // #if FETCH_GEOM_CHECK_4_SUBPIXELS (:314-358)
      uvec4 gInfo00 = loadG(gBuffer, imageCoords * 2 + ivec2(0, 0));
      uvec4 gInfo10 = loadG(gBuffer, imageCoords * 2 + ivec2(1, 0));
      uvec4 gInfo11 = loadG(gBuffer, imageCoords * 2 + ivec2(1, 1));
      uvec4 gInfo01 = loadG(gBuffer, imageCoords * 2 + ivec2(0, 1));
      depth = (uintBitsToFloat(gInfo00.x) + uintBitsToFloat(gInfo10.x) + uintBitsToFloat(gInfo11.x) + uintBitsToFloat(gInfo01.x)) * 0.25f;
      if (depth >= INFINITY_ - EPS * 10.0f) return false;
      state.position = ray.origin + ray.direction * depth;
      state.normal = (decompress_unit_vec(gInfo00.y) + decompress_unit_vec(gInfo10.y) + decompress_unit_vec(gInfo11.y) + decompress_unit_vec(gInfo01.y)) * 0.25f;
      state.ffnormal = dot(state.normal, ray.direction) <= 0.0f ? state.normal : -state.normal;
      state.mat.albedo = (unpackUnorm4x8(gInfo00.w).xyz() + unpackUnorm4x8(gInfo10.w).xyz() + unpackUnorm4x8(gInfo11.w).xyz() + unpackUnorm4x8(gInfo01.w).xyz()) * 0.25f;
      vec4 matInfo00 = unpackUnorm4x8(gInfo00.z), matInfo10 = unpackUnorm4x8(gInfo10.z), matInfo11 = unpackUnorm4x8(gInfo11.z), matInfo01 = unpackUnorm4x8(gInfo01.z);
      state.mat.metallic = (matInfo00.x + matInfo10.x + matInfo11.x + matInfo01.x) * 0.25f;
      state.mat.roughness = (matInfo00.y + matInfo10.y + matInfo11.y + matInfo01.y) * 0.25f;
      state.mat.ior = (matInfo00.z + matInfo10.z + matInfo11.z + matInfo01.z) * 0.25f * MAX_IOR_MINUS_ONE + 1.f;
      state.mat.transmission = (matInfo00.w + matInfo01.w + matInfo11.w + matInfo10.w) * 0.25f;
      float r = rand();
      if (r < 0.25f) state.matID = gInfo00.w >> 24;
      else if (r < 0.5f) state.matID = gInfo10.w >> 24;
      else if (r < 0.75f) state.matID = gInfo11.w >> 24;
      else state.matID = gInfo01.w >> 24;
      return true;
    }
    uvec4 gInfo = loadG(gBuffer, imageCoords * 2);
    depth = uintBitsToFloat(gInfo.x);
    if (depth >= INFINITY_ * 0.8f) return false;
    state.position = ray.origin + ray.direction * depth;
    state.normal = decompress_unit_vec(gInfo.y);
    state.ffnormal = dot(state.normal, ray.direction) <= 0.0f ? state.normal : -state.normal;
    state.mat.albedo = unpackUnorm4x8(gInfo.w).xyz();
    vec4 matInfo = unpackUnorm4x8(gInfo.z);
    state.mat.metallic = matInfo.x;
    state.mat.roughness = matInfo.y;
    state.mat.ior = matInfo.z * MAX_IOR_MINUS_ONE + 1.f;
    state.mat.transmission = matInfo.w;
    state.matID = gInfo.w >> 24;
    return true;
  }
  vec3 DebugInfo(const State& state) {                                                         // :362-380
    switch (rtxState.debugging_mode) {
      case eMetallic: return vec3(state.mat.metallic);
      case eNormal: return (state.normal + vec3(1)) * .5f;
      case eDepth: return vec3(0.0f);
      case eBaseColor: return state.mat.albedo;
      case eEmissive: return state.mat.emission;
      case eRoughness: return vec3(state.mat.roughness);
      case eTexcoord: return vec3(state.texCoord, 0);
    }
    return vec3(1000, 0, 0);
  }

  // ---- direct_stage.comp ------------------------------------------------------------------------
  uvec4 encodeGeometryInfo(const State& state, float depth) {                                  // :37-45
    uvec4 gInfo;
    gInfo.x = floatBitsToUint(depth);
    gInfo.y = compress_unit_vec(state.normal);
    gInfo.z = packUnorm4x8(vec4(state.mat.metallic, state.mat.roughness, (state.mat.ior - 1.0f) / MAX_IOR_MINUS_ONE, state.mat.transmission));
    gInfo.w = packUnorm4x8(vec4(state.mat.albedo, 1.0f)) & 0xFFFFFF;
    gInfo.w += hash8bit(state.matID);
    return gInfo;
  }
  bool findTemporalNeighborDirect(vec3 norm, float depth, float reprojDepth, uint matId, ivec2 lastCoord, DirectReservoir& resv) {   // :47-84
    vec3 pnorm; float pdepth; uint matHash;
    (void)depth;
    if (!inBound(lastCoord, ivec2(2, 0), size())) return false;
    loadLastGeometryInfo(lastCoord, pnorm, pdepth, matHash);
    if (inBound(lastCoord, size())) {
      if (hash8bit(matId) == matHash) {
        if (dot(norm, pnorm) > 0.9f && reprojDepth < pdepth * 1.05f) {
          resv = lastDirectResv[(size_t)lastCoord.y * rtxState.size.x + lastCoord.x];
          return true;
        }
      }
    }
    return false;
  }
  void loadThisGeometryInfo(ivec2 c, vec3& normal, float& depth) {                            // pathtrace.glsl:247-251
    uvec4 gInfo = loadG(thisGbuffer, c);
    normal = decompress_unit_vec(gInfo.y);
    depth = uintBitsToFloat(gInfo.x);
  }
  bool findSpatialNeighbor(vec3 norm, float depth, uint matId, DirectReservoir& resv) {       // :86-108 (Radius is unused there too)
    (void)matId;
    const float r0 = rand(), r1 = rand();
    vec2 p = toConcentricDisk(vec2(r0, r1));
    int px = f2i(float((float)imageCoords.x + p.x) + 0.5f);
    int py = f2i(float((float)imageCoords.y + p.y) + 0.5f);
    int pidx = py * rtxState.size.x + px;
    vec3 pnorm; float pdepth;
    loadThisGeometryInfo(imageCoords, pnorm, pdepth);      // the pixel's own G-buffer entry, as in the reference
    if (!inBound(ivec2(px, py), size())) return false;
    else if (dot(norm, pnorm) < 0.5f || gabs(depth - pdepth) > depth * 0.1f) return false;
    resv = rr.tempDirectResv[(size_t)pidx];
    return true;
  }
  bool mergeSpatialNeighbors(vec3 norm, float depth, uint matId, DirectReservoir& resv) {     // :110-123
    bool valid = false;
    resvReset(resv);
    for (int i = 0; i < 5; i++) {
      DirectReservoir spatial{};
      if (findSpatialNeighbor(norm, depth, matId, spatial)) {
        if (!resvInvalid(spatial)) {
          resvMerge(resv, spatial, rand());
          valid = true;
        }
      }
    }
    return valid;
  }
  ivec2 createMotionIndex(vec3 wpos) {                                                         // :125-139
    const mat4& LPV = *reinterpret_cast<const mat4*>(&cam.lastProjView);
    vec4 proj = mul(LPV, vec4(wpos, 1.0f));
    vec3 ndc = proj.xyz() / proj.w;
    vec2 mv = vec2(ndc.x, ndc.y) * 0.5f + 0.5f;
    vec2 s = mv * vec2((float)rtxState.size.x, (float)rtxState.size.y);
    return ivec2(f2i(s.x), f2i(s.y));
  }
  void storeMotion(ivec2 c, ivec2 mv) {   // RG16_SINT image store: saturating conversion (contract)
    if (c.x < 0 || c.y < 0 || c.x >= (int)rr.width || c.y >= (int)rr.height) return;
    auto sat = [](int v) { return (int16_t)imax(-32768, imin(32767, v)); };
    rr.motion[2 * ((size_t)c.y * pitch + c.x)] = sat(mv.x);
    rr.motion[2 * ((size_t)c.y * pitch + c.x) + 1] = sat(mv.y);
  }

  vec3 ReSTIRDirect(const Ray& r) {                                                            // :150-270
    ClosestHit(r);
    if (prd.hitT >= INFINITY_) {
      thisGbuffer[(size_t)imageCoords.y * pitch + imageCoords.x] = uvec4(floatBitsToUint(INFINITY_), 0, 0, InvalidMatId);
      storeMotion(imageCoords, ivec2(0, 0));
      return EnvRadiance(r.direction);
    }
    rr.primaryHits.fetch_add(1, std::memory_order_relaxed);
    State state = GetState(prd, r.direction);
    GetMaterials(state, r);

    ivec2 motionIdx = createMotionIndex(state.position);
    uvec4 gInfo = encodeGeometryInfo(state, prd.hitT);
    storeMotion(imageCoords, motionIdx);
    thisGbuffer[(size_t)imageCoords.y * pitch + imageCoords.x] = gInfo;

    if (rtxState.debugging_mode > eIndirectStage) return DebugInfo(state);
    if (state.isEmitter) return state.mat.emission;

    vec3 wo = -r.direction;
    vec3 direct = vec3(0.0f);
    state.mat.albedo = vec3(1.0f);

    if (rtxState.ReSTIRState == eNone) {
      direct = DirectLight(state, wo);
    } else {
      DirectReservoir resv{};
      resvReset(resv);
      for (int i = 0; i < rtxState.RISSampleNum; i++) {
        LightSample lsample{};
        float p = SampleDirectLightNoVisibility(state.position, lsample);
        vec3 pHat = V(lsample.Li) * Eval(state, wo, state.ffnormal, V(lsample.wi)) * gabs(dot(state.ffnormal, V(lsample.wi)));
        float weight = resvToScalar(pHat / p);
        if (IsPdfInvalid(p) || gisnan(weight)) weight = 0.0f;
        resvUpdate(resv, lsample, weight, rand());
      }
      LightSample lsample = resv.lightSample;
      Ray shadowRay;
      shadowRay.origin = OffsetRay(state.position, state.ffnormal);
      shadowRay.direction = V(lsample.wi);
      if (Occlusion(shadowRay, state, lsample.dist)) resv.weight = 0.0f;

      if (rtxState.ReSTIRState == eTemporal || rtxState.ReSTIRState == eSpatiotemporal) {
        float reprojDepth = length(V(cam.lastPosition) - state.position);
        DirectReservoir temporal{};
        if (findTemporalNeighborDirect(state.normal, prd.hitT, reprojDepth, state.matID, motionIdx, temporal)) {
          if (!resvInvalid(temporal)) resvMerge(resv, temporal, rand());
        }
      }
      DirectReservoir tempResv = resv;
      resvCheckValidity(tempResv);
      resvClamp(tempResv, rtxState.RISSampleNum * rtxState.reservoirClamp);
      thisDirectResv[(size_t)imageCoords.y * rtxState.size.x + imageCoords.x] = tempResv;   // saveNewReservoir

      if (rtxState.ReSTIRState == eSpatial || rtxState.ReSTIRState == eSpatiotemporal) {       // :224-255
        // The reference separates its writes of tempDirectResv from the neighbours' reads by barrier(), which only orders one 8x8 work
        // group.  The contract is the race-free reading: every pixel's write happens before any pixel's read (Renderer::runDirect
        // dispatches the stage twice; a pixel writes the same value both times, so this is what the reference computes whenever its
        // neighbours' writes have landed).  Pixels that never get here (sky, emitters, debug views) keep their older entry, as there.
        DirectReservoir spatial{};
        resvReset(spatial);
        resvCheckValidity(resv);
        rr.tempDirectResv[(size_t)imageCoords.y * rtxState.size.x + imageCoords.x] = resv;      // cacheTempReservoir
        if (primeOnly) return vec3(0.0f);
        DirectReservoir spatialAggregate{};
        if (mergeSpatialNeighbors(state.normal, prd.hitT, state.matID, spatialAggregate)) {
          if (!resvInvalid(spatialAggregate)) resvMerge(spatial, spatialAggregate, rand());
        }
        resvCheckValidity(resv);
        rr.tempDirectResv[(size_t)imageCoords.y * rtxState.size.x + imageCoords.x] = resv;
        if (mergeSpatialNeighbors(state.normal, prd.hitT, state.matID, spatialAggregate)) {
          if (!resvInvalid(spatialAggregate)) resvMerge(spatial, spatialAggregate, rand());
        }
        if (!resvInvalid(spatial)) resvMerge(resv, spatial, rand());
      }
      lsample = resv.lightSample;
      if (!resvInvalid(resv)) {
        vec3 LiBsdf = V(lsample.Li) * Eval(state, wo, state.ffnormal, V(lsample.wi));
        direct = LiBsdf / resvToScalar(LiBsdf) * resv.weight / float(resv.num);
      }
    }
    if (gisnan(direct.x) || gisnan(direct.y) || gisnan(direct.z)) direct = vec3(0.0f);
    vec3 res = clampRadiance(state.mat.emission + direct);
    res = HDRToLDR(res);
    return res;
  }

  // ---- direct_gen.comp / direct_reuse.comp: the two-kernel form of the direct stage (pipelines created at renderer.cpp:129-132, never
  // dispatched by Renderer::run; EID_VARIANT_DIRECT_SPLIT runs them in place of direct_stage) ---------------------------------------
  void updateGeometryAlbedo(uvec4& gInfo, vec3 albedo) {                                       // direct_gen.comp:62-65
    uint matId = gInfo.w & 0xff000000u;
    gInfo.w = (packUnorm4x8(vec4(albedo, 1.0f)) & 0x00ffffffu) | matId;
  }
  void directGenMain(int gx, int gy) {                                                          // direct_gen.comp:77-137 + main :139-149
    ivec2 imageRes = size();
    imageCoords = ivec2(gx, gy);
    if (imageCoords.x >= imageRes.x || imageCoords.y >= imageRes.y) return;
    prd.seed = tea((uint)rtxState.size.x * (uint)gy + (uint)gx, rtxState.time);
    Ray r = raySpawn(imageCoords, imageRes);
    const size_t index = (size_t)imageCoords.y * rtxState.size.x + imageCoords.x;
    ClosestHit(r);
    DirectReservoir resv{};
    resvReset(resv);
    if (prd.hitT >= INFINITY_ * 0.8f) {
      uvec4 gInfo = uvec4(floatBitsToUint(INFINITY_), 0, 0, InvalidMatId);
      updateGeometryAlbedo(gInfo, EnvRadiance(r.direction));
      thisGbuffer[(size_t)imageCoords.y * pitch + imageCoords.x] = gInfo;
      storeMotion(imageCoords, ivec2(0, 0));
      thisDirectResv[index] = resv;
      return;
    }
    rr.primaryHits.fetch_add(1, std::memory_order_relaxed);
    State state = GetState(prd, r.direction);
    GetMaterials(state, r);
    ivec2 motionIdx = createMotionIndex(state.position);
    storeMotion(imageCoords, motionIdx);
    uvec4 gInfo = encodeGeometryInfo(state, prd.hitT);
    if (rtxState.debugging_mode > eIndirectStage) updateGeometryAlbedo(gInfo, DebugInfo(state));
    else if (state.isEmitter) updateGeometryAlbedo(gInfo, state.mat.emission);
    else {
      vec3 wo = -r.direction;
      state.mat.albedo = vec3(1.0f);
      for (int i = 0; i < rtxState.RISSampleNum; i++) {
        LightSample lsample{};
        float p = SampleDirectLightNoVisibility(state.position, lsample);
        vec3 pHat = V(lsample.Li) * Eval(state, wo, state.ffnormal, V(lsample.wi)) * gabs(dot(state.ffnormal, V(lsample.wi)));
        float weight = resvToScalar(pHat / p);
        if (IsPdfInvalid(p) || gisnan(weight)) weight = 0.0f;
        resvUpdate(resv, lsample, weight, rand());
      }
      LightSample lsample = resv.lightSample;
      Ray shadowRay;
      shadowRay.origin = OffsetRay(state.position, state.ffnormal);
      shadowRay.direction = V(lsample.wi);
      if (Occlusion(shadowRay, state, lsample.dist)) resv.weight = 0.0f;
    }
    thisGbuffer[(size_t)imageCoords.y * pitch + imageCoords.x] = gInfo;
    thisDirectResv[index] = resv;
  }
  bool getDirectStateFromGBuffer(const std::vector<uvec4>& gBuffer, const Ray& ray, State& state, float& depth) {   // pathtrace.glsl:277-294
    uvec4 gInfo = loadG(gBuffer, imageCoords);
    depth = uintBitsToFloat(gInfo.x);
    if (depth >= INFINITY_ * 0.8f) return false;
    state.position = ray.origin + ray.direction * depth;
    state.normal = decompress_unit_vec(gInfo.y);
    state.ffnormal = dot(state.normal, ray.direction) <= 0.0f ? state.normal : -state.normal;
    state.mat.albedo = unpackUnorm4x8(gInfo.w).xyz();
    vec4 matInfo = unpackUnorm4x8(gInfo.z);
    state.mat.metallic = matInfo.x;
    state.mat.roughness = matInfo.y;
    state.mat.ior = matInfo.z * MAX_IOR_MINUS_ONE + 1.f;
    state.mat.transmission = matInfo.w;
    state.matID = gInfo.w >> 24;
    return true;
  }
  void directReuseMain(int gx, int gy) {                                                        // direct_reuse.comp:102-153
    imageCoords = ivec2(gx, gy);
    if (imageCoords.x >= rtxState.size.x || imageCoords.y >= rtxState.size.y) return;
    int index = imageCoords.y * rtxState.size.x + imageCoords.x;
    prd.seed = tea((uint)(index + rtxState.size.x * rtxState.size.y), rtxState.time);
    Ray ray = raySpawn(imageCoords, size());
    State state{};
    float depth;
    if (!getDirectStateFromGBuffer(thisGbuffer, ray, state, depth)) {
      storeImg(rr.directResult, imageCoords, vec4(0.0f));
      return;
    }
    state.mat.albedo = vec3(1.0f);
    vec3 direct = vec3(0.0f);
    DirectReservoir resv = thisDirectResv[(size_t)index];
    LightSample lsample = resv.lightSample;
    ivec2 motionIdx(0, 0);
    if (imageCoords.x < (int)rr.width && imageCoords.y < (int)rr.height)
      motionIdx = ivec2(rr.motion[2 * ((size_t)imageCoords.y * pitch + imageCoords.x)], rr.motion[2 * ((size_t)imageCoords.y * pitch + imageCoords.x) + 1]);
    if (rtxState.ReSTIRState == eTemporal || rtxState.ReSTIRState == eSpatiotemporal) {
      float reprojDepth = length(V(cam.lastPosition) - state.position);
      DirectReservoir temporal{};
      if (findTemporalNeighborDirect(state.normal, depth, reprojDepth, state.matID, motionIdx, temporal)) {
        if (!resvInvalid(temporal)) resvMerge(resv, temporal, rand());
      }
    }
    if (!resvInvalid(resv)) direct = V(lsample.Li);      // (the LiBSDF of :134 is computed and not used)
    resvClamp(resv, rtxState.RISSampleNum * rtxState.reservoirClamp);
    resvCheckValidity(resv);
    if (gisnan(direct.x) || gisnan(direct.y) || gisnan(direct.z)) direct = vec3(0.0f);
    thisDirectResv[(size_t)index] = resv;
    storeImg(rr.directResult, imageCoords, vec4(HDRToLDR(clampRadiance(direct)), 1.0f));
  }

  void directMain(int gx, int gy) {                                                            // :272-289
    ivec2 imageRes = size();
    imageCoords = ivec2(gx, gy);
    if (imageCoords.x >= imageRes.x || imageCoords.y >= imageRes.y) return;
    prd.seed = tea((uint)rtxState.size.x * (uint)gy + (uint)gx, rtxState.time);
    Ray ray = raySpawn(imageCoords, imageRes);
    vec3 radiance = ReSTIRDirect(ray);
    vec3 pixelColor = clampRadiance(radiance);
    if (rr.variant & EID_VARIANT_DIRECT_BILATERAL) storeImg(rr.denoiseTemp[0], imageCoords, vec4(pixelColor, 1));   // #if DENOISER_DIRECT_BILATERAL (:284-288)
    else storeImg(rr.directResult, imageCoords, vec4(pixelColor, 1));
  }

  // ---- indirect_stage.comp ----------------------------------------------------------------------
  float MIS(float f, float g) { return (rtxState.MIS > 0) ? powerHeuristic(f, g) : 1.0f; }      // :59-61
  static float pHatIndirect(const GISample& g) { return resvToScalar(V(g.L)); }                 // :63-64 (rest is dead code)
  static float bigWIndirect(const IndirectReservoir& resv) { return resv.weight / (pHatIndirect(resv.giSample) * float(resv.num)); }   // :70-72
  bool findTemporalNeighborIndirect(vec3 norm, float depth, float reprojDepth, uint matId, ivec2 lastCoord, IndirectReservoir& resv) {   // :74-108
    vec3 pnorm; float pdepth; uint matHash;
    (void)depth;
    loadLastGeometryInfo(lastCoord, pnorm, pdepth, matHash);
    ivec2 coord = lastCoord / 2;
    if (inBound(coord, indSize())) {
      if (hash8bit(matId) == matHash) {
        if (dot(norm, pnorm) > 0.5f && reprojDepth < pdepth * 1.1f) {
          resv = lastIndirectResv[(size_t)coord.y * indSize().x + coord.x];
          return true;
        }
      }
    }
    return false;
  }
  static GISample newGISample() {                                                               // :110-115
    GISample g{};
    g.nv = {100.0f, 100.0f, 100.0f};
    g.L = {0, 0, 0};
    return g;
  }
  static bool GISampleValid(const GISample& g) { return g.nv.x < 1.1f && !hasNan(V(g.L)); }     // :117-119

  void pathTraceIndirect(State state, Ray ray, bool multiBounce, float& primSamplePdf, vec3& primWo, State& primState, GISample& giSample) {   // :129-226
    vec3 throughput = vec3(multiBounce ? 4.0f : 1.0f);
    primWo = -ray.direction;
    primState = state;
    giSample = newGISample();
    state.mat.albedo = vec3(1.0f);
    vec3 gL = V(giSample.L);

    for (int depth = 1; depth <= rtxState.maxDepth; depth++) {
      vec3 wo = -ray.direction;
      if (depth > 1 && rtxState.MIS > 0) {
        vec3 Li, wi;
        float lightPdf = SampleDirectLight(state, Li, wi);
        if (!IsPdfInvalid(lightPdf)) {
          float BSDFPdf = Pdf(state, wo, state.ffnormal, wi);
          float weight = MIS(lightPdf, BSDFPdf);
          gL += Li * BSDF(state, wo, state.ffnormal, wi) * absDot(state.ffnormal, wi) * throughput / lightPdf * weight;
        }
      }
      vec3 sampleWi;
      float samplePdf;
      vec3 sampleBSDF = Sample(state, wo, state.ffnormal, sampleWi, samplePdf);
      if (IsPdfInvalid(samplePdf)) break;

      if (depth > 1) {
        if (!multiBounce) { giSample.L = E(gL); return; }
        throughput *= sampleBSDF / samplePdf * absDot(state.ffnormal, sampleWi);
      } else {
        primSamplePdf = samplePdf;
        giSample.xv = E(state.position);
        giSample.nv = E(state.ffnormal);
      }
      ray.origin = OffsetRay(state.position, state.ffnormal);
      ray.direction = sampleWi;
      ClosestHit(ray);

      if (prd.hitT >= INFINITY_ - 1e-4f) {
        if (depth > 1) {
          float lightPdf;
          vec3 Li = EnvEval(sampleWi, lightPdf);
          float weight = MIS(samplePdf, lightPdf);
          gL += Li * throughput * weight;
        } else {
          giSample.xs = E(state.position + sampleWi * INFINITY_ * 0.8f);
          giSample.ns = E(-sampleWi);
        }
        break;
      }
      state = GetState(prd, ray.direction);
      GetMaterials(state, ray);

      if (state.isEmitter) {
        if (depth > 1) {
          float lightPdf;
          vec3 Li = LightEval(state, prd.hitT, sampleWi, lightPdf);
          float weight = MIS(samplePdf, lightPdf);
          gL += Li * throughput * weight;
        } else {
          giSample.xs = E(state.position);
          giSample.ns = E(state.ffnormal);
        }
        break;
      }
      if (depth == 1) {
        giSample.xs = E(state.position);
        giSample.ns = E(state.ffnormal);
      }
      // Russian roulette is compiled out (#ifndef RR while pathtrace.glsl:2 defines RR) (:218-224)
    }
    giSample.L = E(gL);
  }

  vec3 ReSTIRIndirect(float dist, float primSamplePdf, vec3 primWo, State primState, GISample giSample) {   // :228-268
    vec3 indirect = vec3(0.0f);
    IndirectReservoir resv{};
    resvReset(resv);
    if (rtxState.ReSTIRState == eTemporal || rtxState.ReSTIRState == eSpatiotemporal) {
      float reprojDepth = length(V(cam.lastPosition) - primState.position);
      ivec2 c2 = imageCoords * 2;
      ivec2 motionIdx(0, 0);
      if (c2.x >= 0 && c2.y >= 0 && c2.x < (int)rr.width && c2.y < (int)rr.height)
        motionIdx = ivec2(rr.motion[2 * ((size_t)c2.y * pitch + c2.x)], rr.motion[2 * ((size_t)c2.y * pitch + c2.x) + 1]);
      findTemporalNeighborIndirect(primState.ffnormal, dist, reprojDepth, primState.matID, motionIdx, resv);
    }
    float sampleWeight = 0.0f;
    if (GISampleValid(giSample)) {
      giSample.pHat = pHatIndirect(giSample);
      sampleWeight = (giSample.pHat / primSamplePdf);
      if (gisnan(sampleWeight) || sampleWeight < 0.0f) sampleWeight = 0.0f;
    }
    resvUpdate(resv, giSample, sampleWeight, rand());
    resvCheckValidity(resv);
    resvClamp(resv, rtxState.reservoirClamp * 2);
    thisIndirectResv[(size_t)imageCoords.y * indSize().x + imageCoords.x] = resv;   // saveNewReservoir

    giSample = resv.giSample;
    if (!resvInvalid(resv) && GISampleValid(giSample)) {
      vec3 primWi = normalize(V(giSample.xs) - V(giSample.xv));
      primState.mat.albedo = vec3(1.0f);
      indirect = V(giSample.L) * BSDF(primState, primWo, V(giSample.nv), primWi) * satDot(V(giSample.nv), primWi) * bigWIndirect(resv);
    }
    vec3 res = clampRadiance(indirect);
    res = HDRToLDR(res);
    return res;
  }

  void indirectMain(int gx, int gy) {                                                           // :270-309
    imageCoords = ivec2(gx, gy);
    if (!inBound(imageCoords, indSize())) return;
    prd.seed = tea((uint)indSize().x * (uint)gy + (uint)gx, rtxState.time);
    Ray ray = raySpawn(imageCoords, indSize());
    // TILED_MULTIBOUNCE (:283-288): invocation 0 of the 8x8 group draws once and shares the flag
    bool multiBounce;
    {
      int tx = (gx / 8) * 8, ty = (gy / 8) * 8;
      if (gx == tx && gy == ty) {
        multiBounce = rand() < 0.25f;
      } else {
        uint s0 = tea((uint)indSize().x * (uint)ty + (uint)tx, rtxState.time);
        multiBounce = rnd(s0) < 0.25f;
      }
    }
    State state;
    float depth;
    if (!getIndirectStateFromGBuffer(thisGbuffer, ray, state, depth)) {
      storeImg(rr.denoiseTemp[2], imageCoords, vec4(0.0f));
      return;
    }
    state.position += state.ffnormal * 2e-2f;
    float primSamplePdf = 0.f; vec3 primWo; State primState; GISample giSample;
    pathTraceIndirect(state, ray, multiBounce, primSamplePdf, primWo, primState, giSample);
    vec3 pixelColor = ReSTIRIndirect(depth, primSamplePdf, primWo, primState, giSample);
    pixelColor = clampRadiance(pixelColor);
    storeImg(rr.denoiseTemp[2], imageCoords, vec4(pixelColor, 1.0f));
  }

  // ---- denoise_common.glsl ----------------------------------------------------------------------
  Ray raySpawnDenoise(ivec2 coord, ivec2 sizeImage) {                                           // :27-35 (direction NOT re-normalised)
    const vec2 pixelCenter = vec2((float)coord.x, (float)coord.y) + 0.5f;
    const vec2 inUV = pixelCenter / vec2((float)sizeImage.x, (float)sizeImage.y);
    vec2 d = inUV * 2.0f - 1.0f;
    const mat4& VI = *reinterpret_cast<const mat4*>(&cam.viewInverse);
    const mat4& PI = *reinterpret_cast<const mat4*>(&cam.projInverse);
    vec3 origin(VI.m[12], VI.m[13], VI.m[14]);
    vec4 target = mul(PI, vec4(d.x, d.y, 1, 1));
    vec3 direction = mulDir(VI, normalize(target.xyz()));
    Ray r; r.origin = origin; r.direction = direction;
    return r;
  }
  void loadThisGeometry(ivec2 coord, vec3& normal, vec3& pos, uint& matHash, ivec2 imageSize) {  // :42-47
    uvec4 gInfo = loadG(thisGbuffer, coord);
    normal = decompress_unit_vec(gInfo.y);
    Ray ray = raySpawnDenoise(coord, imageSize);
    pos = ray.origin + ray.direction * uintBitsToFloat(gInfo.x);
    matHash = gInfo.w & 0xFF000000;
  }
  static float Gaussian5x5(int a, int b) {
    static const float G[5][5] = {{.0030f, .0133f, .0219f, .0133f, .0030f},
                                  {.0133f, .0596f, .0983f, .0596f, .0133f},
                                  {.0219f, .0983f, .1621f, .0983f, .0219f},
                                  {.0133f, .0596f, .0983f, .0596f, .0133f},
                                  {.0030f, .0133f, .0219f, .0133f, .0030f}};
    return G[a][b];
  }
  // denoise_direct.comp:19-71 (indirectMode=false) / denoise_indirect.comp:23-75 (indirectMode=true)
  vec3 waveletFilter(const std::vector<vec4>& inImage, ivec2 coord, vec3 norm, vec3 pos, uint matHash,
                     float sigLumin, float sigNormal, float sigDepth, int level, bool indirectMode) {
    if (matHash == InvalidMatId) return vec3(0.0f);
    int step = 1 << level;
    vec3 sum = vec3(0.0f);
    float sumWeight = 0.0f;
    vec3 color = loadImg(inImage, coord).xyz();
    ivec2 bound = indirectMode ? indSize() : size();
    for (int j = -2; j <= 2; j++) {
      for (int i = -2; i <= 2; i++) {
        ivec2 q = coord + ivec2(i, j) * step;
        if (q.x >= bound.x || q.y >= bound.y || q.x < 0 || q.y < 0) continue;
        vec3 normQ, posQ; uint matHashQ;
        if (indirectMode) loadThisGeometry(q * 2, normQ, posQ, matHashQ, indSize());
        else loadThisGeometry(q, normQ, posQ, matHashQ, size());
        vec3 colorQ = loadImg(inImage, q).xyz();
        if (matHash != matHashQ || matHashQ == InvalidMatId) continue;
        float var = sigLumin;
        float distColor = indirectMode ? dot(color - colorQ, color - colorQ) : gabs(luminance(color) - luminance(colorQ));
        float wColor = eid_expf(-distColor / var) + 1e-2f;
        float distNorm2 = dot(norm - normQ, norm - normQ);
        float wNorm = gmin(1.0f, eid_expf(-distNorm2 / sigNormal));
        float distPos2 = dot(pos - posQ, pos - posQ);
        float wDepth = eid_expf(-distPos2 / sigDepth) + 1e-2f;
        float weight = wColor * wNorm * wDepth * Gaussian5x5(i + 2, j + 2);
        sum += colorQ * weight;
        sumWeight += weight;
      }
    }
    vec3 res = (sumWeight < 1e-5f) ? vec3(0.0f) : sum / sumWeight;
    if (hasNan(res) || res.x < 0 || res.y < 0 || res.z < 0 || res.x > 1e8f || res.y > 1e8f || res.z > 1e8f) res = vec3(0.0f);
    return res;
  }
  // #if DENOISER_DIRECT_BILATERAL (denoise_direct.comp:73-137, Radius 4, with the spatial term) / #if DENOISER_INDIRECT_BILATERAL
  // (denoise_indirect.comp:77-130, Radius 5, no spatial term, sky pixels return 0 at once)
  vec3 bilateralFilter(const std::vector<vec4>& inImage, ivec2 coord, vec3 norm, vec3 pos, uint matHash,
                       float sigLumin, float sigNormal, float sigDepth, bool indirectMode) {
    const int Radius = indirectMode ? 5 : 4;
    if (indirectMode && matHash == InvalidMatId) return vec3(0.0f);
    vec3 sum = vec3(0.0f);
    float sumWeight = 0.0f;
    vec3 color = loadImg(inImage, coord).xyz();
    ivec2 bound = indirectMode ? indSize() : size();
    for (int j = -Radius; j <= Radius; j++) {
      for (int i = -Radius; i <= Radius; i++) {
        ivec2 q = coord + ivec2(i, j);
        if (q.x >= bound.x || q.y >= bound.y || q.x < 0 || q.y < 0) continue;
        vec3 normQ, posQ; uint matHashQ;
        if (indirectMode) loadThisGeometry(q * 2, normQ, posQ, matHashQ, indSize());
        else loadThisGeometry(q, normQ, posQ, matHashQ, size());
        vec3 colorQ = loadImg(inImage, q).xyz();
        if (matHash != matHashQ || matHashQ == InvalidMatId) continue;
        float var = sigLumin;
        float distColor = dot(color - colorQ, color - colorQ);
        float wColor = eid_expf(-distColor / var) + 1e-2f;
        float distNorm2 = dot(norm - normQ, norm - normQ);
        float wNorm = gmin(1.0f, eid_expf(-distNorm2 / sigNormal));
        float distPos2 = dot(pos - posQ, pos - posQ);
        float wDepth = eid_expf(-distPos2 / sigDepth) + 1e-2f;
        float weight = wColor * wNorm * wDepth;
        if (!indirectMode) {
          float dist2 = float(i * i + j * j);
          float wDist = eid_expf(-dist2 / 10.0f) + 1e-2f;
          weight = weight * wDist;
        }
        sum += colorQ * weight;
        sumWeight += weight;
      }
    }
    vec3 res = (sumWeight < 1e-5f) ? vec3(0.0f) : sum / sumWeight;
    if (hasNan(res) || res.x < 0 || res.y < 0 || res.z < 0 || res.x > 1e8f || res.y > 1e8f || res.z > 1e8f) res = vec3(0.0f);
    return res;
  }
  void denoiseDirectMain(int gx, int gy) {                                                       // denoise_direct.comp:139-173
    ivec2 coord(gx, gy);
    if (!inBound(coord, size())) return;
    vec3 norm, pos; uint matHash;
    loadThisGeometry(coord, norm, pos, matHash, size());
    const float sl = rtxState.sigLuminDirect, sn = rtxState.sigNormalDirect, sd = rtxState.sigDepthDirect;
    auto& A = rr.denoiseTemp[0]; auto& B = rr.denoiseTemp[1];
    if (rr.variant & EID_VARIANT_DIRECT_BILATERAL) {
      vec3 res = bilateralFilter(A, coord, norm, pos, matHash, sl, sn, sd, false);
      res = LDRToHDR(res);
      storeImg(rr.directResult, coord, vec4(res, 1.0f));
      return;
    }
    if (rtxState.denoiseLevel == 0) storeImg(A, coord, vec4(waveletFilter(rr.directResult, coord, norm, pos, matHash, sl, sn, sd, 0, false), 1.0f));
    else if (rtxState.denoiseLevel == 1) storeImg(B, coord, vec4(waveletFilter(A, coord, norm, pos, matHash, sl, sn, sd, 1, false), 1.0f));
    else if (rtxState.denoiseLevel == 2) storeImg(A, coord, vec4(waveletFilter(B, coord, norm, pos, matHash, sl, sn, sd, 2, false), 1.0f));
    else if (rtxState.denoiseLevel == 3) {
      vec3 res = waveletFilter(A, coord, norm, pos, matHash, sl, sn, sd, 3, false);
      res = LDRToHDR(res);
      storeImg(rr.directResult, coord, vec4(res, 1.0f));
    }
  }
  void denoiseIndirectMain(int gx, int gy) {                                                     // denoise_indirect.comp:132-172
    ivec2 coord(gx, gy);
    if (coord.x >= indSize().x || coord.y >= indSize().y || rtxState.denoise == 0) return;
    vec3 norm, pos; uint matHash;
    loadThisGeometry(coord * 2, norm, pos, matHash, indSize());
    const float sl = rtxState.sigLuminIndirect, sn = rtxState.sigNormalIndirect, sd = rtxState.sigDepthIndirect;
    auto& A = rr.denoiseTemp[2]; auto& B = rr.denoiseTemp[3];
    if (rr.variant & EID_VARIANT_INDIRECT_BILATERAL) {
      vec3 res = bilateralFilter(A, coord, norm, pos, matHash, sl, sn, sd, true);
      res = LDRToHDR(res);
      storeImg(B, coord, vec4(res, 1.0f));
      return;
    }
    if (rtxState.denoiseLevel == 0) storeImg(B, coord, vec4(waveletFilter(A, coord, norm, pos, matHash, sl, sn, sd, 0, true), 1.0f));
    else if (rtxState.denoiseLevel == 1) storeImg(A, coord, vec4(waveletFilter(B, coord, norm, pos, matHash, sl, sn, sd, 1, true), 1.0f));
    else if (rtxState.denoiseLevel == 2) storeImg(rr.indirectResult, coord, vec4(waveletFilter(A, coord, norm, pos, matHash, sl, sn, sd, 2, true), 1.0f));
    else if (rtxState.denoiseLevel == 3) storeImg(A, coord, vec4(waveletFilter(rr.indirectResult, coord, norm, pos, matHash, sl, sn, sd, 3, true), 1.0f));
    else if (rtxState.denoiseLevel == 4) {
      vec3 res = waveletFilter(A, coord, norm, pos, matHash, sl, sn, sd, 4, true);
      res = LDRToHDR(res);
      storeImg(B, coord, vec4(res, 1.0f));
    }
  }
  void composeMain(int gx, int gy) {                                                             // compose.comp:23-42
    ivec2 coord(gx, gy);
    if (coord.x >= rtxState.size.x || coord.y >= rtxState.size.y) return;
    const auto& indSrc = (rtxState.denoise > 0) ? rr.denoiseTemp[3] : rr.denoiseTemp[2];
    if (rtxState.modulate == 0) {
      storeImg(rr.indirectResult, coord, loadImg(indSrc, coord / 2));
    } else {
      vec3 albedo = unpackUnorm4x8(loadG(thisGbuffer, coord).w).xyz();
      vec3 direct = loadImg(rr.directResult, coord).xyz() * albedo;
      vec3 indirect = loadImg(indSrc, coord / 2).xyz() * albedo;
      storeImg(rr.directResult, coord, vec4(direct, 1.0f));
      storeImg(rr.indirectResult, coord, vec4(indirect, 1.0f));
    }
  }
};

// =================================================================================================
// Renderer (renderer.cpp)
// =================================================================================================
void Renderer::create(const Scene* s, uint32_t w, uint32_t h) {    // renderer.cpp:97-148, 227-302; history zero-initialised
  scene = s; width = w; height = h;
  size_t n = (size_t)w * h, ni = (size_t)(w / 2) * (h / 2);
  for (int i = 0; i < 2; ++i) {
    gbuffer[i].assign(n, uvec4());
    directResv[i].assign(n, DirectReservoir{});
    if (i == 0) tempDirectResv.assign(n, DirectReservoir{});
    indirectResv[i].assign(ni, IndirectReservoir{});
  }
  motion.assign(2 * n, 0);
  directResult.assign(n, vec4()); indirectResult.assign(n, vec4());
  for (auto& t : denoiseTemp) t.assign(n, vec4());
  envConstant = vec3(0.f);
}

template <class F>
static void dispatch(int w, int h, int y0, int y1, F&& f) {
  // 8x8 work groups like the reference's vkCmdDispatch; groups are independent.
  int gy0 = y0 / 8, gy1 = (y1 + 7) / 8;
  int gxn = (w + 7) / 8;
#pragma omp parallel for schedule(dynamic, 4) collapse(2)
  for (int gy = gy0; gy < gy1; ++gy)
    for (int gx = 0; gx < gxn; ++gx)
      for (int ly = 0; ly < 8; ++ly)
        for (int lx = 0; lx < 8; ++lx) {
          int x = gx * 8 + lx, y = gy * 8 + ly;
          if (x < w && y < h && y >= y0 && y < y1) f(x, y);
        }
}

static double nowMs() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

void Renderer::runDirect(const RtxState& st, int frames, int y0, int y1) {
  int set = (frames + 1) % 2;   // renderer.cpp:157
  double t0 = nowMs();
  if (st.ReSTIRState == eSpatial || st.ReSTIRState == eSpatiotemporal) {
    // spatial reuse reads tempDirectResv of pixels up to one row / column away: a first dispatch (one halo row beyond the band) carries
    // every pixel up to its write of tempDirectResv, so the second one reads completed entries whatever the group order.  The first
    // dispatch's rays are not counted (they are the second one's, traced twice).
    const uint64_t c0 = closestRays, a0 = anyRays, p0 = primaryHits;
    dispatch(st.size.x, st.size.y, y0 > 0 ? y0 - 1 : 0, y1 < st.size.y ? y1 + 1 : y1, [&](int x, int y) { Ctx c(*scene, *this, st, set); c.primeOnly = true; c.directMain(x, y); });
    closestRays = c0; anyRays = a0; primaryHits = p0;
  }
  if (variant & EID_VARIANT_DIRECT_SPLIT) {     // direct_gen.comp then direct_reuse.comp in place of direct_stage.comp
    dispatch(st.size.x, st.size.y, y0, y1, [&](int x, int y) { Ctx c(*scene, *this, st, set); c.directGenMain(x, y); });
    dispatch(st.size.x, st.size.y, y0, y1, [&](int x, int y) { Ctx c(*scene, *this, st, set); c.directReuseMain(x, y); });
    kernelMs[0] += nowMs() - t0;
    return;
  }
  dispatch(st.size.x, st.size.y, y0, y1, [&](int x, int y) { Ctx c(*scene, *this, st, set); c.directMain(x, y); });
  kernelMs[0] += nowMs() - t0;
}
void Renderer::runIndirect(const RtxState& st, int frames, int y0, int y1) {
  int set = (frames + 1) % 2;
  double t0 = nowMs();
  dispatch(st.size.x / 2, st.size.y / 2, y0, y1, [&](int x, int y) { Ctx c(*scene, *this, st, set); c.indirectMain(x, y); });
  kernelMs[1] += nowMs() - t0;
}
void Renderer::runPost(const RtxState& st, int frames) {
  int set = (frames + 1) % 2;
  RtxState cState = st;
  double t0 = nowMs();
  if (st.denoise > 0 && (variant & EID_VARIANT_DIRECT_BILATERAL)) {   // #if DENOISER_DIRECT_BILATERAL: one dispatch, the caller's push constants (renderer.cpp:186-188)
    dispatch(st.size.x, st.size.y, 0, st.size.y, [&](int x, int y) { Ctx c(*scene, *this, st, set); c.denoiseDirectMain(x, y); });
  } else if (st.denoise > 0) {
    for (int i = 0; i < 4; i++) {   // renderer.cpp:178-189
      cState.denoiseLevel = i;
      dispatch(st.size.x, st.size.y, 0, st.size.y, [&](int x, int y) { Ctx c(*scene, *this, cState, set); c.denoiseDirectMain(x, y); });
    }
  }
  double t1 = nowMs();
  kernelMs[2] += t1 - t0;
  if (st.denoise > 0 && (variant & EID_VARIANT_INDIRECT_BILATERAL)) {
    dispatch(st.size.x / 2, st.size.y / 2, 0, st.size.y / 2, [&](int x, int y) { Ctx c(*scene, *this, cState, set); c.denoiseIndirectMain(x, y); });
  } else if (st.denoise > 0) {
    for (int i = 0; i < 5; i++) {   // renderer.cpp:191-202
      cState.denoiseLevel = i;
      dispatch(st.size.x / 2, st.size.y / 2, 0, st.size.y / 2, [&](int x, int y) { Ctx c(*scene, *this, cState, set); c.denoiseIndirectMain(x, y); });
    }
  }
  double t2 = nowMs();
  kernelMs[3] += t2 - t1;
  // compose is dispatched with the un-modified push constants (renderer.cpp:161, 204-205)
  dispatch(st.size.x, st.size.y, 0, st.size.y, [&](int x, int y) { Ctx c(*scene, *this, cState, set); c.composeMain(x, y); });
  kernelMs[4] += nowMs() - t2;
}
void Renderer::run(const RtxState& st, int frames) {   // renderer.cpp:154-206, strict serial order K1..K5
  lastSet = (frames + 1) % 2;
  runDirect(st, frames, 0, st.size.y);
  runIndirect(st, frames, 0, st.size.y / 2);
  runPost(st, frames);
}

// ---- function taps: the restated shader functions, one call per item, for the bit-for-bit comparison against the reference's own
// GLSL text compiled as C++ (oracle/ref_shim/ref_glsl.cpp, same numbering and arity) and for the golden vectors made from it
int fn_arity(int which, int* nin, int* nout) {
  static const int A[][2] = {{2, 2}, {2, 1}, {3, 2}, {3, 6}, {3, 3}, {3, 3}, {14, 3}, {14, 1}, {14, 7}, {30, 27}, {39, 19}, {4, 3}, {6, 3}, {2, 1}, {2, 3}};
  if (which < 0 || which >= (int)(sizeof(A) / sizeof(A[0]))) return -1;
  *nin = A[which][0]; *nout = A[which][1];
  return 0;
}
static State tapState(const float* p) { State s{}; s.mat.albedo = vec3(p[0], p[1], p[2]); s.mat.roughness = p[3]; s.mat.metallic = p[4]; return s; }
static vec3 tv3(const float* p) { return vec3(p[0], p[1], p[2]); }
static void tput(float* o, vec3 v) { o[0] = v.x; o[1] = v.y; o[2] = v.z; }
static DirectReservoir tapD(const float* p) {
  DirectReservoir r{}; r.lightSample.Li = E(tv3(p)); r.lightSample.wi = E(tv3(p + 3)); r.lightSample.dist = p[6]; r.num = floatBitsToUint(p[7]); r.weight = p[8]; return r;
}
static void tputD(float* o, const DirectReservoir& r) {
  tput(o, V(r.lightSample.Li)); tput(o + 3, V(r.lightSample.wi)); o[6] = r.lightSample.dist; o[7] = uintBitsToFloat(r.num); o[8] = r.weight;
}
static GISample tapG(const float* p) { GISample g{}; g.L = E(tv3(p)); g.xv = E(tv3(p + 3)); g.nv = E(tv3(p + 6)); g.xs = E(tv3(p + 9)); g.ns = E(tv3(p + 12)); g.pHat = p[15]; return g; }
int fn(int which, const float* in, int n, float* out) {
  int ni, no;
  if (fn_arity(which, &ni, &no)) return -1;
  for (int i = 0; i < n; ++i) {
    const float* p = in + (size_t)i * ni;
    float* o = out + (size_t)i * no;
    switch (which) {
      case 0: { vec2 d = toConcentricDisk(vec2(p[0], p[1])); o[0] = d.x; o[1] = d.y; break; }
      case 1: o[0] = powerHeuristic(p[0], p[1]); break;
      case 2: { vec2 uv = Ctx::GetSphericalUv(tv3(p)); o[0] = uv.x; o[1] = uv.y; break; }
      case 3: { vec3 t, b; Ctx::CreateCoordinateSystem(tv3(p), t, b); tput(o, t); tput(o + 3, b); break; }
      case 4: tput(o, HDRToLDR(tv3(p))); break;
      case 5: tput(o, LDRToHDR(tv3(p))); break;
      case 6: tput(o, metallicWorkflowBSDF(tapState(p), tv3(p + 5), tv3(p + 8), tv3(p + 11))); break;
      case 7: o[0] = metallicWorkflowPdf(tapState(p), tv3(p + 5), tv3(p + 8), tv3(p + 11)); break;
      case 8: { vec3 bsdf(0.0f), dir(0.0f); o[0] = metallicWorkflowSample(tapState(p), tv3(p + 5), tv3(p + 8), tv3(p + 11), bsdf, dir); tput(o + 1, bsdf); tput(o + 4, dir); break; }
      case 9: {
        DirectReservoir r = tapD(p);
        LightSample s{}; s.Li = E(tv3(p + 9)); s.wi = E(tv3(p + 12)); s.dist = p[15];
        resvUpdate(r, s, p[16], p[17]); tputD(o, r);
        resvMerge(r, tapD(p + 18), p[27]); tputD(o + 9, r);
        resvCheckValidity(r); resvClamp(r, (int)p[28]); tputD(o + 18, r);
        break;
      }
      case 10: {
        IndirectReservoir r{}; r.giSample = tapG(p); r.num = floatBitsToUint(p[16]); r.weight = p[17]; r.bigW = p[18];
        resvUpdate(r, tapG(p + 19), p[35], p[36]); resvCheckValidity(r); resvClamp(r, (int)p[37]);
        tput(o, V(r.giSample.L)); tput(o + 3, V(r.giSample.xv)); tput(o + 6, V(r.giSample.nv)); tput(o + 9, V(r.giSample.xs)); tput(o + 12, V(r.giSample.ns));
        o[15] = r.giSample.pHat; o[16] = uintBitsToFloat(r.num); o[17] = r.weight; o[18] = r.bigW;
        break;
      }
      case 11: tput(o, post_toneMap(tv3(p), p[3])); break;
      case 12: tput(o, OffsetRay(tv3(p), tv3(p + 3))); break;
      case 13: o[0] = uintBitsToFloat(tea(floatBitsToUint(p[0]), floatBitsToUint(p[1]))); break;
      case 14: { uint s = floatBitsToUint(p[0]); float a = rnd(s); float b = rnd(s); o[0] = a; o[1] = b; o[2] = uintBitsToFloat(s); break; }
    }
  }
  return 0;
}

// scene-dependent taps (same numbering as ref_ctx_fn of oracle/ref_shim/ref_glsl.cpp): 0 SampleDirectLightNoVisibility, 1 LightEval,
// 2 EnvEval, 3 EnvRadiance, 4 raySpawn, 5 clampRadiance, 6 Sample
int ctx_fn(Renderer& rr, const RtxState& st, int which, const float* in, int n, float* out) {
  static const int A[][2] = {{4, 9}, {9, 4}, {3, 4}, {3, 3}, {4, 6}, {3, 3}, {12, 8}};
  if (which < 0 || which >= 7) return -1;
  const int ni = A[which][0], no = A[which][1];
  for (int i = 0; i < n; ++i) {
    const float* p = in + (size_t)i * ni;
    float* o = out + (size_t)i * no;
    Ctx c(*rr.scene, rr, st, 0);
    switch (which) {
      case 0: {
        c.prd.seed = floatBitsToUint(p[0]);
        LightSample ls{};
        o[0] = c.SampleDirectLightNoVisibility(tv3(p + 1), ls);
        tput(o + 1, V(ls.Li)); tput(o + 4, V(ls.wi)); o[7] = ls.dist; o[8] = uintBitsToFloat(c.prd.seed);
        break;
      }
      case 1: { State s{}; s.matID = floatBitsToUint(p[0]); s.ffnormal = tv3(p + 5); s.area = p[8]; float pdf = 0.0f; tput(o, c.LightEval(s, p[1], tv3(p + 2), pdf)); o[3] = pdf; break; }
      case 2: { float pdf = 0.0f; tput(o, c.EnvEval(tv3(p), pdf)); o[3] = pdf; break; }
      case 3: tput(o, c.EnvRadiance(tv3(p))); break;
      case 4: { Ray r = c.raySpawn(ivec2((int)p[0], (int)p[1]), ivec2((int)p[2], (int)p[3])); tput(o, r.origin); tput(o + 3, r.direction); break; }
      case 5: tput(o, c.clampRadiance(tv3(p))); break;
      case 6: {
        c.prd.seed = floatBitsToUint(p[0]);
        State s = tapState(p + 1); vec3 L(0.0f); float pdf = 0.0f;
        tput(o, c.Sample(s, tv3(p + 6), tv3(p + 9), L, pdf)); tput(o + 3, L); o[6] = pdf; o[7] = uintBitsToFloat(c.prd.seed);
        break;
      }
    }
  }
  return 0;
}

}  // namespace orc

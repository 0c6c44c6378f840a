// shade.cuh — device-side arithmetic of the reference's shader includes, written for CUDA registers:
//   random.glsl (tea/pcg/rand), common.glsl (OffsetRay, toConcentricDisk, powerHeuristic, HDR<->LDR),
//   shade_state.glsl (GetState), gltf_material.glsl (GetMaterials, texture-less),
//   pbr_metallicworkflow.glsl (BSDF / Pdf / Sample), reservoir.glsl, pathtrace.glsl (light sampling,
//   Occlusion, clampRadiance, raySpawn, G-buffer decode).
// RNG draw order and floating-point evaluation order follow SURVEY.md §8 a.3 / DESIGN.md §3 exactly;
// every function names the reference lines it implements.
#pragma once
#include "dmath.cuh"
#include "sunsky.cuh"
#include "trace.cuh"

namespace eid {

#define EID_INFINITY 1e28f          // globals.glsl:29
#define EID_PI 3.14159265358979323846f
#define EID_INVALID_PDF (-1.0f)     // common.glsl:30
#define EID_INVALID_MAT 0xff000000u // globals.glsl:106

// ---- random.glsl ---------------------------------------------------------------------------------
DEV uint32_t tea(uint32_t v0, uint32_t v1) {            // :34-48
  uint32_t s0 = 0;
#pragma unroll
  for (int n = 0; n < 16; n++) {
    s0 += 0x9e3779b9u;
    v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + s0) ^ ((v1 >> 5) + 0xc8013ea4u);
    v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + s0) ^ ((v0 >> 5) + 0x7e95761eu);
  }
  return v0;
}
DEV float rnd(uint32_t& seed) {                         // pcg :59-65 + rand :98-102
  uint32_t prev = seed * 747796405u + 2891336453u;
  uint32_t word = ((prev >> ((prev >> 28u) + 4u)) ^ prev) * 277803737u;
  seed = prev;
  uint32_t r = (word >> 22u) ^ word;
  return __fsub_rn(__uint_as_float(0x3f800000u | (r >> 9)), 1.0f);
}

// ---- common.glsl ---------------------------------------------------------------------------------
DEV f3 offsetRay(f3 p, f3 n) {                          // :98-113 (Ray Tracing Gems ch. 6)
  const float intScale = 256.0f, floatScale = 1.0f / 65536.0f, origin = 1.0f / 32.0f;
  int ox = f2i_sat(__fmul_rn(intScale, n.x)), oy = f2i_sat(__fmul_rn(intScale, n.y)), oz = f2i_sat(__fmul_rn(intScale, n.z));
  float ix = __int_as_float(__float_as_int(p.x) + ((p.x < 0) ? -ox : ox));
  float iy = __int_as_float(__float_as_int(p.y) + ((p.y < 0) ? -oy : oy));
  float iz = __int_as_float(__float_as_int(p.z) + ((p.z < 0) ? -oz : oz));
  return mk3(fabsf(p.x) < origin ? __fadd_rn(p.x, __fmul_rn(floatScale, n.x)) : ix,
             fabsf(p.y) < origin ? __fadd_rn(p.y, __fmul_rn(floatScale, n.y)) : iy,
             fabsf(p.z) < origin ? __fadd_rn(p.z, __fmul_rn(floatScale, n.z)) : iz);
}
DEV void toConcentricDisk(float rx_, float ry_, float& dx, float& dy) {   // :171-175 (polar, not Shirley)
  float rx = __fsqrt_rn(rx_);
  float theta = __fmul_rn(__fmul_rn(ry_, 2.0f), EID_PI);
  float s, c;
  eid_sincosf(theta, &s, &c);
  dx = __fmul_rn(c, rx); dy = __fmul_rn(s, rx);
}
DEV float powerHeuristic(float f, float g) { float f2 = __fmul_rn(f, f); return __fdiv_rn(f2, __fadd_rn(f2, __fmul_rn(g, g))); }   // :177-180
DEV f3 hdrToLdr(f3 c) { return c / (c + 1.0f); }       // :194-196
DEV f3 ldrToHdr(f3 c) { return c / (1.01f - c); }      // :198-200

// ---- globals.glsl:64-104 -------------------------------------------------------------------------
struct Material { f3 albedo, emission; float metallic, ior, roughness, transmission; };
struct State {
  f3 position, normal, ffnormal;
  f3 tangent, bitangent;   // only filled (and only needed) when the material has a normal map
  float u, v;          // texCoord
  float eta, area;
  uint32_t matID;
  bool isEmitter;
  Material mat;
};

// ---- pbr_metallicworkflow.glsl -------------------------------------------------------------------
struct M3 { f3 c0, c1, c2; };   // column-major mat3
DEV f3 m3mul(const M3& m, f3 v) { return (m.c0 * v.x + m.c1 * v.y) + m.c2 * v.z; }
DEV M3 localRefMatrix(f3 n) {                           // :11-16
  f3 t = (fabsf(n.y) > 0.9999f) ? mk3(0.f, 0.f, 1.f) : mk3(0.f, 1.f, 0.f);
  f3 b = norm3(cross3(n, t));
  t = cross3(b, n);
  M3 m = {t, b, n};
  return m;
}
DEV M3 m3inverse(const M3& m) {                          // GLSL inverse(mat3), cofactor form (contract)
  float a00 = m.c0.x, a01 = m.c0.y, a02 = m.c0.z, a10 = m.c1.x, a11 = m.c1.y, a12 = m.c1.z, a20 = m.c2.x, a21 = m.c2.y, a22 = m.c2.z;
  float b01 = __fsub_rn(__fmul_rn(a22, a11), __fmul_rn(a12, a21));
  float b11 = __fsub_rn(__fmul_rn(a12, a20), __fmul_rn(a22, a10));
  float b21 = __fsub_rn(__fmul_rn(a21, a10), __fmul_rn(a11, a20));
  float det = __fadd_rn(__fadd_rn(__fmul_rn(a00, b01), __fmul_rn(a01, b11)), __fmul_rn(a02, b21));
  float id = __fdiv_rn(1.0f, det);
  M3 r;
  r.c0 = mk3(__fmul_rn(b01, id), __fmul_rn(__fsub_rn(__fmul_rn(a02, a21), __fmul_rn(a22, a01)), id), __fmul_rn(__fsub_rn(__fmul_rn(a12, a01), __fmul_rn(a02, a11)), id));
  r.c1 = mk3(__fmul_rn(b11, id), __fmul_rn(__fsub_rn(__fmul_rn(a22, a00), __fmul_rn(a02, a20)), id), __fmul_rn(__fsub_rn(__fmul_rn(a02, a10), __fmul_rn(a12, a00)), id));
  r.c2 = mk3(__fmul_rn(b21, id), __fmul_rn(__fsub_rn(__fmul_rn(a01, a20), __fmul_rn(a21, a00)), id), __fmul_rn(__fsub_rn(__fmul_rn(a11, a00), __fmul_rn(a01, a10)), id));
  return r;
}
DEV float satDot(f3 a, f3 b) { return gmax(dot3(a, b), 0.0f); }
DEV float absDot(f3 a, f3 b) { return fabsf(dot3(a, b)); }
DEV f3 sampleHemisphereCosine(f3 n, float r0, float r1) {   // :22-26 + localToWorld :18-20
  float dx, dy;
  toConcentricDisk(r0, r1, dx, dy);
  float z = __fsqrt_rn(__fsub_rn(1.0f, __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy))));
  return norm3(m3mul(localRefMatrix(n), mk3(dx, dy, z)));
}
DEV f3 fresnelSchlick(float cosTheta, f3 f0) {           // :36-41
  float c4 = __fsub_rn(1.0f, cosTheta);
  c4 = __fmul_rn(c4, c4);
  c4 = __fmul_rn(c4, c4);
  return mix3(f0, mk3(1.0f), __fmul_rn(c4, __fsub_rn(1.0f, cosTheta)));
}
DEV float schlickG(float cosTheta, float alpha) {        // :43-46
  float a = __fmul_rn(alpha, 0.5f);
  return __fdiv_rn(cosTheta, __fadd_rn(__fmul_rn(cosTheta, __fsub_rn(1.0f, a)), a));
}
DEV float smithG(float cosWo, float cosWi, float alpha) { return __fmul_rn(schlickG(fabsf(cosWo), alpha), schlickG(fabsf(cosWi), alpha)); }   // :48-50
DEV float gtr2Distrib(float cosTheta, float alpha) {     // :52-61
  if (cosTheta < 1e-6f) return 0.0f;
  float aa = __fmul_rn(alpha, alpha);
  float denom = __fadd_rn(__fmul_rn(__fmul_rn(cosTheta, cosTheta), __fsub_rn(aa, 1.0f)), 1.0f);
  denom = __fmul_rn(__fmul_rn(denom, denom), EID_PI);
  return __fdiv_rn(aa, denom);
}
DEV float gtr2Pdf(f3 n, f3 m, f3 wo, float alpha) {      // :63-65
  return __fdiv_rn(__fmul_rn(__fmul_rn(gtr2Distrib(dot3(n, m), alpha), schlickG(dot3(n, wo), alpha)), absDot(m, wo)), absDot(n, wo));
}
DEV f3 gtr2Sample(f3 n, f3 wo, float alpha, float r0, float r1) {   // :67-84
  M3 transMat = localRefMatrix(n);
  M3 transInv = m3inverse(transMat);
  f3 vh = norm3(m3mul(transInv, wo) * mk3(alpha, alpha, 1.0f));
  float lenSq = __fadd_rn(__fmul_rn(vh.x, vh.x), __fmul_rn(vh.y, vh.y));
  f3 t = lenSq > 0.0f ? mk3(-vh.y, vh.x, 0.0f) / __fsqrt_rn(lenSq) : mk3(1.0f, 0.0f, 0.0f);
  f3 b = cross3(vh, t);
  float px, py;
  toConcentricDisk(r0, r1, px, py);
  float s = __fmul_rn(0.5f, __fadd_rn(vh.z, 1.0f));
  py = __fadd_rn(__fmul_rn(__fsub_rn(1.0f, s), __fsqrt_rn(__fsub_rn(1.0f, __fmul_rn(px, px)))), __fmul_rn(s, py));
  float pp = __fadd_rn(__fmul_rn(px, px), __fmul_rn(py, py));
  f3 h = (t * px + b * py) + vh * __fsqrt_rn(gmax(0.0f, __fsub_rn(1.0f, pp)));
  h = mk3(__fmul_rn(h.x, alpha), __fmul_rn(h.y, alpha), gmax(0.0f, h.z));
  return norm3(m3mul(transMat, h));
}
// metallicWorkflowBSDF :86-106 (== the value of metallicWorkflowEval :123-144; its pdf output is dead on the live path)
DEV f3 bsdfEval(f3 albedo, float roughness, float metallic, f3 n, f3 wo, f3 wi) {
  const float PiInv = 1.0f / EID_PI;
  float alpha = roughness;
  f3 h = norm3(wo + wi);
  float cosO = dot3(n, wo), cosI = dot3(n, wi);
  if (__fmul_rn(cosI, cosO) < 1e-7f) return mk3(0.0f);
  f3 f = fresnelSchlick(dot3(h, wo), mix3(mk3(.08f), albedo, metallic));
  float g = smithG(cosO, cosI, alpha);
  float d = gtr2Distrib(dot3(n, h), alpha);
  float spec = __fdiv_rn(__fmul_rn(g, d), __fmul_rn(__fmul_rn(4.0f, cosI), cosO));
  return mix3((albedo * PiInv) * __fsub_rn(1.0f, metallic), mk3(spec), f);
}
DEV float bsdfPdf(float roughness, float metallic, f3 n, f3 wo, f3 wi) {   // :108-121
  const float PiInv = 1.0f / EID_PI;
  float alpha = roughness;
  f3 h = norm3(wo + wi);
  return mixf(__fmul_rn(satDot(n, wi), PiInv), __fdiv_rn(gtr2Pdf(n, h, wo, alpha), __fmul_rn(4.0f, absDot(h, wo))),
              __fdiv_rn(1.0f, __fsub_rn(2.0f, metallic)));
}
// Sample (pathtrace.glsl:36-38) -> metallicWorkflowSample (:146-166): three draws, returns pdf (or InvalidPdf)
// metallicWorkflowSample :146-166 with the three random numbers given
DEV float bsdfSampleR(const State& s, f3 n, f3 wo, float r0, float r1, float r2, f3& bsdf, f3& dir) {
  float roughness = s.mat.roughness, metallic = s.mat.metallic, alpha = roughness;
  if (r2 > __fdiv_rn(1.0f, __fsub_rn(2.0f, metallic))) {
    dir = sampleHemisphereCosine(n, r0, r1);
  } else {
    f3 h = gtr2Sample(n, wo, alpha, r0, r1);
    dir = -reflect3(wo, h);
  }
  if (dot3(n, dir) < 0.0f) { bsdf = mk3(0.f); return EID_INVALID_PDF; }
  bsdf = bsdfEval(s.mat.albedo, roughness, metallic, n, wo, dir);
  return bsdfPdf(roughness, metallic, n, wo, dir);
}
DEV float bsdfSample(const State& s, f3 n, f3 wo, uint32_t& seed, f3& bsdf, f3& dir) {
  const float r0 = rnd(seed), r1 = rnd(seed), r2 = rnd(seed);   // GLSL evaluates the vec3(rand, rand, rand) constructor left to right
  return bsdfSampleR(s, n, wo, r0, r1, r2, bsdf, dir);
}

// ---- reservoir.glsl ------------------------------------------------------------------------------
DEV bool resvInvalidW(float w) { return (w != w) || w < 0.0f; }   // :28-34
struct DResv { f3 Li, wi; float dist; uint32_t num; float weight; };
DEV void resvUpdate(DResv& r, f3 Li, f3 wi, float dist, float newWeight, float rv) {   // :47-53
  r.weight = __fadd_rn(r.weight, newWeight);
  r.num += 1;
  if (__fmul_rn(rv, r.weight) < newWeight) { r.Li = Li; r.wi = wi; r.dist = dist; }
}

// ---- texturesMap[] taps (layouts.glsl:51): textureLod(sampler2D, uv, 0) on RGBA8 UNORM ------------------------------------
DEV int wrapCoord(int i, int n, int mode) {              // 0 REPEAT, 1 MIRRORED_REPEAT, 2 CLAMP_TO_EDGE
  if (mode == 2) return i < 0 ? 0 : (i >= n ? n - 1 : i);
  if (mode == 1) { const int p = 2 * n; int m = i % p; if (m < 0) m += p; return m < n ? m : p - 1 - m; }
  int m = i % n; return m < 0 ? m + n : m;
}
DEV float4 texel8(const TextureDev& T, int x, int y) {
  const uint32_t t = __ldg(T.texels + (size_t)y * T.width + x);
  return make_float4(unormToFloat(t & 0xffu), unormToFloat((t >> 8) & 0xffu), unormToFloat((t >> 16) & 0xffu), unormToFloat(t >> 24));
}
DEV float4 mix4(float4 a, float4 b, float t) { return make_float4(mixf(a.x, b.x, t), mixf(a.y, b.y, t), mixf(a.z, b.z, t), mixf(a.w, b.w, t)); }
DEV float4 textureLod0(const DeviceSceneView& sc, int index, float u, float v) {
  const TextureDev T = sc.textures[index];
  if (!T.linear) {
    const int x = wrapCoord(f2i_sat(eid_floorf(__fmul_rn(u, (float)T.width))), T.width, T.wrapS);
    const int y = wrapCoord(f2i_sat(eid_floorf(__fmul_rn(v, (float)T.height))), T.height, T.wrapT);
    return texel8(T, x, y);
  }
  const float x = __fsub_rn(__fmul_rn(u, (float)T.width), 0.5f), y = __fsub_rn(__fmul_rn(v, (float)T.height), 0.5f);
  const float x0f = eid_floorf(x), y0f = eid_floorf(y);
  const float fx = __fsub_rn(x, x0f), fy = __fsub_rn(y, y0f);
  const int x0 = f2i_sat(x0f), y0 = f2i_sat(y0f);
  const int xa = wrapCoord(x0, T.width, T.wrapS), xb = wrapCoord(x0 + 1, T.width, T.wrapS);
  const int ya = wrapCoord(y0, T.height, T.wrapT), yb = wrapCoord(y0 + 1, T.height, T.wrapT);
  return mix4(mix4(texel8(T, xa, ya), texel8(T, xb, ya), fx), mix4(texel8(T, xa, yb), texel8(T, xb, yb), fx), fy);
}
DEV f3 srgbToLinear3(float4 c) { return mk3(eid_powf(c.x, 2.2f), eid_powf(c.y, 2.2f), eid_powf(c.z, 2.2f)); }   // gltf_material.glsl:36-46
DEV void createCoordinateSystem(f3 N, f3& Nt, f3& Nb) {   // common.glsl:81-93
  const f3 a = (fabsf(N.z) > 0.99999f) ? mk3(__fmul_rn(-N.x, N.y), __fsub_rn(1.0f, __fmul_rn(N.y, N.y)), __fmul_rn(-N.y, N.z))
                                        : mk3(__fmul_rn(-N.x, N.z), __fmul_rn(-N.y, N.z), __fsub_rn(1.0f, __fmul_rn(N.z, N.z)));
  Nt = norm3(a);
  Nb = cross3(Nt, N);
}

// ---- scene access --------------------------------------------------------------------------------
struct Payload {       // PtPayload (globals.glsl:48-58) minus the matrices, which are fetched by instanceID
  float hitT, baryU, baryV;
  int primitiveID, instanceID, instanceCustomIndex;
};

// shade_state.glsl:147-221 GetState (tangent frame is only needed by normal mapping: later scope row)
template <bool TEX>
DEV State getState(const DeviceSceneView& sc, const Payload& h, f3 rayDir) {
  State st;
  const InstanceXform& X = sc.instances[h.instanceID];
  const InstanceData gi = sc.geoInfo[h.instanceCustomIndex];
  const uint32_t* idx = (const uint32_t*)(uintptr_t)gi.indexAddress + 3 * (size_t)h.primitiveID;
  const uint32_t i0 = __ldg(idx), i1 = __ldg(idx + 1), i2 = __ldg(idx + 2);
  const float4* vb = (const float4*)(uintptr_t)gi.vertexAddress;   // 32 B / vertex = 2 x float4
  const float4 a0 = __ldg(vb + 2 * (size_t)i0), a1 = __ldg(vb + 2 * (size_t)i0 + 1);
  const float4 b0 = __ldg(vb + 2 * (size_t)i1), b1 = __ldg(vb + 2 * (size_t)i1 + 1);
  const float4 c0 = __ldg(vb + 2 * (size_t)i2), c1 = __ldg(vb + 2 * (size_t)i2 + 1);
  const float bx = __fsub_rn(__fsub_rn(1.0f, h.baryU), h.baryV), by = h.baryU, bz = h.baryV;
  const int mi = gi.materialIndex;
  st.matID = (uint32_t)(mi < 0 ? 0 : mi);

  const f3 pos0 = mk3(a0.x, a0.y, a0.z), pos1 = mk3(b0.x, b0.y, b0.z), pos2 = mk3(c0.x, c0.y, c0.z);
  const f3 position = (pos0 * bx + pos1 * by) + pos2 * bz;
  st.position = xfPoint(X.objectToWorld, position);
  const f3 w0 = xfPoint(X.objectToWorld, pos0), w1 = xfPoint(X.objectToWorld, pos1), w2 = xfPoint(X.objectToWorld, pos2);

  const f3 n0 = octDecode(__float_as_uint(a0.w)), n1 = octDecode(__float_as_uint(b0.w)), n2 = octDecode(__float_as_uint(c0.w));
  const f3 normal = norm3((n0 * bx + n1 * by) + n2 * bz);
  const f3 world_normal = norm3(xfTransposed(normal, X.worldToObject));
  const f3 geom_normal = norm3(cross3(pos1 - pos0, pos2 - pos0));
  const f3 wgeom_normal = norm3(xfTransposed(geom_normal, X.worldToObject));

  // decode_texture (:53-56): clear the handedness bit of v
  const float v0 = __uint_as_float(__float_as_uint(a1.y) & ~1u), v1 = __uint_as_float(__float_as_uint(b1.y) & ~1u), v2 = __uint_as_float(__float_as_uint(c1.y) & ~1u);
  st.u = __fadd_rn(__fadd_rn(__fmul_rn(a1.x, bx), __fmul_rn(b1.x, by)), __fmul_rn(c1.x, bz));
  st.v = __fadd_rn(__fadd_rn(__fmul_rn(v0, bx), __fmul_rn(v1, by)), __fmul_rn(v2, bz));

  // Tangent and binormal (:194-204) feed only the TBN of normal mapping (gltf_material.glsl:135-146): evaluated on demand
  if (TEX && __ldg(&sc.materials[st.matID].normalTexture) > -1) {
    const float h0 = (__float_as_int(a1.y) & 1) == 1 ? 1.0f : -1.0f;
    const f3 t0 = octDecode(__float_as_uint(a1.z)), t1 = octDecode(__float_as_uint(b1.z)), t2 = octDecode(__float_as_uint(c1.z));
    f3 tangent = norm3((t0 * bx + t1 * by) + t2 * bz);
    f3 world_tangent = norm3(xfVector(X.objectToWorld, tangent));
    world_tangent = norm3(world_tangent - world_normal * dot3(world_tangent, world_normal));
    st.tangent = world_tangent;
    st.bitangent = cross3(world_normal, world_tangent) * h0;
  } else { st.tangent = mk3(0.f); st.bitangent = mk3(0.f); }

  st.normal = (dot3(world_normal, wgeom_normal) > 0.0f) ? world_normal : -world_normal;
  st.ffnormal = dot3(st.normal, rayDir) <= 0.0f ? st.normal : -st.normal;
  st.area = __fmul_rn(len3(cross3(w1 - w0, w2 - w0)), 0.5f);
  return st;
}

// gltf_material.glsl:130-176 GetMaterials + GetMetallicRoughness :52-91
template <bool TEX>
DEV void getMaterials(const DeviceSceneView& sc, State& st, f3 rayDir) {
  const float4* m = (const float4*)(sc.materials + st.matID);   // 80 B = 5 x float4
  const float4 q0 = __ldg(m), q1 = __ldg(m + 1), q2 = __ldg(m + 2), q3 = __ldg(m + 3), q4 = __ldg(m + 4);
  const int baseTex = __float_as_int(q1.x), mrTex = __float_as_int(q1.w), emisTex = __float_as_int(q2.x);
  const int nrmTex = __float_as_int(q3.x), transTex = __float_as_int(q3.w);
  if (TEX && nrmTex > -1) {                             // :135-146 normal mapping
    const float4 t = textureLod0(sc, nrmTex, st.u, st.v);
    f3 nv = norm3(mk3(t.x, t.y, t.z) * 2.0f + (-1.0f));
    nv = nv * mk3(q3.y, q3.y, 1.0f);
    const M3 TBN = {st.tangent, st.bitangent, st.normal};
    st.normal = norm3(m3mul(TBN, nv));
    st.ffnormal = dot3(st.normal, rayDir) <= 0.0f ? st.normal : -st.normal;
    createCoordinateSystem(st.ffnormal, st.tangent, st.bitangent);
  }
  st.mat.emission = mk3(q2.y, q2.z, q2.w);
  if (TEX && emisTex > -1) st.mat.emission = st.mat.emission * srgbToLinear3(textureLod0(sc, emisTex, st.u, st.v));
  st.isEmitter = __fadd_rn(__fadd_rn(st.mat.emission.x, st.mat.emission.y), st.mat.emission.z) > 1e-3f;
  float roughness = q1.z, metallic = q1.y;
  if (TEX && mrTex > -1) {
    const float4 t = textureLod0(sc, mrTex, st.u, st.v);
    roughness = __fmul_rn(t.y, roughness);
    metallic = __fmul_rn(t.z, metallic);
  }
  st.mat.albedo = mk3(q0.x, q0.y, q0.z);
  if (TEX && baseTex > -1) st.mat.albedo = st.mat.albedo * srgbToLinear3(textureLod0(sc, baseTex, st.u, st.v));
  st.mat.metallic = metallic;
  st.mat.roughness = gmax(roughness, 0.001f);
  st.mat.transmission = q3.z;
  if (TEX && transTex > -1) st.mat.transmission = __fmul_rn(st.mat.transmission, textureLod0(sc, transTex, st.u, st.v).x);
  st.mat.ior = q4.x;
  st.eta = dot3(st.normal, st.ffnormal) > 0.0f ? __fdiv_rn(1.0f, st.mat.ior) : st.mat.ior;
}

// ---- environment (env_sampling.glsl, common.glsl:69-76, hdr_sampling.cpp sampler state; sun & sky in sunsky.cuh) --
struct EnvView {
  const float4* tex;                 // RGBA32F lat-long map, null = constant environment `constant`
  const ImptSampData* accel;
  int width, height;
  float constant[3];
  SunAndSky sunSky;                  // _sunAndSky uniform (layouts.glsl:53); in_use == 1 replaces the map / the constant
};
DEV void sphericalUv(f3 v, float& u, float& w) {                     // GetSphericalUv (common.glsl:69-76)
  const float gamma = eid_asinf(-v.y);
  const float theta = eid_atan2f(v.z, v.x);
  const float M_1_OVER_PI = 0.318309886183790671538f;
  u = __fadd_rn(__fmul_rn(__fmul_rn(theta, M_1_OVER_PI), 0.5f), 0.5f);
  w = __fadd_rn(__fmul_rn(gamma, M_1_OVER_PI), 0.5f);
}
// texture(environmentTexture, uv).rgb: LINEAR, REPEAT in u, CLAMP_TO_EDGE in v, LOD 0; full-float bilinear weights (contract)
DEV f3 envTexel(const EnvView& E, int x, int y) { const float4 t = __ldg(E.tex + (size_t)y * E.width + x); return mk3(t.x, t.y, t.z); }
DEV f3 envTextureUv(const EnvView& E, float u, float v) {
  const float x = __fsub_rn(__fmul_rn(u, (float)E.width), 0.5f), y = __fsub_rn(__fmul_rn(v, (float)E.height), 0.5f);
  const float x0f = eid_floorf(x), y0f = eid_floorf(y);
  const float fx = __fsub_rn(x, x0f), fy = __fsub_rn(y, y0f);
  const int x0 = f2i_sat(x0f), y0 = f2i_sat(y0f);
  const int xa = wrapCoord(x0, E.width, 0), xb = wrapCoord(x0 + 1, E.width, 0);
  const int ya = wrapCoord(y0, E.height, 2), yb = wrapCoord(y0 + 1, E.height, 2);
  return mix3(mix3(envTexel(E, xa, ya), envTexel(E, xb, ya), fx), mix3(envTexel(E, xa, yb), envTexel(E, xb, yb), fx), fy);
}
DEV f3 envTextureDir(const EnvView& E, f3 dir) {
  if (!E.tex) return mk3(E.constant[0], E.constant[1], E.constant[2]);
  float u, v;
  sphericalUv(dir, u, v);
  return envTextureUv(E, u, v);
}
// Sun & sky is compiled into the FULL kernel variants only (the ones that also carry texture taps and stochastic alpha): the host
// selects them whenever SunAndSky.in_use == 1, and the lean variants keep their register / stack budget.
// EnvRadiance (pathtrace.glsl:40-47)
template <bool FULL>
DEV f3 envRadianceOf(const EnvView& E, const RtxState& rs, f3 dir) {
  if (FULL && E.sunSky.in_use == 1) return sunAndSky(E.sunSky, dir) * rs.hdrMultiplier;
  return envTextureDir(E, dir) * rs.hdrMultiplier;
}
// EnvEval (pathtrace.glsl:60-72): the sun & sky branch returns radiance x hdrMultiplier and pdf 0.5, the HDR branch the bare texel
template <bool FULL>
DEV f3 envEvalOf(const EnvView& E, const RtxState& rs, f3 dir, float& pdf) {
  if (FULL && E.sunSky.in_use == 1) {
    pdf = __fmul_rn(0.5f, rs.environmentProb);
    return sunAndSky(E.sunSky, dir) * rs.hdrMultiplier;
  }
  const f3 radiance = envTextureDir(E, dir);
  pdf = __fmul_rn(__fmul_rn(lum3(radiance), rs.envMapLuminIntegInv), rs.environmentProb);
  return radiance;
}
// EnvSample (env_sampling.glsl:100-135) -> Environment_sample (:38-94): three draws; returns the texel pdf, fills direction + radiance
template <bool FULL>
DEV float envSample(const EnvView& E, float hdrMultiplier, uint32_t& seed, f3& radiance, f3& toLight) {
  if (FULL && E.sunSky.in_use == 1) {                              // env_sampling.glsl:111-125: a direction inside the sun's glow disc, two draws
    const float sunRadius = __fmul_rn(__fmul_rn(0.00465f, 10.0f), E.sunSky.sun_disk_scale);
    const f3 sd = ld3(E.sunSky.sun_direction);
    f3 T, B;
    createCoordinateSystem(sd, T, B);
    const float dx = __fmul_rn(rnd(seed), sunRadius), dy = __fmul_rn(rnd(seed), sunRadius);
    const float dz = __fsqrt_rn(gmax(0.0f, __fsub_rn(__fsub_rn(1.0f, __fmul_rn(dx, dx)), __fmul_rn(dy, dy))));
    toLight = norm3((T * dx + B * dy) + sd * dz);
    radiance = sunAndSky(E.sunSky, toLight) * hdrMultiplier;
    return 0.5f;
  }
  float xx = rnd(seed), xy = rnd(seed), xz = rnd(seed);
  const uint32_t width = (uint32_t)E.width, height = (uint32_t)E.height, size = width * height;
  const uint32_t idx = (uint32_t)min((int)f2u_sat(__fmul_rn(xx, (float)size)), (int)size - 1);
  const ImptSampData sd = E.accel[idx];
  uint32_t envIdx; float pdf;
  if (xy < sd.q) { envIdx = idx; xy = __fdiv_rn(xy, sd.q); pdf = sd.pdf; }
  else { envIdx = (uint32_t)sd.alias; xy = __fdiv_rn(__fsub_rn(xy, sd.q), __fsub_rn(1.0f, sd.q)); pdf = sd.aliasPdf; }
  const uint32_t px = envIdx % width, py = envIdx / width;
  const float u = __fdiv_rn(__fadd_rn((float)px, xy), (float)width);
  const float phi = __fsub_rn(__fmul_rn(u, __fmul_rn(2.0f, EID_PI)), EID_PI);
  float sinPhi, cosPhi;
  eid_sincosf(phi, &sinPhi, &cosPhi);
  const float stepTheta = __fdiv_rn(EID_PI, (float)height);
  const float theta0 = __fmul_rn((float)py, stepTheta);
  const float cosTheta = __fadd_rn(__fmul_rn(eid_cosf(theta0), __fsub_rn(1.0f, xz)), __fmul_rn(eid_cosf(__fadd_rn(theta0, stepTheta)), xz));
  const float theta = eid_acosf(cosTheta);
  const float sinTheta = eid_sinf(theta);
  const float v = __fmul_rn(theta, 0.318309886183790671538f);
  toLight = mk3(__fmul_rn(cosPhi, sinTheta), cosTheta, __fmul_rn(sinPhi, sinTheta));
  radiance = envTextureUv(E, u, v) * hdrMultiplier;
  return pdf;
}

// ---- pathtrace.glsl ------------------------------------------------------------------------------
DEV bool isPdfInvalid(float p) { return p <= 1e-8f || p != p; }   // :14-16

struct LightSampleD { f3 Li, wi; float dist; };

// SampleTriangleLight :103-139 (+ SampleTriangleUniform :90-97): 4 draws
template <bool TEX>
DEV float sampleTriangleLight(const DeviceSceneView& sc, f3 x, uint32_t& seed, LightSampleD& ls) {
  const uint32_t n = sc.lightBufInfo.trigLightSize;
  if (n == 0) return EID_INVALID_PDF;
  int id = min(f2i_sat(__fmul_rn((float)n, rnd(seed))), (int)n - 1);
  const float q = __ldg(&sc.trigLights[id].impSamp.q);
  if (rnd(seed) > q) id = __ldg(&sc.trigLights[id].impSamp.alias);
  const float4* L = (const float4*)(sc.trigLights + id);   // 96 B = 6 x float4
  const float4 l0 = __ldg(L), l1 = __ldg(L + 1), l2 = __ldg(L + 2), l4 = __ldg(L + 4);
  const uint32_t matIndex = __float_as_uint(l0.x);
  const f3 v0 = mk3(l0.z, l0.w, l1.x), v1 = mk3(l1.y, l1.z, l1.w), v2 = mk3(l2.x, l2.y, l2.z);
  const float lightPdf = l4.w;   // impSamp.pdf
  f3 normal = cross3(v1 - v0, v2 - v0);
  const float area = __fmul_rn(len3(normal), 0.5f);
  normal = norm3(normal);
  const float ru = rnd(seed), rv = rnd(seed);
  const float r = __fsqrt_rn(rv);
  const float bu = __fsub_rn(1.0f, r), bv = __fmul_rn(ru, r);
  const f3 y = (bu * v0 + bv * v1) + __fsub_rn(__fsub_rn(1.0f, bu), bv) * v2;
  const float4 em = __ldg((const float4*)(sc.materials + matIndex) + 2);   // emissiveTexture, emissiveFactor.xyz
  f3 emission = mk3(em.y, em.z, em.w);
  if (TEX && __float_as_int(em.x) > -1) {                // :127-132 textured emitter: uv from the light's own uv0..2
    const float4 l3 = __ldg(L + 3);
    const float w2 = __fsub_rn(__fsub_rn(1.0f, bu), bv);
    const float tu = __fadd_rn(__fadd_rn(__fmul_rn(bu, l2.w), __fmul_rn(bv, l3.y)), __fmul_rn(w2, l3.w));
    const float tv = __fadd_rn(__fadd_rn(__fmul_rn(bu, l3.x), __fmul_rn(bv, l3.z)), __fmul_rn(w2, l4.x));
    emission = emission * srgbToLinear3(textureLod0(sc, __float_as_int(em.x), tu, tv));
  }
  const f3 dir = y - x;
  const float dist = len3(dir);
  ls.Li = emission / area;
  ls.wi = dir / dist;
  ls.dist = dist;
  return __fdiv_rn(__fmul_rn(lightPdf, __fmul_rn(dist, dist)), __fmul_rn(area, fabsf(dot3(ls.wi, normal))));
}
// SamplePuncLight :141-159: 2 draws; type / range / cone are ignored by the reference
DEV float samplePuncLight(const DeviceSceneView& sc, f3 x, uint32_t& seed, LightSampleD& ls) {
  const uint32_t n = sc.lightBufInfo.puncLightSize;
  if (n == 0) return EID_INVALID_PDF;
  int id = min(f2i_sat(__fmul_rn((float)n, rnd(seed))), (int)n - 1);
  if (rnd(seed) > sc.puncLights[id].impSamp.q) id = sc.puncLights[id].impSamp.alias;
  const PuncLight& L = sc.puncLights[id];
  const f3 dir = ld3(L.position) - x;
  const float dist = len3(dir);
  ls.Li = (ld3(L.color) * L.intensity) / __fmul_rn(dist, dist);
  ls.wi = dir / dist;
  ls.dist = dist;
  return L.impSamp.pdf;
}
// SampleDirectLightNoVisibility :161-183
template <bool TEX>
DEV float sampleDirectLightNoVisibility(const DeviceSceneView& sc, const EnvView& env, const RtxState& rs, f3 pos, uint32_t& seed, LightSampleD& ls) {
  const float r = rnd(seed);
  const float envProb = rs.environmentProb;
  if (r < envProb) {                         // sample the environment (:163-172)
    if (!env.tex && !(TEX && env.sunSky.in_use == 1)) return EID_INVALID_PDF;    // (unreachable: the host refuses environmentProb > 0 without an environment)
    const float pdf = envSample<TEX>(env, rs.hdrMultiplier, seed, ls.Li, ls.wi);
    if (isPdfInvalid(pdf)) return EID_INVALID_PDF;
    ls.dist = EID_INFINITY;
    return __fmul_rn(pdf, envProb);
  }
  const float lightProb = __fsub_rn(1.0f, envProb);
  const float tsp = sc.lightBufInfo.trigSampProb;
  if (r < __fadd_rn(envProb, __fmul_rn(lightProb, tsp)))
    return __fmul_rn(__fmul_rn(lightProb, sampleTriangleLight<TEX>(sc, pos, seed, ls)), tsp);
  return __fmul_rn(__fmul_rn(lightProb, samplePuncLight(sc, pos, seed, ls)), __fsub_rn(1.0f, tsp));
}
// clampRadiance :222-232
DEV f3 clampRadiance(f3 radiance, float threshold) {
  if (nan3(radiance)) return mk3(0.0f);
  float lum = lum3(radiance);
  if (lum > threshold) radiance = radiance * __fdiv_rn(threshold, lum);
  return radiance;
}
// raySpawn :260-270 (normalizeDir = true) and the denoiser's variant denoise_common.glsl:27-35 (false)
template <bool NORMALIZE>
DEV void raySpawn(const SceneCamera& cam, int cx, int cy, int sw, int sh, f3& origin, f3& dir) {
  const float ux = __fdiv_rn(__fadd_rn((float)cx, 0.5f), (float)sw), uy = __fdiv_rn(__fadd_rn((float)cy, 0.5f), (float)sh);
  const float dx = __fsub_rn(__fmul_rn(ux, 2.0f), 1.0f), dy = __fsub_rn(__fmul_rn(uy, 2.0f), 1.0f);
  origin = mk3(cam.viewInverse.m[12], cam.viewInverse.m[13], cam.viewInverse.m[14]);
  float tg[4];
  mat4MulV(cam.projInverse, dx, dy, 1.0f, 1.0f, tg);
  f3 d = mat4MulDir(cam.viewInverse, norm3(mk3(tg[0], tg[1], tg[2])));
  dir = NORMALIZE ? norm3(d) : d;
}

}  // namespace eid

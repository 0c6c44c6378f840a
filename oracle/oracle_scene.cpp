/*
 * oracle/oracle_scene.cpp — TEST INFRASTRUCTURE (CPU oracle).
 * Host-side table builders of the reference restated: src/scene.cpp (vertex compression, material,
 * punctual / emissive-triangle light tables, camera update), src/alias_table.hpp, the TLAS instance
 * flags of src/accelstruct.cpp, and a reference intersector (brute force or a binned-SAH BVH2)
 * implementing the ray-query semantics documented in SURVEY.md §8c.
 */
#include "oracle.h"
#include "eid_vecmath.h"
#include <algorithm>
#include <cmath>
#include <cstdio>

namespace orc {

// ---- tools.hpp:57-61 ------------------------------------------------------------------------
static inline float luminanceHost(const float* c) { return c[0] * 0.2126f + c[1] * 0.7152f + c[2] * 0.0722f; }

// ---- compress.glsl:76-98 (C++ twin of roundEven) ---------------------------------------------
static float roundEvenCpp(float x) {
  int Integer = static_cast<int>(x);
  float IntegerPart = static_cast<float>(Integer);
  float FractionalPart = (x - floorf(x));
  if (FractionalPart > 0.5f || FractionalPart < 0.5f) return roundf(x);
  else if ((Integer % 2) == 0) return IntegerPart;
  else if (x <= 0) return IntegerPart - 1;
  else return IntegerPart + 1;
}

// ---- compress.glsl:111-139 -------------------------------------------------------------------
uint compress_unit_vec(vec3 nv) {
  const float C_Stack_Max = 3.402823466e+38f;
  if ((nv.x < C_Stack_Max) && !gisinf(nv.x)) {
    const float d = 32767.0f / ((gabs(nv.x) + gabs(nv.y)) + gabs(nv.z));
    int x = f2i(roundEvenCpp(nv.x * d));
    int y = f2i(roundEvenCpp(nv.y * d));
    if (nv.z < 0.0f) {
      const int maskx = x >> 31;
      const int masky = y >> 31;
      const int tmp = 32767 + maskx + masky;
      const int tmpx = x;
      x = (tmp - (y ^ masky)) ^ maskx;
      y = (tmp - (tmpx ^ maskx)) ^ masky;
    }
    uint packed = (uint(y + 32767) << 16) | uint(x + 32767);
    if (packed == ~0u) return ~0x1u;
    return packed;
  }
  return ~0u;
}

// ---- compress.glsl:143-180 -------------------------------------------------------------------
static float short_to_floatm11(const int v) {
  return (v >= 0) ? (uintBitsToFloat(0x3F800000u | (uint(v) << 8)) - 1.0f)
                  : (uintBitsToFloat((0x80000000u | 0x3F800000u) | (uint(-v) << 8)) + 1.0f);
}
vec3 decompress_unit_vec(uint packed) {
  if (packed != ~0u) {
    int x = int(packed & 0xFFFFu) - 32767;
    int y = int(packed >> 16) - 32767;
    const int maskx = x >> 31;
    const int masky = y >> 31;
    const int tmp0 = 32767 + maskx + masky;
    const int ymask = y ^ masky;
    const int tmp1 = tmp0 - (x ^ maskx);
    const int z = tmp1 - ymask;
    float zf;
    if (z < 0) {
      x = (tmp0 - ymask) ^ maskx;
      y = tmp1 ^ masky;
      zf = uintBitsToFloat((0x80000000u | 0x3F800000u) | (uint(-z) << 8)) + 1.0f;
    } else {
      zf = uintBitsToFloat(0x3F800000u | (uint(z) << 8)) - 1.0f;
    }
    return normalize(vec3(short_to_floatm11(x), short_to_floatm11(y), zf));
  }
  return vec3(3.402823466e+38f);
}

// ---- alias_table.hpp:21-63 -------------------------------------------------------------------
void discreteSampler1D(std::vector<float> values, std::vector<float>& prob, std::vector<int>& failId) {
  struct D { float prob; int failId; };
  float sumAll = 0.f;
  for (const auto& val : values) sumAll += val;
  float sumInv = static_cast<float>(values.size()) / sumAll;
  for (auto& val : values) val *= sumInv;
  std::vector<D> binom(values.size());
  std::vector<D> stackGtOne(values.size() * 2), stackLsOne(values.size() * 2);
  int topGtOne = 0, topLsOne = 0;
  for (int i = 0; i < (int)values.size(); i++) {
    float val = values[i];
    (val > 1.0f ? stackGtOne[topGtOne++] : stackLsOne[topLsOne++]) = D{val, i};
  }
  while (topGtOne && topLsOne) {
    D gt = stackGtOne[--topGtOne];
    D ls = stackLsOne[--topLsOne];
    binom[ls.failId] = D{ls.prob, gt.failId};
    gt.prob -= (1.0f - ls.prob);
    (gt.prob > 1.0f ? stackGtOne[topGtOne++] : stackLsOne[topLsOne++]) = gt;
  }
  for (int i = topGtOne - 1; i >= 0; i--) { D gt = stackGtOne[i]; binom[gt.failId] = gt; }
  for (int i = topLsOne - 1; i >= 0; i--) { D ls = stackLsOne[i]; binom[ls.failId] = ls; }
  prob.resize(values.size());
  failId.resize(values.size());
  for (size_t i = 0; i < values.size(); ++i) { prob[i] = binom[i].prob; failId[i] = binom[i].failId; }
}

static mat4x3 toMat4x3(const float* m) {   // upper 3 rows of a column-major 4x4
  mat4x3 r;
  for (int c = 0; c < 4; ++c) r.c[c] = vec3(m[c * 4 + 0], m[c * 4 + 1], m[c * 4 + 2]);
  return r;
}

// textureLod(sampler2D, uv, 0) on an 8-bit UNORM image: LOD 0 selects the magnification filter; unnormalised coordinates
// u*w - 0.5, floor / fract, per-axis wrap, full-float bilinear weights (contract)
static int wrapCoord(int i, int n, int mode) {
  if (mode == 2) return i < 0 ? 0 : (i >= n ? n - 1 : i);
  if (mode == 1) { int p = 2 * n; int m = i % p; if (m < 0) m += p; return m < n ? m : p - 1 - m; }
  int m = i % n; return m < 0 ? m + n : m;
}
vec4 Texture::sample(vec2 uv) const {
  auto texel = [&](int x, int y) {
    const uint8_t* p = &rgba[4 * ((size_t)y * width + x)];
    return vec4(float(p[0]) / 255.0f, float(p[1]) / 255.0f, float(p[2]) / 255.0f, float(p[3]) / 255.0f);
  };
  if (!linear) {
    int x = wrapCoord(f2i(eid_floorf(uv.x * float(width))), width, wrapS);
    int y = wrapCoord(f2i(eid_floorf(uv.y * float(height))), height, wrapT);
    return texel(x, y);
  }
  const float x = uv.x * float(width) - 0.5f, y = uv.y * float(height) - 0.5f;
  const float x0f = eid_floorf(x), y0f = eid_floorf(y);
  const float fx = x - x0f, fy = y - y0f;
  const int x0 = f2i(x0f), y0 = f2i(y0f);
  const int xa = wrapCoord(x0, width, wrapS), xb = wrapCoord(x0 + 1, width, wrapS);
  const int ya = wrapCoord(y0, height, wrapT), yb = wrapCoord(y0 + 1, height, wrapT);
  auto mix4 = [](vec4 a, vec4 b, float t) { return vec4(mix(a.x, b.x, t), mix(a.y, b.y, t), mix(a.z, b.z, t), mix(a.w, b.w, t)); };
  return mix4(mix4(texel(xa, ya), texel(xb, ya), fx), mix4(texel(xa, yb), texel(xb, yb), fx), fy);
}

// Scene::createTextureImages + gltfSamplerToVulkan (scene.cpp:513-646)
static void buildTextures(const eid_scene_desc& d, std::vector<Texture>& out) {
  out.clear();
  if (d.imageCount == 0) { out.push_back(Texture{}); return; }   // "No images, add a default one" (:570-576)
  auto filt = [](int f) { return (f == 9729 || f == 9985 || f == 9987) ? 1 : 0; };   // std::map default = NEAREST for unknown keys
  auto wrap = [](int w) { return w == 33071 ? 2 : (w == 33648 ? 1 : 0); };
  for (uint32_t i = 0; i < d.textureCount; ++i) {
    const eid_texture_desc& t = d.textures[i];
    Texture tex;
    if (t.image < 0 || (uint32_t)t.image >= d.imageCount) { out.push_back(tex); continue; }   // default white texture
    const eid_image_desc& im = d.images[t.image];
    if (im.rgba8 && im.width && im.height) {
      tex.width = (int)im.width; tex.height = (int)im.height;
      tex.rgba.assign(im.rgba8, im.rgba8 + 4 * (size_t)im.width * im.height);
    }
    if (t.hasSampler) { tex.linear = filt(t.magFilter); tex.wrapS = wrap(t.wrapS); tex.wrapT = wrap(t.wrapT); }
    out.push_back(tex);
  }
}

void Scene::load(const eid_scene_desc& d) {
  buildTextures(d, textures);
  positions.assign(d.positions, d.positions + 3 * (size_t)d.vertexCount);
  normals.assign(d.normals, d.normals + 3 * (size_t)d.vertexCount);
  tangents.assign(d.tangents, d.tangents + 4 * (size_t)d.vertexCount);
  texcoords0.assign(d.texcoords0, d.texcoords0 + 2 * (size_t)d.vertexCount);
  colors0.assign(d.colors0, d.colors0 + 4 * (size_t)d.vertexCount);
  indices.assign(d.indices, d.indices + d.indexCount);
  primMeshes.assign(d.primMeshes, d.primMeshes + d.primMeshCount);
  nodes.assign(d.nodes, d.nodes + d.nodeCount);
  materials.assign(d.materials, d.materials + d.materialCount);
  lights.assign(d.lights, d.lights + d.lightCount);

  // setCameraFromScene (scene.cpp:295-314)
  if (d.hasCamera) {
    for (int i = 0; i < 3; ++i) { eye[i] = d.camEye[i]; center[i] = d.camCenter[i]; up[i] = d.camUp[i]; }
    fovDeg = d.camYfovRad * 57.29577951308232f;   // rad2deg
  }
  camera = SceneCamera{};
  camera.nbLights = (int)lights.size();   // scene.cpp:78
  staticEye[0] = staticEye[1] = staticEye[2] = 0.f;

  // createMaterialBuffer (scene.cpp:415-448)
  shadeMaterials.clear();
  for (auto& m : materials) {
    GltfShadeMaterial s{};
    s.pbrBaseColorFactor = {m.baseColorFactor[0], m.baseColorFactor[1], m.baseColorFactor[2], m.baseColorFactor[3]};
    s.pbrBaseColorTexture = m.baseColorTexture;
    s.pbrMetallicFactor = m.metallicFactor;
    s.pbrRoughnessFactor = m.roughnessFactor;
    s.pbrMetallicRoughnessTexture = m.metallicRoughnessTexture;
    s.emissiveTexture = m.emissiveTexture;
    s.emissiveFactor = {m.emissiveFactor[0], m.emissiveFactor[1], m.emissiveFactor[2]};
    s.normalTexture = m.normalTexture;
    s.normalTextureScale = m.normalTextureScale;
    s.transmissionFactor = m.transmissionFactor;
    s.transmissionTexture = m.transmissionTexture;
    s.ior = gmin(gmax(m.ior, 1.f), MAX_IOR_MINUS_ONE + 1.f);   // nv_clamp
    s.alphaMode = m.alphaMode;
    s.alphaCutoff = m.alphaCutoff;
    shadeMaterials.push_back(s);
  }

  // createPuncLightBuffer + createPuncLightImptSampAccel (scene.cpp:319-353, 700-728)
  puncLights.clear();
  puncLightWeight = 0.f;
  trigLightWeight = 0.f;
  for (auto& l : lights) {
    PuncLight p{};
    const float* w = l.worldMatrix;
    p.position = {w[12], w[13], w[14]};
    p.direction = {-w[8], -w[9], -w[10]};
    p.color = {l.color[0], l.color[1], l.color[2]};
    p.innerConeCos = static_cast<float>(cos((double)l.innerConeAngle));
    p.outerConeCos = static_cast<float>(cos((double)l.outerConeAngle));
    p.range = l.range;
    p.intensity = l.intensity;
    p.type = l.type;
    puncLights.push_back(p);
  }
  lightBufInfo = LightBufInfo{};
  lightBufInfo.puncLightSize = (uint32_t)puncLights.size();
  if (!puncLights.empty()) {
    float total = 0.f;
    std::vector<float> distrib;
    for (auto& p : puncLights) {
      float power = luminanceHost(&p.color.x) * p.intensity * 3.1416f * 4.f;
      distrib.push_back(power);
      total += power;
    }
    std::vector<float> prob; std::vector<int> fail;
    discreteSampler1D(distrib, prob, fail);
    for (size_t i = 0; i < distrib.size(); ++i) {
      auto& a = puncLights[i].impSamp;
      a.alias = fail[i]; a.q = prob[i]; a.pdf = distrib[i] / total; a.aliasPdf = distrib[fail[i]] / total;
    }
    puncLightWeight = total;
  }
  if (puncLights.empty()) puncLights.push_back(PuncLight{});

  // createVertexBuffer (scene.cpp:209-289)
  vertexBufs.clear(); indexBufs.clear(); instMaterial.clear();
  for (auto& pm : primMeshes) {
    std::vector<VertexAttributes> verts;
    verts.reserve(pm.vertexCount);
    for (size_t v = 0; v < pm.vertexCount; ++v) {
      size_t idx = pm.vertexOffset + v;
      VertexAttributes a{};
      a.position = {positions[3 * idx], positions[3 * idx + 1], positions[3 * idx + 2]};
      a.normal = compress_unit_vec(vec3(normals[3 * idx], normals[3 * idx + 1], normals[3 * idx + 2]));
      a.tangent = compress_unit_vec(vec3(tangents[4 * idx], tangents[4 * idx + 1], tangents[4 * idx + 2]));
      a.texcoord = {texcoords0[2 * idx], texcoords0[2 * idx + 1]};
      a.color = packUnorm4x8(vec4(colors0[4 * idx], colors0[4 * idx + 1], colors0[4 * idx + 2], colors0[4 * idx + 3]));
      uint32_t value = floatBitsToUint(a.texcoord.y);
      if (tangents[4 * idx + 3] > 0) value |= 1; else value &= ~1u;
      a.texcoord.y = uintBitsToFloat(value);
      verts.push_back(a);
    }
    vertexBufs.push_back(std::move(verts));
    indexBufs.emplace_back(indices.begin() + pm.firstIndex, indices.begin() + pm.firstIndex + pm.indexCount);
    instMaterial.push_back(pm.materialIndex);   // createInstanceDataBuffer (scene.cpp:179-195)
  }

  // createTrigLightBuffer + createTrigLightImptSampAccel (scene.cpp:355-409, 742-772)
  trigLights.clear();
  for (auto& node : nodes) {
    const auto& pm = primMeshes[node.primMesh];
    const auto& mat = materials[pm.materialIndex];
    if (luminanceHost(mat.emissiveFactor) > 1e-2f) {
      for (uint32_t idx = pm.firstIndex; idx < pm.firstIndex + pm.indexCount - 1; idx += 3) {
        TrigLight t{};
        uint32_t i0 = indices[idx] + pm.vertexOffset, i1 = indices[idx + 1] + pm.vertexOffset, i2 = indices[idx + 2] + pm.vertexOffset;
        t.transformIndex = 0xFFFFFFFFu;   // transforms.size()-1 with an always-empty vector (scene.cpp:380)
        t.matIndex = pm.materialIndex;
        mat4x3 w = toMat4x3(node.worldMatrix);
        vec3 p0 = mulPoint(w, vec3(positions[3 * i0], positions[3 * i0 + 1], positions[3 * i0 + 2]));
        vec3 p1 = mulPoint(w, vec3(positions[3 * i1], positions[3 * i1 + 1], positions[3 * i1 + 2]));
        vec3 p2 = mulPoint(w, vec3(positions[3 * i2], positions[3 * i2 + 1], positions[3 * i2 + 2]));
        t.v0 = {p0.x, p0.y, p0.z}; t.v1 = {p1.x, p1.y, p1.z}; t.v2 = {p2.x, p2.y, p2.z};
        t.uv0 = {texcoords0[2 * i0], texcoords0[2 * i0 + 1]};
        t.uv1 = {texcoords0[2 * i1], texcoords0[2 * i1 + 1]};
        t.uv2 = {texcoords0[2 * i2], texcoords0[2 * i2 + 1]};
        trigLights.push_back(t);
      }
    }
  }
  {
    float total = 0.f;
    std::vector<float> distrib;
    for (auto& t : trigLights) {
      float power = luminanceHost(materials[t.matIndex].emissiveFactor);
      distrib.push_back(power);
      total += power;
    }
    if (!trigLights.empty()) {
      std::vector<float> prob; std::vector<int> fail;
      discreteSampler1D(distrib, prob, fail);
      for (size_t i = 0; i < trigLights.size(); ++i) {
        auto& a = trigLights[i].impSamp;
        a.alias = fail[i]; a.q = prob[i]; a.pdf = distrib[i] / total; a.aliasPdf = distrib[fail[i]] / total;
      }
    }
    trigLightWeight = total;
  }
  lightBufInfo.trigLightSize = (uint32_t)trigLights.size();
  if (trigLights.empty()) trigLights.push_back(TrigLight{});
  // scene.cpp:101-103
  if (lightBufInfo.puncLightSize > 0 || lightBufInfo.trigLightSize > 0)
    lightBufInfo.trigSampProb = trigLightWeight / (trigLightWeight + puncLightWeight);

  buildAccel();
}

// Scene::updateCamera (scene.cpp:777-826); CameraManip.getMatrix() = right-handed look-at
void Scene::updateCamera(uint32_t w, uint32_t h) {
  const float aspectRatio = w / (float)h;
  float jx = .5f / w, jy = .5f / h;
  eid_mat4 view = eid_look_at({eye[0], eye[1], eye[2]}, {center[0], center[1], center[2]}, {up[0], up[1], up[2]});
  eid_mat4 proj = eid_perspectiveVK(fovDeg, aspectRatio, CAMERA_NEAR, CAMERA_FAR);
  proj.m[2 * 4 + 0] += jx;   // a02
  proj.m[2 * 4 + 1] += jy;   // a12
  camera.lastProjView = camera.projView;
  camera.lastView = eid_mat4_invert(camera.viewInverse);
  camera.viewInverse = eid_mat4_invert(view);
  camera.projInverse = eid_mat4_invert(proj);
  camera.projView = eid_mat4_mul(proj, view);
  camera.lastPosition = {staticEye[0], staticEye[1], staticEye[2]};
  for (int i = 0; i < 3; ++i) staticEye[i] = eye[i];
}

// accelstruct.cpp:132-162 — one instance per node, world-space triangle soup
void Scene::buildAccel() {
  objectToWorld.clear(); worldToObject.clear(); tris.clear(); hasNonOpaque = false;
  for (int k = 0; k < 3; ++k) { bboxMin[k] = 1e30f; bboxMax[k] = -1e30f; }
  for (size_t n = 0; n < nodes.size(); ++n) {
    const auto& node = nodes[n];
    eid_mat4 wm; memcpy(wm.m, node.worldMatrix, 64);
    eid_mat4 inv = eid_mat4_invert(wm);
    mat4x3 o2w = toMat4x3(wm.m), w2o = toMat4x3(inv.m);
    objectToWorld.push_back(o2w); worldToObject.push_back(w2o);
    const auto& pm = primMeshes[node.primMesh];
    const auto& mat = materials[pm.materialIndex];
    const auto& vb = vertexBufs[node.primMesh];
    const auto& ib = indexBufs[node.primMesh];
    // facing is decided in object space (Vulkan); a mirroring instance transform flips the world-space winding
    vec3 c0 = o2w.c[0], c1 = o2w.c[1], c2 = o2w.c[2];
    const int flip = (dot(c0, cross(c1, c2)) < 0.0f) ? 1 : 0;
    for (uint32_t t = 0; t + 3 <= pm.indexCount; t += 3) {
      vec3 p[3];
      for (int k = 0; k < 3; ++k) {
        const auto& v = vb[ib[t + k]];
        p[k] = mulPoint(o2w, vec3(v.position.x, v.position.y, v.position.z));
        for (int a = 0; a < 3; ++a) { bboxMin[a] = std::min(bboxMin[a], p[k][a]); bboxMax[a] = std::max(bboxMax[a], p[k][a]); }
      }
      OTri tr;
      tr.v0 = p[0]; tr.e1 = p[1] - p[0]; tr.e2 = p[2] - p[0];
      tr.prim = (int)(t / 3); tr.inst = (int)n; tr.customIndex = node.primMesh;
      tr.cullDisable = (mat.doubleSided == 1) ? 1 : 0;
      tr.opaque = (mat.alphaMode == 0 || (mat.baseColorFactor[3] == 1.0f && mat.baseColorTexture == -1)) ? 1 : 0;   // accelstruct.cpp:145-147
      if (!tr.opaque) hasNonOpaque = true;
      tr.flip = flip;
      tris.push_back(tr);
    }
  }
  float diag = 0.f;
  for (int a = 0; a < 3; ++a) { float e = bboxMax[a] - bboxMin[a]; diag += e * e; }
  diag = sqrtf(diag);
  if (useBvh && tris.size() > 64) bvh.build(tris, 1e-5f * diag + 1e-7f);
  else { bvh.nodes.clear(); bvh.order.clear(); }
}

// ---- ray / triangle test: the exact arithmetic both sides of the parity contract use ---------
// (replaces the driver's ray-query traversal, traceray_rq.glsl:108-185; semantics SURVEY.md §8c:
//  tmin=0 exclusive, tmax exclusive, back faces culled unless the instance disables culling,
//  barycentrics (u,v) weight vertices 1 and 2, ties broken by lowest (instance, primitive).)
static inline bool triTest(const OTri& T, vec3 o, vec3 d, float tmax, float& t, float& u, float& v) {
  vec3 pvec = cross(d, T.e2);
  float det = dot(T.e1, pvec);
  if (T.cullDisable ? (det == 0.0f) : !((T.flip ? -det : det) > 0.0f)) return false;
  float inv = 1.0f / det;
  vec3 tvec = o - T.v0;
  u = dot(tvec, pvec) * inv;
  if (!(u >= 0.0f && u <= 1.0f)) return false;
  vec3 qvec = cross(tvec, T.e1);
  v = dot(d, qvec) * inv;
  if (!(v >= 0.0f && u + v <= 1.0f)) return false;
  t = dot(T.e2, qvec) * inv;
  return t > 0.0f && t < tmax;
}

static inline bool better(float t, int inst, int prim, const Hit& h) {
  if (t < h.hitT) return true;
  if (t > h.hitT) return false;
  if (inst != h.instanceID) return inst < h.instanceID;
  return prim < h.primitiveID;
}

static inline bool boxTest(const BvhNode& n, vec3 o, vec3 id, float tmax) {
  float tn = 0.f, tf = tmax;
  for (int a = 0; a < 3; ++a) {
    float t0 = (n.lo[a] - o[a]) * id[a], t1 = (n.hi[a] - o[a]) * id[a];
    if (t0 != t0 || t1 != t1) continue;   // 0*inf: origin on the slab plane, direction parallel -> inside
    if (t0 > t1) std::swap(t0, t1);
    tn = std::max(tn, t0); tf = std::min(tf, t1);
  }
  return tn <= tf * 1.0000004f + 1e-30f;
}

Hit Scene::closestHit(vec3 o, vec3 d, float tmax, std::atomic<uint64_t>* ctr, const HitKey* after) const {
  if (ctr) ctr->fetch_add(1, std::memory_order_relaxed);
  Hit h; h.hitT = tmax; h.primitiveID = h.instanceID = h.instanceCustomIndex = -1; h.bary = vec2(0, 0); h.opaque = 1;
  bool found = false;
  auto consider = [&](const OTri& T) {
    float t, u, v;
    // candidates at exactly the current best t must still be examined for the tie-break
    if (triTest(T, o, d, tmax, t, u, v)) {
      if (after) {   // only candidates strictly after the key (t, inst, prim)
        if (t < after->t) return;
        if (t == after->t && (T.inst < after->inst || (T.inst == after->inst && T.prim <= after->prim))) return;
      }
      if (!found || better(t, T.inst, T.prim, h)) {
        h.hitT = t; h.primitiveID = T.prim; h.instanceID = T.inst; h.instanceCustomIndex = T.customIndex; h.bary = vec2(u, v);
        h.opaque = T.opaque;
        found = true;
      }
    }
  };
  if (bvh.nodes.empty()) {
    for (const auto& T : tris) consider(T);
  } else {
    vec3 id(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    int stack[128]; int sp = 0; stack[sp++] = 0;
    while (sp) {
      const BvhNode& n = bvh.nodes[stack[--sp]];
      if (!boxTest(n, o, id, found ? h.hitT : tmax)) continue;
      if (n.count) { for (int i = 0; i < n.count; ++i) consider(tris[bvh.order[n.first + i]]); }
      else { stack[sp++] = n.left; stack[sp++] = n.right; }
    }
  }
  if (!found) h.hitT = 1e28f;
  return h;
}

bool Scene::anyHit(vec3 o, vec3 d, float tmax, std::atomic<uint64_t>* ctr) const {
  if (ctr) ctr->fetch_add(1, std::memory_order_relaxed);
  float t, u, v;
  if (bvh.nodes.empty()) {
    for (const auto& T : tris) if (triTest(T, o, d, tmax, t, u, v)) return true;
    return false;
  }
  vec3 id(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
  int stack[128]; int sp = 0; stack[sp++] = 0;
  while (sp) {
    const BvhNode& n = bvh.nodes[stack[--sp]];
    if (!boxTest(n, o, id, tmax)) continue;
    if (n.count) { for (int i = 0; i < n.count; ++i) if (triTest(tris[bvh.order[n.first + i]], o, d, tmax, t, u, v)) return true; }
    else { stack[sp++] = n.left; stack[sp++] = n.right; }
  }
  return false;
}

// ---- hdr_sampling.cpp:107-176 buildAliasmap --------------------------------------------------------------
static float buildAliasmap(const std::vector<float>& data, std::vector<ImptSampData>& accel) {
  auto size = static_cast<uint32_t>(data.size());
  float sum = 0.f;
  for (float d : data) sum += d;                      // std::accumulate(..., 0.f)
  auto fSize = static_cast<float>(size);
  float inverseAverage = fSize / sum;
  for (uint32_t i = 0; i < size; ++i) { accel[i].q = data[i] * inverseAverage; accel[i].alias = (int)i; }
  std::vector<uint32_t> partitionTable(size);
  uint32_t s = 0u, large = size;
  for (uint32_t i = 0; i < size; ++i) {
    if (accel[i].q < 1.f) partitionTable[s++] = i;
    else partitionTable[--large] = i;
  }
  for (s = 0; s < large && large < size; ++s) {
    const uint32_t smallEnergyIndex = partitionTable[s];
    const uint32_t highEnergyIndex = partitionTable[large];
    accel[smallEnergyIndex].alias = (int)highEnergyIndex;
    const float differenceWithAverage = 1.f - accel[smallEnergyIndex].q;
    accel[highEnergyIndex].q -= differenceWithAverage;
    if (accel[highEnergyIndex].q < 1.0f) large++;
  }
  return sum;
}

// ---- hdr_sampling.cpp:181-242 createEnvironmentAccel ----------------------------------------------------
void Environment::create(const float* rgba, uint32_t w, uint32_t h) {
  width = w; height = h;
  pixels.assign(rgba, rgba + 4 * (size_t)w * h);
  const uint32_t rx = w, ry = h;
  accel.assign((size_t)rx * ry, ImptSampData{});
  std::vector<float> importanceData((size_t)rx * ry);
  float cosTheta0 = 1.0f;
  const float stepPhi = float(2.0 * M_PI) / float(rx);
  const float stepTheta = float(M_PI) / float(ry);
  double total = 0;
  for (uint32_t y = 0; y < ry; ++y) {
    const float theta1 = float(y + 1) * stepTheta;
    const float cosTheta1 = std::cos(theta1);
    const float area = (cosTheta0 - cosTheta1) * stepPhi;
    cosTheta0 = cosTheta1;
    for (uint32_t x = 0; x < rx; ++x) {
      const uint32_t idx = y * rx + x, idx4 = idx * 4;
      float cieLuminance = luminanceHost(&pixels[idx4]);
      importanceData[idx] = area * std::max(pixels[idx4], std::max(pixels[idx4 + 1], pixels[idx4 + 2]));
      total += cieLuminance;
    }
  }
  average = static_cast<float>(total) / static_cast<float>(rx * ry);
  integral = buildAliasmap(importanceData, accel);
  const float invEnvIntegral = 1.0f / integral;
  for (uint32_t i = 0; i < rx * ry; ++i) {
    const uint32_t idx4 = i * 4;
    accel[i].pdf = std::max(pixels[idx4], std::max(pixels[idx4 + 1], pixels[idx4 + 2])) * invEnvIntegral;
  }
  for (uint32_t i = 0; i < rx * ry; ++i) accel[i].aliasPdf = accel[accel[i].alias].pdf;
}

// texture(environmentTexture, uv): bilinear, REPEAT in u / CLAMP_TO_EDGE in v, full-float weights (contract)
vec3 Environment::texture(vec2 uv) const {
  const float x = uv.x * float(width) - 0.5f, y = uv.y * float(height) - 0.5f;
  const float x0f = eid_floorf(x), y0f = eid_floorf(y);
  const float fx = x - x0f, fy = y - y0f;
  int x0 = f2i(x0f), y0 = f2i(y0f);
  auto wrap = [&](int v) { int m = v % (int)width; return m < 0 ? m + (int)width : m; };
  auto clampv = [&](int v) { return v < 0 ? 0 : (v >= (int)height ? (int)height - 1 : v); };
  const int xa = wrap(x0), xb = wrap(x0 + 1), ya = clampv(y0), yb = clampv(y0 + 1);
  auto tex = [&](int xx, int yy) { const float* p = &pixels[4 * ((size_t)yy * width + xx)]; return vec3(p[0], p[1], p[2]); };
  return mix(mix(tex(xa, ya), tex(xb, ya), fx), mix(tex(xa, yb), tex(xb, yb), fx), fy);
}

// ---- binned-SAH BVH2 (oracle-only acceleration; results do not depend on it) ------------------
void Bvh::build(const std::vector<OTri>& tris, float pad) {
  const int N = (int)tris.size();
  std::vector<float> lo(3 * (size_t)N), hi(3 * (size_t)N), cen(3 * (size_t)N);
  for (int i = 0; i < N; ++i) {
    const OTri& T = tris[i];
    vec3 a = T.v0, b = T.v0 + T.e1, c = T.v0 + T.e2;
    for (int k = 0; k < 3; ++k) {
      float mn = std::min(a[k], std::min(b[k], c[k])) - pad, mx = std::max(a[k], std::max(b[k], c[k])) + pad;
      lo[3 * (size_t)i + k] = mn; hi[3 * (size_t)i + k] = mx; cen[3 * (size_t)i + k] = 0.5f * (mn + mx);
    }
  }
  order.resize(N);
  for (int i = 0; i < N; ++i) order[i] = i;
  nodes.clear(); nodes.reserve(2 * (size_t)N);
  struct Task { int node, first, count; };
  std::vector<Task> todo;
  nodes.push_back(BvhNode{});
  todo.push_back({0, 0, N});
  while (!todo.empty()) {
    Task tk = todo.back(); todo.pop_back();
    BvhNode nd{};
    float clo[3] = {1e30f, 1e30f, 1e30f}, chi[3] = {-1e30f, -1e30f, -1e30f};
    for (int k = 0; k < 3; ++k) { nd.lo[k] = 1e30f; nd.hi[k] = -1e30f; }
    for (int i = 0; i < tk.count; ++i) {
      int t = order[tk.first + i];
      for (int k = 0; k < 3; ++k) {
        nd.lo[k] = std::min(nd.lo[k], lo[3 * (size_t)t + k]); nd.hi[k] = std::max(nd.hi[k], hi[3 * (size_t)t + k]);
        clo[k] = std::min(clo[k], cen[3 * (size_t)t + k]); chi[k] = std::max(chi[k], cen[3 * (size_t)t + k]);
      }
    }
    nd.first = tk.first; nd.count = tk.count; nd.left = nd.right = -1;
    if (tk.count > 4) {
      const int NB = 16;
      float bestCost = 1e30f; int bestAxis = -1, bestBin = -1;
      for (int ax = 0; ax < 3; ++ax) {
        float ext = chi[ax] - clo[ax];
        if (!(ext > 0.f)) continue;
        int cnt[NB] = {0}; float blo[NB][3], bhi[NB][3];
        for (int b = 0; b < NB; ++b) for (int k = 0; k < 3; ++k) { blo[b][k] = 1e30f; bhi[b][k] = -1e30f; }
        float scale = NB / ext;
        for (int i = 0; i < tk.count; ++i) {
          int t = order[tk.first + i];
          int b = std::min(NB - 1, (int)((cen[3 * (size_t)t + ax] - clo[ax]) * scale));
          cnt[b]++;
          for (int k = 0; k < 3; ++k) { blo[b][k] = std::min(blo[b][k], lo[3 * (size_t)t + k]); bhi[b][k] = std::max(bhi[b][k], hi[3 * (size_t)t + k]); }
        }
        float la[NB], ra[NB]; int lc[NB], rc[NB];
        float l0[3] = {1e30f, 1e30f, 1e30f}, h0[3] = {-1e30f, -1e30f, -1e30f}; int c = 0;
        auto area = [](const float* l, const float* h) { float x = h[0] - l[0], y = h[1] - l[1], z = h[2] - l[2]; return (x < 0) ? 0.f : x * y + y * z + z * x; };
        for (int b = 0; b < NB; ++b) { for (int k = 0; k < 3; ++k) { l0[k] = std::min(l0[k], blo[b][k]); h0[k] = std::max(h0[k], bhi[b][k]); } c += cnt[b]; la[b] = area(l0, h0); lc[b] = c; }
        for (int k = 0; k < 3; ++k) { l0[k] = 1e30f; h0[k] = -1e30f; } c = 0;
        for (int b = NB - 1; b >= 0; --b) { for (int k = 0; k < 3; ++k) { l0[k] = std::min(l0[k], blo[b][k]); h0[k] = std::max(h0[k], bhi[b][k]); } c += cnt[b]; ra[b] = area(l0, h0); rc[b] = c; }
        for (int b = 0; b < NB - 1; ++b) {
          if (lc[b] == 0 || rc[b + 1] == 0) continue;
          float cost = la[b] * lc[b] + ra[b + 1] * rc[b + 1];
          if (cost < bestCost) { bestCost = cost; bestAxis = ax; bestBin = b; }
        }
      }
      int mid = -1;
      if (bestAxis >= 0) {
        float ext = chi[bestAxis] - clo[bestAxis]; float scale = NB / ext;
        auto it = std::partition(order.begin() + tk.first, order.begin() + tk.first + tk.count, [&](int t) {
          int b = std::min(NB - 1, (int)((cen[3 * (size_t)t + bestAxis] - clo[bestAxis]) * scale));
          return b <= bestBin;
        });
        mid = (int)(it - order.begin());
      }
      if (mid <= tk.first || mid >= tk.first + tk.count) {
        if (tk.count > 16) {   // degenerate centroids: median split on the largest axis
          int ax = 0; for (int k = 1; k < 3; ++k) if (nd.hi[k] - nd.lo[k] > nd.hi[ax] - nd.lo[ax]) ax = k;
          mid = tk.first + tk.count / 2;
          std::nth_element(order.begin() + tk.first, order.begin() + mid, order.begin() + tk.first + tk.count,
                           [&](int a, int b) { return cen[3 * (size_t)a + ax] < cen[3 * (size_t)b + ax]; });
        } else mid = -1;
      }
      if (mid > 0) {
        nd.count = 0;
        nd.left = (int)nodes.size(); nodes.push_back(BvhNode{});
        nd.right = (int)nodes.size(); nodes.push_back(BvhNode{});
        todo.push_back({nd.left, tk.first, mid - tk.first});
        todo.push_back({nd.right, mid, tk.first + tk.count - mid});
      }
    }
    nodes[tk.node] = nd;
  }
}

}  // namespace orc

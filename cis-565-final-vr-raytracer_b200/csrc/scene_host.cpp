// scene_host.cpp — builds every table the kernels read from the imported glTF scene.
// Reference: Scene::load and its helpers (src/scene.cpp:57-125, 179-195, 209-289, 319-448, 700-772),
// the alias-method table (src/alias_table.hpp:21-63), TLAS instance flags (src/accelstruct.cpp:132-162)
// and Scene::updateCamera (src/scene.cpp:777-826).
#include "scene_host.h"
#include <cmath>
#include <map>
#include "common.h"
#include "eid_vecmath.h"
#include "pack.h"

namespace eid {

// Vose-style alias table with two explicit stacks, in the reference's exact push/pop order
// (alias_table.hpp:21-63) so that (q, alias) per cell — and therefore every light pick — matches.
static void buildAliasCells(const std::vector<float>& weights, std::vector<ImptSampData>& cells) {
  const size_t n = weights.size();
  cells.assign(n, ImptSampData{});
  if (!n) return;
  float total = 0.f;
  for (float w : weights) total += w;
  const float scale = static_cast<float>(n) / total;
  struct Cell { float p; int id; };
  std::vector<Cell> big(2 * n), small(2 * n);
  int nb = 0, ns = 0;
  for (size_t i = 0; i < n; ++i) {
    float p = weights[i] * scale;
    if (p > 1.0f) big[nb++] = {p, (int)i}; else small[ns++] = {p, (int)i};
  }
  std::vector<Cell> table(n);
  while (nb && ns) {
    Cell g = big[--nb], s = small[--ns];
    table[s.id] = {s.p, g.id};
    g.p -= (1.0f - s.p);
    if (g.p > 1.0f) big[nb++] = g; else small[ns++] = g;
  }
  for (int i = nb - 1; i >= 0; --i) table[big[i].id] = big[i];
  for (int i = ns - 1; i >= 0; --i) table[small[i].id] = small[i];
  for (size_t i = 0; i < n; ++i) {
    cells[i].alias = table[i].id;
    cells[i].q = table[i].p;
    cells[i].pdf = weights[i] / total;
    cells[i].aliasPdf = weights[table[i].id] / total;
  }
}

static inline void xformPoint(const float* m /*4x4 col-major*/, const float* p, float* o) {
  for (int r = 0; r < 3; ++r) o[r] = ((m[r] * p[0] + m[4 + r] * p[1]) + m[8 + r] * p[2]) + m[12 + r];
}

// Scene::createTextureImages + gltfSamplerToVulkan (scene.cpp:513-646): missing images / bad sources become a 1x1 white texel,
// a file without images gets one default texture, sampler enums outside the filter map fall back to NEAREST (std::map default).
static void buildTextureTable(const HostGltf& g, std::vector<TextureHost>& out, std::vector<uint32_t>& texels) {
  out.clear(); texels.clear();
  texels.push_back(0xFFFFFFFFu);                       // texel 0: the shared white default
  if (g.images.empty()) { out.push_back(TextureHost{}); return; }
  auto isLinear = [](int f) { return (f == 9729 || f == 9985 || f == 9987) ? 1 : 0; };
  auto wrapMode = [](int w) { return w == 33071 ? 2 : (w == 33648 ? 1 : 0); };
  std::vector<int64_t> imageOffset(g.images.size(), -1);
  for (const auto& t : g.textures) {
    TextureHost h;
    if (t.image >= 0 && (size_t)t.image < g.images.size()) {
      const auto& im = g.images[t.image];
      if (!im.rgba8.empty()) {
        if (imageOffset[t.image] < 0) {
          imageOffset[t.image] = (int64_t)texels.size();
          const size_t n = (size_t)im.width * im.height;
          texels.resize(texels.size() + n);
          memcpy(&texels[imageOffset[t.image]], im.rgba8.data(), 4 * n);
        }
        h.width = im.width; h.height = im.height; h.texelOffset = (uint64_t)imageOffset[t.image];
      }
      if (t.hasSampler) { h.linear = isLinear(t.magFilter); h.wrapS = wrapMode(t.wrapS); h.wrapT = wrapMode(t.wrapT); }
    }
    out.push_back(h);
  }
}

void SceneHost::build() {
  const HostGltf& g = gltf;
  buildTextureTable(g, textures, texels);
  hasTextures = false;
  for (const auto& m : g.materials)
    for (int t : {m.baseColorTexture, m.metallicRoughnessTexture, m.emissiveTexture, m.normalTexture, m.transmissionTexture}) {
      if (t >= (int)textures.size()) raise(EID_ERR_INVALID, "material references texture %d but only %zu exist", t, textures.size());
      if (t > -1) hasTextures = true;
    }
  // ---- materials (scene.cpp:415-448)
  materials.clear();
  for (const auto& m : g.materials) {
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
    float ior = m.ior;
    if (ior < 1.f) ior = 1.f;
    if (ior > MAX_IOR_MINUS_ONE + 1.f) ior = MAX_IOR_MINUS_ONE + 1.f;
    s.ior = ior;
    s.alphaMode = m.alphaMode;
    s.alphaCutoff = m.alphaCutoff;
    materials.push_back(s);
  }

  // ---- punctual lights + alias cells (scene.cpp:319-353, 700-728)
  puncLights.clear();
  puncLightWeight = trigLightWeight = 0.f;
  for (const auto& l : g.lights) {
    PuncLight p{};
    p.position = {l.worldMatrix[12], l.worldMatrix[13], l.worldMatrix[14]};
    p.direction = {-l.worldMatrix[8], -l.worldMatrix[9], -l.worldMatrix[10]};
    p.color = {l.color[0], l.color[1], l.color[2]};
    p.innerConeCos = static_cast<float>(std::cos((double)l.innerConeAngle));
    p.outerConeCos = static_cast<float>(std::cos((double)l.outerConeAngle));
    p.range = l.range;
    p.intensity = l.intensity;
    p.type = l.type;
    puncLights.push_back(p);
  }
  lightInfo = LightBufInfo{};
  lightInfo.puncLightSize = (uint32_t)puncLights.size();
  if (!puncLights.empty()) {
    std::vector<float> w;
    for (const auto& p : puncLights) {
      float power = lum709(p.color.x, p.color.y, p.color.z) * p.intensity * 3.1416f * 4.f;
      w.push_back(power);
      puncLightWeight += power;
    }
    std::vector<ImptSampData> cells;
    buildAliasCells(w, cells);
    for (size_t i = 0; i < cells.size(); ++i) puncLights[i].impSamp = cells[i];
  } else {
    puncLights.push_back(PuncLight{});   // the buffer "cannot be null"
  }

  // ---- compressed vertices + indices (scene.cpp:209-289) and InstanceData bases (scene.cpp:179-195)
  vertices.clear(); indices.clear(); vtxBase.clear(); idxBase.clear();
  std::map<std::pair<uint32_t, uint32_t>, uint64_t> cache;   // (vertexOffset, vertexCount) -> base
  for (const auto& pm : g.primMeshes) {
    auto key = std::make_pair(pm.vertexOffset, pm.vertexCount);
    auto it = cache.find(key);
    if (it == cache.end()) {
      uint64_t base = vertices.size();
      for (uint32_t v = 0; v < pm.vertexCount; ++v) {
        size_t i = (size_t)pm.vertexOffset + v;
        VertexAttributes a{};
        a.position = {g.positions[3 * i], g.positions[3 * i + 1], g.positions[3 * i + 2]};
        a.normal = octEncode(g.normals[3 * i], g.normals[3 * i + 1], g.normals[3 * i + 2]);
        a.tangent = octEncode(g.tangents[4 * i], g.tangents[4 * i + 1], g.tangents[4 * i + 2]);
        a.color = packUnorm4(g.colors0[4 * i], g.colors0[4 * i + 1], g.colors0[4 * i + 2], g.colors0[4 * i + 3]);
        uint32_t vbits = eid_f2u(g.texcoords0[2 * i + 1]);
        vbits = (g.tangents[4 * i + 3] > 0) ? (vbits | 1u) : (vbits & ~1u);   // tangent handedness in the LSB of v
        a.texcoord = {g.texcoords0[2 * i], eid_u2f(vbits)};
        vertices.push_back(a);
      }
      cache[key] = base;
      vtxBase.push_back(base);
    } else {
      vtxBase.push_back(it->second);
    }
    idxBase.push_back(indices.size());
    indices.insert(indices.end(), g.indices.begin() + pm.firstIndex, g.indices.begin() + pm.firstIndex + pm.indexCount);
  }

  // ---- emissive-triangle lights (scene.cpp:355-409, 742-772): one record per triangle of every node whose
  // material has luminance(emissiveFactor) > 1e-2, vertices in world space, weight = luminance only.
  trigLights.clear();
  std::vector<float> tw;
  for (const auto& node : g.nodes) {
    const auto& pm = g.primMeshes[node.primMesh];
    const auto& mat = g.materials[pm.materialIndex];
    float power = lum709(mat.emissiveFactor[0], mat.emissiveFactor[1], mat.emissiveFactor[2]);
    if (!(power > 1e-2f)) continue;
    for (uint32_t k = pm.firstIndex; k < pm.firstIndex + pm.indexCount - 1; k += 3) {   // the reference's loop bound
      uint32_t id[3] = {g.indices[k] + pm.vertexOffset, g.indices[k + 1] + pm.vertexOffset, g.indices[k + 2] + pm.vertexOffset};
      TrigLight t{};
      t.matIndex = (uint32_t)pm.materialIndex;
      t.transformIndex = 0xFFFFFFFFu;   // size()-1 of a vector that is never filled (scene.cpp:380)
      float w[3][3];
      for (int c = 0; c < 3; ++c) xformPoint(node.worldMatrix, &g.positions[3 * (size_t)id[c]], w[c]);
      t.v0 = {w[0][0], w[0][1], w[0][2]}; t.v1 = {w[1][0], w[1][1], w[1][2]}; t.v2 = {w[2][0], w[2][1], w[2][2]};
      t.uv0 = {g.texcoords0[2 * (size_t)id[0]], g.texcoords0[2 * (size_t)id[0] + 1]};
      t.uv1 = {g.texcoords0[2 * (size_t)id[1]], g.texcoords0[2 * (size_t)id[1] + 1]};
      t.uv2 = {g.texcoords0[2 * (size_t)id[2]], g.texcoords0[2 * (size_t)id[2] + 1]};
      trigLights.push_back(t);
      tw.push_back(power);
      trigLightWeight += power;
    }
  }
  lightInfo.trigLightSize = (uint32_t)trigLights.size();
  if (!trigLights.empty()) {
    std::vector<ImptSampData> cells;
    buildAliasCells(tw, cells);
    for (size_t i = 0; i < cells.size(); ++i) trigLights[i].impSamp = cells[i];
  } else {
    trigLights.push_back(TrigLight{});
  }
  if (lightInfo.puncLightSize > 0 || lightInfo.trigLightSize > 0)   // scene.cpp:101-103
    lightInfo.trigSampProb = trigLightWeight / (trigLightWeight + puncLightWeight);

  // ---- TLAS instances (accelstruct.cpp:132-162)
  instances.clear();
  triangleInstances = 0;
  hasNonOpaque = false;
  for (const auto& node : g.nodes) {
    const auto& pm = g.primMeshes[node.primMesh];
    const auto& mat = g.materials[pm.materialIndex];
    InstanceXform x{};
    eid_mat4 wm;
    memcpy(wm.m, node.worldMatrix, sizeof(wm.m));
    eid_mat4 inv = eid_mat4_invert(wm);
    for (int c = 0; c < 4; ++c) for (int r = 0; r < 3; ++r) { x.objectToWorld[c * 3 + r] = wm.m[c * 4 + r]; x.worldToObject[c * 3 + r] = inv.m[c * 4 + r]; }
    x.primMesh = node.primMesh;
    bool opaque = mat.alphaMode == 0 || (mat.baseColorFactor[3] == 1.0f && mat.baseColorTexture == -1);
    if (opaque) x.flags |= INST_FORCE_OPAQUE; else hasNonOpaque = true;
    if (mat.doubleSided == 1) x.flags |= INST_CULL_DISABLE;
    // det of the upper 3x3 with the contract's dot/cross order: facing is decided in object space
    const float* c0 = &x.objectToWorld[0]; const float* c1 = &x.objectToWorld[3]; const float* c2 = &x.objectToWorld[6];
    float cr[3] = {c1[1] * c2[2] - c1[2] * c2[1], c1[2] * c2[0] - c1[0] * c2[2], c1[0] * c2[1] - c1[1] * c2[0]};
    float det = (c0[0] * cr[0] + c0[1] * cr[1]) + c0[2] * cr[2];
    if (det < 0.0f) x.flags |= INST_MIRROR;
    x.firstTriangle = (uint32_t)triangleInstances;
    x.triangleCount = pm.indexCount / 3;
    triangleInstances += x.triangleCount;
    instances.push_back(x);
  }
  if (triangleInstances >= (1ull << 28)) raise(EID_ERR_UNSUPPORTED, "more than 2^28 triangle instances");

  // ---- camera defaults (scene.cpp:295-314, 78)
  if (g.hasCamera) {
    for (int i = 0; i < 3; ++i) { eye[i] = g.camEye[i]; center[i] = g.camCenter[i]; up[i] = g.camUp[i]; }
    fovDeg = g.camYfovRad * 57.29577951308232f;
  }
  camera = SceneCamera{};
  camera.nbLights = (int)g.lights.size();
  prevEye[0] = prevEye[1] = prevEye[2] = 0.f;
}

void SceneHost::updateCamera(uint32_t w, uint32_t h) {
  const float aspect = w / (float)h;
  eid_mat4 view = eid_look_at({eye[0], eye[1], eye[2]}, {center[0], center[1], center[2]}, {up[0], up[1], up[2]});
  eid_mat4 proj = eid_perspectiveVK(fovDeg, aspect, CAMERA_NEAR, CAMERA_FAR);
  proj.m[8] += .5f / w;    // a02: constant sub-pixel shift (scene.cpp:783-787)
  proj.m[9] += .5f / h;    // a12
  camera.lastProjView = camera.projView;
  camera.lastView = eid_mat4_invert(camera.viewInverse);
  camera.viewInverse = eid_mat4_invert(view);
  camera.projInverse = eid_mat4_invert(proj);
  camera.projView = eid_mat4_mul(proj, view);
  camera.lastPosition = {prevEye[0], prevEye[1], prevEye[2]};   // the previous call's eye (function-static in the reference)
  for (int i = 0; i < 3; ++i) prevEye[i] = eye[i];
}

}  // namespace eid

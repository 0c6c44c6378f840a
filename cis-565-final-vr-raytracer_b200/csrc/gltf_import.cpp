// gltf_import.cpp — minimal JSON parser + glTF 2.0 importer (host only).
//
// Covers what the reference's loader path (scene.cpp:130-173 tinygltf, scene.cpp:72-74 nvh::GltfScene)
// feeds into the hot path: node hierarchy flattening, TRIANGLES primitives with POSITION / NORMAL /
// TANGENT / TEXCOORD_0 / COLOR_0 (+ defaults for the missing ones), uint8/16/32 indices, materials with
// KHR_materials_ior / KHR_materials_transmission, KHR_lights_punctual, cameras; .gltf with external
// .bin or data: URIs, and .glb; textures / samplers / images, PNG images decoded in place (png_decode.cpp), other formats provided
// by the host.  Every index and byte range taken from the file is validated: a malformed file is EID_ERR_PARSE, never an
// out-of-bounds access.
#include "gltf_import.h"
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include "common.h"

namespace eid {

// ------------------------------------------------------------------------------------------------
// JSON
// ------------------------------------------------------------------------------------------------
struct JValue;
using JPtr = std::shared_ptr<JValue>;
struct JValue {
  enum Type { Null, Bool, Num, Str, Arr, Obj } type = Null;
  bool b = false;
  double num = 0;
  std::string str;
  std::vector<JPtr> arr;
  std::vector<std::pair<std::string, JPtr>> obj;

  const JValue* get(const char* key) const {
    if (type != Obj) return nullptr;
    for (auto& kv : obj) if (kv.first == key) return kv.second.get();
    return nullptr;
  }
  bool has(const char* key) const { return get(key) != nullptr; }
  double number(const char* key, double def) const { auto v = get(key); return (v && v->type == Num) ? v->num : def; }
  int integer(const char* key, int def) const { auto v = get(key); return (v && v->type == Num) ? (int)v->num : def; }
  bool boolean(const char* key, bool def) const { auto v = get(key); return (v && v->type == Bool) ? v->b : def; }
  std::string string(const char* key, const char* def) const { auto v = get(key); return (v && v->type == Str) ? v->str : std::string(def); }
  size_t size() const { return type == Arr ? arr.size() : 0; }
  const JValue& at(size_t i) const {
    if (type != Arr || i >= arr.size()) raise(EID_ERR_PARSE, "glTF: array index %zu out of range (size %zu)", i, size());
    return *arr[i];
  }
  // element i as an index into another glTF array (must be a non-negative integral number)
  int indexAt(size_t i) const {
    const JValue& v = at(i);
    if (v.type != Num || !(v.num >= 0) || v.num > 2147483647.0 || v.num != (double)(int)v.num) raise(EID_ERR_PARSE, "glTF: element %zu is not a valid index", i);
    return (int)v.num;
  }
  // a byte count / offset / element count: non-negative, integral, below 2^48
  size_t sizeField(const char* key, size_t def) const {
    auto v = get(key);
    if (!v) return def;
    if (v->type != Num || !(v->num >= 0) || v->num > 281474976710656.0 || v->num != (double)(uint64_t)v->num) raise(EID_ERR_PARSE, "glTF: '%s' is not a valid size", key);
    return (size_t)v->num;
  }
};

class JParser {
 public:
  JParser(const char* p, size_t n) : p_(p), e_(p + n) {}
  JPtr parse() {
    JPtr v = value();
    ws();
    if (p_ != e_) err("trailing characters");
    return v;
  }

 private:
  const char *p_, *e_;
  int depth_ = 0;
  struct Nest { int& d; explicit Nest(int& x) : d(x) { if (++d > 200) raise(EID_ERR_PARSE, "JSON parse error: nesting deeper than 200 levels"); } ~Nest() { --d; } };
  [[noreturn]] void err(const char* m) { raise(EID_ERR_PARSE, "JSON parse error: %s", m); }
  void ws() { while (p_ < e_ && (*p_ == ' ' || *p_ == '\n' || *p_ == '\r' || *p_ == '\t')) ++p_; }
  JPtr value() {
    Nest nest(depth_);
    ws();
    if (p_ >= e_) err("unexpected end");
    auto v = std::make_shared<JValue>();
    char c = *p_;
    if (c == '{') {
      v->type = JValue::Obj; ++p_; ws();
      if (p_ < e_ && *p_ == '}') { ++p_; return v; }
      for (;;) {
        ws();
        std::string k = str();
        ws();
        if (p_ >= e_ || *p_ != ':') err("expected ':'");
        ++p_;
        v->obj.emplace_back(std::move(k), value());
        ws();
        if (p_ < e_ && *p_ == ',') { ++p_; continue; }
        if (p_ < e_ && *p_ == '}') { ++p_; break; }
        err("expected ',' or '}'");
      }
    } else if (c == '[') {
      v->type = JValue::Arr; ++p_; ws();
      if (p_ < e_ && *p_ == ']') { ++p_; return v; }
      for (;;) {
        v->arr.push_back(value());
        ws();
        if (p_ < e_ && *p_ == ',') { ++p_; continue; }
        if (p_ < e_ && *p_ == ']') { ++p_; break; }
        err("expected ',' or ']'");
      }
    } else if (c == '"') {
      v->type = JValue::Str; v->str = str();
    } else if (c == 't' && e_ - p_ >= 4 && !strncmp(p_, "true", 4)) { v->type = JValue::Bool; v->b = true; p_ += 4;
    } else if (c == 'f' && e_ - p_ >= 5 && !strncmp(p_, "false", 5)) { v->type = JValue::Bool; v->b = false; p_ += 5;
    } else if (c == 'n' && e_ - p_ >= 4 && !strncmp(p_, "null", 4)) { p_ += 4;
    } else {
      char* end = nullptr;
      std::string tmp(p_, std::min<size_t>(64, e_ - p_));
      double d = strtod(tmp.c_str(), &end);
      if (end == tmp.c_str()) err("bad token");
      v->type = JValue::Num; v->num = d; p_ += (end - tmp.c_str());
    }
    return v;
  }
  std::string str() {
    if (p_ >= e_ || *p_ != '"') err("expected string");
    ++p_;
    std::string s;
    while (p_ < e_ && *p_ != '"') {
      if (*p_ == '\\') {
        ++p_;
        if (p_ >= e_) err("bad escape");
        switch (*p_) {
          case 'n': s += '\n'; break; case 't': s += '\t'; break; case 'r': s += '\r'; break;
          case 'b': s += '\b'; break; case 'f': s += '\f'; break;
          case 'u': {
            if (e_ - p_ < 5) err("bad \\u");
            unsigned cp = (unsigned)strtoul(std::string(p_ + 1, 4).c_str(), nullptr, 16);
            p_ += 4;
            if (cp < 0x80) s += (char)cp;
            else if (cp < 0x800) { s += (char)(0xC0 | (cp >> 6)); s += (char)(0x80 | (cp & 0x3F)); }
            else { s += (char)(0xE0 | (cp >> 12)); s += (char)(0x80 | ((cp >> 6) & 0x3F)); s += (char)(0x80 | (cp & 0x3F)); }
            break;
          }
          default: s += *p_;
        }
        ++p_;
      } else s += *p_++;
    }
    if (p_ >= e_) err("unterminated string");
    ++p_;
    return s;
  }
};

// ------------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------------
static std::vector<uint8_t> readFile(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) raise(EID_ERR_IO, "cannot open '%s'", path.c_str());
  f.seekg(0, std::ios::end);
  std::streamoff n = f.tellg();
  f.seekg(0);
  std::vector<uint8_t> d((size_t)n);
  if (n) f.read((char*)d.data(), n);
  return d;
}
static std::vector<uint8_t> base64Decode(const char* s, size_t n) {
  static int8_t T[256]; static bool init = false;
  if (!init) {
    memset(T, -1, sizeof(T));
    const char* A = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
    for (int i = 0; i < 64; ++i) T[(uint8_t)A[i]] = (int8_t)i;
    init = true;
  }
  std::vector<uint8_t> out; out.reserve(n * 3 / 4);
  uint32_t acc = 0; int bits = 0;
  for (size_t i = 0; i < n; ++i) {
    int v = T[(uint8_t)s[i]];
    if (v < 0) continue;
    acc = (acc << 6) | (uint32_t)v; bits += 6;
    if (bits >= 8) { bits -= 8; out.push_back((uint8_t)((acc >> bits) & 0xff)); }
  }
  return out;
}
static std::string dirOf(const std::string& p) {
  size_t k = p.find_last_of("/\\");
  return k == std::string::npos ? std::string(".") : p.substr(0, k);
}

// column-major 4x4 helpers (double precision for hierarchy composition, rounded once at the end)
struct M4 { double m[16]; };
static M4 identity() { M4 r{}; r.m[0] = r.m[5] = r.m[10] = r.m[15] = 1; return r; }
static M4 mul(const M4& a, const M4& b) {
  M4 r{};
  for (int c = 0; c < 4; ++c) for (int rr = 0; rr < 4; ++rr) {
    double s = 0; for (int k = 0; k < 4; ++k) s += a.m[k * 4 + rr] * b.m[c * 4 + k];
    r.m[c * 4 + rr] = s;
  }
  return r;
}
static M4 nodeLocal(const JValue& n) {
  if (auto m = n.get("matrix")) {
    if (m->size() != 16) raise(EID_ERR_PARSE, "node.matrix must have 16 elements");
    M4 r; for (int i = 0; i < 16; ++i) r.m[i] = m->at(i).num; return r;
  }
  double t[3] = {0, 0, 0}, q[4] = {0, 0, 0, 1}, s[3] = {1, 1, 1};
  if (auto v = n.get("translation")) for (int i = 0; i < 3 && i < (int)v->size(); ++i) t[i] = v->at(i).num;
  if (auto v = n.get("rotation")) for (int i = 0; i < 4 && i < (int)v->size(); ++i) q[i] = v->at(i).num;
  if (auto v = n.get("scale")) for (int i = 0; i < 3 && i < (int)v->size(); ++i) s[i] = v->at(i).num;
  double x = q[0], y = q[1], z = q[2], w = q[3];
  M4 r = identity();
  r.m[0] = (1 - 2 * (y * y + z * z)) * s[0]; r.m[1] = (2 * (x * y + z * w)) * s[0]; r.m[2] = (2 * (x * z - y * w)) * s[0];
  r.m[4] = (2 * (x * y - z * w)) * s[1]; r.m[5] = (1 - 2 * (x * x + z * z)) * s[1]; r.m[6] = (2 * (y * z + x * w)) * s[1];
  r.m[8] = (2 * (x * z + y * w)) * s[2]; r.m[9] = (2 * (y * z - x * w)) * s[2]; r.m[10] = (1 - 2 * (x * x + y * y)) * s[2];
  r.m[12] = t[0]; r.m[13] = t[1]; r.m[14] = t[2];
  return r;
}

struct Doc {
  JPtr root;
  std::vector<std::vector<uint8_t>> buffers;
  const JValue& top(const char* k) const {
    static JValue empty;
    auto v = root->get(k);
    return v ? *v : empty;
  }
};

static int componentSize(int ct) {
  switch (ct) { case 5120: case 5121: return 1; case 5122: case 5123: return 2; case 5125: case 5126: return 4; }
  raise(EID_ERR_UNSUPPORTED, "accessor componentType %d", ct);
}
static int typeComponents(const std::string& t) {
  if (t == "SCALAR") return 1; if (t == "VEC2") return 2; if (t == "VEC3") return 3; if (t == "VEC4") return 4;
  raise(EID_ERR_UNSUPPORTED, "accessor type %s", t.c_str());
}

// Start of accessor `idx`'s data inside its buffer after validating bufferView / buffer indices and that `count` elements of
// `elemBytes` at the view's stride lie inside both the bufferView and the buffer (overflow-safe: every quantity is < 2^48).
static size_t accessorStride(const Doc& d, int bv, size_t packed) {
  const size_t stride = d.top("bufferViews").at((size_t)bv).sizeField("byteStride", 0);
  if (stride && (stride < packed || stride > 65536)) raise(EID_ERR_PARSE, "bufferView %d: byteStride %zu is invalid", bv, stride);
  return stride ? stride : packed;
}
static const uint8_t* accessorData(const Doc& d, const JValue& a, int idx, int bv, size_t count, size_t elemBytes, size_t packed) {
  const JValue& views = d.top("bufferViews");
  if ((size_t)bv >= views.size()) raise(EID_ERR_PARSE, "accessor %d: bufferView %d out of range", idx, bv);
  const JValue& view = views.at((size_t)bv);
  const int buf = view.integer("buffer", -1);
  if (buf < 0 || (size_t)buf >= d.buffers.size()) raise(EID_ERR_PARSE, "bufferView %d: buffer %d out of range", bv, buf);
  const auto& B = d.buffers[(size_t)buf];
  const size_t vOff = view.sizeField("byteOffset", 0), vLen = view.sizeField("byteLength", B.size() > vOff ? B.size() - vOff : 0);
  const size_t aOff = a.sizeField("byteOffset", 0);
  const size_t stride = accessorStride(d, bv, packed);
  if (vOff > B.size() || vLen > B.size() - vOff) raise(EID_ERR_PARSE, "bufferView %d overruns buffer %d", bv, buf);
  if (count > (size_t)1 << 40) raise(EID_ERR_PARSE, "accessor %d: count %zu is not plausible", idx, count);
  const size_t need = count ? aOff + (count - 1) * stride + elemBytes : 0;     // < 2^48 + 2^40 * 2^16 + ...: no wrap in 64 bits
  if (need > vLen) raise(EID_ERR_PARSE, "accessor %d overruns its bufferView", idx);
  return B.data() + vOff + aOff;
}

// Reads accessor `idx` as floats with `want` components per element (extra components dropped, missing
// ones filled with `fill`); integer types are converted (normalized -> [0,1] / [-1,1]).
static size_t readAccessorFloat(const Doc& d, int idx, int want, float fill, std::vector<float>& out) {
  const JValue& accs = d.top("accessors");
  if (idx < 0 || (size_t)idx >= accs.size()) raise(EID_ERR_PARSE, "accessor %d out of range", idx);
  const JValue& a = accs.at(idx);
  if (a.has("sparse")) raise(EID_ERR_UNSUPPORTED, "sparse accessors are not supported");
  int ct = a.integer("componentType", 5126);
  int nc = typeComponents(a.string("type", "SCALAR"));
  size_t count = a.sizeField("count", 0);
  bool normalized = a.boolean("normalized", false);
  int bv = a.integer("bufferView", -1);
  if (bv < 0) raise(EID_ERR_UNSUPPORTED, "accessor without bufferView");
  int cs = componentSize(ct);
  const uint8_t* data = accessorData(d, a, idx, bv, count, (size_t)cs * nc, (size_t)cs * nc);
  const size_t stride = accessorStride(d, bv, (size_t)cs * nc);
  size_t base = out.size();
  out.resize(base + count * want);
  for (size_t i = 0; i < count; ++i) {
    const uint8_t* p = data + i * stride;
    for (int c = 0; c < want; ++c) {
      float v = fill;
      if (c < nc) {
        switch (ct) {
          case 5126: { float f; memcpy(&f, p + 4 * c, 4); v = f; break; }
          case 5121: v = normalized ? p[c] / 255.0f : (float)p[c]; break;
          case 5123: { uint16_t u; memcpy(&u, p + 2 * c, 2); v = normalized ? u / 65535.0f : (float)u; break; }
          case 5120: { int8_t s = (int8_t)p[c]; v = normalized ? std::max(s / 127.0f, -1.0f) : (float)s; break; }
          case 5122: { int16_t s; memcpy(&s, p + 2 * c, 2); v = normalized ? std::max(s / 32767.0f, -1.0f) : (float)s; break; }
          case 5125: { uint32_t u; memcpy(&u, p + 4 * c, 4); v = (float)u; break; }
        }
      }
      out[base + i * want + c] = v;
    }
  }
  return count;
}
static size_t readAccessorIndices(const Doc& d, int idx, std::vector<uint32_t>& out) {
  const JValue& accs = d.top("accessors");
  if (idx < 0 || (size_t)idx >= accs.size()) raise(EID_ERR_PARSE, "index accessor %d out of range", idx);
  const JValue& a = accs.at(idx);
  if (a.has("sparse")) raise(EID_ERR_UNSUPPORTED, "sparse accessors are not supported");
  int ct = a.integer("componentType", 5125);
  size_t count = a.sizeField("count", 0);
  int bv = a.integer("bufferView", -1);
  if (bv < 0) raise(EID_ERR_UNSUPPORTED, "index accessor without bufferView");
  int cs = componentSize(ct);
  const uint8_t* data = accessorData(d, a, idx, bv, count, (size_t)cs, (size_t)cs);
  const size_t stride = accessorStride(d, bv, (size_t)cs);
  size_t base = out.size();
  out.resize(base + count);
  for (size_t i = 0; i < count; ++i) {
    const uint8_t* p = data + i * stride;
    uint32_t v = 0;
    if (ct == 5121) v = p[0];
    else if (ct == 5123) { uint16_t u; memcpy(&u, p, 2); v = u; }
    else if (ct == 5125) memcpy(&v, p, 4);
    else raise(EID_ERR_UNSUPPORTED, "index componentType %d", ct);
    out[base + i] = v;
  }
  return count;
}

// importMaterials (nvh::GltfScene) — glTF defaults
static void importMaterials(const Doc& d, HostGltf& g) {
  const JValue& mats = d.top("materials");
  auto texIndex = [](const JValue* t) { return t ? t->integer("index", -1) : -1; };
  for (size_t i = 0; i < mats.size(); ++i) {
    const JValue& m = mats.at(i);
    eid_material_desc o{};
    o.baseColorFactor[0] = o.baseColorFactor[1] = o.baseColorFactor[2] = o.baseColorFactor[3] = 1.0f;
    o.baseColorTexture = o.metallicRoughnessTexture = o.emissiveTexture = o.normalTexture = o.transmissionTexture = -1;
    o.metallicFactor = 1.0f; o.roughnessFactor = 1.0f;
    o.alphaMode = 0; o.alphaCutoff = 0.5f; o.doubleSided = 0; o.normalTextureScale = 1.0f;
    o.transmissionFactor = 0.0f; o.ior = 1.5f;
    if (auto pbr = m.get("pbrMetallicRoughness")) {
      if (auto f = pbr->get("baseColorFactor")) for (int k = 0; k < 4 && k < (int)f->size(); ++k) o.baseColorFactor[k] = (float)f->at(k).num;
      o.metallicFactor = (float)pbr->number("metallicFactor", 1.0);
      o.roughnessFactor = (float)pbr->number("roughnessFactor", 1.0);
      o.baseColorTexture = texIndex(pbr->get("baseColorTexture"));
      o.metallicRoughnessTexture = texIndex(pbr->get("metallicRoughnessTexture"));
    }
    if (auto f = m.get("emissiveFactor")) for (int k = 0; k < 3 && k < (int)f->size(); ++k) o.emissiveFactor[k] = (float)f->at(k).num;
    o.emissiveTexture = texIndex(m.get("emissiveTexture"));
    if (auto nt = m.get("normalTexture")) { o.normalTexture = nt->integer("index", -1); o.normalTextureScale = (float)nt->number("scale", 1.0); }
    std::string am = m.string("alphaMode", "OPAQUE");
    o.alphaMode = (am == "MASK") ? 1 : (am == "BLEND") ? 2 : 0;
    o.alphaCutoff = (float)m.number("alphaCutoff", 0.5);
    o.doubleSided = m.boolean("doubleSided", false) ? 1 : 0;
    if (auto ext = m.get("extensions")) {
      if (auto e = ext->get("KHR_materials_ior")) o.ior = (float)e->number("ior", 1.5);
      if (auto e = ext->get("KHR_materials_transmission")) {
        o.transmissionFactor = (float)e->number("transmissionFactor", 0.0);
        o.transmissionTexture = texIndex(e->get("transmissionTexture"));
      }
    }
    g.materials.push_back(o);
  }
  if (g.materials.empty()) {   // a default material is appended when the file has none
    eid_material_desc o{};
    o.baseColorFactor[0] = o.baseColorFactor[1] = o.baseColorFactor[2] = o.baseColorFactor[3] = 1.0f;
    o.baseColorTexture = o.metallicRoughnessTexture = o.emissiveTexture = o.normalTexture = o.transmissionTexture = -1;
    o.metallicFactor = 1.0f; o.roughnessFactor = 1.0f; o.alphaCutoff = 0.5f; o.normalTextureScale = 1.0f; o.ior = 1.5f;
    g.materials.push_back(o);
  }
}

static void generateNormals(HostGltf& g, size_t v0, size_t nv, size_t i0, size_t ni) {
  std::vector<double> acc(nv * 3, 0.0);
  for (size_t t = 0; t + 2 < ni; t += 3) {
    uint32_t a = g.indices[i0 + t], b = g.indices[i0 + t + 1], c = g.indices[i0 + t + 2];
    const float* pa = &g.positions[3 * (v0 + a)]; const float* pb = &g.positions[3 * (v0 + b)]; const float* pc = &g.positions[3 * (v0 + c)];
    double e1[3] = {pb[0] - pa[0], pb[1] - pa[1], pb[2] - pa[2]}, e2[3] = {pc[0] - pa[0], pc[1] - pa[1], pc[2] - pa[2]};
    double n[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
    for (uint32_t v : {a, b, c}) for (int k = 0; k < 3; ++k) acc[3 * v + k] += n[k];
  }
  for (size_t v = 0; v < nv; ++v) {
    double l = std::sqrt(acc[3 * v] * acc[3 * v] + acc[3 * v + 1] * acc[3 * v + 1] + acc[3 * v + 2] * acc[3 * v + 2]);
    if (l > 0) for (int k = 0; k < 3; ++k) g.normals[3 * (v0 + v) + k] = (float)(acc[3 * v + k] / l);
    else { g.normals[3 * (v0 + v)] = 0; g.normals[3 * (v0 + v) + 1] = 1; g.normals[3 * (v0 + v) + 2] = 0; }
  }
}
static void generateTangents(HostGltf& g, size_t v0, size_t nv, size_t i0, size_t ni) {
  std::vector<double> tan(nv * 3, 0.0), bit(nv * 3, 0.0);
  for (size_t t = 0; t + 2 < ni; t += 3) {
    uint32_t ia = g.indices[i0 + t], ib = g.indices[i0 + t + 1], ic = g.indices[i0 + t + 2];
    const float* pa = &g.positions[3 * (v0 + ia)]; const float* pb = &g.positions[3 * (v0 + ib)]; const float* pc = &g.positions[3 * (v0 + ic)];
    const float* ua = &g.texcoords0[2 * (v0 + ia)]; const float* ub = &g.texcoords0[2 * (v0 + ib)]; const float* uc = &g.texcoords0[2 * (v0 + ic)];
    double e1[3] = {pb[0] - pa[0], pb[1] - pa[1], pb[2] - pa[2]}, e2[3] = {pc[0] - pa[0], pc[1] - pa[1], pc[2] - pa[2]};
    double du1 = ub[0] - ua[0], dv1 = ub[1] - ua[1], du2 = uc[0] - ua[0], dv2 = uc[1] - ua[1];
    double det = du1 * dv2 - du2 * dv1;
    if (std::fabs(det) < 1e-20) continue;
    double r = 1.0 / det;
    for (uint32_t v : {ia, ib, ic}) for (int k = 0; k < 3; ++k) {
      tan[3 * v + k] += (e1[k] * dv2 - e2[k] * dv1) * r;
      bit[3 * v + k] += (e2[k] * du1 - e1[k] * du2) * r;
    }
  }
  for (size_t v = 0; v < nv; ++v) {
    const float* n = &g.normals[3 * (v0 + v)];
    double t[3] = {tan[3 * v], tan[3 * v + 1], tan[3 * v + 2]};
    double d = t[0] * n[0] + t[1] * n[1] + t[2] * n[2];
    for (int k = 0; k < 3; ++k) t[k] -= n[k] * d;
    double l = std::sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]);
    float w = 1.0f;
    if (l > 1e-12) {
      for (int k = 0; k < 3; ++k) t[k] /= l;
      double c[3] = {n[1] * t[2] - n[2] * t[1], n[2] * t[0] - n[0] * t[2], n[0] * t[1] - n[1] * t[0]};
      w = (c[0] * bit[3 * v] + c[1] * bit[3 * v + 1] + c[2] * bit[3 * v + 2]) < 0 ? -1.0f : 1.0f;
    } else {   // fallback: any vector orthogonal to the normal
      if (std::fabs(n[0]) > std::fabs(n[1])) { double s = std::sqrt((double)n[0] * n[0] + (double)n[2] * n[2]); t[0] = -n[2] / s; t[1] = 0; t[2] = n[0] / s; }
      else { double s = std::sqrt((double)n[1] * n[1] + (double)n[2] * n[2]); t[0] = 0; t[1] = n[2] / s; t[2] = -n[1] / s; }
    }
    float* o = &g.tangents[4 * (v0 + v)];
    o[0] = (float)t[0]; o[1] = (float)t[1]; o[2] = (float)t[2]; o[3] = w;
  }
}

struct ImportCtx {
  const Doc& d;
  HostGltf& g;
  std::map<int, std::vector<int>> meshToPrimMeshes;   // mesh index -> prim mesh indices
  std::map<std::string, std::pair<uint32_t, uint32_t>> vertexCache;   // attribute accessor set -> (vertexOffset, vertexCount)
  struct CamNode { M4 world; float yfov; };
  std::vector<CamNode> cameras;
};

static void importMesh(ImportCtx& c, int meshIdx) {
  if (c.meshToPrimMeshes.count(meshIdx)) return;
  std::vector<int>& list = c.meshToPrimMeshes[meshIdx];
  const JValue& mesh = c.d.top("meshes").at(meshIdx);
  const JValue* prims = mesh.get("primitives");
  if (!prims) return;
  HostGltf& g = c.g;
  for (size_t p = 0; p < prims->size(); ++p) {
    const JValue& prim = prims->at(p);
    if (prim.integer("mode", 4) != 4) continue;   // only TRIANGLES
    const JValue* attrs = prim.get("attributes");
    if (!attrs || !attrs->has("POSITION")) continue;
    int aPos = attrs->integer("POSITION", -1), aNrm = attrs->integer("NORMAL", -1), aTan = attrs->integer("TANGENT", -1);
    int aUv = attrs->integer("TEXCOORD_0", -1), aCol = attrs->integer("COLOR_0", -1);
    std::ostringstream key;
    key << aPos << ":" << aNrm << ":" << aTan << ":" << aUv << ":" << aCol;
    eid_prim_mesh pm{};
    pm.materialIndex = std::max(0, prim.integer("material", -1));
    bool needNormals = false, needTangents = false;
    auto it = c.vertexCache.find(key.str());
    bool fresh = (it == c.vertexCache.end());
    if (fresh) {
      pm.vertexOffset = (uint32_t)(g.positions.size() / 3);
      size_t nv = readAccessorFloat(c.d, aPos, 3, 0.f, g.positions);
      pm.vertexCount = (uint32_t)nv;
      if (aNrm >= 0) { if (readAccessorFloat(c.d, aNrm, 3, 0.f, g.normals) != nv) raise(EID_ERR_PARSE, "NORMAL count mismatch"); }
      else { g.normals.resize(g.normals.size() + 3 * nv, 0.f); needNormals = true; }
      if (aUv >= 0) { if (readAccessorFloat(c.d, aUv, 2, 0.f, g.texcoords0) != nv) raise(EID_ERR_PARSE, "TEXCOORD_0 count mismatch"); }
      else g.texcoords0.resize(g.texcoords0.size() + 2 * nv, 0.f);
      if (aTan >= 0) { if (readAccessorFloat(c.d, aTan, 4, 1.f, g.tangents) != nv) raise(EID_ERR_PARSE, "TANGENT count mismatch"); }
      else { g.tangents.resize(g.tangents.size() + 4 * nv, 0.f); needTangents = true; }
      if (aCol >= 0) { if (readAccessorFloat(c.d, aCol, 4, 1.f, g.colors0) != nv) raise(EID_ERR_PARSE, "COLOR_0 count mismatch"); }
      else g.colors0.resize(g.colors0.size() + 4 * nv, 1.f);
      c.vertexCache[key.str()] = {pm.vertexOffset, pm.vertexCount};
    } else {
      pm.vertexOffset = it->second.first; pm.vertexCount = it->second.second;
    }
    pm.firstIndex = (uint32_t)g.indices.size();
    int aIdx = prim.integer("indices", -1);
    if (aIdx >= 0) pm.indexCount = (uint32_t)readAccessorIndices(c.d, aIdx, g.indices);
    else { for (uint32_t i = 0; i < pm.vertexCount; ++i) g.indices.push_back(i); pm.indexCount = pm.vertexCount; }
    for (uint32_t i = 0; i < pm.indexCount; ++i)
      if (g.indices[pm.firstIndex + i] >= pm.vertexCount) raise(EID_ERR_PARSE, "index out of range in mesh %d", meshIdx);
    if (fresh && needNormals) generateNormals(g, pm.vertexOffset, pm.vertexCount, pm.firstIndex, pm.indexCount);
    if (fresh && needTangents) generateTangents(g, pm.vertexOffset, pm.vertexCount, pm.firstIndex, pm.indexCount);
    list.push_back((int)g.primMeshes.size());
    g.primMeshes.push_back(pm);
  }
}

static void processNode(ImportCtx& c, int nodeIdx, const M4& parent, int depth) {
  if (depth > 256) raise(EID_ERR_PARSE, "node hierarchy too deep (cycle?)");
  const JValue& nodes = c.d.top("nodes");
  if (nodeIdx < 0 || (size_t)nodeIdx >= nodes.size()) raise(EID_ERR_PARSE, "node %d out of range", nodeIdx);
  const JValue& n = nodes.at(nodeIdx);
  M4 world = mul(parent, nodeLocal(n));
  int mesh = n.integer("mesh", -1);
  if (mesh >= 0) {
    if ((size_t)mesh >= c.d.top("meshes").size()) raise(EID_ERR_PARSE, "mesh %d out of range", mesh);
    importMesh(c, mesh);
    for (int pm : c.meshToPrimMeshes[mesh]) {
      eid_node o{};
      for (int i = 0; i < 16; ++i) o.worldMatrix[i] = (float)world.m[i];
      o.primMesh = pm;
      c.g.nodes.push_back(o);
    }
  }
  if (auto ext = n.get("extensions")) {
    if (auto kl = ext->get("KHR_lights_punctual")) {
      int li = kl->integer("light", -1);
      const JValue* lights = nullptr;
      if (auto te = c.d.root->get("extensions")) if (auto e = te->get("KHR_lights_punctual")) lights = e->get("lights");
      if (lights && li >= 0 && (size_t)li < lights->size()) {
        const JValue& L = lights->at(li);
        eid_light_desc o{};
        for (int i = 0; i < 16; ++i) o.worldMatrix[i] = (float)world.m[i];
        std::string t = L.string("type", "point");
        o.type = (t == "directional") ? LightType_Directional : (t == "spot") ? LightType_Spot : LightType_Point;
        o.color[0] = o.color[1] = o.color[2] = 1.0f;
        if (auto col = L.get("color")) for (int k = 0; k < 3 && k < (int)col->size(); ++k) o.color[k] = (float)col->at(k).num;
        o.intensity = (float)L.number("intensity", 1.0);
        o.range = (float)L.number("range", 0.0);
        o.innerConeAngle = 0.0f; o.outerConeAngle = 0.7853981633974483f;
        if (auto sp = L.get("spot")) { o.innerConeAngle = (float)sp->number("innerConeAngle", 0.0); o.outerConeAngle = (float)sp->number("outerConeAngle", 0.7853981633974483); }
        c.g.lights.push_back(o);
      }
    }
  }
  int cam = n.integer("camera", -1);
  if (cam >= 0 && (size_t)cam < c.d.top("cameras").size()) {
    const JValue& C = c.d.top("cameras").at(cam);
    float yfov = 1.0471975512f;
    if (auto p = C.get("perspective")) yfov = (float)p->number("yfov", yfov);
    c.cameras.push_back({world, yfov});
  }
  if (auto ch = n.get("children")) for (size_t i = 0; i < ch->size(); ++i) processNode(c, ch->indexAt(i), world, depth + 1);
}

void HostGltf::computeDimensions() {
  for (int k = 0; k < 3; ++k) { bboxMin[k] = 3.4e38f; bboxMax[k] = -3.4e38f; }
  for (const auto& n : nodes) {
    const auto& pm = primMeshes[n.primMesh];
    const float* m = n.worldMatrix;
    for (uint32_t v = 0; v < pm.vertexCount; ++v) {
      const float* p = &positions[3 * (size_t)(pm.vertexOffset + v)];
      for (int k = 0; k < 3; ++k) {
        float w = m[0 + k] * p[0] + m[4 + k] * p[1] + m[8 + k] * p[2] + m[12 + k];
        bboxMin[k] = std::min(bboxMin[k], w); bboxMax[k] = std::max(bboxMax[k], w);
      }
    }
  }
  if (nodes.empty()) for (int k = 0; k < 3; ++k) bboxMin[k] = bboxMax[k] = 0.f;
}

void HostGltf::fromDesc(const eid_scene_desc& d) {
  if (!d.positions || !d.normals || !d.tangents || !d.texcoords0 || !d.colors0 || !d.indices)
    raise(EID_ERR_INVALID, "eid_scene_desc: null vertex/index arrays");
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
  images.clear(); textures.clear();
  for (uint32_t i = 0; i < d.imageCount; ++i) {
    Image im;
    const eid_image_desc& src = d.images[i];
    if (src.rgba8 && src.width && src.height) { im.width = src.width; im.height = src.height; im.rgba8.assign(src.rgba8, src.rgba8 + 4 * (size_t)src.width * src.height); }
    images.push_back(std::move(im));
  }
  if (d.textureCount) textures.assign(d.textures, d.textures + d.textureCount);
  for (const auto& pm : primMeshes) {
    if ((uint64_t)pm.firstIndex + pm.indexCount > indices.size() || (uint64_t)pm.vertexOffset + pm.vertexCount > d.vertexCount)
      raise(EID_ERR_INVALID, "prim mesh range outside the vertex/index arrays");
    if (pm.materialIndex < 0 || (size_t)pm.materialIndex >= materials.size()) raise(EID_ERR_INVALID, "prim mesh material index out of range");
    for (uint32_t i = 0; i < pm.indexCount; ++i)
      if (indices[pm.firstIndex + i] >= pm.vertexCount) raise(EID_ERR_INVALID, "index out of range");
  }
  for (const auto& n : nodes)
    if (n.primMesh < 0 || (size_t)n.primMesh >= primMeshes.size()) raise(EID_ERR_INVALID, "node prim mesh index out of range");
  hasCamera = d.hasCamera != 0;
  if (hasCamera) {
    for (int i = 0; i < 3; ++i) { camEye[i] = d.camEye[i]; camCenter[i] = d.camCenter[i]; camUp[i] = d.camUp[i]; }
    camYfovRad = d.camYfovRad;
  }
  computeDimensions();
}

void importGltfFile(const std::string& path, HostGltf& g, const std::vector<HostGltf::Image>& provided) {
  g = HostGltf();
  std::vector<uint8_t> file = readFile(path);
  Doc d;
  std::vector<uint8_t> glbBin;
  bool isGlb = file.size() >= 12 && !memcmp(file.data(), "glTF", 4);
  if (isGlb) {
    uint32_t jsonLen, jsonType;
    if (file.size() < 20) raise(EID_ERR_PARSE, "truncated .glb");
    memcpy(&jsonLen, file.data() + 12, 4); memcpy(&jsonType, file.data() + 16, 4);
    if (jsonType != 0x4E4F534A || 20 + (size_t)jsonLen > file.size()) raise(EID_ERR_PARSE, "bad .glb JSON chunk");
    d.root = JParser((const char*)file.data() + 20, jsonLen).parse();
    size_t off = 20 + jsonLen;
    if (off + 8 <= file.size()) {
      uint32_t binLen, binType;
      memcpy(&binLen, file.data() + off, 4); memcpy(&binType, file.data() + off + 4, 4);
      if (binType == 0x004E4942 && off + 8 + binLen <= file.size()) glbBin.assign(file.begin() + off + 8, file.begin() + off + 8 + binLen);
    }
  } else {
    d.root = JParser((const char*)file.data(), file.size()).parse();
  }
  if (d.root->type != JValue::Obj) raise(EID_ERR_PARSE, "glTF root is not an object");
  const JValue& bufs = d.top("buffers");
  for (size_t i = 0; i < bufs.size(); ++i) {
    const JValue& b = bufs.at(i);
    std::string uri = b.string("uri", "");
    if (uri.empty()) {
      if (isGlb && i == 0) d.buffers.push_back(glbBin);
      else raise(EID_ERR_PARSE, "buffer %zu has no uri", i);
    } else if (uri.rfind("data:", 0) == 0) {
      size_t k = uri.find(";base64,");
      if (k == std::string::npos) raise(EID_ERR_UNSUPPORTED, "data: URI without base64");
      d.buffers.push_back(base64Decode(uri.c_str() + k + 8, uri.size() - k - 8));
    } else {
      d.buffers.push_back(readFile(dirOf(path) + "/" + uri));
    }
  }
  importMaterials(d, g);
  {   // textures / samplers / images (tinygltf::Model::textures, samplers, images)
    const JValue& imgs = d.top("images");
    g.images.resize(imgs.size());
    for (size_t i = 0; i < imgs.size() && i < provided.size(); ++i) g.images[i] = provided[i];
    // images the host did not provide: PNG files / data URIs / bufferViews are decoded here (png_decode.cpp); anything else stays empty
    for (size_t i = 0; i < imgs.size(); ++i) {
      if (!g.images[i].rgba8.empty()) continue;
      const JValue& im = imgs.at(i);
      std::vector<uint8_t> bytes;
      const std::string uri = im.string("uri", "");
      if (!uri.empty()) {
        if (uri.rfind("data:", 0) == 0) {
          const size_t k = uri.find(";base64,");
          if (k != std::string::npos) bytes = base64Decode(uri.c_str() + k + 8, uri.size() - k - 8);
        } else {
          try { bytes = readFile(dirOf(path) + "/" + uri); } catch (const Error&) { bytes.clear(); }   // missing file: only an error if the image is used
        }
      } else if (im.has("bufferView")) {
        const int bv = im.integer("bufferView", -1);
        const JValue& views = d.top("bufferViews");
        if (bv < 0 || (size_t)bv >= views.size()) raise(EID_ERR_PARSE, "image %zu: bufferView %d out of range", i, bv);
        const JValue& view = views.at((size_t)bv);
        const int buf = view.integer("buffer", -1);
        if (buf < 0 || (size_t)buf >= d.buffers.size()) raise(EID_ERR_PARSE, "bufferView %d: buffer %d out of range", bv, buf);
        const auto& B = d.buffers[(size_t)buf];
        const size_t off = view.sizeField("byteOffset", 0), len = view.sizeField("byteLength", 0);
        if (off > B.size() || len > B.size() - off) raise(EID_ERR_PARSE, "image %zu overruns buffer %d", i, buf);
        bytes.assign(B.begin() + off, B.begin() + off + len);
      }
      if (isPng(bytes.data(), bytes.size())) decodePng(bytes.data(), bytes.size(), g.images[i]);
    }
    const JValue& texs = d.top("textures");
    const JValue& samps = d.top("samplers");
    for (size_t i = 0; i < texs.size(); ++i) {
      eid_texture_desc t{};
      t.image = texs.at(i).integer("source", -1);
      int si = texs.at(i).integer("sampler", -1);
      t.hasSampler = (si >= 0 && (size_t)si < samps.size()) ? 1 : 0;
      t.magFilter = t.minFilter = -1; t.wrapS = t.wrapT = 10497;
      if (t.hasSampler) {
        const JValue& sm = samps.at(si);
        t.magFilter = sm.integer("magFilter", -1); t.minFilter = sm.integer("minFilter", -1);
        t.wrapS = sm.integer("wrapS", 10497); t.wrapT = sm.integer("wrapT", 10497);
      }
      g.textures.push_back(t);
    }
    auto used = [&](int tex) {
      if (tex < 0 || (size_t)tex >= g.textures.size()) return;
      int im = g.textures[tex].image;
      if (im >= 0 && (size_t)im < g.images.size() && g.images[im].rgba8.empty())
        raise(EID_ERR_UNSUPPORTED, "texture %d uses image %d, which is not a PNG (or is missing): decode it on the host and pass the texels "
                                   "with eid_scene_provide_image before eid_scene_load_gltf (the library decodes PNG only)", tex, im);
    };
    for (const auto& m : g.materials) { used(m.baseColorTexture); used(m.metallicRoughnessTexture); used(m.emissiveTexture); used(m.normalTexture); used(m.transmissionTexture); }
  }
  ImportCtx c{d, g, {}, {}, {}};
  int sceneIdx = d.root->integer("scene", 0);
  const JValue& scenes = d.top("scenes");
  if (scenes.size()) {
    if (sceneIdx < 0 || (size_t)sceneIdx >= scenes.size()) sceneIdx = 0;
    if (auto roots = scenes.at(sceneIdx).get("nodes"))
      for (size_t i = 0; i < roots->size(); ++i) processNode(c, roots->indexAt(i), identity(), 0);
  } else {
    // no `scenes`: the roots are the nodes that are nobody's child (a child is reached through its parent, with the parent's transform)
    const JValue& nodes = d.top("nodes");
    std::vector<char> isChild(nodes.size(), 0);
    for (size_t i = 0; i < nodes.size(); ++i)
      if (auto ch = nodes.at(i).get("children"))
        for (size_t k = 0; k < ch->size(); ++k) { const int ci = ch->indexAt(k); if ((size_t)ci < nodes.size()) isChild[(size_t)ci] = 1; }
    for (size_t i = 0; i < nodes.size(); ++i) if (!isChild[i]) processNode(c, (int)i, identity(), 0);
  }
  for (const auto& pm : g.primMeshes)
    if ((size_t)pm.materialIndex >= g.materials.size()) raise(EID_ERR_PARSE, "material index %d out of range", pm.materialIndex);
  g.computeDimensions();
  if (!c.cameras.empty()) {   // camera 0: eye = translation, center = eye + R*(0,0,-distance-to-scene-centre), up = +Y
    const auto& cn = c.cameras[0];
    g.hasCamera = true;
    double eye[3] = {cn.world.m[12], cn.world.m[13], cn.world.m[14]};
    double ctr[3] = {0.5 * (g.bboxMin[0] + g.bboxMax[0]), 0.5 * (g.bboxMin[1] + g.bboxMax[1]), 0.5 * (g.bboxMin[2] + g.bboxMax[2])};
    double dist = std::sqrt((ctr[0] - eye[0]) * (ctr[0] - eye[0]) + (ctr[1] - eye[1]) * (ctr[1] - eye[1]) + (ctr[2] - eye[2]) * (ctr[2] - eye[2]));
    if (!(dist > 0)) dist = 1.0;
    double zl = std::sqrt(cn.world.m[8] * cn.world.m[8] + cn.world.m[9] * cn.world.m[9] + cn.world.m[10] * cn.world.m[10]);
    if (!(zl > 0)) zl = 1.0;
    for (int k = 0; k < 3; ++k) {
      g.camEye[k] = (float)eye[k];
      g.camCenter[k] = (float)(eye[k] - cn.world.m[8 + k] / zl * dist);
    }
    g.camUp[0] = 0.f; g.camUp[1] = 1.f; g.camUp[2] = 0.f;
    g.camYfovRad = cn.yfov;
  }
}

}  // namespace eid

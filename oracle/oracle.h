/*
 * oracle/oracle.h — TEST INFRASTRUCTURE: CPU oracle of the reference's per-frame render loop.
 *
 * A plain C++17 restatement (no CUDA, no GPU) of IwakuraRein/CIS-565-Final-VR-Raytracer's
 * hot path: Scene table builders (src/scene.cpp), the alias table (src/alias_table.hpp),
 * Renderer::run's dispatch schedule (src/renderer.cpp:154-206) and the five live compute
 * shaders with all their includes (shaders/ *.glsl and *.comp).  Each function cites the
 * reference file:line it follows.
 *
 * PARITY PINNING.  The reference ships no tests, golden vectors or fixtures, and its application cannot be built or
 * run here (no Vulkan ICD, no glslang, nvpro_core un-vendored — SURVEY.md §8c).  But its SOURCE can be compiled where
 * it lies, and this oracle is pinned to it (oracle/Makefile target `ref` -> oracle/_ref/libref.so; outputs committed as
 * tests/golden/ref_vectors.npz, ref_post.npz, ref_trace.npz for machines without /root/reference):
 *   - shaders/host_device.h (struct sizes), shaders/compress.glsl C++ branch (oct codec, packUnorm4x8),
 *     src/alias_table.hpp (light alias tables), src/hdr_sampling.cpp (environment alias map, integral, average);
 *   - ALL FIVE STAGE SHADERS, main() included, with everything they include — direct_stage.comp, indirect_stage.comp,
 *     denoise_direct.comp, denoise_indirect.comp, compose.comp; globals, random, common, pathtrace,
 *     pbr_metallicworkflow, gltf_material, env_sampling, sun_and_sky, shade_state, reservoir, denoise_common,
 *     compress, tonemapping — transliterated token by token into compilable C++ (oracle/ref_shim/glsl_prep.py:
 *     parameter qualifiers, literal suffixes, swizzle calls, built-in names; the expressions are the reference's text)
 *     and dispatched over whole frames in 8x8 work groups like Renderer::run (ref_trace.cpp, ref_post.cpp).  Every
 *     buffer this oracle leaves after every frame — G-buffer, motion vectors, both reservoir buffers, pre-denoise and
 *     final images — is BIT-IDENTICAL to what the reference's text leaves (tests/test_oracle_kat.py), for point /
 *     triangle / HDR / sun & sky lighting, ReSTIR off / RIS / temporal, ragged sizes, textured materials, instanced /
 *     rotated / mirrored nodes; and so is the CUDA path (tests/test_gpu_parity.py).
 * What remains the numerical CONTRACT of DESIGN.md §3 rather than the reference's own arithmetic: the ray queries
 * (they run inside the Vulkan driver: hit acceptance, tie-break, candidate order), the rounding of the GLSL built-ins
 * (pow, exp, sin, normalize ...), the fixed-function samplers, and the un-vendored glTF import of nvpro_core.  The pin
 * scenes are opaque (HitTest, the stochastic-alpha callback, lives with the ray queries in traceray_rq.glsl).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (libeidola.so) never links or calls it.
 */
#pragma once
#include <cstdint>
#include <vector>
#include <atomic>
#include "host_device.h"
#include "eidola.h"      // eid_scene_desc & friends: the interchange structs only
#include "glsl_types.h"

namespace orc {

struct OTri {             // world-space triangle prepared for the intersector
  vec3 v0, e1, e2;
  int prim, inst;         // primitiveID inside the prim mesh, instanceID (node index)
  int customIndex;        // prim mesh index
  int cullDisable;        // VK_GEOMETRY_INSTANCE_TRIANGLE_FACING_CULL_DISABLE (accelstruct.cpp:148-149)
  int opaque;             // VK_GEOMETRY_INSTANCE_FORCE_OPAQUE (accelstruct.cpp:145-147)
  int flip;               // instance transform mirrors (det < 0): object-space facing is the opposite of world-space
};

struct Hit {
  float hitT; int primitiveID, instanceID, instanceCustomIndex; vec2 bary; int opaque;
};
// candidates are ordered by (t, instanceID, primitiveID); `after` = only candidates strictly greater than this key
struct HitKey { float t; int inst, prim; };

struct BvhNode { float lo[3], hi[3]; int left, right, first, count; };

struct Bvh {
  std::vector<BvhNode> nodes;
  std::vector<int> order;    // triangle indices
  void build(const std::vector<OTri>& tris, float pad);
};

// one entry of texturesMap[] (layouts.glsl:51): BGRA8/RGBA8 UNORM image + the sampler state the reference derives from glTF
struct Texture {
  int width = 1, height = 1;
  std::vector<uint8_t> rgba{255, 255, 255, 255};
  int linear = 1;           // magnification filter (all taps are textureLod(..., 0))
  int wrapS = 0, wrapT = 0; // 0 REPEAT, 1 MIRRORED_REPEAT, 2 CLAMP_TO_EDGE
  vec4 sample(vec2 uv) const;
};

struct Scene {
  std::vector<Texture> textures;
  // flat glTF-import data (nvh::GltfScene shape)
  std::vector<float> positions, normals, tangents, texcoords0, colors0;
  std::vector<uint32_t> indices;
  std::vector<eid_prim_mesh> primMeshes;
  std::vector<eid_node> nodes;
  std::vector<eid_material_desc> materials;
  std::vector<eid_light_desc> lights;
  // device tables of the reference
  std::vector<std::vector<VertexAttributes>> vertexBufs;
  std::vector<std::vector<uint32_t>> indexBufs;
  std::vector<int> instMaterial;                 // InstanceData.materialIndex
  std::vector<GltfShadeMaterial> shadeMaterials;
  std::vector<PuncLight> puncLights;
  std::vector<TrigLight> trigLights;
  LightBufInfo lightBufInfo{};
  float trigLightWeight = 0.f, puncLightWeight = 0.f;
  // camera
  SceneCamera camera{};
  float eye[3] = {2, 2, -5}, center[3] = {-1, 2, -1}, up[3] = {0, 1, 0};   // main.cpp:68
  float fovDeg = 60.f;
  float staticEye[3] = {0, 0, 0};   // function-static `eye` of Scene::updateCamera (scene.cpp:780)
  // acceleration structure inputs
  std::vector<mat4x3> objectToWorld, worldToObject;
  std::vector<OTri> tris;
  Bvh bvh;
  bool useBvh = true;
  float bboxMin[3], bboxMax[3];

  void load(const eid_scene_desc& d);
  void updateCamera(uint32_t w, uint32_t h);
  void buildAccel();
  Hit closestHit(vec3 o, vec3 d, float tmax, std::atomic<uint64_t>* ctr, const HitKey* after = nullptr) const;
  bool hasNonOpaque = false;
  bool anyHit(vec3 o, vec3 d, float tmax, std::atomic<uint64_t>* ctr) const;
};

// HdrSampling (src/hdr_sampling.{hpp,cpp}): RGBA32F lat-long environment + per-texel alias map
struct Environment {
  uint32_t width = 0, height = 0;
  std::vector<float> pixels;              // rgba
  std::vector<ImptSampData> accel;
  float integral = 1.f, average = 1.f;
  void create(const float* rgba, uint32_t w, uint32_t h);
  vec3 texture(vec2 uv) const;            // sampler: LINEAR, REPEAT in u, CLAMP_TO_EDGE in v (hdr_sampling.cpp:67-75)
};

struct Renderer {
  const Scene* scene = nullptr;
  const Environment* env = nullptr;      // null: constant environment `envConstant`
  uint32_t width = 0, height = 0;      // allocation size
  std::vector<uvec4> gbuffer[2];
  std::vector<int16_t> motion;         // 2 per pixel
  std::vector<DirectReservoir> directResv[2];
  std::vector<DirectReservoir> tempDirectResv;          // spatial reuse scratch (renderer.hpp: m_tempDirectResv); persists across frames
  std::vector<IndirectReservoir> indirectResv[2];
  std::vector<vec4> directResult, indirectResult;       // thisDirectResultImage / thisIndirectResultImage
  std::vector<vec4> denoiseTemp[4];                     // DirTempA, DirTempB, IndTempA, IndTempB
  vec3 envConstant;
  SunAndSky sunSky{};                  // _sunAndSky uniform (layouts.glsl:53); in_use == 1 replaces the HDR map
  int lastSet = 0;
  int variant = 0;                     // EID_VARIANT_* bits: the reference's compile-time switches (host_device.h:27-29, indirect_stage.comp:35)
  std::atomic<uint64_t> closestRays{0}, anyRays{0}, primaryHits{0};
  double kernelMs[5] = {0, 0, 0, 0, 0};

  void create(const Scene* s, uint32_t w, uint32_t h);
  void run(const RtxState& st, int frames);
  // stage-wise entry points (used by the band-sharded tests)
  void runDirect(const RtxState& st, int frames, int y0, int y1);
  void runIndirect(const RtxState& st, int frames, int y0, int y1);
  void runPost(const RtxState& st, int frames);
};

// function taps (oracle_shaders.cpp): which = 0 toConcentricDisk, 1 powerHeuristic, 2 GetSphericalUv, 3 CreateCoordinateSystem, 4 HDRToLDR,
// 5 LDRToHDR, 6 metallicWorkflowBSDF, 7 metallicWorkflowPdf, 8 metallicWorkflowSample, 9 DirectReservoir update/merge/validity/clamp,
// 10 IndirectReservoir update/validity/clamp, 11 toneMap, 12 OffsetRay, 13 tea, 14 rand x2
int fn_arity(int which, int* nin, int* nout);
int fn(int which, const float* in, int n, float* out);
int ctx_fn(Renderer& rr, const RtxState& st, int which, const float* in, int n, float* out);   // scene-dependent taps
vec3 post_toneMap(vec3 color, float exposure);   // tonemapping.glsl:78-95 (oracle_post.cpp)
// shaders/post.frag main for one pixel (oracle_post.cpp)
vec4 post_frag(const Tonemapper& tm, int debugging_mode, vec4 direct, vec4 indirect, int px, int py, int width, int height, vec4 avgDirect = vec4(), vec4 avgIndirect = vec4());
vec4 mip_chain_average(const vec4* img, int width, int height, int pitch);   // 1x1 level of RenderOutput::genMipmap's chain (oracle_post.cpp)
vec3 post_toneExposure(const Tonemapper& tm, vec3 RGB, float logAvgLum);      // post.frag:65-70
// shaders/sun_and_sky.glsl:453-601 (oracle_sunsky.cpp)
vec3 sun_and_sky(const SunAndSky& ss, vec3 in_direction);

// alias table (src/alias_table.hpp:21-63)
void discreteSampler1D(std::vector<float> values, std::vector<float>& prob, std::vector<int>& failId);

// free functions exposed for known-answer tests
uint tea(uint val0, uint val1);
uint pcg(uint& state);
float rnd(uint& seed);
uint hash8bit(uint a);
uint compress_unit_vec(vec3 nv);
vec3 decompress_unit_vec(uint packed);
vec3 OffsetRay(vec3 p, vec3 n);

}  // namespace orc

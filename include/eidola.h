/*
 * eidola.h — C-ABI of libeidola.so, the B200-native drop-in for the reference's per-frame
 * render loop (Renderer::run and the acceleration structure it traverses).
 *
 * The reference has no FFI layer; its seam is the C++ class API that SampleExample calls
 * (SURVEY.md §8b).  Every export below names the reference method it replaces (file:line are
 * relative to the reference repository).  Conventions:
 *   - plain pointers and sizes only, no C++/torch types; handles are opaque;
 *   - return 0 (EID_OK) on success, a negative eid_status otherwise; eid_last_error() returns a
 *     thread-local message.  No exceptions cross the boundary.  (The reference ignores VkResults
 *     and uses bool/assert: scene.cpp:164-169, renderer.cpp:48.)
 *   - host inputs are borrowed for the duration of the call; the library owns all device memory;
 *   - a handle is used by one thread at a time (reference: loader thread + render thread with the
 *     m_busy flag, sample_example.cpp:120-156); eid_renderer_run is asynchronous on the
 *     renderer's CUDA stream, like command-buffer recording.
 *   - there is NO CPU fallback: every compute entry point fails with EID_ERR_CUDA when no
 *     sm_100-class device is usable.
 */
#ifndef EIDOLA_H
#define EIDOLA_H

#include <stddef.h>
#include <stdint.h>
#include "host_device.h"

#ifdef __cplusplus
extern "C" {
#endif

#define EID_API __attribute__((visibility("default")))

typedef enum eid_status {
  EID_OK = 0,
  EID_ERR_INVALID = -1,   /* bad argument / null handle */
  EID_ERR_IO = -2,        /* file not found / unreadable */
  EID_ERR_PARSE = -3,     /* malformed glTF / JSON */
  EID_ERR_UNSUPPORTED = -4,
  EID_ERR_CUDA = -5,      /* CUDA runtime failure, no usable GPU */
  EID_ERR_STATE = -6      /* call order violated (e.g. run before accel build) */
} eid_status;

typedef struct eid_scene eid_scene;
typedef struct eid_accel eid_accel;
typedef struct eid_renderer eid_renderer;
typedef struct eid_env eid_env;

EID_API const char* eid_last_error(void);
EID_API int eid_version(void);
/* number of CUDA devices visible (0 when none); never fails */
EID_API int eid_device_count(void);

/* ---------------------------------------------------------------------------------------------
 * Flat scene description == what nvh::GltfScene holds after importMaterials/importDrawableNodes
 * (reference scene.cpp:72-74; members m_positions/m_normals/m_tangents/m_texcoords0/m_colors0/
 * m_indices/m_primMeshes/m_nodes/m_materials/m_lights/m_cameras).
 * ------------------------------------------------------------------------------------------- */
typedef struct eid_prim_mesh {       /* nvh::GltfPrimMesh */
  uint32_t firstIndex, indexCount, vertexOffset, vertexCount;
  int32_t  materialIndex;
} eid_prim_mesh;

typedef struct eid_node {            /* nvh::GltfNode */
  float   worldMatrix[16];           /* column-major */
  int32_t primMesh;
} eid_node;

typedef struct eid_material_desc {   /* nvh::GltfMaterial, the fields scene.cpp:415-448 reads */
  float   baseColorFactor[4];
  int32_t baseColorTexture;
  float   metallicFactor, roughnessFactor;
  int32_t metallicRoughnessTexture;
  int32_t emissiveTexture;
  float   emissiveFactor[3];
  int32_t alphaMode;                 /* 0 OPAQUE, 1 MASK, 2 BLEND */
  float   alphaCutoff;
  int32_t doubleSided;
  int32_t normalTexture;
  float   normalTextureScale;
  float   transmissionFactor;
  int32_t transmissionTexture;
  float   ior;
} eid_material_desc;

typedef struct eid_light_desc {      /* nvh::GltfLight (KHR_lights_punctual) */
  float   worldMatrix[16];
  int32_t type;                      /* LightType_* */
  float   color[3];
  float   intensity;
  float   range;
  float   innerConeAngle, outerConeAngle;
} eid_light_desc;

typedef struct eid_image_desc {      /* tinygltf::Image after decoding: 4 bytes per texel, row 0 first */
  uint32_t       width, height;
  const uint8_t* rgba8;              /* NULL / 0x0 = "image not present" -> 1x1 white (scene.cpp:575-582) */
} eid_image_desc;

typedef struct eid_texture_desc {    /* tinygltf::Texture + its Sampler (glTF enums; -1 = no sampler object) */
  int32_t image;                     /* index into images[]; out of range -> 1x1 white default texture (scene.cpp:604-610) */
  int32_t hasSampler;                /* 0: Vulkan defaults LINEAR / REPEAT (scene.cpp:613-616) */
  int32_t magFilter, minFilter;      /* 9728 NEAREST, 9729 LINEAR, 9984..9987 mip variants; anything else -> NEAREST (scene.cpp:519-525) */
  int32_t wrapS, wrapT;              /* 10497 REPEAT, 33648 MIRRORED_REPEAT, 33071 CLAMP_TO_EDGE */
} eid_texture_desc;

typedef struct eid_scene_desc {
  const float*    positions;   /* 3 floats / vertex */
  const float*    normals;     /* 3 */
  const float*    tangents;    /* 4 (w = handedness) */
  const float*    texcoords0;  /* 2 */
  const float*    colors0;     /* 4 */
  uint32_t        vertexCount;
  const uint32_t* indices;
  uint32_t        indexCount;
  const eid_prim_mesh*     primMeshes;  uint32_t primMeshCount;
  const eid_node*          nodes;       uint32_t nodeCount;
  const eid_material_desc* materials;   uint32_t materialCount;
  const eid_light_desc*    lights;      uint32_t lightCount;
  int32_t hasCamera;           /* glTF camera 0, scene.cpp:298-308 */
  float   camEye[3], camCenter[3], camUp[3];
  float   camYfovRad;
  const eid_image_desc*   images;    uint32_t imageCount;     /* Scene::createTextureImages (scene.cpp:554-646) */
  const eid_texture_desc* textures;  uint32_t textureCount;
} eid_scene_desc;

typedef struct eid_scene_info {
  uint32_t primMeshCount, nodeCount, materialCount, puncLightCount, trigLightCount;
  uint32_t vertexCount, indexCount;
  uint64_t triangleInstances;   /* sum over nodes of indexCount/3 == triangles the BVH holds */
  float    trigLightWeight;     /* Scene::m_trigLightWeight (scene.hpp:82-83) */
  float    puncLightWeight;     /* Scene::m_puncLightWeight */
  float    bboxMin[3], bboxMax[3];
} eid_scene_info;

typedef enum eid_scene_table {
  EID_TABLE_MATERIALS = 0,     /* GltfShadeMaterial[materialCount]            scene.cpp:415-448 */
  EID_TABLE_PUNC_LIGHTS = 1,   /* PuncLight[max(1,n)]                         scene.cpp:319-353 */
  EID_TABLE_TRIG_LIGHTS = 2,   /* TrigLight[max(1,n)]                         scene.cpp:355-409 */
  EID_TABLE_LIGHT_INFO = 3,    /* LightBufInfo                                scene.cpp:101-105 */
  EID_TABLE_INSTANCE_DATA = 4, /* InstanceData[primMeshCount]                 scene.cpp:179-195 */
  EID_TABLE_VERTICES = 5,      /* VertexAttributes[] of prim mesh `index`     scene.cpp:209-289 */
  EID_TABLE_INDICES = 6,       /* uint32[] of prim mesh `index`                                  */
  EID_TABLE_CAMERA = 7,        /* SceneCamera                                 scene.cpp:777-826 */
  EID_TABLE_TEXELS = 8         /* RGBA8 texels (uint32 each, row-major) of texture `index` = texturesMap[index], scene.cpp:554-646 */
} eid_scene_table;

/* Scene::setup + constructor (scene.hpp:60). device = CUDA ordinal, or EID_DEVICE_NONE for a host-only
 * scene: glTF import + table building + camera maths work (and can be read back), nothing is uploaded,
 * and eid_accel_build / eid_renderer_create on it fail with EID_ERR_CUDA.  This is NOT a CPU render
 * path — it exists so the host-side logic can be unit-tested on machines without a GPU. */
#define EID_DEVICE_NONE (-1)
EID_API int  eid_scene_create(eid_scene** out, int device);
/* Scene::load(filename) (scene.cpp:57-125): glTF 2.0 (.gltf + external .bin / data: URIs). */
EID_API int  eid_scene_load_gltf(eid_scene* s, const char* path);
/* Hands the loader image `imageIndex` of the next eid_scene_load_gltf already decoded (RGBA8, row 0 first).  The reference decodes
 * PNG/JPEG with FreeImage/stb (scene.cpp:152,159); this build has no image decoder, so the host decodes and provides. */
EID_API int  eid_scene_provide_image(eid_scene* s, uint32_t imageIndex, const uint8_t* rgba8, uint32_t width, uint32_t height);
/* Same import, from arrays already in the GltfScene shape (harness-generated scenes). */
EID_API int  eid_scene_load_desc(eid_scene* s, const eid_scene_desc* desc);
/* Scene::destroy (scene.cpp:453-511) */
EID_API void eid_scene_destroy(eid_scene* s);
/* CameraManip.setCamera({eye, center, up, fov}) (scene.cpp:298-308, main.cpp:67-68) */
EID_API int  eid_scene_set_lookat(eid_scene* s, const float eye[3], const float center[3], const float up[3], float fovDeg);
/* Scene::updateCamera(cmdBuf, size) (scene.cpp:777-826): rolls last* fields, applies the constant
 * sub-pixel shift, uploads the UBO. */
EID_API int  eid_scene_update_camera(eid_scene* s, uint32_t width, uint32_t height);
/* replay path: install a complete SceneCamera verbatim */
EID_API int  eid_scene_set_camera(eid_scene* s, const SceneCamera* cam);
EID_API int  eid_scene_get_camera(eid_scene* s, SceneCamera* out);
/* Scene::getStat / getScene / m_trigLightWeight / m_puncLightWeight (scene.hpp:74-83) */
EID_API int  eid_scene_get_info(eid_scene* s, eid_scene_info* out);
/* size in bytes of a table (index only used for VERTICES/INDICES) */
EID_API int64_t eid_scene_table_bytes(eid_scene* s, int table, uint32_t index);
/* device -> host copy of one table (debug / parity tap) */
EID_API int  eid_scene_read_table(eid_scene* s, int table, uint32_t index, void* dst, size_t bytes);

/* ---------------------------------------------------------------------------------------------
 * AccelStructure (accelstruct.hpp:40-46; accelstruct.cpp:55-162).  The Vulkan driver BVH is
 * replaced by a 4-wide BVH (binned-SAH or Morton topology; one flat world-space tree, or BLAS per
 * prim mesh + TLAS) built by csrc/accel.cu and walked in software (csrc/trace.cuh).
 * ------------------------------------------------------------------------------------------- */
typedef struct eid_accel_info {
  uint64_t triangleCount;
  uint32_t nodeCount;          /* wide nodes */
  uint32_t maxDepth;
  uint64_t nodeBytes, triBytes;
  float    buildMs;
  int32_t  twoLevel;           /* 1: BLAS per prim mesh + TLAS over the instances (triangleCount = unique triangles) */
  uint32_t blasCount, tlasNodeCount, instanceCount;
  int32_t  fastTrace;          /* 1: tree topology from the binned-SAH builder (EID_ACCEL_FAST_TRACE, the default), 0: Morton / LBVH build on the GPU */
} eid_accel_info;

typedef struct eid_hit {       /* PtPayload subset (globals.glsl:48-58) */
  float   hitT;                /* 1e28 on miss */
  int32_t primitiveID;
  int32_t instanceID;          /* node index */
  int32_t instanceCustomIndex; /* prim mesh index */
  float   baryU, baryV;
} eid_hit;

/* AccelStructure::create(gltfScene, vertexBufs, indexBufs) (accelstruct.cpp:55-162).
 * EID_ACCEL_FLAT: every instance's triangles baked into ONE world-space BVH (fastest walk; memory x instance count).
 * EID_ACCEL_TWO_LEVEL: the reference's structure — one bottom-level tree per prim mesh over object-space triangles, shared by all its
 * instances, + a top-level tree over the instances; hits are bit-identical to the flat form (the triangle test runs on the same
 * world-space triangle, computed from the instance matrix at test time).  EID_ACCEL_AUTO (eid_accel_build): two-level once the flat list
 * would hold at least twice the unique triangles. */
enum { EID_ACCEL_AUTO = 0, EID_ACCEL_FLAT = 1, EID_ACCEL_TWO_LEVEL = 2 };
/* Build quality, or-ed into `mode` (the reference builds every BLAS and the TLAS with VK_BUILD_ACCELERATION_STRUCTURE_PREFER_FAST_TRACE_BIT_KHR,
 * accelstruct.cpp:125-126,161).  EID_ACCEL_FAST_TRACE (the default when neither is given): binned-SAH topology built by the host's threads
 * (~0.3 s per million triangles), refit + 4-wide collapse on the GPU; fewer node visits per ray.  EID_ACCEL_FAST_BUILD: Morton codes + radix
 * sort + Karras tree, all on the GPU (~10 ms per million triangles).  Results never depend on the choice (closest hit is a total order).
 * The environment variable EIDOLA_ACCEL_BUILD=lbvh|sah overrides the default (not an explicit flag). */
enum { EID_ACCEL_FAST_TRACE = 0x100, EID_ACCEL_FAST_BUILD = 0x200 };
EID_API int  eid_accel_build(eid_scene* s, eid_accel** out);
EID_API int  eid_accel_build_ex(eid_scene* s, int mode, eid_accel** out);
EID_API void eid_accel_destroy(eid_accel* a);
EID_API int  eid_accel_get_info(eid_accel* a, eid_accel_info* out);
/* Batch ray query with HOST buffers (ClosestHit / AnyHit of traceray_rq.glsl:108-185).
 * rays: 8 floats each {ox,oy,oz,tmax, dx,dy,dz,unused}.  any_hit != 0 -> hits[i].hitT is 0 when
 * occluded and 1e28 when free, other fields undefined. */
EID_API int  eid_accel_trace(eid_accel* a, const float* rays, uint32_t n, int any_hit, eid_hit* hits);
/* Tap of the EID_ACCEL_FAST_TRACE topology builder (host only, needs no device; test infrastructure like the eid_*_tap functions below):
 * the binary tree over n boxes (lo / hi: n x 3 floats) exactly as the build hands it to the GPU refit + collapse.  order[n]: position ->
 * source box; per inner node (n - 1 of them, root = 0) left / right (>= 0 inner node, < 0 leaf ~position), parentInner, rangeFirst,
 * rangeLast; parentLeaf[n].  threads <= 0: all hardware threads.  The result does not depend on `threads`. */
EID_API int  eid_accel_sah_tap(const float* lo, const float* hi, uint32_t n, int threads, uint32_t* order, int32_t* left, int32_t* right,
                               int32_t* parentInner, int32_t* parentLeaf, int32_t* rangeFirst, int32_t* rangeLast);

/* ---------------------------------------------------------------------------------------------
 * Renderer (renderer.hpp:52-61; renderer.cpp:97-375)
 * ------------------------------------------------------------------------------------------- */
typedef enum eid_buffer {
  EID_BUF_THIS_GBUFFER = 0,        /* uint32 x4 / pixel  (allocation pitch)               */
  EID_BUF_LAST_GBUFFER = 1,
  EID_BUF_MOTION = 2,              /* int16 x2 / pixel                                     */
  EID_BUF_THIS_DIRECT_RESV = 3,    /* DirectReservoir[size.x*size.y]                       */
  EID_BUF_LAST_DIRECT_RESV = 4,
  EID_BUF_THIS_INDIRECT_RESV = 5,  /* IndirectReservoir[(size.x/2)*(size.y/2)]             */
  EID_BUF_LAST_INDIRECT_RESV = 6,
  EID_BUF_DIRECT = 7,              /* thisDirectResultImage   float x4 / pixel             */
  EID_BUF_INDIRECT = 8,            /* thisIndirectResultImage float x4 / pixel             */
  EID_BUF_DENOISE_DIR_A = 9, EID_BUF_DENOISE_DIR_B = 10,
  EID_BUF_DENOISE_IND_A = 11, EID_BUF_DENOISE_IND_B = 12,
  EID_BUF_DISPLAY_F32 = 13,        /* output of eid_renderer_run_output: float x4 / pixel    */
  EID_BUF_DISPLAY_RGBA8 = 14,      /* ... and its UNORM8 packing, 4 bytes / pixel            */
  EID_BUF_TEMP_DIRECT_RESV = 15    /* tempDirectResv (spatial reuse scratch, one buffer for both sets, renderer.cpp:235);
                                      exists once a frame has run with eSpatial / eSpatiotemporal */
} eid_buffer;

/* kernels of one Renderer::run, in launch order */
enum { EID_K_DIRECT = 0, EID_K_INDIRECT = 1, EID_K_DENOISE_DIRECT = 2, EID_K_DENOISE_INDIRECT = 3,
       EID_K_COMPOSE = 4, EID_K_COUNT = 5 };

typedef struct eid_frame_stats {
  uint64_t closestHitRays;     /* rays issued by ClosestHit in the last frame (counted on device) */
  uint64_t anyHitRays;         /* rays issued by AnyHit */
  uint64_t primaryHits;        /* pixels whose primary ray hit geometry */
  uint32_t launches;           /* CUDA kernels launched by the last eid_renderer_run */
  float    kernelMs[EID_K_COUNT];   /* CUDA-event time per stage (sum over its passes); needs profiling on */
  uint32_t kernelLaunches[EID_K_COUNT];
  uint64_t nodeVisits;         /* BVH inner nodes fetched / triangles tested in the last frame; only counted */
  uint64_t triangleTests;      /*   at profiling level 2 (instrumented kernel variants)                        */
  uint64_t totalClosestHitRays;/* rays issued since the renderer was created (never reset) */
  uint64_t totalAnyHitRays;
  float    exchangeMs;             /* profiling: gap between the end of run_trace and the start of run_post* (multi-GPU exchange 1) */
  uint64_t maxNodeVisitsPerThread; /* profiling level 2: most inner-node visits any single thread (pixel) needed so far */
  uint64_t maxNodeVisitsPerQueuedRay[2]; /* profiling level 2, wavefront indirect stage: longest closest-hit / any-hit ray of the queues so far */
} eid_frame_stats;

/* Renderer::setup + create(size, layouts, scene) (renderer.cpp:50-57, 97-148).
 * cuda_stream: a cudaStream_t to enqueue on, or NULL for a renderer-owned stream.
 * All history buffers are zero-initialised. */
EID_API int  eid_renderer_create(eid_renderer** out, eid_scene* s, eid_accel* a,
                                 uint32_t width, uint32_t height, void* cuda_stream);
/* Renderer::update(size) (renderer.cpp:209-225): re-allocates, drops history */
EID_API int  eid_renderer_resize(eid_renderer* r, uint32_t width, uint32_t height);
EID_API void eid_renderer_destroy(eid_renderer* r);
/* ---- HdrSampling (hdr_sampling.hpp:43-48; hdr_sampling.cpp:55-242) -------------------------------------------------------
 * loadEnvironment(file) = eid_env_load_hdr (Radiance .hdr / RGBE; replaces stbi_loadf) or eid_env_create from RGBA32F texels
 * already in memory (row 0 = +Y pole).  Builds the per-texel alias table exactly like createEnvironmentAccel/buildAliasmap.
 * getIntegral()/getAverage() feed RtxState.fireflyClampThreshold / envMapLuminIntegInv (sample_example.cpp:104-105).
 * device = CUDA ordinal or EID_DEVICE_NONE (host tables only, for tests). */
EID_API int   eid_env_create(eid_env** out, int device, const float* rgba, uint32_t width, uint32_t height);
EID_API int   eid_env_load_hdr(eid_env** out, int device, const char* path);
EID_API void  eid_env_destroy(eid_env* e);
EID_API float eid_env_integral(eid_env* e);
EID_API float eid_env_average(eid_env* e);
EID_API int   eid_env_get_size(eid_env* e, uint32_t* width, uint32_t* height);
/* what = 0: ImptSampData[w*h] alias table, 1: RGBA32F texels (host copies; parity taps) */
EID_API int   eid_env_read(eid_env* e, int what, void* dst, size_t bytes);
/* install / remove (NULL) the HDR environment: EnvRadiance, EnvEval and EnvSample (pathtrace.glsl:40-72, env_sampling.glsl) then
 * use it; RtxState.environmentProb > 0 requires one */
EID_API int   eid_renderer_set_env(eid_renderer* r, eid_env* e);
/* the `_sunAndSky` uniform (layouts.glsl:53; SampleExample::m_sunAndSky, updated every frame at sample_example.cpp:172): with
 * in_use == 1 the procedural sun & sky of shaders/sun_and_sky.glsl replaces the HDR map in EnvRadiance, EnvEval and EnvSample
 * (pathtrace.glsl:40-72, env_sampling.glsl:111-125).  The struct is copied; default in_use = 0. */
EID_API int   eid_renderer_set_sun_and_sky(eid_renderer* r, const SunAndSky* ss);
/* parity tap of the device-side shader functions, n items of a fixed number of floats each (which: 0 toConcentricDisk, 1 powerHeuristic,
 * 2 GetSphericalUv, 3 CreateCoordinateSystem, 4 HDRToLDR, 5 LDRToHDR, 6 metallicWorkflowBSDF, 7 metallicWorkflowPdf,
 * 8 metallicWorkflowSample, 11 toneMap, 12 OffsetRay, 13 tea, 14 rand x2; layouts in tests/ref_fn_inputs.py): compared bit for bit
 * with the reference's own GLSL text compiled as C++ (oracle/ref_shim) */
EID_API int   eid_fn_tap(int device, int which, const float* in, uint32_t n, float* out);
/* the scene-dependent ones, evaluated with the renderer's scene, camera, environment and the given RtxState (which: 0
 * SampleDirectLightNoVisibility, 2 EnvEval, 3 EnvRadiance, 4 raySpawn, 5 clampRadiance, 6 Sample) */
EID_API int   eid_renderer_fn_tap(eid_renderer* r, const RtxState* state, int which, const float* in, uint32_t n, float* out);
/* parity tap: sun_and_sky(ss, dir) (sun_and_sky.glsl:453-601) for n host directions (3 floats each) -> n RGB triples, on `device` */
EID_API int   eid_sun_and_sky_eval(int device, const SunAndSky* ss, const float* dirs, uint32_t n, float* rgb);
/* constant environment radiance used by EnvRadiance/EnvEval (pathtrace.glsl:40-72) until an HDR
 * map is installed; default (0,0,0). */
EID_API int  eid_renderer_set_env_constant(eid_renderer* r, const float rgb[3]);
/* Numerics of the denoiser's edge-stopping exponentials.  0 (default): hardware ex2 (MUFU), images within ~1e-6 relative of
 * the bit-reproducible path.  1: the deterministic polynomial exp shared with the CPU oracle — every buffer of a frame is then
 * bit-identical to the oracle's (used by the parity tests).  G-buffer, motion indices and reservoirs never depend on this. */
EID_API int  eid_renderer_set_strict_math(eid_renderer* r, int enabled);
/* A-Trous kernel shape: every thread filters `rowsPerThread` pixels of one column that are 2^level rows apart and shares their
 * overlapping tap rows: 1 = 25 loads per pixel, 2 (default, measured fastest on B200) = 15, 4 = 10 but 110 registers.
 * Results are bit-identical for every setting. */
EID_API int  eid_renderer_set_denoise_rows(eid_renderer* r, int rowsPerThread);
/* Form of the A-Trous passes (denoise_direct.comp / denoise_indirect.comp).  mode 1 (default): shared-memory tile kernel — a block owns a
 * 32 x (4 rowsPerThread) tile of ONE phase of the level's dilated lattice, its three input planes + halo arrive by TMA
 * (cp.async.bulk.tensor.5d through a lattice-view tensor map) and every tap is an LDS at an immediate offset; mode 2: the same kernel with
 * the tiles loaded by cp.async (A/B measurement); mode 0: the round-1 kernel (taps served by L1, eid_renderer_set_denoise_rows applies).
 * rowsPerThread: 2 or 4 (0 keeps the current value).  With strict math every mode is bit-identical to the oracle. */
EID_API int  eid_renderer_set_denoise_tiles(eid_renderer* r, int mode, int rowsPerThread);
/* RenderOutput::run (render_output.cpp:224-240) -> shaders/post.frag as a compute pass over the frame rendered last: direct +
 * indirect (or the debug view selected by that frame's RtxState.debugging_mode), Uncharted-2 tonemap, dither, contrast /
 * brightness / saturation / vignette, evaluated 1:1 (one output pixel per rendered pixel, zoom 1, renderingRatio (1,1) unless set
 * in `tm`).  Enqueued on the renderer's stream; results in EID_BUF_DISPLAY_F32 / EID_BUF_DISPLAY_RGBA8.
 * tm->autoExposure bit 0 (the GUI's "Auto Exposure", sample_gui.cpp:238-259): RenderOutput::genMipmap (render_output.cpp:243-253) runs first —
 * the mip chain of both result images down to 1 x 1 — and post.frag's toneExposure uses that average; bit 1 (toneLocalExposure) is
 * EID_ERR_UNSUPPORTED. */
EID_API int  eid_renderer_run_output(eid_renderer* r, const Tonemapper* tm);
/* Form of indirect_stage (K2).  enabled = 1 (default): wavefront — the stage is cut at its ray queries into ray queues that a
 * persistent dynamic-fetch traversal kernel drains (every lane takes the next queued ray when its own ends); used whenever the
 * scene has no stochastic-alpha instance and maxDepth <= 25, otherwise (and with enabled = 0) the one-thread-per-pixel kernel
 * runs.  enabled = 2: wavefront with the shadow-ray queues on the main stream too (default: a second stream, beside the deeper
 * bounces).  All forms produce bit-identical buffers.  traceBlocks = grid of the traversal kernel in 128-thread blocks (0 = default). */
EID_API int  eid_renderer_set_wavefront(eid_renderer* r, int enabled, int traceBlocks);
/* Renderer::run(cmdBuf, state, profiler, descSets, frames) (renderer.cpp:154-206): enqueues
 * direct_stage, indirect_stage, denoise_direct x4, denoise_indirect x5, compose and returns.
 * `frames` selects the ping-pong set exactly as (frames+1)%2 (renderer.cpp:157). */
EID_API int  eid_renderer_run(eid_renderer* r, const RtxState* state, int frames);
EID_API int  eid_renderer_sync(eid_renderer* r);
/* device pointers of thisDirectResultImage / thisIndirectResultImage, valid until the next run */
EID_API int  eid_renderer_get_outputs(eid_renderer* r, const float** direct_rgba32f, const float** indirect_rgba32f);
/* size in bytes of a buffer for the size used by the last run (or the allocation when none) */
EID_API int64_t eid_renderer_buffer_bytes(eid_renderer* r, int which);
/* synchronising device -> host copy (parity tap / checkpoint of the cross-frame state) */
EID_API int  eid_renderer_read(eid_renderer* r, int which, void* host_dst, size_t bytes);
/* host -> device restore of a history buffer (checkpoint/resume, tests) */
EID_API int  eid_renderer_write(eid_renderer* r, int which, const void* host_src, size_t bytes);
/* End-to-end convenience used by hosts that keep everything in host memory: uploads `cam`
 * (may be NULL to keep the scene's camera), runs one frame, copies the two result images into
 * pinned or pageable host buffers (width*height*16 bytes each, either may be NULL) and syncs. */
EID_API int  eid_renderer_render_host(eid_renderer* r, const SceneCamera* cam, const RtxState* state, int frames,
                                      float* direct_host, float* indirect_host);
/* Pipelined form of eid_renderer_render_host: returns after enqueueing; the device->host copies of frame f run on a copy stream
 * while frame f+1 renders.  The host buffers are valid after eid_renderer_wait_host.  Alternate two host buffer pairs. */
EID_API int  eid_renderer_render_host_async(eid_renderer* r, const SceneCamera* cam, const RtxState* state, int frames,
                                            float* direct_host, float* indirect_host);
EID_API int  eid_renderer_wait_host(eid_renderer* r);
/* The reference's compile-time shader variants as run-time switches (all 0 = the reference as shipped):
 *   EID_VARIANT_DIRECT_BILATERAL    #define DENOISER_DIRECT_BILATERAL 1 (host_device.h:28): direct_stage writes denoiseDirTempA (direct_stage.comp:284-288),
 *                                   ONE 9 x 9 cross-bilateral pass with a spatial term replaces the four A-Trous levels (denoise_direct.comp:73-137, renderer.cpp:186-188)
 *   EID_VARIANT_INDIRECT_BILATERAL  #define DENOISER_INDIRECT_BILATERAL 1 (host_device.h:29): ONE 11 x 11 pass at quarter resolution (denoise_indirect.comp:77-130)
 *   EID_VARIANT_FETCH_4_SUBPIXELS   #define FETCH_GEOM_CHECK_4_SUBPIXELS 1 (indirect_stage.comp:35): the quarter-res stage averages the four G-buffer texels of
 *                                   its 2 x 2 footprint and draws the material id among them (pathtrace.glsl:314-358; one more RNG draw per pixel)
 *   EID_VARIANT_DIRECT_SPLIT        the two-kernel form of the direct stage, direct_gen.comp (:77-149: primary ray, G-buffer — sky / emitter /
 *                                   debug radiance parked in its albedo bits —, RIS candidates, shadow ray, reservoir) followed by direct_reuse.comp
 *                                   (:102-153: state rebuilt from the G-buffer, temporal merge, clamp, `direct = Li` of the pre-merge sample), in place
 *                                   of direct_stage.comp.  The reference builds both pipelines (renderer.cpp:129-132) and never dispatches them; the
 *                                   pair is reproduced as written (it is visibly work in progress there).  No spatial reuse in this form.
 * (INDIRECT_PRE_UPSCALE, host_device.h:27, is defined but referenced nowhere in the reference: there is nothing to switch.) */
enum { EID_VARIANT_DIRECT_BILATERAL = 1, EID_VARIANT_INDIRECT_BILATERAL = 2, EID_VARIANT_FETCH_4_SUBPIXELS = 4, EID_VARIANT_DIRECT_SPLIT = 8 };
EID_API int  eid_renderer_set_variant(eid_renderer* r, int flags);
/* 1 (default): run the direct denoiser (K3) on a second CUDA stream concurrently with indirect_stage (K2) + the indirect
 * denoiser (K4) — the reference's true dependencies are K1->{K2,K3}, K2->K4, {K3,K4}->K5.  0: strict K1..K5 order on one stream.
 * Results are identical either way; per-stage times (kernelMs) overlap when enabled. */
EID_API int  eid_renderer_set_overlap(eid_renderer* r, int enabled);
/* Frames in flight.  1 (default): a frame's stages only start when the previous eid_renderer_run has finished (the reference records one
 * command buffer per frame).  2: direct_stage of frame f + 1 runs on an internal stream while indirect_stage / denoise / compose of frame f
 * are still on the renderer's stream — the direct stage only needs the G-buffer and the direct reservoirs of frame f, both complete after
 * ITS direct stage; what the later stages of frame f read is double-buffered per ping-pong parity.  Results are bit-identical to mode 1;
 * the host must call eid_renderer_run with consecutive `frames` values and not modify the cross-frame buffers in between.  Work the host
 * enqueued on the renderer's stream before a run is NOT waited for by that frame's direct stage in mode 2. */
EID_API int  eid_renderer_set_pipeline(eid_renderer* r, int framesInFlight);
/* 0: off (default; ray counters are always kept), 1: per-stage CUDA-event timing,
 * 2: additionally run the instrumented trace kernels that count BVH node visits / triangle tests */
EID_API int  eid_renderer_set_profiling(eid_renderer* r, int enabled);
EID_API int  eid_renderer_get_stats(eid_renderer* r, eid_frame_stats* out);

/* ---- multi-GPU band sharding (new; no reference analogue) --------------------
 * Rank `rank` of `world` traces (direct_stage + indirect_stage) only full-res rows
 * [y0, y1) of the frame (band edges multiples of 8: the direct stage works in 8x8 pixel tiles, the quarter-res stage lays its
 * 8x8 tiles over absolute tile rows and masks what belongs to the neighbour).
 * eid_renderer_run_trace enqueues the two trace kernels for the band; the caller then
 * all-gathers the three exchange buffers below (the ONE collective) and calls
 * eid_renderer_run_post, which runs denoise+compose on the full frame. */
EID_API int  eid_renderer_set_band(eid_renderer* r, uint32_t y0, uint32_t y1);
/* Interleaved ownership for load balance: rows are cut into stripes of `stripeRows` (multiple of 16); stripe i belongs to
 * rank i % world.  `world` consecutive stripes form an exchange group: a contiguous region of equal-sized chunks, completed on
 * every rank by one in-place all-gather.  The renderer's allocation height must be a multiple of world*stripeRows. */
EID_API int  eid_renderer_set_stripes(eid_renderer* r, uint32_t rank, uint32_t world, uint32_t stripeRows);
EID_API int  eid_renderer_exchange_groups(eid_renderer* r);
/* device pointer, byte offset and byte size of this rank's chunk in exchange group `group` of buffer `which` */
EID_API int  eid_renderer_exchange_range(eid_renderer* r, int which, uint32_t group, void** dev_base, uint64_t* offset, uint64_t* bytes);
EID_API int  eid_renderer_run_trace(eid_renderer* r, const RtxState* state, int frames);
/* eid_renderer_run_trace == run_direct followed by run_indirect; split so the host can overlap the exchange of the G-buffer and
 * the direct image (complete after run_direct) with indirect_stage */
EID_API int  eid_renderer_run_direct(eid_renderer* r, const RtxState* state, int frames);
EID_API int  eid_renderer_run_indirect(eid_renderer* r, const RtxState* state, int frames);
EID_API int  eid_renderer_run_post(eid_renderer* r, const RtxState* state, int frames);
/* Mode B (two exchange steps): denoise + compose only what this rank's band of the FINAL images needs (the A-Trous levels are
 * evaluated on the band plus the reach of the later levels); the caller then all-gathers EID_BUF_DIRECT and EID_BUF_INDIRECT. */
EID_API int  eid_renderer_run_post_band(eid_renderer* r, const RtxState* state, int frames);
/* = eid_renderer_exchange_range(group 0): device pointer + byte range of this rank's band inside buffer `which`
 * (EID_BUF_THIS_GBUFFER, EID_BUF_DIRECT, EID_BUF_DENOISE_IND_A, EID_BUF_MOTION, reservoirs) */
EID_API int  eid_renderer_band_range(eid_renderer* r, int which, void** dev_base, uint64_t* offset, uint64_t* bytes);

/* ---- eid_group: one frame on the N GPUs of a node (SURVEY.md §8(b) last row, §8(e); new, no reference analogue) -------------------------
 * One process or thread per GPU.  Every rank creates its scene / accel / renderer on its own device, the renderer with the padded height of
 * eid_group_layout (N equal bands, multiples of 8 rows), then joins the group: rank 0 obtains a 128-byte id with eid_group_unique_id and
 * the host hands it to the other ranks by whatever means it has (MPI, a socket, torch.distributed, a file).  The library owns the NCCL
 * communicator (libnccl.so.2, loaded at the first eid_group call) and a communication stream; eid_group_run enqueues the whole frame —
 * band trace stages, exchange A (G-buffer + direct image, behind indirect_stage), exchange B (quarter-res indirect image), denoise +
 * compose, exchange C (all-gather of the composed images) — each exchange ONE NCCL launch.  Every rank ends with the complete composed
 * frame, bit-identical to eid_renderer_run on one GPU (also with a moving camera: see eid_group_set_mode). */
typedef struct eid_group eid_group;
typedef struct eid_group_info {
  int32_t  rank, world;
  uint32_t y0, y1;            /* full-res rows this rank owns (y1 may exceed the rendered height on the last rank) */
  uint32_t bandRows;
  int32_t  ncclVersion;       /* e.g. 22809 */
  uint64_t collectives;       /* NCCL launches since creation */
  /* stage pipeline (eid_group_create_pipeline) only; zero otherwise */
  int32_t  stages;            /* EID_STAGE_* bits this rank runs */
  int32_t  nDirect, nIndirect, nPost;
  int32_t  streamMemOps;      /* 1: flags are awaited with cuStreamWaitValue32, 0: with a polling kernel */
  int32_t  pad_;
  uint64_t peerCopies;        /* peer (NVLink) copies enqueued since creation, and their bytes */
  uint64_t peerBytes;
} eid_group_info;
EID_API int  eid_group_layout(uint32_t height, int world, int rank, uint32_t* y0, uint32_t* y1, uint32_t* padded_height);
EID_API int  eid_group_unique_id(void* id128);
EID_API int  eid_group_create(eid_group** out, eid_renderer* r, int rank, int world, const void* id128);
EID_API void eid_group_destroy(eid_group* g);
/* post_sharded 1 (default): every rank denoises + composes its band only; 0: the whole frame on every rank (no exchange C).
 * history: how last frame's reservoirs cross band edges for temporal reuse.  0: never (exact for a static camera only); 1: gathered every
 * frame behind the post stages; 2 (default): gathered lazily, before the direct stage of a frame whose camera moved (projView != lastProjView).
 * gather_final 1 (default): exchange C; 0: every rank keeps only its band of the composed images. */
EID_API int  eid_group_set_mode(eid_group* g, int post_sharded, int history, int gather_final);
EID_API int  eid_group_run(eid_group* g, const RtxState* state, int frames);
/* host delivery without a funnel: every rank copies ITS band of the composed images into the (shared, pinned) host images over its own
 * PCIe link, pipelined behind the next frame; exchange C is skipped.  Images: size.x * 16 bytes per row, full frame. */
EID_API int  eid_group_render_host_async(eid_group* g, const SceneCamera* cam, const RtxState* state, int frames,
                                         float* direct_host, float* indirect_host);
EID_API int  eid_group_wait_host(eid_group* g);
EID_API int  eid_group_sync(eid_group* g);
EID_API int  eid_group_get_info(eid_group* g, eid_group_info* out);

/* ---- eid_group as a STAGE PIPELINE (csrc/pipeline.cu; new, no reference analogue) -------------------------------------------------------
 * The stages of Renderer::run (renderer.cpp:154-206) only depend on their own history: direct_stage(f) on direct_stage(f-1), indirect_stage(f)
 * on direct_stage(f) + indirect_stage(f-1), denoise / compose(f) on the two trace stages of f.  The N ranks therefore form a pipeline:
 * n_direct ranks run direct_stage on row bands, n_indirect ranks indirect_stage, n_post ranks denoise + compose, each group one frame behind
 * the one before it — throughput is set by the slowest stage instead of by the whole frame plus exchanges, every frame stays bit-identical
 * to eid_renderer_run, and latency grows by the two hand-overs (eid_group_create's row bands remain the low-latency mode).
 * There is no collective: ranks map each other's buffers (CUDA IPC handles, exchanged once through a rendezvous file in /dev/shm that is
 * keyed by id128), producers write the rows a consumer needs straight into the consumer's buffers with peer copies over NVLink and raise a
 * sequence flag the consumer's stream waits for (cuStreamWaitValue32); acknowledgements flow back the same way.  One PROCESS per rank.
 * n_direct = n_indirect = n_post = 0 selects the default split (2: 1|0|1 — the post rank also runs indirect_stage —, 4: 2|1|1, 8: 3|3|2).
 * Every rank creates its renderer with eid_pipeline_layout.paddedHeight rows; `height` is the rendered height the bands are cut from.
 * eid_group_run / render_host_async / wait_host / sync / get_info / destroy work on such a group; eid_group_set_mode only uses `history`.
 * After the last frame, direct ranks hold the G-buffer / direct reservoirs of their band, indirect ranks the indirect reservoirs, post
 * ranks their band of the two composed images. */
enum { EID_STAGE_DIRECT = 1, EID_STAGE_INDIRECT = 2, EID_STAGE_POST = 4 };
typedef struct eid_pipeline_layout {
  int32_t  nDirect, nIndirect, nPost;
  int32_t  stages;            /* EID_STAGE_* bits of this rank */
  int32_t  index, count;      /* this rank's band: `index` of `count` within its stage */
  uint32_t y0, y1;            /* full-res rows of that band (y1 may exceed the rendered height) */
  uint32_t paddedHeight;      /* allocation height for eid_renderer_create on every rank */
} eid_pipeline_layout;
EID_API int  eid_group_pipeline_layout(uint32_t height, int world, int rank, int n_direct, int n_indirect, int n_post, eid_pipeline_layout* out);
EID_API int  eid_group_random_id(void* id128);      /* 128 random bytes: a group id that does not need NCCL */
EID_API int  eid_group_create_pipeline(eid_group** out, eid_renderer* r, int rank, int world, const void* id128, uint32_t height,
                                       int n_direct, int n_indirect, int n_post);

#ifdef __cplusplus
}
#endif
#endif /* EIDOLA_H */

/*
 * host_device.h — shared ABI of the EIDOLA-B200 render loop.
 *
 * Plain-C restatement of the struct layouts that the reference shares between its
 * C++ host and GLSL device code (reference: shaders/host_device.h:68-375).  Field order,
 * field names and byte sizes are kept identical ("scalar" block layout == C packing of
 * 4-byte scalars) so a host that fills the reference's structs can hand them to this
 * library unchanged.  Everything here is POD and usable from C, C++ and CUDA.
 *
 * Sizes are pinned by the static asserts at the bottom (values from SURVEY.md §4, obtained
 * by compiling the reference header with g++).
 */
#ifndef EIDOLA_HOST_DEVICE_H
#define EIDOLA_HOST_DEVICE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- small POD vector types (GLSL vecN / nvmath::vecNf twins) ------------------------ */
typedef struct { float x, y; } eid_vec2;
typedef struct { float x, y, z; } eid_vec3;
typedef struct { float x, y, z, w; } eid_vec4;
typedef struct { int32_t x, y; } eid_ivec2;
/* column-major 4x4 (nvmath::mat4f): m[c*4 + r] is row r, column c */
typedef struct { float m[16]; } eid_mat4;

/* ---- compile-time switches of the reference (host_device.h:27-29), all off ------------ */
#define EID_INDIRECT_PRE_UPSCALE 0
#define EID_DENOISER_DIRECT_BILATERAL 0
#define EID_DENOISER_INDIRECT_BILATERAL 0

/* work-group shape of every dispatch (host_device.h:31-38) */
#define EID_BLOCK_X 8
#define EID_BLOCK_Y 8

/* DebugMode (host_device.h:128-139) */
enum {
  eNoDebug = 0, eDirectStage = 1, eIndirectStage = 2, eBaseColor = 3, eNormal = 4,
  eDepth = 5, eMetallic = 6, eEmissive = 7, eRoughness = 8, eTexcoord = 9
};

/* ReSTIRState (host_device.h:142-148) */
enum { eNone = 0, eRIS = 1, eSpatial = 2, eTemporal = 3, eSpatiotemporal = 4 };

#define CAMERA_NEAR 0.001f
#define CAMERA_FAR 1000.0f

/* host_device.h:153-165 — uniform block read by every kernel */
typedef struct SceneCamera {
  eid_mat4 viewInverse;
  eid_mat4 projInverse;
  eid_mat4 projView;
  eid_mat4 lastView;
  eid_mat4 lastProjView;
  eid_vec3 lastPosition;
  int32_t  nbLights;
} SceneCamera;

/* host_device.h:167-174 — compressed vertex, 32 B */
typedef struct VertexAttributes {
  eid_vec3 position;
  uint32_t normal;    /* oct-encoded unit vector */
  eid_vec2 texcoord;  /* LSB of .y carries the tangent handedness */
  uint32_t tangent;   /* oct-encoded unit vector */
  uint32_t color;     /* RGBA8 unorm */
} VertexAttributes;

#define ALPHA_OPAQUE 0
#define ALPHA_MASK 1
#define ALPHA_BLEND 2
#define MAX_IOR_MINUS_ONE 3.f

/* host_device.h:183-204 — 80 B */
typedef struct GltfShadeMaterial {
  eid_vec4 pbrBaseColorFactor;
  int32_t  pbrBaseColorTexture;
  float    pbrMetallicFactor;
  float    pbrRoughnessFactor;
  int32_t  pbrMetallicRoughnessTexture;
  int32_t  emissiveTexture;
  eid_vec3 emissiveFactor;
  int32_t  normalTexture;
  float    normalTextureScale;
  float    transmissionFactor;
  int32_t  transmissionTexture;
  float    ior;
  int32_t  alphaMode;
  float    alphaCutoff;
  int32_t  pad;
} GltfShadeMaterial;

/* host_device.h:207-238 — the 100-byte per-frame "push constant" */
typedef struct RtxState {
  int32_t  frame;
  int32_t  maxDepth;
  int32_t  modulate;
  float    fireflyClampThreshold;
  float    hdrMultiplier;
  int32_t  debugging_mode;
  float    environmentProb;
  uint32_t time;
  int32_t  ReSTIRState;
  int32_t  RISSampleNum;
  int32_t  reservoirClamp;
  int32_t  accumulate;
  eid_ivec2 size;
  float    envMapLuminIntegInv;
  float    lightLuminIntegInv;
  int32_t  MIS;
  float    sigLuminDirect;
  float    sigNormalDirect;
  float    sigDepthDirect;
  int32_t  denoise;
  float    sigLuminIndirect;
  float    sigNormalIndirect;
  float    sigDepthIndirect;
  int32_t  denoiseLevel;
} RtxState;

/* host_device.h:242-247 — 20 B of payload, 24 B stride (uint64 alignment) */
typedef struct InstanceData {
  uint64_t vertexAddress;
  uint64_t indexAddress;
  int32_t  materialIndex;
} InstanceData;

enum { LightType_Directional = 0, LightType_Point = 1, LightType_Spot = 2, LightType_Triangle = 3 };

/* host_device.h:260-284 — ReSTIR sample / reservoir records */
typedef struct LightSample { eid_vec3 Li; eid_vec3 wi; float dist; } LightSample;
typedef struct GISample { eid_vec3 L; eid_vec3 xv, nv; eid_vec3 xs, ns; float pHat; } GISample;
typedef struct DirectReservoir { LightSample lightSample; uint32_t num; float weight; } DirectReservoir;
typedef struct IndirectReservoir { GISample giSample; uint32_t num; float weight; float bigW; } IndirectReservoir;

/* host_device.h:287-293 — alias-method cell */
typedef struct ImptSampData { int32_t alias; float q; float pdf; float aliasPdf; } ImptSampData;

/* host_device.h:295-311 — 80 B */
typedef struct PuncLight {
  int32_t  type;
  eid_vec3 direction;
  float    intensity;
  eid_vec3 color;
  eid_vec3 position;
  float    range;
  float    outerConeCos;
  float    innerConeCos;
  eid_vec2 padding;
  ImptSampData impSamp;
} PuncLight;

/* host_device.h:313-325 — 96 B, vertices in WORLD space */
typedef struct TrigLight {
  uint32_t matIndex;
  uint32_t transformIndex;
  eid_vec3 v0, v1, v2;
  eid_vec2 uv0, uv1, uv2;
  ImptSampData impSamp;
  eid_vec3 pad;
} TrigLight;

/* host_device.h:327-333 */
typedef struct LightBufInfo {
  uint32_t puncLightSize;
  uint32_t trigLightSize;
  float    trigSampProb;
  int32_t  pad;
} LightBufInfo;

/* host_device.h:336-375 — display-side structs, kept for layout completeness only */
typedef struct Tonemapper {
  float brightness, contrast, saturation, vignette;
  float avgLum, zoom;
  eid_vec2 renderingRatio;
  int32_t autoExposure;
  float Ywhite, key;
  int32_t pad;
} Tonemapper;

typedef struct SunAndSky {
  eid_vec3 rgb_unit_conversion; float multiplier;
  float haze, redblueshift, saturation, horizon_height;
  eid_vec3 ground_color; float horizon_blur;
  eid_vec3 night_color; float sun_disk_intensity;
  eid_vec3 sun_direction; float sun_disk_scale;
  float sun_glow_intensity; int32_t y_is_up; int32_t physically_scaled_sun; int32_t in_use;
} SunAndSky;

#ifdef __cplusplus
}  /* extern "C" */
#define EID_SA(T, n) static_assert(sizeof(T) == (n), #T " must be " #n " bytes (reference host_device.h layout)")
EID_SA(SceneCamera, 336);
EID_SA(VertexAttributes, 32);
EID_SA(GltfShadeMaterial, 80);
EID_SA(RtxState, 100);
EID_SA(InstanceData, 24);
EID_SA(LightSample, 28);
EID_SA(GISample, 64);
EID_SA(DirectReservoir, 36);
EID_SA(IndirectReservoir, 76);
EID_SA(ImptSampData, 16);
EID_SA(PuncLight, 80);
EID_SA(TrigLight, 96);
EID_SA(LightBufInfo, 16);
EID_SA(Tonemapper, 48);
EID_SA(SunAndSky, 96);
#undef EID_SA
#endif

#endif /* EIDOLA_HOST_DEVICE_H */

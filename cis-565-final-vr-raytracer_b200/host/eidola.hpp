// eidola.hpp — C++17 host-side mirror of the reference's Scene / AccelStructure / Renderer classes on top of the
// C-ABI (include/eidola.h).  Same method names and argument meaning as the reference (src/scene.hpp:60-83,
// src/accelstruct.hpp:40-46, src/renderer.hpp:52-61); Vulkan handles are gone, errors surface as eidola::Error
// (the reference returns bool / asserts, scene.cpp:164-169).  Header-only; link against libeidola.so.
#pragma once
#include <stdexcept>
#include <string>
#include <utility>
#include "eidola.h"

namespace eidola {

struct Error : std::runtime_error {
  int code;
  Error(int c, const char* what) : std::runtime_error(what ? what : "eidola error"), code(c) {}
};
inline void check(int rc) { if (rc != EID_OK) throw Error(rc, eid_last_error()); }

struct Extent2D { uint32_t width, height; };   // VkExtent2D stand-in

class Scene {
 public:
  // Scene::setup(device, physicalDevice, queue, allocator) -> a CUDA ordinal is all that is left (scene.cpp:45-51)
  void setup(int cudaDevice = 0) { destroy(); check(eid_scene_create(&m_h, cudaDevice)); }
  // Scene::load(filename) -> bool (scene.cpp:57-125)
  bool load(const std::string& filename) { return eid_scene_load_gltf(m_h, filename.c_str()) == EID_OK; }
  bool load(const eid_scene_desc& desc) { return eid_scene_load_desc(m_h, &desc) == EID_OK; }
  // CameraManip.setCamera / setLookat (scene.cpp:298-308, main.cpp:67-68)
  void setLookat(const float eye[3], const float center[3], const float up[3], float fovDeg) { check(eid_scene_set_lookat(m_h, eye, center, up, fovDeg)); }
  // Scene::updateCamera(cmdBuf, size) (scene.cpp:777-826)
  void updateCamera(Extent2D size) { check(eid_scene_update_camera(m_h, size.width, size.height)); }
  SceneCamera getCamera() const { SceneCamera c; check(eid_scene_get_camera(m_h, &c)); return c; }   // scene.hpp:80
  void setCamera(const SceneCamera& c) { check(eid_scene_set_camera(m_h, &c)); }
  eid_scene_info getStat() const { eid_scene_info i; check(eid_scene_get_info(m_h, &i)); return i; }  // getStat / m_*LightWeight
  void destroy() { if (m_h) { eid_scene_destroy(m_h); m_h = nullptr; } }
  ~Scene() { destroy(); }
  eid_scene* handle() const { return m_h; }
 private:
  eid_scene* m_h = nullptr;
};

class AccelStructure {
 public:
  // AccelStructure::create(gltfScene, vertexBuffers, indexBuffers) (accelstruct.cpp:55-65): the buffers live in the Scene
  // mode: EID_ACCEL_AUTO (default), EID_ACCEL_FLAT (one world-space BVH), EID_ACCEL_TWO_LEVEL (BLAS per prim mesh + TLAS, as the reference builds),
  // optionally | EID_ACCEL_FAST_BUILD (Morton build on the GPU) instead of the default EID_ACCEL_FAST_TRACE (binned SAH, the reference's build flag)
  void create(Scene& scene, int mode = EID_ACCEL_AUTO) { destroy(); check(eid_accel_build_ex(scene.handle(), mode, &m_h)); }
  eid_accel_info info() const { eid_accel_info i; check(eid_accel_get_info(m_h, &i)); return i; }
  void destroy() { if (m_h) { eid_accel_destroy(m_h); m_h = nullptr; } }
  ~AccelStructure() { destroy(); }
  eid_accel* getTlas() const { return m_h; }   // accelstruct.hpp:44
 private:
  eid_accel* m_h = nullptr;
};

// HdrSampling (src/hdr_sampling.hpp:43-48): the environment map, its importance-sampling alias map, integral and average
class HdrSampling {
 public:
  void setup(int cudaDevice = 0) { m_device = cudaDevice; }
  // HdrSampling::loadEnvironment(hdrImage) (hdr_sampling.cpp:48-105): Radiance .hdr file -> RGBA32F texels + createEnvironmentAccel
  void loadEnvironment(const std::string& hdrImage) { destroy(); check(eid_env_load_hdr(&m_h, m_device, hdrImage.c_str())); }
  void setPixels(const float* rgba, uint32_t width, uint32_t height) { destroy(); check(eid_env_create(&m_h, m_device, rgba, width, height)); }
  float getIntegral() const { return eid_env_integral(m_h); }   // hdr_sampling.hpp:47
  float getAverage() const { return eid_env_average(m_h); }     // hdr_sampling.hpp:48
  void destroy() { if (m_h) { eid_env_destroy(m_h); m_h = nullptr; } }
  ~HdrSampling() { destroy(); }
  eid_env* handle() const { return m_h; }
 private:
  eid_env* m_h = nullptr;
  int m_device = 0;
};

class Renderer {
 public:
  // Renderer::create(size, rtDescSetLayouts, scene) (renderer.cpp:97-148): descriptor-set layouts become the two handles
  void create(Extent2D size, Scene& scene, AccelStructure& accel, void* cudaStream = nullptr) {
    destroy();
    check(eid_renderer_create(&m_h, scene.handle(), accel.getTlas(), size.width, size.height, cudaStream));
  }
  // Renderer::run(cmdBuf, state, profiler, descSets, frames) (renderer.cpp:154-206): enqueues the 12 dispatches, asynchronous
  void run(const RtxState& state, int frames) { check(eid_renderer_run(m_h, &state, frames)); }
  // Renderer::update(size) (renderer.cpp:209-225)
  void update(Extent2D size) { check(eid_renderer_resize(m_h, size.width, size.height)); }
  void sync() { check(eid_renderer_sync(m_h)); }
  void setEnvironmentConstant(const float rgb[3]) { check(eid_renderer_set_env_constant(m_h, rgb)); }
  void setEnvironment(HdrSampling* env) { check(eid_renderer_set_env(m_h, env ? env->handle() : nullptr)); }   // the S_ENV descriptor set (sample_example.cpp:318)
  void runOutput(const Tonemapper& tm) { check(eid_renderer_run_output(m_h, &tm)); }   // RenderOutput::run -> post.frag (render_output.cpp:224-240)
  void setSunAndSky(const SunAndSky& ss) { check(eid_renderer_set_sun_and_sky(m_h, &ss)); }   // SampleExample::m_sunAndSky (sample_example.cpp:172)
  void setWavefront(bool on, int traceBlocks = 0) { check(eid_renderer_set_wavefront(m_h, on ? 1 : 0, traceBlocks)); }
  void setStrictMath(bool on) { check(eid_renderer_set_strict_math(m_h, on ? 1 : 0)); }
  // the reference's compile-time shader switches (host_device.h:27-29, indirect_stage.comp:35, the never-dispatched direct_gen / direct_reuse pair)
  void setVariant(int flags) { check(eid_renderer_set_variant(m_h, flags)); }
  void setFramesInFlight(int n) { check(eid_renderer_set_pipeline(m_h, n)); }          // 2: direct_stage of frame f + 1 beside the later stages of frame f
  void setDenoiseTiles(int mode, int rowsPerThread = 0) { check(eid_renderer_set_denoise_tiles(m_h, mode, rowsPerThread)); }
  // the two RGBA32F images RenderOutput hands to post.frag (render_output.cpp:195-215): device pointers
  std::pair<const float*, const float*> outputs() const { const float *d, *i; check(eid_renderer_get_outputs(m_h, &d, &i)); return {d, i}; }
  void renderToHost(const SceneCamera* cam, const RtxState& st, int frames, float* direct, float* indirect) {
    check(eid_renderer_render_host(m_h, cam, &st, frames, direct, indirect));
  }
  eid_frame_stats stats() const { eid_frame_stats s; check(eid_renderer_get_stats(m_h, &s)); return s; }
  const std::string name() { return std::string("RQ"); }   // renderer.hpp:56
  void destroy() { if (m_h) { eid_renderer_destroy(m_h); m_h = nullptr; } }
  ~Renderer() { destroy(); }
  eid_renderer* handle() const { return m_h; }
 private:
  eid_renderer* m_h = nullptr;
};

// One rank of an N-GPU frame (eid_group, include/eidola.h): the renderer must have been created with the padded height of Group::layout.
// Rank 0 calls Group::uniqueId() and hands the 128 bytes to the other ranks (MPI_Bcast, a socket, a file ...).
class Group {
 public:
  struct Layout { uint32_t y0, y1, paddedHeight; };
  static Layout layout(uint32_t height, int world, int rank = 0) { Layout l{}; check(eid_group_layout(height, world, rank, &l.y0, &l.y1, &l.paddedHeight)); return l; }
  static void uniqueId(unsigned char id[128]) { check(eid_group_unique_id(id)); }
  void create(Renderer& r, int rank, int world, const unsigned char* id128 = nullptr) { destroy(); check(eid_group_create(&m_h, r.handle(), rank, world, id128)); }
  // stage pipeline (csrc/pipeline.cu): direct | indirect | post ranks joined by NVLink peer copies; the renderer needs pipelineLayout().paddedHeight rows
  static eid_pipeline_layout pipelineLayout(uint32_t height, int world, int rank, int nDirect = 0, int nIndirect = 0, int nPost = 0) {
    eid_pipeline_layout l{}; check(eid_group_pipeline_layout(height, world, rank, nDirect, nIndirect, nPost, &l)); return l;
  }
  static void randomId(unsigned char id[128]) { check(eid_group_random_id(id)); }
  void createPipeline(Renderer& r, int rank, int world, const unsigned char* id128, uint32_t height, int nDirect = 0, int nIndirect = 0, int nPost = 0) {
    destroy(); check(eid_group_create_pipeline(&m_h, r.handle(), rank, world, id128, height, nDirect, nIndirect, nPost));
  }
  void setMode(bool postSharded, int history, bool gatherFinal) { check(eid_group_set_mode(m_h, postSharded, history, gatherFinal)); }
  void run(const RtxState& state, int frames) { check(eid_group_run(m_h, &state, frames)); }       // the whole multi-GPU frame, asynchronous
  void renderToHostAsync(const SceneCamera* cam, const RtxState& st, int frames, float* direct, float* indirect) {
    check(eid_group_render_host_async(m_h, cam, &st, frames, direct, indirect));                  // every rank delivers ITS band into the shared host images
  }
  void waitHost() { check(eid_group_wait_host(m_h)); }
  void sync() { check(eid_group_sync(m_h)); }
  eid_group_info info() const { eid_group_info i; check(eid_group_get_info(m_h, &i)); return i; }
  void destroy() { if (m_h) { eid_group_destroy(m_h); m_h = nullptr; } }
  ~Group() { destroy(); }
 private:
  eid_group* m_h = nullptr;
};

// RenderOutput (src/render_output.hpp:44-60, render_output.cpp:224-240): owns the tonemapper settings; run() = post.frag over the renderer's two
// result images — m_tm, or m_depthTm in the depth view, with zoom and renderingRatio overwritten per call, exactly like the reference
class RenderOutput {
 public:
  Tonemapper m_tm{1.0f, 1.0f, 1.0f, 0.0f, 1.0f, 1.0f, {1.0f, 1.0f}, 0, 0.5f, 0.5f, 0};
  Tonemapper m_depthTm{0.0f, 2.2f, 0.0f, 0.0f, 0.0f, 0.0f, {0.0f, 0.0f}, 0, 0.0f, 0.0f, 0};
  void run(Renderer& renderer, const RtxState& state, float zoom = 1.0f, eid_vec2 ratio = eid_vec2{1.0f, 1.0f}) {
    Tonemapper tm = (state.debugging_mode == eDepth) ? m_depthTm : m_tm;
    tm.zoom = zoom;
    tm.renderingRatio = ratio;
    renderer.runOutput(tm);          // (RenderOutput::genMipmap runs inside when tm.autoExposure bit 0 is set)
  }
};

// SampleExample::m_rtxState defaults (sample_example.hpp:154-184)
inline RtxState defaultRtxState(uint32_t w, uint32_t h) {
  RtxState s{};
  s.maxDepth = 4; s.modulate = 1; s.fireflyClampThreshold = 1.f; s.hdrMultiplier = 1.f; s.environmentProb = 0.25f;
  s.ReSTIRState = eTemporal; s.RISSampleNum = 4; s.reservoirClamp = 80; s.size = {(int32_t)w, (int32_t)h}; s.MIS = 1;
  s.sigLuminDirect = 0.4f; s.sigNormalDirect = 0.1f; s.sigDepthDirect = 0.02f; s.denoise = 1;
  s.sigLuminIndirect = 4.f; s.sigNormalIndirect = 0.4f; s.sigDepthIndirect = 1.f;
  return s;
}

}  // namespace eidola

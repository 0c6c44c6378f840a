// renderer.h — the Renderer object (buffers + launch schedule state) shared by render.cu (C-ABI of the reference's Renderer class)
// and group.cu (multi-GPU group: one renderer per rank + the NCCL exchange steps).
#pragma once
#ifndef EID_L2_PERSIST_DEFAULT
#define EID_L2_PERSIST_DEFAULT 0   // access-policy window of the acceleration structure (render.cu: applyL2Policy); EIDOLA_L2_PERSIST overrides
#endif
#include <map>
#include <tuple>
#include <vector>
#include "stages.h"

using namespace eid;   // internal header of the two translation units that implement the C-ABI

// ------------------------------------------------------------------------------------------------
// Renderer object
// ------------------------------------------------------------------------------------------------
struct eid_renderer {
  eid_scene* scene = nullptr;
  eid_accel* accel = nullptr;
  int device = 0;
  uint32_t width = 0, height = 0;
  cudaStream_t stream = nullptr;
  bool ownStream = false;
  uint4* gbuffer[2] = {nullptr, nullptr};
  short2* motion = nullptr;
  float* directResv[2] = {nullptr, nullptr};
  float* indirectResv[2] = {nullptr, nullptr};
  float4* directImgs[2] = {nullptr, nullptr};   // thisDirectResultImage, one per ping-pong parity (a frame in flight keeps its own while the next one traces)
  float4* directImg = nullptr;                  // = directImgs[parity of the frame enqueued last]
  float4* indirectImgs[2] = {nullptr, nullptr};
  float4* indirectImg = nullptr;                // = indirectImgs[parity of the frame enqueued last]
  uint4* k2G[2] = {nullptr, nullptr}; short2* k2Mv[2] = {nullptr, nullptr};   // FrameParams::k2G / k2Mv, per parity
  // frames in flight (eid_renderer_set_pipeline): direct_stage of frame f + 1 runs on `k1Stream` while indirect_stage / denoise / compose of
  // frame f are still on the render stream; evK1Done orders K2 / K3 after K1, evFrameDone[parity] lets K1 reuse a parity's buffers
  int variant = 0;            // EID_VARIANT_* bits
  int pipeline = 0;
  cudaStream_t k1Stream = nullptr;
  cudaEvent_t evK1Done = nullptr, evFrameDone2[2] = {nullptr, nullptr};
  bool frameDoneValid[2] = {false, false};
  // host delivery straight from a parity's result images (no staging copy): the next frame of the same parity waits for that copy
  cudaEvent_t evCopyDone2[2] = {nullptr, nullptr};
  bool copyPending2[2] = {false, false};
  bool k1MustWaitStream = false;     // a strictly ordered entry point (run_trace, run_direct, the group schedule ...) ran on the render stream since the last pipelined frame
  cudaEvent_t evOrder = nullptr;
  cudaStream_t groupStream = nullptr;   // communication stream of the eid_group this renderer belongs to (synchronised with the render stream by sync / read / get_stats)
  float* tempDirectResv = nullptr; float4* spatialCont = nullptr;   // spatial reuse (eSpatial / eSpatiotemporal), allocated on first use
  float4* denoiseTemp[4] = {nullptr, nullptr, nullptr, nullptr};
  float4* indIn[2] = {nullptr, nullptr};        // stage pipeline (pipeline.cu): pre-denoise indirect image received from the indirect ranks, per parity
  float4* geom[4] = {nullptr, nullptr, nullptr, nullptr};   // geomPos, geomNrm, geomPosH, geomNrmH
  float4* displayF = nullptr; uchar4* display8 = nullptr;   // output of the display pass (post.frag), allocated on first use
  float4* mipScratch = nullptr;                             // auto exposure: two ping-pong mip levels + the two 1x1 averages
  // wavefront K2 scratch (WaveView): sized for the allocation and for `waveTerms` NEE depths; (re)allocated on demand
  void* waveMem = nullptr; uint32_t waveSlots = 0; int waveTerms = 0; uint32_t* waveCtr = nullptr;
  cudaStream_t shadowStream = nullptr; cudaEvent_t evWave = nullptr, evWaveJoin = nullptr; bool waveOverlap = true;
  // a second scratch + shadow stream: the stage pipeline's indirect ranks trace the paths of frame f + 1 while frame f is still in flight
  // (only k_gi_finish depends on the previous frame), pipeline.cu
  void* waveMem2 = nullptr; uint32_t waveSlots2 = 0; int waveTerms2 = 0; uint32_t* waveCtr2 = nullptr;
  cudaStream_t shadowStream2 = nullptr; cudaEvent_t evWave2 = nullptr, evWaveJoin2 = nullptr;
  int wavefront = 1;          // 1 (default): K2 runs as ray queues + dynamic-fetch traversal when the scene allows it; 0: one mega-kernel
  int traceBlocks = 0;        // grid of k_trace_queue (blocks of 128 threads); 0 = EID_TQ_MIN_BLOCKS per SM
  int smCount = 0;
  int denoiseRowBlock = 2;    // legacy A-Trous kernel: pixels of one column filtered per thread (1, 2 or 4; 2 measured fastest)
  int denoiseTiles = 1;       // 1 (default): shared-memory tile kernel fed by TMA; 2: same, tiles loaded with cp.async; 0: legacy kernel (L1-served taps)
  int denoiseTileRows = 2;    // tile kernel: lattice rows per thread (2 or 4; 2 measured faster: 66 vs 96 registers)
  std::map<std::tuple<const void*, int, int, int, int>, CUtensorMap> tmaps;   // (buffer, pitch, rows, level, tile height) -> lattice-view tensor map
  bool strictMath = false;    // bit-reproducible exp in the denoiser (parity runs) instead of MUFU ex2
  unsigned long long* counters = nullptr;
  unsigned long long* countersHost = nullptr;   // pinned
  float env[3] = {0.f, 0.f, 0.f};
  eid_env* envMap = nullptr;
  SunAndSky sunSky{};         // SampleExample::m_sunAndSky (sample_example.hpp:186-203); in_use = 0 until the host sets it
  int lastSet = 0;
  RtxState lastState{};
  bool hasRun = false;
  uint32_t sFirst = 0, sStride = 0, sRows = 0; bool stripesSet = false;   // multi-GPU row ownership (see FrameParams)
  bool profiling = false;
  size_t l2PersistBytes = 0;  // L2 set-aside holding the acceleration structure (applyL2Policy), 0 = none
  bool countVisits = false;   // profiling level 2: STATS kernels (node / triangle visit counters)
  cudaEvent_t ev[2 * EID_K_COUNT] = {};   // start/stop per stage
  cudaEvent_t evFork = nullptr, evJoin = nullptr, evPost = nullptr;
  cudaStream_t aux = nullptr;             // second stream: K3 runs beside K2/K4 (see launchFrame)
  bool overlap = true;
  bool postStarted = false;
  cudaStream_t copyStream = nullptr;      // eid_renderer_render_host_async: D2H of frame f overlaps the kernels of frame f+1
  cudaEvent_t evFrameDone = nullptr, evCopyDone = nullptr;
  float4* staging[2] = {nullptr, nullptr};
  bool copyPending = false;
  eid_frame_stats stats{};
  bool statsPending = false;

  void allocate();
  void release();
  void ensureWave(int terms, int which = 0);
  WaveView waveView(int which = 0) const;
};


// per-frame pieces of eid_renderer_run (render.cu), reused by the multi-GPU schedule of group.cu
void fillParams(eid_renderer* r, const RtxState& st, int frames, FrameParams& P);
void beginFrame(eid_renderer* r, cudaStream_t st = nullptr);
void stageDirect(eid_renderer* r, const FrameParams& P, cudaStream_t st, bool mark = true);
void stageIndirect(eid_renderer* r, const FrameParams& P, cudaStream_t st);
// indirect_stage in two halves (wavefront form only): the path tracing (k_gi_begin, ray queues, k_gi_bounce; needs this frame's G-buffer only) and
// k_gi_finish (temporal reuse: needs the previous frame's reservoirs).  `ctx` selects the shadow stream / events of scratch 0 or 1.
bool indirectIsWavefront(eid_renderer* r, const FrameParams& P);
void stageIndirectTrace(eid_renderer* r, const FrameParams& P, cudaStream_t st, int ctx);
void stageIndirectFinish(eid_renderer* r, const FrameParams& P, cudaStream_t st);
void launchPost(eid_renderer* r, const FrameParams& P, bool sharded);
void* bufferPtr(eid_renderer* r, int which, size_t& bytes);
// called by every entry point that enqueues stages on the render stream in strict order: orders them after any direct_stage still on the K1 stream
void strictOrder(eid_renderer* r);
void ensureCopyStream(eid_renderer* r);
// post stages one by one (the multi-GPU schedule starts the direct denoiser as soon as exchange A has landed, beside indirect_stage)
struct PostLayout { int first, stride, srows, count; bool sharded; };
PostLayout postLayout(const FrameParams& P, bool sharded);
void stagePrep(eid_renderer* r, const FrameParams& P, const PostLayout& L, cudaStream_t st);
void stageDenoiseDirect(eid_renderer* r, const FrameParams& P, const PostLayout& L, cudaStream_t st);
void stageDenoiseIndirect(eid_renderer* r, const FrameParams& P, const PostLayout& L, cudaStream_t st);
void stageCompose(eid_renderer* r, const FrameParams& P, const PostLayout& L, cudaStream_t st);
void endFrame(eid_renderer* r);
void markStart(eid_renderer* r, int stage, cudaStream_t st);
void markStop(eid_renderer* r, int stage, cudaStream_t st);

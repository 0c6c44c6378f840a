// render.cu — the per-frame render loop: CUDA kernels for the five live compute shaders of the reference
// and the Renderer object that owns their buffers and launch schedule.
//
//   k_direct_stage      shaders/direct_stage.comp    primary ray, G-buffer, motion index, RIS over M light
//                                                    candidates, shadow ray, temporal reservoir merge, shade
//   k_indirect_stage    shaders/indirect_stage.comp  quarter-res ReSTIR GI: BSDF-sampled path (NEE+MIS from
//                                                    depth 2), per-8x8-tile multibounce lottery, temporal reuse
//   k_denoise<false>    shaders/denoise_direct.comp  A-Trous level 0..3 (one launch per level)
//   k_denoise<true>     shaders/denoise_indirect.comp A-Trous level 0..4 at quarter res
//   k_compose           shaders/compose.comp         re-modulate by albedo, 2x nearest upsample of indirect
//
// Renderer::run (src/renderer.cpp:154-206) = strict stream order of the above: 1 + 1 + 4 + 5 + 1 launches.
#include <algorithm>
#include <cstring>
#include "renderer.h"

void eid_renderer::release() {
  for (int i = 0; i < 2; ++i) { cudaFree(gbuffer[i]); cudaFree(directResv[i]); cudaFree(indirectResv[i]); gbuffer[i] = nullptr; directResv[i] = nullptr; indirectResv[i] = nullptr; }
  cudaFree(motion); motion = nullptr;
  cudaFree(tempDirectResv); cudaFree(spatialCont); tempDirectResv = nullptr; spatialCont = nullptr;
  cudaFree(displayF); cudaFree(display8); displayF = nullptr; display8 = nullptr;
  cudaFree(mipScratch); mipScratch = nullptr;
  cudaFree(waveMem); waveMem = nullptr; cudaFree(waveCtr); waveCtr = nullptr; waveSlots = 0; waveTerms = 0;
  cudaFree(waveMem2); waveMem2 = nullptr; cudaFree(waveCtr2); waveCtr2 = nullptr; waveSlots2 = 0; waveTerms2 = 0;
  for (int i = 0; i < 2; ++i) { cudaFree(directImgs[i]); cudaFree(k2G[i]); cudaFree(k2Mv[i]); directImgs[i] = nullptr; k2G[i] = nullptr; k2Mv[i] = nullptr; }
  for (int i = 0; i < 2; ++i) { cudaFree(indirectImgs[i]); indirectImgs[i] = nullptr; }
  directImg = indirectImg = nullptr;
  copyPending2[0] = copyPending2[1] = false;
  frameDoneValid[0] = frameDoneValid[1] = false;
  tmaps.clear();
  for (auto& t : denoiseTemp) { cudaFree(t); t = nullptr; }
  for (auto& t : indIn) { cudaFree(t); t = nullptr; }
  for (auto& t : geom) { cudaFree(t); t = nullptr; }
}

// Renderer::createBuffer / createImage (renderer.cpp:227-302): 2 G-buffers, 1 motion image, 2+2 reservoir buffers,
// 4 denoise temporaries, 2 result images — all zero-initialised (the reference leaves them undefined).
void eid_renderer::allocate() {
  const size_t n = (size_t)width * height, ni = (size_t)(width / 2) * (height / 2);
  auto zalloc = [&](void** p, size_t bytes) {
    bytes = std::max<size_t>(bytes, 16);
    CUDA_CHECK(cudaMalloc(p, bytes));
    CUDA_CHECK(cudaMemsetAsync(*p, 0, bytes, stream));
  };
  for (int i = 0; i < 2; ++i) {
    zalloc((void**)&gbuffer[i], n * 16);
    zalloc((void**)&directResv[i], n * sizeof(DirectReservoir));
    zalloc((void**)&indirectResv[i], ni * sizeof(IndirectReservoir));
  }
  zalloc((void**)&motion, n * 4);
  // everything the A-Trous tile kernel reads through a lattice-view tensor map carries 17 rows of slack: the view of level l rounds
  // the row / column counts up to multiples of 2^l <= 16, so its last lattice row / column may address up to 15 rows + 15 texels
  // past the image (never used: the kernel invalidates texels outside the rendered size)
  const size_t slack = (size_t)17 * width * 16;
  for (int i = 0; i < 2; ++i) { zalloc((void**)&directImgs[i], n * 16 + slack); zalloc((void**)&k2G[i], ni * 16); zalloc((void**)&k2Mv[i], ni * 4); }
  directImg = directImgs[0];
  for (int i = 0; i < 2; ++i) zalloc((void**)&indirectImgs[i], n * 16 + slack);
  indirectImg = indirectImgs[0];
  for (auto& t : denoiseTemp) zalloc((void**)&t, n * 16 + slack);
  zalloc((void**)&geom[0], n * 16 + slack); zalloc((void**)&geom[1], n * 16 + slack);
  zalloc((void**)&geom[2], ni * 16 + slack); zalloc((void**)&geom[3], ni * 16 + slack);
  CUDA_CHECK(cudaStreamSynchronize(stream));
  hasRun = false; lastSet = 0;
  if (!stripesSet) { sFirst = 0; sRows = (height + 15) / 16 * 16; sStride = 1u << 20; }
}

// Scratch of the wavefront K2: planes of `waveSlots` float4 (one slot per thread of the largest possible K2 grid of this
// allocation).  Layout of waveMem in float4 units: rayQ0 2S | rayQ1 2S | hitQ S | misc S | thr S | gsXv S | gsNv S | gsXs S |
// gsNs S | hitL S | neeTerm T*S | shadowQ 2*T*S | occl (T*S uint32).  1080p, maxDepth 3: ~125 MB.
void eid_renderer::ensureWave(int terms, int which) {
  terms = std::max(terms, 1);
  const uint32_t tilesX = (width / 2 + 7) / 8, tilesY = ((height + 15) / 16 * 16 / 2 + 7) / 8 + 1;
  const uint32_t S = tilesX * tilesY * 64u;
  void*& mem = which ? waveMem2 : waveMem; uint32_t*& ctr = which ? waveCtr2 : waveCtr;
  uint32_t& slots = which ? waveSlots2 : waveSlots; int& have = which ? waveTerms2 : waveTerms;
  if (mem && slots == S && have >= terms) return;
  CUDA_CHECK(cudaDeviceSynchronize());
  cudaFree(mem); mem = nullptr;
  const size_t f4 = (size_t)S * (12 + 3 * (size_t)terms);
  CUDA_CHECK(cudaMalloc(&mem, f4 * 16 + (size_t)S * terms * 4));
  if (!ctr) CUDA_CHECK(cudaMalloc(&ctr, 128 * sizeof(uint32_t)));
  slots = S; have = terms;
}
WaveView eid_renderer::waveView(int which) const {
  WaveView V;
  const size_t S = which ? waveSlots2 : waveSlots, T = (size_t)(which ? waveTerms2 : waveTerms);
  float4* b = (float4*)(which ? waveMem2 : waveMem);
  V.slots = (uint32_t)S;
  V.rayQ[0] = b; V.rayQ[1] = b + 2 * S; V.hitQ = b + 4 * S; V.misc = (uint4*)(b + 5 * S); V.thr = b + 6 * S;
  V.gsXv = b + 7 * S; V.gsNv = b + 8 * S; V.gsXs = b + 9 * S; V.gsNs = b + 10 * S; V.hitL = b + 11 * S;
  V.neeTerm = b + 12 * S; V.shadowQ = b + (12 + T) * S; V.occl = (uint32_t*)(b + (12 + 3 * T) * S);
  V.ctr = which ? waveCtr2 : waveCtr;
  return V;
}

void fillParams(eid_renderer* r, const RtxState& st, int frames, FrameParams& P) {
  if (st.size.x <= 0 || st.size.y <= 0 || (uint32_t)st.size.x > r->width || (uint32_t)st.size.y > r->height)
    raise(EID_ERR_INVALID, "RtxState.size %dx%d outside the renderer allocation %ux%u", st.size.x, st.size.y, r->width, r->height);
  if (st.environmentProb > 0.0f && !r->envMap && r->sunSky.in_use != 1)
    raise(EID_ERR_UNSUPPORTED, "environmentProb > 0 needs an environment to sample: an HDR map (eid_env_create + eid_renderer_set_env) or sun & sky (eid_renderer_set_sun_and_sky with in_use = 1)");
  if (st.RISSampleNum < 0 || st.maxDepth < 0) raise(EID_ERR_INVALID, "negative RISSampleNum / maxDepth");
  const int set = (frames + 1) % 2;   // renderer.cpp:157; set i: last* = [i], this* = [!i] (renderer.cpp:341-375)
  P.st = st;
  P.cam = r->scene->host.camera;
  P.sc = r->scene->dev.view(r->scene->host);
  P.accel = r->accel->view();
  P.thisG = r->gbuffer[!set]; P.lastG = r->gbuffer[set];
  P.motion = r->motion;
  P.thisDR = r->directResv[!set]; P.lastDR = r->directResv[set];
  P.thisIR = r->indirectResv[!set]; P.lastIR = r->indirectResv[set];
  r->directImg = r->directImgs[set]; r->indirectImg = r->indirectImgs[set];
  if (r->copyPending2[set]) {        // a device-to-host copy of this parity's result images (two frames ago) must have drained before they are rewritten
    CUDA_CHECK(cudaStreamWaitEvent(r->stream, r->evCopyDone2[set], 0));
    if (r->pipeline) CUDA_CHECK(cudaStreamWaitEvent(r->k1Stream, r->evCopyDone2[set], 0));
    r->copyPending2[set] = false;
  }
  P.directImg = r->directImg; P.indirectImg = r->indirectImg;
  P.k2G = r->k2G[set]; P.k2Mv = r->k2Mv[set];
  P.variant = r->variant;
  P.directOut = (r->variant & EID_VARIANT_DIRECT_BILATERAL) ? r->denoiseTemp[0] : r->directImg;
  if ((st.ReSTIRState == eSpatial || st.ReSTIRState == eSpatiotemporal) && !r->tempDirectResv) {   // m_directTempResv (renderer.cpp:235), on first use
    const size_t n = (size_t)r->width * r->height;
    CUDA_CHECK(cudaMalloc((void**)&r->tempDirectResv, n * sizeof(DirectReservoir)));
    CUDA_CHECK(cudaMemsetAsync(r->tempDirectResv, 0, n * sizeof(DirectReservoir), r->stream));
    CUDA_CHECK(cudaMalloc((void**)&r->spatialCont, 3 * n * sizeof(float4)));
  }
  P.tempDR = r->tempDirectResv; P.spCont = r->spatialCont;
  P.dirA = r->denoiseTemp[0]; P.dirB = r->denoiseTemp[1]; P.indA = r->denoiseTemp[2]; P.indB = r->denoiseTemp[3];
  P.indIn = r->indIn[set] ? r->indIn[set] : P.indA;
  P.geomPos = r->geom[0]; P.geomNrm = r->geom[1]; P.geomPosH = r->geom[2]; P.geomNrmH = r->geom[3];
  for (int k = 0; k < 3; ++k) P.env.constant[k] = r->env[k];
  P.env.sunSky = r->sunSky;
  P.env.tex = r->envMap ? r->envMap->tex : nullptr; P.env.accel = r->envMap ? r->envMap->accel : nullptr;
  P.env.width = r->envMap ? (int)r->envMap->host.width : 0; P.env.height = r->envMap ? (int)r->envMap->host.height : 0;
  P.hasNonOpaque = r->scene->host.hasNonOpaque ? 1 : 0;
  P.pitch = (int)r->width; P.allocH = (int)r->height;
  P.sFirst = (int)r->sFirst; P.sStride = (int)r->sStride; P.sRows = (int)r->sRows;
  P.sCount = ((int)r->sFirst < st.size.y) ? (st.size.y - 1 - (int)r->sFirst) / (int)r->sStride + 1 : 0;   // stripes that start inside the frame
  P.counters = r->counters + EID_NUM_COUNTERS * set; P.totals = r->counters + 2 * EID_NUM_COUNTERS;
  memset(&P.wv, 0, sizeof(P.wv));
  if (r->wavefront && !P.hasNonOpaque && !P.accel.twoLevel && st.maxDepth <= GI_MAX_WAVE_DEPTH) { r->ensureWave(st.maxDepth - 1); P.wv = r->waveView(); }
  r->lastSet = set; r->lastState = st; r->hasRun = true;
}

// ---- stage launchers -----------------------------------------------------------------------------------------------------------
// Every stage records a start/stop CUDA-event pair on the stream it runs on when profiling is enabled (stage k: ev[2k], ev[2k+1]).
void markStart(eid_renderer* r, int stage, cudaStream_t st) { if (r->profiling) CUDA_CHECK(cudaEventRecord(r->ev[2 * stage], st)); }
void markStop(eid_renderer* r, int stage, cudaStream_t st) { if (r->profiling) CUDA_CHECK(cudaEventRecord(r->ev[2 * stage + 1], st)); }

void beginFrame(eid_renderer* r, cudaStream_t st) {
  CUDA_CHECK(cudaMemsetAsync(r->counters + EID_NUM_COUNTERS * r->lastSet, 0, 5 * sizeof(unsigned long long), st ? st : r->stream));   // this parity's per-frame counters
  memset(&r->stats, 0, sizeof(r->stats));
  r->postStarted = false;
}

void stageDirect(eid_renderer* r, const FrameParams& P, cudaStream_t st, bool mark) {
  if (mark) markStart(r, EID_K_DIRECT, st);
  if (P.sCount > 0) {
    dim3 g((P.st.size.x + 7) / 8, P.sCount * (P.sRows / 8));
    // TEX = false: lean variant for scenes without a single textured material (no texture branches, no tangent frame)
    const bool tex = r->scene->host.hasTextures || r->scene->host.hasNonOpaque || P.env.sunSky.in_use == 1 || P.accel.twoLevel;   // (the two-level walk lives in the full variants)
    const bool spatial = P.st.ReSTIRState == eSpatial || P.st.ReSTIRState == eSpatiotemporal;
    if (P.variant & EID_VARIANT_DIRECT_SPLIT) {
      launchDirectSplit(P, g, st, r->countVisits, tex);
      r->stats.kernelLaunches[EID_K_DIRECT] += 2;
    } else if (!spatial) {
      launchDirectStage(P, g, st, r->countVisits, tex, false, 0);
      r->stats.kernelLaunches[EID_K_DIRECT]++;
    } else {
      // spatial reuse: every pixel up to its tempDirectResv write (owned stripes, then — when the stripes do not cover the frame — the
      // row above and the row below each stripe), then the neighbour merge and the shading
      auto first = [&](dim3 grid, int halo) {
        launchDirectStage(P, grid, st, r->countVisits, tex, true, halo);
        r->stats.kernelLaunches[EID_K_DIRECT]++;
      };
      first(g, 0);
      if (P.sFirst > 0 || P.sCount > 1 || P.sFirst + P.sRows < P.st.size.y) first(dim3((P.st.size.x + 63) / 64, 2 * P.sCount), 1);
      launchDirectSpatial(P, g, st);
      r->stats.kernelLaunches[EID_K_DIRECT]++;
    }
  }
  if (mark) markStop(r, EID_K_DIRECT, st);
}

static void traceQueue(eid_renderer* r, bool any, const FrameParams& P, const float4* rays, const uint32_t* count, uint32_t* cursor, cudaStream_t st) {
  const int g = r->traceBlocks > 0 ? r->traceBlocks : r->smCount * EID_TQ_MIN_BLOCKS;
  launchTraceQueue(any, r->countVisits, g, st, P.accel, rays, count, cursor, P.wv.hitQ, P.wv.occl, P.counters, P.totals);
  r->stats.kernelLaunches[EID_K_INDIRECT]++;
}

static dim3 indirectGrid(const FrameParams& P) {
  // 8 x 8 quarter-res tiles on ABSOLUTE tile rows: a band that starts inside a tile row gets one more (masked) block row
  return dim3((P.st.size.x / 2 + 7) / 8, P.sCount * ((((P.sFirst / 2) & 7) + P.sRows / 2 + 7) / 8));
}
static bool indirectHasWork(const FrameParams& P) { return P.sCount > 0 && P.st.size.x / 2 > 0 && P.st.size.y / 2 > 0; }
bool indirectIsWavefront(eid_renderer* r, const FrameParams& P) {
  const dim3 g = indirectGrid(P);
  return indirectHasWork(P) && P.wv.slots && (size_t)g.x * g.y * 64 <= P.wv.slots;
}

// wavefront form: begin, then per depth (closest-hit queue, bounce); the shadow queue a bounce fills is traced on the
// `shadow` stream while the main stream goes on with the next depth (small queues are latency-bound: the longest ray
// of the C3 scene needs ~150 dependent node visits, ~80 us, however few rays there are); join before finish
void stageIndirectTrace(eid_renderer* r, const FrameParams& P, cudaStream_t st, int ctx) {
  const dim3 g = indirectGrid(P);
  const bool tex = r->scene->host.hasTextures || r->scene->host.hasNonOpaque || P.env.sunSky.in_use == 1 || P.accel.twoLevel;
  cudaStream_t shadow = ctx ? r->shadowStream2 : r->shadowStream;
  cudaEvent_t evWave = ctx ? r->evWave2 : r->evWave, evJoin = ctx ? r->evWaveJoin2 : r->evWaveJoin;
  cudaStream_t sh = r->waveOverlap ? shadow : st;
  CUDA_CHECK(cudaMemsetAsync(P.wv.ctr, 0, 128 * sizeof(uint32_t), st));
  launchGiBegin(P, g, st, tex);
  r->stats.kernelLaunches[EID_K_INDIRECT]++;
  const int gb = r->smCount * 8;
  bool forked = false;
  for (int d = 1; d <= P.st.maxDepth; ++d) {
    traceQueue(r, false, P, P.wv.rayQ[d & 1], P.wv.ctr + d, P.wv.ctr + 64 + d, st);
    launchGiBounce(P, gb, st, tex, d);
    r->stats.kernelLaunches[EID_K_INDIRECT]++;
    if (d + 1 <= P.st.maxDepth && P.st.MIS > 0) {
      if (sh != st) { CUDA_CHECK(cudaEventRecord(evWave, st)); CUDA_CHECK(cudaStreamWaitEvent(sh, evWave, 0)); forked = true; }
      traceQueue(r, true, P, P.wv.shadowQ + 2 * (size_t)(d - 1) * P.wv.slots, P.wv.ctr + 32 + d - 1, P.wv.ctr + 96 + d - 1, sh);
    }
  }
  if (forked) { CUDA_CHECK(cudaEventRecord(evJoin, sh)); CUDA_CHECK(cudaStreamWaitEvent(st, evJoin, 0)); }
}
void stageIndirectFinish(eid_renderer* r, const FrameParams& P, cudaStream_t st) {
  launchGiFinish(P, indirectGrid(P), st);
  r->stats.kernelLaunches[EID_K_INDIRECT]++;
}

void stageIndirect(eid_renderer* r, const FrameParams& P, cudaStream_t st) {
  markStart(r, EID_K_INDIRECT, st);
  if (indirectHasWork(P)) {
    if (indirectIsWavefront(r, P)) {
      stageIndirectTrace(r, P, st, 0);
      stageIndirectFinish(r, P, st);
    } else {
      const bool tex = r->scene->host.hasTextures || r->scene->host.hasNonOpaque || P.env.sunSky.in_use == 1 || P.accel.twoLevel;
      launchIndirectMega(P, indirectGrid(P), st, r->countVisits, tex);
      r->stats.kernelLaunches[EID_K_INDIRECT]++;
    }
  }
  markStop(r, EID_K_INDIRECT, st);
}

// Row layout of the post stages.  sharded = false: the whole frame (single GPU, or the replicated post of multi-GPU mode A).
// sharded = true (multi-GPU mode B): only what this rank's stripes of the FINAL images need.  An A-Trous level l reaches 2*2^l
// rows, so level l is evaluated on the stripe plus the summed reach of the levels after it (direct: 28/24/16/0 rows for levels
// 0..3; indirect: 60/56/48/32/0 quarter-res rows for levels 0..4); the inputs of level 0 (pre-denoise images, G-buffer) are
// complete on every rank after the first exchange step.  Values are identical to the full-frame evaluation, only the evaluated
// row ranges shrink (overlapping ranges of neighbouring stripes recompute identical values).
PostLayout postLayout(const FrameParams& P, bool sharded) {
  PostLayout L;
  L.sharded = sharded;
  L.first = sharded ? P.sFirst : 0; L.stride = sharded ? P.sStride : (1 << 20);
  L.srows = sharded ? P.sRows : (P.st.size.y + 15) / 16 * 16; L.count = sharded ? P.sCount : 1;
  return L;
}
static inline unsigned gridRows(const PostLayout& L, int rows, int bh) { return (unsigned)(L.count * ((rows + bh - 1) / bh)); }

// one A-Trous level; `rowBlock` (default 4) selects the row-blocked kernel, 1 the one-pixel-per-thread kernel
template <bool INDIRECT>
static void launchDenoise(eid_renderer* r, const FrameParams& P, const float4* src, float4* dst, int level, int lastLevel, int width,
                          const PostLayout& L, int first, int stride, int rows, cudaStream_t st);

static bool fastSigmas(const RtxState& st);
// geometry planes: +-30 full-res rows for K3, +-62 quarter-res rows (= 124 full-res rows) for K4; accounted to the direct denoiser
void stagePrep(eid_renderer* r, const FrameParams& P, const PostLayout& L, cudaStream_t st) {
  if (P.st.denoise <= 0 || L.count <= 0) return;
  const int rows = L.srows + 2 * 124, W = P.st.size.x;
  dim3 g((W + 31) / 32, gridRows(L, rows, 8));
  const bool fast = r->denoiseTiles != 0 && !r->strictMath && fastSigmas(P.st);     // pre-scaled planes feed the tile kernel's fast path only
  launchDenoisePrep(P, g, st, L.first - 124, L.stride, rows, fast && !(P.variant & EID_VARIANT_DIRECT_BILATERAL), fast && !(P.variant & EID_VARIANT_INDIRECT_BILATERAL));
  r->stats.kernelLaunches[EID_K_DENOISE_DIRECT]++;
}

// ---- lattice-view tensor maps of the A-Trous tile kernel (stage_denoise.cuh) ------------------------------------------------------
// cuTensorMapEncodeTiled is a driver-API entry point; it is resolved through the runtime so that libeidola.so does not link libcuda.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encodeTiled() {
  static EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) { cudaGetLastError(); f = nullptr; }
    return (EncodeTiledFn)f;
  }();
  return fn;
}
// View of a pitch-linear float4 image (pitch P texels, `rows` rows) as the lattice of A-Trous level l (s = 2^l): 5-D tensor of floats
// (component 4, phase x s, lattice x ceil(P / s), phase y s, lattice y ceil(rows / s)), box = (4, 1, 36, 1, tileH) = one phase's tile + halo.
static const CUtensorMap& tensorMapFor(eid_renderer* r, const void* base, int pitch, int rows, int level, int tileH) {
  const auto key = std::make_tuple(base, pitch, rows, level, tileH);
  auto it = r->tmaps.find(key);
  if (it != r->tmaps.end()) return it->second;
  EncodeTiledFn enc = encodeTiled();
  if (!enc) raise(EID_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver (TMA needs sm_90+ and a CUDA 12 driver)");
  const cuuint64_t s = 1ull << level;
  const cuuint64_t dims[5] = {4, s, ((cuuint64_t)pitch + s - 1) / s, s, ((cuuint64_t)rows + s - 1) / s};
  const cuuint64_t strides[4] = {16, 16 * s, 16ull * pitch, 16ull * pitch * s};
  const cuuint32_t box[5] = {4, 1, EID_TILE_PW, 1, (cuuint32_t)tileH};
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  alignas(64) CUtensorMap m;
  const CUresult e = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (e != CUDA_SUCCESS) raise(EID_ERR_CUDA, "cuTensorMapEncodeTiled failed with %d (pitch %d rows %d level %d)", (int)e, pitch, rows, level);
  return r->tmaps.emplace(key, m).first->second;
}

// the fast path pre-scales by sqrt(log2e / sigma): needs finite positive sigmas (anything else runs the reference arithmetic)
static bool fastSigmas(const RtxState& st) {
  const float v[6] = {st.sigLuminDirect, st.sigNormalDirect, st.sigDepthDirect, st.sigLuminIndirect, st.sigNormalIndirect, st.sigDepthIndirect};
  for (float f : v) if (!(f > 1e-12f && f < 1e12f)) return false;
  return true;
}

template <bool INDIRECT>
static void launchDenoise(eid_renderer* r, const FrameParams& P, const float4* src, float4* dst, int level, int lastLevel, int width,
                          const PostLayout& L, int first, int stride, int rows, cudaStream_t st) {
  if (r->denoiseTiles != 0) {
    const int R = r->denoiseTileRows, TR = 4 * R, TH = TR + 4, s = 1 << level;
    const bool strict = r->strictMath || !fastSigmas(P.st);
    const int gPitch = INDIRECT ? P.pitch / 2 : P.pitch, gRows = INDIRECT ? P.allocH / 2 : P.allocH;
    const float4* gPos = INDIRECT ? P.geomPosH : P.geomPos;
    const float4* gNrm = INDIRECT ? P.geomNrmH : P.geomNrm;
    AtrousArgs a;
    a.gPos = gPos; a.gNrm = gNrm; a.inImg = src; a.outImg = dst;
    a.gPitch = gPitch; a.iPitch = P.pitch; a.allocRows = gRows;
    a.level = level; a.lastLevel = lastLevel; a.first = first; a.stride = stride; a.rows = rows;
    a.nTy = ((rows - 1) >> level) / TR + 2;                       // upper bound for any stripe alignment; surplus blocks exit at once
    if (L.count == 1) {                                           // exact for the single-stripe layouts
      const int bh = INDIRECT ? P.st.size.y / 2 : P.st.size.y, ylo = std::max(first, 0), yhi = std::min(first + rows, bh);
      if (yhi <= ylo) return;
      a.nTy = ((yhi - 1) >> level) / TR - (ylo >> level) / TR + 1;
    }
    a.useTma = r->denoiseTiles == 1;
    static const CUtensorMap none{};
    const CUtensorMap& mp = a.useTma ? tensorMapFor(r, gPos, gPitch, gRows, level, TH) : none;
    const CUtensorMap& mn = a.useTma ? tensorMapFor(r, gNrm, gPitch, gRows, level, TH) : none;
    const CUtensorMap& mc = a.useTma ? tensorMapFor(r, src, P.pitch, P.allocH, level, TH) : none;
    const int latticeW = (width + s - 1) >> level;
    dim3 g((unsigned)(((latticeW + EID_TILE_W - 1) / EID_TILE_W) << level), (unsigned)((L.count * a.nTy) << level));
    launchAtrousTile(INDIRECT, strict, R, g, st, P, mp, mn, mc, a);
    return;
  }
  const int R = r->denoiseRowBlock;
  const int step = 1 << level, vrows = step * ((((rows + step - 1) >> level) + R - 1) / R);
  dim3 g((width + 31) / 32, gridRows(L, vrows, 4));
  eid::launchDenoise(INDIRECT, r->strictMath, R, g, st, P, src, dst, level, lastLevel, first, stride, rows);
}

void stageDenoiseDirect(eid_renderer* r, const FrameParams& P, const PostLayout& L, cudaStream_t st) {   // renderer.cpp:178-189
  if (P.st.denoise > 0 && L.count > 0 && (P.variant & EID_VARIANT_DIRECT_BILATERAL)) {   // ONE pass: denoiseDirTempA -> thisDirect (renderer.cpp:186-188)
    dim3 g((P.st.size.x + 31) / 32, gridRows(L, L.srows, 4));
    launchBilateral(false, r->strictMath, g, st, P, P.dirA, P.directImg, L.first, L.stride, L.srows);
    r->stats.kernelLaunches[EID_K_DENOISE_DIRECT]++;
  } else
  if (P.st.denoise > 0 && L.count > 0) {   // thisDirect -> A -> B -> A -> thisDirect
    const int W = P.st.size.x;
    const float4* src[4] = {P.directImg, P.dirA, P.dirB, P.dirA};
    float4* dst[4] = {P.dirA, P.dirB, P.dirA, P.directImg};
    const int halo[4] = {28, 24, 16, 0};
    for (int i = 0; i < 4; ++i) {
      const int h = L.sharded ? halo[i] : 0, rows = L.srows + 2 * h;
      launchDenoise<false>(r, P, src[i], dst[i], i, 3, W, L, L.first - h, L.stride, rows, st);
      r->stats.kernelLaunches[EID_K_DENOISE_DIRECT]++;
    }
  }
}

void stageDenoiseIndirect(eid_renderer* r, const FrameParams& P, const PostLayout& L, cudaStream_t st) {   // renderer.cpp:191-202
  const int Wi = P.st.size.x / 2, Hi = P.st.size.y / 2;
  if (P.st.denoise > 0 && Wi > 0 && Hi > 0 && L.count > 0 && (P.variant & EID_VARIANT_INDIRECT_BILATERAL)) {   // ONE pass: IndA -> IndB
    dim3 g((Wi + 31) / 32, gridRows(L, L.srows / 2, 4));
    launchBilateral(true, r->strictMath, g, st, P, P.indA, P.indB, L.first / 2, L.stride / 2, L.srows / 2);
    r->stats.kernelLaunches[EID_K_DENOISE_INDIRECT]++;
  } else
  if (P.st.denoise > 0 && Wi > 0 && Hi > 0 && L.count > 0) {   // IndA -> IndB -> IndA -> thisIndirect -> IndA -> IndB
    const float4* src[5] = {P.indIn, P.indB, P.indA, P.indirectImg, P.indA};
    float4* dst[5] = {P.indB, P.indA, P.indirectImg, P.indA, P.indB};
    const int halo[5] = {60, 56, 48, 32, 0};
    for (int i = 0; i < 5; ++i) {
      const int h = L.sharded ? halo[i] : 0, rows = L.srows / 2 + 2 * h;
      launchDenoise<true>(r, P, src[i], dst[i], i, 4, Wi, L, L.first / 2 - h, L.stride / 2, rows, st);
      r->stats.kernelLaunches[EID_K_DENOISE_INDIRECT]++;
    }
  }
}

void stageCompose(eid_renderer* r, const FrameParams& P, const PostLayout& L, cudaStream_t st) {
  markStart(r, EID_K_COMPOSE, st);
  if (L.count > 0) {
    dim3 g((P.st.size.x + 31) / 32, gridRows(L, L.srows, 8));
    launchCompose(P, g, st, P.st.denoise > 0 ? P.indB : P.indIn, L.first, L.stride, L.srows);
    r->stats.kernelLaunches[EID_K_COMPOSE]++;
  }
  markStop(r, EID_K_COMPOSE, st);
}

void endFrame(eid_renderer* r) {
  CUDA_CHECK(cudaMemcpyAsync(r->countersHost, r->counters + EID_NUM_COUNTERS * r->lastSet, 5 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, r->stream));
  CUDA_CHECK(cudaMemcpyAsync(r->countersHost + 5, r->counters + 2 * EID_NUM_COUNTERS + 5, (EID_NUM_TOTALS - 5) * sizeof(unsigned long long), cudaMemcpyDeviceToHost, r->stream));
  r->statsPending = true;
  CUDA_CHECK(cudaGetLastError());
}

// run `work` on the auxiliary stream between a fork after everything enqueued so far on the main stream and a later join
static void forkAux(eid_renderer* r) { CUDA_CHECK(cudaEventRecord(r->evFork, r->stream)); CUDA_CHECK(cudaStreamWaitEvent(r->aux, r->evFork, 0)); }
static void joinAux(eid_renderer* r) { CUDA_CHECK(cudaEventRecord(r->evJoin, r->aux)); CUDA_CHECK(cudaStreamWaitEvent(r->stream, r->evJoin, 0)); }

// The reference's schedule is K1, K2, K3x4, K4x5, K5 in strict order (renderer.cpp:154-206), but its true dependencies are
// K1 -> {K2, K3}, K2 -> K4, {K3, K4} -> K5.  K2 (quarter-res, long dependent chains, few threads) leaves most issue slots
// idle, so with `overlap` the direct denoiser K3 runs on a second stream concurrently with K2 + K4.  Results are identical.
static void launchTrace(eid_renderer* r, const FrameParams& P) {
  beginFrame(r);
  stageDirect(r, P, r->stream);
  stageIndirect(r, P, r->stream);
}

void launchPost(eid_renderer* r, const FrameParams& P, bool sharded) {
  const PostLayout L = postLayout(P, sharded);
  if (r->profiling) CUDA_CHECK(cudaEventRecord(r->evPost, r->stream));   // everything between the trace stages and here = exchange
  r->postStarted = true;
  cudaStream_t sd = r->overlap ? r->aux : r->stream;
  markStart(r, EID_K_DENOISE_DIRECT, r->stream);
  stagePrep(r, P, L, r->stream);
  if (r->overlap) forkAux(r);
  stageDenoiseDirect(r, P, L, sd);
  markStop(r, EID_K_DENOISE_DIRECT, sd);
  markStart(r, EID_K_DENOISE_INDIRECT, r->stream);
  stageDenoiseIndirect(r, P, L, r->stream);
  markStop(r, EID_K_DENOISE_INDIRECT, r->stream);
  if (r->overlap) joinAux(r);
  stageCompose(r, P, L, r->stream);
  endFrame(r);
}

void strictOrder(eid_renderer* r) {
  if (!r->pipeline) return;
  CUDA_CHECK(cudaEventRecord(r->evOrder, r->k1Stream));
  CUDA_CHECK(cudaStreamWaitEvent(r->stream, r->evOrder, 0));
  r->k1MustWaitStream = true;
}

static void markFrameDone(eid_renderer* r) {     // everything enqueued so far on the render stream belongs to the frame of parity lastSet
  if (!r->pipeline) return;
  CUDA_CHECK(cudaEventRecord(r->evFrameDone2[r->lastSet], r->stream));
  r->frameDoneValid[r->lastSet] = true;
}

static void launchFrame(eid_renderer* r, const FrameParams& P) {
  if (!r->overlap) { strictOrder(r); launchTrace(r, P); launchPost(r, P, false); return; }
  const PostLayout L = postLayout(P, false);
  // Frames in flight: direct_stage goes to its own stream and only waits for the frame before last (which used this parity's G-buffer,
  // direct image, K2 gather buffers and counters); indirect_stage / denoise / compose follow it on the render stream.  K1 of frame f + 1
  // therefore overlaps K2 .. K5 of frame f.  Nothing K1 writes is read by a later stage of the PREVIOUS frame (see FrameParams::k2G).
  cudaStream_t k1 = r->pipeline ? r->k1Stream : r->stream;
  if (r->pipeline && r->k1MustWaitStream) {          // strictly ordered stages ran on the render stream in between: K1 follows them
    CUDA_CHECK(cudaEventRecord(r->evOrder, r->stream)); CUDA_CHECK(cudaStreamWaitEvent(k1, r->evOrder, 0));
    r->k1MustWaitStream = false;
  }
  if (r->pipeline && r->frameDoneValid[r->lastSet]) CUDA_CHECK(cudaStreamWaitEvent(k1, r->evFrameDone2[r->lastSet], 0));
  beginFrame(r, k1);
  stageDirect(r, P, k1);
  if (r->pipeline) { CUDA_CHECK(cudaEventRecord(r->evK1Done, k1)); CUDA_CHECK(cudaStreamWaitEvent(r->stream, r->evK1Done, 0)); }
  markStart(r, EID_K_DENOISE_DIRECT, r->stream);
  stagePrep(r, P, L, r->stream);
  forkAux(r);
  stageDenoiseDirect(r, P, L, r->aux);
  markStop(r, EID_K_DENOISE_DIRECT, r->aux);
  stageIndirect(r, P, r->stream);
  markStart(r, EID_K_DENOISE_INDIRECT, r->stream);
  stageDenoiseIndirect(r, P, L, r->stream);
  markStop(r, EID_K_DENOISE_INDIRECT, r->stream);
  joinAux(r);
  stageCompose(r, P, L, r->stream);
  endFrame(r);
  markFrameDone(r);
}

void* bufferPtr(eid_renderer* r, int which, size_t& bytes) {
  const size_t n = (size_t)r->width * r->height, ni = (size_t)(r->width / 2) * (r->height / 2);
  const int set = r->lastSet;
  switch (which) {
    case EID_BUF_THIS_GBUFFER: bytes = n * 16; return r->gbuffer[!set];
    case EID_BUF_LAST_GBUFFER: bytes = n * 16; return r->gbuffer[set];
    case EID_BUF_MOTION: bytes = n * 4; return r->motion;
    case EID_BUF_THIS_DIRECT_RESV: bytes = n * sizeof(DirectReservoir); return r->directResv[!set];
    case EID_BUF_LAST_DIRECT_RESV: bytes = n * sizeof(DirectReservoir); return r->directResv[set];
    case EID_BUF_THIS_INDIRECT_RESV: bytes = ni * sizeof(IndirectReservoir); return r->indirectResv[!set];
    case EID_BUF_LAST_INDIRECT_RESV: bytes = ni * sizeof(IndirectReservoir); return r->indirectResv[set];
    case EID_BUF_DIRECT: bytes = n * 16; return r->directImg;
    case EID_BUF_INDIRECT: bytes = n * 16; return r->indirectImg;
    case EID_BUF_DENOISE_DIR_A: case EID_BUF_DENOISE_DIR_B: case EID_BUF_DENOISE_IND_A: case EID_BUF_DENOISE_IND_B:
      bytes = n * 16; return r->denoiseTemp[which - EID_BUF_DENOISE_DIR_A];
    case EID_BUF_DISPLAY_F32: bytes = n * 16; return r->displayF;
    case EID_BUF_DISPLAY_RGBA8: bytes = n * 4; return r->display8;
    case EID_BUF_TEMP_DIRECT_RESV: bytes = n * sizeof(DirectReservoir); return r->tempDirectResv;
    default: return nullptr;
  }
}

extern "C" {

static void envRelease(eid_env* e);

// L2 residency of the acceleration structure.  A frame streams ~1 GB of screen-space buffers through the 126 MB L2, which keeps evicting the
// BVH (ncu: 30-40 % of the trace kernels' L2 requests miss, and every miss is a DRAM round trip inside a dependent node walk).  An access-policy
// window on the kernels' streams marks the node array (mode 1), the triangle array (2) or the span of both (3, when they are close enough) as
// persisting in the L2 set-aside.  EIDOLA_L2_PERSIST=0 disables it.
static void applyL2Policy(eid_renderer* r) {
  const char* e = getenv("EIDOLA_L2_PERSIST");
  const int mode = e ? atoi(e) : EID_L2_PERSIST_DEFAULT;
  r->l2PersistBytes = 0;
  if (mode <= 0 || !r->accel) return;
  int maxPersist = 0, maxWin = 0;
  if (cudaDeviceGetAttribute(&maxPersist, cudaDevAttrMaxPersistingL2CacheSize, r->device) != cudaSuccess ||
      cudaDeviceGetAttribute(&maxWin, cudaDevAttrMaxAccessPolicyWindowSize, r->device) != cudaSuccess || maxPersist <= 0 || maxWin <= 0) { cudaGetLastError(); return; }
  const char* nodes = (const char*)r->accel->nodes; const size_t nodeBytes = (size_t)std::max(r->accel->nodeAlloc, 1u) * EID_NODE_BYTES;
  const char* tris = (const char*)r->accel->tris; const size_t triBytes = (size_t)r->accel->triCount * 48;
  const char* base = nodes; size_t bytes = nodeBytes, useful = nodeBytes;
  if (mode == 2) { base = tris; bytes = useful = triBytes; }
  if (mode == 3) {
    const char* lo = std::min(nodes, tris); const char* hi = std::max(nodes + nodeBytes, tris + triBytes);
    if ((size_t)(hi - lo) <= (size_t)maxWin && (size_t)(hi - lo) <= 2 * (nodeBytes + triBytes)) { base = lo; bytes = (size_t)(hi - lo); useful = nodeBytes + triBytes; }
  }
  if (!base || !bytes) return;
  bytes = std::min(bytes, (size_t)maxWin);
  const size_t setAside = std::min(useful, (size_t)maxPersist);
  if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, setAside) != cudaSuccess) { cudaGetLastError(); return; }
  cudaStreamAttrValue v;
  memset(&v, 0, sizeof(v));
  v.accessPolicyWindow.base_ptr = (void*)base;
  v.accessPolicyWindow.num_bytes = bytes;
  v.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)setAside / (double)bytes);
  v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
  v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
  for (cudaStream_t st : {r->stream, r->aux, r->shadowStream, r->shadowStream2, r->k1Stream})
    if (st && cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &v) != cudaSuccess) { cudaGetLastError(); return; }
  r->l2PersistBytes = setAside;
}

int eid_renderer_create(eid_renderer** out, eid_scene* s, eid_accel* a, uint32_t width, uint32_t height, void* cuda_stream) {
  EID_TRY
  if (!out || !s || !a) raise(EID_ERR_INVALID, "eid_renderer_create: null argument");
  if (!s->loaded || a->scene != s) raise(EID_ERR_STATE, "eid_renderer_create: scene not loaded or accel built for another scene");
  if (!width || !height || width > 32768 || height > 32768) raise(EID_ERR_INVALID, "bad render size %ux%u", width, height);
  eid_renderer* r = new eid_renderer();
  try {
    r->scene = s; r->accel = a; r->device = s->dev.device; r->width = width; r->height = height;
    CUDA_CHECK(cudaSetDevice(r->device));
    CUDA_CHECK(cudaDeviceGetAttribute(&r->smCount, cudaDevAttrMultiProcessorCount, r->device));
    if (cuda_stream) r->stream = (cudaStream_t)cuda_stream;
    else { CUDA_CHECK(cudaStreamCreateWithFlags(&r->stream, cudaStreamNonBlocking)); r->ownStream = true; }
    CUDA_CHECK(cudaMalloc(&r->counters, (2 * EID_NUM_COUNTERS + EID_NUM_TOTALS) * sizeof(unsigned long long)));
    CUDA_CHECK(cudaMemset(r->counters, 0, (2 * EID_NUM_COUNTERS + EID_NUM_TOTALS) * sizeof(unsigned long long)));
    CUDA_CHECK(cudaMallocHost(&r->countersHost, EID_NUM_TOTALS * sizeof(unsigned long long)));
    memset(r->countersHost, 0, EID_NUM_TOTALS * sizeof(unsigned long long));
    for (auto& e : r->ev) CUDA_CHECK(cudaEventCreate(&e));
    CUDA_CHECK(cudaEventCreateWithFlags(&r->evFork, cudaEventDisableTiming)); CUDA_CHECK(cudaEventCreateWithFlags(&r->evJoin, cudaEventDisableTiming));
    CUDA_CHECK(cudaEventCreate(&r->evPost));
    CUDA_CHECK(cudaStreamCreateWithFlags(&r->aux, cudaStreamNonBlocking));
    CUDA_CHECK(cudaStreamCreateWithFlags(&r->shadowStream, cudaStreamNonBlocking));
    {   // the K1 stream has the LOWEST priority: with frames in flight, the next frame's direct stage fills what the latency-bound
        // chain of the current frame's indirect stage leaves idle, never the other way round
      int lo = 0, hi = 0;
      CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      CUDA_CHECK(cudaStreamCreateWithPriority(&r->k1Stream, cudaStreamNonBlocking, lo));
    }
    CUDA_CHECK(cudaEventCreateWithFlags(&r->evK1Done, cudaEventDisableTiming));
    CUDA_CHECK(cudaEventCreateWithFlags(&r->evOrder, cudaEventDisableTiming));
    for (auto& e : r->evFrameDone2) CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    CUDA_CHECK(cudaEventCreateWithFlags(&r->evWave, cudaEventDisableTiming)); CUDA_CHECK(cudaEventCreateWithFlags(&r->evWaveJoin, cudaEventDisableTiming));
    r->allocate();
    applyL2Policy(r);
  } catch (...) { eid_renderer_destroy(r); throw; }
  *out = r;
  return EID_OK;
  EID_CATCH
}

int eid_renderer_resize(eid_renderer* r, uint32_t width, uint32_t height) {
  EID_TRY
  if (!r) raise(EID_ERR_INVALID, "eid_renderer_resize: null renderer");
  if (!width || !height || width > 32768 || height > 32768) raise(EID_ERR_INVALID, "bad render size %ux%u", width, height);
  CUDA_CHECK(cudaSetDevice(r->device));
  CUDA_CHECK(cudaStreamSynchronize(r->stream));
  if (r->copyStream) CUDA_CHECK(cudaStreamSynchronize(r->copyStream));
  cudaFree(r->staging[0]); cudaFree(r->staging[1]); r->staging[0] = r->staging[1] = nullptr; r->copyPending = false;
  r->release();
  r->width = width; r->height = height; r->stripesSet = false;
  r->allocate();
  return EID_OK;
  EID_CATCH
}

void eid_renderer_destroy(eid_renderer* r) {
  if (!r) return;
  cudaSetDevice(r->device);
  if (r->stream) cudaStreamSynchronize(r->stream);
  if (r->envMap) { envRelease(r->envMap); r->envMap = nullptr; }
  r->release();
  cudaFree(r->counters);
  if (r->countersHost) cudaFreeHost(r->countersHost);
  for (auto& e : r->ev) if (e) cudaEventDestroy(e);
  if (r->evFork) cudaEventDestroy(r->evFork); if (r->evJoin) cudaEventDestroy(r->evJoin); if (r->evPost) cudaEventDestroy(r->evPost);
  if (r->aux) { cudaStreamSynchronize(r->aux); cudaStreamDestroy(r->aux); }
  if (r->shadowStream) { cudaStreamSynchronize(r->shadowStream); cudaStreamDestroy(r->shadowStream); }
  if (r->shadowStream2) { cudaStreamSynchronize(r->shadowStream2); cudaStreamDestroy(r->shadowStream2); }
  if (r->evWave2) cudaEventDestroy(r->evWave2);
  if (r->evWaveJoin2) cudaEventDestroy(r->evWaveJoin2);
  if (r->k1Stream) { cudaStreamSynchronize(r->k1Stream); cudaStreamDestroy(r->k1Stream); }
  if (r->evK1Done) cudaEventDestroy(r->evK1Done);
  if (r->evOrder) cudaEventDestroy(r->evOrder);
  for (auto& e : r->evFrameDone2) if (e) cudaEventDestroy(e);
  if (r->evWave) cudaEventDestroy(r->evWave); if (r->evWaveJoin) cudaEventDestroy(r->evWaveJoin);
  if (r->copyStream) { cudaStreamSynchronize(r->copyStream); cudaStreamDestroy(r->copyStream); }
  if (r->evFrameDone) cudaEventDestroy(r->evFrameDone); if (r->evCopyDone) cudaEventDestroy(r->evCopyDone);
  for (auto& e : r->evCopyDone2) if (e) cudaEventDestroy(e);
  cudaFree(r->staging[0]); cudaFree(r->staging[1]);
  if (r->ownStream && r->stream) cudaStreamDestroy(r->stream);
  delete r;
}

// ---- HdrSampling (hdr_sampling.hpp:43-48) ---------------------------------------------------------------------------------
static int envUpload(eid_env* e) {
  CUDA_CHECK(cudaSetDevice(e->device));
  const size_t n = (size_t)e->host.width * e->host.height;
  CUDA_CHECK(cudaMalloc(&e->tex, n * 16));
  CUDA_CHECK(cudaMalloc(&e->accel, n * sizeof(ImptSampData)));
  CUDA_CHECK(cudaMemcpy(e->tex, e->host.pixels.data(), n * 16, cudaMemcpyHostToDevice));
  CUDA_CHECK(cudaMemcpy(e->accel, e->host.accel.data(), n * sizeof(ImptSampData), cudaMemcpyHostToDevice));
  return EID_OK;
}

int eid_env_create(eid_env** out, int device, const float* rgba, uint32_t width, uint32_t height) {
  EID_TRY
  if (!out) raise(EID_ERR_INVALID, "eid_env_create: out is null");
  eid_env* e = new eid_env();
  try {
    e->device = device;
    e->host.build(rgba, width, height);
    if (device != EID_DEVICE_NONE) envUpload(e);
  } catch (...) { eid_env_destroy(e); throw; }
  *out = e;
  return EID_OK;
  EID_CATCH
}

int eid_env_load_hdr(eid_env** out, int device, const char* path) {
  EID_TRY
  if (!out || !path) raise(EID_ERR_INVALID, "eid_env_load_hdr: null argument");
  eid_env* e = new eid_env();
  try {
    e->device = device;
    e->host.loadRadianceHdr(path);
    if (device != EID_DEVICE_NONE) envUpload(e);
  } catch (...) { eid_env_destroy(e); throw; }
  *out = e;
  return EID_OK;
  EID_CATCH
}

static void envFree(eid_env* e);
void eid_env_destroy(eid_env* e) {
  if (!e) return;
  if (e->users > 0) { e->destroyRequested = true; return; }     // a renderer still samples it: freed when the last one lets go
  envFree(e);
}
static void envRelease(eid_env* e) {                             // a renderer lets go of `e`
  if (e && --e->users <= 0 && e->destroyRequested) envFree(e);
}
static void envFree(eid_env* e) {
  if (e->tex || e->accel) { cudaSetDevice(e->device); cudaFree(e->tex); cudaFree(e->accel); }
  delete e;
}

float eid_env_integral(eid_env* e) { return e ? e->host.integral : 0.f; }
float eid_env_average(eid_env* e) { return e ? e->host.average : 0.f; }

int eid_env_get_size(eid_env* e, uint32_t* width, uint32_t* height) {
  EID_TRY
  if (!e || !width || !height) raise(EID_ERR_INVALID, "eid_env_get_size: null argument");
  *width = e->host.width; *height = e->host.height;
  return EID_OK;
  EID_CATCH
}

int eid_env_read(eid_env* e, int what, void* dst, size_t bytes) {
  EID_TRY
  if (!e || !dst) raise(EID_ERR_INVALID, "eid_env_read: null argument");
  const size_t n = (size_t)e->host.width * e->host.height;
  const void* src = what == 0 ? (const void*)e->host.accel.data() : (const void*)e->host.pixels.data();
  const size_t total = what == 0 ? n * sizeof(ImptSampData) : n * 16;
  if (what != 0 && what != 1) raise(EID_ERR_INVALID, "eid_env_read: what must be 0 (alias table) or 1 (pixels)");
  if (bytes > total) raise(EID_ERR_INVALID, "read of %zu bytes from %zu", bytes, total);
  memcpy(dst, src, bytes);
  return EID_OK;
  EID_CATCH
}

int eid_renderer_set_env(eid_renderer* r, eid_env* e) {
  EID_TRY
  if (!r) raise(EID_ERR_INVALID, "eid_renderer_set_env: null renderer");
  if (e && (e->device != r->device || !e->tex)) raise(EID_ERR_INVALID, "environment map lives on another device (or is host-only)");
  if (e == r->envMap) return EID_OK;
  if (r->envMap) { CUDA_CHECK(cudaSetDevice(r->device)); CUDA_CHECK(cudaStreamSynchronize(r->stream)); envRelease(r->envMap); }   // frames in flight may still sample it
  if (e) e->users++;
  r->envMap = e;
  return EID_OK;
  EID_CATCH
}

int eid_renderer_set_sun_and_sky(eid_renderer* r, const SunAndSky* ss) {
  EID_TRY
  if (!r || !ss) raise(EID_ERR_INVALID, "eid_renderer_set_sun_and_sky: null argument");
  r->sunSky = *ss;
  return EID_OK;
  EID_CATCH
}

int eid_renderer_run_output(eid_renderer* r, const Tonemapper* tm) {
  EID_TRY
  if (!r || !tm) raise(EID_ERR_INVALID, "eid_renderer_run_output: null argument");
  if (!r->hasRun) raise(EID_ERR_STATE, "eid_renderer_run_output: no frame has been rendered");
  if (tm->autoExposure & 2) raise(EID_ERR_UNSUPPORTED, "post.frag toneLocalExposure (autoExposure bit 1) is never selected by the reference's GUI and reads an uninitialised variable there: not implemented");
  CUDA_CHECK(cudaSetDevice(r->device));
  const size_t n = (size_t)r->width * r->height;
  if (!r->displayF) {
    CUDA_CHECK(cudaMalloc(&r->displayF, n * 16)); CUDA_CHECK(cudaMalloc(&r->display8, n * 4));
    CUDA_CHECK(cudaMemsetAsync(r->displayF, 0, n * 16, r->stream)); CUDA_CHECK(cudaMemsetAsync(r->display8, 0, n * 4, r->stream));
  }
  FrameParams P;
  memset(&P, 0, sizeof(P));
  P.st = r->lastState; P.pitch = (int)r->width; P.allocH = (int)r->height;
  P.directImg = r->directImg; P.indirectImg = r->indirectImg;
  dim3 g((P.st.size.x + 31) / 32, (P.st.size.y + 7) / 8);
  if (tm->autoExposure & 1) {   // RenderOutput::genMipmap (render_output.cpp:243-253): the chain of the whole images (m_size = the allocation) down to 1 x 1
    if (!r->mipScratch) CUDA_CHECK(cudaMalloc(&r->mipScratch, (2 * ((size_t)(r->width / 2 + 1) * (r->height / 2 + 1)) + 2) * 16));
    const size_t half = (size_t)(r->width / 2 + 1) * (r->height / 2 + 1);
    float4* pp[2] = {r->mipScratch, r->mipScratch + half};
    float4* avg = r->mipScratch + 2 * half;
    for (int img = 0; img < 2; ++img) {
      const float4* src = img ? r->indirectImg : r->directImg;
      int sw = (int)r->width, sh = (int)r->height, spitch = (int)r->width, k = 0;
      if (sw == 1 && sh == 1) CUDA_CHECK(cudaMemcpyAsync(avg + img, src, 16, cudaMemcpyDeviceToDevice, r->stream));
      while (sw > 1 || sh > 1) {
        const int dw = sw > 1 ? sw / 2 : 1, dh = sh > 1 ? sh / 2 : 1;
        float4* dst = (dw == 1 && dh == 1) ? avg + img : pp[k & 1];
        launchMipBlit(dim3((dw + 31) / 32, (dh + 7) / 8), r->stream, src, sw, sh, spitch, dst, dw, dh);
        src = dst; sw = dw; sh = dh; spitch = dw; ++k;
      }
    }
    launchPost(P, g, r->stream, *tm, r->displayF, r->display8, avg);
  } else {
    launchPost(P, g, r->stream, *tm, r->displayF, r->display8, nullptr);
  }
  CUDA_CHECK(cudaGetLastError());
  return EID_OK;
  EID_CATCH
}

int eid_renderer_fn_tap(eid_renderer* r, const RtxState* st, int which, const float* in, uint32_t n, float* out) {
  EID_TRY
  static const int A[][2] = {{4, 9}, {0, 0}, {3, 4}, {3, 3}, {4, 6}, {3, 3}, {12, 8}};
  if (!r || !st || !in || !out) raise(EID_ERR_INVALID, "eid_renderer_fn_tap: null argument");
  if (which < 0 || which >= (int)(sizeof(A) / sizeof(A[0])) || A[which][0] == 0) raise(EID_ERR_INVALID, "eid_renderer_fn_tap: no device tap %d", which);
  if (n == 0) return EID_OK;
  const int ni = A[which][0], no = A[which][1];
  CUDA_CHECK(cudaSetDevice(r->device));
  FrameParams P;
  memset(&P, 0, sizeof(P));
  P.st = *st; P.cam = r->scene->host.camera; P.sc = r->scene->dev.view(r->scene->host);
  for (int k = 0; k < 3; ++k) P.env.constant[k] = r->env[k];
  P.env.sunSky = r->sunSky;
  P.env.tex = r->envMap ? r->envMap->tex : nullptr; P.env.accel = r->envMap ? r->envMap->accel : nullptr;
  P.env.width = r->envMap ? (int)r->envMap->host.width : 0; P.env.height = r->envMap ? (int)r->envMap->host.height : 0;
  float *din = nullptr, *dout = nullptr;
  CUDA_CHECK(cudaMalloc(&din, (size_t)n * ni * 4));
  if (cudaMalloc(&dout, (size_t)n * no * 4) != cudaSuccess) { cudaFree(din); raise(EID_ERR_CUDA, "cudaMalloc failed"); }
  cudaError_t e = cudaMemcpy(din, in, (size_t)n * ni * 4, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemset(dout, 0, (size_t)n * no * 4);
  if (e == cudaSuccess) {
    launchCtxTap(P, which, ni, no, din, n, dout);
    e = cudaMemcpy(out, dout, (size_t)n * no * 4, cudaMemcpyDeviceToHost);
  }
  cudaFree(din); cudaFree(dout);
  if (e != cudaSuccess) raise(EID_ERR_CUDA, "eid_renderer_fn_tap: %s", cudaGetErrorString(e));
  return EID_OK;
  EID_CATCH
}

int eid_fn_tap(int device, int which, const float* in, uint32_t n, float* out) {
  EID_TRY
  static const int A[][2] = {{2, 2}, {2, 1}, {3, 2}, {3, 6}, {3, 3}, {3, 3}, {14, 3}, {14, 1}, {14, 7}, {0, 0}, {0, 0}, {4, 3}, {6, 3}, {2, 1}, {2, 3}};
  if (!in || !out) raise(EID_ERR_INVALID, "eid_fn_tap: null argument");
  if (which < 0 || which >= (int)(sizeof(A) / sizeof(A[0])) || A[which][0] == 0) raise(EID_ERR_INVALID, "eid_fn_tap: no device tap %d", which);
  if (n == 0) return EID_OK;
  const int ni = A[which][0], no = A[which][1];
  CUDA_CHECK(cudaSetDevice(device));
  float *din = nullptr, *dout = nullptr;
  CUDA_CHECK(cudaMalloc(&din, (size_t)n * ni * 4));
  if (cudaMalloc(&dout, (size_t)n * no * 4) != cudaSuccess) { cudaFree(din); raise(EID_ERR_CUDA, "cudaMalloc failed"); }
  cudaError_t e = cudaMemcpy(din, in, (size_t)n * ni * 4, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemset(dout, 0, (size_t)n * no * 4);
  if (e == cudaSuccess) {
    launchFnTap(which, ni, no, din, n, dout);
    e = cudaMemcpy(out, dout, (size_t)n * no * 4, cudaMemcpyDeviceToHost);
  }
  cudaFree(din); cudaFree(dout);
  if (e != cudaSuccess) raise(EID_ERR_CUDA, "eid_fn_tap: %s", cudaGetErrorString(e));
  return EID_OK;
  EID_CATCH
}

int eid_sun_and_sky_eval(int device, const SunAndSky* ss, const float* dirs, uint32_t n, float* rgb) {
  EID_TRY
  if (!ss || !dirs || !rgb) raise(EID_ERR_INVALID, "eid_sun_and_sky_eval: null argument");
  if (n == 0) return EID_OK;
  CUDA_CHECK(cudaSetDevice(device));
  float *dd = nullptr, *dout = nullptr;
  CUDA_CHECK(cudaMalloc(&dd, (size_t)n * 12));
  if (cudaMalloc(&dout, (size_t)n * 12) != cudaSuccess) { cudaFree(dd); raise(EID_ERR_CUDA, "cudaMalloc failed"); }
  cudaError_t e = cudaMemcpy(dd, dirs, (size_t)n * 12, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    launchSunAndSky(*ss, dd, n, dout);
    e = cudaMemcpy(rgb, dout, (size_t)n * 12, cudaMemcpyDeviceToHost);
  }
  cudaFree(dd); cudaFree(dout);
  if (e != cudaSuccess) raise(EID_ERR_CUDA, "eid_sun_and_sky_eval: %s", cudaGetErrorString(e));
  return EID_OK;
  EID_CATCH
}

int eid_renderer_set_env_constant(eid_renderer* r, const float rgb[3]) {
  EID_TRY
  if (!r || !rgb) raise(EID_ERR_INVALID, "eid_renderer_set_env_constant: null argument");
  for (int k = 0; k < 3; ++k) r->env[k] = rgb[k];
  return EID_OK;
  EID_CATCH
}

int eid_renderer_run(eid_renderer* r, const RtxState* state, int frames) {
  EID_TRY
  if (!r || !state) raise(EID_ERR_INVALID, "eid_renderer_run: null argument");
  CUDA_CHECK(cudaSetDevice(r->device));
  FrameParams P;
  fillParams(r, *state, frames, P);
  launchFrame(r, P);
  return EID_OK;
  EID_CATCH
}

int eid_renderer_run_trace(eid_renderer* r, const RtxState* state, int frames) {
  EID_TRY
  if (!r || !state) raise(EID_ERR_INVALID, "eid_renderer_run_trace: null argument");
  CUDA_CHECK(cudaSetDevice(r->device));
  FrameParams P;
  strictOrder(r);
  fillParams(r, *state, frames, P);
  launchTrace(r, P);
  CUDA_CHECK(cudaGetLastError());
  return EID_OK;
  EID_CATCH
}

// run_trace split in two so a multi-GPU host can start exchanging the G-buffer + direct image while indirect_stage runs
int eid_renderer_run_direct(eid_renderer* r, const RtxState* state, int frames) {
  EID_TRY
  if (!r || !state) raise(EID_ERR_INVALID, "eid_renderer_run_direct: null argument");
  CUDA_CHECK(cudaSetDevice(r->device));
  FrameParams P;
  strictOrder(r);
  fillParams(r, *state, frames, P);
  beginFrame(r);
  stageDirect(r, P, r->stream);
  CUDA_CHECK(cudaGetLastError());
  return EID_OK;
  EID_CATCH
}

int eid_renderer_run_indirect(eid_renderer* r, const RtxState* state, int frames) {
  EID_TRY
  if (!r || !state) raise(EID_ERR_INVALID, "eid_renderer_run_indirect: null argument");
  CUDA_CHECK(cudaSetDevice(r->device));
  FrameParams P;
  strictOrder(r);
  fillParams(r, *state, frames, P);
  stageIndirect(r, P, r->stream);
  CUDA_CHECK(cudaGetLastError());
  return EID_OK;
  EID_CATCH
}

int eid_renderer_run_post(eid_renderer* r, const RtxState* state, int frames) {
  EID_TRY
  if (!r || !state) raise(EID_ERR_INVALID, "eid_renderer_run_post: null argument");
  CUDA_CHECK(cudaSetDevice(r->device));
  FrameParams P;
  strictOrder(r);
  fillParams(r, *state, frames, P);
  launchPost(r, P, false);
  return EID_OK;
  EID_CATCH
}

int eid_renderer_run_post_band(eid_renderer* r, const RtxState* state, int frames) {
  EID_TRY
  if (!r || !state) raise(EID_ERR_INVALID, "eid_renderer_run_post_band: null argument");
  CUDA_CHECK(cudaSetDevice(r->device));
  FrameParams P;
  strictOrder(r);
  fillParams(r, *state, frames, P);
  launchPost(r, P, true);
  return EID_OK;
  EID_CATCH
}

int eid_renderer_sync(eid_renderer* r) {
  EID_TRY
  if (!r) raise(EID_ERR_INVALID, "eid_renderer_sync: null renderer");
  CUDA_CHECK(cudaSetDevice(r->device));
  CUDA_CHECK(cudaStreamSynchronize(r->stream));   // the aux stream is always joined into the main stream before a frame ends,
  if (r->pipeline) CUDA_CHECK(cudaStreamSynchronize(r->k1Stream));   // ... and every direct_stage is followed by its frame on the main stream
  if (r->groupStream) CUDA_CHECK(cudaStreamSynchronize(r->groupStream));   // exchange C of an eid_group frame
  return EID_OK;
  EID_CATCH
}

int eid_renderer_get_outputs(eid_renderer* r, const float** direct, const float** indirect) {
  EID_TRY
  if (!r) raise(EID_ERR_INVALID, "eid_renderer_get_outputs: null renderer");
  if (direct) *direct = (const float*)r->directImg;
  if (indirect) *indirect = (const float*)r->indirectImg;
  return EID_OK;
  EID_CATCH
}

int64_t eid_renderer_buffer_bytes(eid_renderer* r, int which) {
  if (!r) return -1;
  size_t b = 0;
  return bufferPtr(r, which, b) ? (int64_t)b : -1;
}

int eid_renderer_read(eid_renderer* r, int which, void* host_dst, size_t bytes) {
  EID_TRY
  if (!r || !host_dst) raise(EID_ERR_INVALID, "eid_renderer_read: null argument");
  size_t b = 0; void* p = bufferPtr(r, which, b);
  if (!p) raise(EID_ERR_INVALID, "no such buffer %d", which);
  if (bytes > b) raise(EID_ERR_INVALID, "read of %zu bytes from a %zu-byte buffer", bytes, b);
  CUDA_CHECK(cudaSetDevice(r->device));
  if (r->groupStream) CUDA_CHECK(cudaStreamSynchronize(r->groupStream));
  CUDA_CHECK(cudaMemcpyAsync(host_dst, p, bytes, cudaMemcpyDeviceToHost, r->stream));
  CUDA_CHECK(cudaStreamSynchronize(r->stream));
  return EID_OK;
  EID_CATCH
}

int eid_renderer_write(eid_renderer* r, int which, const void* host_src, size_t bytes) {
  EID_TRY
  if (!r || !host_src) raise(EID_ERR_INVALID, "eid_renderer_write: null argument");
  size_t b = 0; void* p = bufferPtr(r, which, b);
  if (!p) raise(EID_ERR_INVALID, "no such buffer %d", which);
  if (bytes > b) raise(EID_ERR_INVALID, "write of %zu bytes into a %zu-byte buffer", bytes, b);
  CUDA_CHECK(cudaSetDevice(r->device));
  CUDA_CHECK(cudaMemcpyAsync(p, host_src, bytes, cudaMemcpyHostToDevice, r->stream));
  CUDA_CHECK(cudaStreamSynchronize(r->stream));
  return EID_OK;
  EID_CATCH
}

int eid_renderer_render_host(eid_renderer* r, const SceneCamera* cam, const RtxState* state, int frames, float* direct_host, float* indirect_host) {
  EID_TRY
  if (!r || !state) raise(EID_ERR_INVALID, "eid_renderer_render_host: null argument");
  CUDA_CHECK(cudaSetDevice(r->device));
  if (cam) r->scene->host.camera = *cam;
  FrameParams P;
  fillParams(r, *state, frames, P);
  launchFrame(r, P);
  const size_t rowBytes = (size_t)state->size.x * 16;
  if (direct_host) CUDA_CHECK(cudaMemcpy2DAsync(direct_host, rowBytes, r->directImg, (size_t)r->width * 16, rowBytes, state->size.y, cudaMemcpyDeviceToHost, r->stream));
  if (indirect_host) CUDA_CHECK(cudaMemcpy2DAsync(indirect_host, rowBytes, r->indirectImg, (size_t)r->width * 16, rowBytes, state->size.y, cudaMemcpyDeviceToHost, r->stream));
  CUDA_CHECK(cudaStreamSynchronize(r->stream));
  return EID_OK;
  EID_CATCH
}

// Pipelined variant: the frame is enqueued and its two result images — which exist once per ping-pong parity, so the next frame
// renders into the other pair — are copied to the host IN PLACE on a dedicated copy stream while the next frame's kernels run; the
// frame after that (same parity) waits for this copy before it rewrites them.  The host buffers are complete after
// eid_renderer_wait_host.  Use two host buffer pairs alternately.
extern "C++" {
void ensureCopyStream(eid_renderer* r) {
  if (r->copyStream) return;
  CUDA_CHECK(cudaStreamCreateWithFlags(&r->copyStream, cudaStreamNonBlocking));
  CUDA_CHECK(cudaEventCreateWithFlags(&r->evFrameDone, cudaEventDisableTiming));
  for (auto& e : r->evCopyDone2) CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
}
}

int eid_renderer_render_host_async(eid_renderer* r, const SceneCamera* cam, const RtxState* state, int frames, float* direct_host, float* indirect_host) {
  EID_TRY
  if (!r || !state) raise(EID_ERR_INVALID, "eid_renderer_render_host_async: null argument");
  CUDA_CHECK(cudaSetDevice(r->device));
  ensureCopyStream(r);
  if (cam) r->scene->host.camera = *cam;
  FrameParams P;
  fillParams(r, *state, frames, P);                 // (waits for the copy that read this parity's images two frames ago)
  launchFrame(r, P);
  // the result images are per ping-pong parity: the copy stream reads them in place while the next frame renders into the other pair
  const int set = r->lastSet;
  CUDA_CHECK(cudaEventRecord(r->evFrameDone, r->stream));
  CUDA_CHECK(cudaStreamWaitEvent(r->copyStream, r->evFrameDone, 0));
  const size_t rowBytes = (size_t)state->size.x * 16;
  if (direct_host) CUDA_CHECK(cudaMemcpy2DAsync(direct_host, rowBytes, r->directImg, (size_t)r->width * 16, rowBytes, state->size.y, cudaMemcpyDeviceToHost, r->copyStream));
  if (indirect_host) CUDA_CHECK(cudaMemcpy2DAsync(indirect_host, rowBytes, r->indirectImg, (size_t)r->width * 16, rowBytes, state->size.y, cudaMemcpyDeviceToHost, r->copyStream));
  CUDA_CHECK(cudaEventRecord(r->evCopyDone2[set], r->copyStream));
  r->copyPending2[set] = true;
  r->copyPending = true;
  return EID_OK;
  EID_CATCH
}

int eid_renderer_wait_host(eid_renderer* r) {
  EID_TRY
  if (!r) raise(EID_ERR_INVALID, "eid_renderer_wait_host: null renderer");
  CUDA_CHECK(cudaSetDevice(r->device));
  if (r->copyStream) CUDA_CHECK(cudaStreamSynchronize(r->copyStream));
  r->copyPending = false;
  return EID_OK;
  EID_CATCH
}

int eid_renderer_set_strict_math(eid_renderer* r, int enabled) {
  EID_TRY
  if (!r) raise(EID_ERR_INVALID, "eid_renderer_set_strict_math: null renderer");
  r->strictMath = enabled != 0;
  return EID_OK;
  EID_CATCH
}

int eid_renderer_set_denoise_rows(eid_renderer* r, int rowsPerThread) {
  EID_TRY
  if (!r) raise(EID_ERR_INVALID, "eid_renderer_set_denoise_rows: null renderer");
  if (rowsPerThread != 1 && rowsPerThread != 2 && rowsPerThread != 4) raise(EID_ERR_INVALID, "eid_renderer_set_denoise_rows: 1, 2 or 4");
  r->denoiseRowBlock = rowsPerThread;
  return EID_OK;
  EID_CATCH
}

int eid_renderer_set_denoise_tiles(eid_renderer* r, int mode, int rowsPerThread) {
  EID_TRY
  if (!r) raise(EID_ERR_INVALID, "eid_renderer_set_denoise_tiles: null renderer");
  if (mode < 0 || mode > 2) raise(EID_ERR_INVALID, "eid_renderer_set_denoise_tiles: mode 0 (legacy), 1 (TMA tiles) or 2 (cp.async tiles)");
  if (rowsPerThread != 0 && rowsPerThread != 2 && rowsPerThread != 4) raise(EID_ERR_INVALID, "eid_renderer_set_denoise_tiles: rowsPerThread 2 or 4 (0 = keep)");
  r->denoiseTiles = mode;
  if (rowsPerThread) r->denoiseTileRows = rowsPerThread;
  return EID_OK;
  EID_CATCH
}

int eid_renderer_set_wavefront(eid_renderer* r, int enabled, int traceBlocks) {
  EID_TRY
  if (!r) raise(EID_ERR_INVALID, "eid_renderer_set_wavefront: null renderer");
  if (traceBlocks < 0 || traceBlocks > 65535) raise(EID_ERR_INVALID, "eid_renderer_set_wavefront: traceBlocks out of range");
  r->wavefront = enabled != 0;
  r->waveOverlap = enabled != 2;      // 2: wavefront with every queue on the main stream (strictly serial stages)
  r->traceBlocks = traceBlocks;
  return EID_OK;
  EID_CATCH
}

int eid_renderer_set_variant(eid_renderer* r, int flags) {
  EID_TRY
  if (!r) raise(EID_ERR_INVALID, "eid_renderer_set_variant: null renderer");
  if (flags & ~(EID_VARIANT_DIRECT_BILATERAL | EID_VARIANT_INDIRECT_BILATERAL | EID_VARIANT_FETCH_4_SUBPIXELS | EID_VARIANT_DIRECT_SPLIT)) raise(EID_ERR_INVALID, "eid_renderer_set_variant: unknown variant bits 0x%x", flags);
  r->variant = flags;
  return EID_OK;
  EID_CATCH
}

int eid_renderer_set_pipeline(eid_renderer* r, int framesInFlight) {
  EID_TRY
  if (!r) raise(EID_ERR_INVALID, "eid_renderer_set_pipeline: null renderer");
  if (framesInFlight < 1 || framesInFlight > 2) raise(EID_ERR_INVALID, "eid_renderer_set_pipeline: 1 (strict) or 2 frames in flight");
  CUDA_CHECK(cudaSetDevice(r->device));
  CUDA_CHECK(cudaStreamSynchronize(r->k1Stream));
  CUDA_CHECK(cudaStreamSynchronize(r->stream));
  r->pipeline = framesInFlight == 2;
  r->k1MustWaitStream = true;
  r->frameDoneValid[0] = r->frameDoneValid[1] = false;
  return EID_OK;
  EID_CATCH
}

int eid_renderer_set_overlap(eid_renderer* r, int enabled) {
  EID_TRY
  if (!r) raise(EID_ERR_INVALID, "eid_renderer_set_overlap: null renderer");
  r->overlap = enabled != 0;
  return EID_OK;
  EID_CATCH
}

int eid_renderer_set_profiling(eid_renderer* r, int enabled) {
  EID_TRY
  if (!r) raise(EID_ERR_INVALID, "eid_renderer_set_profiling: null renderer");
  r->profiling = enabled != 0;
  r->countVisits = enabled >= 2;
  return EID_OK;
  EID_CATCH
}

int eid_renderer_get_stats(eid_renderer* r, eid_frame_stats* out) {
  EID_TRY
  if (!r || !out) raise(EID_ERR_INVALID, "eid_renderer_get_stats: null argument");
  CUDA_CHECK(cudaSetDevice(r->device));
  CUDA_CHECK(cudaStreamSynchronize(r->stream));
  if (r->statsPending) {
    r->stats.closestHitRays = r->countersHost[0]; r->stats.anyHitRays = r->countersHost[1]; r->stats.primaryHits = r->countersHost[2];
    r->stats.nodeVisits = r->countersHost[3]; r->stats.triangleTests = r->countersHost[4];
    r->stats.totalClosestHitRays = r->countersHost[5]; r->stats.totalAnyHitRays = r->countersHost[6];
    r->stats.maxNodeVisitsPerThread = r->countersHost[7];
    r->stats.maxNodeVisitsPerQueuedRay[0] = r->countersHost[8]; r->stats.maxNodeVisitsPerQueuedRay[1] = r->countersHost[9];
    r->stats.launches = 0;
    for (int k = 0; k < EID_K_COUNT; ++k) r->stats.launches += r->stats.kernelLaunches[k];
    if (r->profiling) {
      CUDA_CHECK(cudaStreamSynchronize(r->aux));
      for (int k = 0; k < EID_K_COUNT; ++k) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r->ev[2 * k], r->ev[2 * k + 1]) == cudaSuccess) r->stats.kernelMs[k] = ms; else cudaGetLastError();
      }
      float ms = 0.f;   // multi-GPU: time between the end of the trace stages and the start of the post stages = exchange step 1
      if (r->postStarted && cudaEventElapsedTime(&ms, r->ev[2 * EID_K_INDIRECT + 1], r->evPost) == cudaSuccess) r->stats.exchangeMs = ms; else cudaGetLastError();
    }
    r->statsPending = false;
  }
  *out = r->stats;
  return EID_OK;
  EID_CATCH
}

int eid_renderer_set_band(eid_renderer* r, uint32_t y0, uint32_t y1) {
  EID_TRY
  if (!r) raise(EID_ERR_INVALID, "eid_renderer_set_band: null renderer");
  if (y0 > y1 || y1 > r->height) raise(EID_ERR_INVALID, "band [%u,%u) outside 0..%u", y0, y1, r->height);
  if ((y0 % 8) != 0 || (y1 % 8) != 0) raise(EID_ERR_INVALID, "band edges must be multiples of 8 rows (the direct stage works in 8 x 8 pixel tiles)");
  if (y1 == y0) raise(EID_ERR_INVALID, "empty band");
  r->sFirst = y0; r->sRows = y1 - y0; r->sStride = 1u << 20; r->stripesSet = true;
  return EID_OK;
  EID_CATCH
}

int eid_renderer_set_stripes(eid_renderer* r, uint32_t rank, uint32_t world, uint32_t stripeRows) {
  EID_TRY
  if (!r) raise(EID_ERR_INVALID, "eid_renderer_set_stripes: null renderer");
  if (!world || rank >= world) raise(EID_ERR_INVALID, "rank %u outside world %u", rank, world);
  if (!stripeRows || (stripeRows % 16) != 0) raise(EID_ERR_INVALID, "stripeRows must be a positive multiple of 16");
  if (r->height % (world * stripeRows) != 0) raise(EID_ERR_INVALID, "renderer height %u is not a multiple of world*stripeRows = %u (pad the allocation)", r->height, world * stripeRows);
  r->sFirst = rank * stripeRows; r->sRows = stripeRows; r->sStride = world * stripeRows; r->stripesSet = true;
  return EID_OK;
  EID_CATCH
}

int eid_renderer_exchange_groups(eid_renderer* r) {
  if (!r) return -1;
  if (!r->stripesSet || r->sStride >= (1u << 20)) return 1;
  return (int)(r->height / r->sStride);
}

// Byte range of this rank's stripe in exchange group `group` of buffer `which`: the group occupies
// [groupOffset, groupOffset + world*chunkBytes) and rank k's chunk starts at groupOffset + k*chunkBytes, so an in-place
// all-gather over that region (equal chunks) completes the group on every rank.
int eid_renderer_exchange_range(eid_renderer* r, int which, uint32_t group, void** dev_base, uint64_t* offset, uint64_t* bytes) {
  EID_TRY
  if (!r || !dev_base || !offset || !bytes) raise(EID_ERR_INVALID, "eid_renderer_exchange_range: null argument");
  size_t total = 0; void* p = bufferPtr(r, which, total);
  if (!p) raise(EID_ERR_INVALID, "no such buffer %d", which);
  if ((int)group >= eid_renderer_exchange_groups(r)) raise(EID_ERR_INVALID, "exchange group %u out of range", group);
  const bool single = r->sStride >= (1u << 20);
  const uint64_t y0 = (uint64_t)r->sFirst + (single ? 0 : (uint64_t)group * r->sStride), y1 = y0 + r->sRows;
  const uint64_t W = r->width;
  const uint64_t sw = r->hasRun ? (uint64_t)r->lastState.size.x : W;   // reservoir buffers are pitched by RtxState.size.x
  uint64_t rowBytes = 0, a = y0, b = y1;
  switch (which) {
    case EID_BUF_THIS_GBUFFER: case EID_BUF_LAST_GBUFFER: case EID_BUF_DIRECT: case EID_BUF_INDIRECT:
    case EID_BUF_DENOISE_DIR_A: case EID_BUF_DENOISE_DIR_B: rowBytes = W * 16; break;
    case EID_BUF_MOTION: rowBytes = W * 4; break;
    case EID_BUF_DENOISE_IND_A: case EID_BUF_DENOISE_IND_B: rowBytes = W * 16; a = y0 / 2; b = y1 / 2; break;   // half-res rows, full-res pitch
    case EID_BUF_THIS_DIRECT_RESV: case EID_BUF_LAST_DIRECT_RESV: rowBytes = sw * sizeof(DirectReservoir); break;
    case EID_BUF_THIS_INDIRECT_RESV: case EID_BUF_LAST_INDIRECT_RESV: rowBytes = (sw / 2) * sizeof(IndirectReservoir); a = y0 / 2; b = y1 / 2; break;
    default: raise(EID_ERR_INVALID, "no band layout for buffer %d", which);
  }
  if (b * rowBytes > total) raise(EID_ERR_INVALID, "stripe exceeds the buffer (allocation not padded to the stripe layout?)");
  *dev_base = p; *offset = a * rowBytes; *bytes = (b - a) * rowBytes;
  return EID_OK;
  EID_CATCH
}

int eid_renderer_band_range(eid_renderer* r, int which, void** dev_base, uint64_t* offset, uint64_t* bytes) {
  return eid_renderer_exchange_range(r, which, 0, dev_base, offset, bytes);
}

}  // extern "C"

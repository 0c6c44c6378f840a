// group.cu — one frame on N GPUs of a node behind the C-ABI (no reference analogue: the reference renders on one GPU).
//
// One process (or thread) per GPU owns one eid_renderer whose allocation is padded to N equal row bands; eid_group joins the N
// renderers through an NCCL communicator that the LIBRARY owns (libnccl.so.2 is resolved with dlopen at the first eid_group call, so
// libeidola.so itself does not link NCCL and single-GPU hosts never load it).  eid_group_run enqueues the whole multi-GPU frame:
//
//   direct_stage on the band            -> exchange A (one NCCL group call: G-buffer + pre-denoise direct image [+ direct reservoirs]),
//                                          on the group's communication stream, while ...
//   indirect_stage on the band             ... runs on the render stream
//                                       -> exchange B (quarter-res indirect image [+ indirect reservoirs])
//   denoise + compose on the band (+ the reach of the A-Trous levels), or on the whole frame (replicated post)
//                                       -> exchange C (optional): all-gather of the two composed images = the one collective of the
//                                          north star; skipped when every rank delivers its own band to the host
//
// Every exchange is ONE NCCL launch (ncclGroupStart / End aggregates its in-place all-gathers).  Reservoir history crosses band
// edges only when the camera moves: it is gathered (lazily, before the direct stage of the frame that needs it, or eagerly behind the
// post stages) so that a moving camera stays bit-identical to the single-GPU frame.
#include <dlfcn.h>
#include <cstdlib>
#include "group.h"

namespace {

struct NcclId { char internal[128]; };   // ncclUniqueId
struct Nccl {
  void* lib = nullptr;
  int (*GetUniqueId)(NcclId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclId, int) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*GetVersion)(int*) = nullptr;
};

Nccl& nccl() {
  static Nccl n;
  if (n.lib) return n;
  const char* names[] = {getenv("EID_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    if (!nm || !*nm) continue;
    n.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);   // a host that already loaded NCCL (e.g. through torch) gets that copy: same soname
    if (n.lib) break;
  }
  if (!n.lib) raise(EID_ERR_UNSUPPORTED, "eid_group needs NCCL: libnccl.so.2 could not be loaded (%s); set EID_NCCL_LIB", dlerror());
  auto sym = [&](const char* s) { void* p = dlsym(n.lib, s); if (!p) raise(EID_ERR_UNSUPPORTED, "NCCL symbol %s not found", s); return p; };
  n.GetUniqueId = (int (*)(NcclId*))sym("ncclGetUniqueId");
  n.CommInitRank = (int (*)(NcclComm*, int, NcclId, int))sym("ncclCommInitRank");
  n.CommDestroy = (int (*)(NcclComm))sym("ncclCommDestroy");
  n.AllGather = (int (*)(const void*, void*, size_t, int, NcclComm, cudaStream_t))sym("ncclAllGather");
  n.GroupStart = (int (*)())sym("ncclGroupStart");
  n.GroupEnd = (int (*)())sym("ncclGroupEnd");
  n.GetErrorString = (const char* (*)(int))sym("ncclGetErrorString");
  n.GetVersion = (int (*)(int*))sym("ncclGetVersion");
  return n;
}
#define NCCL_CHECK(x)                                                                                   \
  do {                                                                                                  \
    int e_ = (x);                                                                                       \
    if (e_ != 0) raise(EID_ERR_CUDA, "%s failed: %s (%s:%d)", #x, nccl().GetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

}  // namespace

static void gatherList(eid_group* g, const int* which, int n) {
  if (g->world == 1 || n == 0) return;
  Nccl& N = nccl();
  NCCL_CHECK(N.GroupStart());
  for (int i = 0; i < n; ++i) {
    void* base = nullptr; uint64_t off = 0, bytes = 0;
    if (eid_renderer_exchange_range(g->r, which[i], 0, &base, &off, &bytes) != EID_OK) raise(EID_ERR_INVALID, "%s", eid_last_error());
    char* mine = (char*)base + off;
    NCCL_CHECK(N.AllGather(mine, mine - (size_t)g->rank * bytes, (size_t)bytes, /*ncclChar*/ 0, g->comm, g->cs));   // in place
  }
  NCCL_CHECK(N.GroupEnd());
  g->collectives++;
}
static void commAfter(eid_group* g, cudaEvent_t ev) { CUDA_CHECK(cudaEventRecord(ev, g->r->stream)); CUDA_CHECK(cudaStreamWaitEvent(g->cs, ev, 0)); }
static void renderAfter(eid_group* g, cudaEvent_t ev) { CUDA_CHECK(cudaEventRecord(ev, g->cs)); CUDA_CHECK(cudaStreamWaitEvent(g->r->stream, ev, 0)); }

static bool temporalReuse(const RtxState& st) { return st.ReSTIRState == eTemporal || st.ReSTIRState == eSpatiotemporal; }
static bool cameraMoved(const SceneCamera& c) { return memcmp(&c.projView, &c.lastProjView, sizeof(c.projView)) != 0; }

// the whole multi-GPU frame; `finalGather` = exchange C
static void groupFrame(eid_group* g, const RtxState& st, int frames, bool finalGather) {
  eid_renderer* r = g->r;
  FrameParams P;
  strictOrder(r);
  fillParams(r, st, frames, P);
  const bool temporal = temporalReuse(st);
  if (g->world > 1 && temporal && g->history == 2 && cameraMoved(P.cam) && !g->historyComplete) {
    // lazily complete last frame's reservoirs before the stage that reprojects into them
    const int h[2] = {EID_BUF_LAST_DIRECT_RESV, EID_BUF_LAST_INDIRECT_RESV};
    commAfter(g, g->evH);
    gatherList(g, h, 2);
    renderAfter(g, g->evH);
  }
  g->historyComplete = false;
  if (g->histPending) { CUDA_CHECK(cudaStreamWaitEvent(r->stream, g->evHist, 0)); g->histPending = false; }   // K1 reprojects into the gathered history
  beginFrame(r);
  stageDirect(r, P, r->stream);
  const bool eager = g->world > 1 && temporal && g->history == 1;
  if (g->world > 1) {
    // the pre-denoise direct image lives in denoiseDirTempA with EID_VARIANT_DIRECT_BILATERAL (direct_stage.comp:284-288)
    const int a[3] = {EID_BUF_THIS_GBUFFER, (r->variant & EID_VARIANT_DIRECT_BILATERAL) ? EID_BUF_DENOISE_DIR_A : EID_BUF_DIRECT, EID_BUF_THIS_DIRECT_RESV};
    commAfter(g, g->evA);
    gatherList(g, a, 2);                                    // exchange A, behind indirect_stage
  }
  // the direct denoiser needs exchange A only: geometry planes + K3 go to the renderer's second stream as soon as it has landed, beside
  // indirect_stage; the indirect denoiser follows exchange B on the render stream (true dependencies K1 -> {K2, K3}, K2 -> K4, {K3, K4} -> K5)
  const PostLayout L = postLayout(P, g->world > 1 && g->post == 1);
  if (g->world > 1) { CUDA_CHECK(cudaEventRecord(g->evA2, g->cs)); CUDA_CHECK(cudaStreamWaitEvent(r->aux, g->evA2, 0)); }
  else { CUDA_CHECK(cudaEventRecord(g->evA2, r->stream)); CUDA_CHECK(cudaStreamWaitEvent(r->aux, g->evA2, 0)); }
  markStart(r, EID_K_DENOISE_DIRECT, r->aux);
  stagePrep(r, P, L, r->aux);
  CUDA_CHECK(cudaEventRecord(g->evPrep, r->aux));
  stageDenoiseDirect(r, P, L, r->aux);
  markStop(r, EID_K_DENOISE_DIRECT, r->aux);
  CUDA_CHECK(cudaEventRecord(g->evK3, r->aux));
  stageIndirect(r, P, r->stream);
  if (g->world > 1) {
    const int b[1] = {EID_BUF_DENOISE_IND_A};
    commAfter(g, g->evB);
    gatherList(g, b, 1);                                    // exchange B
    renderAfter(g, g->evX);
    if (eager) {                                            // history for the NEXT frame: behind the post stages
      const int h[2] = {EID_BUF_THIS_DIRECT_RESV, EID_BUF_THIS_INDIRECT_RESV};
      gatherList(g, h, 2);
      CUDA_CHECK(cudaEventRecord(g->evHist, g->cs));
      g->historyComplete = true; g->histPending = true;
    }
  }
  if (r->profiling) CUDA_CHECK(cudaEventRecord(r->evPost, r->stream));   // K2 end .. here = what the render stream waited for exchange B
  r->postStarted = true;
  CUDA_CHECK(cudaStreamWaitEvent(r->stream, g->evPrep, 0));
  // exchange C of the PREVIOUS frame ran behind this frame's trace stages (it only touches the previous parity's direct image and the
  // indirect result image); the first stage that writes the indirect result image is ordered after it here
  if (g->finalPending) { CUDA_CHECK(cudaStreamWaitEvent(r->stream, g->evD, 0)); g->finalPending = false; }
  markStart(r, EID_K_DENOISE_INDIRECT, r->stream);
  stageDenoiseIndirect(r, P, L, r->stream);
  markStop(r, EID_K_DENOISE_INDIRECT, r->stream);
  CUDA_CHECK(cudaStreamWaitEvent(r->stream, g->evK3, 0));
  stageCompose(r, P, L, r->stream);
  endFrame(r);
  if (g->world > 1 && g->post == 1 && finalGather) {
    const int c[2] = {EID_BUF_DIRECT, EID_BUF_INDIRECT};
    commAfter(g, g->evC);
    gatherList(g, c, 2);                                    // exchange C: the composed frame on every rank; complete after eid_group_sync
    CUDA_CHECK(cudaEventRecord(g->evD, g->cs));             // (the render stream is ordered after it when the next frame needs the images)
    g->finalPending = true;
  }
}

extern "C" {

int eid_group_layout(uint32_t height, int world, int rank, uint32_t* y0, uint32_t* y1, uint32_t* padded_height) {
  EID_TRY
  if (!height || world < 1 || rank < 0 || rank >= world) raise(EID_ERR_INVALID, "eid_group_layout: bad height / world / rank");
  const uint32_t b = world == 1 ? height : ((height + world - 1) / world + 7) / 8 * 8;
  if (y0) *y0 = (uint32_t)rank * b;
  if (y1) *y1 = (uint32_t)(rank + 1) * b;
  if (padded_height) *padded_height = world == 1 ? height : b * world;
  return EID_OK;
  EID_CATCH
}

int eid_group_unique_id(void* id128) {
  EID_TRY
  if (!id128) raise(EID_ERR_INVALID, "eid_group_unique_id: null argument");
  NcclId id;
  NCCL_CHECK(nccl().GetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
  return EID_OK;
  EID_CATCH
}

int eid_group_create(eid_group** out, eid_renderer* r, int rank, int world, const void* id128) {
  EID_TRY
  if (!out || !r) raise(EID_ERR_INVALID, "eid_group_create: null argument");
  if (world < 1 || rank < 0 || rank >= world) raise(EID_ERR_INVALID, "eid_group_create: rank %d outside world %d", rank, world);
  if (world > 1 && !id128) raise(EID_ERR_INVALID, "eid_group_create: world > 1 needs the 128-byte id of eid_group_unique_id");
  if (world > 1 && (r->height % world != 0 || (r->height / world) % 8 != 0))
    raise(EID_ERR_INVALID, "eid_group_create: the renderer's allocation height %u is not %d equal bands of a multiple of 8 rows (use eid_group_layout's padded_height)", r->height, world);
  eid_group* g = new eid_group();
  try {
    g->r = r; g->rank = rank; g->world = world; g->bandRows = r->height / world;
    CUDA_CHECK(cudaSetDevice(r->device));
    CUDA_CHECK(cudaStreamCreateWithFlags(&g->cs, cudaStreamNonBlocking));
    r->groupStream = g->cs;                         // eid_renderer_sync / read / get_stats also wait for the exchanges
    for (cudaEvent_t* e : {&g->evA, &g->evB, &g->evX, &g->evC, &g->evD, &g->evH, &g->evA2, &g->evPrep, &g->evK3, &g->evHist}) CUDA_CHECK(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    if (world > 1) {
      NcclId id;
      memcpy(&id, id128, sizeof(id));
      NCCL_CHECK(nccl().CommInitRank(&g->comm, world, id, rank));
      if (eid_renderer_set_band(r, (uint32_t)rank * g->bandRows, (uint32_t)(rank + 1) * g->bandRows) != EID_OK) raise(EID_ERR_INVALID, "%s", eid_last_error());
    }
  } catch (...) { eid_group_destroy(g); throw; }
  *out = g;
  return EID_OK;
  EID_CATCH
}

void eid_group_destroy(eid_group* g) {
  if (!g) return;
  if (g->r) cudaSetDevice(g->r->device);
  if (g->r && g->r->stream) cudaStreamSynchronize(g->r->stream);
  if (g->r && g->r->copyStream) cudaStreamSynchronize(g->r->copyStream);
  if (g->cs) cudaStreamSynchronize(g->cs);
  if (g->pipe) { try { pipelineSync(g); } catch (...) {} pipelineDestroy(g); }
  if (g->r && g->r->groupStream == g->cs) g->r->groupStream = nullptr;
  if (g->copyStream) { cudaStreamSynchronize(g->copyStream); cudaStreamDestroy(g->copyStream); }
  if (g->comm) nccl().CommDestroy(g->comm);
  for (cudaEvent_t e : {g->evA, g->evB, g->evX, g->evC, g->evD, g->evH, g->evA2, g->evPrep, g->evK3, g->evHist, g->evFrameDone, g->evCopyDone}) if (e) cudaEventDestroy(e);
  cudaFree(g->staging[0]); cudaFree(g->staging[1]);
  if (g->cs) cudaStreamDestroy(g->cs);
  delete g;
}

int eid_group_set_mode(eid_group* g, int post_sharded, int history, int gather_final) {
  EID_TRY
  if (!g) raise(EID_ERR_INVALID, "eid_group_set_mode: null group");
  if (history < 0 || history > 2) raise(EID_ERR_INVALID, "eid_group_set_mode: history 0 (never), 1 (every frame) or 2 (when the camera moved)");
  if (g->pipe) { g->history = history; return EID_OK; }     // a stage pipeline has no post / gather modes: only the history policy applies
  g->post = post_sharded != 0; g->history = history; g->gatherFinal = gather_final != 0;
  g->historyComplete = false;
  return EID_OK;
  EID_CATCH
}

int eid_group_run(eid_group* g, const RtxState* state, int frames) {
  EID_TRY
  if (!g || !state) raise(EID_ERR_INVALID, "eid_group_run: null argument");
  CUDA_CHECK(cudaSetDevice(g->r->device));
  if (g->pipe) { pipelineFrame(g, *state, frames, true); return EID_OK; }
  groupFrame(g, *state, frames, g->gatherFinal != 0);
  CUDA_CHECK(cudaGetLastError());
  return EID_OK;
  EID_CATCH
}

// Host delivery without a funnel: every rank copies the rows of ITS band of the two composed images into the host images (one image
// pair shared by all ranks, e.g. POSIX shared memory that each process registered with cudaHostRegister; pitch = size.x * 16 bytes).
// Device-side snapshot on the render stream, device-to-host copy on a copy stream while the next frame renders; exchange C is skipped.
int eid_group_render_host_async(eid_group* g, const SceneCamera* cam, const RtxState* state, int frames, float* direct_host, float* indirect_host) {
  EID_TRY
  if (!g || !state) raise(EID_ERR_INVALID, "eid_group_render_host_async: null argument");
  eid_renderer* r = g->r;
  CUDA_CHECK(cudaSetDevice(r->device));
  ensureCopyStream(r);
  if (cam) r->scene->host.camera = *cam;
  uint32_t y0 = 0, y1 = 0;
  if (g->pipe) {
    // stage pipeline: the composed rows exist on the post ranks only, and they leave through every rank's PCIe link (pipeline.cu)
    pipelineFrame(g, *state, frames, false);
    pipelineDeliver(g, *state, direct_host, indirect_host);
    return EID_OK;
  } else {
    groupFrame(g, *state, frames, false);           // (fillParams waits for the copy that read this parity's images two frames ago)
    y0 = g->world > 1 ? (uint32_t)g->rank * g->bandRows : 0;
    y1 = std::min<uint32_t>(y0 + g->bandRows, (uint32_t)state->size.y);
  }
  if (y1 > y0) {
    // the band is copied IN PLACE from this parity's result images on the renderer's copy stream, behind the next frame (other parity)
    const int set = r->lastSet;
    const size_t first = (size_t)y0 * r->width;
    CUDA_CHECK(cudaEventRecord(r->evFrameDone, r->stream));
    CUDA_CHECK(cudaStreamWaitEvent(r->copyStream, r->evFrameDone, 0));
    const size_t rowBytes = (size_t)state->size.x * 16;
    if (direct_host) CUDA_CHECK(cudaMemcpy2DAsync((char*)direct_host + (size_t)y0 * rowBytes, rowBytes, r->directImg + first, (size_t)r->width * 16, rowBytes, y1 - y0, cudaMemcpyDeviceToHost, r->copyStream));
    if (indirect_host) CUDA_CHECK(cudaMemcpy2DAsync((char*)indirect_host + (size_t)y0 * rowBytes, rowBytes, r->indirectImg + first, (size_t)r->width * 16, rowBytes, y1 - y0, cudaMemcpyDeviceToHost, r->copyStream));
    CUDA_CHECK(cudaEventRecord(r->evCopyDone2[set], r->copyStream));
    r->copyPending2[set] = true;
  }
  return EID_OK;
  EID_CATCH
}

int eid_group_wait_host(eid_group* g) {
  EID_TRY
  if (!g) raise(EID_ERR_INVALID, "eid_group_wait_host: null group");
  CUDA_CHECK(cudaSetDevice(g->r->device));
  if (g->r->copyStream) CUDA_CHECK(cudaStreamSynchronize(g->r->copyStream));
  return EID_OK;
  EID_CATCH
}

int eid_group_sync(eid_group* g) {
  EID_TRY
  if (!g) raise(EID_ERR_INVALID, "eid_group_sync: null group");
  CUDA_CHECK(cudaSetDevice(g->r->device));
  CUDA_CHECK(cudaStreamSynchronize(g->r->stream));
  CUDA_CHECK(cudaStreamSynchronize(g->cs));
  if (g->pipe) pipelineSync(g);
  return EID_OK;
  EID_CATCH
}

int eid_group_get_info(eid_group* g, eid_group_info* out) {
  EID_TRY
  if (!g || !out) raise(EID_ERR_INVALID, "eid_group_get_info: null argument");
  memset(out, 0, sizeof(*out));
  if (g->pipe) { pipelineInfo(g, out); return EID_OK; }
  out->rank = g->rank; out->world = g->world; out->bandRows = g->bandRows;
  out->y0 = g->world > 1 ? (uint32_t)g->rank * g->bandRows : 0; out->y1 = out->y0 + g->bandRows;
  out->collectives = g->collectives;
  if (g->world > 1) { int v = 0; if (nccl().GetVersion(&v) == 0) out->ncclVersion = v; }
  return EID_OK;
  EID_CATCH
}

}  // extern "C"

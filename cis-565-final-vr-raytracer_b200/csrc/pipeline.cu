// pipeline.cu — eid_group as a STAGE PIPELINE over NVLink peer memory (no reference analogue: the reference renders on one GPU).
//
// Row bands (group.cu) split ONE frame over N GPUs; every band still pays the latency floor of indirect_stage (three dependent ray
// queries of ~80 us worst-case latency each, whatever the band height) and two or three exchange steps, so the frame rate saturates
// near 2.2 x at N = 8 (profiles/README.md).  The stages of Renderer::run, however, only depend on their OWN history:
//
//     direct_stage(f)    reads  direct_stage(f-1)            (last G-buffer, last direct reservoirs)
//     indirect_stage(f)  reads  direct_stage(f), indirect_stage(f-1)   (this G-buffer; last indirect reservoirs)
//     denoise/compose(f) reads  direct_stage(f), indirect_stage(f)     (no history at all)
//
// so the N ranks form a pipeline instead: nDirect ranks run direct_stage on row bands, nIndirect ranks run indirect_stage on row
// bands, nPost ranks run denoise + compose on row bands — each group one frame behind the previous one.  Throughput is set by the
// slowest STAGE (e.g. 3 | 3 | 2 ranks at N = 8), not by the sum of all stages plus exchanges; every frame is bit-identical to
// eid_renderer_run on one GPU (the stages run the same kernels on the same buffers; rows travel verbatim).  Latency per frame grows
// by the two hand-overs; the band mode of group.cu stays the low-latency alternative.
//
// Transport: no collective.  At creation every rank exports its buffers as CUDA IPC handles through a rendezvous file in /dev/shm
// (keyed by the group's 128-byte id) and maps its peers' buffers.  A producer writes the rows a consumer needs STRAIGHT INTO THE
// CONSUMER'S BUFFERS with peer copies on its communication stream (copy engines over NVLink / NVSwitch: no SM is taken from the
// kernels), then raises a sequence flag in the consumer's memory; the consumer's stream waits for the flag with a stream memory
// operation (cuStreamWaitValue32: no host round trip, no spinning SM).  Back-pressure is the same mechanism in the other direction:
// a consumer acknowledges a frame once it has read it, and a producer only overwrites a (per-parity) buffer after that.
//
//   direct rank d  --[G-buffer rows, pre-denoise direct rows, K2's gathered temporal lookups]-->  indirect ranks, post ranks
//   indirect rank  --[quarter-res pre-denoise indirect rows]-->  post ranks
//   same-stage ranks exchange their band of the G-buffer / reservoirs only when temporal reuse can cross a band edge (moving camera)
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>
#include <algorithm>
#include <string>
#include <cuda.h>
#include "group.h"

namespace {

constexpr int MAXW = 16;
#ifndef EID_PIPE_HISTORY_CHUNKS
#define EID_PIPE_HISTORY_CHUNKS 4      // row chunks of direct_stage when its band's history is on the peers' critical path (moving camera)
#endif
enum { ROLE_D = EID_STAGE_DIRECT, ROLE_I = EID_STAGE_INDIRECT, ROLE_P = EID_STAGE_POST };
enum FlagKind { F_READY_D, F_READY_I, F_READY_H, F_ACK_D, F_ACK_I, F_ACK_H, F_READY_V, F_ACK_V, F_KINDS };
enum BufIdx { B_G0, B_G1, B_DIR0, B_DIR1, B_K2G0, B_K2G1, B_K2MV0, B_K2MV1, B_INDIN0, B_INDIN1, B_DR0, B_DR1, B_IR0, B_IR1, B_FLAGS, B_DELIV, B_COUNT };

struct RankLayout { int role = 0, index = 0, count = 1; uint32_t y0 = 0, y1 = 0; };

// ---- layout: who runs what on which rows ------------------------------------------------------------------------------------
uint32_t bandRowsOf(uint32_t height, int n) { return ((height + n - 1) / n + 7) / 8 * 8; }

void stageCounts(int world, int& nD, int& nI, int& nP) {
  if (nD == 0 && nI == 0 && nP == 0) {
    // default split by the measured stage times of the 1080p workload (direct 0.89 ms, indirect 0.60 ms with a ~0.28 ms floor per band,
    // denoise + compose 0.48 ms): 2: 1|-|1 (post rank also traces the indirect stage), 3: 1|1|1, 4: 2|1|1, 5: 2|2|1, 8: 3|3|2
    nP = std::max(1, world / 4);
    nD = (world - nP + 1) / 2;
    nI = world - nP - nD;
  }
  if (nD < 1 || nP < 1 || nI < 0 || nD + nI + nP != world) raise(EID_ERR_INVALID, "pipeline stages %d | %d | %d do not add up to %d ranks (direct >= 1, post >= 1)", nD, nI, nP, world);
  if (nI == 0 && nP != 1) raise(EID_ERR_INVALID, "pipeline without indirect ranks needs exactly one post rank (it runs indirect_stage, denoise and compose on the whole frame)");
  if (world > MAXW) raise(EID_ERR_INVALID, "pipeline supports at most %d ranks", MAXW);
}

RankLayout rankLayout(uint32_t height, int rank, int nD, int nI, int nP) {
  RankLayout L;
  if (rank < nD) { L.role = ROLE_D; L.index = rank; L.count = nD; }
  else if (rank < nD + nI) { L.role = ROLE_I; L.index = rank - nD; L.count = nI; }
  else { L.role = nI == 0 ? (ROLE_I | ROLE_P) : ROLE_P; L.index = rank - nD - nI; L.count = nP; }
  const uint32_t b = bandRowsOf(height, L.count);
  L.y0 = (uint32_t)L.index * b; L.y1 = L.y0 + b;
  return L;
}
uint32_t paddedHeightOf(uint32_t height, int nD, int nI, int nP) {
  uint32_t p = (height + 7) / 8 * 8;
  for (int n : {nD, nI, nP}) if (n > 0) p = std::max(p, bandRowsOf(height, n) * (uint32_t)n);
  return p;
}

// rows [a, b) of a buffer that a consumer needs from the producers of a stage, in the buffer's own row units
struct Range { int a = 0, b = 0; bool empty() const { return b <= a; } };
Range clampRange(int a, int b, int lo, int hi) { Range r; r.a = std::max(a, lo); r.b = std::min(b, hi); return r; }
Range intersect(Range x, int a, int b) { return clampRange(x.a, x.b, a, b); }
// what a consumer needs of direct_stage's outputs: G-buffer rows, pre-denoise direct rows (full res), gathered lookups (quarter-res rows)
struct DirectNeeds { Range g, d, q; };
DirectNeeds directNeeds(const RankLayout& c, int padded) {
  DirectNeeds n;
  if ((c.role & ROLE_I) && (c.role & ROLE_P)) { n.g = n.d = clampRange(0, padded, 0, padded); n.q = clampRange(0, padded / 2, 0, padded / 2); }
  else if (c.role & ROLE_I) { n.g = clampRange((int)c.y0, (int)c.y1, 0, padded); n.q = clampRange((int)c.y0 / 2, (int)c.y1 / 2, 0, padded / 2); }
  else if (c.role & ROLE_P) {
    // denoise_prep evaluates the geometry planes on the band +- 124 rows (render.cu: stagePrep), A-Trous level 0 of the direct image on the
    // band +- 28 rows with taps +- 2 rows beyond
    n.g = clampRange((int)c.y0 - 124, (int)c.y1 + 124, 0, padded);
    n.d = clampRange((int)c.y0 - 32, (int)c.y1 + 32, 0, padded);
  }
  return n;
}
// quarter-res rows of the pre-denoise indirect image a post rank needs: band / 2 +- 60 rows (level 0) + 2 rows of taps
Range indirectNeeds(const RankLayout& c, int padded) {
  if (!(c.role & ROLE_P) || (c.role & ROLE_I)) return Range{};
  return clampRange((int)c.y0 / 2 - 64, (int)c.y1 / 2 + 64, 0, padded / 2);
}

// ---- rendezvous through a file in /dev/shm ----------------------------------------------------------------------------------------
struct ShmBuf { cudaIpcMemHandle_t h; unsigned long long offset, bytes; };
struct ShmRank {
  uint32_t ready, opened, closed, pad;
  int32_t pid, device;
  uint32_t width, height;
  ShmBuf buf[B_COUNT];
};
struct ShmHdr { ShmRank ranks[MAXW]; };

double nowSec() { timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }
uint32_t loadAcq(const uint32_t* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
void storeRel(uint32_t* p, uint32_t v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }

// driver-API entry points resolved through the runtime (libeidola.so does not link libcuda)
typedef CUresult (*WaitValueFn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
typedef CUresult (*AddrRangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);
void* driverEntry(const char* name) {
  void* f = nullptr;
  cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
  if (cudaGetDriverEntryPoint(name, &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) { cudaGetLastError(); return nullptr; }
  return f;
}

__global__ void k_flag_set(uint32_t* flag, uint32_t value) {       // release: everything this stream did before is visible before the flag
  __threadfence_system();
  *(volatile uint32_t*)flag = value;
}
// Peer copy by the SMs (16-byte stores over NVLink).  The copy engines move a band at ~0.25 TB/s; that is fine for the hand-overs that
// run beside the next frame's kernels, but the same-stage history is ON the critical path (the peer's next direct_stage waits for it) while
// this rank's SMs are idle anyway: a grid-stride copy kernel delivers it at NVLink speed.
__global__ void __launch_bounds__(256) k_peer_copy(uint4* __restrict__ dst, const uint4* __restrict__ src, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
__global__ void k_flag_wait(const uint32_t* flag, uint32_t value) {   // fallback when stream memory operations are unavailable
  while ((int32_t)(*(volatile const uint32_t*)flag - value) < 0) __nanosleep(200);
  __threadfence_system();
}

}  // namespace

struct EidPipe {
  int nD = 0, nI = 0, nP = 0;
  uint32_t height = 0, padded = 0;
  RankLayout me;
  RankLayout ranks[MAXW];
  void* peer[MAXW][B_COUNT] = {};          // peer[j][b]: rank j's buffer b in THIS process's address space (own rank: the local pointers)
  std::vector<void*> openedBases;
  std::vector<std::pair<std::string, void*>> openedByHandle;
  uint32_t* flags = nullptr;               // local: [kind][peer]
  uint32_t seq = 0;                        // frames enqueued so far
  int lastSeqOfParity[2] = {-1, -1};
  // Communication streams: pushes to different consumer GROUPS must not queue behind each other — a push waits for its consumer's
  // acknowledgement of the frame before last, and the post ranks (two stages behind) acknowledge just in time, so on one stream the
  // next frame's rows for the indirect ranks (or the history the peer's next direct_stage waits for) would stall behind them.
  // g->cs: to the indirect ranks (direct ranks) / to the post ranks (indirect ranks, post ranks' delivery); csP: direct -> post; csH: history
  cudaStream_t csP = nullptr, csH = nullptr;
  cudaEvent_t evStage = nullptr, evPush[2] = {nullptr, nullptr}, evPushP[2] = {nullptr, nullptr}, evPushH[2] = {nullptr, nullptr}, evPushI[2] = {nullptr, nullptr},
              evPrep = nullptr, evK3 = nullptr, evFork = nullptr;
  bool pushValid[2] = {false, false}, pushHValid[2] = {false, false}, pushIValid[2] = {false, false};
  bool historyComplete = false;            // the LAST buffers of the next frame already hold the peers' rows
  // host delivery spread over every rank's PCIe link: post ranks write the composed rows of rank k's delivery band into k's staging
  // buffer ([parity][direct | indirect][delivRows x width] float4), k copies them to the host
  float4* deliv = nullptr;
  uint32_t delivRows = 0, dseq = 0;
  cudaEvent_t evDelivPush = nullptr, evDone = nullptr;
  // indirect ranks: the path tracing of frame f + 1 (k_gi_begin .. k_gi_bounce; needs this frame's G-buffer only) runs on its own stream
  // and scratch while frame f is still in flight; only k_gi_finish (temporal reuse) is ordered after the previous frame's
  cudaStream_t k2Stream[2] = {nullptr, nullptr};
  cudaEvent_t evTrace[2] = {nullptr, nullptr}, evFinish[2] = {nullptr, nullptr};
  bool finishValid[2] = {false, false};
  bool ackPending = false;                 // eid_group_run: the frame's acknowledgement is enqueued with the NEXT frame (see pipelineFrame)
  ShmHdr* shm = nullptr;
  std::string shmPath;
  WaitValueFn waitValue = nullptr;
  unsigned long long peerCopies = 0, peerBytes = 0;
};

namespace {

uint32_t* flagOf(EidPipe* p, int owner, int kind, int from) { return (uint32_t*)p->peer[owner][B_FLAGS] + kind * MAXW + from; }

void waitFlag(eid_group* g, cudaStream_t st, int kind, int from, uint32_t value) {
  EidPipe* p = g->pipe;
  uint32_t* f = p->flags + kind * MAXW + from;
  if (p->waitValue) {
    const CUresult e = p->waitValue((CUstream)st, (CUdeviceptr)(uintptr_t)f, value, CU_STREAM_WAIT_VALUE_GEQ);
    if (e == CUDA_SUCCESS) return;
    p->waitValue = nullptr;                // not supported on this device / driver: fall back to the polling kernel for good
  }
  k_flag_wait<<<1, 1, 0, st>>>(f, value);
}
void setFlag(eid_group* g, cudaStream_t st, int owner, int kind, uint32_t value) {
  k_flag_set<<<1, 1, 0, st>>>(flagOf(g->pipe, owner, kind, g->rank), value);
}
void pushRows(eid_group* g, cudaStream_t cs, int to, int buf, size_t rowBytes, Range r) {
  if (r.empty()) return;
  EidPipe* p = g->pipe;
  const size_t off = (size_t)r.a * rowBytes, bytes = (size_t)(r.b - r.a) * rowBytes;
  CUDA_CHECK(cudaMemcpyAsync((char*)p->peer[to][buf] + off, (const char*)p->peer[g->rank][buf] + off, bytes, cudaMemcpyDeviceToDevice, cs));
  p->peerCopies++; p->peerBytes += bytes;
}
void pushRowsSM(eid_group* g, cudaStream_t cs, int to, int buf, size_t rowBytes, Range r) {
  if (r.empty()) return;
  EidPipe* p = g->pipe;
  const size_t off = (size_t)r.a * rowBytes, bytes = (size_t)(r.b - r.a) * rowBytes;
  if ((off | bytes) & 15) { pushRows(g, cs, to, buf, rowBytes, r); return; }
  k_peer_copy<<<g->r->smCount * 2, 256, 0, cs>>>((uint4*)((char*)p->peer[to][buf] + off), (const uint4*)((const char*)p->peer[g->rank][buf] + off), bytes / 16);
  p->peerCopies++; p->peerBytes += bytes;
}
// a local buffer pushed under another index at the consumer (indA -> the consumer's per-parity indIn)
void pushRowsFrom(eid_group* g, int to, int dstBuf, const void* src, size_t rowBytes, Range r) {
  if (r.empty()) return;
  EidPipe* p = g->pipe;
  const size_t off = (size_t)r.a * rowBytes, bytes = (size_t)(r.b - r.a) * rowBytes;
  CUDA_CHECK(cudaMemcpyAsync((char*)p->peer[to][dstBuf] + off, (const char*)src + off, bytes, cudaMemcpyDeviceToDevice, g->cs));
  p->peerCopies++; p->peerBytes += bytes;
}

bool temporalReuse(const RtxState& st) { return st.ReSTIRState == eTemporal || st.ReSTIRState == eSpatiotemporal; }
bool cameraMoved(const SceneCamera& c) { return memcmp(&c.projView, &c.lastProjView, sizeof(c.projView)) != 0; }

// Same-stage history (moving camera): this rank's band of the buffers the NEXT frame reprojects into goes to the other ranks of the stage.
// `last` = false: the buffers this frame wrote (eager, behind the stage); true: the LAST buffers of this frame (lazy, before the stage).
// `rows` (full-res rows, a part of the band): direct_stage runs in row chunks when the history is on the critical path, and every chunk leaves
// as soon as it is done; `first` waits for the peers' acknowledgement, `final` raises their flag.  Chunks that leave while the stage is
// still running use the copy engines (the SMs are busy); a whole band behind an idle stage uses the SM copy kernel.
void pushHistory(eid_group* g, const FrameParams& P, int set, bool last, uint32_t readyValue, Range rows, bool first = true, bool final = true, bool sm = true) {
  EidPipe* p = g->pipe;
  const RankLayout& me = p->me;
  const int thisIdx = last ? set : !set;                      // fillParams: last* = [set], this* = [!set]
  const size_t sw = (size_t)P.st.size.x;
  auto push = [&](int j, int buf, size_t rowBytes, Range r) { if (sm) pushRowsSM(g, p->csH, j, buf, rowBytes, r); else pushRows(g, p->csH, j, buf, rowBytes, r); };
  for (int j = 0; j < g->world; ++j) {
    if (j == g->rank || p->ranks[j].role != me.role) continue;
    // the peer read these rows' buffer (as `last`) in every frame up to the previous one: wait until it has finished that stage
    if (first && p->seq > 0) waitFlag(g, p->csH, F_ACK_H, j, p->seq);
    if (me.role & ROLE_D) {
      push(j, B_G0 + thisIdx, (size_t)P.pitch * 16, rows);
      push(j, B_DR0 + thisIdx, sw * sizeof(DirectReservoir), rows);
    } else {
      push(j, B_IR0 + thisIdx, (sw / 2) * sizeof(IndirectReservoir), Range{rows.a / 2, rows.b / 2});
    }
    if (final) setFlag(g, p->csH, j, F_READY_H, readyValue);
  }
}

// indirect_stage of frame n on a rank that traces it: the path-tracing half on the stream / scratch of the frame's parity — it only waits
// for this frame's G-buffer rows (`waitInputs` enqueues those waits) and for the frame before last to have left that scratch —, then
// k_gi_finish on the render stream, behind the previous frame's.  Falls back to the one-stream form when the wavefront form is not in use.
template <class WaitInputs>
bool indirectTraceAsync(eid_group* g, FrameParams& P, uint32_t n, WaitInputs waitInputs) {
  eid_renderer* r = g->r;
  EidPipe* p = g->pipe;
  const int par = (int)(n & 1u);
  const bool serial = getenv("EID_PIPE_K2_SERIAL") != nullptr;
  if (P.wv.slots && !serial && par) { r->ensureWave(P.st.maxDepth - 1, 1); P.wv = r->waveView(1); }
  if (serial || !indirectIsWavefront(r, P)) {
    waitInputs(r->stream);
    beginFrame(r);
    return false;
  }
  cudaStream_t ts = p->k2Stream[par];
  // what the render stream did up to here (history waits, acknowledgements of earlier frames) need not hold the tracing back; what must:
  // the inputs, and finish(n - 2) — it read this scratch, and its frame's ray counters went to the host after it
  if (p->finishValid[par]) CUDA_CHECK(cudaStreamWaitEvent(ts, p->evFinish[par], 0));
  waitInputs(ts);
  beginFrame(r, ts);
  markStart(r, EID_K_INDIRECT, ts);
  stageIndirectTrace(r, P, ts, par);
  CUDA_CHECK(cudaEventRecord(p->evTrace[par], ts));
  return true;
}
void indirectFinish(eid_group* g, const FrameParams& P, uint32_t n, bool split) {
  eid_renderer* r = g->r;
  EidPipe* p = g->pipe;
  if (!split) { stageIndirect(r, P, r->stream); return; }
  CUDA_CHECK(cudaStreamWaitEvent(r->stream, p->evTrace[n & 1u], 0));
  stageIndirectFinish(r, P, r->stream);
  markStop(r, EID_K_INDIRECT, r->stream);
}
void indirectDone(eid_group* g, uint32_t n) {        // after endFrame: the scratch of this parity and its counters are free again
  EidPipe* p = g->pipe;
  const int par = (int)(n & 1u);
  if (!p->evFinish[par]) return;
  CUDA_CHECK(cudaEventRecord(p->evFinish[par], g->r->stream));
  p->finishValid[par] = true;
}

}  // namespace

// ---- the frame --------------------------------------------------------------------------------------------------------------
void pipelineAck(eid_group* g, cudaStream_t st) {
  EidPipe* p = g->pipe;
  const uint32_t v = p->seq;               // = index of the frame just enqueued + 1
  if (p->me.role & ROLE_P) {
    for (int j = 0; j < g->world; ++j) {
      const int role = p->ranks[j].role;
      if (role == ROLE_D) {
        const DirectNeeds n = directNeeds(p->me, (int)p->padded);
        const RankLayout& d = p->ranks[j];
        if (!intersect(n.g, (int)d.y0, (int)d.y1).empty() || !intersect(n.d, (int)d.y0, (int)d.y1).empty()) setFlag(g, st, j, F_ACK_D, v);
      } else if (role == ROLE_I) {
        const RankLayout& i = p->ranks[j];
        if (!intersect(indirectNeeds(p->me, (int)p->padded), (int)i.y0 / 2, (int)i.y1 / 2).empty()) setFlag(g, st, j, F_ACK_I, v);
      }
    }
  }
}

void pipelineFrame(eid_group* g, const RtxState& st, int frames, bool ackNow) {
  eid_renderer* r = g->r;
  EidPipe* p = g->pipe;
  if (r->variant) raise(EID_ERR_UNSUPPORTED, "the stage pipeline runs the reference's live shaders only (eid_renderer_set_variant must be 0)");
  if (st.ReSTIRState == eSpatial || st.ReSTIRState == eSpatiotemporal) {
    if (p->nD > 1) raise(EID_ERR_UNSUPPORTED, "spatial reuse across the bands of several direct ranks is not built for the stage pipeline (use one direct rank, or the band mode of eid_group_create)");
  }
  if ((uint32_t)st.size.y > p->height) raise(EID_ERR_INVALID, "RtxState.size.y %d exceeds the height %u the pipeline was laid out for", st.size.y, p->height);
  // A post rank's composed images live in the buffers its producers write into, so the acknowledgement that frees them for the frame
  // after next is only enqueued now, with the next frame: until then eid_renderer_read / get_outputs see the finished frame.  (Stream
  // order is unchanged — the acknowledgement follows that frame's compose —, and a host that enqueues ahead loses nothing.)
  if (p->ackPending) { pipelineAck(g, r->stream); p->ackPending = false; }
  FrameParams P;
  strictOrder(r);
  fillParams(r, st, frames, P);
  if (p->me.role & ROLE_I) P.indIn = P.indA;                  // this rank traces indirect_stage itself: its denoiser reads the local image
  const int set = r->lastSet;
  const uint32_t n = p->seq;                                  // index of this frame
  const int prev = p->lastSeqOfParity[set];                   // last frame that used this parity's buffers
  const RankLayout& me = p->me;
  const int padded = (int)p->padded;
  const bool temporal = temporalReuse(st);
  const int stagePeers = me.count;
  const bool needHistory = stagePeers > 1 && temporal && g->history != 0 && (g->history == 1 || cameraMoved(P.cam)) && !(me.role & ROLE_P);
  const bool eagerHistory = stagePeers > 1 && temporal && (g->history == 1 || (g->history == 2 && cameraMoved(P.cam))) && !(me.role & ROLE_P);
  const size_t rowG = (size_t)P.pitch * 16;

  if (me.role == ROLE_D) {
    // K1 rewrites this parity's G-buffer / direct image / gathered lookups: the peer copies of the frame that used them last must have drained
    if (p->pushValid[set]) { CUDA_CHECK(cudaStreamWaitEvent(r->stream, p->evPush[set], 0)); CUDA_CHECK(cudaStreamWaitEvent(r->stream, p->evPushP[set], 0)); }
    if (p->pushHValid[set]) CUDA_CHECK(cudaStreamWaitEvent(r->stream, p->evPushH[set], 0));
    if (needHistory) {
      if (!p->historyComplete) {                              // lazily: first frame, or the camera started moving
        CUDA_CHECK(cudaEventRecord(p->evStage, r->stream)); CUDA_CHECK(cudaStreamWaitEvent(p->csH, p->evStage, 0));
        pushHistory(g, P, set, true, n + 1, Range{(int)me.y0, (int)me.y1});
        // the rows just sent are rewritten by the NEXT frame's stage, which is only ordered after the pushes of its own parity: order this one too
        CUDA_CHECK(cudaEventRecord(p->evPrep, p->csH)); CUDA_CHECK(cudaStreamWaitEvent(r->stream, p->evPrep, 0));
      }
      for (int j = 0; j < g->world; ++j) if (j != g->rank && p->ranks[j].role == ROLE_D) waitFlag(g, r->stream, F_READY_H, j, n + 1);
    }
    beginFrame(r);
    const bool spatial = st.ReSTIRState == eSpatial || st.ReSTIRState == eSpatiotemporal;
    const int chunks = (eagerHistory && !spatial) ? EID_PIPE_HISTORY_CHUNKS : 1;
    if (chunks <= 1) {
      stageDirect(r, P, r->stream);
    } else {
      // The peers' next direct_stage waits for this band's G-buffer / reservoir rows: the band is traced in row chunks and every chunk is
      // on its way (copy engines) while the next one is still being traced; only the last chunk's copy is exposed.
      const int rc = (((int)me.y1 - (int)me.y0 + chunks - 1) / chunks + 7) / 8 * 8;
      markStart(r, EID_K_DIRECT, r->stream);
      bool firstChunk = true;
      for (int a = (int)me.y0; a < (int)me.y1; a += rc) {
        const int b = std::min(a + rc, (int)me.y1);
        FrameParams Pk = P;
        Pk.sFirst = a; Pk.sRows = b - a; Pk.sStride = 1 << 20; Pk.sCount = a < st.size.y ? 1 : 0;
        stageDirect(r, Pk, r->stream, false);
        CUDA_CHECK(cudaEventRecord(p->evStage, r->stream)); CUDA_CHECK(cudaStreamWaitEvent(p->csH, p->evStage, 0));
        pushHistory(g, P, set, false, n + 2, Range{a, b}, firstChunk, b >= (int)me.y1, false);
        firstChunk = false;
      }
      markStop(r, EID_K_DIRECT, r->stream);
    }
    if (stagePeers > 1) for (int j = 0; j < g->world; ++j) if (j != g->rank && p->ranks[j].role == ROLE_D) setFlag(g, r->stream, j, F_ACK_H, n + 1);
    CUDA_CHECK(cudaEventRecord(p->evStage, r->stream));
    CUDA_CHECK(cudaStreamWaitEvent(g->cs, p->evStage, 0)); CUDA_CHECK(cudaStreamWaitEvent(p->csP, p->evStage, 0)); CUDA_CHECK(cudaStreamWaitEvent(p->csH, p->evStage, 0));
    if (eagerHistory && chunks <= 1) pushHistory(g, P, set, false, n + 2, Range{(int)me.y0, (int)me.y1});   // the peer's next direct_stage waits for it: first
    p->historyComplete = eagerHistory;
    for (int j = 0; j < g->world; ++j) {
      const RankLayout& c = p->ranks[j];
      if (c.role == ROLE_D) continue;
      const DirectNeeds need = directNeeds(c, padded);
      const Range rg = intersect(need.g, (int)me.y0, (int)me.y1), rd = intersect(need.d, (int)me.y0, (int)me.y1), rq = intersect(need.q, (int)me.y0 / 2, (int)me.y1 / 2);
      if (rg.empty() && rd.empty() && rq.empty()) continue;
      cudaStream_t cs = (c.role & ROLE_I) ? g->cs : p->csP;
      if (prev >= 0) waitFlag(g, cs, F_ACK_D, j, (uint32_t)prev + 1);
      pushRows(g, cs, j, B_G0 + !set, rowG, rg);
      pushRows(g, cs, j, B_DIR0 + set, rowG, rd);
      pushRows(g, cs, j, B_K2G0 + set, (size_t)(P.pitch / 2) * 16, rq);
      pushRows(g, cs, j, B_K2MV0 + set, (size_t)(P.pitch / 2) * 4, rq);
      setFlag(g, cs, j, F_READY_D, n + 1);
    }
    CUDA_CHECK(cudaEventRecord(p->evPush[set], g->cs)); CUDA_CHECK(cudaEventRecord(p->evPushP[set], p->csP)); p->pushValid[set] = true;
    CUDA_CHECK(cudaEventRecord(p->evPushH[set], p->csH)); p->pushHValid[set] = true;
    endFrame(r);
  } else if (me.role == ROLE_I) {
    // indirect_stage writes its image into this rank's own (otherwise unused) per-parity landing buffer instead of the single denoiseIndTempA,
    // so that only the peer copy of the frame before last has to have drained — not the one the post ranks may still be acknowledging
    P.indA = r->indIn[set];
    if (p->pushIValid[set]) CUDA_CHECK(cudaStreamWaitEvent(r->stream, p->evPushI[set], 0));
    if (p->pushHValid[set]) CUDA_CHECK(cudaStreamWaitEvent(r->stream, p->evPushH[set], 0));   // ... and the history push that read this parity's reservoirs
    if (needHistory) {
      if (!p->historyComplete) {
        CUDA_CHECK(cudaEventRecord(p->evStage, r->stream)); CUDA_CHECK(cudaStreamWaitEvent(p->csH, p->evStage, 0));
        pushHistory(g, P, set, true, n + 1, Range{(int)me.y0, (int)me.y1});
        // the rows just sent are rewritten by the NEXT frame's stage, which is only ordered after the pushes of its own parity: order this one too
        CUDA_CHECK(cudaEventRecord(p->evPrep, p->csH)); CUDA_CHECK(cudaStreamWaitEvent(r->stream, p->evPrep, 0));
      }
      for (int j = 0; j < g->world; ++j) if (j != g->rank && p->ranks[j].role == ROLE_I) waitFlag(g, r->stream, F_READY_H, j, n + 1);
    }
    const DirectNeeds need = directNeeds(me, padded);
    const bool split = indirectTraceAsync(g, P, n, [&](cudaStream_t s) {
      for (int j = 0; j < g->world; ++j) {
        const RankLayout& d = p->ranks[j];
        if (d.role != ROLE_D) continue;
        if (intersect(need.g, (int)d.y0, (int)d.y1).empty() && intersect(need.q, (int)d.y0 / 2, (int)d.y1 / 2).empty()) continue;
        waitFlag(g, s, F_READY_D, j, n + 1);
      }
    });
    indirectFinish(g, P, n, split);
    for (int j = 0; j < g->world; ++j) {
      const RankLayout& d = p->ranks[j];
      if (d.role == ROLE_D && !(intersect(need.g, (int)d.y0, (int)d.y1).empty() && intersect(need.q, (int)d.y0 / 2, (int)d.y1 / 2).empty())) setFlag(g, r->stream, j, F_ACK_D, n + 1);
      if (d.role == ROLE_I && j != g->rank) setFlag(g, r->stream, j, F_ACK_H, n + 1);
    }
    CUDA_CHECK(cudaEventRecord(p->evStage, r->stream)); CUDA_CHECK(cudaStreamWaitEvent(g->cs, p->evStage, 0)); CUDA_CHECK(cudaStreamWaitEvent(p->csH, p->evStage, 0));
    if (eagerHistory) pushHistory(g, P, set, false, n + 2, Range{(int)me.y0, (int)me.y1});
    p->historyComplete = eagerHistory;
    for (int j = 0; j < g->world; ++j) {
      const RankLayout& c = p->ranks[j];
      const Range ri = intersect(indirectNeeds(c, padded), (int)me.y0 / 2, (int)me.y1 / 2);
      if (ri.empty()) continue;
      if (prev >= 0) waitFlag(g, g->cs, F_ACK_I, j, (uint32_t)prev + 1);
      pushRowsFrom(g, j, B_INDIN0 + set, P.indA, rowG, ri);              // quarter-res rows at the full-res pitch (renderer.cpp:267-281)
      setFlag(g, g->cs, j, F_READY_I, n + 1);
    }
    CUDA_CHECK(cudaEventRecord(p->evPushI[set], g->cs)); p->pushIValid[set] = true;
    CUDA_CHECK(cudaEventRecord(p->evPushH[set], p->csH)); p->pushHValid[set] = true;
    endFrame(r);
    indirectDone(g, n);
  } else {
    // post rank (denoise + compose on its band), or — without indirect ranks — indirect_stage + denoise + compose on the whole frame
    const bool alsoIndirect = (me.role & ROLE_I) != 0;
    const DirectNeeds need = directNeeds(me, padded);
    auto waitDirect = [&](cudaStream_t s) {
      for (int j = 0; j < g->world; ++j) {
        const RankLayout& d = p->ranks[j];
        if (d.role != ROLE_D) continue;
        if (intersect(need.g, (int)d.y0, (int)d.y1).empty() && intersect(need.d, (int)d.y0, (int)d.y1).empty() && intersect(need.q, (int)d.y0 / 2, (int)d.y1 / 2).empty()) continue;
        waitFlag(g, s, F_READY_D, j, n + 1);
      }
    };
    bool split = false;
    if (alsoIndirect) {
      // the path tracing of this frame starts on its own stream as soon as the direct ranks' rows are here — also while the previous frame
      // is still being denoised on the render stream
      split = indirectTraceAsync(g, P, n, waitDirect);
      if (split) waitDirect(r->stream);             // (the denoiser reads the same rows)
    } else {
      waitDirect(r->stream);
      beginFrame(r);
    }
    const PostLayout L = postLayout(P, me.count > 1);
    markStart(r, EID_K_DENOISE_DIRECT, r->stream);
    stagePrep(r, P, L, r->stream);
    CUDA_CHECK(cudaEventRecord(p->evFork, r->stream)); CUDA_CHECK(cudaStreamWaitEvent(r->aux, p->evFork, 0));
    stageDenoiseDirect(r, P, L, r->aux);
    markStop(r, EID_K_DENOISE_DIRECT, r->aux);
    CUDA_CHECK(cudaEventRecord(p->evK3, r->aux));
    if (alsoIndirect) indirectFinish(g, P, n, split);
    else {
      const Range ni = indirectNeeds(me, padded);
      for (int j = 0; j < g->world; ++j) {
        const RankLayout& i = p->ranks[j];
        if (i.role == ROLE_I && !intersect(ni, (int)i.y0 / 2, (int)i.y1 / 2).empty()) waitFlag(g, r->stream, F_READY_I, j, n + 1);
      }
    }
    r->postStarted = true;
    markStart(r, EID_K_DENOISE_INDIRECT, r->stream);
    stageDenoiseIndirect(r, P, L, r->stream);
    markStop(r, EID_K_DENOISE_INDIRECT, r->stream);
    CUDA_CHECK(cudaStreamWaitEvent(r->stream, p->evK3, 0));
    stageCompose(r, P, L, r->stream);
    endFrame(r);
    if (alsoIndirect) indirectDone(g, n);
  }
  p->lastSeqOfParity[set] = (int)n;
  p->seq = n + 1;
  if (ackNow) p->ackPending = true;
  CUDA_CHECK(cudaGetLastError());
}

// Host delivery of the frame just enqueued (eid_group_render_host_async).  The composed rows only exist on the post ranks, and one PCIe
// link moves a 1080p image pair (66 MB) in ~1.6 ms — longer than a pipeline stage.  So the frame leaves through ALL links: rank k owns the
// delivery band [k B, (k + 1) B); a post rank copies its rows of its own band to the host itself and writes the rest into the owners'
// staging buffers over NVLink; every owner copies what it received to the (shared, pinned) host images on its copy stream.
void pipelineDeliver(eid_group* g, const RtxState& st, float* directHost, float* indirectHost) {
  eid_renderer* r = g->r;
  EidPipe* p = g->pipe;
  const int H = st.size.y, set = r->lastSet;
  const uint32_t d = p->dseq, par = d & 1u;
  const size_t pitchB = (size_t)r->width * 16, rowBytes = (size_t)st.size.x * 16, img = (size_t)p->delivRows * r->width;
  float* hostImg[2] = {directHost, indirectHost};
  auto delivBand = [&](int k) { return clampRange(k * (int)p->delivRows, (k + 1) * (int)p->delivRows, 0, H); };
  const bool post = (p->me.role & ROLE_P) != 0;
  if (post) {
    const float4* srcImg[2] = {r->directImg, r->indirectImg};
    CUDA_CHECK(cudaEventRecord(p->evDone, r->stream));
    CUDA_CHECK(cudaStreamWaitEvent(g->cs, p->evDone, 0));
    CUDA_CHECK(cudaStreamWaitEvent(r->copyStream, p->evDone, 0));
    for (int j = 0; j < g->world; ++j) {
      const Range b = delivBand(j);
      const Range rows = intersect(b, (int)p->me.y0, (int)p->me.y1);
      if (rows.empty()) continue;
      if (j == g->rank) {
        for (int k = 0; k < 2; ++k)
          if (hostImg[k]) CUDA_CHECK(cudaMemcpy2DAsync((char*)hostImg[k] + (size_t)rows.a * rowBytes, rowBytes, srcImg[k] + (size_t)rows.a * r->width, pitchB, rowBytes,
                                                        rows.b - rows.a, cudaMemcpyDeviceToHost, r->copyStream));
        continue;
      }
      if (d >= 2) waitFlag(g, g->cs, F_ACK_V, j, d - 1);              // the owner has sent the frame that used this staging parity last
      for (int k = 0; k < 2; ++k) {
        float4* dst = (float4*)p->peer[j][B_DELIV] + (2 * par + k) * img + (size_t)(rows.a - j * (int)p->delivRows) * r->width;
        CUDA_CHECK(cudaMemcpyAsync(dst, srcImg[k] + (size_t)rows.a * r->width, (size_t)(rows.b - rows.a) * pitchB, cudaMemcpyDeviceToDevice, g->cs));
        p->peerCopies++; p->peerBytes += (size_t)(rows.b - rows.a) * pitchB;
      }
      setFlag(g, g->cs, j, F_READY_V, d + 1);
    }
    CUDA_CHECK(cudaEventRecord(p->evDelivPush, g->cs));
  }
  // rows of MY delivery band that other post ranks composed
  const Range mine = delivBand(g->rank);
  for (int j = 0; j < g->world; ++j) {
    if (j == g->rank || !(p->ranks[j].role & ROLE_P)) continue;
    const Range rows = intersect(mine, (int)p->ranks[j].y0, (int)p->ranks[j].y1);
    if (rows.empty()) continue;
    waitFlag(g, r->copyStream, F_READY_V, j, d + 1);
    for (int k = 0; k < 2; ++k) {
      const float4* src = p->deliv + (2 * par + k) * img + (size_t)(rows.a - g->rank * (int)p->delivRows) * r->width;
      if (hostImg[k]) CUDA_CHECK(cudaMemcpy2DAsync((char*)hostImg[k] + (size_t)rows.a * rowBytes, rowBytes, src, pitchB, rowBytes, rows.b - rows.a, cudaMemcpyDeviceToHost, r->copyStream));
    }
    setFlag(g, r->copyStream, j, F_ACK_V, d + 1);
  }
  if (post) {
    // this parity's images (the producers' landing buffers) are free once the local copy AND the pushes have read them
    CUDA_CHECK(cudaStreamWaitEvent(r->copyStream, p->evDelivPush, 0));
    CUDA_CHECK(cudaEventRecord(r->evCopyDone2[set], r->copyStream));
    r->copyPending2[set] = true;
    pipelineAck(g, r->copyStream);
  }
  p->dseq = d + 1;
}

bool pipelineDelivers(eid_group* g, uint32_t* y0, uint32_t* y1) {
  EidPipe* p = g->pipe;
  if (!(p->me.role & ROLE_P)) return false;
  *y0 = p->me.y0; *y1 = p->me.y1;
  return true;
}

void pipelineSync(eid_group* g) {
  EidPipe* p = g->pipe;
  CUDA_CHECK(cudaStreamSynchronize(p->csP));
  CUDA_CHECK(cudaStreamSynchronize(p->csH));
  for (cudaStream_t st : p->k2Stream) if (st) CUDA_CHECK(cudaStreamSynchronize(st));
}

void pipelineInfo(eid_group* g, eid_group_info* out) {
  EidPipe* p = g->pipe;
  out->rank = g->rank; out->world = g->world;
  out->y0 = p->me.y0; out->y1 = p->me.y1; out->bandRows = p->me.y1 - p->me.y0;
  out->collectives = 0;
  out->stages = p->me.role; out->nDirect = p->nD; out->nIndirect = p->nI; out->nPost = p->nP;
  out->peerCopies = p->peerCopies; out->peerBytes = p->peerBytes;
  out->streamMemOps = p->waitValue ? 1 : 0;
}

void pipelineDestroy(eid_group* g) {
  EidPipe* p = g->pipe;
  if (!p) return;
  // nobody may free memory a peer can still write a flag into: every rank announces that its streams have drained and waits for the others
  if (p->shm) {
    storeRel(&p->shm->ranks[g->rank].closed, 1u);
    const double t0 = nowSec();
    for (int j = 0; j < g->world; ++j) while (!loadAcq(&p->shm->ranks[j].closed) && nowSec() - t0 < 20.0) usleep(200);
  }
  for (void* b : p->openedBases) cudaIpcCloseMemHandle(b);
  for (cudaStream_t* st : {&p->csP, &p->csH, &p->k2Stream[0], &p->k2Stream[1]}) if (*st) { cudaStreamSynchronize(*st); cudaStreamDestroy(*st); *st = nullptr; }
  for (cudaEvent_t e : {p->evTrace[0], p->evTrace[1], p->evFinish[0], p->evFinish[1]}) if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : {p->evPushP[0], p->evPushP[1], p->evPushH[0], p->evPushH[1]}) if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : {p->evStage, p->evPush[0], p->evPush[1], p->evPushI[0], p->evPushI[1], p->evPrep, p->evK3, p->evFork, p->evDelivPush, p->evDone}) if (e) cudaEventDestroy(e);
  cudaFree(p->flags); cudaFree(p->deliv);
  if (p->shm) munmap(p->shm, sizeof(ShmHdr));
  if (!p->shmPath.empty() && g->rank == 0) unlink(p->shmPath.c_str());
  delete p;
  g->pipe = nullptr;
}

extern "C" {

int eid_group_pipeline_layout(uint32_t height, int world, int rank, int n_direct, int n_indirect, int n_post, eid_pipeline_layout* out) {
  EID_TRY
  if (!out || !height || world < 2 || rank < 0 || rank >= world) raise(EID_ERR_INVALID, "eid_group_pipeline_layout: bad height / world (>= 2) / rank, or null output");
  int nD = n_direct, nI = n_indirect, nP = n_post;
  stageCounts(world, nD, nI, nP);
  const RankLayout L = rankLayout(height, rank, nD, nI, nP);
  memset(out, 0, sizeof(*out));
  out->nDirect = nD; out->nIndirect = nI; out->nPost = nP;
  out->stages = L.role; out->index = L.index; out->count = L.count; out->y0 = L.y0; out->y1 = L.y1;
  out->paddedHeight = paddedHeightOf(height, nD, nI, nP);
  return EID_OK;
  EID_CATCH
}

int eid_group_random_id(void* id128) {
  EID_TRY
  if (!id128) raise(EID_ERR_INVALID, "eid_group_random_id: null argument");
  const int fd = open("/dev/urandom", O_RDONLY);
  if (fd < 0 || read(fd, id128, 128) != 128) { if (fd >= 0) close(fd); raise(EID_ERR_IO, "cannot read /dev/urandom"); }
  close(fd);
  return EID_OK;
  EID_CATCH
}

int eid_group_create_pipeline(eid_group** out, eid_renderer* r, int rank, int world, const void* id128, uint32_t height, int n_direct, int n_indirect, int n_post) {
  EID_TRY
  if (!out || !r || !id128) raise(EID_ERR_INVALID, "eid_group_create_pipeline: null argument");
  if (world < 2 || rank < 0 || rank >= world) raise(EID_ERR_INVALID, "eid_group_create_pipeline: rank %d outside world %d (a pipeline needs at least 2 ranks)", rank, world);
  int nD = n_direct, nI = n_indirect, nP = n_post;
  stageCounts(world, nD, nI, nP);
  const uint32_t padded = paddedHeightOf(height, nD, nI, nP);
  if (!height || r->height < padded) raise(EID_ERR_INVALID, "eid_group_create_pipeline: the renderer's allocation height %u is below the padded height %u of this layout (eid_group_pipeline_layout)", r->height, padded);
  eid_group* g = new eid_group();
  EidPipe* p = new EidPipe();
  g->pipe = p;
  try {
    g->r = r; g->rank = rank; g->world = world;
    p->nD = nD; p->nI = nI; p->nP = nP; p->height = height; p->padded = padded;
    for (int j = 0; j < world; ++j) p->ranks[j] = rankLayout(height, j, nD, nI, nP);
    p->me = p->ranks[rank];
    g->bandRows = p->me.y1 - p->me.y0;
    CUDA_CHECK(cudaSetDevice(r->device));
    CUDA_CHECK(cudaStreamCreateWithFlags(&g->cs, cudaStreamNonBlocking));
    r->groupStream = g->cs;
    CUDA_CHECK(cudaStreamCreateWithFlags(&p->csP, cudaStreamNonBlocking));
    CUDA_CHECK(cudaStreamCreateWithFlags(&p->csH, cudaStreamNonBlocking));
    for (cudaEvent_t* e : {&p->evPushP[0], &p->evPushP[1], &p->evPushH[0], &p->evPushH[1]}) CUDA_CHECK(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    for (cudaEvent_t* e : {&p->evStage, &p->evPush[0], &p->evPush[1], &p->evPushI[0], &p->evPushI[1], &p->evPrep, &p->evK3, &p->evFork, &p->evDelivPush, &p->evDone}) CUDA_CHECK(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    if (p->me.role & ROLE_I) {
      for (int k = 0; k < 2; ++k) {
        CUDA_CHECK(cudaStreamCreateWithFlags(&p->k2Stream[k], cudaStreamNonBlocking));
        CUDA_CHECK(cudaEventCreateWithFlags(&p->evTrace[k], cudaEventDisableTiming)); CUDA_CHECK(cudaEventCreateWithFlags(&p->evFinish[k], cudaEventDisableTiming));
      }
      if (!r->shadowStream2) CUDA_CHECK(cudaStreamCreateWithFlags(&r->shadowStream2, cudaStreamNonBlocking));
      if (!r->evWave2) CUDA_CHECK(cudaEventCreateWithFlags(&r->evWave2, cudaEventDisableTiming));
      if (!r->evWaveJoin2) CUDA_CHECK(cudaEventCreateWithFlags(&r->evWaveJoin2, cudaEventDisableTiming));
    }
    p->delivRows = bandRowsOf(height, world);
    CUDA_CHECK(cudaMalloc((void**)&p->deliv, (size_t)4 * p->delivRows * r->width * 16));
    if (eid_renderer_set_band(r, p->me.y0, std::min(p->me.y1, r->height)) != EID_OK) raise(EID_ERR_INVALID, "%s", eid_last_error());
    // per-parity landing buffers of the pre-denoise indirect image (post ranks read them in place of denoiseIndTempA)
    const size_t nImg = (size_t)r->width * r->height * 16 + (size_t)17 * r->width * 16;
    for (int k = 0; k < 2; ++k) if (!r->indIn[k]) { CUDA_CHECK(cudaMalloc((void**)&r->indIn[k], nImg)); CUDA_CHECK(cudaMemsetAsync(r->indIn[k], 0, nImg, r->stream)); }
    CUDA_CHECK(cudaMalloc((void**)&p->flags, F_KINDS * MAXW * sizeof(uint32_t)));
    CUDA_CHECK(cudaMemsetAsync(p->flags, 0, F_KINDS * MAXW * sizeof(uint32_t), r->stream));
    CUDA_CHECK(cudaStreamSynchronize(r->stream));
    p->waitValue = getenv("EID_PIPE_NO_MEMOPS") ? nullptr : (WaitValueFn)driverEntry("cuStreamWaitValue32");
    AddrRangeFn addrRange = (AddrRangeFn)driverEntry("cuMemGetAddressRange");

    void* local[B_COUNT] = {r->gbuffer[0], r->gbuffer[1], r->directImgs[0], r->directImgs[1], r->k2G[0], r->k2G[1], r->k2Mv[0], r->k2Mv[1],
                            r->indIn[0], r->indIn[1], r->directResv[0], r->directResv[1], r->indirectResv[0], r->indirectResv[1], p->flags, p->deliv};
    for (int b = 0; b < B_COUNT; ++b) p->peer[rank][b] = local[b];

    // ---- rendezvous: publish this rank's IPC handles, wait for everybody's, map them ----
    char name[64];
    const unsigned char* id = (const unsigned char*)id128;
    snprintf(name, sizeof(name), "/dev/shm/eidola_pipe_%02x%02x%02x%02x%02x%02x%02x%02x%02x%02x%02x%02x", id[0] ^ id[12], id[1] ^ id[13], id[2] ^ id[14], id[3] ^ id[15],
             id[4], id[5], id[6], id[7], id[8], id[9], id[10], id[11]);
    p->shmPath = name;
    const int fd = open(name, O_CREAT | O_RDWR, 0600);
    if (fd < 0) raise(EID_ERR_IO, "cannot create the rendezvous file %s", name);
    if (ftruncate(fd, sizeof(ShmHdr)) != 0) { close(fd); raise(EID_ERR_IO, "cannot size the rendezvous file %s", name); }
    void* m = mmap(nullptr, sizeof(ShmHdr), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (m == MAP_FAILED) raise(EID_ERR_IO, "cannot map the rendezvous file %s", name);
    p->shm = (ShmHdr*)m;
    ShmRank& mine = p->shm->ranks[rank];
    mine.pid = (int32_t)getpid(); mine.device = r->device; mine.width = r->width; mine.height = r->height;
    for (int b = 0; b < B_COUNT; ++b) {
      CUdeviceptr base = (CUdeviceptr)(uintptr_t)local[b]; size_t size = 0;
      if (addrRange && addrRange(&base, &size, (CUdeviceptr)(uintptr_t)local[b]) != CUDA_SUCCESS) base = (CUdeviceptr)(uintptr_t)local[b];
      CUDA_CHECK(cudaIpcGetMemHandle(&mine.buf[b].h, (void*)(uintptr_t)base));
      mine.buf[b].offset = (unsigned long long)((uintptr_t)local[b] - (uintptr_t)base);
      mine.buf[b].bytes = size;
    }
    storeRel(&mine.ready, 1u);
    const double t0 = nowSec();
    const double timeout = getenv("EID_PIPE_TIMEOUT") ? atof(getenv("EID_PIPE_TIMEOUT")) : 120.0;
    for (int j = 0; j < world; ++j) {
      while (!loadAcq(&p->shm->ranks[j].ready)) {
        if (nowSec() - t0 > timeout) raise(EID_ERR_IO, "eid_group_create_pipeline: rank %d did not join within %.0f s", j, timeout);
        usleep(200);
      }
    }
    for (int j = 0; j < world; ++j) {
      if (j == rank) continue;
      const ShmRank& o = p->shm->ranks[j];
      if (o.pid == mine.pid) raise(EID_ERR_UNSUPPORTED, "the stage pipeline needs one PROCESS per rank (CUDA IPC cannot map memory of the same process)");
      if (o.width != r->width || o.height != r->height) raise(EID_ERR_INVALID, "rank %d allocated %ux%u, this rank %ux%u: every rank must use the padded height", j, o.width, o.height, r->width, r->height);
      for (int b = 0; b < B_COUNT; ++b) {
        const std::string key((const char*)&o.buf[b].h, sizeof(cudaIpcMemHandle_t));
        void* base = nullptr;
        for (auto& kv : p->openedByHandle) if (kv.first == key) base = kv.second;
        if (!base) {
          CUDA_CHECK(cudaIpcOpenMemHandle(&base, o.buf[b].h, cudaIpcMemLazyEnablePeerAccess));
          p->openedBases.push_back(base);
          p->openedByHandle.emplace_back(key, base);
        }
        p->peer[j][b] = (char*)base + o.buf[b].offset;
      }
    }
    storeRel(&mine.opened, 1u);
    for (int j = 0; j < world; ++j) {
      while (!loadAcq(&p->shm->ranks[j].opened)) {
        if (nowSec() - t0 > timeout) raise(EID_ERR_IO, "eid_group_create_pipeline: rank %d did not finish mapping within %.0f s", j, timeout);
        usleep(200);
      }
    }
    if (rank == 0) unlink(name);                              // every rank holds its mapping; the name can go
  } catch (...) { eid_group_destroy(g); throw; }
  *out = g;
  return EID_OK;
  EID_CATCH
}

}  // extern "C"

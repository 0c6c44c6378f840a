// group.h — the eid_group object shared by group.cu (row bands joined by NCCL all-gathers) and pipeline.cu (stage pipeline over CUDA-IPC
// peer mappings).
#pragma once
#include "renderer.h"

typedef void* NcclComm;
struct EidPipe;                                     // pipeline.cu

struct eid_group {
  eid_renderer* r = nullptr;
  int rank = 0, world = 1;
  NcclComm comm = nullptr;
  cudaStream_t cs = nullptr;                       // communication stream
  cudaEvent_t evA = nullptr, evB = nullptr, evX = nullptr, evC = nullptr, evD = nullptr, evH = nullptr, evA2 = nullptr, evPrep = nullptr, evK3 = nullptr;
  int post = 1;                                    // 1: denoise + compose per band (default), 0: replicated on every rank
  int history = 2;                                 // reservoir history across band edges: 0 never, 1 every frame (behind the post stages), 2 lazily when the camera moved
  int gatherFinal = 1;                             // exchange C
  bool historyComplete = false;                    // the LAST reservoirs of the next frame are complete on this rank
  bool histPending = false, finalPending = false;  // the render stream has not yet been ordered after the eager history gather / exchange C
  cudaEvent_t evHist = nullptr;
  uint32_t bandRows = 0;
  // host delivery of this rank's band
  cudaStream_t copyStream = nullptr;
  cudaEvent_t evFrameDone = nullptr, evCopyDone = nullptr;
  float4* staging[2] = {nullptr, nullptr};
  bool copyPending = false;
  unsigned long long collectives = 0;              // NCCL launches since creation
  EidPipe* pipe = nullptr;                         // non-null: this group is a stage pipeline (eid_group_create_pipeline)
};

// pipeline.cu
void pipelineFrame(eid_group* g, const RtxState& st, int frames, bool ackNow);
void pipelineAck(eid_group* g, cudaStream_t st);
void pipelineDestroy(eid_group* g);
void pipelineSync(eid_group* g);
void pipelineInfo(eid_group* g, eid_group_info* out);
bool pipelineDelivers(eid_group* g, uint32_t* y0, uint32_t* y1);
void pipelineDeliver(eid_group* g, const RtxState& st, float* directHost, float* indirectHost);

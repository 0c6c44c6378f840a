// k_wave.cu — translation unit of K2 in its wavefront form: k_gi_begin / k_gi_bounce / k_gi_finish + the ray-queue traversal kernel.
#include "stages.h"
#include "stage_wave.cuh"

namespace eid {

void launchGiBegin(const FrameParams& P, dim3 g, cudaStream_t st, bool tex) {
  if (tex) k_gi_begin<true><<<g, dim3(8, 8), 0, st>>>(P); else k_gi_begin<false><<<g, dim3(8, 8), 0, st>>>(P);
}
void launchGiBounce(const FrameParams& P, int blocks, cudaStream_t st, bool tex, int d) {
  if (tex) k_gi_bounce<true><<<blocks, 128, 0, st>>>(P, d); else k_gi_bounce<false><<<blocks, 128, 0, st>>>(P, d);
}
void launchGiFinish(const FrameParams& P, dim3 g, cudaStream_t st) { k_gi_finish<<<g, dim3(8, 8), 0, st>>>(P); }

void launchTraceQueue(bool any, bool stats, int g, cudaStream_t st, const AccelView& A, const float4* rays, const uint32_t* count,
                      uint32_t* cursor, float4* hits, uint32_t* occl, unsigned long long* counters, unsigned long long* totals) {
  if (any) {
    if (stats) k_trace_queue<true, true><<<g, 128, 0, st>>>(A, rays, count, cursor, hits, occl, counters, totals);
    else k_trace_queue<true, false><<<g, 128, 0, st>>>(A, rays, count, cursor, hits, occl, counters, totals);
  } else {
    if (stats) k_trace_queue<false, true><<<g, 128, 0, st>>>(A, rays, count, cursor, hits, occl, counters, totals);
    else k_trace_queue<false, false><<<g, 128, 0, st>>>(A, rays, count, cursor, hits, occl, counters, totals);
  }
}

}  // namespace eid

// k_indirect.cu — translation unit of K2 in its one-thread-per-pixel form (indirect_stage.comp).
#include "stages.h"
#include "stage_indirect.cuh"

namespace eid {

void launchIndirectMega(const FrameParams& P, dim3 g, cudaStream_t st, bool stats, bool tex) {
  const dim3 b(8, 8);
  if (stats) { if (tex) k_indirect_stage<true, true><<<g, b, 0, st>>>(P); else k_indirect_stage<true, false><<<g, b, 0, st>>>(P); }
  else { if (tex) k_indirect_stage<false, true><<<g, b, 0, st>>>(P); else k_indirect_stage<false, false><<<g, b, 0, st>>>(P); }
}

}  // namespace eid

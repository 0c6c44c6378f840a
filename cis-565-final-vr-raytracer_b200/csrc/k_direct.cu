// k_direct.cu — translation unit of K1 (direct_stage.comp): kernel instantiations + launchers.
#include "stages.h"
#include "stage_direct.cuh"

namespace eid {

template <bool STATS, bool TEX>
static void directVariant(const FrameParams& P, dim3 g, cudaStream_t st, bool spatial, int halo) {
  const dim3 b(8, 8);
  if (spatial) k_direct_stage<STATS, TEX, true><<<g, b, 0, st>>>(P, halo);
  else k_direct_stage<STATS, TEX, false><<<g, b, 0, st>>>(P, halo);
}

void launchDirectStage(const FrameParams& P, dim3 g, cudaStream_t st, bool stats, bool tex, bool spatial, int halo) {
  if (stats) { if (tex) directVariant<true, true>(P, g, st, spatial, halo); else directVariant<true, false>(P, g, st, spatial, halo); }
  else { if (tex) directVariant<false, true>(P, g, st, spatial, halo); else directVariant<false, false>(P, g, st, spatial, halo); }
}

void launchDirectSplit(const FrameParams& P, dim3 g, cudaStream_t st, bool stats, bool tex) {
  const dim3 b(8, 8);
  if (stats) { if (tex) k_direct_gen<true, true><<<g, b, 0, st>>>(P); else k_direct_gen<true, false><<<g, b, 0, st>>>(P); }
  else { if (tex) k_direct_gen<false, true><<<g, b, 0, st>>>(P); else k_direct_gen<false, false><<<g, b, 0, st>>>(P); }
  k_direct_reuse<<<g, b, 0, st>>>(P);
}

void launchDirectSpatial(const FrameParams& P, dim3 g, cudaStream_t st) { k_direct_spatial<<<g, dim3(8, 8), 0, st>>>(P); }

}  // namespace eid

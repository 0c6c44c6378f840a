// stages.h — host-callable launchers of the stage kernels.  Every kernel family lives in its own translation unit
// (k_direct.cu, k_indirect.cu, k_wave.cu, k_post.cu) so that they compile in parallel; render.cu owns the Renderer
// object, the launch schedule and the C-ABI and only sees these declarations.
#pragma once
#include <cuda.h>
#include "frame.cuh"

namespace eid {

// K1 — direct_stage.comp (k_direct.cu)
void launchDirectStage(const FrameParams& P, dim3 grid, cudaStream_t st, bool stats, bool tex, bool spatial, int halo);
void launchDirectSpatial(const FrameParams& P, dim3 grid, cudaStream_t st);
void launchDirectSplit(const FrameParams& P, dim3 grid, cudaStream_t st, bool stats, bool tex);   // direct_gen.comp + direct_reuse.comp

// K2 — indirect_stage.comp, one thread per pixel (k_indirect.cu)
void launchIndirectMega(const FrameParams& P, dim3 grid, cudaStream_t st, bool stats, bool tex);

// K2 — wavefront form (k_wave.cu)
void launchGiBegin(const FrameParams& P, dim3 grid, cudaStream_t st, bool tex);
void launchGiBounce(const FrameParams& P, int blocks, cudaStream_t st, bool tex, int depth);
void launchGiFinish(const FrameParams& P, dim3 grid, cudaStream_t st);
void launchTraceQueue(bool any, bool stats, int blocks, cudaStream_t st, const AccelView& A, const float4* rays, const uint32_t* count,
                      uint32_t* cursor, float4* hits, uint32_t* occl, unsigned long long* counters, unsigned long long* totals);

// K3 / K4 / K5 + display pass + parity taps (k_post.cu)
void launchDenoisePrep(const FrameParams& P, dim3 grid, cudaStream_t st, int first, int stride, int rows, bool fastFull, bool fastQuarter);
void launchBilateral(bool indirect, bool strict, dim3 grid, cudaStream_t st, const FrameParams& P, const float4* src, float4* dst, int first, int stride, int rows);
#define EID_TILE_W 32                  // lattice points per tile row = lanes of a warp
#define EID_TILE_PW (EID_TILE_W + 4)   // + 2-point halo on both sides
// one level of the shared-memory tile A-Trous kernel (stage_denoise.cuh)
struct AtrousArgs {
  const float4* gPos; const float4* gNrm; const float4* inImg;   // plain pointers of the three planes (cp.async loader, centre-free paths)
  float4* outImg;
  int gPitch, iPitch;        // pitch of the geometry planes / of the colour images, in texels
  int allocRows;             // rows of the allocation (cp.async loader: nothing beyond is touched)
  int level, lastLevel;
  int first, stride, rows;   // stripe layout of the OUTPUT rows (kernel's own resolution), see stripeRow
  int nTy;                   // tile rows per stripe (upper bound; surplus blocks exit)
  int useTma;                // 1: TMA tile loads, 0: cp.async (LDGSTS) tile loads (A/B + fallback when no tensor map could be encoded)
};
void launchAtrousTile(bool indirect, bool strict, int rowBlock, dim3 grid, cudaStream_t st, const FrameParams& P, const CUtensorMap& mapPos,
                      const CUtensorMap& mapNrm, const CUtensorMap& mapCol, const AtrousArgs& args);
void launchDenoise(bool indirect, bool strict, int rowBlock, dim3 grid, cudaStream_t st, const FrameParams& P, const float4* src, float4* dst,
                   int level, int lastLevel, int first, int stride, int rows);
void launchCompose(const FrameParams& P, dim3 grid, cudaStream_t st, const float4* indSrc, int first, int stride, int rows);
void launchMipBlit(dim3 grid, cudaStream_t st, const float4* src, int sw, int sh, int spitch, float4* dst, int dw, int dh);
void launchPost(const FrameParams& P, dim3 grid, cudaStream_t st, const Tonemapper& tm, float4* outF, uchar4* out8, const float4* avg);
void launchFnTap(int which, int ni, int no, const float* in, uint32_t n, float* out);
void launchCtxTap(const FrameParams& P, int which, int ni, int no, const float* in, uint32_t n, float* out);
void launchSunAndSky(const SunAndSky& ss, const float* dirs, uint32_t n, float* out);

}  // namespace eid

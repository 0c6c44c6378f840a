// k_post.cu — translation unit of K3 / K4 (A-Trous denoisers), K5 (compose), the display pass and the parity taps.
#include "stages.h"
#include "stage_post.cuh"
#include "stage_denoise.cuh"
#include "taps.cuh"

namespace eid {

void launchDenoisePrep(const FrameParams& P, dim3 g, cudaStream_t st, int first, int stride, int rows, bool fastFull, bool fastQuarter) {
  const dim3 b(32, 8);
  if (fastFull) { if (fastQuarter) k_denoise_prep<true, true><<<g, b, 0, st>>>(P, first, stride, rows); else k_denoise_prep<true, false><<<g, b, 0, st>>>(P, first, stride, rows); }
  else { if (fastQuarter) k_denoise_prep<false, true><<<g, b, 0, st>>>(P, first, stride, rows); else k_denoise_prep<false, false><<<g, b, 0, st>>>(P, first, stride, rows); }
}
void launchBilateral(bool indirect, bool strict, dim3 g, cudaStream_t st, const FrameParams& P, const float4* src, float4* dst, int first, int stride, int rows) {
  const dim3 b(32, 4);
  if (indirect) { if (strict) k_bilateral<true, true><<<g, b, 0, st>>>(P, src, dst, first, stride, rows); else k_bilateral<true, false><<<g, b, 0, st>>>(P, src, dst, first, stride, rows); }
  else { if (strict) k_bilateral<false, true><<<g, b, 0, st>>>(P, src, dst, first, stride, rows); else k_bilateral<false, false><<<g, b, 0, st>>>(P, src, dst, first, stride, rows); }
}

template <bool INDIRECT, bool STRICT>
static void atrousVariant(int R, dim3 g, cudaStream_t st, const FrameParams& P, const CUtensorMap& mp, const CUtensorMap& mn, const CUtensorMap& mc,
                          const AtrousArgs& a) {
  const dim3 b(32, 4);
  if (R == 4) k_atrous_tile<INDIRECT, STRICT, 4><<<g, b, 0, st>>>(P, mp, mn, mc, a);
  else k_atrous_tile<INDIRECT, STRICT, 2><<<g, b, 0, st>>>(P, mp, mn, mc, a);
}
void launchAtrousTile(bool indirect, bool strict, int R, dim3 g, cudaStream_t st, const FrameParams& P, const CUtensorMap& mp, const CUtensorMap& mn,
                      const CUtensorMap& mc, const AtrousArgs& a) {
  if (indirect) { if (strict) atrousVariant<true, true>(R, g, st, P, mp, mn, mc, a); else atrousVariant<true, false>(R, g, st, P, mp, mn, mc, a); }
  else { if (strict) atrousVariant<false, true>(R, g, st, P, mp, mn, mc, a); else atrousVariant<false, false>(R, g, st, P, mp, mn, mc, a); }
}

template <bool INDIRECT, bool STRICT>
static void denoiseVariant(int R, dim3 g, cudaStream_t st, const FrameParams& P, const float4* src, float4* dst, int level, int lastLevel,
                           int first, int stride, int rows) {
  const dim3 b(32, 4);
  if (R == 4) k_denoise<INDIRECT, STRICT, 4><<<g, b, 0, st>>>(P, src, dst, level, lastLevel, first, stride, rows);
  else if (R == 2) k_denoise<INDIRECT, STRICT, 2><<<g, b, 0, st>>>(P, src, dst, level, lastLevel, first, stride, rows);
  else k_denoise<INDIRECT, STRICT, 1><<<g, b, 0, st>>>(P, src, dst, level, lastLevel, first, stride, rows);
}
void launchDenoise(bool indirect, bool strict, int R, dim3 g, cudaStream_t st, const FrameParams& P, const float4* src, float4* dst,
                   int level, int lastLevel, int first, int stride, int rows) {
  if (indirect) { if (strict) denoiseVariant<true, true>(R, g, st, P, src, dst, level, lastLevel, first, stride, rows); else denoiseVariant<true, false>(R, g, st, P, src, dst, level, lastLevel, first, stride, rows); }
  else { if (strict) denoiseVariant<false, true>(R, g, st, P, src, dst, level, lastLevel, first, stride, rows); else denoiseVariant<false, false>(R, g, st, P, src, dst, level, lastLevel, first, stride, rows); }
}

void launchCompose(const FrameParams& P, dim3 g, cudaStream_t st, const float4* indSrc, int first, int stride, int rows) {
  k_compose<<<g, dim3(32, 8), 0, st>>>(P, indSrc, first, stride, rows);
}
void launchMipBlit(dim3 g, cudaStream_t st, const float4* src, int sw, int sh, int spitch, float4* dst, int dw, int dh) {
  k_mip_blit<<<g, dim3(32, 8), 0, st>>>(src, sw, sh, spitch, dst, dw, dh);
}
void launchPost(const FrameParams& P, dim3 g, cudaStream_t st, const Tonemapper& tm, float4* outF, uchar4* out8, const float4* avg) {
  k_post<<<g, dim3(32, 8), 0, st>>>(P, tm, outF, out8, avg);
}
void launchFnTap(int which, int ni, int no, const float* in, uint32_t n, float* out) { k_fn_tap<<<(n + 63) / 64, 64>>>(which, ni, no, in, n, out); }
void launchCtxTap(const FrameParams& P, int which, int ni, int no, const float* in, uint32_t n, float* out) {
  k_ctx_tap<<<(n + 63) / 64, 64>>>(P, which, ni, no, in, n, out);
}
void launchSunAndSky(const SunAndSky& ss, const float* dirs, uint32_t n, float* out) { k_sun_and_sky<<<(n + 63) / 64, 64>>>(ss, dirs, n, out); }

}  // namespace eid

/*
 * oracle/ref_shim/ref_post.cpp — TEST INFRASTRUCTURE.
 * The reference's three post stages — shaders/denoise_direct.comp, denoise_indirect.comp, compose.comp with denoise_common.glsl and
 * compress.glsl — compiled WHOLE (their main() included) as C++ from the transliterations of glsl_prep.py, and run over a frame with
 * the dispatch schedule of Renderer::run (src/renderer.cpp:178-205).  What a shader build binds is provided here as plain globals: the
 * storage images (imageLoad / imageStore on arrays; out-of-bounds loads return 0, stores are dropped — Vulkan robust access), the
 * push constant, the camera UBO, gl_GlobalInvocationID.  Every expression evaluated is the reference's own text.
 */
#include <cmath>
#include <cstdint>
#include <cstring>
#include "glsl/glsl_builtins.h"
#include "host_device.h"          // /root/reference/shaders/host_device.h (C++ branch)
// REF_VARIANT: a second build with DENOISER_DIRECT_BILATERAL / DENOISER_INDIRECT_BILATERAL = 1 (host_device.h:28-29): the bilateralFilter
// branches of denoise_direct.comp / denoise_indirect.comp and the one-dispatch schedule of renderer.cpp:186-188, 199-201; exported with
// the suffix _variant.  `which` of ref_post_run_variant: bit 0 = the direct stage is the bilateral build, bit 1 = the indirect one
// (the other stage then runs its A-Trous levels from the regular build's schedule — see the wrapper at the end of this file).
#ifdef REF_VARIANT
#undef DENOISER_DIRECT_BILATERAL
#undef DENOISER_INDIRECT_BILATERAL
#define DENOISER_DIRECT_BILATERAL 1
#define DENOISER_INDIRECT_BILATERAL 1
#define refpost refpost_v
#define refpost_compress refpost_compress_v
#define ref_post_run ref_post_run_variant_all
#define ref_post_dispatch ref_post_dispatch_variant
#define shared static
#endif
namespace refpost_compress {      // (its non-inline functions also exist in ref_wrap.cpp's translation unit)
#include "compress.glsl"          // /root/reference/shaders/compress.glsl: decompress_unit_vec as the shaders see it
}
#undef M_PI
#undef M_PI_2
#undef M_PI_4
#undef INFINITY
#undef PI

namespace refpost {
using orc::vec2; using orc::vec3; using orc::vec4; using orc::ivec2; using orc::uvec4;
using refpost_compress::decompress_unit_vec;

struct image2D { vec4* data; int w, h, pitch; };
struct uimage2D { uvec4* data; int w, h, pitch; };
static vec4 imageLoad(const image2D& im, ivec2 c) { return (c.x < 0 || c.y < 0 || c.x >= im.w || c.y >= im.h) ? vec4() : im.data[(size_t)c.y * im.pitch + c.x]; }
static uvec4 imageLoad(const uimage2D& im, ivec2 c) { return (c.x < 0 || c.y < 0 || c.x >= im.w || c.y >= im.h) ? uvec4() : im.data[(size_t)c.y * im.pitch + c.x]; }
static void imageStore(const image2D& im, ivec2 c, vec4 v) { if (c.x >= 0 && c.y >= 0 && c.x < im.w && c.y < im.h) im.data[(size_t)c.y * im.pitch + c.x] = v; }
struct GlobalId { unsigned int x, y, z; ivec2 xy() const { return ivec2((int)x, (int)y); } };

// layouts.glsl bindings used by the post stages + the push constant
static uimage2D thisGbuffer;
static image2D thisDirectResultImage, thisIndirectResultImage, denoiseDirTempA, denoiseDirTempB, denoiseIndTempA, denoiseIndTempB;
static RtxState rtxState;
static SceneCamera sceneCamera;
static GlobalId gl_GlobalInvocationID, gl_LocalInvocationID;

#include "../_ref/gen/globals.hpp"
#include "../_ref/gen/common_post.hpp"
namespace dd {
#include "../_ref/gen/denoise_common.hpp"
#include "../_ref/gen/denoise_direct.hpp"
}
#undef DENOISE_COMMON_GLSL
namespace di {
#include "../_ref/gen/denoise_common.hpp"
#include "../_ref/gen/denoise_indirect.hpp"
}
namespace cp {
#include "../_ref/gen/compose.hpp"
}

template <class F>
static void dispatch(int w, int h, F&& mainFn) {   // vkCmdDispatch(CEIL_DIV(w, 8), CEIL_DIV(h, 8), 1) of 8x8 groups
  const int gw = (w + 7) / 8 * 8, gh = (h + 7) / 8 * 8;
  for (int y = 0; y < gh; ++y)
    for (int x = 0; x < gw; ++x) { gl_GlobalInvocationID = GlobalId{(unsigned)x, (unsigned)y, 0u}; gl_LocalInvocationID = GlobalId{(unsigned)(x & 7), (unsigned)(y & 7), 0u}; mainFn(); }
}
}  // namespace refpost

using namespace refpost;

// Renderer::run, renderer.cpp:178-205: denoise_direct x4, denoise_indirect x5 (denoiseLevel = i), compose with the caller's state
extern "C" __attribute__((visibility("default")))
void ref_post_run(const RtxState* st, const SceneCamera* cam, int allocW, int allocH, void* gbuffer, void* direct, void* indirect,
                  void* dirA, void* dirB, void* indA, void* indB) {
  thisGbuffer = uimage2D{(uvec4*)gbuffer, allocW, allocH, allocW};
  auto img = [&](void* p) { return image2D{(vec4*)p, allocW, allocH, allocW}; };
  thisDirectResultImage = img(direct); thisIndirectResultImage = img(indirect);
  denoiseDirTempA = img(dirA); denoiseDirTempB = img(dirB); denoiseIndTempA = img(indA); denoiseIndTempB = img(indB);
  sceneCamera = *cam;
  rtxState = *st;
  const int W = st->size.x, H = st->size.y;
#ifdef REF_VARIANT
  if (st->denoise > 0) dispatch(W, H, [] { dd::main(); });            // renderer.cpp:186-188: ONE dispatch, the caller's push constants
  if (st->denoise > 0) dispatch(W / 2, H / 2, [] { di::main(); });    // renderer.cpp:199-201
#else
  if (st->denoise > 0)
    for (int i = 0; i < 4; ++i) { rtxState.denoiseLevel = i; dispatch(W, H, [] { dd::main(); }); }
  if (st->denoise > 0)
    for (int i = 0; i < 5; ++i) { rtxState.denoiseLevel = i; dispatch(W / 2, H / 2, [] { di::main(); }); }
#endif
  rtxState = *st;
  dispatch(W, H, [] { cp::main(); });
}

// ONE dispatch of a post stage, as a command buffer replays it: stage = the shader tag of ref_renderer.cpp (5 denoise_direct, 6 denoise_indirect,
// 7 compose), st = the push constants bound at that point, gx x gy work groups of 8 x 8 invocations (vkCmdDispatch(gx, gy, 1)).
extern "C" __attribute__((visibility("default")))
int ref_post_dispatch(int stage, const RtxState* st, const SceneCamera* cam, int allocW, int allocH, int gx, int gy, void* gbuffer, void* direct,
                      void* indirect, void* dirA, void* dirB, void* indA, void* indB) {
  thisGbuffer = uimage2D{(uvec4*)gbuffer, allocW, allocH, allocW};
  auto img = [&](void* p) { return image2D{(vec4*)p, allocW, allocH, allocW}; };
  thisDirectResultImage = img(direct); thisIndirectResultImage = img(indirect);
  denoiseDirTempA = img(dirA); denoiseDirTempB = img(dirB); denoiseIndTempA = img(indA); denoiseIndTempB = img(indB);
  sceneCamera = *cam;
  rtxState = *st;
  void (*fn)() = stage == 5 ? (void (*)())[] { dd::main(); } : stage == 6 ? (void (*)())[] { di::main(); } : stage == 7 ? (void (*)())[] { cp::main(); } : nullptr;
  if (!fn) return -1;
  for (int y = 0; y < gy * 8; ++y)
    for (int x = 0; x < gx * 8; ++x) { gl_GlobalInvocationID = GlobalId{(unsigned)x, (unsigned)y, 0u}; gl_LocalInvocationID = GlobalId{(unsigned)(x & 7), (unsigned)(y & 7), 0u}; fn(); }
  return 0;
}

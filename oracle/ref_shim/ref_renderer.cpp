/*
 * oracle/ref_shim/ref_renderer.cpp — TEST INFRASTRUCTURE.
 * Runs the reference's OWN src/renderer.cpp — compiled where it lies against the stand-ins of scene/scene_shim.h, whose Vulkan "device"
 * records instead of executing — and reports what Renderer::create / update / run do: the resources they allocate (kind, byte size,
 * extent, format), which resource each binding of the two S_RAYQ descriptor sets names (the ping-pong of renderer.cpp:341-375), and
 * per frame the exact command sequence of Renderer::run (renderer.cpp:154-206): descriptor set bound, every push-constant upload
 * (the RtxState bytes), every pipeline bind (shader tag 1..7 = direct_stage, direct_gen, direct_reuse, indirect_stage, denoise_direct,
 * denoise_indirect, compose), every dispatch with its group counts.  The tests hold the oracle's and the product's schedule against it.
 * The same for src/render_output.cpp (the display pass): the four result images it allocates, how its two S_OUT descriptor sets name them,
 * what RenderOutput::run pushes / binds / draws (tag 9 = post.frag) and which mip chains RenderOutput::genMipmap asks for.
 */
#include "scene_shim.h"
#include "shaders/host_device.h"      // /root/reference/shaders/host_device.h
#define private public
#include "renderer.hpp"                // /root/reference/src/renderer.hpp
#include "render_output.hpp"           // /root/reference/src/render_output.hpp
#undef private

#define REF_API extern "C" __attribute__((visibility("default")))
#include "../_ref/gen/defaults.hpp"   // the reference's own initialisers of m_rtxState, m_sunAndSky (sample_example.hpp:154-203), m_tm, m_depthTm (render_output.hpp:44-60)
// which: 0 RtxState, 1 SunAndSky, 2 Tonemapper m_tm, 3 Tonemapper m_depthTm -> bytes copied
REF_API int ref_default_state(int which, void* out, int cap) {
  const void* p = which == 0 ? (const void*)&ref_m_rtxState : which == 1 ? (const void*)&ref_m_sunAndSky : which == 2 ? (const void*)&ref_m_tm : which == 3 ? (const void*)&ref_m_depthTm : nullptr;
  const int n = which == 0 ? (int)sizeof(RtxState) : which == 1 ? (int)sizeof(SunAndSky) : (int)sizeof(Tonemapper);
  if (!p || cap < n) return -1;
  memcpy(out, p, n);
  return n;
}
// SampleExample::loadScene / loadEnvironmentHdr's derived RtxState fields (sample_example.cpp:87, 104-105), the lifted statements: out = lightLuminIntegInv,
// fireflyClampThreshold, envMapLuminIntegInv
REF_API void ref_glue_state(float trigWeight, float puncWeight, float envIntegral, float* out3) {
  RtxState st = ref_m_rtxState;
  ref_glue(st, RefGlueScene{trigWeight, puncWeight}, RefGlueSky{envIntegral});
  out3[0] = st.lightLuminIntegInv; out3[1] = st.fireflyClampThreshold; out3[2] = st.envMapLuminIntegInv;
}
struct RefRenderer { nvvk::ResourceAllocator alloc; Renderer r; };

REF_API void* ref_renderer_create(unsigned w, unsigned h) {
  RefRenderer* rr = new RefRenderer;
  rr->r.setup(nullptr, nullptr, 0u, &rr->alloc, 1u);
  rr->r.create(VkExtent2D{w, h}, {}, nullptr);
  return rr;
}
REF_API void ref_renderer_update(void* h, unsigned w, unsigned hh) { ((RefRenderer*)h)->r.update(VkExtent2D{w, hh}); }
REF_API void ref_renderer_destroy(void* h) { delete (RefRenderer*)h; }

static int resId(VkImageView v) { return v ? ((ShimResource*)v)->id : -1; }
static int resId(const nvvk::Buffer& b) { return b.buffer && b.buffer->res ? b.buffer->res->id : -1; }
// resource ids in a fixed order: gbuffer[0], gbuffer[1], directReservoir[0], [1], indirectReservoir[0], [1], directTempResv, indirectTempResv,
// motionVector, denoiseTempBuf[0..3]; then per resource id (kind, bytes, width, height, format) through ref_renderer_resource
REF_API void ref_renderer_roles(void* h, int* out13) {
  Renderer& r = ((RefRenderer*)h)->r;
  int k = 0;
  out13[k++] = resId(r.m_gbuffer[0].descriptor.imageView); out13[k++] = resId(r.m_gbuffer[1].descriptor.imageView);
  out13[k++] = resId(r.m_directReservoir[0]); out13[k++] = resId(r.m_directReservoir[1]);
  out13[k++] = resId(r.m_indirectReservoir[0]); out13[k++] = resId(r.m_indirectReservoir[1]);
  out13[k++] = resId(r.m_directTempResv); out13[k++] = resId(r.m_indirectTempResv);
  out13[k++] = resId(r.m_motionVector.descriptor.imageView);
  for (int i = 0; i < 4; ++i) out13[k++] = resId(r.m_denoiseTempBuf[i].descriptor.imageView);
}
REF_API int ref_renderer_resource(int id, long long* out5) {
  auto& d = ShimDevice::get();
  if (id < 0 || id >= (int)d.resources.size()) return -1;
  const ShimResource& s = *d.resources[id];
  out5[0] = s.kind; out5[1] = (long long)s.bytes; out5[2] = s.width; out5[3] = s.height; out5[4] = s.format;
  return 0;
}
// the descriptor writes of the LAST updateDescriptorSet (2 sets x 13 bindings): rows of (set number 1|2, binding, resource id, range bytes)
static int setNumber(const Renderer& r, int handle) { return handle == (int)(intptr_t)r.m_descSet[0] ? 1 : handle == (int)(intptr_t)r.m_descSet[1] ? 2 : -1; }   // 1 = m_descSet[0], 2 = m_descSet[1]
REF_API int ref_renderer_wiring(void* h, long long* out, int capRows) {
  const Renderer& r = ((RefRenderer*)h)->r;
  auto& w = ShimDevice::get().writes;
  const int n = (int)w.size() < 26 ? (int)w.size() : 26;
  for (int i = 0; i < n && i < capRows; ++i) {
    const auto& x = w[w.size() - n + i];
    out[4 * i] = setNumber(r, x.set); out[4 * i + 1] = x.binding; out[4 * i + 2] = x.resource; out[4 * i + 3] = (long long)x.range;
  }
  return n;
}
// Renderer::run for one frame: rows of (what, a, b, c) — 1 bind sets (first, count, number of the last set), 2 push constants (offset, size, index
// into pushBytes), 3 bind pipeline (shader tag), 4 dispatch (x, y, z) — and the pushed bytes, one RtxState after the other
REF_API int ref_renderer_run(void* h, const void* state, int frames, int* rows, int capRows, unsigned char* pushBytes, int capPush) {
  auto& d = ShimDevice::get();
  d.log.clear();
  nvvk::ProfilerVK prof;
  ((RefRenderer*)h)->r.run(nullptr, *(const RtxState*)state, prof, {}, frames);
  int n = 0, np = 0;
  for (const ShimEvent& e : d.log) {
    if (n >= capRows) break;
    rows[4 * n] = e.what; rows[4 * n + 1] = e.a; rows[4 * n + 2] = e.b; rows[4 * n + 3] = e.c;
    if (e.what == 1) rows[4 * n + 3] = setNumber(((RefRenderer*)h)->r, e.c);
    if (e.what == 2) {
      rows[4 * n + 3] = np;
      if ((np + 1) * (int)sizeof(RtxState) <= capPush && e.data.size() == sizeof(RtxState)) memcpy(pushBytes + np * sizeof(RtxState), e.data.data(), sizeof(RtxState));
      ++np;
    }
    ++n;
  }
  return n;
}

// ---- RenderOutput (src/render_output.cpp) --------------------------------------------------------------------------------------
struct RefOutput { nvvk::ResourceAllocator alloc; RenderOutput o; };
REF_API void* ref_output_create(unsigned w, unsigned h) {
  RefOutput* r = new RefOutput;
  r->o.setup(nullptr, nullptr, 0u, &r->alloc, 1u);
  r->o.create(VkExtent2D{w, h}, nullptr);
  return r;
}
REF_API void ref_output_destroy(void* h) { delete (RefOutput*)h; }
// resource ids of m_directResult[0], [1], m_indirectResult[0], [1] (details through ref_renderer_resource; out[4..7] = their mip counts)
REF_API void ref_output_roles(void* h, int* out8) {
  RenderOutput& o = ((RefOutput*)h)->o;
  const int ids[4] = {resId(o.m_directResult[0].descriptor.imageView), resId(o.m_directResult[1].descriptor.imageView),
                      resId(o.m_indirectResult[0].descriptor.imageView), resId(o.m_indirectResult[1].descriptor.imageView)};
  for (int i = 0; i < 4; ++i) { out8[i] = ids[i]; out8[4 + i] = ids[i] >= 0 ? ShimDevice::get().resources[ids[i]]->mips : -1; }
}
static int outSetNumber(const RenderOutput& o, int handle) { return handle == (int)(intptr_t)o.m_postDescSet[0] ? 1 : handle == (int)(intptr_t)o.m_postDescSet[1] ? 2 : -1; }
// the 12 descriptor writes of createPostDescriptor: rows of (set number 1|2, binding of OutputBindings, resource id)
REF_API int ref_output_wiring(void* h, int* out, int capRows) {
  const RenderOutput& o = ((RefOutput*)h)->o;
  auto& w = ShimDevice::get().writes;
  int n = 0;
  for (const auto& x : w) {
    const int sn = outSetNumber(o, x.set);
    if (sn < 0 || n >= capRows) continue;
    out[3 * n] = sn; out[3 * n + 1] = x.binding; out[3 * n + 2] = x.resource; ++n;
  }
  return n;
}
// RenderOutput::genMipmap (when genMips) then RenderOutput::run: rows of (what, a, b, c) — 5 generate mipmaps (resource id, levels, layers), 2 push
// (offset, size, 0), 3 bind pipeline (tag), 1 bind sets (first, count, set number), 6 draw (vertices, instances, first) — and the pushed bytes
REF_API int ref_output_run(void* h, const void* state, float zoom, float ratioX, float ratioY, int frames, int genMips, int* rows, int capRows, unsigned char* push, int capPush) {
  auto& d = ShimDevice::get();
  d.log.clear();
  RenderOutput& o = ((RefOutput*)h)->o;
  if (genMips) o.genMipmap(nullptr);
  o.run(nullptr, *(const RtxState*)state, zoom, vec2(ratioX, ratioY), frames);
  int n = 0;
  for (const ShimEvent& e : d.log) {
    if (n >= capRows) break;
    rows[4 * n] = e.what; rows[4 * n + 1] = e.a; rows[4 * n + 2] = e.b; rows[4 * n + 3] = e.c;
    if (e.what == 1) rows[4 * n + 3] = outSetNumber(o, e.c);
    if (e.what == 2 && (int)e.data.size() <= capPush) memcpy(push, e.data.data(), e.data.size());
    ++n;
  }
  return n;
}

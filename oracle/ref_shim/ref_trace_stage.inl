// oracle/ref_shim/ref_trace_stage.inl — TEST INFRASTRUCTURE.  Included once per trace stage (inside its namespace) by ref_trace.cpp, after
// globals.glsl: the stage's private globals, the stand-in of traceray_rq.glsl, then the reference's include chain of pathtrace.glsl.
PtPayload prd;
ShadowHitPayload shadow_payload;
ivec2 imageCoords;
#include "../_ref/gen/random_t.hpp"
#include "../_ref/gen/common_t.hpp"
// traceray_rq.glsl:108-147 ClosestHit / :153-185 AnyHit — the ray queries themselves run in the driver; the payload is filled as there
static void ClosestHit(Ray r) {
  const float ray[8] = {r.origin.x, r.origin.y, r.origin.z, INFINITY, r.direction.x, r.direction.y, r.direction.z, 0.0f};
  HitRec h;
  g_trace(g_scene, ray, 1u, 0, &h);
  ++g_closest;
  prd.hitT = h.hitT; prd.primitiveID = h.primitiveID; prd.instanceID = h.instanceID; prd.instanceCustomIndex = h.instanceCustomIndex;
  prd.baryCoord = vec2(h.baryU, h.baryV);
  mat4x3 o2w, w2o;
  o2w.c[0] = w2o.c[0] = vec3(1.0f, 0.0f, 0.0f); o2w.c[1] = w2o.c[1] = vec3(0.0f, 1.0f, 0.0f); o2w.c[2] = w2o.c[2] = vec3(0.0f, 0.0f, 1.0f); o2w.c[3] = w2o.c[3] = vec3(0.0f);
  if (g_xforms && h.instanceID >= 0) {
    const float* m = g_xforms + 24 * (size_t)h.instanceID;
    for (int c = 0; c < 4; ++c) { o2w.c[c] = vec3(m[3 * c], m[3 * c + 1], m[3 * c + 2]); w2o.c[c] = vec3(m[12 + 3 * c], m[12 + 3 * c + 1], m[12 + 3 * c + 2]); }
  }
  prd.objectToWorld = o2w; prd.worldToObject = w2o;
}
static bool AnyHit(Ray r, float maxDist) {
  const float ray[8] = {r.origin.x, r.origin.y, r.origin.z, maxDist, r.direction.x, r.direction.y, r.direction.z, 0.0f};
  HitRec h;
  g_trace(g_scene, ray, 1u, 1, &h);
  ++g_any;
  return h.hitT < 1.0f;     // the tap reports an occluded ray as hitT = 0, a free one as 1e28
}
#include "../_ref/gen/pbr.hpp"
#include "../_ref/gen/gltf_material_t.hpp"
#include "../_ref/gen/sun_and_sky.hpp"
#include "../_ref/gen/env_sampling_t.hpp"
#include "../_ref/gen/shade_state.hpp"
#include "../_ref/gen/reservoir.hpp"
#include "../_ref/gen/pathtrace_t.hpp"

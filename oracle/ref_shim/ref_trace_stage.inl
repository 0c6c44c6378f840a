// oracle/ref_shim/ref_trace_stage.inl — TEST INFRASTRUCTURE.  Included once per trace stage (inside its namespace) by ref_trace.cpp, after
// globals.glsl: the stage's private globals, the emulated
// rayQuery*EXT built-ins + traceray_rq.glsl, then the reference's include chain of pathtrace.glsl.
PtPayload prd;
ShadowHitPayload shadow_payload;
ivec2 imageCoords;
#include "../_ref/gen/random_t.hpp"
#include "../_ref/gen/common_t.hpp"
// ---- GL_EXT_ray_query, emulated: the part of traceray_rq.glsl that executes inside the Vulkan driver ------------------------------------
struct RqHit { float t; int prim, inst, custom; float u, v; int opaque; };
struct rayQueryEXT { float ray[8]; unsigned int flags; bool haveLow, candValid, hasCommitted, done; RqHit cand, committed; };
struct accelerationStructureEXT {};
static accelerationStructureEXT topLevelAS;
static const unsigned int gl_RayFlagsTerminateOnFirstHitEXT = 4u, gl_RayFlagsSkipClosestHitShaderEXT = 8u, gl_RayFlagsCullBackFacingTrianglesEXT = 16u;
static const unsigned int gl_RayQueryCommittedIntersectionNoneEXT = 0u, gl_RayQueryCommittedIntersectionTriangleEXT = 1u, gl_RayQueryCandidateIntersectionTriangleEXT = 0u;
static void rayQueryInitializeEXT(rayQueryEXT& rq, const accelerationStructureEXT&, unsigned int flags, unsigned int, vec3 o, float, vec3 d, float tmax) {
  rq = rayQueryEXT{};
  rq.ray[0] = o.x; rq.ray[1] = o.y; rq.ray[2] = o.z; rq.ray[3] = tmax; rq.ray[4] = d.x; rq.ray[5] = d.y; rq.ray[6] = d.z;
  rq.flags = flags;
  if (flags & gl_RayFlagsTerminateOnFirstHitEXT) ++g_any; else ++g_closest;
}
// true = a non-opaque candidate waits for the shader's decision; false = traversal complete.  Candidates arrive front to back, so a
// committed hit ends the query (closest hit: nothing nearer is left; any hit: terminate on first hit); an opaque candidate commits itself.
static bool rayQueryProceedEXT(rayQueryEXT& rq) {
  if (rq.done || rq.hasCommitted) { rq.done = true; return false; }
  float rec[7];
  const int lowInst = rq.cand.inst, lowPrim = rq.cand.prim;
  if (!g_trace(g_scene, rq.ray, rq.candValid ? 1 : 0, rq.cand.t, lowInst, lowPrim, rec)) { rq.done = true; return false; }
  RqHit c; c.t = rec[0]; c.prim = GLSL_floatBitsToInt(rec[1]); c.inst = GLSL_floatBitsToInt(rec[2]); c.custom = GLSL_floatBitsToInt(rec[3]); c.u = rec[4]; c.v = rec[5]; c.opaque = GLSL_floatBitsToInt(rec[6]);
  rq.cand = c; rq.candValid = true;
  if (c.opaque) { rq.committed = c; rq.hasCommitted = true; rq.done = true; return false; }
  return true;
}
static void rayQueryConfirmIntersectionEXT(rayQueryEXT& rq) { rq.committed = rq.cand; rq.hasCommitted = true; }
static unsigned int rayQueryGetIntersectionTypeEXT(const rayQueryEXT& rq, bool committed) { return committed ? (rq.hasCommitted ? gl_RayQueryCommittedIntersectionTriangleEXT : gl_RayQueryCommittedIntersectionNoneEXT) : gl_RayQueryCandidateIntersectionTriangleEXT; }
static const RqHit& rqSel(const rayQueryEXT& rq, bool committed) { return committed ? rq.committed : rq.cand; }
static float rayQueryGetIntersectionTEXT(const rayQueryEXT& rq, bool c) { return rqSel(rq, c).t; }
static int rayQueryGetIntersectionPrimitiveIndexEXT(const rayQueryEXT& rq, bool c) { return rqSel(rq, c).prim; }
static int rayQueryGetIntersectionInstanceIdEXT(const rayQueryEXT& rq, bool c) { return rqSel(rq, c).inst; }
static int rayQueryGetIntersectionInstanceCustomIndexEXT(const rayQueryEXT& rq, bool c) { return rqSel(rq, c).custom; }
static vec2 rayQueryGetIntersectionBarycentricsEXT(const rayQueryEXT& rq, bool c) { return vec2(rqSel(rq, c).u, rqSel(rq, c).v); }
static mat4x3 rqXform(const rayQueryEXT& rq, bool c, int which) {
  mat4x3 m; m.c[0] = vec3(1.0f, 0.0f, 0.0f); m.c[1] = vec3(0.0f, 1.0f, 0.0f); m.c[2] = vec3(0.0f, 0.0f, 1.0f); m.c[3] = vec3(0.0f);
  const int inst = rqSel(rq, c).inst;
  if (g_xforms && inst >= 0) { const float* p = g_xforms + 24 * (size_t)inst + 12 * which; for (int k = 0; k < 4; ++k) m.c[k] = vec3(p[3 * k], p[3 * k + 1], p[3 * k + 2]); }
  return m;
}
static mat4x3 rayQueryGetIntersectionObjectToWorldEXT(const rayQueryEXT& rq, bool c) { return rqXform(rq, c, 0); }
static mat4x3 rayQueryGetIntersectionWorldToObjectEXT(const rayQueryEXT& rq, bool c) { return rqXform(rq, c, 1); }
#include "../_ref/gen/traceray_rq.hpp"     // HitTest, ClosestHit, AnyHit: the reference's text
#include "../_ref/gen/pbr.hpp"
#include "../_ref/gen/gltf_material_t.hpp"
#include "../_ref/gen/sun_and_sky.hpp"
#include "../_ref/gen/env_sampling_t.hpp"
#include "../_ref/gen/shade_state.hpp"
#include "../_ref/gen/reservoir.hpp"
#include "../_ref/gen/pathtrace_t.hpp"

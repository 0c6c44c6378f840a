/* oracle/ref_shim/glsl/nvmath/nvmath.h — TEST INFRASTRUCTURE.  For the translation unit that compiles the reference's GLSL include
 * files as C++ (ref_glsl.cpp): shaders/host_device.h asks nvmath for its vector types; here they are the oracle's GLSL-like types, so
 * the structs of host_device.h (LightSample, DirectReservoir, SunAndSky ...) hold the same vec3 the transliterated GLSL computes with. */
#pragma once
#include "../../../glsl_types.h"
namespace nvmath {
using vec2i = orc::ivec2; using vec2f = orc::vec2; using vec3f = orc::vec3; using vec4f = orc::vec4; using vec4ui = orc::uvec4;
struct vec2ui { unsigned int x, y; };
struct mat4f {
  float m[16];
  mat4f() = default;
  explicit mat4f(const orc::mat4x3& a) {   // GLSL mat4(mat4x3): the missing row is (0, 0, 0, 1)
    for (int c = 0; c < 4; ++c) { m[4 * c] = a.c[c].x; m[4 * c + 1] = a.c[c].y; m[4 * c + 2] = a.c[c].z; m[4 * c + 3] = (c == 3) ? 1.0f : 0.0f; }
  }
};
}

/* oracle/ref_shim/glsl/nvmath/nvmath.h — TEST INFRASTRUCTURE.  For the translation unit that compiles the reference's GLSL include
 * files as C++ (ref_glsl.cpp): shaders/host_device.h asks nvmath for its vector types; here they are the oracle's GLSL-like types, so
 * the structs of host_device.h (LightSample, DirectReservoir, SunAndSky ...) hold the same vec3 the transliterated GLSL computes with. */
#pragma once
#include "../../../glsl_types.h"
namespace nvmath {
using vec2i = orc::ivec2; using vec2f = orc::vec2; using vec3f = orc::vec3; using vec4f = orc::vec4; using vec4ui = orc::uvec4;
struct vec2ui { unsigned int x, y; };
struct mat4f { float m[16]; };
}

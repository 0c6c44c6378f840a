// TEST INFRASTRUCTURE: stand-in for the glslang-generated SPIR-V array of direct_reuse.comp (one word: the tag the recording device reports)
#pragma once
#include <cstdint>
static const uint32_t direct_reuse_comp[] = {3};

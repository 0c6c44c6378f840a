// TEST INFRASTRUCTURE: stand-in for the glslang-generated SPIR-V array of direct_stage.comp (one word: the tag the recording device reports)
#pragma once
#include <cstdint>
static const uint32_t direct_stage_comp[] = {1};

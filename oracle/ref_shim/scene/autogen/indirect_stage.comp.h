// TEST INFRASTRUCTURE: stand-in for the glslang-generated SPIR-V array of indirect_stage.comp (one word: the tag the recording device reports)
#pragma once
#include <cstdint>
static const uint32_t indirect_stage_comp[] = {4};

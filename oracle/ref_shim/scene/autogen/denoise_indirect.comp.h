// TEST INFRASTRUCTURE: stand-in for the glslang-generated SPIR-V array of denoise_indirect.comp (one word: the tag the recording device reports)
#pragma once
#include <cstdint>
static const uint32_t denoise_indirect_comp[] = {6};

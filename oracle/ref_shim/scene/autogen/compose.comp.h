// TEST INFRASTRUCTURE: stand-in for the glslang-generated SPIR-V array of compose.comp (one word: the tag the recording device reports)
#pragma once
#include <cstdint>
static const uint32_t compose_comp[] = {7};

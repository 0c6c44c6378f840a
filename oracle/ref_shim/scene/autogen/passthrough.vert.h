// TEST INFRASTRUCTURE: stand-in for the glslang-generated SPIR-V array of passthrough.vert (one word: the tag the recording device reports)
#pragma once
#include <cstdint>
static const uint32_t passthrough_vert[] = {8};

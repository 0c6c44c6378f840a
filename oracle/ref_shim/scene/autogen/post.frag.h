// TEST INFRASTRUCTURE: stand-in for the glslang-generated SPIR-V array of post.frag (one word: the tag the recording device reports)
#pragma once
#include <cstdint>
static const uint32_t post_frag[] = {9};

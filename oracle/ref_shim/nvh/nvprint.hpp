#pragma once
#include <cstdio>
#define LOGI(...) std::printf(__VA_ARGS__)

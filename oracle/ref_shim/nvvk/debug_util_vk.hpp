#pragma once
#include "../vk_shim.h"
namespace nvvk { struct DebugUtil { void setup(VkDevice) {} }; }
#define NAME_VK(x) (void)(x)

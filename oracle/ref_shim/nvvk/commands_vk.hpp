#pragma once
#include "../vk_shim.h"
namespace nvvk { struct ScopeCommandBuffer { ScopeCommandBuffer(VkDevice, uint32_t, VkQueue) {} operator VkCommandBuffer() const { return nullptr; } }; }

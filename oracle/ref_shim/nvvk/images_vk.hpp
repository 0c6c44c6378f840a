#pragma once
#include "../vk_shim.h"
namespace nvvk {
inline VkImageCreateInfo makeImage2DCreateInfo(VkExtent2D e, VkFormat f) { return VkImageCreateInfo{e, f}; }
inline VkImageViewCreateInfo makeImageViewCreateInfo(VkImage i, const VkImageCreateInfo&) { return VkImageViewCreateInfo{i}; }
}

/* oracle/ref_shim/vk_shim.h — TEST INFRASTRUCTURE.  The handful of Vulkan names src/hdr_sampling.{hpp,cpp} of the reference
 * mention, as inert stand-ins: enough for that file to COMPILE where it lies, so that its two pure functions
 * (HdrSampling::buildAliasmap / createEnvironmentAccel) can be called; nothing here does anything. */
#pragma once
#include <cstdint>
#include <cstddef>
typedef void* VkDevice; typedef void* VkPhysicalDevice; typedef void* VkQueue; typedef void* VkCommandBuffer;
typedef void* VkImage; typedef void* VkBuffer; typedef uint64_t VkDeviceSize; typedef uint32_t VkFlags;
#define VK_NULL_HANDLE nullptr
struct VkExtent2D { uint32_t width, height; };
enum VkStructureType { VK_STRUCTURE_TYPE_SAMPLER_CREATE_INFO = 31 };
enum VkFilter { VK_FILTER_NEAREST = 0, VK_FILTER_LINEAR = 1 };
enum VkSamplerMipmapMode { VK_SAMPLER_MIPMAP_MODE_NEAREST = 0, VK_SAMPLER_MIPMAP_MODE_LINEAR = 1 };
enum VkSamplerAddressMode { VK_SAMPLER_ADDRESS_MODE_REPEAT = 0, VK_SAMPLER_ADDRESS_MODE_CLAMP_TO_EDGE = 2 };
enum VkFormat { VK_FORMAT_R32G32B32A32_SFLOAT = 109 };
enum { VK_BUFFER_USAGE_STORAGE_BUFFER_BIT = 0x20 };
struct VkSamplerCreateInfo { VkStructureType sType; const void* pNext; VkFilter magFilter, minFilter; VkSamplerMipmapMode mipmapMode;
                             VkSamplerAddressMode addressModeU, addressModeV, addressModeW; float maxLod; };
struct VkImageCreateInfo { VkExtent2D extent; VkFormat format; };
struct VkImageViewCreateInfo { VkImage image; };
inline void vkGetDeviceQueue(VkDevice, uint32_t, uint32_t, VkQueue* q) { *q = nullptr; }

#pragma once
#include <vector>
#include "../vk_shim.h"
namespace nvvk {
struct Image { VkImage image = nullptr; };
struct Texture { VkImage image = nullptr; };
struct Buffer { VkBuffer buffer = nullptr; };
class ResourceAllocator {
public:
  void destroy(Texture&) {}
  void destroy(Buffer&) {}
  Image createImage(VkCommandBuffer, VkDeviceSize, const void*, const VkImageCreateInfo&) { return Image(); }
  Texture createTexture(const Image&, const VkImageViewCreateInfo&, const VkSamplerCreateInfo&) { return Texture(); }
  template <class T> Buffer createBuffer(VkCommandBuffer, const std::vector<T>&, VkFlags) { return Buffer(); }
  void finalizeAndReleaseStaging() {}
};
}

// sah_host.h — host-side binned-SAH builder of the BINARY tree the GPU pipeline of accel.cu refits and collapses into 4-wide nodes.
// The reference builds its acceleration structures with VK_BUILD_ACCELERATION_STRUCTURE_PREFER_FAST_TRACE_BIT_KHR (accelstruct.cpp:125-126,
// 161): a one-off, quality-first build.  This is the counterpart of that flag; the Morton/LBVH build on the GPU is the PREFER_FAST_BUILD one.
// Which tree is walked never changes a result (closest hit = total order on (t, instance, primitive)) — only how many nodes
// a ray visits.
#pragma once
#include <cstdint>
#include <vector>

namespace eid {

// Same shape as what k_hierarchy (Karras) produces: n - 1 inner nodes over the primitives in `order`, root = inner node 0,
// child >= 0: inner node, child < 0: leaf ~position (position in `order`); every inner node covers positions [rangeFirst, rangeLast].
struct BinaryTreeHost {
  std::vector<uint32_t> order;        // position -> source primitive
  std::vector<int> left, right, parentInner, rangeFirst, rangeLast;   // per inner node
  std::vector<int> parentLeaf;        // per position
};

// lo / hi: n x 3 floats (the padded boxes).  threads <= 0: hardware concurrency.  Deterministic whatever the thread count.
void buildSahTree(uint32_t n, const float* lo, const float* hi, BinaryTreeHost& T, int threads = 0);

}  // namespace eid

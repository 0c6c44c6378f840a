// sah_host.cpp — binned surface-area-heuristic build of the binary tree over primitive boxes (see sah_host.h).
// Top-down: every node bins the centroids of its primitives into 32 bins on each axis, evaluates the 3 x 31 candidate planes with
// cost = area(L) * |L| + area(R) * |R| and partitions at the cheapest; down to single primitives (the 4-wide collapse on the GPU turns
// subtrees of <= LEAF_MAX primitives into leaves).  An inner node over positions [f, l] that splits between m - 1 and m gets the index
// m - 1 (every gap between two neighbouring positions belongs to exactly one inner node), with the root swapped to index 0 — so node
// indices need no allocation and subtrees can be built by independent threads.
#include "sah_host.h"
#include "common.h"
#include "eidola.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <exception>
#include <mutex>
#include <thread>

namespace eid {
namespace {

constexpr int NB = 32;          // 16 -> 32 bins: 3-5 % fewer node visits for secondary rays on the C3 scene, same build time

struct Bx {
  float lo[3], hi[3];
  void reset() { for (int k = 0; k < 3; ++k) { lo[k] = 3.0e38f; hi[k] = -3.0e38f; } }
  void grow(const float* l, const float* h) { for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], l[k]); hi[k] = std::max(hi[k], h[k]); } }
  void grow(const Bx& o) { grow(o.lo, o.hi); }
  float area() const { const float ex = hi[0] - lo[0], ey = hi[1] - lo[1], ez = hi[2] - lo[2]; return ex * ey + ey * ez + ez * ex; }
};

struct Item { int first, count, parent, side; };   // side 0 = left child of parent, 1 = right; parent -1 = root

struct Builder {
  uint32_t n;
  const float* lo; const float* hi;
  BinaryTreeHost& T;
  int threads = 1;
  int rootNatural = 0;

  int nodeId(int natural) const { return natural == rootNatural ? 0 : (natural == 0 ? rootNatural : natural); }
  float centroid(uint32_t p, int k) const { return 0.5f * (lo[3 * (size_t)p + k] + hi[3 * (size_t)p + k]); }

  // position of the first primitive of the right part: first < mid < first + count
  int split(int first, int count) {
    if (count == 2) return first + 1;
    uint32_t* ord = T.order.data();
    // large nodes (the top of the tree, built before there are enough independent subtrees) are binned by all threads in chunks
    const int chunks = (count >= (1 << 17) && threads > 1) ? threads : 1;
    const int per = (count + chunks - 1) / chunks;
    auto forChunks = [&](auto&& fn) {
      if (chunks == 1) { fn(0, first, first + count); return; }
      // (fn only reads the boxes and writes its own chunk's slot: it cannot throw; a chunk whose thread could not be created runs here)
      std::vector<std::thread> pool;
      int started = 1;
      try {
        for (; started < chunks; ++started) { const int c = started; pool.emplace_back([&, c]() { fn(c, std::min(first + count, first + c * per), std::min(first + count, first + (c + 1) * per)); }); }
      } catch (...) {
      }
      fn(0, first, std::min(first + count, first + per));
      for (int c = started; c < chunks; ++c) fn(c, std::min(first + count, first + c * per), std::min(first + count, first + (c + 1) * per));
      for (std::thread& t : pool) t.join();
    };
    std::vector<float> cbLo(3 * (size_t)chunks, 3.0e38f), cbHi(3 * (size_t)chunks, -3.0e38f);
    forChunks([&](int c, int b, int e) {
      float l[3] = {3.0e38f, 3.0e38f, 3.0e38f}, h[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
      for (int i = b; i < e; ++i)
        for (int k = 0; k < 3; ++k) { const float v = centroid(ord[i], k); l[k] = std::min(l[k], v); h[k] = std::max(h[k], v); }
      for (int k = 0; k < 3; ++k) { cbLo[3 * (size_t)c + k] = l[k]; cbHi[3 * (size_t)c + k] = h[k]; }
    });
    float cl[3] = {3.0e38f, 3.0e38f, 3.0e38f}, ch[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    for (int c = 0; c < chunks; ++c) for (int k = 0; k < 3; ++k) { cl[k] = std::min(cl[k], cbLo[3 * (size_t)c + k]); ch[k] = std::max(ch[k], cbHi[3 * (size_t)c + k]); }
    float scale[3]; bool use[3];
    for (int k = 0; k < 3; ++k) { const float e = ch[k] - cl[k]; use[k] = e > 0.0f && std::isfinite(e); scale[k] = use[k] ? (float)NB / e : 0.0f; }
    struct Bins { Bx box[3][NB]; int cnt[3][NB]; };
    std::vector<Bins> part((size_t)chunks);
    forChunks([&](int c, int b, int e) {
      Bins& P = part[(size_t)c];
      for (int k = 0; k < 3; ++k) for (int q = 0; q < NB; ++q) { P.box[k][q].reset(); P.cnt[k][q] = 0; }
      for (int i = b; i < e; ++i) {
        const uint32_t p = ord[i];
        const float* l = lo + 3 * (size_t)p; const float* h = hi + 3 * (size_t)p;
        for (int k = 0; k < 3; ++k) {
          if (!use[k]) continue;
          const int q = std::min(NB - 1, std::max(0, (int)((centroid(p, k) - cl[k]) * scale[k])));
          P.box[k][q].grow(l, h); P.cnt[k][q]++;
        }
      }
    });
    Bins& M = part[0];                 // min / max and integer sums: the merge does not depend on the chunking
    for (int c = 1; c < chunks; ++c) for (int k = 0; k < 3; ++k) for (int q = 0; q < NB; ++q) { M.box[k][q].grow(part[(size_t)c].box[k][q]); M.cnt[k][q] += part[(size_t)c].cnt[k][q]; }
    auto& bins = M.box; auto& cnt = M.cnt;
    float bestCost = 3.0e38f; int bestAxis = -1, bestBin = -1;
    for (int k = 0; k < 3; ++k) {
      if (!use[k]) continue;
      float rArea[NB]; int rCnt[NB];
      Bx acc; acc.reset(); int c = 0;
      for (int b = NB - 1; b > 0; --b) { acc.grow(bins[k][b]); c += cnt[k][b]; rArea[b] = c ? acc.area() : 0.0f; rCnt[b] = c; }
      acc.reset(); c = 0;
      for (int b = 0; b < NB - 1; ++b) {
        acc.grow(bins[k][b]); c += cnt[k][b];
        if (c == 0 || rCnt[b + 1] == 0) continue;
        const float cost = acc.area() * (float)c + rArea[b + 1] * (float)rCnt[b + 1];
        if (cost < bestCost) { bestCost = cost; bestAxis = k; bestBin = b; }
      }
    }
    int mid = -1;
    if (bestAxis >= 0) {
      const int k = bestAxis; const float c0 = cl[k], sc = scale[k];
      uint32_t* it = std::partition(ord + first, ord + first + count, [&](uint32_t p) {
        return std::min(NB - 1, std::max(0, (int)((centroid(p, k) - c0) * sc))) <= bestBin; });
      mid = (int)(it - ord);
    }
    if (mid <= first || mid >= first + count) {
      // all centroids coincide (or non-finite boxes): median split in the current order
      mid = first + count / 2;
    }
    return mid;
  }

  // builds the inner node of [first, first + count), count >= 2, and returns its index; children of more than one primitive are pushed onto `defer`
  int node(const Item& it, std::vector<Item>* defer) {
    const int mid = split(it.first, it.count);
    if (it.parent < 0) rootNatural = mid - 1;
    const int id = nodeId(mid - 1);
    T.parentInner[id] = it.parent; T.rangeFirst[id] = it.first; T.rangeLast[id] = it.first + it.count - 1;
    if (it.parent >= 0) { if (it.side) T.right[it.parent] = id; else T.left[it.parent] = id; }
    const Item ch[2] = {{it.first, mid - it.first, id, 0}, {mid, it.first + it.count - mid, id, 1}};
    for (int s = 0; s < 2; ++s) {
      if (ch[s].count == 1) {
        if (s) T.right[id] = ~ch[s].first; else T.left[id] = ~ch[s].first;
        T.parentLeaf[ch[s].first] = id;
      } else if (defer) defer->push_back(ch[s]);
      else subtree(ch[s]);
    }
    return id;
  }
  void subtree(const Item& root) {       // explicit stack: the depth of a degenerate input must not overflow the thread's stack
    std::vector<Item> st; st.push_back(root);
    while (!st.empty()) { const Item it = st.back(); st.pop_back(); node(it, &st); }
  }
};

}  // namespace

void buildSahTree(uint32_t n, const float* lo, const float* hi, BinaryTreeHost& T, int threads) {
  T.order.resize(n);
  for (uint32_t i = 0; i < n; ++i) T.order[i] = i;
  const size_t ni = n > 1 ? n - 1 : 0;
  T.left.assign(ni, 0); T.right.assign(ni, 0); T.parentInner.assign(ni, -1); T.rangeFirst.assign(ni, 0); T.rangeLast.assign(ni, 0);
  T.parentLeaf.assign(n, -1);
  if (n < 2) return;
  if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
  threads = std::max(1, std::min(threads, 64));
  Builder B{n, lo, hi, T, threads};
  // breadth-first on this thread until there are enough independent subtrees
  std::vector<Item> queue; queue.push_back(Item{0, (int)n, -1, 0});
  const int minParallel = 1 << 14;
  size_t head = 0;
  while (head < queue.size() && (int)(queue.size() - head) < 4 * threads) {
    // largest pending range first keeps the task sizes even
    size_t big = head;
    for (size_t i = head + 1; i < queue.size(); ++i) if (queue[i].count > queue[big].count) big = i;
    if (queue[big].count < minParallel || threads == 1) break;
    std::swap(queue[head], queue[big]);
    const Item it = queue[head++];
    B.node(it, &queue);
  }
  std::vector<Item> tasks(queue.begin() + head, queue.end());
  std::sort(tasks.begin(), tasks.end(), [](const Item& a, const Item& b) { return a.count > b.count; });
  if (threads == 1 || tasks.size() < 2) { for (const Item& it : tasks) B.subtree(it); return; }
  std::atomic<size_t> next{0};
  std::atomic<bool> failed{false};
  std::exception_ptr error;                        // an exception must not leave a worker thread (std::terminate): the first one is rethrown here
  std::mutex errorLock;
  auto work = [&]() {
    try {
      for (;;) { const size_t i = next.fetch_add(1); if (i >= tasks.size() || failed.load()) break; B.subtree(tasks[i]); }
    } catch (...) {
      std::lock_guard<std::mutex> g(errorLock);
      if (!error) error = std::current_exception();
      failed.store(true);
    }
  };
  std::vector<std::thread> pool;
  try {
    for (int t = 1; t < threads; ++t) pool.emplace_back(work);
  } catch (...) {                                  // no more threads to be had: the ones that started, and this one, do the work
  }
  work();
  for (std::thread& t : pool) t.join();
  if (error) std::rethrow_exception(error);
}

}  // namespace eid

extern "C" int eid_accel_sah_tap(const float* lo, const float* hi, uint32_t n, int threads, uint32_t* order, int32_t* left, int32_t* right,
                                 int32_t* parentInner, int32_t* parentLeaf, int32_t* rangeFirst, int32_t* rangeLast) {
  EID_TRY
  if (!lo || !hi || !order || !parentLeaf || (n > 1 && (!left || !right || !parentInner || !rangeFirst || !rangeLast))) eid::raise(EID_ERR_INVALID, "eid_accel_sah_tap: null argument");
  if (n >= (1u << 28)) eid::raise(EID_ERR_UNSUPPORTED, "eid_accel_sah_tap: more than 2^28 boxes");
  eid::BinaryTreeHost T;
  eid::buildSahTree(n, lo, hi, T, threads);
  std::copy(T.order.begin(), T.order.end(), order);
  std::copy(T.parentLeaf.begin(), T.parentLeaf.end(), parentLeaf);
  std::copy(T.left.begin(), T.left.end(), left); std::copy(T.right.begin(), T.right.end(), right);
  std::copy(T.parentInner.begin(), T.parentInner.end(), parentInner);
  std::copy(T.rangeFirst.begin(), T.rangeFirst.end(), rangeFirst); std::copy(T.rangeLast.begin(), T.rangeLast.end(), rangeLast);
  return EID_OK;
  EID_CATCH
}

// accel.cu — device upload of the scene tables, the CUDA BVH builder that replaces
// AccelStructure::create (reference src/accelstruct.cpp:55-162: per-prim-mesh BLAS + one TLAS instance per
// node, built by the Vulkan driver) and the C-ABI entry points for Scene and AccelStructure.
//
// Builder (all on the GPU, one stream):
//   1. k_emit_triangles   every TLAS instance's triangles -> world space (contract arithmetic), 48 B
//                         Moller-Trumbore records, padded AABBs, scene bounds (atomics)
//   2. k_morton           63-bit Morton code of each AABB centre
//   3. cub radix sort     (key, triangle id)
//   4. k_reorder          triangles + AABBs into Morton order
//   5. k_hierarchy        Karras 2012 binary radix tree, one thread per inner node
//   6. k_refit            bottom-up AABB / leaf-count / height, atomic arrival counters
//   7. k_collapse4        level-by-level collapse into 128 B 4-wide nodes (surface-area greedy); subtrees with <= 4
//                         triangles become leaves.  (EID_BVH_WIDTH=2 keeps the binary tree: k_pack_nodes, 64 B two-box nodes)
#include <cub/device/device_radix_sort.cuh>
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "accel.h"
#include "sah_host.h"
#include "common.h"
#include "trace.cuh"

namespace eid {

#ifndef LEAF_MAX
#define LEAF_MAX 2u   // triangles per leaf (<= 7: the count lives in 3 bits of the reference); measured 1/2/3/4/6 on C3:
                      // 2 is fastest (13.5 nodes + 3.3 triangle tests per ray; 4 gave 13.0 + 5.6 and a 5% slower frame)
#endif

// ------------------------------------------------------------------------------------------------
// scene upload
// ------------------------------------------------------------------------------------------------
template <class T>
static T* uploadVec(const std::vector<T>& v) {
  T* d = nullptr;
  size_t bytes = std::max<size_t>(1, v.size()) * sizeof(T);
  CUDA_CHECK(cudaMalloc(&d, bytes));
  if (!v.empty()) CUDA_CHECK(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return d;
}

void SceneDevice::release() {
  cudaFree(vertices); cudaFree(indices); cudaFree(geoInfo); cudaFree(materials); cudaFree(puncLights);
  cudaFree(trigLights); cudaFree(instances); cudaFree(instFirstTri); cudaFree(textures); cudaFree(texels);
  textures = nullptr; texels = nullptr;
  vertices = nullptr; indices = nullptr; geoInfo = nullptr; materials = nullptr; puncLights = nullptr;
  trigLights = nullptr; instances = nullptr; instFirstTri = nullptr;
}

void SceneDevice::upload(const SceneHost& h) {
  release();
  CUDA_CHECK(cudaSetDevice(device));
  vertices = uploadVec(h.vertices);
  indices = uploadVec(h.indices);
  // InstanceData with real device addresses, like the reference's buffer_reference pointers (scene.cpp:179-195)
  std::vector<InstanceData> inst(h.gltf.primMeshes.size());
  for (size_t p = 0; p < inst.size(); ++p) {
    inst[p].vertexAddress = (uint64_t)(uintptr_t)(vertices + h.vtxBase[p]);
    inst[p].indexAddress = (uint64_t)(uintptr_t)(indices + h.idxBase[p]);
    inst[p].materialIndex = h.gltf.primMeshes[p].materialIndex;
  }
  geoInfo = uploadVec(inst);
  materials = uploadVec(h.materials);
  puncLights = uploadVec(h.puncLights);
  trigLights = uploadVec(h.trigLights);
  instances = uploadVec(h.instances);
  texels = uploadVec(h.texels);
  std::vector<TextureDev> td(h.textures.size());
  for (size_t i = 0; i < td.size(); ++i)
    td[i] = TextureDev{texels + h.textures[i].texelOffset, (int32_t)h.textures[i].width, (int32_t)h.textures[i].height, h.textures[i].linear,
                       h.textures[i].wrapS, h.textures[i].wrapT, 0};
  textures = uploadVec(td);
  std::vector<uint32_t> first(h.instances.size() + 1, 0);
  for (size_t i = 0; i < h.instances.size(); ++i) first[i + 1] = first[i] + h.instances[i].triangleCount;
  instFirstTri = uploadVec(first);
}

DeviceSceneView SceneDevice::view(const SceneHost& h) const {
  DeviceSceneView v;
  v.textures = textures;
  v.geoInfo = geoInfo; v.materials = materials; v.puncLights = puncLights; v.trigLights = trigLights;
  v.instances = instances; v.lightBufInfo = h.lightInfo;
  return v;
}

// ------------------------------------------------------------------------------------------------
// builder kernels
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int floatFlip(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float floatUnflip(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

struct BuildBounds { int lo[3]; int hi[3]; };

__global__ void k_init_bounds(BuildBounds* b) {
  for (int k = 0; k < 3; ++k) { b->lo[k] = floatFlip(3.0e38f); b->hi[k] = floatFlip(-3.0e38f); }
}

__global__ void k_emit_triangles(DeviceSceneView sc, const uint32_t* __restrict__ instFirst, uint32_t nInst, uint32_t nTri,
                                 float4* __restrict__ triOut, float* __restrict__ boxLo, float* __restrict__ boxHi, BuildBounds* bounds) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
  if (t < nTri) {
    uint32_t a = 0, b = nInst;                       // instance owning triangle t (upper bound search)
    while (b - a > 1) { uint32_t m = (a + b) >> 1; if (instFirst[m] <= t) a = m; else b = m; }
    const InstanceXform& X = sc.instances[a];
    const uint32_t prim = t - instFirst[a];
    const InstanceData gi = sc.geoInfo[X.primMesh];
    const uint32_t* idx = (const uint32_t*)(uintptr_t)gi.indexAddress;
    const VertexAttributes* vtx = (const VertexAttributes*)(uintptr_t)gi.vertexAddress;
    f3 p[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) p[k] = xfPoint(X.objectToWorld, ld3(vtx[idx[3 * prim + k]].position));
    f3 e1 = p[1] - p[0], e2 = p[2] - p[0];
    triOut[3 * (size_t)t + 0] = make_float4(p[0].x, p[0].y, p[0].z, e1.x);
    triOut[3 * (size_t)t + 1] = make_float4(e1.y, e1.z, e2.x, e2.y);
    triOut[3 * (size_t)t + 2] = make_float4(e2.z, __int_as_float((int)prim), __int_as_float((int)a),
                                            __uint_as_float(X.flags & (INST_CULL_DISABLE | INST_MIRROR | INST_FORCE_OPAQUE)));
    lo[0] = fminf(p[0].x, fminf(p[1].x, p[2].x)); hi[0] = fmaxf(p[0].x, fmaxf(p[1].x, p[2].x));
    lo[1] = fminf(p[0].y, fminf(p[1].y, p[2].y)); hi[1] = fmaxf(p[0].y, fmaxf(p[1].y, p[2].y));
    lo[2] = fminf(p[0].z, fminf(p[1].z, p[2].z)); hi[2] = fmaxf(p[0].z, fmaxf(p[1].z, p[2].z));
#pragma unroll
    for (int k = 0; k < 3; ++k) { boxLo[3 * (size_t)t + k] = lo[k]; boxHi[3 * (size_t)t + k] = hi[k]; }
  }
  // warp-reduce the bounds, one atomic per warp
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float l = lo[k], h = hi[k];
    for (int o = 16; o > 0; o >>= 1) { l = fminf(l, __shfl_xor_sync(0xffffffffu, l, o)); h = fmaxf(h, __shfl_xor_sync(0xffffffffu, h, o)); }
    if ((threadIdx.x & 31) == 0) { atomicMin(&bounds->lo[k], floatFlip(l)); atomicMax(&bounds->hi[k], floatFlip(h)); }
  }
}

__device__ __forceinline__ unsigned long long expand21(unsigned long long v) {   // spread 21 bits to every 3rd bit
  v &= 0x1fffffull;
  v = (v | v << 32) & 0x1f00000000ffffull;
  v = (v | v << 16) & 0x1f0000ff0000ffull;
  v = (v | v << 8) & 0x100f00f00f00f00full;
  v = (v | v << 4) & 0x10c30c30c30c30c3ull;
  v = (v | v << 2) & 0x1249249249249249ull;
  return v;
}

// pads every triangle box (conservative w.r.t. the fp32 Moller-Trumbore test) and computes its Morton key
__global__ void k_morton(uint32_t nTri, float* __restrict__ boxLo, float* __restrict__ boxHi, const BuildBounds* bounds,
                         unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nTri) return;
  float sl[3], sh[3];
  for (int k = 0; k < 3; ++k) { sl[k] = floatUnflip(bounds->lo[k]); sh[k] = floatUnflip(bounds->hi[k]); }
  float ex = sh[0] - sl[0], ey = sh[1] - sl[1], ez = sh[2] - sl[2];
  float diag = sqrtf(ex * ex + ey * ey + ez * ez);
  float amax = fmaxf(fmaxf(fmaxf(fabsf(sl[0]), fabsf(sh[0])), fmaxf(fabsf(sl[1]), fabsf(sh[1]))), fmaxf(fabsf(sl[2]), fabsf(sh[2])));
  const float pad = 1e-5f * diag + 4e-7f * amax + 1e-7f;
  unsigned long long key = 0;
  for (int k = 0; k < 3; ++k) {
    float l = boxLo[3 * (size_t)t + k] - pad, h = boxHi[3 * (size_t)t + k] + pad;
    boxLo[3 * (size_t)t + k] = l; boxHi[3 * (size_t)t + k] = h;
    float e = sh[k] - sl[k];
    float c = (e > 0.f) ? (0.5f * (l + h) - sl[k]) / e : 0.5f;
    c = fminf(fmaxf(c, 0.f), 1.f);
    unsigned long long q = (unsigned long long)fminf(c * 2097152.0f, 2097151.0f);
    key |= expand21(q) << (2 - k);
  }
  keys[t] = key; vals[t] = t;
}

__global__ void k_reorder(uint32_t nTri, const uint32_t* __restrict__ order, const float4* __restrict__ triIn, const float* __restrict__ loIn,
                          const float* __restrict__ hiIn, float4* __restrict__ triOut, float* __restrict__ loOut, float* __restrict__ hiOut) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nTri) return;
  uint32_t s = order[t];
  for (int k = 0; k < 3; ++k) { triOut[3 * (size_t)t + k] = triIn[3 * (size_t)s + k]; loOut[3 * (size_t)t + k] = loIn[3 * (size_t)s + k]; hiOut[3 * (size_t)t + k] = hiIn[3 * (size_t)s + k]; }
}

__device__ __forceinline__ int prefixLen(const unsigned long long* __restrict__ keys, int n, int i, int j) {
  if (j < 0 || j >= n) return -1;
  unsigned long long a = keys[i], b = keys[j];
  if (a == b) return 64 + __clz(i ^ j);
  return __clzll((long long)(a ^ b));
}

// Karras, "Maximizing Parallelism in the Construction of BVHs, Octrees, and k-d Trees" (2012).
// child encoding here: >= 0 inner node, < 0 leaf ~id
__global__ void k_hierarchy(int n, const unsigned long long* __restrict__ keys, int* __restrict__ left, int* __restrict__ right,
                            int* __restrict__ parentInner, int* __restrict__ parentLeaf, int* __restrict__ rangeFirst, int* __restrict__ rangeLast) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n - 1) return;
  int d = (prefixLen(keys, n, i, i + 1) - prefixLen(keys, n, i, i - 1)) >= 0 ? 1 : -1;
  int dmin = prefixLen(keys, n, i, i - d);
  int lmax = 2;
  while (prefixLen(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
  int l = 0;
  for (int t = lmax >> 1; t >= 1; t >>= 1)
    if (prefixLen(keys, n, i, i + (l + t) * d) > dmin) l += t;
  int j = i + l * d;
  int dnode = prefixLen(keys, n, i, j);
  int s = 0;
  for (int t = (l + 1) >> 1;; t = (t + 1) >> 1) {
    if (prefixLen(keys, n, i, i + (s + t) * d) > dnode) s += t;
    if (t == 1) break;
  }
  int gamma = i + s * d + min(d, 0);
  int lo = min(i, j), hi = max(i, j);
  rangeFirst[i] = lo; rangeLast[i] = hi;
  if (lo == gamma) { left[i] = ~gamma; parentLeaf[gamma] = i; } else { left[i] = gamma; parentInner[gamma] = i; }
  if (hi == gamma + 1) { right[i] = ~(gamma + 1); parentLeaf[gamma + 1] = i; } else { right[i] = gamma + 1; parentInner[gamma + 1] = i; }
  if (i == 0) parentInner[0] = -1;
}

__global__ void k_refit(int n, const int* __restrict__ left, const int* __restrict__ right, const int* __restrict__ parentInner,
                        const int* __restrict__ parentLeaf, const float* __restrict__ leafLo, const float* __restrict__ leafHi,
                        float* nodeLo, float* nodeHi, int* height, unsigned int* __restrict__ arrived) {
  int leaf = blockIdx.x * blockDim.x + threadIdx.x;
  if (leaf >= n) return;
  int node = parentLeaf[leaf];
  while (node >= 0) {
    __threadfence();
    if (atomicAdd(&arrived[node], 1u) == 0u) return;   // first child to arrive stops; the second one owns both boxes
    __threadfence();
    float lo[3], hi[3];
    int h = 0;
    const int cs[2] = {left[node], right[node]};
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const volatile float* cl; const volatile float* ch; int chh;
      if (cs[c] < 0) { cl = leafLo + 3 * (size_t)(~cs[c]); ch = leafHi + 3 * (size_t)(~cs[c]); chh = 0; }
      else { cl = nodeLo + 3 * (size_t)cs[c]; ch = nodeHi + 3 * (size_t)cs[c]; chh = ((volatile int*)height)[cs[c]]; }
      for (int k = 0; k < 3; ++k) {
        lo[k] = c ? fminf(lo[k], cl[k]) : cl[k];
        hi[k] = c ? fmaxf(hi[k], ch[k]) : ch[k];
      }
      h = max(h, chh);
    }
    for (int k = 0; k < 3; ++k) { nodeLo[3 * (size_t)node + k] = lo[k]; nodeHi[3 * (size_t)node + k] = hi[k]; }
    height[node] = h + 1;
    node = parentInner[node];
  }
}

__global__ void k_pack_nodes(int n, const int* __restrict__ left, const int* __restrict__ right, const int* __restrict__ rangeFirst,
                             const int* __restrict__ rangeLast, const float* __restrict__ leafLo, const float* __restrict__ leafHi,
                             const float* __restrict__ nodeLo, const float* __restrict__ nodeHi, float4* __restrict__ out, unsigned int* __restrict__ liveNodes) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n - 1) return;
  int size = rangeLast[i] - rangeFirst[i] + 1;
  if (i != 0 && size <= (int)LEAF_MAX) return;   // absorbed into a leaf of its parent
  atomicAdd(liveNodes, 1u);
  float b[2][6]; int ref[2];
  const int cs[2] = {left[i], right[i]};
  for (int c = 0; c < 2; ++c) {
    if (cs[c] < 0) {
      int l = ~cs[c];
      for (int k = 0; k < 3; ++k) { b[c][k] = leafLo[3 * (size_t)l + k]; b[c][3 + k] = leafHi[3 * (size_t)l + k]; }
      ref[c] = ~((l << 3) | 1);
    } else {
      int ch = cs[c];
      for (int k = 0; k < 3; ++k) { b[c][k] = nodeLo[3 * (size_t)ch + k]; b[c][3 + k] = nodeHi[3 * (size_t)ch + k]; }
      int csz = rangeLast[ch] - rangeFirst[ch] + 1;
      ref[c] = (csz <= (int)LEAF_MAX) ? ~((rangeFirst[ch] << 3) | csz) : ch;
    }
  }
  out[4 * (size_t)i + 0] = make_float4(b[0][0], b[0][1], b[0][2], b[0][3]);
  out[4 * (size_t)i + 1] = make_float4(b[0][4], b[0][5], b[1][0], b[1][1]);
  out[4 * (size_t)i + 2] = make_float4(b[1][2], b[1][3], b[1][4], b[1][5]);
  out[4 * (size_t)i + 3] = make_float4(__int_as_float(ref[0]), __int_as_float(ref[1]), 0.f, 0.f);
}

// Collapse of the binary radix tree into 4-wide nodes (EID_BVH_WIDTH == 4), level by level from the root: a wide node starts
// from a binary node's two children and twice replaces its largest-area expandable child by that child's two children
// (the surface-area-greedy collapse used by wide-BVH CPU tracers).  Subtrees with <= LEAF_MAX triangles are leaves.
// 128-byte node = 8 x float4: lo.x[4], lo.y[4], lo.z[4], hi.x[4], hi.y[4], hi.z[4], child refs[4], unused.
// Empty slots: box at +3e38 (never entered) and ref ~0 (a leaf of 0 triangles).
__global__ void k_collapse4(int nIn, const int2* __restrict__ in, int2* __restrict__ out, unsigned int* __restrict__ counters,
                            const int* __restrict__ left, const int* __restrict__ right, const int* __restrict__ rangeFirst,
                            const int* __restrict__ rangeLast, const float* __restrict__ leafLo, const float* __restrict__ leafHi,
                            const float* __restrict__ nodeLo, const float* __restrict__ nodeHi, float4* __restrict__ wide) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nIn) return;
  const int b = in[i].x, w = in[i].y;
  int c[4] = {left[b], right[b], 0, 0};
  int n = 2;
  auto sizeOf = [&](int r) { return r < 0 ? 1 : rangeLast[r] - rangeFirst[r] + 1; };
  auto boxOf = [&](int r, float* lo, float* hi) {
    const float* l = r < 0 ? leafLo + 3 * (size_t)(~r) : nodeLo + 3 * (size_t)r;
    const float* h = r < 0 ? leafHi + 3 * (size_t)(~r) : nodeHi + 3 * (size_t)r;
    for (int k = 0; k < 3; ++k) { lo[k] = l[k]; hi[k] = h[k]; }
  };
  while (n < 4) {
    int best = -1; float bestA = -1.f;
    for (int j = 0; j < n; ++j) {
      if (c[j] >= 0 && sizeOf(c[j]) > (int)LEAF_MAX) {
        float lo[3], hi[3]; boxOf(c[j], lo, hi);
        float ex = hi[0] - lo[0], ey = hi[1] - lo[1], ez = hi[2] - lo[2];
        float a = ex * ey + ey * ez + ez * ex;
        if (a > bestA) { bestA = a; best = j; }
      }
    }
    if (best < 0) break;
    const int r = c[best];
    c[best] = left[r];
    c[n++] = right[r];
  }
  int inner = 0;
  for (int j = 0; j < n; ++j) inner += (c[j] >= 0 && sizeOf(c[j]) > (int)LEAF_MAX) ? 1 : 0;
  unsigned int wbase = 0, obase = 0;
  if (inner) { wbase = atomicAdd(&counters[1], (unsigned int)inner); obase = atomicAdd(&counters[0], (unsigned int)inner); }
  float lo[4][3], hi[4][3]; int ref[4];
  int k = 0;
  for (int j = 0; j < 4; ++j) {
    if (j >= n) { for (int a = 0; a < 3; ++a) { lo[j][a] = 3e38f; hi[j][a] = 3e38f; } ref[j] = ~0; continue; }
    boxOf(c[j], lo[j], hi[j]);
    const int sz = sizeOf(c[j]);
    if (c[j] < 0) ref[j] = ~(((~c[j]) << 3) | 1);
    else if (sz <= (int)LEAF_MAX) ref[j] = ~((rangeFirst[c[j]] << 3) | sz);
    else { ref[j] = (int)(wbase + k); out[obase + k] = make_int2(c[j], (int)(wbase + k)); ++k; }
  }
  float4* o = wide + 8 * (size_t)w;
  for (int a = 0; a < 3; ++a) {
    o[a] = make_float4(lo[0][a], lo[1][a], lo[2][a], lo[3][a]);
    o[3 + a] = make_float4(hi[0][a], hi[1][a], hi[2][a], hi[3][a]);
  }
  o[6] = make_float4(__int_as_float(ref[0]), __int_as_float(ref[1]), __int_as_float(ref[2]), __int_as_float(ref[3]));
  o[7] = make_float4(0.f, 0.f, 0.f, 0.f);
}

#if EID_NODE_Q8
// 128-byte wide node -> 64-byte node with 8-bit child planes (accel.h).  origin = lower corner of the union of the node's child boxes,
// scale = extent / 255 nudged up until origin + 255 * scale reaches the upper corner; a lower plane is rounded down and an upper plane up,
// each checked against the value the traversal decodes (fmaf(q, scale, origin)), so the decoded box always contains the exact (padded) one.
__global__ void k_quantize4(uint32_t n, const float4* __restrict__ wide, uint4* __restrict__ q) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4* w = wide + 8 * (size_t)i;
  float lo[3][4], hi[3][4];
  for (int a = 0; a < 3; ++a) {
    const float4 l = w[a], h = w[3 + a];
    lo[a][0] = l.x; lo[a][1] = l.y; lo[a][2] = l.z; lo[a][3] = l.w; hi[a][0] = h.x; hi[a][1] = h.y; hi[a][2] = h.z; hi[a][3] = h.w;
  }
  const float4 rf = w[6];
  const int ref[4] = {__float_as_int(rf.x), __float_as_int(rf.y), __float_as_int(rf.z), __float_as_int(rf.w)};
  float org[3], scl[3]; uint32_t qlo[3] = {0, 0, 0}, qhi[3] = {0, 0, 0};
  for (int a = 0; a < 3; ++a) {
    float mn = 3e38f, mx = -3e38f;
    for (int c = 0; c < 4; ++c) if (ref[c] != ~0) { mn = fminf(mn, lo[a][c]); mx = fmaxf(mx, hi[a][c]); }
    if (!(mn <= mx)) { mn = 0.f; mx = 0.f; }
    float sc = (mx - mn) / 255.0f;
    if (!(sc > 0.0f)) sc = 1e-30f;
    while (fmaf(255.0f, sc, mn) < mx) sc = __int_as_float(__float_as_int(sc) + 1);
    org[a] = mn; scl[a] = sc;
    for (int c = 0; c < 4; ++c) {
      int l = 255, h = 0;                                          // empty slot: inverted box (the traversal also tests the reference)
      if (ref[c] != ~0) {
        l = (int)floorf((lo[a][c] - mn) / sc); l = max(0, min(255, l));
        while (l > 0 && fmaf((float)l, sc, mn) > lo[a][c]) --l;
        h = (int)ceilf((hi[a][c] - mn) / sc); h = max(0, min(255, h));
        while (h < 255 && fmaf((float)h, sc, mn) < hi[a][c]) ++h;
      }
      qlo[a] |= (uint32_t)l << (8 * c); qhi[a] |= (uint32_t)h << (8 * c);
    }
  }
  uint4* o = q + 4 * (size_t)i;
  o[0] = make_uint4(__float_as_uint(org[0]), __float_as_uint(org[1]), __float_as_uint(org[2]), __float_as_uint(scl[0]));
  o[1] = make_uint4(__float_as_uint(scl[1]), __float_as_uint(scl[2]), qlo[0], qlo[1]);
  o[2] = make_uint4(qlo[2], qhi[0], qhi[1], qhi[2]);
  o[3] = make_uint4((uint32_t)ref[0], (uint32_t)ref[1], (uint32_t)ref[2], (uint32_t)ref[3]);
}
#endif

// ------------------------------------------------------------------------------------------------
// batch ray query (parity tap for ClosestHit / AnyHit)
// ------------------------------------------------------------------------------------------------
__global__ void k_trace_batch(AccelView A, const InstanceXform* __restrict__ instances, const float* __restrict__ rays, uint32_t n, int anyHit,
                              eid_hit* __restrict__ hits) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* r = rays + 8 * (size_t)i;
  f3 o = mk3(r[0], r[1], r[2]), d = mk3(r[4], r[5], r[6]);
  RayHit h;
  eid_hit out;
  if (anyHit) {
    bool occ = A.twoLevel ? traverse2<true>(A, o, d, r[3], h) : traverse<true>(A, o, d, r[3], h);
    out.hitT = occ ? 0.f : 1e28f; out.primitiveID = out.instanceID = out.instanceCustomIndex = -1; out.baryU = out.baryV = 0.f;
  } else {
    bool ok = A.twoLevel ? traverse2<false>(A, o, d, r[3], h) : traverse<false>(A, o, d, r[3], h);
    if (ok) { out.hitT = h.t; out.primitiveID = h.prim; out.instanceID = h.inst; out.instanceCustomIndex = instances[h.inst].primMesh; out.baryU = h.u; out.baryV = h.v; }
    else { out.hitT = 1e28f; out.primitiveID = out.instanceID = out.instanceCustomIndex = -1; out.baryU = out.baryV = 0.f; }
  }
  hits[i] = out;
}

// ------------------------------------------------------------------------------------------------
// Two-level build (AccelStructure::create, accelstruct.cpp:55-162: one BLAS per prim mesh, one TLAS instance per node)
// ------------------------------------------------------------------------------------------------
// object-space triangles of ONE prim mesh: 48-byte records (p0, p1, p2, primitiveID), unpadded boxes, mesh bounds
__global__ void k_emit_object(DeviceSceneView sc, int primMesh, uint32_t nTri, float4* __restrict__ triOut, float* __restrict__ boxLo, float* __restrict__ boxHi,
                              BuildBounds* bounds) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
  if (t < nTri) {
    const InstanceData gi = sc.geoInfo[primMesh];
    const uint32_t* idx = (const uint32_t*)(uintptr_t)gi.indexAddress;
    const VertexAttributes* vtx = (const VertexAttributes*)(uintptr_t)gi.vertexAddress;
    f3 p[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) p[k] = ld3(vtx[idx[3 * t + k]].position);
    triOut[3 * (size_t)t + 0] = make_float4(p[0].x, p[0].y, p[0].z, p[1].x);
    triOut[3 * (size_t)t + 1] = make_float4(p[1].y, p[1].z, p[2].x, p[2].y);
    triOut[3 * (size_t)t + 2] = make_float4(p[2].z, __int_as_float((int)t), 0.f, 0.f);
    lo[0] = fminf(p[0].x, fminf(p[1].x, p[2].x)); hi[0] = fmaxf(p[0].x, fmaxf(p[1].x, p[2].x));
    lo[1] = fminf(p[0].y, fminf(p[1].y, p[2].y)); hi[1] = fmaxf(p[0].y, fmaxf(p[1].y, p[2].y));
    lo[2] = fminf(p[0].z, fminf(p[1].z, p[2].z)); hi[2] = fmaxf(p[0].z, fmaxf(p[1].z, p[2].z));
#pragma unroll
    for (int k = 0; k < 3; ++k) { boxLo[3 * (size_t)t + k] = lo[k]; boxHi[3 * (size_t)t + k] = hi[k]; }
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float l = lo[k], h = hi[k];
    for (int o = 16; o > 0; o >>= 1) { l = fminf(l, __shfl_xor_sync(0xffffffffu, l, o)); h = fmaxf(h, __shfl_xor_sync(0xffffffffu, h, o)); }
    if ((threadIdx.x & 31) == 0) { atomicMin(&bounds->lo[k], floatFlip(l)); atomicMax(&bounds->hi[k], floatFlip(h)); }
  }
}
// a tree that was built on its own is appended to the shared arrays: inner references move by nodeBase, leaf ranges by primBase
__global__ void k_rebase(float4* __restrict__ nodes, uint32_t count, int nodeBase, int primBase) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  float4 rf = nodes[8 * (size_t)i + 6];
  int r[4] = {__float_as_int(rf.x), __float_as_int(rf.y), __float_as_int(rf.z), __float_as_int(rf.w)};
  for (int k = 0; k < 4; ++k) {
    if (r[k] >= 0) r[k] += nodeBase;
    else { const uint32_t u = ~(uint32_t)r[k]; if (u & 7u) r[k] = ~(int)((((u >> 3) + (uint32_t)primBase) << 3) | (u & 7u)); }
  }
  nodes[8 * (size_t)i + 6] = make_float4(__int_as_float(r[0]), __int_as_float(r[1]), __int_as_float(r[2]), __int_as_float(r[3]));
}
static int rebaseRef(int r, int nodeBase, int primBase) {
  if (r >= 0) return r + nodeBase;
  const uint32_t u = ~(uint32_t)r;
  return (u & 7u) ? ~(int)((((u >> 3) + (uint32_t)primBase) << 3) | (u & 7u)) : r;
}

struct TreeBuild { float4* nodes = nullptr; uint32_t nodeCount = 0, levels = 0; int32_t rootRef = ~0; };

// PREFER_FAST_TRACE build (sah_host.h): the padded boxes go to the host, the binned-SAH builder returns the primitive order and the
// binary tree in the arrays k_hierarchy would have filled; refit and the 4-wide collapse run on the GPU as for the Morton build.
static void sahOrder(uint32_t n, const float* lo0, const float* hi0, uint32_t* valsSorted, BinaryTreeHost& H) {
  std::vector<float> hlo(3 * (size_t)n), hhi(3 * (size_t)n);
  CUDA_CHECK(cudaMemcpy(hlo.data(), lo0, hlo.size() * 4, cudaMemcpyDeviceToHost));
  CUDA_CHECK(cudaMemcpy(hhi.data(), hi0, hhi.size() * 4, cudaMemcpyDeviceToHost));
  buildSahTree(n, hlo.data(), hhi.data(), H);
  CUDA_CHECK(cudaMemcpy(valsSorted, H.order.data(), (size_t)n * 4, cudaMemcpyHostToDevice));
}
static void sahUpload(const BinaryTreeHost& H, int* left, int* right, int* parI, int* parL, int* rf, int* rl) {
  const size_t ni = H.left.size() * 4;
  CUDA_CHECK(cudaMemcpy(left, H.left.data(), ni, cudaMemcpyHostToDevice)); CUDA_CHECK(cudaMemcpy(right, H.right.data(), ni, cudaMemcpyHostToDevice));
  CUDA_CHECK(cudaMemcpy(parI, H.parentInner.data(), ni, cudaMemcpyHostToDevice)); CUDA_CHECK(cudaMemcpy(parL, H.parentLeaf.data(), H.parentLeaf.size() * 4, cudaMemcpyHostToDevice));
  CUDA_CHECK(cudaMemcpy(rf, H.rangeFirst.data(), ni, cudaMemcpyHostToDevice)); CUDA_CHECK(cudaMemcpy(rl, H.rangeLast.data(), ni, cudaMemcpyHostToDevice));
}

// The flat builder's pipeline (Morton keys of the padded boxes, radix sort, Karras tree, refit, 4-wide collapse) over ANY list of
// 48-byte primitive records with boxes: the triangles of one prim mesh (BLAS) or the instances (TLAS).  primOut receives the records in
// Morton order (what the leaves index); lo0 / hi0 are padded in place; `bounds` must hold the union of the boxes.
static void buildTree(uint32_t n, const float4* primTmp, float* lo0, float* hi0, const BuildBounds* bounds, float4* primOut, TreeBuild& T, bool sah) {
#if EID_BVH_WIDTH != 4 || EID_NODE_Q8
  raise(EID_ERR_UNSUPPORTED, "the two-level build needs the default 128-byte BVH4 node");
#else
  const int B = 256;
  const uint32_t G = (n + B - 1) / B;
  float *lo1 = nullptr, *hi1 = nullptr; unsigned long long *keys = nullptr, *keysSorted = nullptr; uint32_t *vals = nullptr, *valsSorted = nullptr;
  int *left = nullptr, *right = nullptr, *parI = nullptr, *parL = nullptr, *rf = nullptr, *rl = nullptr, *height = nullptr;
  float *nlo = nullptr, *nhi = nullptr; unsigned int *arrived = nullptr; void* tmp = nullptr;
  float4* wideTmp = nullptr; int2 *q0 = nullptr, *q1 = nullptr; unsigned int* counters = nullptr;
  auto freeAll = [&]() {
    cudaFree(lo1); cudaFree(hi1); cudaFree(keys); cudaFree(keysSorted); cudaFree(vals); cudaFree(valsSorted); cudaFree(left); cudaFree(right); cudaFree(parI);
    cudaFree(parL); cudaFree(rf); cudaFree(rl); cudaFree(height); cudaFree(nlo); cudaFree(nhi); cudaFree(arrived); cudaFree(tmp); cudaFree(wideTmp); cudaFree(q0);
    cudaFree(q1); cudaFree(counters);
  };
  try {
    CUDA_CHECK(cudaMalloc(&lo1, (size_t)n * 12)); CUDA_CHECK(cudaMalloc(&hi1, (size_t)n * 12));
    CUDA_CHECK(cudaMalloc(&keys, (size_t)n * 8)); CUDA_CHECK(cudaMalloc(&keysSorted, (size_t)n * 8));
    CUDA_CHECK(cudaMalloc(&vals, (size_t)n * 4)); CUDA_CHECK(cudaMalloc(&valsSorted, (size_t)n * 4));
    k_morton<<<G, B>>>(n, lo0, hi0, bounds, keys, vals);
    BinaryTreeHost H;
    sah = sah && n > 1;
    if (sah) sahOrder(n, lo0, hi0, valsSorted, H);
    else {
      size_t tmpBytes = 0;
      CUDA_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, tmpBytes, keys, keysSorted, vals, valsSorted, (int)n, 0, 63));
      CUDA_CHECK(cudaMalloc(&tmp, tmpBytes));
      CUDA_CHECK(cub::DeviceRadixSort::SortPairs(tmp, tmpBytes, keys, keysSorted, vals, valsSorted, (int)n, 0, 63));
    }
    k_reorder<<<G, B>>>(n, valsSorted, primTmp, lo0, hi0, primOut, lo1, hi1);
    if (n == 1) { T.nodes = nullptr; T.nodeCount = 0; T.levels = 0; T.rootRef = ~((0 << 3) | 1); freeAll(); return; }
    const uint32_t nInner = n - 1;
    CUDA_CHECK(cudaMalloc(&left, (size_t)nInner * 4)); CUDA_CHECK(cudaMalloc(&right, (size_t)nInner * 4));
    CUDA_CHECK(cudaMalloc(&parI, (size_t)nInner * 4)); CUDA_CHECK(cudaMalloc(&parL, (size_t)n * 4));
    CUDA_CHECK(cudaMalloc(&rf, (size_t)nInner * 4)); CUDA_CHECK(cudaMalloc(&rl, (size_t)nInner * 4));
    CUDA_CHECK(cudaMalloc(&height, (size_t)nInner * 4));
    CUDA_CHECK(cudaMalloc(&nlo, (size_t)nInner * 12)); CUDA_CHECK(cudaMalloc(&nhi, (size_t)nInner * 12));
    CUDA_CHECK(cudaMalloc(&arrived, (size_t)nInner * 4));
    CUDA_CHECK(cudaMemset(arrived, 0, (size_t)nInner * 4));
    if (sah) sahUpload(H, left, right, parI, parL, rf, rl);
    else k_hierarchy<<<(nInner + B - 1) / B, B>>>((int)n, keysSorted, left, right, parI, parL, rf, rl);
    k_refit<<<G, B>>>((int)n, left, right, parI, parL, lo1, hi1, nlo, nhi, height, arrived);
    if (n <= LEAF_MAX) { T.nodes = nullptr; T.nodeCount = 0; T.levels = 0; T.rootRef = ~(int)((0u << 3) | n); freeAll(); return; }   // the whole list is one leaf
    CUDA_CHECK(cudaMalloc(&wideTmp, (size_t)nInner * 128));
    CUDA_CHECK(cudaMalloc(&q0, (size_t)nInner * 8)); CUDA_CHECK(cudaMalloc(&q1, (size_t)nInner * 8));
    CUDA_CHECK(cudaMalloc(&counters, 8));
    const int2 rootItem = make_int2(0, 0);
    const unsigned int init[2] = {0u, 1u};
    CUDA_CHECK(cudaMemcpy(q0, &rootItem, 8, cudaMemcpyHostToDevice));
    CUDA_CHECK(cudaMemcpy(counters, init, 8, cudaMemcpyHostToDevice));
    unsigned int nIn = 1, levels = 0;
    while (nIn) {
      CUDA_CHECK(cudaMemsetAsync(counters, 0, 4));
      k_collapse4<<<(nIn + 127) / 128, 128>>>((int)nIn, q0, q1, counters, left, right, rf, rl, lo1, hi1, nlo, nhi, wideTmp);
      CUDA_CHECK(cudaMemcpy(&nIn, counters, 4, cudaMemcpyDeviceToHost));
      std::swap(q0, q1);
      if (++levels > 200) raise(EID_ERR_UNSUPPORTED, "BVH collapse did not terminate");
    }
    unsigned int wideCount = 0;
    CUDA_CHECK(cudaMemcpy(&wideCount, counters + 1, 4, cudaMemcpyDeviceToHost));
    CUDA_CHECK(cudaMalloc(&T.nodes, (size_t)wideCount * 128));
    CUDA_CHECK(cudaMemcpy(T.nodes, wideTmp, (size_t)wideCount * 128, cudaMemcpyDeviceToDevice));
    T.nodeCount = wideCount; T.levels = levels; T.rootRef = 0;
  } catch (...) { freeAll(); throw; }
  freeAll();
#endif
}

static void buildAccelTwoLevel(eid_scene* s, eid_accel* a, bool sah) {
  const SceneHost& H = s->host;
  CUDA_CHECK(cudaSetDevice(s->dev.device));
  const size_t nMesh = H.gltf.primMeshes.size(), nInst = H.instances.size();
  a->scene = s; a->twoLevel = true;
  cudaEvent_t ev0, ev1;
  CUDA_CHECK(cudaEventCreate(&ev0)); CUDA_CHECK(cudaEventCreate(&ev1));
  CUDA_CHECK(cudaEventRecord(ev0, 0));
  // ---- bottom level: one tree per prim mesh that some instance uses ----
  std::vector<uint32_t> meshTris(nMesh, 0);
  std::vector<char> used(nMesh, 0);
  for (const InstanceXform& X : H.instances) { used[X.primMesh] = 1; meshTris[X.primMesh] = X.triangleCount; }
  uint64_t unique = 0;
  for (size_t m = 0; m < nMesh; ++m) if (used[m]) unique += meshTris[m];
  if (unique >= (1ull << 28) - 8) raise(EID_ERR_UNSUPPORTED, "eid_accel_build: more than 2^28 - 8 unique triangles");
  a->triCount = (uint32_t)unique; a->uniqueTriangles = unique;
  struct Blas { TreeBuild T; uint32_t primBase = 0, nodeBase = 0; float lo[3], hi[3]; bool valid = false; };
  std::vector<Blas> blas(nMesh);
  float4* triTmp = nullptr; float *lo0 = nullptr, *hi0 = nullptr; BuildBounds* bounds = nullptr;
  auto freeTmp = [&]() { cudaFree(triTmp); cudaFree(lo0); cudaFree(hi0); triTmp = nullptr; lo0 = hi0 = nullptr; };
  try {
    CUDA_CHECK(cudaMalloc(&a->tris, std::max<size_t>(1, unique) * 48));
    CUDA_CHECK(cudaMalloc(&bounds, sizeof(BuildBounds)));
    uint32_t primBase = 0, nodeTotal = 0, maxBlasLevels = 0;
    for (size_t m = 0; m < nMesh; ++m) {
      if (!used[m] || meshTris[m] == 0) continue;
      const uint32_t n = meshTris[m];
      CUDA_CHECK(cudaMalloc(&triTmp, (size_t)n * 48)); CUDA_CHECK(cudaMalloc(&lo0, (size_t)n * 12)); CUDA_CHECK(cudaMalloc(&hi0, (size_t)n * 12));
      k_init_bounds<<<1, 1>>>(bounds);
      k_emit_object<<<(n + 255) / 256, 256>>>(s->dev.view(H), (int)m, n, triTmp, lo0, hi0, bounds);
      BuildBounds hb;
      CUDA_CHECK(cudaMemcpy(&hb, bounds, sizeof(hb), cudaMemcpyDeviceToHost));
      Blas& b = blas[m];
      for (int k = 0; k < 3; ++k) {
        const int l = hb.lo[k], h = hb.hi[k];
        int li = l >= 0 ? l : l ^ 0x7fffffff, hi_ = h >= 0 ? h : h ^ 0x7fffffff;
        memcpy(&b.lo[k], &li, 4); memcpy(&b.hi[k], &hi_, 4);
      }
      buildTree(n, triTmp, lo0, hi0, bounds, a->tris + 3 * (size_t)primBase, b.T, sah);
      b.primBase = primBase; b.nodeBase = nodeTotal; b.valid = true;
      primBase += n; nodeTotal += b.T.nodeCount; maxBlasLevels = std::max(maxBlasLevels, b.T.levels);
      a->blasCount++;
      freeTmp();
    }
    CUDA_CHECK(cudaMalloc(&a->nodes, std::max<size_t>(1, nodeTotal) * 128));
    for (size_t m = 0; m < nMesh; ++m) {
      Blas& b = blas[m];
      if (!b.valid) continue;
      if (b.T.nodeCount) {
        k_rebase<<<(b.T.nodeCount + 127) / 128, 128>>>(b.T.nodes, b.T.nodeCount, (int)b.nodeBase, (int)b.primBase);
        CUDA_CHECK(cudaMemcpy(a->nodes + 8 * (size_t)b.nodeBase, b.T.nodes, (size_t)b.T.nodeCount * 128, cudaMemcpyDeviceToDevice));
        cudaFree(b.T.nodes); b.T.nodes = nullptr;
      }
      b.T.rootRef = rebaseRef(b.T.rootRef, (int)b.nodeBase, (int)b.primBase);
    }
    a->nodeCount = nodeTotal; a->nodeAlloc = nodeTotal; a->rootRef = ~0;
    // ---- top level: the instances' world boxes (the object bounds through objectToWorld, 8 corners, padded) ----
    std::vector<float> ilo, ihi; std::vector<float4> irec;
    float slo[3] = {3e38f, 3e38f, 3e38f}, shi[3] = {-3e38f, -3e38f, -3e38f};
    for (size_t i = 0; i < nInst; ++i) {
      const InstanceXform& X = H.instances[i];
      const Blas& b = blas[X.primMesh];
      if (!b.valid) continue;
      double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
      for (int c = 0; c < 8; ++c) {
        const double p[3] = {(c & 1) ? b.hi[0] : b.lo[0], (c & 2) ? b.hi[1] : b.lo[1], (c & 4) ? b.hi[2] : b.lo[2]};
        for (int k = 0; k < 3; ++k) {
          const double w = (double)X.objectToWorld[k] * p[0] + (double)X.objectToWorld[3 + k] * p[1] + (double)X.objectToWorld[6 + k] * p[2] + (double)X.objectToWorld[9 + k];
          lo[k] = std::min(lo[k], w); hi[k] = std::max(hi[k], w);
        }
      }
      for (int k = 0; k < 3; ++k) {
        // the flat build transforms every vertex in fp32: pad by a few ulps of the largest coordinate plus a fraction of the extent
        const double pad = 1e-5 * (hi[k] - lo[k]) + 2e-6 * std::max(std::fabs(lo[k]), std::fabs(hi[k])) + 1e-7;
        const float l = (float)(lo[k] - pad), h = (float)(hi[k] + pad);
        ilo.push_back(l); ihi.push_back(h);
        slo[k] = std::min(slo[k], l); shi[k] = std::max(shi[k], h);
      }
      float fi, fr; const int ii = (int)i, rr = b.T.rootRef;
      memcpy(&fi, &ii, 4); memcpy(&fr, &rr, 4);
      irec.push_back(make_float4(fi, fr, 0.f, 0.f));
      irec.push_back(make_float4(0.f, 0.f, 0.f, 0.f)); irec.push_back(make_float4(0.f, 0.f, 0.f, 0.f));
    }
    const uint32_t nTop = (uint32_t)(ilo.size() / 3);
    a->tlasPrimCount = nTop;
    CUDA_CHECK(cudaMalloc(&a->tlasPrims, std::max<size_t>(1, nTop) * 48));
    uint32_t tlasLevels = 0;
    if (nTop) {
      BuildBounds hb;
      for (int k = 0; k < 3; ++k) {
        int li, hi_; memcpy(&li, &slo[k], 4); memcpy(&hi_, &shi[k], 4);
        hb.lo[k] = li >= 0 ? li : li ^ 0x7fffffff; hb.hi[k] = hi_ >= 0 ? hi_ : hi_ ^ 0x7fffffff;
      }
      CUDA_CHECK(cudaMemcpy(bounds, &hb, sizeof(hb), cudaMemcpyHostToDevice));
      CUDA_CHECK(cudaMalloc(&triTmp, (size_t)nTop * 48)); CUDA_CHECK(cudaMalloc(&lo0, (size_t)nTop * 12)); CUDA_CHECK(cudaMalloc(&hi0, (size_t)nTop * 12));
      CUDA_CHECK(cudaMemcpy(triTmp, irec.data(), (size_t)nTop * 48, cudaMemcpyHostToDevice));
      CUDA_CHECK(cudaMemcpy(lo0, ilo.data(), (size_t)nTop * 12, cudaMemcpyHostToDevice));
      CUDA_CHECK(cudaMemcpy(hi0, ihi.data(), (size_t)nTop * 12, cudaMemcpyHostToDevice));
      TreeBuild T;
      buildTree(nTop, triTmp, lo0, hi0, bounds, a->tlasPrims, T, sah);
      a->tlasNodes = T.nodes; a->tlasNodeCount = T.nodeCount; a->tlasRootRef = T.rootRef; tlasLevels = T.levels;
      freeTmp();
    }
    if (!a->tlasNodes) CUDA_CHECK(cudaMalloc(&a->tlasNodes, 128));
    a->maxDepth = tlasLevels + maxBlasLevels;
    if (3 * (tlasLevels + maxBlasLevels) + 8 >= EID_STACK_SIZE) raise(EID_ERR_UNSUPPORTED, "two-level BVH depth %u + %u exceeds the traversal stack (%d)", tlasLevels, maxBlasLevels, EID_STACK_SIZE);
    CUDA_CHECK(cudaEventRecord(ev1, 0));
    CUDA_CHECK(cudaDeviceSynchronize());
    CUDA_CHECK(cudaGetLastError());
    CUDA_CHECK(cudaEventElapsedTime(&a->buildMs, ev0, ev1));
  } catch (...) {
    freeTmp(); cudaFree(bounds);
    for (Blas& b : blas) cudaFree(b.T.nodes);
    cudaEventDestroy(ev0); cudaEventDestroy(ev1);
    throw;
  }
  cudaFree(bounds);
  cudaEventDestroy(ev0); cudaEventDestroy(ev1);
}

static void buildAccel(eid_scene* s, eid_accel* a, bool sah) {
  const SceneHost& H = s->host;
  CUDA_CHECK(cudaSetDevice(s->dev.device));
  // leaf references carry (firstTriangle << 3 | count) in 31 bits, and 0x80000000 is the traversal's "done" sentinel
  if (H.triangleInstances >= (1ull << 28) - 8) raise(EID_ERR_UNSUPPORTED, "eid_accel_build: more than 2^28 - 8 triangle instances");
  const uint32_t nTri = (uint32_t)H.triangleInstances;
  a->scene = s; a->triCount = nTri;
  cudaEvent_t ev0, ev1;
  CUDA_CHECK(cudaEventCreate(&ev0)); CUDA_CHECK(cudaEventCreate(&ev1));
  CUDA_CHECK(cudaEventRecord(ev0, 0));
  if (nTri == 0) {
    CUDA_CHECK(cudaMalloc(&a->nodes, EID_NODE_BYTES)); CUDA_CHECK(cudaMalloc(&a->tris, 48));
    a->rootRef = ~0; a->nodeCount = 0; a->maxDepth = 0;   // empty leaf
    CUDA_CHECK(cudaEventDestroy(ev0)); CUDA_CHECK(cudaEventDestroy(ev1));
    return;
  }
  const int B = 256;
  const uint32_t G = (nTri + B - 1) / B;
  float4 *triTmp = nullptr; float *lo0 = nullptr, *hi0 = nullptr, *lo1 = nullptr, *hi1 = nullptr;
  unsigned long long *keys = nullptr, *keysSorted = nullptr; uint32_t *vals = nullptr, *valsSorted = nullptr;
  BuildBounds* bounds = nullptr;
  int *left = nullptr, *right = nullptr, *parI = nullptr, *parL = nullptr, *rf = nullptr, *rl = nullptr, *height = nullptr;
  float *nlo = nullptr, *nhi = nullptr; unsigned int *arrived = nullptr, *live = nullptr; void* tmp = nullptr;
  auto freeAll = [&]() {
    cudaFree(triTmp); cudaFree(lo0); cudaFree(hi0); cudaFree(lo1); cudaFree(hi1); cudaFree(keys); cudaFree(keysSorted); cudaFree(vals);
    cudaFree(valsSorted); cudaFree(bounds); cudaFree(left); cudaFree(right); cudaFree(parI); cudaFree(parL); cudaFree(rf); cudaFree(rl);
    cudaFree(height); cudaFree(nlo); cudaFree(nhi); cudaFree(arrived); cudaFree(live); cudaFree(tmp);
  };
  try {
    CUDA_CHECK(cudaMalloc(&triTmp, (size_t)nTri * 48)); CUDA_CHECK(cudaMalloc(&a->tris, (size_t)nTri * 48));
    CUDA_CHECK(cudaMalloc(&lo0, (size_t)nTri * 12)); CUDA_CHECK(cudaMalloc(&hi0, (size_t)nTri * 12));
    CUDA_CHECK(cudaMalloc(&lo1, (size_t)nTri * 12)); CUDA_CHECK(cudaMalloc(&hi1, (size_t)nTri * 12));
    CUDA_CHECK(cudaMalloc(&keys, (size_t)nTri * 8)); CUDA_CHECK(cudaMalloc(&keysSorted, (size_t)nTri * 8));
    CUDA_CHECK(cudaMalloc(&vals, (size_t)nTri * 4)); CUDA_CHECK(cudaMalloc(&valsSorted, (size_t)nTri * 4));
    CUDA_CHECK(cudaMalloc(&bounds, sizeof(BuildBounds)));
    k_init_bounds<<<1, 1>>>(bounds);
    k_emit_triangles<<<G, B>>>(s->dev.view(H), s->dev.instFirstTri, (uint32_t)H.instances.size(), nTri, triTmp, lo0, hi0, bounds);
    k_morton<<<G, B>>>(nTri, lo0, hi0, bounds, keys, vals);
    BinaryTreeHost H;
    sah = sah && nTri > 1;
    if (sah) sahOrder(nTri, lo0, hi0, valsSorted, H);
    else {
      size_t tmpBytes = 0;
      CUDA_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, tmpBytes, keys, keysSorted, vals, valsSorted, (int)nTri, 0, 63));
      CUDA_CHECK(cudaMalloc(&tmp, tmpBytes));
      CUDA_CHECK(cub::DeviceRadixSort::SortPairs(tmp, tmpBytes, keys, keysSorted, vals, valsSorted, (int)nTri, 0, 63));
    }
    k_reorder<<<G, B>>>(nTri, valsSorted, triTmp, lo0, hi0, a->tris, lo1, hi1);
    const uint32_t nInner = nTri > 1 ? nTri - 1 : 1;
    if (nTri == 1) {
      // single triangle: the root reference is the leaf itself, no inner node exists
      CUDA_CHECK(cudaMalloc(&a->nodes, EID_NODE_BYTES));
      CUDA_CHECK(cudaMemset(a->nodes, 0, EID_NODE_BYTES));
      a->rootRef = ~((0 << 3) | 1); a->nodeCount = 0; a->maxDepth = 0;
    } else {
      CUDA_CHECK(cudaMalloc(&left, (size_t)nInner * 4)); CUDA_CHECK(cudaMalloc(&right, (size_t)nInner * 4));
      CUDA_CHECK(cudaMalloc(&parI, (size_t)nInner * 4)); CUDA_CHECK(cudaMalloc(&parL, (size_t)nTri * 4));
      CUDA_CHECK(cudaMalloc(&rf, (size_t)nInner * 4)); CUDA_CHECK(cudaMalloc(&rl, (size_t)nInner * 4));
      CUDA_CHECK(cudaMalloc(&height, (size_t)nInner * 4));
      CUDA_CHECK(cudaMalloc(&nlo, (size_t)nInner * 12)); CUDA_CHECK(cudaMalloc(&nhi, (size_t)nInner * 12));
      CUDA_CHECK(cudaMalloc(&arrived, (size_t)nInner * 4)); CUDA_CHECK(cudaMalloc(&live, 4));
      CUDA_CHECK(cudaMemset(arrived, 0, (size_t)nInner * 4)); CUDA_CHECK(cudaMemset(live, 0, 4));
      if (sah) sahUpload(H, left, right, parI, parL, rf, rl);
      else k_hierarchy<<<(nInner + B - 1) / B, B>>>((int)nTri, keysSorted, left, right, parI, parL, rf, rl);
      k_refit<<<G, B>>>((int)nTri, left, right, parI, parL, lo1, hi1, nlo, nhi, height, arrived);
#if EID_BVH_WIDTH == 2
      CUDA_CHECK(cudaMalloc(&a->nodes, (size_t)nInner * 64));
      CUDA_CHECK(cudaMemset(a->nodes, 0, (size_t)nInner * 64));
      k_pack_nodes<<<(nInner + B - 1) / B, B>>>((int)nTri, left, right, rf, rl, lo1, hi1, nlo, nhi, a->nodes, live);
      int rootHeight = 0; unsigned int liveNodes = 0;
      CUDA_CHECK(cudaMemcpy(&rootHeight, height, 4, cudaMemcpyDeviceToHost));
      CUDA_CHECK(cudaMemcpy(&liveNodes, live, 4, cudaMemcpyDeviceToHost));
      a->rootRef = 0; a->nodeCount = liveNodes; a->maxDepth = (uint32_t)rootHeight; a->nodeAlloc = nInner;
      if (rootHeight >= EID_STACK_SIZE) raise(EID_ERR_UNSUPPORTED, "BVH height %d exceeds the traversal stack (%d)", rootHeight, EID_STACK_SIZE);
#else
      // level-by-level collapse into 4-wide nodes; queues hold (binary node, wide node index)
      float4* wideTmp = nullptr; int2 *q0 = nullptr, *q1 = nullptr; unsigned int* counters = nullptr;
      try {
        CUDA_CHECK(cudaMalloc(&wideTmp, (size_t)nInner * 128));
        CUDA_CHECK(cudaMalloc(&q0, (size_t)nInner * 8)); CUDA_CHECK(cudaMalloc(&q1, (size_t)nInner * 8));
        CUDA_CHECK(cudaMalloc(&counters, 8));
        const int2 rootItem = make_int2(0, 0);
        const unsigned int init[2] = {0u, 1u};
        CUDA_CHECK(cudaMemcpy(q0, &rootItem, 8, cudaMemcpyHostToDevice));
        CUDA_CHECK(cudaMemcpy(counters, init, 8, cudaMemcpyHostToDevice));
        unsigned int nIn = 1, levels = 0;
        while (nIn) {
          CUDA_CHECK(cudaMemsetAsync(counters, 0, 4));
          k_collapse4<<<(nIn + 127) / 128, 128>>>((int)nIn, q0, q1, counters, left, right, rf, rl, lo1, hi1, nlo, nhi, wideTmp);
          CUDA_CHECK(cudaMemcpy(&nIn, counters, 4, cudaMemcpyDeviceToHost));
          std::swap(q0, q1);
          if (++levels > 200) raise(EID_ERR_UNSUPPORTED, "BVH collapse did not terminate");
        }
        unsigned int wideCount = 0;
        CUDA_CHECK(cudaMemcpy(&wideCount, counters + 1, 4, cudaMemcpyDeviceToHost));
        CUDA_CHECK(cudaMalloc(&a->nodes, (size_t)wideCount * EID_NODE_BYTES));      // trim to the live node count
#if EID_NODE_Q8
        k_quantize4<<<(wideCount + 127) / 128, 128>>>(wideCount, wideTmp, (uint4*)a->nodes);
#else
        CUDA_CHECK(cudaMemcpy(a->nodes, wideTmp, (size_t)wideCount * EID_NODE_BYTES, cudaMemcpyDeviceToDevice));
#endif
        a->rootRef = 0; a->nodeCount = wideCount; a->maxDepth = levels; a->nodeAlloc = wideCount;
        if (3 * levels + 1 >= EID_STACK_SIZE) raise(EID_ERR_UNSUPPORTED, "wide BVH depth %u exceeds the traversal stack (%d)", levels, EID_STACK_SIZE);
      } catch (...) { cudaFree(wideTmp); cudaFree(q0); cudaFree(q1); cudaFree(counters); throw; }
      cudaFree(wideTmp); cudaFree(q0); cudaFree(q1); cudaFree(counters);
#endif
    }
    CUDA_CHECK(cudaEventRecord(ev1, 0));
    CUDA_CHECK(cudaDeviceSynchronize());
    CUDA_CHECK(cudaGetLastError());
    CUDA_CHECK(cudaEventElapsedTime(&a->buildMs, ev0, ev1));
  } catch (...) {
    freeAll(); cudaEventDestroy(ev0); cudaEventDestroy(ev1);
    throw;
  }
  freeAll();
  cudaEventDestroy(ev0); cudaEventDestroy(ev1);
#if EID_FETCH_TEX
  int maxLinear = 0;
  CUDA_CHECK(cudaDeviceGetAttribute(&maxLinear, cudaDevAttrMaxTexture1DLinearWidth, s->dev.device));
  auto mkTex = [&](const float4* p, size_t count, bool asUint = false) {
    if (count > (size_t)maxLinear) raise(EID_ERR_UNSUPPORTED, "eid_accel_build: BVH array of %zu float4 exceeds the linear-texture limit (%d)", count, maxLinear);
    cudaResourceDesc rd{}; rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = (void*)p;
    rd.res.linear.desc = asUint ? cudaCreateChannelDesc<uint4>() : cudaCreateChannelDesc<float4>(); rd.res.linear.sizeInBytes = std::max<size_t>(count, 1) * 16;
    cudaTextureDesc td{}; td.readMode = cudaReadModeElementType;
    cudaTextureObject_t t = 0;
    CUDA_CHECK(cudaCreateTextureObject(&t, &rd, &td, nullptr));
    return t;
  };
  a->nodeTex = mkTex(a->nodes, (size_t)std::max(a->nodeAlloc, 1u) * (EID_NODE_BYTES / 16), EID_NODE_Q8 != 0);
  a->triTex = mkTex(a->tris, (size_t)std::max(nTri, 1u) * 3);
#endif
}

}  // namespace eid

using namespace eid;

// ------------------------------------------------------------------------------------------------
// C-ABI: misc + Scene + AccelStructure
// ------------------------------------------------------------------------------------------------
extern "C" {

const char* eid_last_error(void) { return eid::lastError().c_str(); }
int eid_version(void) { return 100; }
int eid_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int eid_scene_create(eid_scene** out, int device) {
  EID_TRY
  if (!out) raise(EID_ERR_INVALID, "eid_scene_create: out is null");
  if (device != EID_DEVICE_NONE) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) { cudaGetLastError(); raise(EID_ERR_CUDA, "no CUDA device is usable (libeidola has no CPU fallback)"); }
    if (device < 0 || device >= n) raise(EID_ERR_INVALID, "device %d out of range (0..%d)", device, n - 1);
  }
  eid_scene* s = new eid_scene();
  s->dev.device = device;
  *out = s;
  return EID_OK;
  EID_CATCH
}

static int finishLoad(eid_scene* s) {
  s->host.build();
  if (s->dev.device != EID_DEVICE_NONE) s->dev.upload(s->host);
  s->loaded = true;
  return EID_OK;
}

int eid_scene_load_gltf(eid_scene* s, const char* path) {
  EID_TRY
  if (!s || !path) raise(EID_ERR_INVALID, "eid_scene_load_gltf: null argument");
  s->loaded = false;
  importGltfFile(path, s->host.gltf, s->providedImages);
  return finishLoad(s);
  EID_CATCH
}

int eid_scene_provide_image(eid_scene* s, uint32_t imageIndex, const uint8_t* rgba8, uint32_t width, uint32_t height) {
  EID_TRY
  if (!s || !rgba8 || !width || !height) raise(EID_ERR_INVALID, "eid_scene_provide_image: bad argument");
  if (imageIndex >= 65536) raise(EID_ERR_INVALID, "image index %u too large", imageIndex);
  if (s->providedImages.size() <= imageIndex) s->providedImages.resize(imageIndex + 1);
  auto& im = s->providedImages[imageIndex];
  im.width = width; im.height = height;
  im.rgba8.assign(rgba8, rgba8 + 4 * (size_t)width * height);
  return EID_OK;
  EID_CATCH
}

int eid_scene_load_desc(eid_scene* s, const eid_scene_desc* desc) {
  EID_TRY
  if (!s || !desc) raise(EID_ERR_INVALID, "eid_scene_load_desc: null argument");
  s->loaded = false;
  s->host.gltf = HostGltf();
  s->host.gltf.fromDesc(*desc);
  return finishLoad(s);
  EID_CATCH
}

void eid_scene_destroy(eid_scene* s) {
  if (!s) return;
  s->dev.release();
  delete s;
}

int eid_scene_set_lookat(eid_scene* s, const float eye[3], const float center[3], const float up[3], float fovDeg) {
  EID_TRY
  if (!s || !eye || !center || !up) raise(EID_ERR_INVALID, "eid_scene_set_lookat: null argument");
  for (int i = 0; i < 3; ++i) { s->host.eye[i] = eye[i]; s->host.center[i] = center[i]; s->host.up[i] = up[i]; }
  s->host.fovDeg = fovDeg;
  return EID_OK;
  EID_CATCH
}

int eid_scene_update_camera(eid_scene* s, uint32_t w, uint32_t h) {
  EID_TRY
  if (!s || !w || !h) raise(EID_ERR_INVALID, "eid_scene_update_camera: bad argument");
  s->host.updateCamera(w, h);
  return EID_OK;
  EID_CATCH
}

int eid_scene_set_camera(eid_scene* s, const SceneCamera* cam) {
  EID_TRY
  if (!s || !cam) raise(EID_ERR_INVALID, "eid_scene_set_camera: null argument");
  s->host.camera = *cam;
  return EID_OK;
  EID_CATCH
}

int eid_scene_get_camera(eid_scene* s, SceneCamera* out) {
  EID_TRY
  if (!s || !out) raise(EID_ERR_INVALID, "eid_scene_get_camera: null argument");
  *out = s->host.camera;
  return EID_OK;
  EID_CATCH
}

int eid_scene_get_info(eid_scene* s, eid_scene_info* o) {
  EID_TRY
  if (!s || !o) raise(EID_ERR_INVALID, "eid_scene_get_info: null argument");
  if (!s->loaded) raise(EID_ERR_STATE, "scene not loaded");
  const SceneHost& H = s->host;
  memset(o, 0, sizeof(*o));
  o->primMeshCount = (uint32_t)H.gltf.primMeshes.size(); o->nodeCount = (uint32_t)H.gltf.nodes.size();
  o->materialCount = (uint32_t)H.materials.size();
  o->puncLightCount = H.lightInfo.puncLightSize; o->trigLightCount = H.lightInfo.trigLightSize;
  o->vertexCount = (uint32_t)(H.gltf.positions.size() / 3); o->indexCount = (uint32_t)H.gltf.indices.size();
  o->triangleInstances = H.triangleInstances;
  o->trigLightWeight = H.trigLightWeight; o->puncLightWeight = H.puncLightWeight;
  for (int i = 0; i < 3; ++i) { o->bboxMin[i] = H.gltf.bboxMin[i]; o->bboxMax[i] = H.gltf.bboxMax[i]; }
  return EID_OK;
  EID_CATCH
}

static const void* tableSource(eid_scene* s, int table, uint32_t index, size_t& bytes, bool& onDevice) {
  const SceneHost& H = s->host;
  onDevice = true;
  if (s->dev.device == EID_DEVICE_NONE) {   // host-only scene: serve the host copies (InstanceData carries no addresses)
    onDevice = false;
    static thread_local std::vector<InstanceData> hostInst;
    switch (table) {
      case EID_TABLE_MATERIALS: bytes = H.materials.size() * sizeof(GltfShadeMaterial); return H.materials.data();
      case EID_TABLE_PUNC_LIGHTS: bytes = H.puncLights.size() * sizeof(PuncLight); return H.puncLights.data();
      case EID_TABLE_TRIG_LIGHTS: bytes = H.trigLights.size() * sizeof(TrigLight); return H.trigLights.data();
      case EID_TABLE_LIGHT_INFO: bytes = sizeof(LightBufInfo); return &H.lightInfo;
      case EID_TABLE_INSTANCE_DATA:
        hostInst.assign(H.gltf.primMeshes.size(), InstanceData{});
        for (size_t p = 0; p < hostInst.size(); ++p) hostInst[p].materialIndex = H.gltf.primMeshes[p].materialIndex;
        bytes = hostInst.size() * sizeof(InstanceData); return hostInst.data();
      case EID_TABLE_VERTICES:
        if (index >= H.gltf.primMeshes.size()) return nullptr;
        bytes = (size_t)H.gltf.primMeshes[index].vertexCount * sizeof(VertexAttributes); return H.vertices.data() + H.vtxBase[index];
      case EID_TABLE_INDICES:
        if (index >= H.gltf.primMeshes.size()) return nullptr;
        bytes = (size_t)H.gltf.primMeshes[index].indexCount * 4; return H.indices.data() + H.idxBase[index];
      case EID_TABLE_CAMERA: bytes = sizeof(SceneCamera); return &H.camera;
      case EID_TABLE_TEXELS:
        if (index >= H.textures.size()) return nullptr;
        bytes = (size_t)H.textures[index].width * H.textures[index].height * 4; return H.texels.data() + H.textures[index].texelOffset;
      default: return nullptr;
    }
  }
  switch (table) {
    case EID_TABLE_MATERIALS: bytes = H.materials.size() * sizeof(GltfShadeMaterial); return s->dev.materials;
    case EID_TABLE_PUNC_LIGHTS: bytes = H.puncLights.size() * sizeof(PuncLight); return s->dev.puncLights;
    case EID_TABLE_TRIG_LIGHTS: bytes = H.trigLights.size() * sizeof(TrigLight); return s->dev.trigLights;
    case EID_TABLE_LIGHT_INFO: onDevice = false; bytes = sizeof(LightBufInfo); return &H.lightInfo;
    case EID_TABLE_INSTANCE_DATA: bytes = H.gltf.primMeshes.size() * sizeof(InstanceData); return s->dev.geoInfo;
    case EID_TABLE_VERTICES:
      if (index >= H.gltf.primMeshes.size()) return nullptr;
      bytes = (size_t)H.gltf.primMeshes[index].vertexCount * sizeof(VertexAttributes); return s->dev.vertices + H.vtxBase[index];
    case EID_TABLE_INDICES:
      if (index >= H.gltf.primMeshes.size()) return nullptr;
      bytes = (size_t)H.gltf.primMeshes[index].indexCount * 4; return s->dev.indices + H.idxBase[index];
    case EID_TABLE_CAMERA: onDevice = false; bytes = sizeof(SceneCamera); return &H.camera;
    case EID_TABLE_TEXELS:
      if (index >= H.textures.size()) return nullptr;
      onDevice = false; bytes = (size_t)H.textures[index].width * H.textures[index].height * 4; return H.texels.data() + H.textures[index].texelOffset;
    default: return nullptr;
  }
}

int64_t eid_scene_table_bytes(eid_scene* s, int table, uint32_t index) {
  if (!s || !s->loaded) return -1;
  size_t b = 0; bool dev;
  return tableSource(s, table, index, b, dev) ? (int64_t)b : -1;
}

int eid_scene_read_table(eid_scene* s, int table, uint32_t index, void* dst, size_t bytes) {
  EID_TRY
  if (!s || !dst) raise(EID_ERR_INVALID, "eid_scene_read_table: null argument");
  if (!s->loaded) raise(EID_ERR_STATE, "scene not loaded");
  size_t b = 0; bool dev;
  const void* src = tableSource(s, table, index, b, dev);
  if (!src) raise(EID_ERR_INVALID, "no such table %d[%u]", table, index);
  if (bytes > b) raise(EID_ERR_INVALID, "read of %zu bytes from a %zu-byte table", bytes, b);
  if (dev) { CUDA_CHECK(cudaSetDevice(s->dev.device)); CUDA_CHECK(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost)); }
  else memcpy(dst, src, bytes);
  return EID_OK;
  EID_CATCH
}

int eid_accel_build_ex(eid_scene* s, int mode, eid_accel** out) {
  EID_TRY
  if (!s || !out) raise(EID_ERR_INVALID, "eid_accel_build: null argument");
  const int build = mode & (EID_ACCEL_FAST_TRACE | EID_ACCEL_FAST_BUILD);
  mode &= ~(EID_ACCEL_FAST_TRACE | EID_ACCEL_FAST_BUILD);
  if (mode < EID_ACCEL_AUTO || mode > EID_ACCEL_TWO_LEVEL) raise(EID_ERR_INVALID, "eid_accel_build_ex: mode must be EID_ACCEL_AUTO, _FLAT or _TWO_LEVEL (| EID_ACCEL_FAST_TRACE or _FAST_BUILD)");
  if (build == (EID_ACCEL_FAST_TRACE | EID_ACCEL_FAST_BUILD)) raise(EID_ERR_INVALID, "eid_accel_build_ex: EID_ACCEL_FAST_TRACE and EID_ACCEL_FAST_BUILD exclude each other");
  // neither flag: the reference's choice (PREFER_FAST_TRACE, accelstruct.cpp:125-126,161) unless EIDOLA_ACCEL_BUILD=lbvh|sah says otherwise (A/B runs)
  bool sah = build != EID_ACCEL_FAST_BUILD;
  if (!build) { const char* e = getenv("EIDOLA_ACCEL_BUILD"); if (e && !strcmp(e, "lbvh")) sah = false; }
  if (EID_BVH_WIDTH != 4) sah = false;
  if (!s->loaded) raise(EID_ERR_STATE, "eid_accel_build before a scene was loaded");
  if (s->dev.device == EID_DEVICE_NONE) raise(EID_ERR_CUDA, "eid_accel_build on a host-only scene: the BVH build and every kernel need a CUDA device (no CPU fallback)");
  bool two = mode == EID_ACCEL_TWO_LEVEL;
  if (mode == EID_ACCEL_AUTO) {
    // instancing pays once the flattened list would be at least twice the unique triangles (the flat tree traces faster: one level, FFMA2 walk)
    const SceneHost& H = s->host;
    std::vector<uint32_t> meshTris(H.gltf.primMeshes.size(), 0);
    for (const InstanceXform& X : H.instances) meshTris[X.primMesh] = X.triangleCount;
    uint64_t unique = 0;
    for (uint32_t t : meshTris) unique += t;
    two = unique > 0 && H.triangleInstances >= 2 * unique;
  }
  eid_accel* a = new eid_accel();
  try { if (two) buildAccelTwoLevel(s, a, sah); else buildAccel(s, a, sah); }
  catch (...) { cudaFree(a->nodes); cudaFree(a->tris); cudaFree(a->tlasNodes); cudaFree(a->tlasPrims); delete a; throw; }
  a->sahBuild = sah;
  *out = a;
  return EID_OK;
  EID_CATCH
}

int eid_accel_build(eid_scene* s, eid_accel** out) { return eid_accel_build_ex(s, EID_ACCEL_AUTO, out); }

void eid_accel_destroy(eid_accel* a) {
  if (!a) return;
  if (a->nodeTex) cudaDestroyTextureObject(a->nodeTex);
  if (a->triTex) cudaDestroyTextureObject(a->triTex);
  cudaFree(a->nodes); cudaFree(a->tris); cudaFree(a->tlasNodes); cudaFree(a->tlasPrims);
  delete a;
}

int eid_accel_get_info(eid_accel* a, eid_accel_info* o) {
  EID_TRY
  if (!a || !o) raise(EID_ERR_INVALID, "eid_accel_get_info: null argument");
  o->triangleCount = a->triCount; o->nodeCount = a->nodeCount; o->maxDepth = a->maxDepth;
  o->nodeBytes = (uint64_t)std::max<uint32_t>(1u, a->nodeAlloc) * EID_NODE_BYTES; o->triBytes = (uint64_t)a->triCount * 48;
  o->buildMs = a->buildMs;
  o->fastTrace = a->sahBuild ? 1 : 0;
  o->twoLevel = a->twoLevel ? 1 : 0; o->blasCount = a->blasCount; o->tlasNodeCount = a->tlasNodeCount; o->instanceCount = a->twoLevel ? a->tlasPrimCount : (uint32_t)a->scene->host.instances.size();
  if (a->twoLevel) o->nodeBytes += (uint64_t)std::max<uint32_t>(1u, a->tlasNodeCount) * 128 + (uint64_t)a->tlasPrimCount * 48;
  return EID_OK;
  EID_CATCH
}

int eid_accel_trace(eid_accel* a, const float* rays, uint32_t n, int any_hit, eid_hit* hits) {
  EID_TRY
  if (!a || !rays || !hits) raise(EID_ERR_INVALID, "eid_accel_trace: null argument");
  if (n == 0) return EID_OK;
  CUDA_CHECK(cudaSetDevice(a->scene->dev.device));
  float* dr = nullptr; eid_hit* dh = nullptr;
  CUDA_CHECK(cudaMalloc(&dr, (size_t)n * 32));
  if (cudaMalloc(&dh, (size_t)n * sizeof(eid_hit)) != cudaSuccess) { cudaFree(dr); raise(EID_ERR_CUDA, "cudaMalloc failed"); }
  cudaError_t e = cudaMemcpy(dr, rays, (size_t)n * 32, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    k_trace_batch<<<(n + 127) / 128, 128>>>(a->view(), a->scene->dev.instances, dr, n, any_hit, dh);
    e = cudaMemcpy(hits, dh, (size_t)n * sizeof(eid_hit), cudaMemcpyDeviceToHost);
  }
  cudaFree(dr); cudaFree(dh);
  if (e != cudaSuccess) raise(EID_ERR_CUDA, "eid_accel_trace: %s", cudaGetErrorString(e));
  return EID_OK;
  EID_CATCH
}

}  // extern "C"

// tools/lane_model.cpp — CPU model of the lane occupancy of direct_stage's walk (not part of the product); shares the tree code of bvh_quality.cpp:
//   python tools/bvh_quality_dump.py [quads]   (writes tris_<quads>.bin: the C3 scene's world-space triangles)
//   g++ -O2 -std=c++17 -pthread -o /tmp/lanem tools/lane_model.cpp cis-565-final-vr-raytracer_b200/csrc/sah_host.cpp && /tmp/bvhq tris_707.bin 2
// BVH quality experiment: LBVH (Morton, per-axis normalised) vs binned SAH, both collapsed to BVH4 with the product's greedy rule,
// node visits / triangle tests per ray for primary, shadow and diffuse-bounce rays of the C3 view.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>
#include <string>
#include <chrono>
#include "../cis-565-final-vr-raytracer_b200/csrc/sah_host.h"
using namespace std;
struct V { float x, y, z; };
static V operator-(V a, V b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static V operator+(V a, V b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static V operator*(V a, float s) { return {a.x * s, a.y * s, a.z * s}; }
static V cross(V a, V b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
static float dot(V a, V b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static V norm(V a) { return a * (1.0f / sqrtf(dot(a, a))); }
struct Box { float lo[3], hi[3]; };
static Box emptyBox() { return {{3e38f, 3e38f, 3e38f}, {-3e38f, -3e38f, -3e38f}}; }
static void grow(Box& b, const Box& o) { for (int k = 0; k < 3; ++k) { b.lo[k] = min(b.lo[k], o.lo[k]); b.hi[k] = max(b.hi[k], o.hi[k]); } }
static float area(const Box& b) { float ex = b.hi[0] - b.lo[0], ey = b.hi[1] - b.lo[1], ez = b.hi[2] - b.lo[2]; return ex * ey + ey * ez + ez * ex; }

struct Tri { V v0, v1, v2; };
vector<Tri> tris; vector<Box> tbox;
int LEAF_MAX = 2;

// binary tree over a permutation
struct BNode { Box box; int left, right; int first, count; };   // children: index into nodes; leaf when count <= LEAF_MAX (left = -1)
struct BTree { vector<BNode> nodes; vector<int> order; };

static uint64_t expand21(uint64_t v) {
  v &= 0x1fffffull; v = (v | v << 32) & 0x1f00000000ffffull; v = (v | v << 16) & 0x1f0000ff0000ffull; v = (v | v << 8) & 0x100f00f00f00f00full;
  v = (v | v << 4) & 0x10c30c30c30c30c3ull; v = (v | v << 2) & 0x1249249249249249ull; return v;
}
static int buildRadix(BTree& T, const vector<uint64_t>& keys, int first, int count) {
  int id = (int)T.nodes.size(); T.nodes.push_back(BNode());
  Box b = emptyBox(); for (int i = first; i < first + count; ++i) grow(b, tbox[T.order[i]]);
  T.nodes[id].box = b; T.nodes[id].first = first; T.nodes[id].count = count; T.nodes[id].left = T.nodes[id].right = -1;
  if (count <= LEAF_MAX) return id;
  uint64_t a = keys[first], c = keys[first + count - 1];
  int split;
  if (a == c) split = first + count / 2;
  else {
    int bit = 63 - __builtin_clzll(a ^ c);
    // first index with that bit set
    int lo = first, hi = first + count - 1;
    while (lo < hi) { int m = (lo + hi) / 2; if ((keys[m] >> bit) & 1) hi = m; else lo = m + 1; }
    split = lo;
  }
  int l = buildRadix(T, keys, first, split - first), r = buildRadix(T, keys, split, first + count - split);
  T.nodes[id].left = l; T.nodes[id].right = r;
  return id;
}
static BTree buildLBVH(int mode) {   // mode 0: per-axis normalised (product), 1: normalised by the largest extent
  int n = (int)tris.size();
  Box sb = emptyBox(); for (auto& b : tbox) grow(sb, b);
  float emax = max(max(sb.hi[0] - sb.lo[0], sb.hi[1] - sb.lo[1]), sb.hi[2] - sb.lo[2]);
  vector<pair<uint64_t, int>> kv(n);
  for (int t = 0; t < n; ++t) {
    uint64_t key = 0;
    for (int k = 0; k < 3; ++k) {
      float e = mode == 0 ? sb.hi[k] - sb.lo[k] : emax;
      float c = (0.5f * (tbox[t].lo[k] + tbox[t].hi[k]) - sb.lo[k]) / e; c = min(max(c, 0.f), 1.f);
      uint64_t q = (uint64_t)min(c * 2097152.0f, 2097151.0f);
      key |= expand21(q) << (2 - k);
    }
    kv[t] = {key, t};
  }
  sort(kv.begin(), kv.end());
  BTree T; T.order.resize(n); vector<uint64_t> keys(n);
  for (int i = 0; i < n; ++i) { T.order[i] = kv[i].second; keys[i] = kv[i].first; }
  T.nodes.reserve(2 * n);
  buildRadix(T, keys, 0, n);
  return T;
}
static int buildSAH(BTree& T, int first, int count) {
  int id = (int)T.nodes.size(); T.nodes.push_back(BNode());
  Box b = emptyBox(), cb = emptyBox();
  for (int i = first; i < first + count; ++i) {
    const Box& tb = tbox[T.order[i]]; grow(b, tb);
    for (int k = 0; k < 3; ++k) { float c = 0.5f * (tb.lo[k] + tb.hi[k]); cb.lo[k] = min(cb.lo[k], c); cb.hi[k] = max(cb.hi[k], c); }
  }
  T.nodes[id].box = b; T.nodes[id].first = first; T.nodes[id].count = count; T.nodes[id].left = T.nodes[id].right = -1;
  if (count <= LEAF_MAX) return id;
  const int NB = 16;
  float bestCost = 3e38f; int bestAxis = -1, bestBin = -1;
  for (int k = 0; k < 3; ++k) {
    float e = cb.hi[k] - cb.lo[k]; if (!(e > 0.f)) continue;
    Box bb[NB]; int bc[NB]; for (int i = 0; i < NB; ++i) { bb[i] = emptyBox(); bc[i] = 0; }
    float sc = NB / e;
    for (int i = first; i < first + count; ++i) {
      const Box& tb = tbox[T.order[i]]; float c = 0.5f * (tb.lo[k] + tb.hi[k]);
      int bi = min(NB - 1, max(0, (int)((c - cb.lo[k]) * sc))); grow(bb[bi], tb); bc[bi]++;
    }
    float ra[NB]; int rc[NB]; Box acc = emptyBox(); int cnt = 0;
    for (int i = NB - 1; i > 0; --i) { grow(acc, bb[i]); cnt += bc[i]; ra[i] = area(acc); rc[i] = cnt; }
    acc = emptyBox(); cnt = 0;
    for (int i = 0; i < NB - 1; ++i) {
      grow(acc, bb[i]); cnt += bc[i];
      if (cnt == 0 || rc[i + 1] == 0) continue;
      float cost = area(acc) * cnt + ra[i + 1] * rc[i + 1];
      if (cost < bestCost) { bestCost = cost; bestAxis = k; bestBin = i; }
    }
  }
  int mid;
  if (bestAxis < 0) mid = first + count / 2;
  else {
    float e = cb.hi[bestAxis] - cb.lo[bestAxis], sc = NB / e;
    auto it = partition(T.order.begin() + first, T.order.begin() + first + count, [&](int t) {
      float c = 0.5f * (tbox[t].lo[bestAxis] + tbox[t].hi[bestAxis]);
      return min(NB - 1, max(0, (int)((c - cb.lo[bestAxis]) * sc))) <= bestBin; });
    mid = (int)(it - T.order.begin());
    if (mid == first || mid == first + count) mid = first + count / 2;
  }
  int l = buildSAH(T, first, mid - first), r = buildSAH(T, mid, first + count - mid);
  T.nodes[id].left = l; T.nodes[id].right = r;
  return id;
}

// BVH4
struct WNode { Box b[4]; int ref[4]; };   // ref >= 0 inner, < 0: ~((first << 3) | count), empty = ~0
struct W4 { vector<WNode> nodes; vector<int> order; int root; };
static W4 collapse(const BTree& T) {
  W4 W; W.order = T.order;
  vector<pair<int, int>> q = {{0, 0}}; W.nodes.push_back(WNode());
  for (size_t qi = 0; qi < q.size(); ++qi) {
    int b = q[qi].first, w = q[qi].second;
    int c[4] = {T.nodes[b].left, T.nodes[b].right, 0, 0}; int n = 2;
    while (n < 4) {
      int best = -1; float bestA = -1.f;
      for (int j = 0; j < n; ++j) if (T.nodes[c[j]].count > LEAF_MAX) { float a = area(T.nodes[c[j]].box); if (a > bestA) { bestA = a; best = j; } }
      if (best < 0) break;
      int r = c[best]; c[best] = T.nodes[r].left; c[n++] = T.nodes[r].right;
    }
    WNode wn;
    for (int j = 0; j < 4; ++j) {
      if (j >= n) { for (int k = 0; k < 3; ++k) wn.b[j].lo[k] = wn.b[j].hi[k] = 3e38f; wn.ref[j] = ~0; continue; }
      const BNode& cn = T.nodes[c[j]]; wn.b[j] = cn.box;
      if (cn.count <= LEAF_MAX) wn.ref[j] = ~((cn.first << 3) | cn.count);
      else { wn.ref[j] = (int)W.nodes.size(); W.nodes.push_back(WNode()); q.push_back({c[j], wn.ref[j]}); }
    }
    W.nodes[w] = wn;
  }
  W.root = 0;
  return W;
}

static int convRec(const eid::BinaryTreeHost& H, BTree& T, int ref) {
  int id = (int)T.nodes.size(); T.nodes.push_back(BNode());
  if (ref < 0) { int p = ~ref; T.nodes[id].box = tbox[T.order[p]]; T.nodes[id].first = p; T.nodes[id].count = 1; T.nodes[id].left = T.nodes[id].right = -1; return id; }
  int l = convRec(H, T, H.left[ref]), r = convRec(H, T, H.right[ref]);
  Box b = T.nodes[l].box; grow(b, T.nodes[r].box);
  T.nodes[id].box = b; T.nodes[id].first = H.rangeFirst[ref]; T.nodes[id].count = H.rangeLast[ref] - H.rangeFirst[ref] + 1; T.nodes[id].left = l; T.nodes[id].right = r;
  // invariants
  if (T.nodes[l].first != T.nodes[id].first || T.nodes[r].first != T.nodes[l].first + T.nodes[l].count || T.nodes[l].count + T.nodes[r].count != T.nodes[id].count) { printf("RANGE BROKEN at %d\n", ref); exit(1); }
  if (H.left[ref] >= 0 && H.parentInner[H.left[ref]] != ref) { printf("PARENT BROKEN\n"); exit(1); }
  if (H.right[ref] >= 0 && H.parentInner[H.right[ref]] != ref) { printf("PARENT BROKEN\n"); exit(1); }
  if (H.left[ref] < 0 && H.parentLeaf[~H.left[ref]] != ref) { printf("PARENTLEAF BROKEN\n"); exit(1); }
  if (H.right[ref] < 0 && H.parentLeaf[~H.right[ref]] != ref) { printf("PARENTLEAF BROKEN\n"); exit(1); }
  return id;
}
static BTree fromHost(int threads) {
  int n = (int)tris.size();
  vector<float> lo(3 * (size_t)n), hi(3 * (size_t)n);
  for (int t = 0; t < n; ++t) for (int k = 0; k < 3; ++k) { lo[3 * (size_t)t + k] = tbox[t].lo[k]; hi[3 * (size_t)t + k] = tbox[t].hi[k]; }
  eid::BinaryTreeHost H;
  auto t0 = chrono::steady_clock::now();
  eid::buildSahTree(n, lo.data(), hi.data(), H, threads);
  double ms = chrono::duration<double, milli>(chrono::steady_clock::now() - t0).count();
  printf("buildSahTree(%d threads): %.1f ms\n", threads, ms);
  vector<char> seen(n, 0); for (int i = 0; i < n; ++i) { if (seen[H.order[i]]) { printf("PERM BROKEN\n"); exit(1); } seen[H.order[i]] = 1; }
  if (H.parentInner[0] != -1) { printf("ROOT BROKEN\n"); exit(1); }
  BTree T; T.order.assign(H.order.begin(), H.order.end()); T.nodes.reserve(2 * (size_t)n);
  // convRec recursion is depth-bounded by tree height; swap root to index 0 is natural here
  convRec(H, T, 0);
  if ((int)T.nodes.size() != 2 * n - 1) { printf("NODE COUNT BROKEN %zu\n", T.nodes.size()); exit(1); }
  return T;
}
struct Ray { V o, d; float tmax; bool any; };
struct Stat { double nodes = 0, tris = 0, rays = 0; std::string* ev = nullptr; };
static bool triTest(const Tri& T, V o, V d, float tmax, float& t) {
  V e1 = T.v1 - T.v0, e2 = T.v2 - T.v0, p = cross(d, e2); float det = dot(e1, p);
  if (!(det > 0.f)) return false;
  float inv = 1.f / det; V tv = o - T.v0; float u = dot(tv, p) * inv; if (u < 0 || u > 1) return false;
  V qv = cross(tv, e1); float v = dot(d, qv) * inv; if (v < 0 || u + v > 1) return false;
  t = dot(e2, qv) * inv; return t > 0 && t < tmax;
}
static bool trace(const W4& W, const Ray& r, Stat& S, float& tHit, int& triHit) {
  float ix = 1.f / r.d.x, iy = 1.f / r.d.y, iz = 1.f / r.d.z;
  float tbest = r.tmax; triHit = -1;
  int stack[256], sp = 0, cur = W.root;
  const int DONE = (int)0x80000000;
  S.rays++;
  for (;;) {
    while (cur >= 0) {
      S.nodes++; if (S.ev) S.ev->push_back('N');
      const WNode& n = W.nodes[cur];
      float e[4];
      for (int j = 0; j < 4; ++j) {
        float x0 = (n.b[j].lo[0] - r.o.x) * ix, x1 = (n.b[j].hi[0] - r.o.x) * ix;
        float y0 = (n.b[j].lo[1] - r.o.y) * iy, y1 = (n.b[j].hi[1] - r.o.y) * iy;
        float z0 = (n.b[j].lo[2] - r.o.z) * iz, z1 = (n.b[j].hi[2] - r.o.z) * iz;
        float tn = max(max(min(x0, x1), min(y0, y1)), max(min(z0, z1), 0.f));
        float tf = min(min(max(x0, x1), max(y0, y1)), min(max(z0, z1), tbest));
        e[j] = (tn <= tf * 1.0000004f && n.ref[j] != ~0) ? tn : INFINITY;
      }
      if (r.any) {
        int next = 0; bool have = false;
        for (int j = 0; j < 4; ++j) if (e[j] < INFINITY) { if (have) stack[sp++] = n.ref[j]; else { next = n.ref[j]; have = true; } }
        cur = have ? next : (sp ? stack[--sp] : DONE);
      } else {
        int idx[4] = {0, 1, 2, 3};
        sort(idx, idx + 4, [&](int a, int b) { return e[a] < e[b]; });
        if (e[idx[0]] < INFINITY) {
          for (int j = 3; j >= 1; --j) if (e[idx[j]] < INFINITY) stack[sp++] = n.ref[idx[j]];
          cur = n.ref[idx[0]];
        } else cur = sp ? stack[--sp] : DONE;
      }
    }
    if (cur == DONE) break;
    uint32_t ref = ~(uint32_t)cur; uint32_t first = ref >> 3, count = ref & 7;
    bool fin = false;
    if (S.ev) S.ev->push_back((char)('0' + count));
    for (uint32_t k = 0; k < count; ++k) {
      S.tris++;
      float t;
      if (triTest(tris[W.order[first + k]], r.o, r.d, tbest, t)) { tbest = t; triHit = W.order[first + k]; if (r.any) { fin = true; break; } }
    }
    if (fin) break;
    cur = sp ? stack[--sp] : DONE;
  }
  tHit = tbest;
  return triHit >= 0;
}
static double sahCost(const W4& W) {
  double c = 0; double ra = 0;
  // root area
  Box rb = emptyBox(); for (int j = 0; j < 4; ++j) if (W.nodes[0].ref[j] != ~0) grow(rb, W.nodes[0].b[j]); ra = area(rb);
  for (auto& n : W.nodes) for (int j = 0; j < 4; ++j) if (n.ref[j] != ~0) { if (n.ref[j] >= 0) c += area(n.b[j]) / ra; else c += 0.5 * area(n.b[j]) / ra * ((~(uint32_t)n.ref[j]) & 7); }
  return c;
}

// ---- warp-level model of the one-ray-per-thread walk (k_direct_stage): 32 rays of an 8 x 4 pixel tile per warp ---------------------------
// events of a ray: 'N' = inner-node visit, '1'..'7' = leaf visit with that many triangles.  Costs in issue slots: node visit 100, triangle 90.
struct WarpCost { double slots = 0, laneSlots = 0, nodeSlots = 0, nodeLaneSlots = 0; };
static void simulate(const vector<string>& ev, int nodeBound, WarpCost& C) {   // nodeBound <= 0: while-while (node phase until every lane holds a leaf)
  const double CN = 100, CT = 90;
  vector<size_t> pos(ev.size(), 0);
  for (;;) {
    bool any = false;
    for (size_t l = 0; l < ev.size(); ++l) any = any || pos[l] < ev[l].size();
    if (!any) break;
    // node phase
    for (int step = 0; nodeBound <= 0 || step < nodeBound; ++step) {
      int act = 0;
      for (size_t l = 0; l < ev.size(); ++l) if (pos[l] < ev[l].size() && ev[l][pos[l]] == 'N') { ++pos[l]; ++act; }
      if (!act) break;
      C.slots += CN; C.laneSlots += CN * act; C.nodeSlots += CN; C.nodeLaneSlots += CN * act;
    }
    // leaf phase: every lane at a leaf tests its triangles; the warp pays for the largest leaf
    int mx = 0, sum = 0;
    for (size_t l = 0; l < ev.size(); ++l) if (pos[l] < ev[l].size() && ev[l][pos[l]] != 'N') { int c = ev[l][pos[l]] - '0'; mx = max(mx, c); sum += c; ++pos[l]; }
    C.slots += CT * mx; C.laneSlots += CT * sum;
  }
}
int main(int argc, char** argv) {
  const char* path = argc > 1 ? argv[1] : "tris_707.bin";
  FILE* f = fopen(path, "rb"); fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET);
  int n = (int)(sz / 36); tris.resize(n); if (fread(tris.data(), 36, n, f) != (size_t)n) return 1; fclose(f);
  tbox.resize(n);
  for (int t = 0; t < n; ++t) {
    const V* v = &tris[t].v0; Box b = emptyBox();
    for (int j = 0; j < 3; ++j) { float p[3] = {v[j].x, v[j].y, v[j].z}; for (int k = 0; k < 3; ++k) { b.lo[k] = min(b.lo[k], p[k] - 1e-4f); b.hi[k] = max(b.hi[k], p[k] + 1e-4f); } }
    tbox[t] = b;
  }
  W4 W = collapse(fromHost(0));
  V eye = {0.f, 5.f, -17.5f}, ctr = {0.f, 1.5f, 4.f}, up = {0, 1, 0};
  V fw = norm(ctr - eye), rt = norm(cross(fw, up)), upv = cross(rt, fw);
  const int Wd = 1920, Hd = 1080; float th = tanf(0.5f * 60.f * 3.14159265f / 180.f), asp = (float)Wd / Hd;
  vector<int> lights; for (int t = n - 1000; t < n; ++t) lights.push_back(t);
  mt19937 rng(1); uniform_real_distribution<float> U(0, 1);
  const int bounds[] = {0, 1, 2, 4, 8};
  WarpCost prim[5], shad[5]; double primEvents = 0, shadEvents = 0, rays = 0, srays = 0;
  for (int ty = 0; ty < Hd / 4; ty += 7) for (int tx = 0; tx < Wd / 8; tx += 5) {          // a sample of 8 x 4 pixel tiles = warps
    vector<string> pe(32), se(32);
    for (int l = 0; l < 32; ++l) {
      const int x = tx * 8 + (l & 7), y = ty * 4 + (l >> 3);
      float u = ((x + 0.5f) / Wd * 2 - 1) * th * asp, v = (1 - (y + 0.5f) / Hd * 2) * th;
      Ray r{eye, norm(fw + rt * u + upv * v), 1e30f, false};
      Stat s; s.ev = &pe[l]; float t; int hit;
      rays++;
      if (!trace(W, r, s, t, hit)) continue;
      const Tri& T = tris[hit]; V nrm = norm(cross(T.v1 - T.v0, T.v2 - T.v0)); if (dot(nrm, r.d) > 0) nrm = nrm * -1.f;
      V p = r.o + r.d * t + nrm * 1e-3f;
      const Tri& L = tris[lights[(int)(U(rng) * 999.99f)]]; V lc = (L.v0 + L.v1 + L.v2) * (1.f / 3.f);
      V dd = lc - p; float dist = sqrtf(dot(dd, dd));
      Ray sr{p, dd * (1.f / dist), dist - 1e-3f, true};
      Stat s2; s2.ev = &se[l]; srays++;
      trace(W, sr, s2, t, hit);
    }
    for (int b = 0; b < 5; ++b) { simulate(pe, bounds[b], prim[b]); simulate(se, bounds[b], shad[b]); }
    for (auto& e : pe) for (char c : e) primEvents += c == 'N' ? 100 : 90 * (c - '0');
    for (auto& e : se) for (char c : e) shadEvents += c == 'N' ? 100 : 90 * (c - '0');
  }
  printf("rays: %.0f primary, %.0f shadow (warps of one 8 x 4 pixel tile)\n", rays, srays);
  printf("perfectly packed (every slot 32 lanes): primary %.0f slots / ray, shadow %.0f\n", primEvents / 32 / rays * 32 / 32, shadEvents / 32 / srays * 32 / 32);
  for (int b = 0; b < 5; ++b) {
    printf("%-28s primary: %6.0f slots / ray, %4.1f lanes (node phase %4.1f) | shadow: %6.0f slots / ray, %4.1f lanes (node phase %4.1f)\n",
           bounds[b] ? (string("node phase <= ") + to_string(bounds[b]) + " visits").c_str() : "while-while (product)",
           prim[b].slots / rays, prim[b].laneSlots / prim[b].slots, prim[b].nodeLaneSlots / prim[b].nodeSlots,
           shad[b].slots / srays, shad[b].laneSlots / shad[b].slots, shad[b].nodeLaneSlots / shad[b].nodeSlots);
  }
  return 0;
}

// knn.cuh — exact k-nearest-neighbour search on the Morton-sorted leaf/box tree (device side).
// Replaces pcl::KdTreeFLANN::nearestKSearch at impl/gicp.hpp:69,196, impl/semantic_icp.hpp:68,
// impl/em_icp.hpp:60,221,296 and impl/semantic_point_cloud.hpp:41.
//
// Contract (bit-exact vs the oracle): the k targets minimising (d2, original index) lexicographically with
//   d2 = fl(fl(fl(dx*dx)+fl(dy*dy))+fl(dz*dz)),  dx = fl(q.x - p.x)            (FLANN L2_Simple<float>, no FMA)
// Pruning uses box lower bounds evaluated with the same rounded operation sequence; rounding is monotone, so a
// box is skipped only if every point in it is strictly farther than the current k-th best.
//
// Execution model: one warp owns 32 Morton-consecutive queries (one leaf of the query cloud).  The warp walks the
// target tree in a fixed left-to-right order with WARP-UNIFORM control flow: a node is entered when any lane still
// needs it; a surviving leaf is staged into shared memory with one coalesced 512-byte load and every lane scans
// the 32 candidates from shared memory (broadcast reads).  A seed leaf near the queries is scanned first so that
// the fixed-order walk starts with a tight bound.
#pragma once
#include "common.cuh"

namespace sicp {

constexpr int kMortonBits = 19;                       // per axis; 57-bit code + 7 bits of class rank in the sort key
constexpr uint64_t kMortonMask = (1ull << 57) - 1;
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ uint64_t spread3(uint32_t v) {  // 21 -> 63 bits, two zero bits between
  uint64_t x = v & 0x1fffff;
  x = (x | x << 32) & 0x1f00000000ffffull;
  x = (x | x << 16) & 0x1f0000ff0000ffull;
  x = (x | x << 8) & 0x100f00f00f00f00full;
  x = (x | x << 4) & 0x10c30c30c30c30c3ull;
  x = (x | x << 2) & 0x1249249249249249ull;
  return x;
}
__device__ __forceinline__ uint64_t morton57(float x, float y, float z, const float* lo, float inv_cell) {
  const float mx = (float)((1u << kMortonBits) - 1);
  const float fx = fminf(fmaxf((x - lo[0]) * inv_cell, 0.f), mx);
  const float fy = fminf(fmaxf((y - lo[1]) * inv_cell, 0.f), mx);
  const float fz = fminf(fmaxf((z - lo[2]) * inv_cell, 0.f), mx);
  return spread3((uint32_t)fx) | (spread3((uint32_t)fy) << 1) | (spread3((uint32_t)fz) << 2);
}

__device__ __forceinline__ float dist2_rn(float qx, float qy, float qz, const float4& p) {
  const float dx = __fsub_rn(qx, p.x), dy = __fsub_rn(qy, p.y), dz = __fsub_rn(qz, p.z);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}
__device__ __forceinline__ float box_lb_rn(float qx, float qy, float qz, const float4& lo, const float4& hi) {
  const float dx = fmaxf(fmaxf(__fsub_rn(lo.x, qx), __fsub_rn(qx, hi.x)), 0.f);
  const float dy = fmaxf(fmaxf(__fsub_rn(lo.y, qy), __fsub_rn(qy, hi.y)), 0.f);
  const float dz = fmaxf(fmaxf(__fsub_rn(lo.z, qz), __fsub_rn(qz, hi.z)), 0.f);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// Sorted (ascending) list of the K best (d2, slot) in registers; ties on d2 are broken by the ORIGINAL index, which
// is fetched from pts[slot].w only when two distances are exactly equal.
template <int K>
struct TopK {
  float d[K];
  int s[K];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int i = 0; i < K; i++) { d[i] = INFINITY; s[i] = -1; }
  }
  __device__ __forceinline__ float worst() const { return d[K - 1]; }
  static __device__ __forceinline__ bool before(float dc, int oc, float de, int se, const float4* __restrict__ pts) {
    if (dc < de) return true;
    if (dc == de) return se >= 0 && oc < __float_as_int(pts[se].w);
    return false;
  }
  __device__ __forceinline__ void consider(float dc, int slot, int orig, const float4* __restrict__ pts) {
    if (!(dc <= d[K - 1])) return;  // also rejects NaN (padding slots)
    if (!before(dc, orig, d[K - 1], s[K - 1], pts)) return;
    bool placed = false;
#pragma unroll
    for (int i = K - 1; i >= 1; --i) {
      const bool mv = !placed && before(dc, orig, d[i - 1], s[i - 1], pts);
      if (mv) { d[i] = d[i - 1]; s[i] = s[i - 1]; }
      else if (!placed) { d[i] = dc; s[i] = slot; placed = true; }
    }
    if (!placed) { d[0] = dc; s[0] = slot; }
  }
};

// All 32 lanes scan one target leaf.  wbuf: 32 float4 of shared memory private to the warp.
template <int K>
__device__ __forceinline__ void scan_leaf(const float4* __restrict__ pts, int slot0, float qx, float qy, float qz, bool valid,
                                          TopK<K>& L, float4* wbuf) {
  const int lane = threadIdx.x & 31;
  const float4 mine = __ldg(&pts[slot0 + lane]);
  __syncwarp();
  wbuf[lane] = mine;
  __syncwarp();
  if (valid) {
#pragma unroll 8
    for (int j = 0; j < kLeaf; j++) {
      const float4 p = wbuf[j];
      L.consider(dist2_rn(qx, qy, qz, p), slot0 + j, __float_as_int(p.w), pts);
    }
  }
}

// Fixed-order walk of segment `sg` of the target; leaves [skip_lo, skip_hi] (segment-relative) were already scanned.
template <int K>
__device__ __forceinline__ void tree_walk(const CloudView& tv, const Segment& sg, float qx, float qy, float qz, bool valid, TopK<K>& L,
                                          int skip_lo, int skip_hi, float4* wbuf) {
  const int top = sg.nlevels - 1;
  int level = top, idx = 0;
  if (sg.nleaf == 0) return;
  for (;;) {
    const int ni = sg.node_off[level] + idx;
    const float4 lo = __ldg(&tv.node_lo[ni]), hi = __ldg(&tv.node_hi[ni]);
    const float lb = box_lb_rn(qx, qy, qz, lo, hi);
    const bool need = valid && !(lb > L.worst());
    bool descend = false;
    if (__any_sync(kFull, need)) {
      if (level == 0) {
        if (idx < skip_lo || idx > skip_hi) scan_leaf<K>(tv.pts, sg.p0 + idx * kLeaf, qx, qy, qz, valid, L, wbuf);
      } else {
        descend = true;
      }
    }
    if (descend) {
      level--;
      idx *= kArity;
      continue;
    }
    // advance to the next node in pre-order
    bool done = false;
    for (;;) {
      idx++;
      if (level == top) { done = idx >= sg.node_cnt[top]; break; }
      if ((idx % kArity) != 0 && idx < sg.node_cnt[level]) break;
      idx = (idx - 1) / kArity;  // parent; the loop increments to its next sibling
      level++;
    }
    if (done) break;
  }
}

// Leaf of segment `sg` whose first Morton code is the last one <= the query's code (any leaf is a valid seed).
__device__ __forceinline__ int seed_leaf(const CloudView& tv, const Segment& sg, float qx, float qy, float qz) {
  const uint64_t code = morton57(qx, qy, qz, sg.lo, sg.inv_cell);
  int lo = 0, hi = sg.nleaf - 1;  // invariant: answer in [lo, hi]
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (__ldg(&tv.leaf_code[sg.leaf0 + mid]) <= code) lo = mid;
    else hi = mid - 1;
  }
  return lo;
}

// Complete search of one warp's 32 queries in segment `sg`: seed leaves [seed-1, seed+1], then the pruned walk.
template <int K>
__device__ __forceinline__ void knn_search(const CloudView& tv, const Segment& sg, float qx, float qy, float qz, bool valid, int seed,
                                           TopK<K>& L, float4* wbuf) {
  if (sg.nleaf == 0) return;
  const int s0 = max(seed - 1, 0), s1 = min(seed + 1, sg.nleaf - 1);
  for (int lf = s0; lf <= s1; lf++) scan_leaf<K>(tv.pts, sg.p0 + lf * kLeaf, qx, qy, qz, valid, L, wbuf);
  tree_walk<K>(tv, sg, qx, qy, qz, valid, L, s0, s1, wbuf);
}

}  // namespace sicp

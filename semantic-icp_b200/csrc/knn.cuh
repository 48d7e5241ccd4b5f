// knn.cuh — exact k-nearest-neighbour search on the Morton-sorted leaf/box tree (device side).
// Replaces pcl::KdTreeFLANN::nearestKSearch at impl/gicp.hpp:69,196, impl/semantic_icp.hpp:68,
// impl/em_icp.hpp:60,221,296 and impl/semantic_point_cloud.hpp:41.
//
// Contract (bit-exact vs the oracle): the k targets minimising (d2, original index) lexicographically with
//   d2 = fl(fl(fl(dx*dx)+fl(dy*dy))+fl(dz*dz)),  dx = fl(q.x - p.x)            (FLANN L2_Simple<float>, no FMA)
// Pruning uses box lower bounds evaluated with the same rounded operation sequence; rounding is monotone, so a
// box is skipped only if every point in it is strictly farther than the current k-th best.
//
// Execution model: one warp owns 32 Morton-consecutive queries (one leaf of the query cloud) and walks the target
// tree as a PACKET with warp-uniform control flow: depth-first, children of a node visited nearest-first (key = the
// minimum box distance over the 32 lanes), a node entered only if some lane still needs it.  The traversal stack is
// warp-uniform and lives in shared memory.  A surviving leaf is staged into shared memory with one coalesced
// 512-byte load and every lane scans the 32 candidates from shared memory (broadcast reads).
#pragma once
#include "common.cuh"

namespace sicp {

constexpr int kMortonBits = 19;                       // per axis; 57-bit code + 7 bits of class rank in the sort key
constexpr uint64_t kMortonMask = (1ull << 57) - 1;
constexpr unsigned kFull = 0xffffffffu;
constexpr int kStackCap = 8 * kMaxLevels;             // DFS over an 8-ary tree: <= 7 pending siblings per level

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

// Shared memory owned by one warp during a search.
struct WarpScratch {
  float4 leaf[kLeaf];        // staged candidates
  int2 stack[kStackCap];     // x = level << 26 | idx, y = key bits (min box distance over the lanes)
};

// Packet search of one warp's 32 queries in segment `sg` (warp-uniform).  Exact for every valid lane.
template <int K>
__device__ __forceinline__ void knn_search(const CloudView& tv, const Segment& sg, float qx, float qy, float qz, bool valid, TopK<K>& L,
                                           WarpScratch& ws) {
  if (sg.nleaf == 0) return;
  const int lane = threadIdx.x & 31;
  const int top = sg.nlevels - 1;
  int sp = 0;
  // virtual root: expand the top level (<= kArity nodes); afterwards pop / expand / scan
  int level = top + 1, idx = 0;
  for (;;) {
    if (level > 0) {
      // ---- expand: test the children of (level, idx), push the needed ones nearest-last so the nearest pops first
      const int cl = level - 1;
      const int c0 = idx * kArity;
      const int nc = min(kArity, sg.node_cnt[cl] - c0);
      unsigned mykey = 0x7f800000u;  // +inf: "not needed"
      int mychild = 0;
#pragma unroll
      for (int j = 0; j < kArity; j++) {
        if (j < nc) {
          const int ni = sg.node_off[cl] + c0 + j;
          const float4 lo = __ldg(&tv.node_lo[ni]), hi = __ldg(&tv.node_hi[ni]);
          const float lb = box_lb_rn(qx, qy, qz, lo, hi);
          const bool need = valid && !(lb > L.worst());
          const unsigned kmin = __reduce_min_sync(kFull, need ? __float_as_uint(lb) : 0x7f800000u);  // lb >= 0: bit order == value order
          if (lane == j) { mykey = kmin; mychild = c0 + j; }
        }
      }
      // rank of child `lane` among the 8 (ties by child index); needed children have the smallest ranks
      int rank = 0;
#pragma unroll
      for (int j = 0; j < kArity; j++) {
        const unsigned kj = __shfl_sync(kFull, mykey, j);
        rank += (kj < mykey) || (kj == mykey && j < lane);
      }
      const unsigned needed = __ballot_sync(kFull, lane < kArity && mykey != 0x7f800000u);
      const int m = __popc(needed);
      if (lane < kArity && mykey != 0x7f800000u) ws.stack[sp + (m - 1 - rank)] = make_int2((cl << 26) | mychild, (int)mykey);
      sp += m;
      __syncwarp();
    } else {
      // ---- leaf: stage 32 candidates through shared memory, every lane scans all of them
      const int slot0 = sg.p0 + idx * kLeaf;
      const float4 mine = __ldg(&tv.pts[slot0 + lane]);
      __syncwarp();
      ws.leaf[lane] = mine;
      __syncwarp();
      if (valid) {
#pragma unroll(K <= 4 ? 8 : 1)
        for (int j = 0; j < kLeaf; j++) {
          const float4 p = ws.leaf[j];
          L.consider(dist2_rn(qx, qy, qz, p), slot0 + j, __float_as_int(p.w), tv.pts);
        }
      }
    }
    // ---- pop the next node some lane still needs
    bool found = false;
    while (sp > 0) {
      const int2 e = ws.stack[--sp];
      // key = min box distance over lanes at push time; bounds only tighten, so this cull is conservative
      const float wmax = __uint_as_float(__reduce_max_sync(kFull, valid ? __float_as_uint(L.worst()) : 0u));
      if (__int_as_float(e.y) > wmax) continue;
      level = e.x >> 26;
      idx = e.x & ((1 << 26) - 1);
      if (level == 0) {  // exact per-lane re-test of the leaf box before paying for the scan
        const int ni = sg.node_off[0] + idx;
        const float4 lo = __ldg(&tv.node_lo[ni]), hi = __ldg(&tv.node_hi[ni]);
        const float lb = box_lb_rn(qx, qy, qz, lo, hi);
        if (!__any_sync(kFull, valid && !(lb > L.worst()))) continue;
      }
      found = true;
      break;
    }
    if (!found) break;
  }
}

}  // namespace sicp

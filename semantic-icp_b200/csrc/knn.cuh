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

constexpr unsigned kFull = 0xffffffffu;
constexpr int kStackCap = 8 * kMaxLevels;             // DFS over an 8-ary tree: <= 7 pending siblings per level

#ifdef SICP_STATS
static __device__ unsigned long long g_stats[8];  // per warp: 0 node expansions, 1 leaf scans, 3 phase-2 iterations; per lane: 2 insertions
#define SICP_STAT(i, v) do { if ((threadIdx.x & 31) == 0 || (i) == 2) atomicAdd(&sicp::g_stats[i], (unsigned long long)(v)); } while (0)
#else
#define SICP_STAT(i, v) do { } while (0)
#endif

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

// Sorted (ascending) list of the K best candidates in registers.  A candidate is ONE 64-bit key
//   key = float_bits(d2) << 32 | original_index
// d2 >= 0, so unsigned order of the bits is numeric order and the 64-bit unsigned compare IS the lexicographic
// (d2, original index) order of the exact-kNN contract — no separate tie-break path.
template <int K>
struct TopK {
  unsigned long long key[K];
  static constexpr unsigned long long kEmpty = (0x7f800000ull << 32) | 0xffffffffull;  // (+inf, no index)
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int i = 0; i < K; i++) key[i] = kEmpty;
  }
  __device__ __forceinline__ float worst() const { return __uint_as_float((unsigned)(key[K - 1] >> 32)); }
  __device__ __forceinline__ float dist(int i) const { return __uint_as_float((unsigned)(key[i] >> 32)); }
  __device__ __forceinline__ int orig(int i) const { return (int)(unsigned)(key[i] & 0xffffffffull); }  // -1 = empty
  __device__ __forceinline__ void consider(float dc, int oc) {
    const unsigned long long k = ((unsigned long long)__float_as_uint(dc) << 32) | (unsigned)oc;
    if (!(k < key[K - 1])) return;  // NaN distances (padding slots) have bit patterns above +inf and never pass
    // one branch-free compare-exchange pass: the candidate sinks to its place, every larger entry moves down one,
    // the previous worst falls off the end
    unsigned long long carry = k;
#pragma unroll
    for (int i = 0; i < K; i++) {
      const unsigned long long cur = key[i];
      const bool lt = carry < cur;
      key[i] = lt ? carry : cur;
      carry = lt ? cur : carry;
    }
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
      SICP_STAT(0, 1);
      __syncwarp();
    } else {
      // ---- leaf: stage 32 candidates through shared memory, every lane scans all of them
      const int slot0 = sg.p0 + idx * kLeaf;
      const float4 mine = __ldg(&tv.pts[slot0 + lane]);
      __syncwarp();
      ws.leaf[lane] = mine;
      __syncwarp();
      // phase 1: all 32 distances, remember which candidates pass this lane's current bound (no insertion yet);
      // phase 2: each lane inserts only its own survivors, so the warp iterates max-over-lanes(popcount) times
      // instead of once per candidate that ANY lane wants.
      unsigned pass = 0;
      if (valid) {
        const float w0 = L.worst();
#pragma unroll 8
        for (int j = 0; j < kLeaf; j++) {
          const float dj = dist2_rn(qx, qy, qz, ws.leaf[j]);
          pass |= (dj <= w0 ? 1u : 0u) << j;
        }
      }
      SICP_STAT(1, 1);
#ifdef SICP_STATS
      { const unsigned mx = __reduce_max_sync(kFull, (unsigned)__popc(pass)); SICP_STAT(3, mx); }
#endif
      while (pass) {
        const int j = __ffs(pass) - 1;
        pass &= pass - 1;
        const float4 p = ws.leaf[j];
        L.consider(dist2_rn(qx, qy, qz, p), __float_as_int(p.w));
        SICP_STAT(2, 1);
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

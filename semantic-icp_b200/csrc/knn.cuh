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
constexpr int kMaxSeed = 24;                           // seed leaves injected before the traversal (see knn_search)
constexpr int kStackCap = 8 * kMaxLevels + kMaxSeed;   // DFS over an 8-ary tree: <= 7 pending siblings per level, plus the seeds
constexpr int kForced = 1 << 25;                       // stack entry flag: seed leaf (nleaf < 2^25)
constexpr float kPrune = 1.0f - 1.0f / 262144.0f;      // packet-level bounds are deflated by 2^-18 before pruning: far above
                                                       // the ~2^-21 relative rounding of a box gap or a point distance

// ---- Morton quantisation shared by the build (cloud.cu) and the query-side home-leaf lookup
__device__ __forceinline__ int f2ord(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }
__device__ __forceinline__ uint32_t spread10(uint32_t v) {
  v &= 0x3ff;
  v = (v | (v << 16)) & 0x030000ff;
  v = (v | (v << 8)) & 0x0300f00f;
  v = (v | (v << 4)) & 0x030c30c3;
  v = (v | (v << 2)) & 0x09249249;
  return v;
}
#ifndef SICP_CURVE_HILBERT
#define SICP_CURVE_HILBERT 0
#endif
// 30-bit space-filling-curve key (10 bits per axis) of (x,y,z) in the bounding cube described by bb (ordered ints, see
// bbox_kernel).  The key only decides the ORDER of the points (which 32 of them share a leaf) and the home-leaf lookup;
// search results do not depend on it.  Default: Z-order (Morton).  -DSICP_CURVE_HILBERT=1 builds the Hilbert variant
// (Skilling's transpose algorithm: consecutive keys are always adjacent cells, so a run of 32 points never straddles one
// of the long jumps of the Z-order curve).  Measured on KITTI-shaped scans (profiles/r2_knn_curve_ab.md): Hilbert order
// cuts the leaves a 32-query packet scans from 12.2 to 9.2 (k = 4) and the instructions of the kernel by 12 %, but the
// kernel's duration is set by its few heaviest packets (queries with no target point nearby in a dense region), which
// get no lighter — a lone k = 4 search runs at 600 M queries/s against 688 M with Z-order, and batch throughput is equal.
__device__ __forceinline__ uint32_t morton30(float x, float y, float z, const int* __restrict__ bb) {
  float lo[3], ext = 0.f;
#pragma unroll
  for (int c = 0; c < 3; c++) { lo[c] = ord2f(bb[c]); ext = fmaxf(ext, ord2f(bb[3 + c]) - lo[c]); }
  if (!(ext > 0.f) || !isfinite(ext)) ext = 1.f;
  const float inv_cell = 1023.f / ext;
  const float p[3] = {x, y, z};
  uint32_t X[3];
#pragma unroll
  for (int c = 0; c < 3; c++) X[c] = (uint32_t)fminf(fmaxf((p[c] - lo[c]) * inv_cell, 0.f), 1023.f);
#if SICP_CURVE_HILBERT
#pragma unroll
  for (uint32_t Q = 512; Q > 1; Q >>= 1) {  // inverse undo
    const uint32_t P = Q - 1;
#pragma unroll
    for (int i = 0; i < 3; i++) {
      if (X[i] & Q) X[0] ^= P;
      else { const uint32_t t = (X[0] ^ X[i]) & P; X[0] ^= t; X[i] ^= t; }
    }
  }
  X[1] ^= X[0]; X[2] ^= X[1];  // Gray encode
  uint32_t t = 0;
#pragma unroll
  for (uint32_t Q = 512; Q > 1; Q >>= 1) if (X[2] & Q) t ^= Q - 1;
  X[0] ^= t; X[1] ^= t; X[2] ^= t;
  return (spread10(X[0]) << 2) | (spread10(X[1]) << 1) | spread10(X[2]);
#else
  return spread10(X[0]) | (spread10(X[1]) << 1) | (spread10(X[2]) << 2);
#endif
}

#ifdef SICP_STATS
static __device__ unsigned long long g_stats[8];  // per warp: 0 node expansions, 1 leaf scans, 3 phase-2 iterations; per lane: 2 insertions
#define SICP_STAT(i, v) do { if ((threadIdx.x & 31) == 0 || (i) == 2) atomicAdd(&sicp::g_stats[i], (unsigned long long)(v)); } while (0)
static __device__ unsigned g_warp_scans[16384], g_warp_cycles[16384];  // per query warp of the LAST search kernel: leaf scans, SM cycles
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
#ifndef SICP_PHASE1_FMA
#define SICP_PHASE1_FMA 1
#endif
#ifndef SICP_TOPK_F64CMP
#define SICP_TOPK_F64CMP 1
#endif
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
    for (int i = 0; i < K; i++) cmpxchg(key[i], carry);
  }
  // slot <- min(slot, carry), carry <- max(slot, carry).  The integer form costs 4 ISETP + 4 SEL per step, all on the
  // half-rate ALU pipe, which is what bounds the k = 20 search (ncu: 60 % of its instructions are this chain).  Keys
  // are bit patterns of FINITE NON-NEGATIVE doubles (high word = float bits of d2 <= 0x7f800000 < 0x7ff00000, sign
  // clear; the early-out above keeps NaN distances out), whose numeric order is their unsigned order, denormals
  // included (f64 never flushes) — so ONE f64 compare on the otherwise idle FP64 pipe decides the step and the ALU
  // pipe is left with the 4 selects.
  static __device__ __forceinline__ void cmpxchg(unsigned long long& slot, unsigned long long& carry) {
#if SICP_TOPK_F64CMP
    unsigned long long lo, hi;
    asm("{\n\t.reg .pred p;\n\t.reg .f64 a, b;\n\tmov.b64 a, %2;\n\tmov.b64 b, %3;\n\tsetp.lt.f64 p, b, a;\n\t"
        "selp.b64 %0, %3, %2, p;\n\tselp.b64 %1, %2, %3, p;\n\t}"
        : "=&l"(lo), "=&l"(hi) : "l"(slot), "l"(carry));  // early-clobber: the second select still reads both inputs
    slot = lo; carry = hi;
#else
    const unsigned long long cur = slot;
    const bool lt = carry < cur;
    slot = lt ? carry : cur;
    carry = lt ? cur : carry;
#endif
  }
};

// Shared memory owned by one warp during a search.
struct WarpScratch {
  float4 leaf[kLeaf];        // staged candidates
  int2 stack[kStackCap];     // x = level << 26 | idx, y = key bits (min box distance over the lanes)
};

// Leaf of segment `sg` whose Morton range contains the query (last leaf whose first key is <= the query's key).
__device__ __forceinline__ int home_leaf(const CloudView& tv, const Segment& sg, float qx, float qy, float qz) {
  const uint32_t key = morton30(qx, qy, qz, tv.bb);
  const uint32_t* lk = tv.leaf_key + sg.leaf0;
  int lo = 0, hi = sg.nleaf - 1;  // invariant: answer in [lo, hi]
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (__ldg(&lk[mid]) <= key) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// Packet search of one warp's 32 queries in segment `sg` (warp-uniform).  Exact for every valid lane.
//
// Seeding: a plain nearest-first descent gives good bounds only to the lanes closest to the first leaves; the others
// keep loose bounds for many scans ("needing" most leaves they meet).  So before the traversal every lane looks up its
// HOME leaf (Morton binary search) and the distinct home leaves of the packet (+- SEED Morton neighbours) are pushed on
// top of the stack: they are scanned first, after which every lane's k-th bound is already close to final.  Seed leaves
// are remembered (one per lane register) and skipped when the traversal reaches them again.
template <int K, class ListT>
__device__ __forceinline__ void knn_search(const CloudView& tv, const Segment& sg, float qx, float qy, float qz, bool valid, ListT& L,
                                           WarpScratch& ws) {
  if (sg.nleaf == 0) return;
  const int lane = threadIdx.x & 31;
  const int top = sg.nlevels - 1;
  int sp = 0;
#ifndef SICP_SEED_BIG
#define SICP_SEED_BIG 1
#endif
#ifndef SICP_SEED_SMALL
#define SICP_SEED_SMALL 0
#endif
#ifndef SICP_SPLIT_SMALL
#define SICP_SPLIT_SMALL 2
#endif
  constexpr int SEED = (K >= 8) ? SICP_SEED_BIG : SICP_SEED_SMALL;  // Morton neighbours of a home leaf seeded with it (tuned: tools/probe_cov.py A/B)
  const int home = valid ? home_leaf(tv, sg, qx, qy, qz) : -1;
  int myseed = -1;   // lane i remembers the i-th seed leaf
  bool seeded = false;
#ifdef SICP_STATS
  unsigned dbg_scans = 0;
  const long long dbg_t0 = clock64();
  const int dbg_w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
#endif
  // AABB of the packet's (valid) queries, for the node tests of the expansion step
  float pk_lo[3], pk_hi[3];
  {
    const float q[3] = {qx, qy, qz};
#pragma unroll
    for (int c = 0; c < 3; c++) {
      pk_lo[c] = ord2f(__reduce_min_sync(kFull, valid ? f2ord(q[c]) : 0x7fffffff));
      pk_hi[c] = ord2f(__reduce_max_sync(kFull, valid ? f2ord(q[c]) : (int)0x80000000));
    }
  }
  if (!__any_sync(kFull, valid)) return;
  // virtual root: expand the top level (<= kArity nodes); afterwards pop / expand / scan
  int level = top + 1, idx = 0;
  for (;;) {
    if (level > 0) {
      // ---- expand: test the children of (level, idx), push the needed ones nearest-last so the nearest pops first
      const int cl = level - 1;
      const int c0 = idx * kArity;
      const int nc = min(kArity, sg.node_cnt[cl] - c0);
      // Stage 1, one lane per child: box-to-box lower bound between the child's AABB and the AABB of the packet's
      // queries against the loosest k-th bound of the packet — conservative for every lane (kPrune covers the rounding
      // of both sides).  Stage 2, only for the children that survive: the exact per-lane bound, whose minimum over the
      // lanes that need the child is the ordering / culling key.
      unsigned mykey = 0x7f800000u;  // +inf: "not needed"
      const int mychild = c0 + lane;
      const float wmax_e = __uint_as_float(__reduce_max_sync(kFull, valid ? __float_as_uint(L.worst()) : 0u));
      bool maybe = false;
      if (lane < nc) {
        const int ni = sg.node_off[cl] + c0 + lane;
        const float4 lo = __ldg(&tv.node_lo[ni]), hi = __ldg(&tv.node_hi[ni]);
        const float gx = fmaxf(fmaxf(lo.x - pk_hi[0], pk_lo[0] - hi.x), 0.f);
        const float gy = fmaxf(fmaxf(lo.y - pk_hi[1], pk_lo[1] - hi.y), 0.f);
        const float gz = fmaxf(fmaxf(lo.z - pk_hi[2], pk_lo[2] - hi.z), 0.f);
        maybe = !((gx * gx + gy * gy + gz * gz) * kPrune > wmax_e);
      }
      unsigned cand = __ballot_sync(kFull, maybe);
      while (cand) {
        const int j = __ffs(cand) - 1;
        cand &= cand - 1;
        const int ni = sg.node_off[cl] + c0 + j;
        const float4 lo = __ldg(&tv.node_lo[ni]), hi = __ldg(&tv.node_hi[ni]);
        const float lb = box_lb_rn(qx, qy, qz, lo, hi);
        const bool need = valid && !(lb > L.worst());
        const unsigned kmin = __reduce_min_sync(kFull, need ? __float_as_uint(lb) : 0x7f800000u);  // lb >= 0: bit order == value order
        if (lane == j) mykey = kmin;
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
      if (!seeded) {  // right after the virtual-root expansion: seed leaves go on top of the stack
        seeded = true;
        unsigned pending = __ballot_sync(kFull, home >= 0);
        int nseed = 0;
        while (pending && nseed + 2 * SEED + 1 <= kMaxSeed) {
          const int h = __shfl_sync(kFull, home, __ffs(pending) - 1);
#pragma unroll
          for (int dl = SEED; dl >= -SEED; dl--) {
            const int lf = h + dl;
            if (lf < 0 || lf >= sg.nleaf) continue;
            if (__any_sync(kFull, myseed == lf)) continue;
            if (lane == nseed) myseed = lf;
            if (lane == 0) ws.stack[sp] = make_int2(kForced | lf, 0);
            nseed++; sp++;
          }
          pending &= ~__ballot_sync(kFull, home == h);
        }
        __syncwarp();
      }
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
      // The leaf is taken in kSplit parts so that later parts are filtered against the bound tightened by earlier ones.
      SICP_STAT(1, 1);
#ifdef SICP_STATS
      dbg_scans++;
#endif
      constexpr int kSplit = (K >= 8) ? 1 : SICP_SPLIT_SMALL, kPart = kLeaf / kSplit;  // measured: helps short lists (k<=4: -3%), not k = 20
#pragma unroll 1
      for (int h = 0; h < kSplit; h++) {
        unsigned pass = 0;
        if (valid) {
#if SICP_PHASE1_FMA
          // phase 1 only FILTERS (phase 2 recomputes the contract's d2 and compares the full key), so it may use the
          // contracted form (3 FADD + FMUL + 2 FFMA instead of 5 FADD + 3 FMUL on the pipe that bounds this loop) against a
          // threshold inflated by far more than the two forms can differ: both are within 2^-22 relative of the real sum of
          // squares of the SAME rounded differences (plus a few 2^-149 when products are denormal, covered by FLT_MIN).
          const float w0 = __fmaf_rn(L.worst(), 1.0f + 1.0f / 262144.0f, 1.17549435e-38f);
#pragma unroll
          for (int j = 0; j < kPart; j++) {
            const float4 p = ws.leaf[h * kPart + j];
            const float dx = __fsub_rn(qx, p.x), dy = __fsub_rn(qy, p.y), dz = __fsub_rn(qz, p.z);
            const float dj = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
            pass |= (dj <= w0 ? 1u : 0u) << j;
          }
#else
          const float w0 = L.worst();
#pragma unroll
          for (int j = 0; j < kPart; j++) {
            const float dj = dist2_rn(qx, qy, qz, ws.leaf[h * kPart + j]);
            pass |= (dj <= w0 ? 1u : 0u) << j;
          }
#endif
        }
#ifdef SICP_STATS
        { const unsigned mx = __reduce_max_sync(kFull, (unsigned)__popc(pass)); SICP_STAT(3, mx); }
#endif
        while (pass) {
          const int j = __ffs(pass) - 1;
          pass &= pass - 1;
          const float4 p = ws.leaf[h * kPart + j];
          L.consider(dist2_rn(qx, qy, qz, p), __float_as_int(p.w));
          SICP_STAT(2, 1);
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
      idx = e.x & (kForced - 1);
      if (level == 0) {  // exact per-lane re-test of the leaf box before paying for the scan
        if (!(e.x & kForced) && __any_sync(kFull, myseed == idx)) continue;  // already scanned as a seed
        const int ni = sg.node_off[0] + idx;
        const float4 lo = __ldg(&tv.node_lo[ni]), hi = __ldg(&tv.node_hi[ni]);
        const float lb = box_lb_rn(qx, qy, qz, lo, hi);
        const unsigned needers = __ballot_sync(kFull, valid && !(lb > L.worst()));
        if (!needers) continue;
        SICP_STAT(4, __popc(needers));
        SICP_STAT(5, __popc(needers) <= 8 ? 1 : 0);
        SICP_STAT(6, __popc(needers) <= 16 ? 1 : 0);
      }
      found = true;
      break;
    }
    if (!found) break;
  }
#ifdef SICP_STATS
  if (lane == 0 && dbg_w < 16384) { g_warp_scans[dbg_w] = dbg_scans; g_warp_cycles[dbg_w] = (unsigned)(clock64() - dbg_t0); }
#endif
}

}  // namespace sicp

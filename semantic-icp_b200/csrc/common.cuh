// common.cuh — shared host/device definitions of libsicp_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <mutex>
#include <string>
#include <vector>
#include "sicp_b200.h"

namespace sicp {

// ------------------------------------------------------------------ error plumbing
void set_error(const std::string& msg);
#define SICP_CUDA(call)                                                                                   \
  do {                                                                                                    \
    cudaError_t _e = (call);                                                                              \
    if (_e != cudaSuccess) {                                                                              \
      ::sicp::set_error(std::string(#call) + " failed: " + cudaGetErrorString(_e) + " (" __FILE__ ":" +   \
                        std::to_string(__LINE__) + ")");                                                  \
      return SICP_ERR_CUDA;                                                                               \
    }                                                                                                     \
  } while (0)
#define SICP_CHECK(st)                     \
  do {                                     \
    sicp_status _s = (st);                 \
    if (_s != SICP_OK) return _s;          \
  } while (0)
#define SICP_REQUIRE(cond, msg)            \
  do {                                     \
    if (!(cond)) {                         \
      ::sicp::set_error(msg);              \
      return SICP_ERR_INVALID;             \
    }                                      \
  } while (0)

cudaStream_t current_stream();
void count_launches(int n);  // per-thread kernel launch counter (sicp_launch_count)
// Sophus::SE3d is a unit quaternion by construction; a pose7 that comes through the C ABI has to be checked
// (finite, |q| = 1) before it reaches quat_to_R.  nullable: a null pointer is accepted (optional pose arguments).
sicp_status validate_pose7(const double* p7, const char* what, bool nullable = false);

// pinned staging blocks for small host->device tables (see cloud.cu)
struct PinnedBlock { void* p; size_t bytes; cudaEvent_t ev; };
void* pinned_stage(size_t bytes, cudaStream_t st, PinnedBlock* out);
void pinned_release(PinnedBlock& b, cudaStream_t st);

// ------------------------------------------------------------------ search structure
constexpr int kLeaf = 32;        // points per leaf == warp size: one warp owns one leaf of queries
constexpr int kArity = 8;        // children per internal node
constexpr int kMaxLevels = 10;   // 32 * 8^9 points
constexpr int kMaxClasses = 64;  // label vectors live in registers of two lanes-worth (N <= 64)
constexpr int kMaxK = 32;
constexpr size_t kMaxCloudPoints = (size_t)1 << 26;  // 32-bit launch / index arithmetic of the per-slot kernels holds up to here (x32 lanes, x4 candidates)

// One searchable point set: the whole cloud, or one semantic class (SemanticPointCloud::labeledKdTrees).
struct Segment {
  int p0;                      // first slot in the sorted arrays (multiple of kLeaf)
  int n;                       // real points
  int nleaf;                   // ceil(n / kLeaf)
  int nlevels;                 // levels in the implicit tree; level 0 = leaves
  int node_off[kMaxLevels];    // offset of each level in the node arrays
  int node_cnt[kMaxLevels];
  int leaf0;                   // index of this segment's first leaf
  int start;                   // first position of this segment in the key-sorted order
  uint32_t label;              // class label (PER_CLASS) or 0
};

// Device view of a cloud, passed to kernels by value.
struct CloudView {
  int n;                 // real points
  int nslots;            // padded slots (multiple of kLeaf per segment)
  int nseg;
  const float4* pts;     // [nslots] sorted: x,y,z, original index bits (pad: NaN, -1)
  const uint32_t* label; // [nslots] sorted labels (0 if none)
  const int* seg_of_leaf;// [nleaf_total] segment id of each leaf (== of each query warp)
  const Segment* seg;    // [nseg]
  const float4* node_lo; // node boxes, all segments / levels
  const float4* node_hi;
  const uint32_t* leaf_key; // [nleaf_total] 30-bit Morton code of each leaf's first point (ascending inside a segment)
  const int* bb;            // [8] ordered-int bounding box of the cloud (the Morton quantisation frame)
  const double* nrm;     // [nslots][4] normals (nx, ny, nz, 0): one 32-byte sector per gathered normal  (after precompute)
  const double* avec;    // [nslots*N] label vectors a_p = CM^T dist_p (EM)
  int N;
};

}  // namespace sicp

struct sicp_cloud {
  int device = 0;
  int layout = 0;
  size_t n = 0;
  int nslots = 0, nseg = 0, nleaf = 0, nnodes = 0;
  std::vector<sicp::Segment> h_seg;
  std::vector<uint32_t> class_labels;  // first-appearance order
  // device buffers (the first group is carved out of one slab)
  void* d_slab = nullptr;
  int* d_bb = nullptr;            // [8] ordered-int bounding box + label min/max
  float4* d_pts = nullptr;
  uint32_t* d_label = nullptr;
  int* d_seg_of_leaf = nullptr;
  sicp::Segment* d_seg = nullptr;
  float4* d_node_lo = nullptr;
  float4* d_node_hi = nullptr;
  int* d_slot_of_orig = nullptr;  // [n] inverse permutation
  uint32_t* d_leaf_key = nullptr; // [nleaf]
  double* d_nrm = nullptr;
  double* d_avec = nullptr;
  // precompute cache key
  bool pre_valid = false;
  int pre_k = 0, pre_N = 0;
  double pre_eps = 0;
  std::vector<double> pre_cm;
  bool has_labels = false;
  uint32_t min_label = 0, max_label = 0;
  bool label_range_known = false;
  cudaEvent_t ready_ev = nullptr;  // recorded after the last precompute; consumers on other streams wait on it
  cudaEvent_t built_ev = nullptr;  // recorded after the build (upload, sort, tree); work on other streams waits on it
  std::mutex mu;                   // guards the precompute cache (pairs of an odometry chain share clouds across host threads)
  // Deferred build: the sort / tree build of a WHOLE cloud runs on the stream of its FIRST consumer (sicp::ensure_built), so
  // that the builds of a batch overlap the registrations already in flight instead of all preceding them on the creating
  // stream.  Until then the inputs live in owned staging buffers; the host-side layout (nslots, nleaf ...) is final at once.
  bool pending_build = false;
  float* stage_xyz = nullptr;
  uint32_t* stage_lab = nullptr;
  cudaEvent_t staged_ev = nullptr;  // recorded on the creating stream after the staging copies
  std::mutex build_mu;
  sicp::CloudView view() const;
};

namespace sicp {
// Every entry point that reads a cloud on stream `st` calls one of these first: the cloud may still have to be built
// (deferred build), or was built / precomputed on another stream or by another host thread (a wait on a completed event
// costs nothing).  ensure_ready also waits for the last covariance precompute.
sicp_status ensure_built(const sicp_cloud* c, cudaStream_t st);
sicp_status ensure_ready(const sicp_cloud* c, cudaStream_t st);
}  // namespace sicp

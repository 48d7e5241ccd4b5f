// cloud.cu — device-resident clouds: SoA upload, Morton sort, implicit 8-ary box tree over 32-point leaves.
// Replaces the kd-tree construction of GICP::setSourceCloud/setTargetCloud (gicp.h:42-63),
// EmIterativeClosestPoint::set*Cloud (em_icp.h:50-66) and pcl_2_semantic + addSemanticCloud
// (pcl_2_semantic.h:14-42, impl/semantic_point_cloud.hpp:17-23).
#include <cub/device/device_radix_sort.cuh>
#include <algorithm>
#include <cmath>
#include <cstring>
#include <mutex>
#include "common.cuh"
#include "knn.cuh"
#include "se3.cuh"

namespace sicp {

static thread_local std::string g_err;
static thread_local cudaStream_t g_stream = nullptr;
void set_error(const std::string& msg) { g_err = msg; }
cudaStream_t current_stream() { return g_stream; }
static thread_local uint64_t g_launches = 0;
void count_launches(int n) { g_launches += (uint64_t)n; }

// ---- pinned staging: small host tables go to the device with truly asynchronous copies (a cudaMemcpyAsync from
// pageable memory synchronises the stream first).  Blocks are recycled once the event recorded after their last
// use has completed.
static std::mutex g_pin_mu;
static std::vector<PinnedBlock> g_pin_pool;
void* pinned_stage(size_t bytes, cudaStream_t st, PinnedBlock* out) {
  bytes = std::max<size_t>(bytes, 4096);
  {
    std::lock_guard<std::mutex> lk(g_pin_mu);
    for (size_t i = 0; i < g_pin_pool.size(); i++)
      if (g_pin_pool[i].bytes >= bytes && cudaEventQuery(g_pin_pool[i].ev) == cudaSuccess) {
        *out = g_pin_pool[i];
        g_pin_pool.erase(g_pin_pool.begin() + i);
        return out->p;
      }
  }
  out->p = nullptr; out->bytes = bytes;
  if (cudaMallocHost(&out->p, bytes) != cudaSuccess) return nullptr;
  cudaEventCreateWithFlags(&out->ev, cudaEventDisableTiming);
  (void)st;
  return out->p;
}
void pinned_release(PinnedBlock& b, cudaStream_t st) {  // call after the last copy that reads the block was enqueued
  cudaEventRecord(b.ev, st);
  std::lock_guard<std::mutex> lk(g_pin_mu);
  g_pin_pool.push_back(b);
}

// ---- bounding box: block reduce + ordered-int atomics (f2ord / ord2f / morton30 live in knn.cuh) --------------------

// bb[0..5] = ordered-int min/max of xyz; bb[6], bb[7] = min/max label
__global__ void bbox_init_kernel(int* bb) {
  if (threadIdx.x < 3) bb[threadIdx.x] = f2ord(INFINITY);
  else if (threadIdx.x < 6) bb[threadIdx.x] = f2ord(-INFINITY);
  else if (threadIdx.x == 6) bb[6] = -1;  // 0xffffffff as unsigned
  else if (threadIdx.x == 7) bb[7] = 0;
}
__global__ void bbox_kernel(const float* __restrict__ xyz, const uint32_t* __restrict__ labels, int n, int* bb) {
  float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  unsigned llo = 0xffffffffu, lhi = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    for (int c = 0; c < 3; c++) {
      float v = xyz[3 * (size_t)i + c];
      lo[c] = fminf(lo[c], v);
      hi[c] = fmaxf(hi[c], v);
    }
    if (labels) { const unsigned l = labels[i]; llo = min(llo, l); lhi = max(lhi, l); }
  }
  for (int c = 0; c < 3; c++) {
    for (int o = 16; o; o >>= 1) {
      lo[c] = fminf(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
      hi[c] = fmaxf(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
    }
    if ((threadIdx.x & 31) == 0) {
      atomicMin(&bb[c], f2ord(lo[c]));
      atomicMax(&bb[3 + c], f2ord(hi[c]));
    }
  }
  if (labels) {
    for (int o = 16; o; o >>= 1) { llo = min(llo, __shfl_xor_sync(0xffffffffu, llo, o)); lhi = max(lhi, __shfl_xor_sync(0xffffffffu, lhi, o)); }
    if ((threadIdx.x & 31) == 0) { atomicMin((unsigned*)&bb[6], llo); atomicMax((unsigned*)&bb[7], lhi); }
  }
}
// 30-bit Morton code in the cloud's bounding cube (morton30, knn.cuh); PER_CLASS clouds put the class rank above it
template <typename KeyT>
__global__ void key_kernel(const float* __restrict__ xyz, const uint8_t* __restrict__ rank, int n, const int* __restrict__ bb, KeyT* keys,
                           uint32_t* vals) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t m = morton30(xyz[3 * (size_t)i], xyz[3 * (size_t)i + 1], xyz[3 * (size_t)i + 2], bb);
  KeyT k = m;
  if (sizeof(KeyT) == 8 && rank) k |= (KeyT)((uint64_t)rank[i] << 32);
  keys[i] = k;
  vals[i] = (uint32_t)i;
}
__global__ void seg_of_leaf_kernel(const Segment* __restrict__ seg, int nseg, int nleaf, int* out) {
  int leaf = blockIdx.x * blockDim.x + threadIdx.x;
  if (leaf >= nleaf) return;
  int s = 0;
  while (s + 1 < nseg && seg[s + 1].leaf0 <= leaf) s++;
  out[leaf] = s;
}
// one thread per slot: gather the sorted point (or a NaN pad) and record the inverse permutation
__global__ void slot_kernel(const float* __restrict__ xyz, const uint32_t* __restrict__ labels, const uint32_t* __restrict__ sorted_idx,
                            const int* __restrict__ seg_of_leaf, const Segment* __restrict__ seg, int nslots, float4* pts,
                            uint32_t* label_out, int* slot_of_orig) {
  int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= nslots) return;
  const int sid = seg_of_leaf[slot / kLeaf];
  const int r = slot - seg[sid].p0;
  float4 p = make_float4(__int_as_float(0x7fc00000), __int_as_float(0x7fc00000), __int_as_float(0x7fc00000), __int_as_float(-1));
  uint32_t lab = 0;
  if (r < seg[sid].n) {
    const uint32_t o = sorted_idx[seg[sid].start + r];
    p = make_float4(xyz[3 * (size_t)o], xyz[3 * (size_t)o + 1], xyz[3 * (size_t)o + 2], __int_as_float((int)o));
    if (labels) lab = labels[o];
    slot_of_orig[o] = slot;
  }
  pts[slot] = p;
  label_out[slot] = lab;
}
// level 0: one warp per leaf
__global__ void leaf_box_kernel(const float4* __restrict__ pts, const int* __restrict__ seg_of_leaf, const Segment* __restrict__ seg,
                                int nleaf, const int* __restrict__ bb, float4* node_lo, float4* node_hi, uint32_t* leaf_key) {
  const int leaf = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) / 32);
  if (leaf >= nleaf) return;
  const float4 p = pts[(size_t)leaf * kLeaf + (threadIdx.x & 31)];
  float lo[3] = {p.x, p.y, p.z}, hi[3] = {p.x, p.y, p.z};
  for (int c = 0; c < 3; c++)
    for (int o = 16; o; o >>= 1) {
      lo[c] = fminf(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));  // fminf/fmaxf drop the NaN pads
      hi[c] = fmaxf(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
    }
  if ((threadIdx.x & 31) == 0) {
    const Segment& sg = seg[seg_of_leaf[leaf]];
    const int li = sg.node_off[0] + (leaf - sg.leaf0);
    node_lo[li] = make_float4(lo[0], lo[1], lo[2], 0.f);
    node_hi[li] = make_float4(hi[0], hi[1], hi[2], 0.f);
    leaf_key[leaf] = morton30(p.x, p.y, p.z, bb);  // lane 0 holds the leaf's first (never padding) point
  }
}
// upper levels: one block per segment, levels separated by __syncthreads
__global__ void upper_box_kernel(const Segment* __restrict__ seg, float4* node_lo, float4* node_hi) {
  const Segment sg = seg[blockIdx.x];
  for (int l = 1; l < sg.nlevels; l++) {
    for (int i = threadIdx.x; i < sg.node_cnt[l]; i += blockDim.x) {
      const int c0 = i * kArity, c1 = min(c0 + kArity, sg.node_cnt[l - 1]);
      float4 lo = node_lo[sg.node_off[l - 1] + c0], hi = node_hi[sg.node_off[l - 1] + c0];
      for (int c = c0 + 1; c < c1; c++) {
        const float4 a = node_lo[sg.node_off[l - 1] + c], b = node_hi[sg.node_off[l - 1] + c];
        lo.x = fminf(lo.x, a.x); lo.y = fminf(lo.y, a.y); lo.z = fminf(lo.z, a.z);
        hi.x = fmaxf(hi.x, b.x); hi.y = fmaxf(hi.y, b.y); hi.z = fmaxf(hi.z, b.z);
      }
      node_lo[sg.node_off[l] + i] = lo;
      node_hi[sg.node_off[l] + i] = hi;
    }
    __syncthreads();
  }
}
struct Mat34f { float m[12]; };  // rows of a float 4x4 (last row 0 0 0 1)
// slots -> original order, p' = M * p in float arithmetic, left to right, no FMA contraction (x86 float code has none)
__global__ void transform_f32_kernel(const float4* __restrict__ pts, int nslots, Mat34f M, float* __restrict__ xyz_out) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nslots) return;
  const float4 p = pts[s];
  const int o = __float_as_int(p.w);
  if (o < 0) return;
#pragma unroll
  for (int r = 0; r < 3; r++)
    xyz_out[3 * (size_t)o + r] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(M.m[4 * r], p.x), __fmul_rn(M.m[4 * r + 1], p.y)), __fmul_rn(M.m[4 * r + 2], p.z)), M.m[4 * r + 3]);
}
// -------------------------------------------------------------------------------------------------------------
static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

// Host-side layout of the search structure (segments, slots, leaves, nodes): final as soon as the class sizes are known.
static void layout_cloud(sicp_cloud* c, const std::vector<int>& class_sizes) {
  const int nseg = (int)class_sizes.size();
  c->nseg = nseg;
  c->h_seg.assign(nseg, Segment());
  int p0 = 0, leaf0 = 0, node0 = 0, start = 0;
  for (int s = 0; s < nseg; s++) {
    Segment& sg = c->h_seg[s];
    std::memset(&sg, 0, sizeof sg);
    sg.p0 = p0; sg.n = class_sizes[s]; sg.nleaf = (sg.n + kLeaf - 1) / kLeaf; sg.leaf0 = leaf0; sg.start = start;
    sg.label = c->layout == SICP_CLOUD_PER_CLASS ? c->class_labels[s] : 0;
    int cnt = sg.nleaf, l = 0;
    for (;;) {
      sg.node_off[l] = node0; sg.node_cnt[l] = cnt; node0 += cnt; l++;
      if (cnt <= kArity || l >= kMaxLevels) break;
      cnt = (cnt + kArity - 1) / kArity;
    }
    sg.nlevels = l;
    start += sg.n; p0 += sg.nleaf * kLeaf; leaf0 += sg.nleaf;
  }
  c->nslots = p0; c->nleaf = leaf0; c->nnodes = node0;
}

// Device side of the build.  Everything here is enqueued on `st`; the function does not wait for the device.
static sicp_status build_cloud(sicp_cloud* c, const float* d_xyz, const uint32_t* d_labels, const uint8_t* d_rank, cudaStream_t st) {
  const int n = (int)c->n;
  const int nseg = c->nseg;

  // one persistent slab: pts | label | seg_of_leaf | seg | node_lo | node_hi | slot_of_orig | bb
  size_t off = 0;
  auto carve = [&](size_t bytes) { size_t o = off; off += align256(std::max<size_t>(bytes, 16)); return o; };
  const size_t o_pts = carve(sizeof(float4) * c->nslots), o_lab = carve(sizeof(uint32_t) * c->nslots), o_sol = carve(sizeof(int) * c->nleaf),
               o_seg = carve(sizeof(Segment) * nseg), o_nlo = carve(sizeof(float4) * c->nnodes), o_nhi = carve(sizeof(float4) * c->nnodes),
               o_soo = carve(sizeof(int) * n), o_bb = carve(sizeof(int) * 8), o_lk = carve(sizeof(uint32_t) * c->nleaf);
  SICP_CUDA(cudaMallocAsync(&c->d_slab, off, st));
  char* base = (char*)c->d_slab;
  c->d_pts = (float4*)(base + o_pts); c->d_label = (uint32_t*)(base + o_lab); c->d_seg_of_leaf = (int*)(base + o_sol);
  c->d_seg = (Segment*)(base + o_seg); c->d_node_lo = (float4*)(base + o_nlo); c->d_node_hi = (float4*)(base + o_nhi);
  c->d_slot_of_orig = (int*)(base + o_soo); c->d_bb = (int*)(base + o_bb);
  c->d_leaf_key = (uint32_t*)(base + o_lk);

  PinnedBlock pb;
  void* stage = pinned_stage(sizeof(Segment) * std::max(1, nseg), st, &pb);
  if (!stage) { set_error("pinned staging allocation failed"); return SICP_ERR_CUDA; }
  std::memcpy(stage, c->h_seg.data(), sizeof(Segment) * nseg);
  SICP_CUDA(cudaMemcpyAsync(c->d_seg, stage, sizeof(Segment) * nseg, cudaMemcpyHostToDevice, st));
  pinned_release(pb, st);
  const int T = 256;
  bbox_init_kernel<<<1, 32, 0, st>>>(c->d_bb);
  if (n == 0) return SICP_OK;

  // temporaries: keys | keys2 | vals | vals2 | cub
  const bool wide = d_rank != nullptr;
  const size_t kb = wide ? 8 : 4;
  size_t cub_bytes = 0;
  if (wide) SICP_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (uint64_t*)nullptr, (uint64_t*)nullptr, (uint32_t*)nullptr, (uint32_t*)nullptr, n, 0, 39, st));
  else SICP_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (uint32_t*)nullptr, (uint32_t*)nullptr, (uint32_t*)nullptr, (uint32_t*)nullptr, n, 0, 30, st));
  size_t toff = 0;
  auto tcarve = [&](size_t bytes) { size_t o = toff; toff += align256(std::max<size_t>(bytes, 16)); return o; };
  const size_t t_k1 = tcarve(kb * n), t_k2 = tcarve(kb * n), t_v1 = tcarve(4 * (size_t)n), t_v2 = tcarve(4 * (size_t)n), t_cub = tcarve(cub_bytes);
  char* tmp;
  SICP_CUDA(cudaMallocAsync(&tmp, toff, st));
  uint32_t* vals = (uint32_t*)(tmp + t_v1); uint32_t* vals2 = (uint32_t*)(tmp + t_v2);

  bbox_kernel<<<std::min((n + T - 1) / T, 296), T, 0, st>>>(d_xyz, d_labels, n, c->d_bb);
  if (wide) {
    key_kernel<uint64_t><<<(n + T - 1) / T, T, 0, st>>>(d_xyz, d_rank, n, c->d_bb, (uint64_t*)(tmp + t_k1), vals);
    SICP_CUDA(cub::DeviceRadixSort::SortPairs(tmp + t_cub, cub_bytes, (uint64_t*)(tmp + t_k1), (uint64_t*)(tmp + t_k2), vals, vals2, n, 0, 39, st));
  } else {
    key_kernel<uint32_t><<<(n + T - 1) / T, T, 0, st>>>(d_xyz, nullptr, n, c->d_bb, (uint32_t*)(tmp + t_k1), vals);
    SICP_CUDA(cub::DeviceRadixSort::SortPairs(tmp + t_cub, cub_bytes, (uint32_t*)(tmp + t_k1), (uint32_t*)(tmp + t_k2), vals, vals2, n, 0, 30, st));
  }
  if (nseg == 1) SICP_CUDA(cudaMemsetAsync(c->d_seg_of_leaf, 0, sizeof(int) * c->nleaf, st));
  else seg_of_leaf_kernel<<<(c->nleaf + T - 1) / T, T, 0, st>>>(c->d_seg, nseg, c->nleaf, c->d_seg_of_leaf);
  slot_kernel<<<(c->nslots + T - 1) / T, T, 0, st>>>(d_xyz, d_labels, vals2, c->d_seg_of_leaf, c->d_seg, c->nslots, c->d_pts, c->d_label,
                                                     c->d_slot_of_orig);
  leaf_box_kernel<<<(unsigned)(((size_t)c->nleaf * 32 + T - 1) / T), T, 0, st>>>(c->d_pts, c->d_seg_of_leaf, c->d_seg, c->nleaf, c->d_bb, c->d_node_lo, c->d_node_hi, c->d_leaf_key);
  upper_box_kernel<<<nseg, 256, 0, st>>>(c->d_seg, c->d_node_lo, c->d_node_hi);
  count_launches(6 + (nseg > 1) + ((wide ? 39 : 30) + 7) / 8 + 2);  // own kernels + CUB onesweep (histogram, scan, one pass per 8 key bits)
  SICP_CUDA(cudaGetLastError());
  SICP_CUDA(cudaFreeAsync(tmp, st));
  return SICP_OK;
}

// ---- first-appearance class order (pcl_2_semantic.h:24-35) on the device -------------------------------------------
// The reference walks the cloud once and opens a new class the first time it meets a label; classes are then used in
// that order.  Here every point inserts its label into a small open-addressing table (64-bit CAS on (1<<32 | label)),
// takes the atomic minimum of the point index per label (= first appearance) and counts; the host sorts the <= 128
// table entries by first appearance and a second kernel maps every point to its class rank.  Only the 6 KB table
// crosses the bus.
constexpr int kClassSlots = 512;
struct ClassTable {
  unsigned long long key[kClassSlots];  // 0 = empty, else (1 << 32) | label
  unsigned first[kClassSlots];          // smallest original index with this label
  unsigned count[kClassSlots];
  unsigned overflow;                    // more distinct labels than slots
  unsigned pad;
};
__device__ __forceinline__ unsigned label_hash(uint32_t l) { return (l * 2654435761u) >> 23; }  // top 9 bits -> [0, 512)
__global__ void classify_count_kernel(const uint32_t* __restrict__ labels, int n, ClassTable* tab, uint16_t* __restrict__ slot_of_point) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t l = labels[i];
  const unsigned long long k = (1ull << 32) | l;
  unsigned h = label_hash(l);
  int probes = 0;
  for (;; h = (h + 1) & (kClassSlots - 1)) {
    const unsigned long long prev = atomicCAS(&tab->key[h], 0ull, k);
    if (prev == 0ull || prev == k) break;
    if (++probes >= kClassSlots) { tab->overflow = 1; slot_of_point[i] = 0; return; }
  }
  atomicMin(&tab->first[h], (unsigned)i);
  atomicAdd(&tab->count[h], 1u);
  slot_of_point[i] = (uint16_t)h;
}
__global__ void classify_rank_kernel(const uint16_t* __restrict__ slot_of_point, int n, const uint8_t* __restrict__ rank_of_slot, uint8_t* __restrict__ rank) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) rank[i] = rank_of_slot[slot_of_point[i]];
}
// d_labels: n packed labels on the device.  Fills c->class_labels / *sizes (first-appearance order) and *d_rank_out.
static sicp_status classify_device(sicp_cloud* c, const uint32_t* d_labels, size_t n, uint8_t** d_rank_out, std::vector<int>* sizes, cudaStream_t st) {
  *d_rank_out = nullptr;
  if (n == 0) return SICP_OK;
  ClassTable* d_tab = nullptr; uint16_t* d_slot = nullptr; uint8_t* d_ros = nullptr;
  PinnedBlock pb{nullptr, 0, nullptr};
  auto body = [&]() -> sicp_status {
    SICP_CUDA(cudaMallocAsync(&d_tab, sizeof(ClassTable), st));
    SICP_CUDA(cudaMallocAsync(&d_slot, sizeof(uint16_t) * n, st));
    SICP_CUDA(cudaMallocAsync(&d_ros, kClassSlots, st));
    SICP_CUDA(cudaMallocAsync(d_rank_out, n, st));
    SICP_CUDA(cudaMemsetAsync(d_tab, 0, sizeof(ClassTable), st));
    SICP_CUDA(cudaMemsetAsync(d_tab->first, 0xff, sizeof(unsigned) * kClassSlots, st));
    classify_count_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_labels, (int)n, d_tab, d_slot);
    char* stage = (char*)pinned_stage(sizeof(ClassTable) + kClassSlots, st, &pb);
    if (!stage) { set_error("pinned staging allocation failed"); return SICP_ERR_CUDA; }
    ClassTable* h = (ClassTable*)stage;
    uint8_t* h_ros = (uint8_t*)(stage + sizeof(ClassTable));
    SICP_CUDA(cudaMemcpyAsync(h, d_tab, sizeof(ClassTable), cudaMemcpyDeviceToHost, st));
    SICP_CUDA(cudaStreamSynchronize(st));
    SICP_REQUIRE(!h->overflow, "PER_CLASS clouds support at most 128 distinct labels");
    std::vector<int> used;
    for (int s = 0; s < kClassSlots; s++) if (h->key[s]) used.push_back(s);
    SICP_REQUIRE(used.size() <= 128, "PER_CLASS clouds support at most 128 distinct labels");
    std::sort(used.begin(), used.end(), [&](int x, int y) { return h->first[x] < h->first[y]; });
    std::memset(h_ros, 0, kClassSlots);
    for (size_t r = 0; r < used.size(); r++) {
      h_ros[used[r]] = (uint8_t)r;
      c->class_labels.push_back((uint32_t)(h->key[used[r]] & 0xffffffffull));
      sizes->push_back((int)h->count[used[r]]);
    }
    SICP_CUDA(cudaMemcpyAsync(d_ros, h_ros, kClassSlots, cudaMemcpyHostToDevice, st));
    classify_rank_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_slot, (int)n, d_ros, *d_rank_out);
    count_launches(2);
    SICP_CUDA(cudaGetLastError());
    return SICP_OK;
  };
  const sicp_status rc = body();
  if (pb.p) pinned_release(pb, st);
  if (d_tab) cudaFreeAsync(d_tab, st);
  if (d_slot) cudaFreeAsync(d_slot, st);
  if (d_ros) cudaFreeAsync(d_ros, st);
  if (rc != SICP_OK && *d_rank_out) { cudaFreeAsync(*d_rank_out, st); *d_rank_out = nullptr; }
  return rc;
}

static sicp_status init_device(int device) {
  static std::mutex mu;
  static bool done[64] = {false};
  int cnt = 0;
  if (cudaGetDeviceCount(&cnt) != cudaSuccess || cnt == 0) {
    set_error("no CUDA device available (libsicp_b200 has no CPU fallback)");
    return SICP_ERR_CUDA;
  }
  SICP_REQUIRE(device >= 0 && device < cnt && device < 64, "device index out of range");
  SICP_CUDA(cudaSetDevice(device));
  std::lock_guard<std::mutex> lk(mu);
  if (!done[device]) {
    cudaMemPool_t pool;
    SICP_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
    uint64_t thr = UINT64_MAX;
    SICP_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));  // keep freed blocks cached
    done[device] = true;
  }
  return SICP_OK;
}

}  // namespace sicp

using namespace sicp;

sicp::CloudView sicp_cloud::view() const {
  CloudView v;
  v.n = (int)n; v.nslots = nslots; v.nseg = nseg;
  v.pts = d_pts; v.label = d_label; v.seg_of_leaf = d_seg_of_leaf; v.seg = d_seg;
  v.node_lo = d_node_lo; v.node_hi = d_node_hi; v.leaf_key = d_leaf_key; v.bb = d_bb;
  v.nrm = d_nrm; v.avec = d_avec; v.N = pre_N;
  return v;
}

extern "C" {

const char* sicp_last_error(void) { return g_err.c_str(); }
const char* sicp_version(void) { return "sicp_b200 0.1 (sm_100a)"; }
sicp_status sicp_device_count(int* count) {
  SICP_REQUIRE(count, "count is null");
  int c = 0;
  if (cudaGetDeviceCount(&c) != cudaSuccess) c = 0;
  *count = c;
  return SICP_OK;
}
uint64_t sicp_launch_count(void) { return g_launches; }
sicp_status sicp_set_stream(void* s) { g_stream = (cudaStream_t)s; return SICP_OK; }

// d_xyz / d_labels are staging buffers OWNED by this call (allocated stream-ordered on the current stream): they are
// freed after an eager build, or kept by the cloud until its deferred build.
static sicp_status create_common(float* d_xyz, uint32_t* d_labels, size_t n, int layout, int device, sicp_cloud** out) {
  cudaStream_t st = current_stream();
  sicp_cloud* c = new sicp_cloud();
  c->device = device; c->layout = layout; c->n = n; c->has_labels = d_labels != nullptr;
  std::vector<int> sizes;
  uint8_t* d_rank = nullptr;
  sicp_status rc = SICP_OK;
  if (layout == SICP_CLOUD_PER_CLASS) rc = classify_device(c, d_labels, n, &d_rank, &sizes, st);
  else sizes.push_back((int)n);
  // The label range (EM-ICP needs labels in 1..N, em_icp.hpp:301) is computed on the device by the build for host- and
  // device-created clouds alike and validated by whoever needs it (knn_cov.cu: precompute_cloud; register.cu:
  // Job::check_labels); a host-side scan here cost 20-80 us per 120k-point cloud on the caller's thread.
  if (rc == SICP_OK) layout_cloud(c, sizes);
  static const bool eager = [] { const char* e = getenv("SICP_EAGER_BUILD"); return e && *e && *e != '0'; }();
  if (rc == SICP_OK && cudaEventCreateWithFlags(&c->built_ev, cudaEventDisableTiming) != cudaSuccess) { set_error("event creation failed"); rc = SICP_ERR_CUDA; }
  if (rc == SICP_OK && layout == SICP_CLOUD_WHOLE && !eager && n > 0) {
    // deferred: keep the staged inputs, build on the first consumer's stream (ensure_built)
    if (cudaEventCreateWithFlags(&c->staged_ev, cudaEventDisableTiming) != cudaSuccess || cudaEventRecord(c->staged_ev, st) != cudaSuccess) {
      set_error("event creation failed"); rc = SICP_ERR_CUDA;
    } else {
      c->stage_xyz = d_xyz; c->stage_lab = d_labels; c->pending_build = true;
      d_xyz = nullptr; d_labels = nullptr;
    }
  } else if (rc == SICP_OK) {
    rc = build_cloud(c, d_xyz, d_labels, d_rank, st);
    if (rc == SICP_OK && cudaEventRecord(c->built_ev, st) != cudaSuccess) { set_error("event record failed"); rc = SICP_ERR_CUDA; }
  }
  if (d_rank) cudaFreeAsync(d_rank, st);
  if (d_xyz) cudaFreeAsync(d_xyz, st);
  if (d_labels) cudaFreeAsync(d_labels, st);
  if (rc != SICP_OK) { sicp_cloud_destroy(c); return rc; }
  *out = c;
  return SICP_OK;
}

}  // extern "C"

sicp_status sicp::ensure_built(const sicp_cloud* cc, cudaStream_t st) {
  sicp_cloud* c = const_cast<sicp_cloud*>(cc);
  std::lock_guard<std::mutex> lk(c->build_mu);
  if (c->pending_build) {
    SICP_CUDA(cudaSetDevice(c->device));
    SICP_CUDA(cudaStreamWaitEvent(st, c->staged_ev, 0));
    SICP_CHECK(build_cloud(c, c->stage_xyz, c->stage_lab, nullptr, st));
    SICP_CUDA(cudaFreeAsync(c->stage_xyz, st));
    if (c->stage_lab) SICP_CUDA(cudaFreeAsync(c->stage_lab, st));
    c->stage_xyz = nullptr; c->stage_lab = nullptr;
    SICP_CUDA(cudaEventRecord(c->built_ev, st));
    c->pending_build = false;
    return SICP_OK;
  }
  if (c->built_ev) SICP_CUDA(cudaStreamWaitEvent(st, c->built_ev, 0));
  return SICP_OK;
}
sicp_status sicp::ensure_ready(const sicp_cloud* c, cudaStream_t st) {
  SICP_CHECK(ensure_built(c, st));
  if (c->ready_ev) SICP_CUDA(cudaStreamWaitEvent(st, c->ready_ev, 0));
  return SICP_OK;
}

extern "C" {

sicp_status sicp_cloud_create(const void* xyz, size_t xyz_stride, const void* labels, size_t label_stride, size_t n, int layout,
                              int device, sicp_cloud** out) {
  SICP_REQUIRE(out, "out is null");
  SICP_REQUIRE(xyz || n == 0, "xyz is null");
  SICP_REQUIRE(xyz_stride >= 12 && (labels == nullptr || label_stride >= 4), "stride too small");
  SICP_REQUIRE(n <= kMaxCloudPoints, "too many points (at most 2^26 per cloud)");
  SICP_REQUIRE(layout == SICP_CLOUD_WHOLE || layout == SICP_CLOUD_PER_CLASS, "bad layout");
  SICP_REQUIRE(layout == SICP_CLOUD_WHOLE || labels, "PER_CLASS layout needs labels");
  SICP_CHECK(init_device(device));
  cudaStream_t st = current_stream();
  float* d_xyz = nullptr; uint32_t* d_lab = nullptr;
  SICP_CUDA(cudaMallocAsync(&d_xyz, std::max<size_t>(1, n) * 12, st));
  if (n) {
    if (xyz_stride == 12) SICP_CUDA(cudaMemcpyAsync(d_xyz, xyz, 12 * n, cudaMemcpyHostToDevice, st));
    else SICP_CUDA(cudaMemcpy2DAsync(d_xyz, 12, xyz, xyz_stride, 12, n, cudaMemcpyHostToDevice, st));
  }
  if (labels) {
    SICP_CUDA(cudaMallocAsync(&d_lab, std::max<size_t>(1, n) * 4, st));
    if (n) {
      if (label_stride == 4) SICP_CUDA(cudaMemcpyAsync(d_lab, labels, 4 * n, cudaMemcpyHostToDevice, st));
      else SICP_CUDA(cudaMemcpy2DAsync(d_lab, 4, labels, label_stride, 4, n, cudaMemcpyHostToDevice, st));
    }
  }
  return create_common(d_xyz, d_lab, n, layout, device, out);  // takes the staging buffers over
}

sicp_status sicp_cloud_create_device(const float* d_xyz, const uint32_t* d_labels, size_t n, int layout, int device, sicp_cloud** out) {
  SICP_REQUIRE(out, "out is null");
  SICP_REQUIRE(d_xyz || n == 0, "d_xyz is null");
  SICP_REQUIRE(n <= kMaxCloudPoints, "too many points (at most 2^26 per cloud)");
  SICP_REQUIRE(layout == SICP_CLOUD_WHOLE || layout == SICP_CLOUD_PER_CLASS, "bad layout");
  SICP_REQUIRE(layout == SICP_CLOUD_WHOLE || d_labels, "PER_CLASS layout needs labels");
  SICP_CHECK(init_device(device));
  // the caller's buffers are only read during this call: stage owned copies (device to device) for the (possibly deferred) build
  cudaStream_t st = current_stream();
  float* s_xyz = nullptr; uint32_t* s_lab = nullptr;
  SICP_CUDA(cudaMallocAsync(&s_xyz, std::max<size_t>(1, n) * 12, st));
  if (n) SICP_CUDA(cudaMemcpyAsync(s_xyz, d_xyz, 12 * n, cudaMemcpyDeviceToDevice, st));
  if (d_labels) {
    SICP_CUDA(cudaMallocAsync(&s_lab, std::max<size_t>(1, n) * 4, st));
    if (n) SICP_CUDA(cudaMemcpyAsync(s_lab, d_labels, 4 * n, cudaMemcpyDeviceToDevice, st));
  }
  return create_common(s_xyz, s_lab, n, layout, device, out);
}

void sicp_cloud_destroy(sicp_cloud* c) {
  if (!c) return;
  cudaStream_t st = current_stream();
  cudaSetDevice(c->device);
  void* bufs[] = {c->d_slab, c->d_nrm, c->d_avec, c->stage_xyz, c->stage_lab};
  for (void* b : bufs) if (b) cudaFreeAsync(b, st);
  if (c->ready_ev) cudaEventDestroy(c->ready_ev);
  if (c->built_ev) cudaEventDestroy(c->built_ev);
  if (c->staged_ev) cudaEventDestroy(c->staged_ev);
  delete c;
}

sicp_status sicp_cloud_size(const sicp_cloud* c, size_t* n) {
  SICP_REQUIRE(c && n, "null argument");
  *n = c->n;
  return SICP_OK;
}

sicp_status sicp_cloud_get_classes(const sicp_cloud* c, uint32_t* labels_out, int32_t* sizes_out, int* n_inout) {
  SICP_REQUIRE(c && n_inout, "null argument");
  SICP_REQUIRE(c->layout == SICP_CLOUD_PER_CLASS, "cloud is not PER_CLASS");
  int cap = *n_inout;
  *n_inout = c->nseg;
  for (int s = 0; s < c->nseg && s < cap; s++) {
    if (labels_out) labels_out[s] = c->class_labels[s];
    if (sizes_out) sizes_out[s] = c->h_seg[s].n;
  }
  return SICP_OK;
}

sicp_status sicp_cloud_transform_f32(const sicp_cloud* c, const double* pose7, void* out_xyz, size_t out_stride) {
  // Matrix4f path of the finalisation (gicp.hpp:166-171, em_icp.hpp:192-197, semantic_point_cloud.hpp:105-111): the pose
  // is cast to a float 4x4 and applied in float arithmetic, on the device, to the points in the caller's original order.
  SICP_REQUIRE(c && pose7 && out_xyz && out_stride >= 12, "bad argument");
  SICP_CHECK(validate_pose7(pose7, "sicp_cloud_transform_f32"));
  cudaStream_t st = current_stream();
  SICP_CUDA(cudaSetDevice(c->device));
  if (c->n == 0) return SICP_OK;
  SICP_CHECK(ensure_built(c, st));
  double Rd[9];
  quat_to_R(pose7, Rd);
  Mat34f M;
  for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) M.m[4 * i + j] = (float)Rd[3 * i + j]; M.m[4 * i + 3] = (float)pose7[4 + i]; }
  std::vector<float> h(3 * c->n);
  float* d_tmp;
  SICP_CUDA(cudaMallocAsync(&d_tmp, c->n * 12, st));
  transform_f32_kernel<<<(c->nslots + 255) / 256, 256, 0, st>>>(c->d_pts, c->nslots, M, d_tmp);
  count_launches(1);
  SICP_CUDA(cudaGetLastError());
  SICP_CUDA(cudaMemcpyAsync(h.data(), d_tmp, c->n * 12, cudaMemcpyDeviceToHost, st));
  SICP_CUDA(cudaStreamSynchronize(st));
  SICP_CUDA(cudaFreeAsync(d_tmp, st));
  if (out_stride == 12) std::memcpy(out_xyz, h.data(), c->n * 12);
  else for (size_t i = 0; i < c->n; i++) std::memcpy((char*)out_xyz + i * out_stride, &h[3 * i], 12);
  return SICP_OK;
}

}  // extern "C"

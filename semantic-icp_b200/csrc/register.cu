// register.cu — host orchestration of align(): covariance precompute, then the outer loop
// [transform + kNN] -> [E-step] -> [cooperative LM solve + convergence test] with every piece of loop state on the device.
// Replaces the bodies of GICP::align (impl/gicp.hpp:29-175), SemanticIterativeClosestPoint::align
// (impl/semantic_icp.hpp:27-166) and EmIterativeClosestPoint::align (impl/em_icp.hpp:24-200).
//
// The outer loop itself runs on the DEVICE: the three kernels of a pass are the body of a CUDA-graph WHILE node whose
// condition the LM kernel sets from the reference's stopping rule (cudaGraphSetConditional), so a registration is ONE
// graph launch and ONE read of the 2 KB control block at the end — no host round trip between passes.  Registrations
// run in SLOTS: a slot owns a stream, a persistent workspace (no allocation per registration) and its graph exec, and the
// batch executor keeps `max_concurrent` slots busy; the host sleeps on a completion queue fed by stream callbacks.
// The pass-by-pass path (kernels enqueued in chunks, control block read back between chunks) remains for
// options.profile (per-stage CUDA events cannot live inside a graph) and as the fallback when SICP_GRAPH=0.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>
#include "common.cuh"
#include "kernels.h"

namespace sicp {

// Batch defaults (tuned on one B200; the environment variables of the same purpose override them for
// tuning runs): registrations in flight, LM kernel shape and LM grid of a solve that shares the GPU.
constexpr int kBatchConcurrent = 8;
constexpr int kBatchLmVariant = 3;  // 256-thread CTAs held to 160 registers: the searches and E-steps of the other registrations in flight
                                    // run beside a solve on its SMs (tools/sweep.py, 32 pairs: 399 -> 423 registrations/s; 176 / 144 registers: 420)
constexpr int kBatchLmGrid = 37;
constexpr int kLoneCtlShare8 = 8, kBatchCtlShare8 = 8;  // sweep share of the LM controller block (eighths)
// Pairing (SICP_PAIR=1): registrations of a batch run in pairs that share one LM launch per pass, whose blocks alternate
// between the two solves so that each control step overlaps the other's sweep.  Measured on one B200 (tools/sweep.py, 16
// KITTI pairs): 388 registrations/s paired vs 390 unpaired, pass loop alone 20.8 vs 20.3 ms — in a batch the control gaps
// are already filled by the other registrations in flight, and the lockstep costs what the overlap gains.  Off by default.
constexpr int kPairDefault = 0;
constexpr int kBatchLmPairGrid = 74;  // CTAs of a paired solve in a batch (2 controllers + 72 sweeping blocks)
constexpr int kGraphReuse = 1;  // 1: keep one graph exec per slot and update it in place for every registration

static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  if (!e || !*e) return dflt;
  return atoi(e);
}

static LMConfig make_cfg(int algo, const sicp_options& o) {
  LMConfig c;
  c.algo = algo;
  c.kc = algo == SICP_ALGO_EM ? 4 : 1;                      // em_icp.hpp:60 / gicp.hpp:69, semantic_icp.hpp:68
  c.eps = o.epsilon;
  c.kappa = 1.0 - o.epsilon; c.hk = 0.5 * c.kappa; c.k4 = 0.25 * c.kappa; c.aa = 1.0 - c.hk;
  c.max_iter = o.max_lm_iterations;
  c.mse_stop = algo == SICP_ALGO_SEMANTIC ? 1e-3 : 1e-5;    // semantic_icp.hpp:152 / gicp.hpp:154, em_icp.hpp:180
  c.outer_cap = algo == SICP_ALGO_SEMANTIC ? 35 : 50;
  c.variant = 0;
  c.ctl_share8 = 8;
  return c;
}

static sicp_status validate(int algo, const sicp_cloud* src, const sicp_cloud* tgt, const sicp_options* o) {
  SICP_REQUIRE(src && tgt && o, "null argument");
  SICP_REQUIRE(algo == SICP_ALGO_GICP || algo == SICP_ALGO_SEMANTIC || algo == SICP_ALGO_EM, "unknown algorithm");
  SICP_REQUIRE(src->device == tgt->device, "clouds live on different devices");
  if (algo == SICP_ALGO_SEMANTIC) SICP_REQUIRE(src->layout == SICP_CLOUD_PER_CLASS && tgt->layout == SICP_CLOUD_PER_CLASS, "SEMANTIC needs PER_CLASS clouds");
  else SICP_REQUIRE(src->layout == SICP_CLOUD_WHOLE && tgt->layout == SICP_CLOUD_WHOLE, "GICP/EM need WHOLE clouds");
  if (algo == SICP_ALGO_EM) {
    SICP_REQUIRE(o->n_classes >= 1 && o->n_classes <= kMaxClasses && o->confusion, "EM needs n_classes in 1..64 and a confusion matrix");
    SICP_REQUIRE(src->has_labels && tgt->has_labels, "EM needs labelled clouds");
  }
  SICP_REQUIRE(o->k_cov >= 1 && o->k_cov <= kMaxK, "k_cov must be in 1..32");
  SICP_REQUIRE((long long)src->nslots * 4 < (1ll << 31), "source cloud too large for one registration (nslots * 4 must fit 31 bits)");
  return SICP_OK;
}

sicp_status validate_pose7(const double* p7, const char* what, bool nullable) {
  if (!p7) {
    if (nullable) return SICP_OK;
    set_error(std::string(what) + ": pose7 is null");
    return SICP_ERR_INVALID;
  }
  double n2 = 0;
  bool finite = true;
  for (int i = 0; i < 7; i++) finite = finite && std::isfinite(p7[i]);
  for (int i = 0; i < 4; i++) n2 += p7[i] * p7[i];
  if (!finite || std::fabs(n2 - 1.0) > 1e-6) { set_error(std::string(what) + ": pose7 must be finite with a unit quaternion [qx,qy,qz,qw]"); return SICP_ERR_INVALID; }
  return SICP_OK;
}

// Covariances / label vectors of both clouds of a pair.  The two clouds are independent, and one covariance kernel is
// a single wave with a long tail (a few slow warps), so for a LONE registration the target runs on a helper stream
// beside the source on `st`; `st` then waits for both ready events.  In a batch (helper == nullptr) both run on `st`:
// the tails are already filled by the other registrations in flight.
static sicp_status precompute_pair(int algo, sicp_cloud* src, sicp_cloud* tgt, const sicp_options* o, cudaStream_t st, cudaStream_t helper,
                                   bool defer_label_check = false) {
  const int N = algo == SICP_ALGO_EM ? o->n_classes : 0;
  cudaStream_t saved = current_stream();
  sicp_status rc = SICP_OK;
  if (tgt != src) {
    bool cached;
    { std::lock_guard<std::mutex> lk(tgt->mu); cached = tgt->pre_valid; }
    sicp_set_stream(helper && !cached ? helper : st);
    rc = precompute_cloud(tgt, o->k_cov, o->epsilon, N, o->confusion, defer_label_check);
  }
  if (rc == SICP_OK) {
    sicp_set_stream(st);
    rc = precompute_cloud(src, o->k_cov, o->epsilon, N, o->confusion, defer_label_check);
  }
  sicp_set_stream(saved);
  SICP_CHECK(rc);
  SICP_CHECK(ensure_ready(src, st));
  SICP_CHECK(ensure_ready(tgt, st));
  return SICP_OK;
}

// ------------------------------------------------------------------ completion queue (stream callbacks -> host)
struct Completion {
  std::mutex mu;
  std::condition_variable cv;
  std::deque<int> q;
  void push(int v) { { std::lock_guard<std::mutex> lk(mu); q.push_back(v); } cv.notify_one(); }
  int pop() {
    std::unique_lock<std::mutex> lk(mu);
    cv.wait(lk, [&] { return !q.empty(); });
    const int v = q.front();
    q.pop_front();
    return v;
  }
};
struct SlotNote { Completion* c; int slot; };
static void CUDART_CB note_done(void* p) {  // runs on a CUDA-internal thread: no CUDA calls here
  SlotNote* n = static_cast<SlotNote*>(p);
  n->c->push(n->slot);
}

// ------------------------------------------------------------------ slots
// A slot is everything one registration in flight needs, kept across registrations and across calls (per host thread
// and device): stream, workspace sized for the largest source cloud seen, pinned control block, graph exec.
struct Slot {
  int device = -1;
  cudaStream_t st = nullptr;      // own stream (batch) — a lone registration runs on the caller's stream instead
  cudaStream_t cap = nullptr;     // capture-only stream for building the pass graph
  size_t cap_nc = 0;              // capacity in candidate records
  int* d_corr = nullptr; float* d_d2 = nullptr;
  char* d_rec = nullptr;         // record group blocks of the current pass (kernels.h: struct Rec), sized for k_c = 4
  RegCtl* d_ctl = nullptr; double* d_partials = nullptr; int* d_map = nullptr;
  RegCtl* h_ctl = nullptr;        // pinned
  int* h_map = nullptr;           // pinned, kMaxSegMap ints (class map staging) + 4 words: label ranges of the pair's clouds
  cudaGraphExec_t exec = nullptr;       // WHILE { kNN, E-step, LM } of one registration
  cudaGraphExec_t exec_pair = nullptr;  // ... of a pair of registrations (two kNN + E-step chains, one paired LM launch)
  SlotNote note{nullptr, 0};
  static constexpr int kMaxSegMap = 256;

  sicp_status ensure(int dev, size_t nc) {
    if (device != dev) { destroy(); device = dev; }
    SICP_CUDA(cudaSetDevice(dev));
    if (!st) SICP_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    if (!cap) SICP_CUDA(cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking));
    if (!d_ctl) {
      SICP_CUDA(cudaMalloc(&d_ctl, sizeof(RegCtl)));
      SICP_CUDA(cudaMalloc(&d_partials, sizeof(double) * kLmPartialsDoubles));
      SICP_CUDA(cudaMemset(d_partials, 0, sizeof(double) * kLmPartialsDoubles));  // LMSync starts at zero; generations continue from there
      SICP_CUDA(cudaMalloc(&d_map, sizeof(int) * kMaxSegMap));
      SICP_CUDA(cudaMallocHost(&h_ctl, sizeof(RegCtl)));
      SICP_CUDA(cudaMallocHost(&h_map, sizeof(int) * (kMaxSegMap + 4)));
    }
    nc = std::max<size_t>(nc, 1);
    if (nc > cap_nc) {
      // growing is rare (largest source cloud seen so far); the old buffers may still be referenced by work in flight
      SICP_CUDA(cudaDeviceSynchronize());
      free_records();
      const size_t want = nc + nc / 8;
      SICP_CUDA(cudaMalloc(&d_corr, sizeof(int) * want));
      SICP_CUDA(cudaMalloc(&d_d2, sizeof(float) * want));
      SICP_CUDA(cudaMalloc(&d_rec, (want / 32 + 1) * (size_t)Rec::group_bytes(4)));  // want >= nslots * k_c: enough groups for either k_c
      cap_nc = want;
    }
    return SICP_OK;
  }
  void free_records() {
    cudaFree(d_corr); cudaFree(d_d2); cudaFree(d_rec);
    d_corr = nullptr; d_d2 = nullptr; d_rec = nullptr; cap_nc = 0;
  }
  void destroy() {
    if (device < 0) return;
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return; }  // context already gone (process exit)
    if (exec) cudaGraphExecDestroy(exec);
    if (exec_pair) cudaGraphExecDestroy(exec_pair);
    free_records();
    cudaFree(d_ctl); cudaFree(d_partials); cudaFree(d_map);
    if (h_ctl) cudaFreeHost(h_ctl);
    if (h_map) cudaFreeHost(h_map);
    if (st) cudaStreamDestroy(st);
    if (cap) cudaStreamDestroy(cap);
    cudaGetLastError();
    *this = Slot();
  }
};
struct SlotPool {
  std::vector<Slot*> slots;
  Slot* get(int i) {
    while ((int)slots.size() <= i) slots.push_back(new Slot());
    return slots[i];
  }
  ~SlotPool() { for (Slot* s : slots) { s->destroy(); delete s; } }
};
static thread_local SlotPool t_pool;
// second stream of a lone registration (target build + covariances beside the source's); owned per host thread and device
struct HelperStream {
  int device = -1;
  cudaStream_t st = nullptr;
  sicp_status ensure(int dev) {
    if (st && device == dev) return SICP_OK;
    release();
    SICP_CUDA(cudaSetDevice(dev));
    SICP_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    device = dev;
    return SICP_OK;
  }
  void release() {
    if (st && cudaSetDevice(device) == cudaSuccess) cudaStreamDestroy(st);
    cudaGetLastError();  // context already gone at process exit: nothing to free
    st = nullptr; device = -1;
  }
  ~HelperStream() { release(); }
};
static thread_local HelperStream t_helper;

// graph path switch: SICP_GRAPH=0 disables it; it also turns itself off for the process if the driver rejects the graph
static std::atomic<int> g_graph_ok{1};
static bool graph_enabled() { return g_graph_ok.load() == 1 && env_int("SICP_GRAPH", 1) != 0; }

struct StageTimer {
  bool on = false;
  std::vector<cudaEvent_t> ev;   // pairs
  std::vector<int> stage;
  void begin(int s, cudaStream_t st) { if (!on) return; cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); ev.push_back(e); stage.push_back(s); }
  void end(cudaStream_t st) { if (!on) return; cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); ev.push_back(e); }
  void collect(sicp_result* out, cudaEvent_t ref = nullptr, int job = 0) {
    FILE* tf = nullptr;  // SICP_TRACE=<file>: append "job stage start_ms end_ms" (relative to the batch start) per timed launch
    if (ref) { if (const char* path = getenv("SICP_TRACE")) tf = fopen(path, "a"); }
    for (size_t i = 0; i + 1 < ev.size(); i += 2) {
      float ms = 0;
      if (cudaEventElapsedTime(&ms, ev[i], ev[i + 1]) == cudaSuccess) { out->stage_ms[stage[i / 2]] += ms; out->stage_launches[stage[i / 2]]++; }
      if (tf) {
        float t0 = 0, t1 = 0;
        cudaEventElapsedTime(&t0, ref, ev[i]); cudaEventElapsedTime(&t1, ref, ev[i + 1]);
        fprintf(tf, "%d %d %.4f %.4f\n", job, stage[i / 2], t0, t1);
      }
    }
    if (tf) fclose(tf);
    for (cudaEvent_t e : ev) cudaEventDestroy(e);
    ev.clear(); stage.clear();
  }
};

// One registration as a resumable state machine so that several can be interleaved on different streams.
struct Job {
  int algo; sicp_cloud* src; sicp_cloud* tgt; const sicp_options* opts; LMConfig cfg; StageTimer tm;
  sicp_result* out; Slot* sl = nullptr; cudaStream_t st = nullptr; int lm_grid = 0;
  int enqueued = 0; bool finished = false; bool graphed = false; int d2h = 0;
  cudaEvent_t trace_ref = nullptr; int trace_id = 0;
  // Pairing (batches): the LEADER of a pair drives both registrations on its stream — their passes run in lockstep and
  // their inner solves share one launch whose blocks alternate between the two problems (lm.cu: lm_pair_kernel).  The
  // partner only contributes its slot's workspace and its own control block.
  Job* partner = nullptr;
  bool follower = false;
  int pair_grid = 0;
  Slot* idle_partner = nullptr;  // odd job of a paired batch: runs the paired kernel against an already-converged dummy, so that
                                 // every registration of the batch is summed in the same block order (bit-identical results)

  // class map of SemanticICP (semantic_icp.hpp:50-51) through the slot's pinned staging: no host synchronisation
  sicp_status stage_class_map() {
    if (cfg.algo != SICP_ALGO_SEMANTIC) return SICP_OK;
    SICP_REQUIRE(src->nseg <= Slot::kMaxSegMap, "too many classes");
    for (int s = 0; s < src->nseg; s++) {
      int m = -1;
      if (src->h_seg[s].n > opts->min_class_points)
        for (int t = 0; t < tgt->nseg; t++)
          if (tgt->class_labels[t] == src->class_labels[s]) { m = t; break; }
      sl->h_map[s] = m;
    }
    SICP_CUDA(cudaMemcpyAsync(sl->d_map, sl->h_map, sizeof(int) * std::max(1, src->nseg), cudaMemcpyHostToDevice, st));
    return SICP_OK;
  }
  sicp_status start(const double* init7) {
    SICP_CHECK(stage_class_map());
    std::memset(sl->h_ctl, 0, sizeof(RegCtl));
    std::memcpy(sl->h_ctl->pose, init7, 56);
    SICP_CUDA(cudaMemcpyAsync(sl->d_ctl, sl->h_ctl, sizeof(RegCtl), cudaMemcpyHostToDevice, st));
    return SICP_OK;
  }
  const int* class_map() const { return cfg.algo == SICP_ALGO_SEMANTIC ? sl->d_map : nullptr; }
  // EM labels must be 1..N (em_icp.hpp:301).  For clouds created from device labels the range is only known on the device:
  // it is read back with the registration (no host synchronisation before the passes) and checked when the job completes.
  bool labels_pending = false;
  sicp_status fetch_label_ranges() {
    if (algo != SICP_ALGO_EM) return SICP_OK;
    unsigned* h = reinterpret_cast<unsigned*>(sl->h_map + Slot::kMaxSegMap);
    const sicp_cloud* cl[2] = {src, tgt};
    for (int i = 0; i < 2; i++) {
      bool known;
      { std::lock_guard<std::mutex> lk(const_cast<sicp_cloud*>(cl[i])->mu); known = cl[i]->label_range_known || cl[i]->nslots == 0; }
      h[2 * i] = 1; h[2 * i + 1] = 1;  // "in range" unless the device says otherwise
      if (known) continue;
      SICP_CUDA(cudaMemcpyAsync(h + 2 * i, cl[i]->d_bb + 6, 8, cudaMemcpyDeviceToHost, st));
      labels_pending = true;
    }
    return SICP_OK;
  }
  sicp_status check_labels() {
    if (!labels_pending) return SICP_OK;
    const unsigned* h = reinterpret_cast<const unsigned*>(sl->h_map + Slot::kMaxSegMap);
    sicp_cloud* cl[2] = {src, tgt};
    for (int i = 0; i < 2; i++) {
      if (cl[i]->nslots == 0) continue;
      { std::lock_guard<std::mutex> lk(cl[i]->mu);
        if (!cl[i]->label_range_known) { cl[i]->min_label = h[2 * i]; cl[i]->max_label = h[2 * i + 1]; cl[i]->label_range_known = true; } }
      SICP_REQUIRE(cl[i]->min_label >= 1, "label 0 found: EM-ICP labels must be 1..N");
      SICP_REQUIRE((int)cl[i]->max_label <= opts->n_classes, "label exceeds n_classes: EM-ICP labels must be 1..N");
    }
    return SICP_OK;
  }
  // correspondences + E-step of one outer pass on stream `s`
  sicp_status enqueue_corr(cudaStream_t s) {
    const int* stop = &sl->d_ctl->converged;
    tm.begin(SICP_STAGE_KNN, s);
    SICP_CHECK(launch_cross_knn(src, tgt, sl->d_ctl->pose, stop, class_map(), cfg.kc, sl->d_corr, sl->d_d2, s));
    tm.end(s);
    tm.begin(SICP_STAGE_ESTEP, s);
    SICP_CHECK(launch_estep(src, tgt, cfg, opts->gate_d2, sl->d_ctl->pose, stop, sl->d_corr, sl->d_d2, sl->d_rec, sl->d_ctl, s));
    tm.end(s);
    return SICP_OK;
  }
  // the kernels of one outer pass on stream `s` (the slot's stream, or the capture stream of the graph build)
  sicp_status enqueue_pass(cudaStream_t s, unsigned long long cond) {
    SICP_CHECK(enqueue_corr(s));
    if (partner) {
      SICP_CHECK(partner->enqueue_corr(s));
      SICP_CHECK(launch_lm_pair(src, cfg, sl->d_rec, sl->d_ctl, sl->d_partials, partner->src, partner->cfg, partner->sl->d_rec, partner->sl->d_ctl,
                                partner->sl->d_partials, pair_grid, s, cond));
      return SICP_OK;
    }
    if (idle_partner) {
      SICP_CHECK(launch_lm_pair(src, cfg, sl->d_rec, sl->d_ctl, sl->d_partials, src, cfg, sl->d_rec, idle_partner->d_ctl, idle_partner->d_partials, pair_grid, s, cond));
      return SICP_OK;
    }
    tm.begin(SICP_STAGE_LM, s);
    SICP_CHECK(launch_lm(src, cfg, sl->d_rec, sl->d_ctl, sl->d_partials, lm_grid, s, cond));
    tm.end(s);
    return SICP_OK;
  }
  // ---- device-resident outer loop: WHILE(not converged) { kNN, E-step, LM } as one graph launch
  sicp_status launch_graph() {
    cudaGraph_t g = nullptr;
    SICP_CUDA(cudaGraphCreate(&g, 0));
    auto build = [&]() -> sicp_status {
      cudaGraphConditionalHandle h;
      SICP_CUDA(cudaGraphConditionalHandleCreate(&h, g, 1, cudaGraphCondAssignDefault));  // every launch starts with "run a pass"
      cudaGraphNodeParams cp = {cudaGraphNodeTypeConditional};
      cp.conditional.handle = h;
      cp.conditional.type = cudaGraphCondTypeWhile;
      cp.conditional.size = 1;
      cudaGraphNode_t node;
      SICP_CUDA(cudaGraphAddNode(&node, g, nullptr, 0, &cp));
      cudaGraph_t body = cp.conditional.phGraph_out[0];
      SICP_CUDA(cudaStreamBeginCaptureToGraph(sl->cap, body, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed));
      const sicp_status rc = enqueue_pass(sl->cap, (unsigned long long)h);
      cudaGraph_t ended = nullptr;
      const cudaError_t e = cudaStreamEndCapture(sl->cap, &ended);
      SICP_CHECK(rc);
      SICP_CUDA(e);
      const int reuse = env_int("SICP_GRAPH_REUSE", kGraphReuse);
      cudaGraphExec_t& ex = partner ? sl->exec_pair : sl->exec;
      if (ex && reuse) {  // same topology as the previous registration of this slot: update the kernel parameters in place
        cudaGraphExecUpdateResultInfo info;
        if (cudaGraphExecUpdate(ex, g, &info) != cudaSuccess) { cudaGetLastError(); cudaGraphExecDestroy(ex); ex = nullptr; }
      } else if (ex) {
        cudaGraphExecDestroy(ex);  // the previous registration of this slot has completed (its readback was consumed)
        ex = nullptr;
      }
      if (!ex) SICP_CUDA(cudaGraphInstantiate(&ex, g, 0));
      SICP_CUDA(cudaGraphLaunch(ex, st));
      return SICP_OK;
    };
    const sicp_status rc = build();
    cudaGraphDestroy(g);
    if (rc == SICP_OK) { graphed = true; enqueued = cfg.outer_cap + 2; }
    return rc;
  }
  // enqueue the rest of the registration (graph) or a chunk of passes (pass-by-pass path), then the control-block
  // readback and the completion callback
  sicp_status advance(int chunk) {
    bool done = false;
    if (!graphed && !tm.on && enqueued == 0 && graph_enabled()) {
      if (launch_graph() == SICP_OK) done = true;
      else {
        cudaGetLastError();
        if (g_graph_ok.exchange(0) == 1) fprintf(stderr, "[sicp] graph outer loop unavailable (%s); using the pass-by-pass path\n", sicp_last_error());
      }
    }
    if (!done) {
      const int cap = cfg.outer_cap + 2;
      for (int i = 0; i < chunk && enqueued < cap; i++) { SICP_CHECK(enqueue_pass(st, 0)); enqueued++; }
    }
    SICP_CUDA(cudaMemcpyAsync(sl->h_ctl, sl->d_ctl, sizeof(RegCtl), cudaMemcpyDeviceToHost, st));
    d2h += (int)sizeof(RegCtl);
    if (partner) {
      SICP_CUDA(cudaMemcpyAsync(partner->sl->h_ctl, partner->sl->d_ctl, sizeof(RegCtl), cudaMemcpyDeviceToHost, st));
      partner->d2h += (int)sizeof(RegCtl);
      partner->graphed = graphed;
    }
    if (sl->note.c) SICP_CUDA(cudaLaunchHostFunc(st, note_done, &sl->note));
    return SICP_OK;
  }
  // after the readback has landed: true when the registration is complete
  bool complete() const {
    return (sl->h_ctl->converged != 0 && (!partner || partner->sl->h_ctl->converged != 0)) || enqueued >= cfg.outer_cap + 2;
  }
  void finish() {
    const RegCtl& c = *sl->h_ctl;
    std::memcpy(out->pose7, c.pose, 56);
    out->outer_iter = c.outer; out->lm_iters_total = c.lm_iters_total; out->final_cost = c.final_cost; out->n_corr_last = c.n_corr_last;
    out->flags = c.flags; out->lm_evals_total = c.lm_evals_total;
    const int np = std::min(c.outer, 64);
    std::memcpy(out->pass_pose7, c.pass_pose, sizeof(double) * 7 * np);
    std::memcpy(out->pass_lm_iters, c.pass_lm_iters, sizeof(int) * np);
    out->d2h_bytes = d2h;
    for (int i = 0; i < 3; i++) out->lm_cycles[i] = (double)c.dbg_cycles[i];
    for (int i = 0; i < 3; i++) out->lm_cycles[3 + i] = (double)c.dbg_cycles[4 + i];
    out->gpu_launches = c.outer * 3;  // kernels that did work (the graph loop launches exactly these; the pass-by-pass path may add early-exit launches)
    if (graphed && partner) count_launches(5 * std::max(c.outer, partner->sl->h_ctl->outer) - 5);  // a pair's pass: 2 kNN + 2 E-step + 1 paired LM
    else if (graphed && !follower) count_launches(c.outer * 3 - 3);  // the capture counted one pass
    tm.collect(out, trace_ref, trace_id);
    finished = true;
  }
};

// Runs all jobs with up to `max_concurrent` registrations in flight, each in its own slot.  A slot that finishes
// immediately picks up the next job (no wave barrier), so the tail of one registration overlaps the bulk of another.
static sicp_status run_jobs(std::vector<Job>& jobs, const double* init7s, int max_concurrent) {
  cudaStream_t base = current_stream();
  const int nj = (int)jobs.size();
  const int S = std::max(1, std::min(max_concurrent, nj));
  const int device = jobs[0].src->device;
  const bool lone = S == 1;
  size_t nc_max = 1;
  for (const Job& jb : jobs) nc_max = std::max(nc_max, (size_t)jb.src->nslots * (jb.algo == SICP_ALGO_EM ? 4 : 1));
  Completion done_q;
  cudaEvent_t fork = nullptr;
  // Pairing: two registrations share a leader slot's stream and graph and one paired LM launch per pass (see Job::partner).
  bool pairing = !lone && nj >= 2 && env_int("SICP_PAIR", kPairDefault) != 0;
  for (const Job& jb : jobs) pairing = pairing && jb.algo == jobs[0].algo && jb.opts->profile == 0;
  const int NS = pairing ? std::max(1, S / 2) : S;  // leader slots in flight; a pair also uses the workspace of slot NS + s
  for (int s = 0; s < (pairing ? 2 * NS : NS); s++) {
    Slot* sl = t_pool.get(s);
    SICP_CHECK(sl->ensure(device, nc_max));
    sl->note = SlotNote{&done_q, s};
  }
  if (lone) SICP_CHECK(t_helper.ensure(device));
  if (!lone) {
    SICP_CUDA(cudaEventCreate(&fork));
    SICP_CUDA(cudaEventRecord(fork, base));
  }
  const int kChunk = std::max(1, env_int("SICP_CHUNK", 3));  // pass-by-pass path: passes enqueued between two readbacks
  // LM kernel shape.  A lone solve takes every SM (shape 0, one 256-thread CTA each at ~240 registers).  Concurrent solves
  // use the register-capped shape and a grid of a quarter of the machine: their sweeps are longer, so the latency-bound
  // control step between sweeps idles a smaller share of the SMs, and the kNN / covariance / E-step kernels of the other
  // registrations in flight fit beside an LM CTA on the same SM (it leaves 24k registers and 55 KB of shared memory).
  const int variant = lone ? 0 : std::min(std::max(env_int("SICP_LM_VARIANT", kBatchLmVariant), 0), kLmVariants - 1);
  sicp_status rc = SICP_OK;
  std::vector<int> slot_job(NS, -1);
  int next = 0, live = 0;
  // everything a registration needs before its passes, on stream `st` with the workspace of `sl`
  auto prepare = [&](int j, Slot* sl, cudaStream_t st) -> sicp_status {
    Job& jb = jobs[j];
    jb.sl = sl;
    jb.st = st;
    jb.cfg = make_cfg(jb.algo, *jb.opts);
    jb.cfg.variant = variant;
    jb.cfg.ctl_share8 = std::min(8, std::max(0, env_int("SICP_LM_CTL_SHARE", lone ? kLoneCtlShare8 : kBatchCtlShare8)));
    const int gmax = lm_max_grid(device, jb.algo, variant);
    jb.lm_grid = lone ? lm_grid_blocks(device) : std::min(gmax, std::max(1, env_int("SICP_LM_GRID", kBatchLmGrid)));
    jb.tm.on = jb.opts->profile != 0;
    jb.trace_ref = fork; jb.trace_id = j;
    // the clouds may still have to be built (deferred), or are building on the stream / host thread that created them
    SICP_CHECK(ensure_built(jb.src, st));
    SICP_CHECK(ensure_built(jb.tgt, st));
    jb.tm.begin(SICP_STAGE_COV, st);
    const sicp_status r = precompute_pair(jb.algo, jb.src, jb.tgt, jb.opts, st, lone ? t_helper.st : nullptr, true);
    jb.tm.end(st);
    SICP_CHECK(r);
    SICP_CHECK(jb.fetch_label_ranges());
    return jb.start(init7s + 7 * (size_t)j);
  };
  auto launch = [&](int slot) -> sicp_status {
    const int j = next++;
    slot_job[slot] = j;
    Slot* sl = t_pool.get(slot);
    cudaStream_t st = lone ? base : sl->st;
    SICP_CHECK(prepare(j, sl, st));
    Job& jb = jobs[j];
    if (pairing && next < nj) {  // take a partner: same stream, its own workspace
      const int j2 = next++;
      SICP_CHECK(prepare(j2, t_pool.get(NS + slot), st));
      jobs[j2].follower = true;
      jb.partner = &jobs[j2];
      jb.pair_grid = std::max(3, env_int("SICP_LM_PAIR_GRID", kBatchLmPairGrid));
    } else if (pairing) {  // odd one out: same kernel, partner slot marked converged
      Slot* idle = t_pool.get(NS + slot);
      std::memset(idle->h_ctl, 0, sizeof(RegCtl));
      idle->h_ctl->pose[3] = 1.0;
      idle->h_ctl->converged = 1;
      SICP_CUDA(cudaMemcpyAsync(idle->d_ctl, idle->h_ctl, sizeof(RegCtl), cudaMemcpyHostToDevice, st));
      jb.idle_partner = idle;
      jb.pair_grid = std::max(3, env_int("SICP_LM_PAIR_GRID", kBatchLmPairGrid));
    }
    SICP_CHECK(jb.advance(kChunk));
    live++;
    return SICP_OK;
  };
  // SICP_HOSTSTAT=1: where the host thread's time goes (issuing work vs asleep waiting for a slot) — tuning aid
  const bool hoststat = env_int("SICP_HOSTSTAT", 0) != 0;
  auto now_ms = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  double t_issue = 0, t_wait = 0, t_begin = now_ms(), t0 = t_begin;
  for (int s = 0; s < NS && next < nj && rc == SICP_OK; s++) rc = launch(s);
  t_issue += now_ms() - t0;
  while (live > 0 && rc == SICP_OK) {
    t0 = now_ms();
    const int s = done_q.pop();  // sleeps until some slot's readback has landed
    t_wait += now_ms() - t0;
    t0 = now_ms();
    const int j = slot_job[s];
    if (j < 0) continue;
    if (jobs[j].complete()) {
      jobs[j].finish();
      if (jobs[j].partner) jobs[j].partner->finish();
      rc = jobs[j].check_labels();
      if (rc == SICP_OK && jobs[j].partner) rc = jobs[j].partner->check_labels();
      live--;
      slot_job[s] = -1;
      if (rc == SICP_OK && next < nj) rc = launch(s);
    } else {
      rc = jobs[j].advance(kChunk);
    }
    t_issue += now_ms() - t0;
  }
  if (hoststat)
    fprintf(stderr, "[sicp] run_jobs: %d jobs, %d slots, %.2f ms wall: host issuing %.2f ms, asleep %.2f ms\n", nj, S, now_ms() - t_begin, t_issue, t_wait);
  if (rc != SICP_OK) cudaDeviceSynchronize();  // error path: nothing may still be writing a pinned control block
  for (int s = 0; s < (pairing ? 2 * NS : NS); s++) t_pool.get(s)->note.c = nullptr;
  if (!lone) {
    for (int s = 0; s < NS; s++) {  // join: later work on the caller's stream sees the results of every slot
      cudaEvent_t e;
      cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
      cudaEventRecord(e, t_pool.get(s)->st);
      cudaStreamWaitEvent(base, e, 0);
      cudaEventDestroy(e);
    }
    cudaEventDestroy(fork);
  }
  return rc;
}

// a slot for the synchronous single-pass entry points below (no registration of this thread is in flight when they run)
static sicp_status scratch_slot(const sicp_cloud* src, int kc, Slot** out) {
  Slot* sl = t_pool.get(0);
  SICP_CHECK(sl->ensure(src->device, (size_t)src->nslots * kc));
  sl->note.c = nullptr;
  *out = sl;
  return SICP_OK;
}

}  // namespace sicp

using namespace sicp;

extern "C" {

void sicp_options_default(int algo, sicp_options* o) {
  if (!o) return;
  std::memset(o, 0, sizeof *o);
  o->k_cov = 20;            // gicp.h:34, em_icp.h:42, semantic_point_cloud.h:31
  o->epsilon = 0.001;
  o->gate_d2 = 250.0;       // gicp.hpp:70
  o->min_class_points = 400;  // semantic_icp.hpp:51
  o->max_lm_iterations = 400; // gicp.hpp:143
  (void)algo;
}

sicp_status sicp_register(int algo, sicp_cloud* src, sicp_cloud* tgt, const sicp_options* opts, const double* init7, sicp_result* out) {
  SICP_REQUIRE(init7 && out, "null argument");
  SICP_CHECK(validate(algo, src, tgt, opts));
  SICP_CHECK(validate_pose7(init7, "sicp_register"));
  SICP_CUDA(cudaSetDevice(src->device));
  std::memset(out, 0, sizeof *out);
  std::vector<Job> jobs(1);
  jobs[0].algo = algo; jobs[0].src = src; jobs[0].tgt = tgt; jobs[0].opts = opts; jobs[0].out = out;
  return run_jobs(jobs, init7, 1);
}

sicp_status sicp_register_batch(int algo, size_t n_pairs, sicp_cloud* const* src, sicp_cloud* const* tgt, const sicp_options* opts,
                                const double* init7s, sicp_result* out) {
  SICP_REQUIRE(src && tgt && init7s && out, "null argument");
  if (n_pairs == 0) return SICP_OK;
  for (size_t i = 0; i < n_pairs; i++) {
    SICP_CHECK(validate(algo, src[i], tgt[i], opts));
    SICP_REQUIRE(src[i]->device == src[0]->device, "all pairs of a batch must live on one device");
    SICP_CHECK(validate_pose7(init7s + 7 * i, "sicp_register_batch"));
  }
  SICP_CUDA(cudaSetDevice(src[0]->device));
  std::memset(out, 0, sizeof(sicp_result) * n_pairs);
  std::vector<Job> jobs(n_pairs);
  for (size_t i = 0; i < n_pairs; i++) { jobs[i].algo = algo; jobs[i].src = src[i]; jobs[i].tgt = tgt[i]; jobs[i].opts = opts; jobs[i].out = out + i; }
  int conc = opts->max_concurrent > 0 ? opts->max_concurrent : env_int("SICP_CONCURRENT", kBatchConcurrent);
  conc = std::max(2, std::min(conc, 64));  // >= 2: a batch always runs in slots on their own streams
  return run_jobs(jobs, init7s, conc);
}

sicp_status sicp_correspondences(int algo, sicp_cloud* src, sicp_cloud* tgt, const sicp_options* opts, const double* pose7, int32_t* idx_out,
                                 double* w_out, float* d2_out) {
  SICP_REQUIRE(pose7 && idx_out, "null argument");
  SICP_CHECK(validate(algo, src, tgt, opts));
  SICP_CHECK(validate_pose7(pose7, "sicp_correspondences"));
  SICP_CUDA(cudaSetDevice(src->device));
  cudaStream_t st = current_stream();
  SICP_CHECK(ensure_built(src, st));
  SICP_CHECK(ensure_built(tgt, st));
  SICP_CHECK(precompute_pair(algo, src, tgt, opts, st, nullptr));
  Job jb;
  jb.algo = algo; jb.src = src; jb.tgt = tgt; jb.opts = opts; jb.cfg = make_cfg(algo, *opts); jb.st = st;
  SICP_CHECK(scratch_slot(src, jb.cfg.kc, &jb.sl));
  Slot* ws = jb.sl;
  LMConfig& cfg = jb.cfg;
  const size_t nslot_c = (size_t)src->nslots * cfg.kc;
  const size_t rec_bytes = (size_t)(src->nslots / 32) * Rec::group_bytes(cfg.kc);
  std::vector<int> h_corr(nslot_c); std::vector<float> h_d2(nslot_c); std::vector<char> h_rec(rec_bytes);
  std::vector<float4> h_spts(src->nslots), h_tpts(tgt->nslots);
  SICP_CHECK(jb.start(pose7));
  SICP_CHECK(launch_cross_knn(src, tgt, ws->d_ctl->pose, nullptr, jb.class_map(), cfg.kc, ws->d_corr, ws->d_d2, st));
  SICP_CHECK(launch_estep(src, tgt, cfg, opts->gate_d2, ws->d_ctl->pose, nullptr, ws->d_corr, ws->d_d2, ws->d_rec, nullptr, st));
  SICP_CUDA(cudaMemcpyAsync(h_corr.data(), ws->d_corr, sizeof(int) * nslot_c, cudaMemcpyDeviceToHost, st));
  SICP_CUDA(cudaMemcpyAsync(h_d2.data(), ws->d_d2, sizeof(float) * nslot_c, cudaMemcpyDeviceToHost, st));
  SICP_CUDA(cudaMemcpyAsync(h_rec.data(), ws->d_rec, rec_bytes, cudaMemcpyDeviceToHost, st));
  SICP_CUDA(cudaMemcpyAsync(h_spts.data(), src->d_pts, sizeof(float4) * src->nslots, cudaMemcpyDeviceToHost, st));
  SICP_CUDA(cudaMemcpyAsync(h_tpts.data(), tgt->d_pts, sizeof(float4) * tgt->nslots, cudaMemcpyDeviceToHost, st));
  SICP_CUDA(cudaStreamSynchronize(st));
  for (int s = 0; s < src->nslots; s++) {
    int o; std::memcpy(&o, &h_spts[s].w, 4);
    if (o < 0) continue;
    for (int c = 0; c < cfg.kc; c++) {
      const int ts = h_corr[(size_t)s * cfg.kc + c];
      int to = -1;
      if (ts >= 0) std::memcpy(&to, &h_tpts[ts].w, 4);
      idx_out[(size_t)o * cfg.kc + c] = to;
      if (w_out) {  // weight of record (slot s, candidate c): group block s / 32, lane s % 32
        double wv;
        std::memcpy(&wv, h_rec.data() + (size_t)(s >> 5) * Rec::group_bytes(cfg.kc) + c * Rec::kCandBytes + Rec::kW + 8 * (s & 31), 8);
        w_out[(size_t)o * cfg.kc + c] = wv;
      }
      if (d2_out) d2_out[(size_t)o * cfg.kc + c] = h_d2[(size_t)s * cfg.kc + c];
    }
  }
  return SICP_OK;
}

sicp_status sicp_evaluate(int algo, sicp_cloud* src, sicp_cloud* tgt, const sicp_options* opts, const double* corr_pose7,
                          const double* eval_pose7, double* cost, double* g6, double* H36) {
  SICP_REQUIRE(corr_pose7 && eval_pose7 && cost && g6 && H36, "null argument");
  SICP_CHECK(validate(algo, src, tgt, opts));
  SICP_CHECK(validate_pose7(corr_pose7, "sicp_evaluate (corr_pose7)"));
  SICP_CHECK(validate_pose7(eval_pose7, "sicp_evaluate (eval_pose7)"));
  SICP_CUDA(cudaSetDevice(src->device));
  cudaStream_t st = current_stream();
  SICP_CHECK(ensure_built(src, st));
  SICP_CHECK(ensure_built(tgt, st));
  SICP_CHECK(precompute_pair(algo, src, tgt, opts, st, nullptr));
  Job jb;
  jb.algo = algo; jb.src = src; jb.tgt = tgt; jb.opts = opts; jb.cfg = make_cfg(algo, *opts); jb.st = st;
  SICP_CHECK(scratch_slot(src, jb.cfg.kc, &jb.sl));
  Slot* ws = jb.sl;
  LMConfig& cfg = jb.cfg;
  SICP_CHECK(jb.stage_class_map());
  std::memset(ws->h_ctl, 0, sizeof(RegCtl));
  std::memcpy(ws->h_ctl->pose, corr_pose7, 56);
  std::memcpy(ws->h_ctl->pass_pose[0], eval_pose7, 56);  // scratch row for the evaluation pose
  SICP_CUDA(cudaMemcpyAsync(ws->d_ctl, ws->h_ctl, sizeof(RegCtl), cudaMemcpyHostToDevice, st));
  SICP_CHECK(launch_cross_knn(src, tgt, ws->d_ctl->pose, nullptr, jb.class_map(), cfg.kc, ws->d_corr, ws->d_d2, st));
  SICP_CHECK(launch_estep(src, tgt, cfg, opts->gate_d2, ws->d_ctl->pose, nullptr, ws->d_corr, ws->d_d2, ws->d_rec, nullptr, st));
  double* d_out = &ws->d_ctl->pass_pose[8][0];
  SICP_CHECK(launch_evaluate(src, cfg, ws->d_rec, &ws->d_ctl->pass_pose[0][0], d_out, ws->d_partials, lm_grid_blocks(src->device), st));
  double* h_out = &ws->h_ctl->pass_pose[8][0];  // pinned
  SICP_CUDA(cudaMemcpyAsync(h_out, d_out, sizeof(double) * 28, cudaMemcpyDeviceToHost, st));
  SICP_CUDA(cudaStreamSynchronize(st));
  *cost = h_out[27];
  for (int a = 0; a < 6; a++) {
    g6[a] = h_out[21 + a];
    for (int b = 0; b < 6; b++) { const int hi = std::max(a, b), lo = std::min(a, b); H36[6 * a + b] = h_out[hi * (hi + 1) / 2 + lo]; }
  }
  return SICP_OK;
}

sicp_status sicp_fused_labels(sicp_cloud* src, sicp_cloud* tgt, const sicp_options* opts, const double* pose7, uint32_t* labels_out) {
  SICP_REQUIRE(pose7 && labels_out, "null argument");
  SICP_CHECK(validate(SICP_ALGO_EM, src, tgt, opts));
  SICP_CHECK(validate_pose7(pose7, "sicp_fused_labels"));
  SICP_CUDA(cudaSetDevice(src->device));
  cudaStream_t st = current_stream();
  SICP_CHECK(ensure_built(src, st));
  SICP_CHECK(ensure_built(tgt, st));
  SICP_CHECK(precompute_pair(SICP_ALGO_EM, src, tgt, opts, st, nullptr));
  Job jb;
  jb.algo = SICP_ALGO_EM; jb.src = src; jb.tgt = tgt; jb.opts = opts; jb.cfg = make_cfg(SICP_ALGO_EM, *opts); jb.st = st;
  SICP_CHECK(scratch_slot(src, 4, &jb.sl));
  Slot* ws = jb.sl;
  SICP_CHECK(jb.start(pose7));
  uint32_t* d_lab = nullptr;
  auto body = [&]() -> sicp_status {
    SICP_CUDA(cudaMallocAsync(&d_lab, sizeof(uint32_t) * std::max<size_t>(1, src->n), st));
    SICP_CHECK(launch_cross_knn(src, tgt, ws->d_ctl->pose, nullptr, nullptr, 4, ws->d_corr, ws->d_d2, st));
    SICP_CHECK(launch_fused_labels(src, tgt, opts->epsilon, opts->gate_d2, ws->d_ctl->pose, ws->d_corr, ws->d_d2, d_lab, st));
    SICP_CUDA(cudaMemcpyAsync(labels_out, d_lab, sizeof(uint32_t) * src->n, cudaMemcpyDeviceToHost, st));
    SICP_CUDA(cudaStreamSynchronize(st));
    return SICP_OK;
  };
  const sicp_status rc = body();
  if (d_lab) cudaFreeAsync(d_lab, st);
  return rc;
}

}  // extern "C"

// register.cu — host orchestration of one align(): covariance precompute, then outer passes
// [transform + kNN] -> [E-step] -> [cooperative LM solve + convergence test], all control state device-resident.
// Replaces the bodies of GICP::align (impl/gicp.hpp:29-175), SemanticIterativeClosestPoint::align
// (impl/semantic_icp.hpp:27-166) and EmIterativeClosestPoint::align (impl/em_icp.hpp:24-200).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>
#include "common.cuh"
#include "kernels.h"

namespace sicp {

static LMConfig make_cfg(int algo, const sicp_options& o) {
  LMConfig c;
  c.algo = algo;
  c.kc = algo == SICP_ALGO_EM ? 4 : 1;                      // em_icp.hpp:60 / gicp.hpp:69, semantic_icp.hpp:68
  c.eps = o.epsilon;
  c.kappa = 1.0 - o.epsilon; c.hk = 0.5 * c.kappa; c.k4 = 0.25 * c.kappa; c.aa = 1.0 - c.hk;
  c.max_iter = o.max_lm_iterations;
  c.mse_stop = algo == SICP_ALGO_SEMANTIC ? 1e-3 : 1e-5;    // semantic_icp.hpp:152 / gicp.hpp:154, em_icp.hpp:180
  c.outer_cap = algo == SICP_ALGO_SEMANTIC ? 35 : 50;
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
  return SICP_OK;
}
// Sophus::SE3d is a unit quaternion by construction; a pose7 coming through the C ABI has to be checked
static sicp_status validate_pose(const double* p7, const char* what) {
  double n2 = 0;
  bool finite = true;
  for (int i = 0; i < 7; i++) finite = finite && std::isfinite(p7[i]);
  for (int i = 0; i < 4; i++) n2 += p7[i] * p7[i];
  if (!finite || std::fabs(n2 - 1.0) > 1e-6) { set_error(std::string(what) + ": pose7 must be finite with a unit quaternion [qx,qy,qz,qw]"); return SICP_ERR_INVALID; }
  return SICP_OK;
}

// Covariances / label vectors of both clouds of a pair.  The two clouds are independent, and one covariance kernel is
// a single wave with a long tail (a few slow warps), so the target runs on a helper stream beside the source on `st`;
// `st` then waits for the target's ready event.  slot: index of the helper stream, or -1 to run both on `st` — the batch
// executor does that: with 8 registrations in flight the tails are already filled by other registrations and the extra
// streams cost 2-4 % of throughput (measured), while a lone registration gains 0.35 ms (cov stage 1.05 -> 0.70 ms).
static thread_local std::vector<cudaStream_t> t_helpers;
static sicp_status precompute_pair(int algo, sicp_cloud* src, sicp_cloud* tgt, const sicp_options* o, cudaStream_t st, int slot) {
  const int N = algo == SICP_ALGO_EM ? o->n_classes : 0;
  cudaStream_t saved = current_stream();
  sicp_status rc = SICP_OK;
  if (tgt != src && !tgt->pre_valid && slot >= 0) {
    while ((int)t_helpers.size() <= slot) {
      cudaStream_t h;
      SICP_CUDA(cudaStreamCreateWithFlags(&h, cudaStreamNonBlocking));
      t_helpers.push_back(h);
    }
    cudaStream_t h = t_helpers[slot];
    if (tgt->built_ev) SICP_CUDA(cudaStreamWaitEvent(h, tgt->built_ev, 0));
    sicp_set_stream(h);
    rc = sicp_cloud_precompute(tgt, o->k_cov, o->epsilon, N, o->confusion);
  } else if (tgt != src) {
    sicp_set_stream(st);
    rc = sicp_cloud_precompute(tgt, o->k_cov, o->epsilon, N, o->confusion);  // cached: checks the parameters only
  }
  if (rc == SICP_OK) {
    sicp_set_stream(st);
    rc = sicp_cloud_precompute(src, o->k_cov, o->epsilon, N, o->confusion);
  }
  sicp_set_stream(saved);
  SICP_CHECK(rc);
  if (src->ready_ev) SICP_CUDA(cudaStreamWaitEvent(st, src->ready_ev, 0));
  if (tgt->ready_ev) SICP_CUDA(cudaStreamWaitEvent(st, tgt->ready_ev, 0));
  return SICP_OK;
}

// Pinned control blocks are pooled: cudaMallocHost / cudaFreeHost synchronise the device and would serialise
// concurrent registrations.
static std::mutex g_pin_mu;
static std::vector<RegCtl*> g_pin_free;
static RegCtl* pin_get() {
  {
    std::lock_guard<std::mutex> lk(g_pin_mu);
    if (!g_pin_free.empty()) { RegCtl* p = g_pin_free.back(); g_pin_free.pop_back(); return p; }
  }
  RegCtl* p = nullptr;
  if (cudaMallocHost(&p, sizeof(RegCtl)) != cudaSuccess) return nullptr;
  return p;
}
static void pin_put(RegCtl* p) { std::lock_guard<std::mutex> lk(g_pin_mu); g_pin_free.push_back(p); }

// Per-registration device workspace
struct Workspace {
  int* d_corr = nullptr; float* d_d2 = nullptr; double* d_w = nullptr; float4* d_gpt = nullptr; double* d_gnt = nullptr; RegCtl* d_ctl = nullptr; double* d_partials = nullptr; int* d_map = nullptr;
  RegCtl* h_ctl = nullptr;  // pinned
  int grid = 0;
  cudaStream_t st = nullptr;
  sicp_status alloc(const sicp_cloud* src, const sicp_cloud* tgt, const LMConfig& cfg, int min_class_points, cudaStream_t s) {
    st = s;
    const size_t nc = (size_t)std::max(1, src->nslots) * cfg.kc;
    grid = lm_grid_blocks(src->device);
    SICP_CUDA(cudaMallocAsync(&d_corr, sizeof(int) * nc, st));
    SICP_CUDA(cudaMallocAsync(&d_d2, sizeof(float) * nc, st));
    SICP_CUDA(cudaMallocAsync(&d_w, sizeof(double) * nc, st));
    SICP_CUDA(cudaMallocAsync(&d_gpt, sizeof(float4) * nc, st));
    SICP_CUDA(cudaMallocAsync(&d_gnt, sizeof(double) * 3 * nc, st));
    SICP_CUDA(cudaMallocAsync(&d_ctl, sizeof(RegCtl), st));
    SICP_CUDA(cudaMallocAsync(&d_partials, sizeof(double) * 2 * 28 * grid, st));  // slab 0: LMSync (zeroed), slab 1: block partials
    SICP_CUDA(cudaMemsetAsync(d_partials, 0, sizeof(double) * 28 * grid, st));
    h_ctl = pin_get();
    if (!h_ctl) { set_error("pinned allocation failed"); return SICP_ERR_CUDA; }
    if (cfg.algo == SICP_ALGO_SEMANTIC) SICP_CHECK(make_class_map(src, tgt, min_class_points, &d_map, st));
    return SICP_OK;
  }
  void release() {
    if (d_corr) cudaFreeAsync(d_corr, st);
    if (d_d2) cudaFreeAsync(d_d2, st);
    if (d_w) cudaFreeAsync(d_w, st);
    if (d_gpt) cudaFreeAsync(d_gpt, st);
    if (d_gnt) cudaFreeAsync(d_gnt, st);
    if (d_ctl) cudaFreeAsync(d_ctl, st);
    if (d_partials) cudaFreeAsync(d_partials, st);
    if (d_map) cudaFreeAsync(d_map, st);
    if (h_ctl) pin_put(h_ctl);
    *this = Workspace();
  }
};

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
  int algo; sicp_cloud* src; sicp_cloud* tgt; const sicp_options* opts; LMConfig cfg; Workspace ws; StageTimer tm;
  sicp_result* out; int enqueued = 0; bool finished = false; int launches = 0; int d2h = 0;

  sicp_status start(const double* init7, cudaStream_t st, int lm_grid) {
    cfg = make_cfg(algo, *opts);
    tm.on = opts->profile != 0;
    SICP_CHECK(ws.alloc(src, tgt, cfg, opts->min_class_points, st));
    ws.grid = lm_grid;
    std::memset(ws.h_ctl, 0, sizeof(RegCtl));
    std::memcpy(ws.h_ctl->pose, init7, 56);
    SICP_CUDA(cudaMemcpyAsync(ws.d_ctl, ws.h_ctl, sizeof(RegCtl), cudaMemcpyHostToDevice, st));
    return SICP_OK;
  }
  static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
  double host_ms[4] = {0, 0, 0, 0};  // host time inside the launch calls of this job (kNN, E-step, LM, readback): diagnostics
  sicp_status enqueue_pass() {
    cudaStream_t st = ws.st;
    const int* stop = &ws.d_ctl->converged;
    double h0 = now_ms();
    tm.begin(SICP_STAGE_KNN, st);
    SICP_CHECK(launch_cross_knn(src, tgt, ws.d_ctl->pose, stop, ws.d_map, cfg.kc, ws.d_corr, ws.d_d2, st));
    tm.end(st);
    host_ms[0] += now_ms() - h0; h0 = now_ms();
    tm.begin(SICP_STAGE_ESTEP, st);
    SICP_CHECK(launch_estep(src, tgt, cfg, opts->gate_d2, ws.d_ctl->pose, stop, ws.d_corr, ws.d_d2, ws.d_w, ws.d_gpt, ws.d_gnt, ws.d_ctl, st));
    tm.end(st);
    host_ms[1] += now_ms() - h0; h0 = now_ms();
    tm.begin(SICP_STAGE_LM, st);
    SICP_CHECK(launch_lm(src, cfg, ws.d_w, ws.d_gpt, ws.d_gnt, ws.d_ctl, ws.d_partials, ws.grid, st));
    tm.end(st);
    host_ms[2] += now_ms() - h0;
    launches += 3;
    enqueued++;
    return SICP_OK;
  }
  // enqueue a chunk of passes, then the control-block readback
  sicp_status enqueue_chunk(int n) {
    const int cap = cfg.outer_cap + 2;
    for (int i = 0; i < n && enqueued < cap; i++) SICP_CHECK(enqueue_pass());
    SICP_CUDA(cudaMemcpyAsync(ws.h_ctl, ws.d_ctl, sizeof(RegCtl), cudaMemcpyDeviceToHost, ws.st));
    d2h += (int)sizeof(RegCtl);
    return SICP_OK;
  }
  // after the stream is synchronised: true when the registration is complete
  bool done_after_sync() const { return ws.h_ctl->converged != 0 || enqueued >= cfg.outer_cap + 2; }
  cudaEvent_t trace_ref = nullptr; int trace_id = 0;
  void finish() {
    const RegCtl& c = *ws.h_ctl;
    std::memcpy(out->pose7, c.pose, 56);
    out->outer_iter = c.outer; out->lm_iters_total = c.lm_iters_total; out->final_cost = c.final_cost; out->n_corr_last = c.n_corr_last;
    out->flags = c.flags; out->lm_evals_total = c.lm_evals_total;
    const int np = std::min(c.outer, 64);
    std::memcpy(out->pass_pose7, c.pass_pose, sizeof(double) * 7 * np);
    std::memcpy(out->pass_lm_iters, c.pass_lm_iters, sizeof(int) * np);
    out->d2h_bytes = d2h;
    for (int i = 0; i < 3; i++) out->lm_cycles[i] = (double)c.dbg_cycles[i];
    for (int i = 0; i < 3; i++) out->lm_cycles[3 + i] = (double)c.dbg_cycles[4 + i];
    out->gpu_launches = c.outer * 3;  // kernels that did work (passes enqueued past convergence return immediately)
    tm.collect(out, trace_ref, trace_id);
    ws.release();
    finished = true;
  }
};

// Streams of the batch executor are created once per host thread and reused.
static thread_local std::vector<cudaStream_t> t_streams;

// Runs all jobs with up to `max_concurrent` registrations in flight, each on its own stream.  A slot that finishes
// immediately picks up the next job (no wave barrier), so tails of one registration overlap the bulk of another.
static sicp_status run_jobs(std::vector<Job>& jobs, const double* init7s, int max_concurrent) {
  cudaStream_t base = current_stream();
  const int nj = (int)jobs.size();
  const int S = std::max(1, std::min(max_concurrent, nj));
  std::vector<cudaStream_t> streams(S, base);
  cudaEvent_t fork = nullptr;
  if (S > 1) {
    while ((int)t_streams.size() < S) {
      cudaStream_t s;
      SICP_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
      t_streams.push_back(s);
    }
    SICP_CUDA(cudaEventCreate(&fork));
    SICP_CUDA(cudaEventRecord(fork, base));
    for (int s = 0; s < S; s++) streams[s] = t_streams[s];  // each job waits for ITS clouds (built_ev), not for the whole base stream
  }
  sicp_status rc = SICP_OK;
  int kChunk = 3;  // passes enqueued between two readbacks of the control block
  if (const char* e = getenv("SICP_CHUNK")) { const int v = atoi(e); if (v > 0) kChunk = v; }
  // A lone solve takes every SM (one CTA each).  Concurrent solves get a quarter of the SMs each: their sweeps are longer,
  // so the latency-bound control step between sweeps idles a smaller share of the machine, and the kNN kernels of other
  // registrations run on the SMs no solve occupies.
  int lm_grid = lm_grid_blocks(jobs[0].src->device) / (S > 1 ? 4 : 1);
  if (const char* e = getenv("SICP_LM_GRID")) { const int g = atoi(e); if (g > 0 && S > 1) lm_grid = std::min(g, lm_grid_blocks(jobs[0].src->device)); }
  std::vector<int> slot_job(S, -1);
  int next = 0, live = 0;
  auto launch = [&](int slot) -> sicp_status {
    const int j = next++;
    slot_job[slot] = j;
    Job& jb = jobs[j];
    cudaStream_t st = streams[slot];
    // covariances / label vectors on this job's stream (no-op when cached); other jobs sharing a cloud wait on its event
    jb.tm.on = jb.opts->profile != 0;
    jb.trace_ref = fork; jb.trace_id = j;
    if (st != base) {  // the clouds may still be building on the stream that created them
      if (jb.src->built_ev) SICP_CUDA(cudaStreamWaitEvent(st, jb.src->built_ev, 0));
      if (jb.tgt->built_ev) SICP_CUDA(cudaStreamWaitEvent(st, jb.tgt->built_ev, 0));
    }
    jb.tm.begin(SICP_STAGE_COV, st);
    const sicp_status r = precompute_pair(jb.algo, jb.src, jb.tgt, jb.opts, st, S == 1 ? 0 : -1);
    jb.tm.end(st);
    SICP_CHECK(r);
    SICP_CHECK(jb.start(init7s + 7 * (size_t)j, st, lm_grid));
    SICP_CHECK(jb.enqueue_chunk(kChunk));
    live++;
    return SICP_OK;
  };
  // one completion event per slot, recorded after the control-block readback of each chunk.  The host POLLS the slots
  // (cudaEventQuery) and serves whichever registration finished its chunk first; blocking on one stream would leave
  // the streams of registrations that are already waiting for their next chunk empty.
  std::vector<cudaEvent_t> ev(S, nullptr);
  for (int s = 0; s < S; s++) SICP_CUDA(cudaEventCreateWithFlags(&ev[s], cudaEventDisableTiming));
  auto launch_and_mark = [&](int slot) -> sicp_status {
    SICP_CHECK(launch(slot));
    SICP_CUDA(cudaEventRecord(ev[slot], streams[slot]));
    return SICP_OK;
  };
  for (int s = 0; s < S && next < nj && rc == SICP_OK; s++) rc = launch_and_mark(s);
  while (live > 0 && rc == SICP_OK) {
    bool served = false;
    for (int s = 0; s < S && rc == SICP_OK; s++) {
      const int j = slot_job[s];
      if (j < 0) continue;
      const cudaError_t q = S == 1 ? cudaEventSynchronize(ev[s]) : cudaEventQuery(ev[s]);
      if (q == cudaErrorNotReady) continue;
      if (q != cudaSuccess) { set_error(std::string("stream failed: ") + cudaGetErrorString(q)); rc = SICP_ERR_CUDA; break; }
      served = true;
      if (jobs[j].done_after_sync()) {
        jobs[j].finish();
        live--;
        slot_job[s] = -1;
        if (next < nj) rc = launch_and_mark(s);
      } else {
        rc = jobs[j].enqueue_chunk(kChunk);
        if (rc == SICP_OK && cudaEventRecord(ev[s], streams[s]) != cudaSuccess) { set_error("event record failed"); rc = SICP_ERR_CUDA; }
      }
    }
    if (!served && rc == SICP_OK) std::this_thread::yield();
  }
  for (int s = 0; s < S; s++) if (ev[s]) cudaEventDestroy(ev[s]);
  if (getenv("SICP_TRACE")) {
    double h[3] = {0, 0, 0};
    for (Job& jb : jobs) for (int i = 0; i < 3; i++) h[i] += jb.host_ms[i];
    fprintf(stderr, "[sicp] host ms inside launches: kNN %.2f E-step %.2f LM %.2f (all jobs)\n", h[0], h[1], h[2]);
  }
  if (rc != SICP_OK) cudaDeviceSynchronize();  // error path: nothing may still be writing a pinned control block we are about to recycle
  for (Job& jb : jobs) if (!jb.finished && jb.ws.st) jb.ws.release();
  if (S > 1) {
    for (int s = 0; s < S; s++) {
      cudaEvent_t e;
      cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
      cudaEventRecord(e, streams[s]);
      cudaStreamWaitEvent(base, e, 0);
      cudaEventDestroy(e);
    }
    cudaEventDestroy(fork);
  }
  return rc;
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
  SICP_CHECK(validate_pose(init7, "sicp_register"));
  SICP_CUDA(cudaSetDevice(src->device));
  cudaStream_t st = current_stream();
  std::memset(out, 0, sizeof *out);
  StageTimer pre;
  pre.on = opts->profile != 0;
  pre.begin(SICP_STAGE_COV, st);
  SICP_CHECK(precompute_pair(algo, src, tgt, opts, st, 0));
  pre.end(st);
  std::vector<Job> jobs(1);
  jobs[0].algo = algo; jobs[0].src = src; jobs[0].tgt = tgt; jobs[0].opts = opts; jobs[0].out = out;
  sicp_status rc = run_jobs(jobs, init7, 1);
  if (rc == SICP_OK) pre.collect(out);
  return rc;
}

sicp_status sicp_register_batch(int algo, size_t n_pairs, sicp_cloud* const* src, sicp_cloud* const* tgt, const sicp_options* opts,
                                const double* init7s, sicp_result* out) {
  SICP_REQUIRE(src && tgt && init7s && out, "null argument");
  if (n_pairs == 0) return SICP_OK;
  for (size_t i = 0; i < n_pairs; i++) {
    SICP_CHECK(validate(algo, src[i], tgt[i], opts));
    SICP_REQUIRE(src[i]->device == src[0]->device, "all pairs of a batch must live on one device");
    SICP_CHECK(validate_pose(init7s + 7 * i, "sicp_register_batch"));
  }
  SICP_CUDA(cudaSetDevice(src[0]->device));
  std::memset(out, 0, sizeof(sicp_result) * n_pairs);
  std::vector<Job> jobs(n_pairs);
  for (size_t i = 0; i < n_pairs; i++) { jobs[i].algo = algo; jobs[i].src = src[i]; jobs[i].tgt = tgt[i]; jobs[i].opts = opts; jobs[i].out = out + i; }
  return run_jobs(jobs, init7s, opts->max_concurrent > 0 ? opts->max_concurrent : 8);
}

sicp_status sicp_correspondences(int algo, sicp_cloud* src, sicp_cloud* tgt, const sicp_options* opts, const double* pose7, int32_t* idx_out,
                                 double* w_out, float* d2_out) {
  SICP_REQUIRE(pose7 && idx_out, "null argument");
  SICP_CHECK(validate(algo, src, tgt, opts));
  SICP_CUDA(cudaSetDevice(src->device));
  SICP_CHECK(precompute_pair(algo, src, tgt, opts, current_stream(), 0));
  cudaStream_t st = current_stream();
  LMConfig cfg = make_cfg(algo, *opts);
  Workspace ws;
  sicp_status rc = ws.alloc(src, tgt, cfg, opts->min_class_points, st);
  if (rc != SICP_OK) { ws.release(); return rc; }
  std::memset(ws.h_ctl, 0, sizeof(RegCtl));
  std::memcpy(ws.h_ctl->pose, pose7, 56);
  const size_t nslot_c = (size_t)src->nslots * cfg.kc, n_c = src->n * cfg.kc;
  std::vector<int> h_corr(nslot_c); std::vector<float> h_d2(nslot_c); std::vector<double> h_w(nslot_c);
  std::vector<float4> h_spts(src->nslots), h_tpts(tgt->nslots);
  auto body = [&]() -> sicp_status {
    SICP_CUDA(cudaMemcpyAsync(ws.d_ctl, ws.h_ctl, sizeof(RegCtl), cudaMemcpyHostToDevice, st));
    SICP_CHECK(launch_cross_knn(src, tgt, ws.d_ctl->pose, nullptr, ws.d_map, cfg.kc, ws.d_corr, ws.d_d2, st));
    SICP_CHECK(launch_estep(src, tgt, cfg, opts->gate_d2, ws.d_ctl->pose, nullptr, ws.d_corr, ws.d_d2, ws.d_w, ws.d_gpt, ws.d_gnt, nullptr, st));
    SICP_CUDA(cudaMemcpyAsync(h_corr.data(), ws.d_corr, sizeof(int) * nslot_c, cudaMemcpyDeviceToHost, st));
    SICP_CUDA(cudaMemcpyAsync(h_d2.data(), ws.d_d2, sizeof(float) * nslot_c, cudaMemcpyDeviceToHost, st));
    SICP_CUDA(cudaMemcpyAsync(h_w.data(), ws.d_w, sizeof(double) * nslot_c, cudaMemcpyDeviceToHost, st));
    SICP_CUDA(cudaMemcpyAsync(h_spts.data(), src->d_pts, sizeof(float4) * src->nslots, cudaMemcpyDeviceToHost, st));
    SICP_CUDA(cudaMemcpyAsync(h_tpts.data(), tgt->d_pts, sizeof(float4) * tgt->nslots, cudaMemcpyDeviceToHost, st));
    SICP_CUDA(cudaStreamSynchronize(st));
    return SICP_OK;
  };
  rc = body();
  ws.release();
  SICP_CHECK(rc);
  (void)n_c;
  for (int s = 0; s < src->nslots; s++) {
    int o; std::memcpy(&o, &h_spts[s].w, 4);
    if (o < 0) continue;
    for (int c = 0; c < cfg.kc; c++) {
      const int ts = h_corr[(size_t)s * cfg.kc + c];
      int to = -1;
      if (ts >= 0) std::memcpy(&to, &h_tpts[ts].w, 4);
      idx_out[(size_t)o * cfg.kc + c] = to;
      if (w_out) w_out[(size_t)o * cfg.kc + c] = h_w[(size_t)c * src->nslots + s];  // gathered arrays are c-major
      if (d2_out) d2_out[(size_t)o * cfg.kc + c] = h_d2[(size_t)s * cfg.kc + c];
    }
  }
  return SICP_OK;
}

sicp_status sicp_evaluate(int algo, sicp_cloud* src, sicp_cloud* tgt, const sicp_options* opts, const double* corr_pose7,
                          const double* eval_pose7, double* cost, double* g6, double* H36) {
  SICP_REQUIRE(corr_pose7 && eval_pose7 && cost && g6 && H36, "null argument");
  SICP_CHECK(validate(algo, src, tgt, opts));
  SICP_CUDA(cudaSetDevice(src->device));
  SICP_CHECK(precompute_pair(algo, src, tgt, opts, current_stream(), 0));
  cudaStream_t st = current_stream();
  LMConfig cfg = make_cfg(algo, *opts);
  Workspace ws;
  sicp_status rc = ws.alloc(src, tgt, cfg, opts->min_class_points, st);
  if (rc != SICP_OK) { ws.release(); return rc; }
  std::memset(ws.h_ctl, 0, sizeof(RegCtl));
  std::memcpy(ws.h_ctl->pose, corr_pose7, 56);
  std::memcpy(ws.h_ctl->pass_pose[0], eval_pose7, 56);  // scratch slot for the evaluation pose
  double h_out[28];
  auto body = [&]() -> sicp_status {
    SICP_CUDA(cudaMemcpyAsync(ws.d_ctl, ws.h_ctl, sizeof(RegCtl), cudaMemcpyHostToDevice, st));
    SICP_CHECK(launch_cross_knn(src, tgt, ws.d_ctl->pose, nullptr, ws.d_map, cfg.kc, ws.d_corr, ws.d_d2, st));
    SICP_CHECK(launch_estep(src, tgt, cfg, opts->gate_d2, ws.d_ctl->pose, nullptr, ws.d_corr, ws.d_d2, ws.d_w, ws.d_gpt, ws.d_gnt, nullptr, st));
    double* d_out = &ws.d_ctl->pass_pose[8][0];
    SICP_CHECK(launch_evaluate(src, cfg, ws.d_w, ws.d_gpt, ws.d_gnt, &ws.d_ctl->pass_pose[0][0], d_out, ws.d_partials, ws.grid, st));
    SICP_CUDA(cudaMemcpyAsync(h_out, d_out, sizeof h_out, cudaMemcpyDeviceToHost, st));
    SICP_CUDA(cudaStreamSynchronize(st));
    return SICP_OK;
  };
  rc = body();
  ws.release();
  SICP_CHECK(rc);
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
  SICP_CUDA(cudaSetDevice(src->device));
  SICP_CHECK(precompute_pair(SICP_ALGO_EM, src, tgt, opts, current_stream(), 0));
  cudaStream_t st = current_stream();
  LMConfig cfg = make_cfg(SICP_ALGO_EM, *opts);
  Workspace ws;
  sicp_status rc = ws.alloc(src, tgt, cfg, 0, st);
  if (rc != SICP_OK) { ws.release(); return rc; }
  std::memset(ws.h_ctl, 0, sizeof(RegCtl));
  std::memcpy(ws.h_ctl->pose, pose7, 56);
  uint32_t* d_lab = nullptr;
  auto body = [&]() -> sicp_status {
    SICP_CUDA(cudaMallocAsync(&d_lab, sizeof(uint32_t) * std::max<size_t>(1, src->n), st));
    SICP_CUDA(cudaMemcpyAsync(ws.d_ctl, ws.h_ctl, sizeof(RegCtl), cudaMemcpyHostToDevice, st));
    SICP_CHECK(launch_cross_knn(src, tgt, ws.d_ctl->pose, nullptr, nullptr, 4, ws.d_corr, ws.d_d2, st));
    SICP_CHECK(launch_fused_labels(src, tgt, opts->epsilon, opts->gate_d2, ws.d_ctl->pose, ws.d_corr, ws.d_d2, d_lab, st));
    SICP_CUDA(cudaMemcpyAsync(labels_out, d_lab, sizeof(uint32_t) * src->n, cudaMemcpyDeviceToHost, st));
    SICP_CUDA(cudaStreamSynchronize(st));
    return SICP_OK;
  };
  rc = body();
  if (d_lab) cudaFreeAsync(d_lab, st);
  ws.release();
  return rc;
}

}  // extern "C"

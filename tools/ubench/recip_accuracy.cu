// Accuracy of the third-order reciprocal / reciprocal-square-root steps and of the table log used by the LM sweep
// (csrc/lm.cu: rcp_pos3, rsqrt_pos3, log_tab — the same expressions, repeated here so that the check stands alone),
// against correctly rounded references computed in double-double / long double on the host.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/recip_accuracy tools/ubench/recip_accuracy.cu
#include <cmath>
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
__device__ __forceinline__ double rcp_pos3(double a) {
  double x;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(a));
  const double e = fma(-a, x, 1.0);
  return fma(x, fma(e, e, e), x);
}
__device__ __forceinline__ double rsqrt_pos3(double a) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  const double e = fma(-(a * y), y, 1.0);
  return fma(y, e * fma(0.375, e, 0.5), y);
}
__device__ __forceinline__ double log_tab(double x, const double2* tab) {
  const int hi = __double2hiint(x), lo = __double2loint(x);
  const int k = (hi >> 20) - 1023;
  const double2 t = tab[(hi >> 12) & 255];
  const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, lo);
  const double r = fma(m, t.x, -1.0);
  double q = fma(r, -1.0 / 6.0, 0.2);
  q = fma(r, q, -0.25);
  q = fma(r, q, 1.0 / 3.0);
  q = fma(r, q, -0.5);
  const double dk = (double)k;
  const double small = fma(r * r, q, fma(dk, 1.90821492927058770002e-10, r));
  return fma(dk, 6.93147180369123816490e-01, t.y + small);
}
__global__ void k(const double* in, int n, double* o_rcp, double* o_rsq, double* o_log, double* o_seed_rcp, double* o_seed_rsq) {
  __shared__ double2 tab[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) { const double c = 1.0 / (1.0 + (i + 0.5) * (1.0 / 256)); tab[i] = make_double2(c, -log(c)); }
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double a = in[i];
  o_rcp[i] = rcp_pos3(a); o_rsq[i] = rsqrt_pos3(a); o_log[i] = log_tab(a < 1.0 ? 1.0 / a : a, tab);
  double x, y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(a));
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  o_seed_rcp[i] = x; o_seed_rsq[i] = y;
}
int main() {
  const int n = 1 << 22;
  std::vector<double> h(n);
  unsigned long long s = 88172645463325252ull;
  for (int i = 0; i < n; i++) {  // log-uniform over [1e-16, 1e16], the range the sweep can see, plus values next to 1
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;
    const double u = (double)(s >> 11) / 9007199254740992.0;
    h[i] = (i & 3) == 0 ? 1.0 + u * 1e-3 : std::pow(10.0, -16.0 + 32.0 * u);
  }
  double *d_in, *d[5];
  cudaMalloc(&d_in, 8 * n); cudaMemcpy(d_in, h.data(), 8 * n, cudaMemcpyHostToDevice);
  for (auto& p : d) cudaMalloc(&p, 8 * n);
  k<<<(n + 255) / 256, 256>>>(d_in, n, d[0], d[1], d[2], d[3], d[4]);
  std::vector<double> r[5];
  for (int j = 0; j < 5; j++) { r[j].resize(n); cudaMemcpy(r[j].data(), d[j], 8 * n, cudaMemcpyDeviceToHost); }
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("CUDA error\n"); return 1; }
  long double worst[5] = {0, 0, 0, 0, 0};
  double worst_log_abs = 0;
  for (int i = 0; i < n; i++) {
    const long double a = h[i];
    const long double ref_rcp = 1.0L / a, ref_rsq = 1.0L / sqrtl(a);
    const long double x = h[i] < 1.0 ? 1.0 / h[i] : h[i];
    const long double ref_log = logl(x);
    worst[0] = fmaxl(worst[0], fabsl(((long double)r[0][i] - ref_rcp) / ref_rcp));
    worst[1] = fmaxl(worst[1], fabsl(((long double)r[1][i] - ref_rsq) / ref_rsq));
    if (ref_log > 1e-3L) worst[2] = fmaxl(worst[2], fabsl(((long double)r[2][i] - ref_log) / ref_log));
    worst_log_abs = fmax(worst_log_abs, (double)fabsl((long double)r[2][i] - ref_log));
    worst[3] = fmaxl(worst[3], fabsl(((long double)r[3][i] - ref_rcp) / ref_rcp));
    worst[4] = fmaxl(worst[4], fabsl(((long double)r[4][i] - ref_rsq) / ref_rsq));
  }
  const double ulp = 1.1102230246251565e-16;  // 2^-53
  printf("max relative error over %d inputs (units of 2^-53):\n", n);
  printf("  rcp_pos3   %.3f   (MUFU seed alone: 2^%.2f)\n", (double)(worst[0] / ulp), log2((double)worst[3]));
  printf("  rsqrt_pos3 %.3f   (MUFU seed alone: 2^%.2f)\n", (double)(worst[1] / ulp), log2((double)worst[4]));
  printf("  log_tab    %.3f for log(x) > 1e-3; max absolute error anywhere %.3e\n", (double)(worst[2] / ulp), worst_log_abs);
  return 0;
}

// Microbenchmark: FP64 FMA (DFMA) vs FP64 tensor MMA (mma.sync m8n8k4 f64, DMMA) issue rates on sm_100a, alone and mixed.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 fp64_pipes.cu -o fp64_pipes
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int MODE, int ILP>
__global__ void k(double* out, int iters, long long* cyc) {
  double f[ILP], c0[ILP], c1[ILP];
  const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int i = 0; i < ILP; i++) { f[i] = i; c0[i] = i; c1[i] = -i; }
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) {
      if (MODE == 0 || MODE == 2) f[i] = fma(f[i], a, b);
      if (MODE == 1 || MODE == 2) dmma(c0[i], c1[i], a, b);
    }
  }
  long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < ILP; i++) s += f[i] + c0[i] + c1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE, int ILP>
void run(const char* name, int warps) {
  double* out; long long* cyc; cudaMalloc(&out, 8 * 148 * 1024); cudaMalloc(&cyc, 8);
  const int iters = 20000;
  k<MODE, ILP><<<148, warps * 32>>>(out, 10, cyc);
  k<MODE, ILP><<<148, warps * 32>>>(out, iters, cyc);
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const double per = (double)h / iters / ILP;  // cycles per (instr per warp)
  printf("%-8s ILP=%d warps/SM=%2d: %.2f cycles per warp-instr-slot; per SMSP: %.2f cycles/instr\n", name, ILP, warps, per, per / (warps / 4.0));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<0, 1>("DFMA", 4); run<0, 8>("DFMA", 4); run<0, 8>("DFMA", 16);
  run<1, 1>("DMMA", 4); run<1, 8>("DMMA", 4); run<1, 8>("DMMA", 16);
  run<2, 8>("DFMA+DMMA", 4); run<2, 8>("DFMA+DMMA", 16);
  return 0;
}

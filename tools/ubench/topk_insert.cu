// Microbenchmark: cost of one sorted insertion into a register-resident list of K 64-bit keys (the TopK chain of
// csrc/knn.cuh) for three forms of the compare-exchange step, at several occupancies.
//   MODE 0: C++ (compiler emits two 64-bit integer compares + 4 selects per step)
//   MODE 1: one 64-bit integer compare (setp.lt.u64 -> 2 ISETP) + 4 selects
//   MODE 2: one f64 compare on the key bits (DSETP) + 4 selects
//   MODE 3: like 2 but the chain starts at slot K/2 when the candidate is not below key[K/2-1] (warp-uniform skip)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 topk_insert.cu -o topk_insert
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
template <int MODE>
__device__ __forceinline__ void cx(u64& slot, u64& carry) {
  if (MODE == 0) {
    const u64 cur = slot; const bool lt = carry < cur; slot = lt ? carry : cur; carry = lt ? cur : carry;
  } else if (MODE == 1) {
    u64 lo, hi;
    asm("{\n\t.reg .pred p;\n\tsetp.lt.u64 p, %3, %2;\n\tselp.b64 %0, %3, %2, p;\n\tselp.b64 %1, %2, %3, p;\n\t}" : "=&l"(lo), "=&l"(hi) : "l"(slot), "l"(carry));
    slot = lo; carry = hi;
  } else {
    u64 lo, hi;
    asm("{\n\t.reg .pred p;\n\t.reg .f64 a, b;\n\tmov.b64 a, %2;\n\tmov.b64 b, %3;\n\tsetp.lt.f64 p, b, a;\n\tselp.b64 %0, %3, %2, p;\n\tselp.b64 %1, %2, %3, p;\n\t}"
        : "=&l"(lo), "=&l"(hi) : "l"(slot), "l"(carry));
    slot = lo; carry = hi;
  }
}
// GEN 0: candidates shrink geometrically (always accepted, land at the head); GEN 1: just below the current worst (land at the tail)
// MODE 4: shift form — every slot compares the ORIGINAL candidate (no carried dependency): key[i] = lt[i-1] ? key[i-1] : lt[i] ? c : key[i]
template <int K>
__device__ __forceinline__ void insert_shift(u64 (&key)[K], u64 c) {
  const double cd = __longlong_as_double((long long)c);
  bool lt[K];
#pragma unroll
  for (int i = 0; i < K; i++) lt[i] = cd < __longlong_as_double((long long)key[i]);
#pragma unroll
  for (int i = K - 1; i > 0; i--) {
    const u64 t = lt[i] ? c : key[i];
    key[i] = lt[i - 1] ? key[i - 1] : t;
  }
  key[0] = lt[0] ? c : key[0];
}
template <int MODE, int K, int GEN>
__global__ void k(const u64* in, u64* out, int iters, long long* cyc) {
  u64 key[K];
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
  for (int i = 0; i < K; i++) key[i] = GEN == 0 ? (~0ull >> 2) : ((u64)(i + 1) << 40);
  u64 x = in[tid & 1023] | 1;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    x = x * 6364136223846793005ull + 1442695040888963407ull;   // pseudo-random candidate (2 IMAD-class ops, off the ALU pipe)
    u64 c = GEN == 0 ? ((x >> 3) >> (it >> 7)) : (key[K - 1] - 1 - (x >> 56));
    if (!(c < key[K - 1])) continue;
    if (MODE == 3) {
      const bool upper = !(c < key[K / 2 - 1]);
      if (__all_sync(0xffffffffu, upper)) {
#pragma unroll
        for (int i = K / 2; i < K; i++) cx<2>(key[i], c);
        continue;
      }
    }
    if (MODE == 4) { insert_shift<K>(key, c); continue; }
#pragma unroll
    for (int i = 0; i < K; i++) cx<MODE == 3 ? 2 : MODE>(key[i], c);
  }
  long long t1 = clock64();
  u64 s = 0;
#pragma unroll
  for (int i = 0; i < K; i++) s ^= key[i];
  out[tid] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE, int K, int GEN>
void run(int warps) {
  u64 *in, *out; long long* cyc; cudaMalloc(&in, 8 * 1024); cudaMalloc(&out, 8 * 148 * 1024); cudaMalloc(&cyc, 8);
  u64 h_in[1024]; for (int i = 0; i < 1024; i++) h_in[i] = 0x9e3779b97f4a7c15ull * (i + 1);
  cudaMemcpy(in, h_in, sizeof h_in, cudaMemcpyHostToDevice);
  const int iters = 4000;
  k<MODE, K, GEN><<<148, warps * 32>>>(in, out, 10, cyc);
  k<MODE, K, GEN><<<148, warps * 32>>>(in, out, iters, cyc);
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("mode %d K=%2d gen %d warps/SM=%2d: %8.1f cycles per candidate per warp; x warps/SMSP: %7.1f SMSP-cycles per candidate\n", MODE, K, GEN, warps, (double)h / iters, (double)h / iters / (warps / 4.0));
  cudaFree(in); cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int w : {4, 8, 16}) { run<0, 20, 0>(w); run<1, 20, 0>(w); run<2, 20, 0>(w); run<3, 20, 0>(w); run<2, 20, 1>(w); run<3, 20, 1>(w); run<4, 20, 0>(w); run<4, 20, 1>(w); }
  for (int w : {4, 16}) { run<0, 4, 0>(w); run<1, 4, 0>(w); run<2, 4, 0>(w); run<4, 4, 0>(w); }
  return 0;
}

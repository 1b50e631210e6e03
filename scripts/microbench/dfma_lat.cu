// DFMA throughput vs independent chains per warp and warps per SM sub-partition (development aid).
#include <cstdio>
#include <cuda_runtime.h>
template <int CH>
__global__ void k(double *out, int iters, double seed) {
  double a[CH];
#pragma unroll
  for (int c = 0; c < CH; c++) a[c] = seed + c;
  const double m = 1.0000001, b = 1e-9;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 16; u++)
#pragma unroll
      for (int c = 0; c < CH; c++) a[c] = fma(a[c], m, b);
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < CH; c++) s += a[c];
  if (s == 12345.678) out[0] = s;
}
template <int CH>
void run(int warps_per_smsp, double *d) {
  int threads = 32 * 4 * warps_per_smsp;   // one block per SM
  int blocks = 148, iters = 2048;
  if (threads > 1024) { blocks *= threads / 1024; threads = 1024; }
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<CH><<<blocks, threads>>>(d, iters, 0.5);
  cudaEventRecord(e0);
  k<CH><<<blocks, threads>>>(d, iters, 0.7);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double inst = (double) iters * 16 * CH * (threads / 32) * blocks;   // warp instructions
  double per_clk_smsp = inst / (ms * 1e-3 * 1.965e9) / (148 * 4);
  printf("chains %d warps/smsp %2d : %.3f DFMA/clk/SMSP  (%.1f TF)\n", CH, warps_per_smsp, per_clk_smsp, inst * 64 / (ms * 1e-3) / 1e12);
}
int main() {
  double *d; cudaMalloc(&d, 8);
  for (int w : {1, 2, 3, 4, 6, 8}) { run<1>(w, d); run<2>(w, d); run<3>(w, d); run<4>(w, d); run<8>(w, d); }
  return 0;
}

// DFMA throughput for different operand patterns (development aid): immediates vs 3 register operands
// vs the Chebyshev recurrence + accumulate pattern of the force kernel.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE, int CH>
__global__ void k(double *out, const double *in, int iters) {
  double a[CH], x[CH], y[CH], s[CH];
#pragma unroll
  for (int c = 0; c < CH; c++) { a[c] = in[c] + threadIdx.x; x[c] = in[c + 8]; y[c] = in[c + 16]; s[c] = 0; }
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++)
#pragma unroll
      for (int c = 0; c < CH; c++) {
        if (MODE == 0) { a[c] = fma(a[c], 1.0000001, 1e-9); a[c] = fma(a[c], 1.0000001, 1e-9); }
        if (MODE == 1) { a[c] = fma(a[c], x[c], y[c]); a[c] = fma(a[c], x[c], y[c]); }
        if (MODE == 2) { const double t = fma(x[c], a[c], -y[c]); s[c] = fma(in[24], t, s[c]); y[c] = a[c]; a[c] = t; }   // 2 DFMA
      }
  }
  double r = 0;
#pragma unroll
  for (int c = 0; c < CH; c++) r += a[c] + s[c] + y[c];
  if (r == 12345.678) out[0] = r;
}
template <int MODE, int CH>
void run(int warps_per_smsp, double *d, double *in) {
  int threads = 32 * 4 * warps_per_smsp, blocks = 148, iters = 2048;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE, CH><<<blocks, threads>>>(d, in, iters);
  cudaEventRecord(e0);
  k<MODE, CH><<<blocks, threads>>>(d, in, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double inst = (double) iters * 8 * 2 * CH * (threads / 32) * blocks;
  printf("mode %d chains %d warps/smsp %d : %.3f DFMA/clk/SMSP\n", MODE, CH, warps_per_smsp, inst / (ms * 1e-3 * 1.965e9) / (148 * 4));
}
int main() {
  double *d, *in; cudaMalloc(&d, 8); cudaMalloc(&in, 256);
  double h[32]; for (int i = 0; i < 32; i++) h[i] = 0.3 + 0.01 * i;
  cudaMemcpy(in, h, 256, cudaMemcpyHostToDevice);
  for (int w : {2, 4}) {
    run<0, 2>(w, d, in); run<1, 2>(w, d, in); run<2, 2>(w, d, in);
    run<0, 4>(w, d, in); run<1, 4>(w, d, in); run<2, 4>(w, d, in);
  }
  return 0;
}

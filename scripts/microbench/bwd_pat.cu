// Which formulation of the backward angular evaluation (A(z), A'(z) of a degree-18 polynomial with per-atom, i.e.
// warp-uniform but run-time, coefficients) issues fastest on the B200 FP64 pipe?  (development aid)
//   F0  Horner with derivative, coefficient from shared memory:   d = d z + b ; b = b z + a_k          (kernel today)
//   F1  two Horner chains with separate coefficient sets (a_k, e_k = (k+1) a_{k+1}), LDS.128 per order, the CH chains
//       of one order issued back to back so that the coefficient is the SAME third operand of consecutive DFMAs
//   F2  F0 with the coefficient in __constant__ memory (upper bound for "third operand not from the register file")
//   F3  F1 with the coefficients in __constant__ memory
// CH = independent triplets per lane (2 in the kernel).  Output: SMSP cycles per warp-triplet (36 or 35 DFMA).
#include <cstdio>
#include <cuda_runtime.h>
__constant__ double2 ccoef[19];
template <int F, int CH>
__global__ void __launch_bounds__(128, 4) k(double *out, const double *in, int iters) {
  __shared__ double2 coef[4][19];
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) < 19) coef[w][threadIdx.x & 31] = make_double2(in[threadIdx.x & 31], in[(threadIdx.x & 31) + 1]);
  __syncthreads();
  double z[CH];
#pragma unroll
  for (int c = 0; c < CH; c++) z[c] = in[c] + 1e-3 * threadIdx.x;
  double acc = 0;
  for (int i = 0; i < iters; i++) {
    double b[CH], d[CH];
    if (F == 0 || F == 2) {
      const double top = (F == 2) ? ccoef[18].x : coef[w][18].x;
#pragma unroll
      for (int c = 0; c < CH; c++) { b[c] = top; d[c] = 0.0; }
#pragma unroll
      for (int n = 17; n >= 0; n--) {
        const double ak = (F == 2) ? ccoef[n].x : coef[w][n].x;
#pragma unroll
        for (int c = 0; c < CH; c++) { d[c] = fma(d[c], z[c], b[c]); b[c] = fma(b[c], z[c], ak); }
      }
    } else {
      const double2 top = (F == 3) ? ccoef[18] : coef[w][18];
#pragma unroll
      for (int c = 0; c < CH; c++) { b[c] = top.x; d[c] = top.y; }
#pragma unroll
      for (int n = 17; n >= 0; n--) {
        const double2 q = (F == 3) ? ccoef[n] : coef[w][n];
#pragma unroll
        for (int c = 0; c < CH; c++) b[c] = fma(b[c], z[c], q.x);
        if (n > 0) {
#pragma unroll
          for (int c = 0; c < CH; c++) d[c] = fma(d[c], z[c], q.y);
        }
      }
    }
#pragma unroll
    for (int c = 0; c < CH; c++) { acc += b[c] * d[c]; z[c] += 1e-9; }
  }
  if (acc == 12345.678) out[0] = acc;
}
template <int F, int CH>
void run(double *d, double *in) {
  int threads = 128, blocks = 148 * 4, iters = 8192;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<F, CH><<<blocks, threads>>>(d, in, iters);
  cudaEventRecord(e0);
  k<F, CH><<<blocks, threads>>>(d, in, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double trip = (double) iters * CH * (threads / 32) * blocks;    // warp-triplets
  const double cyc = ms * 1e-3 * 1.965e9 * 592 / trip;
  printf("F%d chains %d : %.1f SMSP-cycles per warp-triplet = %.2f per DFMA\n", F, CH, cyc, cyc / ((F == 0 || F == 2) ? 36.0 : 35.0));
}
int main() {
  double *d, *in; cudaMalloc(&d, 8); cudaMalloc(&in, 512);
  double h[64]; for (int i = 0; i < 64; i++) h[i] = 0.3 + 0.01 * i;
  cudaMemcpy(in, h, 512, cudaMemcpyHostToDevice);
  cudaMemcpyToSymbol(ccoef, h, sizeof(double2) * 19);
  run<0, 2>(d, in); run<1, 2>(d, in); run<2, 2>(d, in); run<3, 2>(d, in);
  run<0, 4>(d, in); run<1, 4>(d, in); run<2, 4>(d, in); run<3, 4>(d, in);
  run<0, 1>(d, in); run<1, 1>(d, in);
  return 0;
}

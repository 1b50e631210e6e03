// Which formulations of the inner loops escape the 3-register-operand DFMA rate (development aid).
#include <cstdio>
#include <cuda_runtime.h>
// MODE 0: bwd pattern now: U recurrence + A += d_n U + Ap += e_n U, coefficients from shared memory (LDS.128)
// MODE 1: same but coefficients from __constant__ memory (c[bank] operand)
// MODE 2: fwd pattern now: T recurrence + S_n += w T_n (19 register accumulators)
// MODE 3: fwd Z pattern: Z_n = y2 Z_{n-1} - Z_{n-2}; S_n += Z_n
// MODE 4: Horner from shared coefficients: A = A*z + c_k ; Ap = Ap*z + e_k
__constant__ double2 ccoef[19];
template <int MODE, int CH>
__global__ void k(double *out, const double *in, int iters) {
  __shared__ double2 coef[19];
  if (threadIdx.x < 19) coef[threadIdx.x] = make_double2(in[threadIdx.x], in[threadIdx.x + 1]);
  __syncthreads();
  double y2[CH], w[CH], S[19];
#pragma unroll
  for (int c = 0; c < CH; c++) { y2[c] = in[c] + 1e-3 * threadIdx.x; w[c] = in[c + 4] + 1e-4 * threadIdx.x; }
#pragma unroll
  for (int n = 0; n < 19; n++) S[n] = 0;
  double acc = 0;
  for (int i = 0; i < iters; i++) {
    if (MODE == 0 || MODE == 1 || MODE == 4) {
      double u0[CH], u1[CH], A[CH], Ap[CH];
#pragma unroll
      for (int c = 0; c < CH; c++) { u0[c] = 1.0; u1[c] = y2[c]; A[c] = 0.1; Ap[c] = 0.2; }
#pragma unroll
      for (int n = 2; n < 19; n++) {
        const double2 q = (MODE == 1) ? ccoef[n] : coef[n];
#pragma unroll
        for (int c = 0; c < CH; c++) {
          if (MODE == 4) { A[c] = fma(A[c], y2[c], q.x); Ap[c] = fma(Ap[c], y2[c], q.y); }
          else {
            const double un = fma(y2[c], u1[c], -u0[c]);
            A[c] = fma(q.x, un, A[c]); Ap[c] = fma(q.y, un, Ap[c]);
            u0[c] = u1[c]; u1[c] = un;
          }
        }
      }
#pragma unroll
      for (int c = 0; c < CH; c++) { acc += A[c] * Ap[c]; y2[c] += 1e-9; }
    } else {
      double t0[CH], t1[CH];
#pragma unroll
      for (int c = 0; c < CH; c++) { t0[c] = (MODE == 3) ? w[c] : 1.0; t1[c] = (MODE == 3) ? w[c] * y2[c] : y2[c]; }
#pragma unroll
      for (int n = 2; n < 19; n++) {
#pragma unroll
        for (int c = 0; c < CH; c++) {
          const double tn = fma(y2[c], t1[c], -t0[c]);
          if (MODE == 3) S[n] += tn; else S[n] = fma(w[c], tn, S[n]);
          t0[c] = t1[c]; t1[c] = tn;
        }
      }
#pragma unroll
      for (int c = 0; c < CH; c++) y2[c] += 1e-9;
    }
  }
#pragma unroll
  for (int n = 0; n < 19; n++) acc += S[n];
  if (acc == 12345.678) out[0] = acc;
}
template <int MODE, int CH>
void run(double *d, double *in) {
  int threads = 256, blocks = 148 * 4, iters = 4096;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE, CH><<<blocks, threads>>>(d, in, iters);
  cudaEventRecord(e0);
  k<MODE, CH><<<blocks, threads>>>(d, in, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double trip = (double) iters * CH * (threads / 32) * blocks;    // warp-triplets
  printf("mode %d chains %d : %.1f SMSP-cycles per warp-triplet (17 orders)\n", MODE, CH, ms * 1e-3 * 1.965e9 * 592 / trip);
}
int main() {
  double *d, *in; cudaMalloc(&d, 8); cudaMalloc(&in, 512);
  double h[64]; for (int i = 0; i < 64; i++) h[i] = 0.3 + 0.01 * i;
  cudaMemcpy(in, h, 512, cudaMemcpyHostToDevice);
  cudaMemcpyToSymbol(ccoef, h, sizeof(double2) * 19);
  run<0, 1>(d, in); run<1, 1>(d, in); run<4, 1>(d, in); run<2, 1>(d, in); run<3, 1>(d, in);
  run<0, 2>(d, in); run<1, 2>(d, in); run<4, 2>(d, in); run<2, 2>(d, in); run<3, 2>(d, in);
  return 0;
}

// Does ptxas turn warp-uniform coefficients into uniform-register operands? (development aid)
#include <cstdio>
#include <cuda_runtime.h>
// SRC 0: static shared array, 1: dynamic shared + blockIdx-only offset, 2: global scratch indexed by blockIdx,
// SRC 3: dynamic shared + (threadIdx.x>>5) offset (what the kernel does today), 4: __shfl_sync broadcast
template <int SRC>
__global__ void k(double *out, const double *in, double *scratch, int iters) {
  __shared__ double sc[32];
  extern __shared__ double dyn[];
  if (threadIdx.x < 19) { sc[threadIdx.x] = in[threadIdx.x] * 1.5; dyn[threadIdx.x + 64 * (threadIdx.x >> 5)] = in[threadIdx.x] * 1.5; scratch[blockIdx.x * 32 + threadIdx.x] = in[threadIdx.x] * 1.5; }
  __syncthreads();
  double a[19];
#pragma unroll
  for (int n = 0; n < 19; n++) {
    if (SRC == 0) a[n] = sc[n];
    if (SRC == 1) a[n] = dyn[n + (blockIdx.x & 1)];
    if (SRC == 2) a[n] = scratch[blockIdx.x * 32 + n];
    if (SRC == 3) a[n] = dyn[n + 64 * (threadIdx.x >> 5)];
    if (SRC == 4) a[n] = __shfl_sync(0xffffffffu, in[n] + threadIdx.x, 0);
  }
  double z = in[20] + 1e-3 * threadIdx.x, acc = 0;
  for (int i = 0; i < iters; i++) {
    double b = a[18], d = 0.0;
#pragma unroll
    for (int n = 17; n >= 0; n--) { d = fma(d, z, b); b = fma(b, z, a[n]); }
    acc += b * d;
    z += 1e-9;
  }
  if (acc == 12345.678) out[0] = acc;
}
template <int SRC>
void run(double *d, double *in, double *scr) {
  int threads = (SRC == 3) ? 128 : 32, blocks = (SRC == 3) ? 148 * 4 : 148 * 16, iters = 8192;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<SRC><<<blocks, threads, 4096>>>(d, in, scr, iters);
  cudaEventRecord(e0);
  k<SRC><<<blocks, threads, 4096>>>(d, in, scr, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double trip = (double) iters * (threads / 32) * blocks;
  printf("src %d : %.1f SMSP-cycles per Horner-with-derivative (18 orders, 36 DFMA)\n", SRC, ms * 1e-3 * 1.965e9 * 592 / trip);
}
int main() {
  double *d, *in, *scr; cudaMalloc(&d, 8); cudaMalloc(&in, 512); cudaMalloc(&scr, 148 * 16 * 32 * 8);
  double h[64]; for (int i = 0; i < 64; i++) h[i] = 0.3 + 0.01 * i;
  cudaMemcpy(in, h, 512, cudaMemcpyHostToDevice);
  run<0>(d, in, scr); run<1>(d, in, scr); run<2>(d, in, scr); run<3>(d, in, scr); run<4>(d, in, scr);
  return 0;
}

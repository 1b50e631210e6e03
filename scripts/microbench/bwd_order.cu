// Does the ORDER of the backward Horner DFMAs matter on the B200 FP64 pipe?  (development aid)
// Hypothesis: a DFMA costs ~max(2, number of 64-bit register operands NOT served by the operand-reuse cache) cycles
// per SMSP.  ptxas orders the two triplets' chains a,b,a,b (every second DFMA has three fresh operands, 2.5 on average);
// a Gray-code order  P_a(z_a,e) P_b(z_b,e*) A_b(z_b*,a) A_a(z_a,a*) P_a(z_a*,e') ...  has two fresh operands everywhere.
//   F0  Horner with derivative (the kernel's form), coefficients from shared memory
//   F6  decoupled chains (A with a_k, A'/2 with e_k), source order Pa Pb Ab Aa  (ptxas -O3 re-sorts it to Pa Pb Aa Ab;
//       build with -Xptxas -O1 to keep the source order)
//   F7  F6 with a skewed start (true dependencies order the first round; ptxas then keeps the rotation)
#include <cstdio>
#include <cuda_runtime.h>
template <int F>
__global__ void __launch_bounds__(128, 4) k(double *out, const double *in, int iters) {
  __shared__ double2 coef[4][20];
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) < 19) coef[w][threadIdx.x & 31] = make_double2(in[threadIdx.x & 31], in[(threadIdx.x & 31) + 1]);
  __syncthreads();
  double za = in[0] + 1e-3 * threadIdx.x, zb = in[1] + 1e-3 * threadIdx.x;
  double acc = 0;
  for (int i = 0; i < iters; i++) {
    if (F == 0) {
      double Aa = coef[w][18].x, Ab = Aa, Pa = 0.0, Pb = 0.0;
#pragma unroll
      for (int n = 17; n >= 0; n--) {
        const double ak = coef[w][n].x;
        Pa = fma(Pa, za, Aa); Aa = fma(Aa, za, ak);
        Pb = fma(Pb, zb, Ab); Ab = fma(Ab, zb, ak);
      }
      acc += Aa * Pa + Ab * Pb; 
    } else {
      const double2 top = coef[w][18];
      double Aa = top.x, Ab = top.x, Pa = top.y, Pb = top.y;
      if (F == 7) {   // skew: first round ordered by true dependencies Pa -> Pb -> Ab -> Aa
        const double2 q = coef[w][17];
        Pa = fma(Pa, za, q.y);
        Pb = fma(Pb, zb, q.y + 0.0 * Pa);
        Ab = fma(Ab, zb, q.x + 0.0 * Pb);
        Aa = fma(Aa, za, q.x + 0.0 * Ab);
      }
#pragma unroll
      for (int n = (F == 7 ? 16 : 17); n >= 0; n--) {
        const double2 q = coef[w][n];
        if (n > 0) Pa = fma(Pa, za, q.y);
        if (n > 0) Pb = fma(Pb, zb, q.y);
        Ab = fma(Ab, zb, q.x);
        Aa = fma(Aa, za, q.x);
      }
      acc += Aa * Pa + Ab * Pb; 
    }
    za += 1e-9; zb += 1e-9;
  }
  if (acc == 12345.678) out[0] = acc;
}
template <int F>
void run(double *d, double *in) {
  int threads = 128, blocks = 148 * 4, iters = 8192;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<F><<<blocks, threads>>>(d, in, iters);
  cudaEventRecord(e0);
  k<F><<<blocks, threads>>>(d, in, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double trip = (double) iters * 2 * (threads / 32) * blocks;
  const double cyc = ms * 1e-3 * 1.965e9 * 592 / trip;
  printf("F%d : %.1f SMSP-cycles per warp-triplet = %.2f per DFMA\n", F, cyc, cyc / (F==0?36.0:35.0));
}
int main() {
  double *d, *in; cudaMalloc(&d, 8); cudaMalloc(&in, 512);
  double h[64]; for (int i = 0; i < 64; i++) h[i] = 0.3 + 0.01 * i;
  cudaMemcpy(in, h, 512, cudaMemcpyHostToDevice);
  run<0>(d, in); run<6>(d, in); run<7>(d, in);
  return 0;
}

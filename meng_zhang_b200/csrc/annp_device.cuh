// Shared device-side declarations for libannp_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/annp_b200.h"

#define ANNP_NEIGHMASK 0x1FFFFFFF
#define ANNP_MAX_TYPES 8

// Parameter block resident in global memory (one per handle); kernels stage what they need in smem.
struct DevParams {
  int ntypes, nelements, nlayers /* = ntl-1 */, nnod, nsf, npsf, ntsf;
  int flagact[ANNP_B200_MAX_LAYERS];
  int map[ANNP_MAX_TYPES + 1];
  double cutsq[(ANNP_MAX_TYPES + 1) * (ANNP_MAX_TYPES + 1)];
  double rcinv[(ANNP_MAX_TYPES + 1) * (ANNP_MAX_TYPES + 1)];   // 1/sqrt(cutsq): cutoff-function range
  double cut, two_over_cut;                                    // potential-file Rc: Chebyshev argument
  double e_scale, e_shift, e_atom;
  double sf_scale[ANNP_B200_MAX_SF];                           // s_n
  double sf_avg[ANNP_B200_MAX_SF];
  int w_per_elem, b_per_elem;                                  // doubles per element
  int w_off[ANNP_B200_MAX_LAYERS], b_off[ANNP_B200_MAX_LAYERS];
  // weights / biases follow in separate device arrays
  const double *weights;
  const double *bias;
  // [ntsf][ntsf] row-major: monomial coefficient a_k (in z = cos theta) of T_n((z+1)/2) is cheb2mono[k*ntsf+n]
  const double *cheb2mono;
};

// counters written by the force kernel (one block of 8 x 8 bytes)
struct DevCounters {
  unsigned long long work;        // dynamic atom scheduler
  unsigned long long sum_neigh;   // sum of in-cutoff neighbours
  unsigned long long sum_trip;    // sum of N(N-1)/2
  int max_neigh;                  // max in-cutoff neighbours
  int overflow;                   // an atom exceeded the smem capacity
};

struct ForceArgs {
  const DevParams *prm;
  const double4 *xq;          // [nall] x,y,z,(type as double)
  const int *ilist;           // [inum]
  const long long *row_off;   // [inum+1]
  const int *nbr;             // [total]
  double4 *fpair;             // [total] force on the neighbour of every list entry (0 if outside Rc)
  double4 *fself;             // [inum]  x,y,z = -sum_j Fj ; w = E_i
  double *vir_c;              // [inum][6] per-centre pair virial (may be null)
  double *vpair;              // [total][6] per-pair virial for vatom (may be null)
  double *G_dbg, *dEdG_dbg;   // [inum][nsf] (may be null)
  DevCounters *cnt;
  int inum;
  int capacity;               // smem neighbour slots per warp
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

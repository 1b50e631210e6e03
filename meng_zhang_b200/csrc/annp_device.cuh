// Shared device-side declarations for libannp_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/annp_b200.h"

#define ANNP_NEIGHMASK 0x1FFFFFFF
#define ANNP_MAX_TYPES 8
// fixed-point force accumulation: 1 unit = 2^-43 eV/A (1.1e-13); int64 holds +-1.0e6 eV/A; contributions are refused
// (bad_force flag) from 2^18 eV/A on, so 4 billion of them could be summed before the accumulator wraps
#define ANNP_FIX_BITS 43
#define ANNP_FIX_LIMIT_BITS 18

// Parameter block resident in global memory (one per handle); kernels stage what they need in smem.
struct DevParams {
  int ntypes, nelements, nlayers /* = ntl-1 */, nnod, nsf, npsf, ntsf;
  int nout;                                                    // rows of the last layer: 1 (ANNP), 2 (ANNA-ADP: d2, q2)
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
  // [ntsf][ntsf] row-major: T_n((z+1)/2) = sum_j blk2cheb[j*ntsf+n] psi_j(z), psi_{4b+i}(z) = T_{4b}(z) z^i (forward-pass basis)
  const double *blk2cheb;
  // ANNP_B200_VARIANT_NI: Behler-Parrinello coefficients (ni/src/pair_annp.cpp:686-767)
  int variant;
  double rad_eta[ANNP_B200_MAX_SF], rad_rc;                    // radial: eta_m, common Rc (Bohr)
  double ang_eta[ANNP_B200_MAX_SF], ang_lambda[ANNP_B200_MAX_SF], ang_zeta[ANNP_B200_MAX_SF], ang_rc;
  int bp_layout;                                               // 1: product-structured table -> annp_bp_fast_kernel (annp_bp_layout)
  // ANNP_B200_VARIANT_ANNA_ADP: the 17 global ADP parameters and the energy offset (pair_anna_adp.cpp:97-103)
  double gparams[17], e_base;
};

// counters written by the force kernel (one block of 8 x 8 bytes)
struct DevCounters {
  // ---- reset before every step (everything in front of `overflow`)
  unsigned long long work;        // dynamic atom scheduler, first pass
  unsigned long long work2;       // dynamic atom scheduler, overflow pass
  unsigned long long sum_neigh;   // sum of in-cutoff neighbours
  unsigned long long sum_trip;    // sum of N(N-1)/2
  int max_neigh;                  // max in-cutoff neighbours
  int ovf_count;                  // atoms of this step whose in-cutoff neighbours did not fit the first-pass tile: they are
                                  // listed in ForceArgs::ovf_list and redone by the overflow pass with a tile of the list's
                                  // longest row, so the step's results are complete whatever the atoms did inside the skin
  // ---- sticky: only annp_b200_get_stats / the host-mode compute read AND clear them, reporting the failure
  int overflow;                   // an atom has more in-cutoff neighbours than the LARGEST tile (ANNP_B200_MAX_NEIGH)
  int bad_force;                  // a neighbour force left the fixed-point range (|F| >= 2^ANNP_FIX_LIMIT_BITS eV/A) or is NaN
  unsigned long long ovf_total;   // atoms redone by the overflow pass since the counters were last read (statistics)
  unsigned long long stage_clk[8];// ANNP_STAGE_CLOCKS builds: SM cycles per stage summed over warps (filter, radial, forward,
                                  // reduce+MLP, backward, force, scheduler)
};

struct ForceArgs {
  const DevParams *prm;
  const double4 *xq;          // [nall] x,y,z,(type as double)
  const int *ilist;           // [inum]
  const long long *row_off;   // [inum+1]
  const int *nbr;             // [total]
  double4 *fpair;             // [total] force on the neighbour of every list entry (0 if outside Rc); ordered-gather scatter
  long long *facc;            // [nall][3] fixed-point force accumulators (units of 2^-ANNP_FIX_BITS eV/A); when non-null the
                              // neighbour forces are added here with 64-bit integer atomics (exact, order independent)
                              // and fpair is not touched
  double4 *fself;             // [inum]  x,y,z = -sum_j Fj ; w = E_i
  double *vir_c;              // [inum][6] per-centre pair virial (may be null)
  double *vpair;              // [total][6] per-pair virial for vatom (may be null)
  double *G_dbg, *dEdG_dbg;   // [inum][nsf] (may be null)
  DevCounters *cnt;
  int inum;
  int capacity;               // smem neighbour slots per warp
  // work items: first pass = all inum centres (work_list null); overflow pass = the ovf_count centres listed by the first
  // pass (work_list = that list, work_count = its device-side length)
  const int *work_list;
  const int *work_count;
  unsigned long long *work_ctr;
  int *ovf_list;              // [inum] first pass: centres that need the overflow pass; null in the overflow pass itself
  // ANNA-ADP: the 17 global ADP parameters ride in the kernel's parameter block (constant bank), so the tail reads them
  // as instruction operands instead of holding 34 registers
  double gp[17];
  // Peer scatter (multi-GPU, annp_b200_peer_*): the force on a GHOST neighbour is added straight to the accumulator of
  // its OWNER rank over NVLink instead of to this rank's ghost row - the reverse halo exchange fused into the force kernel.
  //   peer_facc[r]  base of rank r's facc (IPC-mapped; entry of this rank = its own facc), null = feature off
  //   ghost_rank / ghost_index [nghost]  owner rank and the owner's local index of ghost atom nlocal + g
  long long *const *peer_facc;
  const int *ghost_rank;
  const int *ghost_index;
  int peer_nlocal;
};

// number of work items of this launch and the centre index of item `item`
__device__ __forceinline__ unsigned long long annp_work_items(const ForceArgs &a) {
  return a.work_count ? (unsigned long long) *a.work_count : (unsigned long long) a.inum;
}
__device__ __forceinline__ int annp_work_centre(const ForceArgs &a, unsigned long long item) {
  return a.work_list ? a.work_list[item] : (int) item;
}
// the in-cutoff neighbours of centre ii do not fit this launch's tile (call from one lane)
__device__ __forceinline__ void annp_note_overflow(const ForceArgs &a, int ii) {
  if (a.ovf_list) a.ovf_list[atomicAdd(&a.cnt->ovf_count, 1)] = ii;
  else atomicExch(&a.cnt->overflow, 1);
}

// activation tables of the reference copies: Fe pair_annp.cpp:709-739, Ni ni/src/pair_annp.cpp:786-807,
// ANNA-ADP pair_anna_adp.cpp:694-718 (3 and 4 = 1.7 tanh(0.3 x); only h is used there)
__device__ __forceinline__ void annp_activation(int variant, int flag, double z, double &h, double &hd) {
  const double ca = 1.7159, cb = 0.666666666666667, cc = 0.1;
  double t;
  if (variant == ANNP_B200_VARIANT_ANNA_ADP && flag >= 3) {
    t = tanh(0.3 * z);
    h = 1.7 * t;
    hd = 1.7 * 0.3 * (1.0 - t * t);
    return;
  }
  switch (flag) {
    case 0: h = z; hd = 1.0; break;
    case 1: h = tanh(z); hd = 1.0 - h * h; break;
    case 2: h = 1.0 / (1.0 + exp(z)); hd = h * (1.0 - h); break;
    case 3:
      if (variant == ANNP_B200_VARIANT_NI) { h = tanh(z); hd = 1.0 - h * h; }
      else { t = tanh(cb * z); h = ca * t; hd = ca * (1.0 - t * t) * cb; }
      break;
    default:
      if (variant == ANNP_B200_VARIANT_NI) { h = tanh(z); hd = 1.0 - h * h; }
      else { t = tanh(cb * z); h = ca * t + cc * z; hd = ca * (1.0 - t * t) * cb + cc; }
      break;
  }
}

// Per-atom MLP executed by one warp: forward pass and reverse-mode backprop of the raw output with respect to the
// (normalised) descriptor.  sG [nsf] in; sdE [nsf] out; returns the raw network output.  sH/sHd: [nl][nnod],
// sDel: [2][nnod] scratch.  Same maths as the reference's forward-mode Jacobian (pair_annp.cpp:741-804).
__device__ __forceinline__ double annp_mlp_warp(const DevParams &P, const double *We, const double *Be, const double *sG,
                                                double *sdE, double *sH, double *sHd, double *sDel, int lane) {
  const int nsf = P.nsf, nnod = P.nnod, nl = P.nlayers;
  const double *in = sG;
  for (int l = 0; l < nl; l++) {
    const int nr = (l == nl - 1) ? 1 : nnod;
    const int nc = (l == 0) ? nsf : nnod;
    const double *W = We + P.w_off[l];
    if (lane < nr) {
      double z = 0.0;
      for (int cidx = 0; cidx < nc; cidx++) z = fma(W[lane * nc + cidx], in[cidx], z);
      z += Be[P.b_off[l] + lane];
      double h, hd;
      annp_activation(P.variant, P.flagact[l], z, h, hd);
      sH[l * nnod + lane] = h;
      sHd[l * nnod + lane] = hd;
    }
    __syncwarp();
    in = sH + l * nnod;
  }
  const double out = sH[(nl - 1) * nnod];
  double *dcur = sDel, *dprev = sDel + nnod;        // delta_l[r] = d out / d z_l[r]
  if (lane == 0) dcur[0] = sHd[(nl - 1) * nnod];
  __syncwarp();
  for (int l = nl - 1; l >= 1; l--) {
    const int nr = (l == nl - 1) ? 1 : nnod;
    const double *W = We + P.w_off[l];               // [nr][nnod]
    if (lane < nnod) {
      double s = 0.0;
      for (int r = 0; r < nr; r++) s = fma(W[r * nnod + lane], dcur[r], s);
      dprev[lane] = s * sHd[(l - 1) * nnod + lane];
    }
    __syncwarp();
    double *tmp = dcur; dcur = dprev; dprev = tmp;
  }
  const int nr0 = (nl == 1) ? 1 : nnod;
  const double *W0 = We + P.w_off[0];                // [nr0][nsf]
  for (int n = lane; n < nsf; n += 32) {
    double s = 0.0;
    for (int r = 0; r < nr0; r++) s = fma(W0[r * nsf + n], dcur[r], s);
    sdE[n] = s;
  }
  __syncwarp();
  return out;
}

// Forward pass only (ANNA-ADP, pair_anna_adp.cpp:720-751): the nout outputs are left in sH[(nl-1)*nnod + k].
__device__ __forceinline__ void annp_mlp_forward_warp(const DevParams &P, const double *We, const double *Be, const double *sG,
                                                      double *sH, int lane) {
  const int nsf = P.nsf, nnod = P.nnod, nl = P.nlayers;
  const double *in = sG;
  for (int l = 0; l < nl; l++) {
    const int nr = (l == nl - 1) ? P.nout : nnod;
    const int nc = (l == 0) ? nsf : nnod;
    const double *W = We + P.w_off[l];
    if (lane < nr) {
      double z = 0.0;
      for (int cidx = 0; cidx < nc; cidx++) z = fma(W[lane * nc + cidx], in[cidx], z);
      z += Be[P.b_off[l] + lane];
      double h, hd;
      annp_activation(P.variant, P.flagact[l], z, h, hd);
      sH[l * nnod + lane] = h;
    }
    __syncwarp();
    in = sH + l * nnod;
  }
}

// add (fx, fy, fz) to the fixed-point accumulator of atom j; false if a component is outside the representable range
// (integer addition is associative and commutative, so the sum does not depend on which GPU adds first: the owner's
// accumulator ends up with the same integer whether a ghost contribution arrives by peer atomic or by reverse exchange)
__device__ __forceinline__ bool annp_fix_add(const ForceArgs &a, int j, double fx, double fy, double fz) {
  const double lim = (double) (1LL << ANNP_FIX_LIMIT_BITS), sc = (double) (1LL << ANNP_FIX_BITS);
  const bool ok = fabs(fx) < lim && fabs(fy) < lim && fabs(fz) < lim;      // false for NaN as well
  if (ok) {
    const unsigned long long ix = (unsigned long long) __double2ll_rn(fx * sc), iy = (unsigned long long) __double2ll_rn(fy * sc),
                             iz = (unsigned long long) __double2ll_rn(fz * sc);
    if (a.peer_facc && j >= a.peer_nlocal) {      // ghost: its owner's accumulator, system-scope atomics over NVLink
      const int g = j - a.peer_nlocal;
      unsigned long long *t = reinterpret_cast<unsigned long long *>(a.peer_facc[a.ghost_rank[g]] + 3 * (size_t) a.ghost_index[g]);
      atomicAdd_system(t, ix);
      atomicAdd_system(t + 1, iy);
      atomicAdd_system(t + 2, iz);
    } else {
      unsigned long long *t = reinterpret_cast<unsigned long long *>(a.facc + 3 * (size_t) j);
      atomicAdd(t, ix);
      atomicAdd(t + 1, iy);
      atomicAdd(t + 2, iz);
    }
  }
  return ok;
}

// Sum 32 per-lane values of EVERY lane across the warp at once: on return lane n holds the warp total of w[n].  In the
// round with lane mask m every lane keeps the half of its values that belongs to its side of the mask and hands the other
// half over, so the number of live values halves each round: 16 + 8 + 4 + 2 + 1 = 31 shuffle-adds instead of 5 per value.
// Fixed order -> deterministic.  (w is consumed.)
__device__ __forceinline__ double warp_transpose_sum32(double (&w)[32], int lane) {
#pragma unroll
  for (int m = 16, cnt = 32; m >= 1; m >>= 1, cnt >>= 1) {
    const bool up = (lane & m) != 0;
#pragma unroll
    for (int n = 0; n < cnt / 2; n++) {
      const double send = up ? w[n] : w[n + cnt / 2];
      const double keep = up ? w[n + cnt / 2] : w[n];
      w[n] = keep + __shfl_xor_sync(0xffffffffu, send, m);
    }
  }
  return w[0];
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

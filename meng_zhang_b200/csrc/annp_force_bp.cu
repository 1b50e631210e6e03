// annp_force_bp.cu -- force kernel for the Ni copy of the reference pair style (ANNP_B200_VARIANT_NI):
// Behler-Parrinello radial G2 and narrow angular G3 in Bohr units.  Reference: annp-gpu-lammps/ni/src/pair_annp.cpp
//   compute 74-212, annp_symmetry_pair 686-708, annp_symmetry_trip 710-767, feed_forward 809-871.
//
// The descriptor couples all three sides of a triplet (r_ij, r_ik, r_jk), has ~18 neighbours inside its 3.9 A
// cutoff (60 contributing triplets in fcc Ni) and costs ~3e4 flop per atom, so the kernel is organised for
// simplicity and determinism, not for the FP64 pipe: one warp per centre atom;
//   1. filter the list row to r*1.889726 < Rc (list order kept), cache x_ij, r, fc, dfc in shared memory
//   2. forward: lane <-> first neighbour j, loop k > j (list order, as the reference's kk > jj) -> G
//   3. MLP + backprop (shared with the Fe kernel); E_i is the raw network output
//   4. backward: lane <-> neighbour a, loop over ALL partners b; lane a accumulates only the force on a, in the
//      role the reference gives it (first or second member of the pair by list order) -> no scatter, fixed order
// Reference quirks kept on purpose (SURVEY.md section 7 "hard parts"):
//   * no pair-level list cutoff test and no r^2 < 1e-12 test (pair_annp.cpp:126-140)
//   * d(r_ij^2+r_ik^2+r_jk^2)/dx uses r_ik where r_jk is meant: term2_drj = 2(r_ij dr_j + r_IK dr_jk),
//     term2_drk = 2(r_ik dr_k - r_IK dr_jk)                                  (pair_annp.cpp:734-735)
//   * 1 + lambda cos(theta) <= 0 skips the term (discontinuous at cos = -+1)    (pair_annp.cpp:748-751)
//   * forces are scaled by 51.422515, the virial tally is not                   (pair_annp.cpp:188-198)
#include "annp_device.cuh"

namespace {

constexpr int kWarps = 4;
constexpr double kPi = 3.14159265358979323846;
constexpr double CFLENGTH = 1.889726;      // ni/src/pair_annp.h:70
constexpr double CFFORCE = 51.422515;      // ni/src/pair_annp.h:71

struct Nb {          // one neighbour inside Rc (Angstrom geometry, Bohr cutoff function)
  double x, y, z;    // x_i - x_j
  double r;          // |x_ij| in Angstrom
  double fc, dfc;    // cutoff function of r*CFLENGTH and its derivative with respect to r*CFLENGTH
};

__device__ __forceinline__ void bp_fc(double rm, double Rc, double &fc, double &dfc) {
  const double a = kPi / Rc * rm;        // pair_annp.cpp:643-647
  double s, c;
  sincos(a, &s, &c);
  fc = 0.5 * (c + 1.0);
  dfc = -0.5 * kPi / Rc * s;
}

// Contribution of the ordered pair (first = a1, second = a2; a1 before a2 in list order) to dOut/dx of ONE of its
// members.  `for_first` selects which member's gradient is returned (gx, gy, gz are accumulated).
// c[n] = dOut/dG_n / (max_n - min_n) for the angular block.
__device__ __forceinline__ void bp_triplet_grad(const DevParams &P, const double *c, const Nb &j, const Nb &k, bool for_first,
                                                double &gx, double &gy, double &gz) {
  const int ntsf = P.ntsf;
  const double Rc = P.ang_rc;
  const double rij_m = j.r * CFLENGTH, rik_m = k.r * CFLENGTH;
  if (!(rij_m < Rc && rik_m < Rc)) return;
  const double xjk0 = k.x - j.x, xjk1 = k.y - j.y, xjk2 = k.z - j.z;       // x_j - x_k = (x_i - x_k) - (x_i - x_j)
  const double rjk = sqrt(xjk0 * xjk0 + xjk1 * xjk1 + xjk2 * xjk2);
  const double rjk_m = rjk * CFLENGTH;
  if (!(rjk_m < Rc)) return;
  double fcjk, dfcjk;
  bp_fc(rjk_m, Rc, fcjk, dfcjk);
  const double rijinv = 1.0 / j.r, rikinv = 1.0 / k.r;
  const double cos_theta = (j.x * rijinv) * (k.x * rikinv) + (j.y * rijinv) * (k.y * rikinv) + (j.z * rijinv) * (k.z * rikinv);
  const double r2sum = rij_m * rij_m + rik_m * rik_m + rjk_m * rjk_m;
  const double term_fc = j.fc * k.fc * fcjk;
  // per-direction geometric factors of the member we differentiate with respect to
  const double xj[3] = {j.x, j.y, j.z}, xk[3] = {k.x, k.y, k.z}, xjk[3] = {xjk0, xjk1, xjk2};
  const double B = j.r * k.r;
  double dct[3], t2[3], t3[3];
#pragma unroll
  for (int m = 0; m < 3; m++) {
    const double dr_dj = -xj[m] / j.r, dr_dk = -xk[m] / k.r, dr_djk = xjk[m] / rjk;      // annp_dr_dij(1,..), (1,..), (0,..)
    if (for_first) {
      dct[m] = (-1.0) * xk[m] / B + cos_theta / (j.r * j.r) * xj[m];                     // annp_dct_djk
      t2[m] = 2.0 * (rij_m * dr_dj + rik_m * dr_djk);
      t3[m] = k.fc * (j.dfc * dr_dj * fcjk + j.fc * dfcjk * dr_djk);
    } else {
      dct[m] = (-1.0) * xj[m] / B + cos_theta / (k.r * k.r) * xk[m];
      t2[m] = 2.0 * (rik_m * dr_dk - rik_m * dr_djk);
      t3[m] = j.fc * (k.dfc * dr_dk * fcjk - k.fc * dfcjk * dr_djk);
    }
  }
  for (int n = 0; n < ntsf; n++) {
    const double eta = P.ang_eta[n], lambda = P.ang_lambda[n], zeta = P.ang_zeta[n];
    const double flag = 1.0 + lambda * cos_theta;
    if (flag <= 0.0) continue;
    const double term_cot = pow(2.0, 1.0 - zeta) * pow(flag, zeta);
    const double term_exp = exp(-eta * r2sum);
    const double term1 = lambda * term_cot * term_exp * term_fc * zeta / flag / CFLENGTH;
    const double term3 = term_cot * term_exp;
    const double term2 = term3 * term_fc * eta;
    const double cn = c[n];
    gx = fma(cn, term1 * dct[0] - term2 * t2[0] + term3 * t3[0], gx);
    gy = fma(cn, term1 * dct[1] - term2 * t2[1] + term3 * t3[1], gy);
    gz = fma(cn, term1 * dct[2] - term2 * t2[2] + term3 * t3[2], gz);
  }
}

__global__ void __launch_bounds__(kWarps * 32) annp_bp_force_kernel(const ForceArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const DevParams &P = *a.prm;
  const int C = a.capacity;
  const int nsf = P.nsf, npsf = P.npsf, nnod = P.nnod, nl = P.nlayers;
  const int wtot = P.nelements * P.w_per_elem, btot = P.nelements * P.b_per_elem;

  double *sW = reinterpret_cast<double *>(smem_raw);
  double *sBias = sW + wtot;
  double *blk_end = sBias + btot;
  for (int t = threadIdx.x; t < wtot; t += blockDim.x) sW[t] = P.weights[t];
  for (int t = threadIdx.x; t < btot; t += blockDim.x) sBias[t] = P.bias[t];
  const size_t per_warp_doubles = (size_t) 6 * C + 3 * nsf + (size_t) 2 * nl * nnod + 2 * nnod;
  size_t warp_bytes = per_warp_doubles * sizeof(double) + (size_t) C * sizeof(int);
  warp_bytes = (warp_bytes + 15) & ~(size_t) 15;
  const size_t blk_bytes = ((size_t) ((unsigned char *) blk_end - smem_raw) + 15) & ~(size_t) 15;
  unsigned char *wbase = smem_raw + blk_bytes + (size_t) warp * warp_bytes;
  Nb *sN = reinterpret_cast<Nb *>(wbase);
  double *sG = reinterpret_cast<double *>(sN + C);
  double *sdE = sG + nsf;
  double *sCn = sdE + nsf;                            // dOut/dG_n / range_n
  double *sH = sCn + nsf;
  double *sHd = sH + nl * nnod;
  double *sDel = sHd + nl * nnod;
  int *spos = reinterpret_cast<int *>(sDel + 2 * nnod);
  __syncthreads();

  const double Rc_rad = P.rad_rc, Rc_ang = P.ang_rc;
  const double Rc_max = fmax(Rc_rad, Rc_ang);

  const unsigned long long nwork = annp_work_items(a);
  if (a.work_list && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&a.cnt->ovf_total, nwork);   // overflow pass: statistics
  for (;;) {
    unsigned long long item = 0;
    if (lane == 0) item = atomicAdd(a.work_ctr, 1ull);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= nwork) break;
    const int ii = annp_work_centre(a, item);
    const int i = a.ilist[ii];
    const double4 xi = a.xq[i];
    const int ti = (int) xi.w;
    const long long p0 = a.row_off[ii];
    const int L = (int) (a.row_off[ii + 1] - p0);

    // ---- 1. filter (list order kept)
    int N = 0;
    for (int base = 0; base < L; base += 32) {
      const int q = base + lane;
      const bool valid = q < L;
      bool in = false;
      double dx = 0, dy = 0, dz = 0, r = 0;
      if (valid) {
        const int j = a.nbr[p0 + q] & ANNP_NEIGHMASK;
        const double4 xj = a.xq[j];
        dx = xi.x - xj.x; dy = xi.y - xj.y; dz = xi.z - xj.z;
        r = sqrt(dx * dx + dy * dy + dz * dz);
        in = (r * CFLENGTH < Rc_max);
      }
      const unsigned mask = __ballot_sync(0xffffffffu, in);
      const int slot = N + __popc(mask & ((1u << lane) - 1u));
      if (in && slot < C) {
        Nb nb;
        nb.x = dx; nb.y = dy; nb.z = dz; nb.r = r;
        bp_fc(r * CFLENGTH, Rc_ang, nb.fc, nb.dfc);
        sN[slot] = nb;
        spos[slot] = q;
      } else if (valid) {
        if (!a.facc) a.fpair[p0 + q] = make_double4(0.0, 0.0, 0.0, 0.0);
        if (a.vpair) {
          double *vp = a.vpair + (size_t) (p0 + q) * 6;
#pragma unroll
          for (int k = 0; k < 6; k++) vp[k] = 0.0;
        }
      }
      N += __popc(mask);
    }
    if (lane == 0 && !a.work_list) {
      atomicMax(&a.cnt->max_neigh, N);
      atomicAdd(&a.cnt->sum_neigh, (unsigned long long) N);
      atomicAdd(&a.cnt->sum_trip, (unsigned long long) N * (unsigned long long) (N > 0 ? N - 1 : 0) / 2ull);
    }
    if (N > C) {
      if (lane == 0) { annp_note_overflow(a, ii); a.fself[ii] = make_double4(0.0, 0.0, 0.0, 0.0); }
      if (!a.facc) for (int q = lane; q < L; q += 32) a.fpair[p0 + q] = make_double4(0.0, 0.0, 0.0, 0.0);
      __syncwarp();
      continue;
    }
    __syncwarp();

    // ---- 2. forward: G_n
    for (int n = lane; n < nsf; n += 32) sG[n] = 0.0;
    __syncwarp();
    for (int n = 0; n < nsf; n++) {
      // lane-parallel over first members, one descriptor component at a time keeps registers small; every lane
      // sums its own (j, k>j) terms in list order, then a fixed butterfly adds the lanes
      double acc = 0.0;
      for (int j = lane; j < N; j += 32) {
        const Nb nj = sN[j];
        const double rij_m = nj.r * CFLENGTH;
        if (n < npsf) {
          if (rij_m < Rc_rad) {
            double fc, dfc;
            bp_fc(rij_m, Rc_rad, fc, dfc);
            acc += exp(-P.rad_eta[n] * rij_m * rij_m) * fc;               // pair_annp.cpp:697-703
          }
        } else {
          const int na = n - npsf;
          const double eta = P.ang_eta[na], lambda = P.ang_lambda[na], zeta = P.ang_zeta[na];
          if (!(rij_m < Rc_ang)) continue;
          const double rijinv = 1.0 / nj.r;
          for (int k = j + 1; k < N; k++) {
            const Nb nk = sN[k];
            const double rik_m = nk.r * CFLENGTH;
            if (!(rik_m < Rc_ang)) continue;
            const double e0 = nk.x - nj.x, e1 = nk.y - nj.y, e2 = nk.z - nj.z;
            const double rjk_m = sqrt(e0 * e0 + e1 * e1 + e2 * e2) * CFLENGTH;
            if (!(rjk_m < Rc_ang)) continue;
            double fcjk, dfcjk;
            bp_fc(rjk_m, Rc_ang, fcjk, dfcjk);
            const double rikinv = 1.0 / nk.r;
            const double cos_theta = (nj.x * rijinv) * (nk.x * rikinv) + (nj.y * rijinv) * (nk.y * rikinv) + (nj.z * rijinv) * (nk.z * rikinv);
            const double flag = 1.0 + lambda * cos_theta;
            if (flag <= 0.0) continue;
            const double r2sum = rij_m * rij_m + rik_m * rik_m + rjk_m * rjk_m;
            acc += pow(2.0, 1.0 - zeta) * pow(flag, zeta) * exp(-eta * r2sum) * (nj.fc * nk.fc * fcjk);   // pair_annp.cpp:752-756
          }
        }
      }
      acc = warp_sum(acc);
      if (lane == 0) sG[n] = (acc - P.sf_avg[n]) * P.sf_scale[n];       // (G - min)/(max - min), pair_annp.cpp:168-170
    }
    __syncwarp();

    // ---- 3. MLP
    const int elem = P.map[ti];
    const double out = annp_mlp_warp(P, sW + elem * P.w_per_elem, sBias + elem * P.b_per_elem, sG, sdE, sH, sHd, sDel, lane);
    const double e_i = out;                                             // pair_annp.cpp:858-860: raw output
    if (a.G_dbg) for (int n = lane; n < nsf; n += 32) { a.G_dbg[(size_t) ii * nsf + n] = sG[n]; a.dEdG_dbg[(size_t) ii * nsf + n] = sdE[n]; }
    for (int n = lane; n < nsf; n += 32) sCn[n] = sdE[n] * P.sf_scale[n];      // dE_dG[n] / sf_max[n] (range)
    __syncwarp();

    // ---- 4. backward: the force on neighbour s, all its pairs, fixed order
    double fix = 0, fiy = 0, fiz = 0;
    double v0 = 0, v1 = 0, v2 = 0, v3 = 0, v4 = 0, v5 = 0;
    for (int s = lane; s < N; s += 32) {
      const Nb ns = sN[s];
      double gx = 0, gy = 0, gz = 0;                   // d out / d x_s  (Bohr^-1 units of the reference)
      const double rm = ns.r * CFLENGTH;
      if (rm < Rc_rad) {                               // annp_symmetry_pair
        double fc, dfc;
        bp_fc(rm, Rc_rad, fc, dfc);
        for (int m = 0; m < npsf; m++) {
          const double eta = P.rad_eta[m];
          const double term1 = exp(-eta * rm * rm);
          const double term2 = term1 * (-fc * 2.0 * eta * rm + dfc);
          const double cm = sCn[m];
          gx = fma(cm, term2 * (-ns.x / ns.r), gx);
          gy = fma(cm, term2 * (-ns.y / ns.r), gy);
          gz = fma(cm, term2 * (-ns.z / ns.r), gz);
        }
      }
      for (int b = 0; b < N; b++) {
        if (b == s) continue;
        const Nb nb = sN[b];
        if (s < b) bp_triplet_grad(P, sCn + npsf, ns, nb, true, gx, gy, gz);     // s is the reference's j
        else bp_triplet_grad(P, sCn + npsf, nb, ns, false, gx, gy, gz);          // s is the reference's k
      }
      const double Fx = -gx, Fy = -gy, Fz = -gz;       // Fj of the reference before CFFORCE
      const int q = spos[s];
      if (a.facc) {     // fixed-point scatter (annp_device.cuh); else the per-entry buffer of the ordered gather
        if (!annp_fix_add(a, a.nbr[p0 + q] & ANNP_NEIGHMASK, Fx * CFFORCE, Fy * CFFORCE, Fz * CFFORCE)) atomicExch(&a.cnt->bad_force, 1);
      } else {
        a.fpair[p0 + q] = make_double4(Fx * CFFORCE, Fy * CFFORCE, Fz * CFFORCE, 0.0);
      }
      fix -= Fx * CFFORCE; fiy -= Fy * CFFORCE; fiz -= Fz * CFFORCE;
      if (a.vir_c || a.vpair) {
        const double w0 = -ns.x * Fx, w1 = -ns.y * Fy, w2 = -ns.z * Fz, w3 = -ns.x * Fy, w4 = -ns.x * Fz, w5 = -ns.y * Fz;
        v0 += w0; v1 += w1; v2 += w2; v3 += w3; v4 += w4; v5 += w5;
        if (a.vpair) {
          double *vp = a.vpair + (size_t) (p0 + q) * 6;
          vp[0] = w0; vp[1] = w1; vp[2] = w2; vp[3] = w3; vp[4] = w4; vp[5] = w5;
        }
      }
    }
    fix = warp_sum(fix); fiy = warp_sum(fiy); fiz = warp_sum(fiz);
    if (lane == 0) a.fself[ii] = make_double4(fix, fiy, fiz, e_i);
    if (a.vir_c) {
      v0 = warp_sum(v0); v1 = warp_sum(v1); v2 = warp_sum(v2);
      v3 = warp_sum(v3); v4 = warp_sum(v4); v5 = warp_sum(v5);
      if (lane == 0) {
        double *vc = a.vir_c + (size_t) ii * 6;
        vc[0] = v0; vc[1] = v1; vc[2] = v2; vc[3] = v3; vc[4] = v4; vc[5] = v5;
      }
    }
    __syncwarp();
  }
}


// ---------------------------------------------------------------------------------------------------------------
// Fast path for coefficient tables with product structure (the shipped ni_annp_potential_2.ann):
//     n = (e * NZ + z) * 2 + l,   eta = eta_e (e < NE),  zeta = 2^kZetaLog2[z],  lambda = -1 (l = 0), +1 (l = 1)
// annp_bp_layout() verifies this at init; any other table runs the generic kernel above.
//   * (1 + lambda cos)^zeta by repeated squaring and one exp per distinct eta instead of 2 pow + 1 exp per component
//   * the gradient of a triplet needs only three sums over the components,
//         S1 = sum_n c_n lambda_n zeta_n T_n / flag_n,  S2 = sum_n c_n eta_n T_n,  S3 = sum_n c_n T_n,
//     (T_n = 2^(1-zeta) flag^zeta exp(-eta R2)), followed by ONE vector assembly
//   * forward and backward are both organised "lane <-> neighbour a, loop over partners b" so every lane only ever
//     adds into its own registers: no scatter, no atomics, fixed summation order.
template <int NZ> __host__ __device__ constexpr int bp_zeta_log2(int z);
template <> __host__ __device__ constexpr int bp_zeta_log2<4>(int z) { return z == 0 ? 0 : (z == 1 ? 1 : (z == 2 ? 2 : 4)); }   // zeta = 1, 2, 4, 16

struct BpGeom {        // neighbour as cached by the fast kernel
  double ux, uy, uz;   // unit vector of x_i - x_j
  double r;            // Angstrom
  double fc, dfc;      // angular cutoff function of r*CFLENGTH
};

// primitives of the pair (a, b): returns false when the triplet does not contribute
template <int NE>
__device__ __forceinline__ bool bp_pair_prims(const BpGeom &A, const BpGeom &B, double Rc, double rcinv, const double *eta,
                                              double &cosv, double &rjk, double &fcjk, double &dfcjk, double (&E)[NE],
                                              double &ex, double &ey, double &ez) {
  // e = x_a - x_b (positions) = (x_i - x_b) - (x_i - x_a)
  ex = B.r * B.ux - A.r * A.ux; ey = B.r * B.uy - A.r * A.uy; ez = B.r * B.uz - A.r * A.uz;
  const double r2 = ex * ex + ey * ey + ez * ez;
  rjk = sqrt(r2);
  const double rjk_m = rjk * CFLENGTH;
  if (!(rjk_m < Rc)) return false;
  double sn, cs;
  sincospi(rjk_m * rcinv, &sn, &cs);
  fcjk = 0.5 * (cs + 1.0);
  dfcjk = -0.5 * kPi * rcinv * sn;
  cosv = A.ux * B.ux + A.uy * B.uy + A.uz * B.uz;
  const double ra = A.r * CFLENGTH, rb = B.r * CFLENGTH;
  const double r2sum = ra * ra + rb * rb + rjk_m * rjk_m;
#pragma unroll
  for (int e = 0; e < NE; e++) E[e] = exp(-eta[e] * r2sum);
  return true;
}

template <int NE, int NZ>
__global__ void __launch_bounds__(kWarps * 32, 3) annp_bp_fast_kernel(const ForceArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int NT = NE * NZ * 2;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const DevParams &P = *a.prm;
  const int C = a.capacity;
  const int nsf = P.nsf, npsf = P.npsf, nnod = P.nnod, nl = P.nlayers;
  const int wtot = P.nelements * P.w_per_elem, btot = P.nelements * P.b_per_elem;

  double *sW = reinterpret_cast<double *>(smem_raw);
  double *sBias = sW + wtot;
  double *blk_end = sBias + btot;
  for (int t = threadIdx.x; t < wtot; t += blockDim.x) sW[t] = P.weights[t];
  for (int t = threadIdx.x; t < btot; t += blockDim.x) sBias[t] = P.bias[t];
  const size_t per_warp_doubles = (size_t) 6 * C + 3 * nsf + (size_t) 2 * nl * nnod + 2 * nnod;
  size_t warp_bytes = per_warp_doubles * sizeof(double) + (size_t) C * sizeof(int);
  warp_bytes = (warp_bytes + 15) & ~(size_t) 15;
  const size_t blk_bytes = ((size_t) ((unsigned char *) blk_end - smem_raw) + 15) & ~(size_t) 15;
  unsigned char *wbase = smem_raw + blk_bytes + (size_t) warp * warp_bytes;
  BpGeom *sN = reinterpret_cast<BpGeom *>(wbase);
  double *sG = reinterpret_cast<double *>(sN + C);
  double *sdE = sG + nsf;
  double *sCn = sdE + nsf;
  double *sH = sCn + nsf;
  double *sHd = sH + nl * nnod;
  double *sDel = sHd + nl * nnod;
  int *spos = reinterpret_cast<int *>(sDel + 2 * nnod);
  __syncthreads();

  const double Rc_rad = P.rad_rc, Rc = P.ang_rc, rcinv = 1.0 / Rc;
  const double Rc_max = fmax(Rc_rad, Rc);
  double eta[NE];
#pragma unroll
  for (int e = 0; e < NE; e++) eta[e] = P.ang_eta[e * NZ * 2];

  const unsigned long long nwork = annp_work_items(a);
  if (a.work_list && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&a.cnt->ovf_total, nwork);   // overflow pass: statistics
  for (;;) {
    unsigned long long item = 0;
    if (lane == 0) item = atomicAdd(a.work_ctr, 1ull);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= nwork) break;
    const int ii = annp_work_centre(a, item);
    const int i = a.ilist[ii];
    const double4 xi = a.xq[i];
    const int ti = (int) xi.w;
    const long long p0 = a.row_off[ii];
    const int L = (int) (a.row_off[ii + 1] - p0);

    // ---- 1. filter (list order kept): r * CFLENGTH < Rc, no pair-level list cutoff (ni/src/pair_annp.cpp:126-140)
    int N = 0;
    for (int base = 0; base < L; base += 32) {
      const int q = base + lane;
      const bool valid = q < L;
      bool in = false;
      double dx = 0, dy = 0, dz = 0, r = 0;
      if (valid) {
        const int j = a.nbr[p0 + q] & ANNP_NEIGHMASK;
        const double4 xj = a.xq[j];
        dx = xi.x - xj.x; dy = xi.y - xj.y; dz = xi.z - xj.z;
        r = sqrt(dx * dx + dy * dy + dz * dz);
        in = (r * CFLENGTH < Rc_max);
      }
      const unsigned mask = __ballot_sync(0xffffffffu, in);
      const int slot = N + __popc(mask & ((1u << lane) - 1u));
      if (in && slot < C) {
        BpGeom g;
        const double rinv = 1.0 / r;
        g.ux = dx * rinv; g.uy = dy * rinv; g.uz = dz * rinv; g.r = r;
        double sn, cs;
        sincospi(r * CFLENGTH * rcinv, &sn, &cs);
        g.fc = 0.5 * (cs + 1.0);
        g.dfc = -0.5 * kPi * rcinv * sn;
        sN[slot] = g;
        spos[slot] = q;
      } else if (valid) {
        if (!a.facc) a.fpair[p0 + q] = make_double4(0.0, 0.0, 0.0, 0.0);
        if (a.vpair) {
          double *vp = a.vpair + (size_t) (p0 + q) * 6;
#pragma unroll
          for (int k = 0; k < 6; k++) vp[k] = 0.0;
        }
      }
      N += __popc(mask);
    }
    if (lane == 0 && !a.work_list) {
      atomicMax(&a.cnt->max_neigh, N);
      atomicAdd(&a.cnt->sum_neigh, (unsigned long long) N);
      atomicAdd(&a.cnt->sum_trip, (unsigned long long) N * (unsigned long long) (N > 0 ? N - 1 : 0) / 2ull);
    }
    if (N > C) {
      if (lane == 0) { annp_note_overflow(a, ii); a.fself[ii] = make_double4(0.0, 0.0, 0.0, 0.0); }
      if (!a.facc) for (int q = lane; q < L; q += 32) a.fpair[p0 + q] = make_double4(0.0, 0.0, 0.0, 0.0);
      __syncwarp();
      continue;
    }
    __syncwarp();

    // ---- 2. forward.  radial: one lane per neighbour, one butterfly per component
    for (int m = 0; m < npsf; m++) {
      double acc = 0.0;
      for (int s = lane; s < N; s += 32) {
        const BpGeom g = sN[s];
        const double rm = g.r * CFLENGTH;
        if (rm < Rc_rad) {
          double fc, dfc;
          bp_fc(rm, Rc_rad, fc, dfc);
          acc += exp(-P.rad_eta[m] * rm * rm) * fc;                      // pair_annp.cpp:697-703
        }
      }
      acc = warp_sum(acc);
      if (lane == 0) sG[m] = (acc - P.sf_avg[m]) * P.sf_scale[m];
    }
    // angular: lane <-> first member a, partners b > a (the reference's kk > jj)
    double G[NT];
#pragma unroll
    for (int n = 0; n < NT; n++) G[n] = 0.0;
    for (int s = lane; s < N; s += 32) {
      const BpGeom A = sN[s];
      if (!(A.r * CFLENGTH < Rc)) continue;
      for (int b = s + 1; b < N; b++) {
        const BpGeom B = sN[b];
        if (!(B.r * CFLENGTH < Rc)) continue;
        double cosv, rjk, fcjk, dfcjk, E[NE], ex, ey, ez;
        if (!bp_pair_prims<NE>(A, B, Rc, rcinv, eta, cosv, rjk, fcjk, dfcjk, E, ex, ey, ez)) continue;
        const double tfc = A.fc * B.fc * fcjk;
        const double fm = 1.0 - cosv, fp = 1.0 + cosv;                    // lambda = -1, +1
        double pm = fm > 0.0 ? fm : 0.0, pp = fp > 0.0 ? fp : 0.0;       // flag <= 0: the term is skipped (:748-751)
        int lg = 0;
#pragma unroll
        for (int z = 0; z < NZ; z++) {
          while (lg < bp_zeta_log2<NZ>(z)) { pm *= pm; pp *= pp; lg++; }
          const double coef = ldexp(1.0, 1 - (1 << bp_zeta_log2<NZ>(z)));   // 2^(1 - zeta)
#pragma unroll
          for (int e = 0; e < NE; e++) {
            const double w = coef * E[e] * tfc;
            G[(e * NZ + z) * 2 + 0] = fma(w, pm, G[(e * NZ + z) * 2 + 0]);
            G[(e * NZ + z) * 2 + 1] = fma(w, pp, G[(e * NZ + z) * 2 + 1]);
          }
        }
      }
    }
#pragma unroll
    for (int n = 0; n < NT; n++) {
      const double v = warp_sum(G[n]);
      if (lane == 0) sG[npsf + n] = (v - P.sf_avg[npsf + n]) * P.sf_scale[npsf + n];   // (G - min)/(max - min), :168-170
    }
    __syncwarp();

    // ---- 3. MLP (raw output is the energy, :858-860)
    const int elem = P.map[ti];
    const double out = annp_mlp_warp(P, sW + elem * P.w_per_elem, sBias + elem * P.b_per_elem, sG, sdE, sH, sHd, sDel, lane);
    const double e_i = out;
    if (a.G_dbg) for (int n = lane; n < nsf; n += 32) { a.G_dbg[(size_t) ii * nsf + n] = sG[n]; a.dEdG_dbg[(size_t) ii * nsf + n] = sdE[n]; }
    for (int n = lane; n < nsf; n += 32) sCn[n] = sdE[n] * P.sf_scale[n];
    __syncwarp();
    double c[NT];
#pragma unroll
    for (int n = 0; n < NT; n++) c[n] = sCn[npsf + n];

    // ---- 4. backward: lane <-> neighbour s, all partners b
    double fix = 0, fiy = 0, fiz = 0;
    double v0 = 0, v1 = 0, v2 = 0, v3 = 0, v4 = 0, v5 = 0;
    for (int s = lane; s < N; s += 32) {
      const BpGeom A = sN[s];
      double gx = 0, gy = 0, gz = 0;                   // d out / d x_s in the reference's units (per Bohr)
      const double ra_m = A.r * CFLENGTH;
      if (ra_m < Rc_rad) {                             // annp_symmetry_pair, :686-708
        double fc, dfc;
        bp_fc(ra_m, Rc_rad, fc, dfc);
        double acc = 0.0;
        for (int m = 0; m < npsf; m++) {
          const double et = P.rad_eta[m];
          acc = fma(sCn[m], exp(-et * ra_m * ra_m) * (-fc * 2.0 * et * ra_m + dfc), acc);
        }
        gx = -acc * A.ux; gy = -acc * A.uy; gz = -acc * A.uz;      // dr/dx_j = -u
      }
      if (ra_m < Rc) {
        for (int b = 0; b < N; b++) {
          if (b == s) continue;
          const BpGeom B = sN[b];
          if (!(B.r * CFLENGTH < Rc)) continue;
          const bool first = s < b;                    // s is the reference's j (first) or k (second) of the pair
          double cosv, rjk, fcjk, dfcjk, E[NE], ex, ey, ez;
          // primitives are symmetric in (a, b) except e = x_a - x_b: evaluate them with (first, second) = (j, k)
          const bool ok = first ? bp_pair_prims<NE>(A, B, Rc, rcinv, eta, cosv, rjk, fcjk, dfcjk, E, ex, ey, ez)
                                : bp_pair_prims<NE>(B, A, Rc, rcinv, eta, cosv, rjk, fcjk, dfcjk, E, ex, ey, ez);
          if (!ok) continue;
          const double tfc = A.fc * B.fc * fcjk;
          const double fm = 1.0 - cosv, fp = 1.0 + cosv;
          const bool okm = fm > 0.0, okp = fp > 0.0;
          double pm = okm ? fm : 0.0, pp = okp ? fp : 0.0;
          const double im = okm ? 1.0 / fm : 0.0, ip = okp ? 1.0 / fp : 0.0;
          double S1 = 0.0, S2 = 0.0, S3 = 0.0;
          int lg = 0;
#pragma unroll
          for (int z = 0; z < NZ; z++) {
            while (lg < bp_zeta_log2<NZ>(z)) { pm *= pm; pp *= pp; lg++; }
            const double zeta = (double) (1 << bp_zeta_log2<NZ>(z));
            const double coef = ldexp(1.0, 1 - (1 << bp_zeta_log2<NZ>(z)));
            double tm = 0.0, tp = 0.0, um = 0.0, up = 0.0;   // sum_e c E (and eta-weighted) for lambda = -1 / +1
#pragma unroll
            for (int e = 0; e < NE; e++) {
              const double cm = c[(e * NZ + z) * 2 + 0] * E[e], cp = c[(e * NZ + z) * 2 + 1] * E[e];
              tm += cm; tp += cp;
              um = fma(eta[e], cm, um); up = fma(eta[e], cp, up);
            }
            const double Tm = coef * pm, Tp = coef * pp;
            S3 = fma(Tm, tm, fma(Tp, tp, S3));
            S2 = fma(Tm, um, fma(Tp, up, S2));
            S1 = fma(zeta * Tp * ip, tp, fma(-zeta * Tm * im, tm, S1));     // lambda zeta T / flag
          }
          const double k1 = S1 * tfc / CFLENGTH, k2 = S2 * tfc;
          // geometry of the member being differentiated (me) and of the other one (ot); e = x_j - x_k
          const double mr = A.r, orr = B.r;
          const double mx = A.ux, my = A.uy, mz = A.uz, ox = B.ux, oy = B.uy, oz = B.uz;
          const double rinv = 1.0 / mr, rjkinv = 1.0 / rjk;
          const double djx = ex * rjkinv, djy = ey * rjkinv, djz = ez * rjkinv;   // dr_djk = (x_j - x_k)/r_jk
          const double sgn = first ? 1.0 : -1.0;
          // d cos / d x_me = (-u_ot + cos u_me) / r_me
          const double cx = (cosv * mx - ox) * rinv, cy = (cosv * my - oy) * rinv, cz = (cosv * mz - oz) * rinv;
          // term2: 2 (r_me_m dr_dme +- r_IK_m dr_djk) -- r_ik in both, as the reference (:734-735)
          const double rik_m = (first ? orr : mr) * CFLENGTH, rme_m = mr * CFLENGTH;
          const double t2x = 2.0 * (rme_m * (-mx) + sgn * rik_m * djx);
          const double t2y = 2.0 * (rme_m * (-my) + sgn * rik_m * djy);
          const double t2z = 2.0 * (rme_m * (-mz) + sgn * rik_m * djz);
          // term3: fc_ot (dfc_me dr_dme fc_jk +- fc_me dfc_jk dr_djk)
          const double q1 = B.fc * A.dfc * fcjk, q2 = sgn * B.fc * A.fc * dfcjk;
          const double t3x = q1 * (-mx) + q2 * djx, t3y = q1 * (-my) + q2 * djy, t3z = q1 * (-mz) + q2 * djz;
          gx += k1 * cx - k2 * t2x + S3 * t3x;
          gy += k1 * cy - k2 * t2y + S3 * t3y;
          gz += k1 * cz - k2 * t2z + S3 * t3z;
        }
      }
      const double Fx = -gx, Fy = -gy, Fz = -gz;       // Fj of the reference before CFFORCE (:180-190)
      const int q = spos[s];
      if (a.facc) {     // fixed-point scatter (annp_device.cuh); else the per-entry buffer of the ordered gather
        if (!annp_fix_add(a, a.nbr[p0 + q] & ANNP_NEIGHMASK, Fx * CFFORCE, Fy * CFFORCE, Fz * CFFORCE)) atomicExch(&a.cnt->bad_force, 1);
      } else {
        a.fpair[p0 + q] = make_double4(Fx * CFFORCE, Fy * CFFORCE, Fz * CFFORCE, 0.0);
      }
      fix -= Fx * CFFORCE; fiy -= Fy * CFFORCE; fiz -= Fz * CFFORCE;
      if (a.vir_c || a.vpair) {
        const double X = A.r * A.ux, Y = A.r * A.uy, Z = A.r * A.uz;
        const double w0 = -X * Fx, w1 = -Y * Fy, w2 = -Z * Fz, w3 = -X * Fy, w4 = -X * Fz, w5 = -Y * Fz;
        v0 += w0; v1 += w1; v2 += w2; v3 += w3; v4 += w4; v5 += w5;
        if (a.vpair) {
          double *vp = a.vpair + (size_t) (p0 + q) * 6;
          vp[0] = w0; vp[1] = w1; vp[2] = w2; vp[3] = w3; vp[4] = w4; vp[5] = w5;
        }
      }
    }
    fix = warp_sum(fix); fiy = warp_sum(fiy); fiz = warp_sum(fiz);
    if (lane == 0) a.fself[ii] = make_double4(fix, fiy, fiz, e_i);
    if (a.vir_c) {
      v0 = warp_sum(v0); v1 = warp_sum(v1); v2 = warp_sum(v2);
      v3 = warp_sum(v3); v4 = warp_sum(v4); v5 = warp_sum(v5);
      if (lane == 0) {
        double *vc = a.vir_c + (size_t) ii * 6;
        vc[0] = v0; vc[1] = v1; vc[2] = v2; vc[3] = v3; vc[4] = v4; vc[5] = v5;
      }
    }
    __syncwarp();
  }
}


// ---------------------------------------------------------------------------------------------------------------
// v2 of the fast path, used when the neighbour tile holds <= 32 atoms (Ni: 18-22 inside 3.9 A): PAIR COMPACTION.
// Only ~40 % of the N(N-1)/2 neighbour pairs have r_jk inside the cutoff, and with "lane <-> neighbour" loops 18 of
// 32 lanes do 17 serial pair evaluations each.  Here the contributing pairs are found first (folded-triangle sweep,
// ballot compaction into a shared list), then forward and backward passes run with lane <-> compacted pair (2-3 full
// iterations instead of 17+9 ragged ones).  The backward pass writes the gradient of both members of a pair into a
// per-pair slot; each neighbour then adds its slots in ascending partner order -> fixed summation order, no atomics.
constexpr int kBpPairCap = 64;       // gradient slots per chunk (6 doubles each); more pairs are processed in chunks

#ifndef ANNP_BP_PAIR_MINBLOCKS
#define ANNP_BP_PAIR_MINBLOCKS 4
#endif
template <int NE, int NZ>
__global__ void __launch_bounds__(kWarps * 32, ANNP_BP_PAIR_MINBLOCKS) annp_bp_pair_kernel(const ForceArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int NT = NE * NZ * 2;
  constexpr int C = 32;                                   // tile size this kernel is launched for
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const DevParams &P = *a.prm;
  const int nsf = P.nsf, npsf = P.npsf, nnod = P.nnod, nl = P.nlayers;
  const int wtot = P.nelements * P.w_per_elem, btot = P.nelements * P.b_per_elem;

  double *sW = reinterpret_cast<double *>(smem_raw);
  double *sBias = sW + wtot;
  double *blk_end = sBias + btot;
  for (int t = threadIdx.x; t < wtot; t += blockDim.x) sW[t] = P.weights[t];
  for (int t = threadIdx.x; t < btot; t += blockDim.x) sBias[t] = P.bias[t];
  // per warp: geometry [C], gradient slots [cap][6], G/dE/c [3 nsf], MLP scratch, pair list [C(C-1)/2] u32, slot table [C][C] u16, spos [C]
  const size_t per_warp_doubles = (size_t) 6 * C + (size_t) 6 * kBpPairCap + 3 * nsf + (size_t) 2 * nl * nnod + 2 * nnod;
  size_t warp_bytes = per_warp_doubles * sizeof(double) + (size_t) (C * (C - 1) / 2) * sizeof(unsigned) + (size_t) C * C * sizeof(unsigned short) + (size_t) C * sizeof(int);
  warp_bytes = (warp_bytes + 15) & ~(size_t) 15;
  const size_t blk_bytes = ((size_t) ((unsigned char *) blk_end - smem_raw) + 15) & ~(size_t) 15;
  unsigned char *wbase = smem_raw + blk_bytes + (size_t) warp * warp_bytes;
  BpGeom *sN = reinterpret_cast<BpGeom *>(wbase);
  double *sT = reinterpret_cast<double *>(sN + C);                 // [kBpPairCap][6]
  double *sG = sT + 6 * kBpPairCap;
  double *sdE = sG + nsf;
  double *sCn = sdE + nsf;
  double *sH = sCn + nsf;
  double *sHd = sH + nl * nnod;
  double *sDel = sHd + nl * nnod;
  unsigned *sPair = reinterpret_cast<unsigned *>(sDel + 2 * nnod);  // a | b << 16
  unsigned short *sSlot = reinterpret_cast<unsigned short *>(sPair + C * (C - 1) / 2);   // [C][C] pair index or 0xffff
  int *spos = reinterpret_cast<int *>(sSlot + C * C);
  __syncthreads();

  const double Rc_rad = P.rad_rc, Rc = P.ang_rc, rcinv = 1.0 / Rc;
  const double Rc_max = fmax(Rc_rad, Rc);
  double eta[NE];
#pragma unroll
  for (int e = 0; e < NE; e++) eta[e] = P.ang_eta[e * NZ * 2];

  const unsigned long long nwork = annp_work_items(a);
  if (a.work_list && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&a.cnt->ovf_total, nwork);   // overflow pass: statistics
  for (;;) {
    unsigned long long item = 0;
    if (lane == 0) item = atomicAdd(a.work_ctr, 1ull);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= nwork) break;
    const int ii = annp_work_centre(a, item);
    const int i = a.ilist[ii];
    const double4 xi = a.xq[i];
    const int ti = (int) xi.w;
    const long long p0 = a.row_off[ii];
    const int L = (int) (a.row_off[ii + 1] - p0);

    // ---- 1. filter (list order kept).  Four row chunks per trip: the index loads, then the position gathers, are
    // issued back to back so their L2 latencies overlap (the kernel is latency-, not throughput-bound)
    int N = 0;
    for (int base = 0; base < L; base += 128) {
      int jv[4];
      double4 xv[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int q = base + 32 * u + lane;
        jv[u] = (q < L) ? (a.nbr[p0 + q] & ANNP_NEIGHMASK) : -1;
      }
#pragma unroll
      for (int u = 0; u < 4; u++) xv[u] = (jv[u] >= 0) ? a.xq[jv[u]] : make_double4(0.0, 0.0, 0.0, 0.0);
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int q = base + 32 * u + lane;
        if (base + 32 * u >= L) break;
        const bool valid = jv[u] >= 0;
        const double dx = xi.x - xv[u].x, dy = xi.y - xv[u].y, dz = xi.z - xv[u].z;
        const double r = sqrt(dx * dx + dy * dy + dz * dz);
        const bool in = valid && (r * CFLENGTH < Rc_max);
        const unsigned mask = __ballot_sync(0xffffffffu, in);
        const int slot = N + __popc(mask & ((1u << lane) - 1u));
        if (in && slot < C) {
          BpGeom g;
          const double rinv = 1.0 / r;
          g.ux = dx * rinv; g.uy = dy * rinv; g.uz = dz * rinv; g.r = r;
          double sn, cs;
          sincospi(r * CFLENGTH * rcinv, &sn, &cs);
          g.fc = 0.5 * (cs + 1.0);
          g.dfc = -0.5 * kPi * rcinv * sn;
          sN[slot] = g;
          spos[slot] = q;
        } else if (valid) {
          if (!a.facc) a.fpair[p0 + q] = make_double4(0.0, 0.0, 0.0, 0.0);
          if (a.vpair) {
            double *vp = a.vpair + (size_t) (p0 + q) * 6;
#pragma unroll
            for (int k = 0; k < 6; k++) vp[k] = 0.0;
          }
        }
        N += __popc(mask);
      }
    }
    if (lane == 0 && !a.work_list) {
      atomicMax(&a.cnt->max_neigh, N);
      atomicAdd(&a.cnt->sum_neigh, (unsigned long long) N);
    }
    if (N > C) {
      if (lane == 0) { annp_note_overflow(a, ii); a.fself[ii] = make_double4(0.0, 0.0, 0.0, 0.0); }
      if (!a.facc) for (int q = lane; q < L; q += 32) a.fpair[p0 + q] = make_double4(0.0, 0.0, 0.0, 0.0);
      __syncwarp();
      continue;
    }
    for (int t = lane; t < N * C / 2; t += 32) reinterpret_cast<unsigned *>(sSlot)[t] = 0xffffffffu;   // rows 0..N-1 of the slot table
    __syncwarp();

    // ---- 2. contributing pairs: folded triangle (rows it and N-2-it together hold N slots), ballot compaction
    int Pv = 0;
    for (int it = 0; 2 * it <= N - 2; it++) {
      const int a1 = it, a2 = N - 2 - it, len1 = N - 1 - a1;
      int pa = -1, pb = -1;
      if (lane < len1) { pa = a1; pb = a1 + 1 + lane; }
      else if (a2 != a1 && lane - len1 < it + 1) { pa = a2; pb = a2 + 1 + (lane - len1); }
      bool ok = false;
      if (pa >= 0) {
        const BpGeom A = sN[pa], B = sN[pb];
        if (A.r * CFLENGTH < Rc && B.r * CFLENGTH < Rc) {
          const double ex = B.r * B.ux - A.r * A.ux, ey = B.r * B.uy - A.r * A.uy, ez = B.r * B.uz - A.r * A.uz;
          ok = sqrt(ex * ex + ey * ey + ez * ez) * CFLENGTH < Rc;      // the very test bp_pair_prims applies
        }
      }
      const unsigned m = __ballot_sync(0xffffffffu, ok);
      if (ok) {
        const int t = Pv + __popc(m & ((1u << lane) - 1u));
        sPair[t] = (unsigned) pa | ((unsigned) pb << 16);
        sSlot[pa * C + pb] = (unsigned short) t;
        sSlot[pb * C + pa] = (unsigned short) t;
      }
      Pv += __popc(m);
    }
    // statistics: this kernel counts the CONTRIBUTING triplets T' (all three sides inside Rc), the quantity SURVEY 8d's
    // flop formula for the Ni copy is written in
    if (lane == 0 && !a.work_list) atomicAdd(&a.cnt->sum_trip, (unsigned long long) Pv);
    __syncwarp();

    // ---- 3. forward.  radial: one lane per neighbour (N <= 32 here), the cutoff function and the npsf exponentials are
    // computed ONCE and stay in registers for the backward pass; angular: one lane per contributing pair
    constexpr int kRadReg = 3;                                   // radial components of the shipped file; other counts take
    const bool rad_fast = npsf == kRadReg;                       // the component-by-component path
    static_assert(kRadReg + NT <= 32, "descriptor sums must fit one transposing reduction");
    double fcr = 0.0, dfcr = 0.0, rm_l = 0.0, er[kRadReg];
#pragma unroll
    for (int m = 0; m < kRadReg; m++) er[m] = 0.0;
    if (rad_fast) {
      if (lane < N) {
        const BpGeom g = sN[lane];
        rm_l = g.r * CFLENGTH;
        if (rm_l < Rc_rad) {
          if (Rc_rad == Rc) { fcr = g.fc; dfcr = g.dfc; }          // same cutoff as the angular part (the shipped file): cached by the filter
          else bp_fc(rm_l, Rc_rad, fcr, dfcr);
#pragma unroll
          for (int m = 0; m < kRadReg; m++) er[m] = exp(-P.rad_eta[m] * rm_l * rm_l);
        }
      }
    } else {
      for (int m = 0; m < npsf; m++) {
        double acc = 0.0;
        for (int s = lane; s < N; s += 32) {
          const BpGeom g = sN[s];
          const double rm = g.r * CFLENGTH;
          if (rm < Rc_rad) {
            double fc, dfc;
            bp_fc(rm, Rc_rad, fc, dfc);
            acc += exp(-P.rad_eta[m] * rm * rm) * fc;
          }
        }
        acc = warp_sum(acc);
        if (lane == 0) sG[m] = (acc - P.sf_avg[m]) * P.sf_scale[m];
      }
    }
    double G[NT];
#pragma unroll
    for (int n = 0; n < NT; n++) G[n] = 0.0;
    for (int t = lane; t < Pv; t += 32) {
      const unsigned pr = sPair[t];
      const BpGeom A = sN[pr & 0xffffu], B = sN[pr >> 16];
      double cosv, rjk, fcjk, dfcjk, E[NE], ex, ey, ez;
      bp_pair_prims<NE>(A, B, Rc, rcinv, eta, cosv, rjk, fcjk, dfcjk, E, ex, ey, ez);
      const double tfc = A.fc * B.fc * fcjk;
      const double fm = 1.0 - cosv, fp = 1.0 + cosv;
      double pm = fm > 0.0 ? fm : 0.0, pp = fp > 0.0 ? fp : 0.0;
      int lg = 0;
#pragma unroll
      for (int z = 0; z < NZ; z++) {
        while (lg < bp_zeta_log2<NZ>(z)) { pm *= pm; pp *= pp; lg++; }
        const double coef = ldexp(1.0, 1 - (1 << bp_zeta_log2<NZ>(z)));
#pragma unroll
        for (int e = 0; e < NE; e++) {
          const double w = coef * E[e] * tfc;
          G[(e * NZ + z) * 2 + 0] = fma(w, pm, G[(e * NZ + z) * 2 + 0]);
          G[(e * NZ + z) * 2 + 1] = fma(w, pp, G[(e * NZ + z) * 2 + 1]);
        }
      }
    }
    if (rad_fast) {
      // all npsf + NT <= 32 descriptor sums are reduced across the warp together (31 shuffle-adds instead of 5 per sum):
      // slot n < npsf radial, npsf + n angular; lane n ends up with the total of slot n
      double w[32];
#pragma unroll
      for (int n = 0; n < 32; n++) w[n] = 0.0;
#pragma unroll
      for (int m = 0; m < kRadReg; m++) w[m] = er[m] * fcr;
#pragma unroll
      for (int n = 0; n < NT; n++) w[kRadReg + n] = G[n];
      const double tot = warp_transpose_sum32(w, lane);
      if (lane < kRadReg + NT) sG[lane] = (tot - P.sf_avg[lane]) * P.sf_scale[lane];
    } else {
#pragma unroll
      for (int n = 0; n < NT; n++) {
        const double v = warp_sum(G[n]);
        if (lane == 0) sG[npsf + n] = (v - P.sf_avg[npsf + n]) * P.sf_scale[npsf + n];
      }
    }
    __syncwarp();

    // ---- 4. MLP
    const int elem = P.map[ti];
    const double out = annp_mlp_warp(P, sW + elem * P.w_per_elem, sBias + elem * P.b_per_elem, sG, sdE, sH, sHd, sDel, lane);
    const double e_i = out;
    if (a.G_dbg) for (int n = lane; n < nsf; n += 32) { a.G_dbg[(size_t) ii * nsf + n] = sG[n]; a.dEdG_dbg[(size_t) ii * nsf + n] = sdE[n]; }
    for (int n = lane; n < nsf; n += 32) sCn[n] = sdE[n] * P.sf_scale[n];
    __syncwarp();
    double c[NT];
#pragma unroll
    for (int n = 0; n < NT; n++) c[n] = sCn[npsf + n];

    // ---- 5. backward.  radial part per neighbour, then pair chunks: gradients of both members -> slots -> ordered sums
    double gx = 0, gy = 0, gz = 0;                      // d out / d x_s for s = lane (per Bohr)
    if (lane < N) {
      const BpGeom A = sN[lane];
      const double ra_m = A.r * CFLENGTH;
      if (ra_m < Rc_rad) {
        double acc = 0.0;
        if (rad_fast) {                                   // cutoff function and exponentials of stage 3
#pragma unroll
          for (int m = 0; m < kRadReg; m++) acc = fma(sCn[m], er[m] * (-fcr * 2.0 * P.rad_eta[m] * ra_m + dfcr), acc);
        } else {
          double fc, dfc;
          bp_fc(ra_m, Rc_rad, fc, dfc);
          for (int m = 0; m < npsf; m++) {
            const double et = P.rad_eta[m];
            acc = fma(sCn[m], exp(-et * ra_m * ra_m) * (-fc * 2.0 * et * ra_m + dfc), acc);
          }
        }
        gx = -acc * A.ux; gy = -acc * A.uy; gz = -acc * A.uz;
      }
    }
    for (int c0 = 0; c0 < Pv; c0 += kBpPairCap) {
      const int c1 = min(Pv, c0 + kBpPairCap);
      for (int t = c0 + lane; t < c1; t += 32) {
        const unsigned pr = sPair[t];
        const BpGeom A = sN[pr & 0xffffu], B = sN[pr >> 16];      // A = the reference's j (first in the row), B = k
        double cosv, rjk, fcjk, dfcjk, E[NE], ex, ey, ez;
        bp_pair_prims<NE>(A, B, Rc, rcinv, eta, cosv, rjk, fcjk, dfcjk, E, ex, ey, ez);
        const double tfc = A.fc * B.fc * fcjk;
        const double fm = 1.0 - cosv, fp = 1.0 + cosv;
        const bool okm = fm > 0.0, okp = fp > 0.0;
        double pm = okm ? fm : 0.0, pp = okp ? fp : 0.0;
        const double im = okm ? 1.0 / fm : 0.0, ip = okp ? 1.0 / fp : 0.0;
        double S1 = 0.0, S2 = 0.0, S3 = 0.0;
        int lg = 0;
#pragma unroll
        for (int z = 0; z < NZ; z++) {
          while (lg < bp_zeta_log2<NZ>(z)) { pm *= pm; pp *= pp; lg++; }
          const double zeta = (double) (1 << bp_zeta_log2<NZ>(z));
          const double coef = ldexp(1.0, 1 - (1 << bp_zeta_log2<NZ>(z)));
          double tm = 0.0, tp = 0.0, um = 0.0, up = 0.0;
#pragma unroll
          for (int e = 0; e < NE; e++) {
            const double cm = c[(e * NZ + z) * 2 + 0] * E[e], cp = c[(e * NZ + z) * 2 + 1] * E[e];
            tm += cm; tp += cp;
            um = fma(eta[e], cm, um); up = fma(eta[e], cp, up);
          }
          const double Tm = coef * pm, Tp = coef * pp;
          S3 = fma(Tm, tm, fma(Tp, tp, S3));
          S2 = fma(Tm, um, fma(Tp, up, S2));
          S1 = fma(zeta * Tp * ip, tp, fma(-zeta * Tm * im, tm, S1));
        }
        const double k1 = S1 * tfc / CFLENGTH, k2 = S2 * tfc;
        const double rjkinv = 1.0 / rjk;
        const double djx = ex * rjkinv, djy = ey * rjkinv, djz = ez * rjkinv;       // dr_djk = (x_j - x_k)/r_jk
        const double rik_m = B.r * CFLENGTH, rij_m = A.r * CFLENGTH;
        double *T = sT + (size_t) (t - c0) * 6;
        {   // first member j = A
          const double rinv = 1.0 / A.r;
          const double cx = (cosv * A.ux - B.ux) * rinv, cy = (cosv * A.uy - B.uy) * rinv, cz = (cosv * A.uz - B.uz) * rinv;
          const double q1 = B.fc * A.dfc * fcjk, q2 = B.fc * A.fc * dfcjk;
          T[0] = k1 * cx - k2 * 2.0 * (rij_m * (-A.ux) + rik_m * djx) + S3 * (q1 * (-A.ux) + q2 * djx);
          T[1] = k1 * cy - k2 * 2.0 * (rij_m * (-A.uy) + rik_m * djy) + S3 * (q1 * (-A.uy) + q2 * djy);
          T[2] = k1 * cz - k2 * 2.0 * (rij_m * (-A.uz) + rik_m * djz) + S3 * (q1 * (-A.uz) + q2 * djz);
        }
        {   // second member k = B: r_ik in both parts of term2, as the reference (:734-735)
          const double rinv = 1.0 / B.r;
          const double cx = (cosv * B.ux - A.ux) * rinv, cy = (cosv * B.uy - A.uy) * rinv, cz = (cosv * B.uz - A.uz) * rinv;
          const double q1 = A.fc * B.dfc * fcjk, q2 = -A.fc * B.fc * dfcjk;
          T[3] = k1 * cx - k2 * 2.0 * (rik_m * (-B.ux) - rik_m * djx) + S3 * (q1 * (-B.ux) + q2 * djx);
          T[4] = k1 * cy - k2 * 2.0 * (rik_m * (-B.uy) - rik_m * djy) + S3 * (q1 * (-B.uy) + q2 * djy);
          T[5] = k1 * cz - k2 * 2.0 * (rik_m * (-B.uz) - rik_m * djz) + S3 * (q1 * (-B.uz) + q2 * djz);
        }
      }
      __syncwarp();
      if (lane < N) {                                   // ordered sum over partners b = 0..N-1 of this chunk's slots
        const unsigned short *row = sSlot + lane * C;
        for (int b = 0; b < N; b++) {
          const int t = row[b];
          if (t >= c0 && t < c1) {
            const double *T = sT + (size_t) (t - c0) * 6 + (lane < b ? 0 : 3);
            gx += T[0]; gy += T[1]; gz += T[2];
          }
        }
      }
      __syncwarp();
    }

    // ---- 6. output
    double fix = 0, fiy = 0, fiz = 0;
    double v0 = 0, v1 = 0, v2 = 0, v3 = 0, v4 = 0, v5 = 0;
    if (lane < N) {
      const BpGeom A = sN[lane];
      const double Fx = -gx, Fy = -gy, Fz = -gz;        // Fj of the reference before CFFORCE (:180-190)
      const int q = spos[lane];
      if (a.facc) {     // fixed-point scatter (annp_device.cuh); else the per-entry buffer of the ordered gather
        if (!annp_fix_add(a, a.nbr[p0 + q] & ANNP_NEIGHMASK, Fx * CFFORCE, Fy * CFFORCE, Fz * CFFORCE)) atomicExch(&a.cnt->bad_force, 1);
      } else {
        a.fpair[p0 + q] = make_double4(Fx * CFFORCE, Fy * CFFORCE, Fz * CFFORCE, 0.0);
      }
      fix = -Fx * CFFORCE; fiy = -Fy * CFFORCE; fiz = -Fz * CFFORCE;
      if (a.vir_c || a.vpair) {
        const double X = A.r * A.ux, Y = A.r * A.uy, Z = A.r * A.uz;
        v0 = -X * Fx; v1 = -Y * Fy; v2 = -Z * Fz; v3 = -X * Fy; v4 = -X * Fz; v5 = -Y * Fz;
        if (a.vpair) {
          double *vp = a.vpair + (size_t) (p0 + q) * 6;
          vp[0] = v0; vp[1] = v1; vp[2] = v2; vp[3] = v3; vp[4] = v4; vp[5] = v5;
        }
      }
    }
    fix = warp_sum(fix); fiy = warp_sum(fiy); fiz = warp_sum(fiz);
    if (lane == 0) a.fself[ii] = make_double4(fix, fiy, fiz, e_i);
    if (a.vir_c) {
      v0 = warp_sum(v0); v1 = warp_sum(v1); v2 = warp_sum(v2);
      v3 = warp_sum(v3); v4 = warp_sum(v4); v5 = warp_sum(v5);
      if (lane == 0) {
        double *vc = a.vir_c + (size_t) ii * 6;
        vc[0] = v0; vc[1] = v1; vc[2] = v2; vc[3] = v3; vc[4] = v4; vc[5] = v5;
      }
    }
    __syncwarp();
  }
}

}    // namespace

// 1 when the angular table has the product structure the fast kernel is instantiated for, else 0 (generic kernel)
int annp_bp_layout(const DevParams &hp) {
  const int NE = 3, NZ = 4;
  static const double zetas[4] = {1.0, 2.0, 4.0, 16.0};
  if (hp.ntsf != NE * NZ * 2) return 0;
  for (int e = 0; e < NE; e++)
    for (int z = 0; z < NZ; z++)
      for (int l = 0; l < 2; l++) {
        const int n = (e * NZ + z) * 2 + l;
        if (hp.ang_eta[n] != hp.ang_eta[e * NZ * 2] || hp.ang_zeta[n] != zetas[z] || hp.ang_lambda[n] != (l ? 1.0 : -1.0)) return 0;
      }
  return 1;
}

size_t annp_bp_smem_bytes(const DevParams &hp, int capacity) {
  size_t blk = (size_t) (hp.nelements * (hp.w_per_elem + hp.b_per_elem)) * sizeof(double);
  blk = (blk + 15) & ~(size_t) 15;
  size_t per_warp = ((size_t) 6 * capacity + 3 * hp.nsf + (size_t) 2 * hp.nlayers * hp.nnod + 2 * hp.nnod) * sizeof(double) + (size_t) capacity * sizeof(int);
  per_warp = (per_warp + 15) & ~(size_t) 15;
  return blk + kWarps * per_warp;
}

static size_t annp_bp_pair_smem_bytes(const DevParams &hp) {
  const int C = 32;
  size_t blk = (size_t) (hp.nelements * (hp.w_per_elem + hp.b_per_elem)) * sizeof(double);
  blk = (blk + 15) & ~(size_t) 15;
  size_t per_warp = ((size_t) 6 * C + (size_t) 6 * kBpPairCap + 3 * hp.nsf + (size_t) 2 * hp.nlayers * hp.nnod + 2 * hp.nnod) * sizeof(double) +
                    (size_t) (C * (C - 1) / 2) * sizeof(unsigned) + (size_t) C * C * sizeof(unsigned short) + (size_t) C * sizeof(int);
  per_warp = (per_warp + 15) & ~(size_t) 15;
  return blk + kWarps * per_warp;
}

cudaError_t annp_bp_force_launch(const ForceArgs &args, const DevParams &hp, int num_sms, cudaStream_t stream, int max_blocks) {
  // product-structured table: pair-compaction kernel while the tile fits one warp, lane-per-neighbour kernel beyond
  const bool pairk = hp.bp_layout == 1 && args.capacity <= 32;
  const size_t smem = pairk ? annp_bp_pair_smem_bytes(hp) : annp_bp_smem_bytes(hp, args.capacity);
  void (*kern)(const ForceArgs) = pairk ? annp_bp_pair_kernel<3, 4> : (hp.bp_layout >= 1 ? annp_bp_fast_kernel<3, 4> : annp_bp_force_kernel);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
  if (e != cudaSuccess) return e;
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kWarps * 32, smem);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) return cudaErrorInvalidConfiguration;
  long long want = ((long long) args.inum + kWarps - 1) / kWarps;
  long long blocks = (long long) per_sm * num_sms;
  if (blocks > want) blocks = want;
  if (max_blocks > 0 && blocks > max_blocks) blocks = max_blocks;       // overflow pass: item count known on the device only
  if (blocks < 1) blocks = 1;
  kern<<<(unsigned) blocks, kWarps * 32, smem, stream>>>(args);
  return cudaGetLastError();
}

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
  const int nsf = P.nsf, npsf = P.npsf, ntsf = P.ntsf, nnod = P.nnod, nl = P.nlayers;
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

  for (;;) {
    unsigned long long item = 0;
    if (lane == 0) item = atomicAdd(&a.cnt->work, 1ull);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= (unsigned long long) a.inum) break;
    const int ii = (int) item;
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
        a.fpair[p0 + q] = make_double4(0.0, 0.0, 0.0, 0.0);
        if (a.vpair) {
          double *vp = a.vpair + (size_t) (p0 + q) * 6;
#pragma unroll
          for (int k = 0; k < 6; k++) vp[k] = 0.0;
        }
      }
      N += __popc(mask);
    }
    if (lane == 0) {
      atomicMax(&a.cnt->max_neigh, N);
      atomicAdd(&a.cnt->sum_neigh, (unsigned long long) N);
      atomicAdd(&a.cnt->sum_trip, (unsigned long long) N * (unsigned long long) (N > 0 ? N - 1 : 0) / 2ull);
    }
    if (N > C) {
      if (lane == 0) { atomicExch(&a.cnt->overflow, 1); a.fself[ii] = make_double4(0.0, 0.0, 0.0, 0.0); }
      for (int q = lane; q < L; q += 32) a.fpair[p0 + q] = make_double4(0.0, 0.0, 0.0, 0.0);
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
      a.fpair[p0 + q] = make_double4(Fx * CFFORCE, Fy * CFFORCE, Fz * CFFORCE, 0.0);
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

}    // namespace

size_t annp_bp_smem_bytes(const DevParams &hp, int capacity) {
  size_t blk = (size_t) (hp.nelements * (hp.w_per_elem + hp.b_per_elem)) * sizeof(double);
  blk = (blk + 15) & ~(size_t) 15;
  size_t per_warp = ((size_t) 6 * capacity + 3 * hp.nsf + (size_t) 2 * hp.nlayers * hp.nnod + 2 * hp.nnod) * sizeof(double) + (size_t) capacity * sizeof(int);
  per_warp = (per_warp + 15) & ~(size_t) 15;
  return blk + kWarps * per_warp;
}

cudaError_t annp_bp_force_launch(const ForceArgs &args, const DevParams &hp, int num_sms, cudaStream_t stream) {
  const size_t smem = annp_bp_smem_bytes(hp, args.capacity);
  cudaError_t e = cudaFuncSetAttribute(annp_bp_force_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
  if (e != cudaSuccess) return e;
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, annp_bp_force_kernel, kWarps * 32, smem);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) return cudaErrorInvalidConfiguration;
  long long want = ((long long) args.inum + kWarps - 1) / kWarps;
  long long blocks = (long long) per_sm * num_sms;
  if (blocks > want) blocks = want;
  if (blocks < 1) blocks = 1;
  annp_bp_force_kernel<<<(unsigned) blocks, kWarps * 32, smem, stream>>>(args);
  return cudaGetLastError();
}

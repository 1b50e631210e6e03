// annp_basis.cpp -- host arithmetic behind the two polynomial bases of the angular passes (no device code).
//
// The reference's angular descriptor is G_n = sum_{j<k} fc_j fc_k T_n(y), y = (cos(theta) + 1) / 2, with Chebyshev
// polynomials T_n on [0, 1] (fe_v2/src/pair_annp.cpp:658-695).  The kernel works in z = cos(theta):
//   * backward pass: A(y) = sum_n c_n T_n(y) is evaluated by Horner's rule in z, so the per-atom coefficients are mapped
//     to monomial coefficients a = cheb2mono . c,   T_n((z+1)/2) = sum_k cheb2mono[k][n] z^k;
//   * forward pass: the sums are accumulated in the block basis psi_{4b+i}(z) = T_{4b}(z) z^i (one accumulate per
//     order) and mapped back,                       T_n((z+1)/2) = sum_j blk2cheb[j][n] psi_j(z).
// All intermediate numbers are dyadic rationals that long double holds exactly for the supported orders; psi_j has degree
// exactly j, so its monomial matrix is upper triangular with power-of-two pivots and blk2cheb is a back substitution.
#include <algorithm>
#include <vector>

#include "../../include/annp_b200.h"

extern "C" int annp_b200_basis_matrices(int ntsf, double *cheb2mono, double *blk2cheb) {
  const int nt = ntsf;
  if (nt < 1 || nt > 24 || !cheb2mono || !blk2cheb) return ANNP_B200_EINVAL;
  const size_t nn = (size_t) nt * nt;
  // Chebyshev coefficients: T[n][k] = coefficient of x^k in T_n(x)
  std::vector<long double> T(nn, 0.0L), M(nn, 0.0L);
  for (int n = 0; n < nt; n++) {
    if (n == 0) T[0] = 1.0L;
    else if (n == 1) T[(size_t) nt + 1] = 1.0L;
    else
      for (int k = 0; k < nt; k++)
        T[(size_t) n * nt + k] = (k > 0 ? 2.0L * T[(size_t) (n - 1) * nt + k - 1] : 0.0L) - T[(size_t) (n - 2) * nt + k];
  }
  // M[q][n] = coefficient of z^q in T_n((z+1)/2): expand ((z+1)/2)^k power by power
  std::vector<long double> pw(nt, 0.0L), nx(nt, 0.0L);
  pw[0] = 1.0L;
  for (int k = 0; k < nt; k++) {
    for (int n = 0; n < nt; n++)
      for (int q = 0; q <= k; q++) M[(size_t) q * nt + n] += T[(size_t) n * nt + k] * pw[q];
    std::fill(nx.begin(), nx.end(), 0.0L);
    for (int q = 0; q <= k && q + 1 < nt; q++) { nx[q] += 0.5L * pw[q]; nx[q + 1] += 0.5L * pw[q]; }
    pw = nx;
  }
  for (size_t q = 0; q < nn; q++) cheb2mono[q] = (double) M[q];
  // B[k][j] = coefficient of z^k in psi_j(z) = T_{4b}(z) z^i, j = 4b + i (T above is also T_n in powers of z)
  std::vector<long double> B(nn, 0.0L), X(nn, 0.0L);
  for (int j = 0; j < nt; j++) {
    const int b4 = (j / 4) * 4, i = j % 4;
    for (int k = 0; k <= b4; k++) B[(size_t) (k + i) * nt + j] = T[(size_t) b4 * nt + k];
  }
  for (int n = 0; n < nt; n++)
    for (int j = nt - 1; j >= 0; j--) {
      long double acc = M[(size_t) j * nt + n];
      for (int q = j + 1; q < nt; q++) acc -= B[(size_t) j * nt + q] * X[(size_t) q * nt + n];
      X[(size_t) j * nt + n] = acc / B[(size_t) j * nt + j];
    }
  for (size_t q = 0; q < nn; q++) blk2cheb[q] = (double) X[q];
  return ANNP_B200_OK;
}

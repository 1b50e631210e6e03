/*
 * annp_oracle.c -- CPU restatement of the reference ANNP (Chebyshev descriptor) force evaluation.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the checker for the CUDA path; it is never linked into
 * libannp_b200.so and the product never calls it.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs load it.
 *
 * Parity status: PINNED.  The restatement is gated in tests/test_oracle.py against the unmodified
 * reference source compiled into oracle/_ref/ref_annp_fe (and against the committed golden vectors
 * under tests/golden/ generated from it by tests/golden/make_golden.py).
 *
 * Every function cites the reference lines it follows; paths are relative to
 * /root/reference/annp-gpu-lammps/fe_v2/src/.  The arithmetic keeps the reference's operation
 * order (compile with -ffp-contract=off) so results agree to the last bits; the only deliberate
 * difference is that the per-atom dG/dx table is indexed by neighbour SLOT (jnum entries, reused
 * across atoms) instead of by atom index over nall+2 freshly new[]'d rows (pair_annp.cpp:127-132),
 * which removes the O(nall) allocation per atom without changing any sum.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "oracle_params.h"

/* pair_annp.cpp:590-594 */
static void annp_fc(double rij, double Rc, double *fc, double *dfc) {
  double coeff_a = MY_PI / Rc * rij;
  *fc = 0.5 * (cos(coeff_a) + 1);
  *dfc = -0.5 * MY_PI / Rc * sin(coeff_a);
}

/* pair_annp.cpp:596-611 */
static void annp_Tx(double x, int n, double *Tx, double *dTx) {
  for (int i = 0; i < n; i++) {
    if (i == 0) { Tx[i] = 1; dTx[i] = 0; }
    else if (i == 1) { Tx[i] = x; dTx[i] = 1; }
    else {
      Tx[i] = 2 * x * Tx[i - 1] - Tx[i - 2];
      dTx[i] = 2 * Tx[i - 1] + 2 * x * dTx[i - 1] - dTx[i - 2];
    }
  }
}

/* pair_annp.cpp:613-616 with f_ijk = 1: pow(-1,1) * rij[i] / r   (argument named rsq is r) */
static void annp_dr_dij(double r, const double *xij, double *dr_dj) {
  for (int i = 0; i < 3; i++) dr_dj[i] = -1.0 * xij[i] / r;
}

/* pair_annp.cpp:618-628 (arguments named rsq are r) */
static void annp_dct_djk(double rij, double rik, const double *xij, const double *xik, double cos_theta,
                         double *dct_dj, double *dct_dk) {
  double B = rij * rik;
  double term1 = cos_theta / (rij * rij);
  double term2 = cos_theta / (rik * rik);
  for (int i = 0; i < 3; i++) {
    dct_dj[i] = (-1.0) * xik[i] / B + term1 * xij[i];
    dct_dk[i] = (-1.0) * xij[i] / B + term2 * xik[i];
  }
}

/* pair_annp.cpp:709-739 (Fe copy of the activation table) */
static void annp_actf(int flag_act, int nr, const double *wxb, double *h, double *hd) {
  const double coeff_a = 1.7159, coeff_b = 0.666666666666667, coeff_c = 0.1;
  for (int i = 0; i < nr; i++) {
    double t;
    switch (flag_act) {
      case 0: h[i] = wxb[i]; hd[i] = 1; break;
      case 1: h[i] = tanh(wxb[i]); hd[i] = 1 - h[i] * h[i]; break;
      case 2: h[i] = 1.0 / (1.0 + exp(wxb[i])); hd[i] = h[i] * (1 - h[i]); break;
      case 3: t = tanh(coeff_b * wxb[i]); h[i] = coeff_a * t; hd[i] = coeff_a * (1.0 - t * t) * coeff_b; break;
      default: t = tanh(coeff_b * wxb[i]); h[i] = coeff_a * t + coeff_c * wxb[i];
               hd[i] = coeff_a * (1.0 - t * t) * coeff_b + coeff_c; break;
    }
  }
}

/* pair_annp.cpp:741-804: forward pass + forward-mode Jacobian J <- diag(act') W J from I_nsf.
 * Same loop nests and summation order as dot_add_wxb (700-707) and dot_mat_2d (822-833). */
static double annp_feed_forward(const oracle_params_t *p, int itype, const double *G, double *dE_dG) {
  const int nsf = p->nsf, nnod = p->nnod, nl = p->ntl - 1;
  const int dim = nsf > nnod ? nsf : nnod;
  double J[ORACLE_MAX_SF][ORACLE_MAX_SF], J1[ORACLE_MAX_SF][ORACLE_MAX_SF];
  double dw[ORACLE_MAX_NOD][ORACLE_MAX_SF];
  double h[ORACLE_MAX_LAYERS][ORACLE_MAX_NOD], hd[ORACLE_MAX_NOD];
  memset(J, 0, sizeof J);
  memset(J1, 0, sizeof J1);
  memset(h, 0, sizeof h);
  (void) dim;
  for (int i = 0; i < nsf; i++) J[i][i] = 1.0;
  for (int l = 0; l < nl; l++) {
    const double *W = p->weights + ((size_t) itype * nl + l) * nnod * nsf;
    const double *b = p->bias + ((size_t) itype * nl + l) * nnod;
    int nr = nnod, nc = nnod;
    const double *in = G;
    if (l == 0) nc = nsf;
    else { if (l == nl - 1) nr = 1; in = h[l - 1]; }
    double wxb[ORACLE_MAX_NOD];
    for (int i = 0; i < nr; i++) {
      double a = 0.0;
      for (int j = 0; j < nc; j++) a += W[i * nsf + j] * in[j];
      a += b[i];
      wxb[i] = a;
    }
    annp_actf(p->flagact[l], nr, wxb, h[l], hd);
    /* hidly_dw = hidly_d (diagonal, other entries are exact zeros) x weight : each entry is the
     * reference's sum over k of hidly_d[i][k]*w[k][j] where only k = i is non-zero, plus 0.0 adds */
    for (int i = 0; i < nr; i++)
      for (int j = 0; j < nc; j++) dw[i][j] = hd[i] * W[i * nsf + j];
    for (int i = 0; i < nr; i++)
      for (int j = 0; j < nsf; j++) {
        double t = 0.0;
        for (int k = 0; k < nc; k++) t += dw[i][k] * J[k][j];
        J1[i][j] = t;
      }
    for (int i = 0; i < nr; i++)
      for (int j = 0; j < nsf; j++) J[i][j] = J1[i][j];
  }
  for (int i = 0; i < nsf; i++) dE_dG[i] = J[0][i];
  double out = h[nl - 1][0];
  return p->e_scale * out + p->e_shift + p->e_atom;   /* pair_annp.cpp:790-793 */
}

/* One centre atom: descriptor, network, per-slot forces.  pair_annp.cpp:110-214.
 * dG is [jnum][nsf][3] scratch (zeroed here); Fj_out is [jnum][3]. Returns E_i. */
static double annp_atom(const oracle_params_t *p, const double *sf_scale, const double *x, const int *type,
                        int i, const int *jlist, int jnum, double *dG, double *Fj_out, double *G_out) {
  const int nsf = p->nsf, npsf = p->npsf, ntsf = p->ntsf, nt1 = p->ntypes + 1;
  const int ritype = type[i], itype = p->map[ritype];
  const double xtmp = x[3 * i], ytmp = x[3 * i + 1], ztmp = x[3 * i + 2];
  double G[ORACLE_MAX_SF], dE_dG[ORACLE_MAX_SF];
  double Tx[ORACLE_MAX_SF], dTx[ORACLE_MAX_SF];
  memset(G, 0, sizeof G);
  memset(dG, 0, sizeof(double) * (size_t) jnum * nsf * 3);

  for (int jj = 0; jj < jnum; jj++) {
    int j = jlist[jj] & NEIGHMASK;
    int rjtype = type[j];
    double xij[3] = {xtmp - x[3 * j], ytmp - x[3 * j + 1], ztmp - x[3 * j + 2]};
    double rsqij = xij[0] * xij[0] + xij[1] * xij[1] + xij[2] * xij[2];
    double cutsq_ij = p->cutsq[ritype * nt1 + rjtype];
    if (rsqij > cutsq_ij || rsqij < 1.0e-12) continue;             /* :144 */
    double rijinv = 1.0 / sqrt(xij[0] * xij[0] + xij[1] * xij[1] + xij[2] * xij[2]);
    double rij_unit[3] = {rijinv * xij[0], rijinv * xij[1], rijinv * xij[2]};
    double rij = sqrt(rsqij);
    double Rc = sqrt(cutsq_ij);
    double fcij, dfcij, dr_dj[3];
    annp_fc(rij, Rc, &fcij, &dfcij);
    annp_dr_dij(rij, xij, dr_dj);
    {                                                               /* annp_symmetry_pair :633-656 */
      double Rcp = p->cut;
      double xx = 2 * rij / Rcp - 1;
      annp_Tx(xx, npsf, Tx, dTx);
      double *dGj = dG + (size_t) jj * nsf * 3;
      for (int m = 0; m < npsf; m++) {
        G[m] += sf_scale[m] * Tx[m] * fcij;
        double term1 = (dTx[m] * 2 / Rcp * fcij + Tx[m] * dfcij) * sf_scale[m];
        for (int n = 0; n < 3; n++) dGj[m * 3 + n] += term1 * dr_dj[n];
      }
    }
    for (int kk = jj + 1; kk < jnum; kk++) {
      int k = jlist[kk];                                            /* not masked in the reference, :157 */
      int rktype = type[k];
      double xik[3] = {xtmp - x[3 * k], ytmp - x[3 * k + 1], ztmp - x[3 * k + 2]};
      double rsqik = xik[0] * xik[0] + xik[1] * xik[1] + xik[2] * xik[2];
      double cutsq_ik = p->cutsq[ritype * nt1 + rktype];
      if (rsqik > cutsq_ik || rsqik < 1.0e-12) continue;           /* :165 */
      double rikinv = 1.0 / sqrt(xik[0] * xik[0] + xik[1] * xik[1] + xik[2] * xik[2]);
      double rik_unit[3] = {rikinv * xik[0], rikinv * xik[1], rikinv * xik[2]};
      double cos_theta = rij_unit[0] * rik_unit[0] + rij_unit[1] * rik_unit[1] + rij_unit[2] * rik_unit[2];
      double rik = sqrt(rsqik);
      double Rck = sqrt(cutsq_ik);
      double fcik, dfcik;
      annp_fc(rik, Rck, &fcik, &dfcik);
      {                                                             /* annp_symmetry_trip :658-695 */
        double dct_dj[3], dct_dk[3], dr_dk[3];
        double yy = 0.5 * (cos_theta + 1);
        annp_Tx(yy, ntsf, Tx, dTx);
        annp_dr_dij(rik, xik, dr_dk);
        annp_dct_djk(rij, rik, xij, xik, cos_theta, dct_dj, dct_dk);
        double *dGj = dG + (size_t) jj * nsf * 3;
        double *dGk = dG + (size_t) kk * nsf * 3;
        for (int n = 0; n < ntsf; n++) {
          double s = sf_scale[n + npsf];
          G[n + npsf] += s * Tx[n] * fcij * fcik;
          double term1 = dTx[n] * 0.5 * fcij * fcik;
          double term2 = Tx[n] * dfcij * fcik;
          double term3 = Tx[n] * fcij * dfcik;
          for (int m = 0; m < 3; m++) {
            double t_dG_dj = term1 * dct_dj[m] + term2 * dr_dj[m];
            double t_dG_dk = term1 * dct_dk[m] + term3 * dr_dk[m];
            dGj[(n + npsf) * 3 + m] += s * t_dG_dj;
            dGk[(n + npsf) * 3 + m] += s * t_dG_dk;
          }
        }
      }
    }
  }
  for (int k = 0; k < nsf; k++) G[k] = G[k] - sf_scale[k] * p->sfnor_avg[k];   /* :178-180 */
  if (G_out) memcpy(G_out, G, sizeof(double) * nsf);
  double e = annp_feed_forward(p, itype, G, dE_dG);
  for (int jj = 0; jj < jnum; jj++) {                                           /* :191-200 */
    const double *dGj = dG + (size_t) jj * nsf * 3;
    for (int k = 0; k < 3; k++) {
      double Fj = 0.0;
      for (int n = 0; n < nsf; n++) Fj += (-1.0) * dE_dG[n] * dGj[n * 3 + k] * p->e_scale;
      Fj_out[jj * 3 + k] = Fj;
    }
  }
  return e;
}

/*
 * Whole PairANNP::compute (pair_annp.cpp:74-222) for newton_pair = 1.
 *   f[nall][3] is ACCUMULATED into (f[j] += Fj, f[i] -= sum Fj), eatom[i] += E_i, eng += sum E_i
 *   virial6 : per-pair tally  sum_{i,j} (xi-xj) (x) (-Fj)   (ev_tally_xyz, :201-209); may be NULL
 *   vatom   : [nall][6], half to i and half to j;            may be NULL
 *   G_dump  : optional [inum][nsf] centred descriptors (test hook)
 * offsets[ii] is the start of row ii (ilist order) in neigh[].
 * Rows are processed in ilist order and all tallies happen in that order, as in the reference.
 * nthreads > 1 evaluates atoms concurrently but still tallies serially in ilist order, so the
 * result is independent of the thread count.
 */
int annp_oracle_compute(const oracle_params_t *p, int nlocal, int nghost, const double *x, const int *type,
                        int inum, const int *ilist, const int *numneigh, const int64_t *offsets,
                        const int *neigh, double *f, double *eng, double *eatom, double *virial6,
                        double *vatom, double *G_dump, int nthreads) {
  (void) nlocal; (void) nghost;
  if (p->nsf > ORACLE_MAX_SF || p->nnod > ORACLE_MAX_NOD || p->ntl - 1 > ORACLE_MAX_LAYERS) return -1;
  if (p->flagsym != 0) return -2;
  const int nsf = p->nsf;
  double sf_scale[ORACLE_MAX_SF];
  for (int i = 0; i < nsf; i++) {                                               /* :98-108 */
    double t_avg = p->sfnor_avg[i];
    double t_scale = sqrt(p->sfnor_cov[i] - t_avg * t_avg);
    sf_scale[i] = (t_scale <= 1.0e-10) ? 0.0 : 1.0 / t_scale;
  }
  int maxj = 0;
  for (int ii = 0; ii < inum; ii++) if (numneigh[ii] > maxj) maxj = numneigh[ii];
  if (nthreads < 1) nthreads = 1;
  const int chunk = 256 * nthreads;
  double *Fj_all = (double *) malloc(sizeof(double) * (size_t) chunk * (maxj + 1) * 3);
  double *E_all = (double *) malloc(sizeof(double) * chunk);
  double *dG_all = (double *) malloc(sizeof(double) * (size_t) nthreads * (maxj + 1) * nsf * 3);
  if (!Fj_all || !E_all || !dG_all) { free(Fj_all); free(E_all); free(dG_all); return -3; }
  double etot = 0.0;
  for (int base = 0; base < inum; base += chunk) {
    int cnt = inum - base < chunk ? inum - base : chunk;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 4) num_threads(nthreads)
#endif
    for (int c = 0; c < cnt; c++) {
      int tid = 0;
#ifdef _OPENMP
      tid = omp_get_thread_num();
#endif
      int ii = base + c;
      E_all[c] = annp_atom(p, sf_scale, x, type, ilist[ii], neigh + offsets[ii], numneigh[ii],
                           dG_all + (size_t) tid * (maxj + 1) * nsf * 3,
                           Fj_all + (size_t) c * (maxj + 1) * 3, G_dump ? G_dump + (size_t) ii * nsf : NULL);
    }
    for (int c = 0; c < cnt; c++) {                                             /* serial tally */
      int ii = base + c, i = ilist[ii];
      const int *jlist = neigh + offsets[ii];
      const double *Fj = Fj_all + (size_t) c * (maxj + 1) * 3;
      etot += E_all[c];
      if (eatom) eatom[i] += E_all[c];
      double Fi[3] = {0.0, 0.0, 0.0};
      for (int jj = 0; jj < numneigh[ii]; jj++) {
        int j = jlist[jj] & NEIGHMASK;
        for (int k = 0; k < 3; k++) { Fi[k] += Fj[jj * 3 + k]; f[3 * j + k] += Fj[jj * 3 + k]; }
        if (virial6 || vatom) {
          double del[3] = {x[3 * i] - x[3 * j], x[3 * i + 1] - x[3 * j + 1], x[3 * i + 2] - x[3 * j + 2]};
          double fx = -Fj[jj * 3], fy = -Fj[jj * 3 + 1], fz = -Fj[jj * 3 + 2];
          double v[6] = {del[0] * fx, del[1] * fy, del[2] * fz, del[0] * fy, del[0] * fz, del[1] * fz};
          if (virial6) for (int k = 0; k < 6; k++) virial6[k] += v[k];
          if (vatom) for (int k = 0; k < 6; k++) { vatom[6 * i + k] += 0.5 * v[k]; vatom[6 * j + k] += 0.5 * v[k]; }
        }
      }
      f[3 * i] -= Fi[0]; f[3 * i + 1] -= Fi[1]; f[3 * i + 2] -= Fi[2];
    }
  }
  if (eng) *eng += etot;
  free(Fj_all); free(E_all); free(dG_all);
  return 0;
}

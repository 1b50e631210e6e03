/*
 * annp_ni_oracle.c -- CPU restatement of the Ni copy of the reference ANNP pair style
 * (Behler-Parrinello radial G2 / narrow angular G3 in Bohr units, min-max normalisation, raw network output).
 *
 * TEST INFRASTRUCTURE ONLY (see annp_oracle.c): the checker for the CUDA path, never linked into
 * libannp_b200.so.  Parity status: PINNED against the unmodified reference source compiled into
 * oracle/_ref/ref_annp_ni (tests/test_oracle.py) and the golden vectors generated from it.
 *
 * Paths are relative to /root/reference/annp-gpu-lammps/ni/src/.  Operation order follows the reference
 * (-ffp-contract=off); as in annp_oracle.c the dG/dx table is indexed by neighbour slot instead of by atom index
 * (pair_annp.cpp:119,124), which gives the same sums whenever an atom appears once per list row (LAMMPS lists do).
 * The restatement is of the FIRST compute() call: the reference overwrites sf_max with sf_max - sf_min on every
 * call (pair_annp.cpp:99-101), so its later calls use a different normalisation.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "oracle_params.h"

#define CFLENGTH 1.889726      /* pair_annp.h:70 */
#define CFFORCE 51.422515      /* pair_annp.h:71 */

/* pair_annp.cpp:643-647 */
static void ni_fc(double rij, double Rc, double *fc, double *dfc) {
  double coeff_a = MY_PI / Rc * rij;
  *fc = 0.5 * (cos(coeff_a) + 1);
  *dfc = -0.5 * MY_PI / Rc * sin(coeff_a);
}

/* pair_annp.cpp:666-669: pow(-1, f_ijk) * rij[i] / r  (the argument named r2ij is r) */
static void ni_dr_dij(int f_ijk, double r, const double *xij, double *dr_dj) {
  for (int i = 0; i < 3; i++) dr_dj[i] = pow(-1, f_ijk) * xij[i] / r;
}

/* pair_annp.cpp:671-681 (arguments named r2 are r) */
static void ni_dct_djk(double rij, double rik, const double *xij, const double *xik, double cos_theta, double *dct_dj,
                       double *dct_dk) {
  double B = rij * rik;
  double term1 = cos_theta / (rij * rij);
  double term2 = cos_theta / (rik * rik);
  for (int i = 0; i < 3; i++) {
    dct_dj[i] = (-1.0) * xik[i] / B + term1 * xij[i];
    dct_dk[i] = (-1.0) * xij[i] / B + term2 * xik[i];
  }
}

/* pair_annp.cpp:786-807: activations 3 and 4 are plain tanh in this copy */
static void ni_actf(int flag_act, int nr, const double *wxb, double *h, double *hd) {
  for (int i = 0; i < nr; i++) {
    switch (flag_act) {
      case 0: h[i] = wxb[i]; hd[i] = 1; break;
      case 2: h[i] = 1.0 / (1.0 + exp(wxb[i])); hd[i] = h[i] * (1 - h[i]); break;
      default: h[i] = tanh(wxb[i]); hd[i] = 1.0 - h[i] * h[i]; break;      /* 1, 3, 4 */
    }
  }
}

/* pair_annp.cpp:809-871; same loop nests as the Fe copy, the energy is the raw output (858-860) */
static double ni_feed_forward(const oracle_params_t *p, int itype, const double *G, double *dE_dG) {
  const int nsf = p->nsf, nnod = p->nnod, nl = p->ntl - 1;
  static _Thread_local double J[ORACLE_MAX_SF][ORACLE_MAX_SF], J1[ORACLE_MAX_SF][ORACLE_MAX_SF], dw[ORACLE_MAX_NOD][ORACLE_MAX_SF];
  double h[ORACLE_MAX_LAYERS][ORACLE_MAX_NOD], hd[ORACLE_MAX_NOD];
  memset(J, 0, sizeof J);
  memset(J1, 0, sizeof J1);
  memset(h, 0, sizeof h);
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
    ni_actf(p->flagact[l], nr, wxb, h[l], hd);
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
  return h[nl - 1][0];
}

/* One centre atom (pair_annp.cpp:104-205).  dG [jnum][nsf][3] scratch; Fj_out [jnum][3] = the reference's Fj
 * BEFORE the CFFORCE factor.  Returns E_i. */
static double ni_atom(const oracle_params_t *p, const double *sf_range, const double *x, const int *type, int i,
                      const int *jlist, int jnum, double *dG, double *Fj_out, double *G_out) {
  const int nsf = p->nsf, npsf = p->npsf, ntsf = p->ntsf;
  const int itype = p->map[type[i]];
  const double xtmp = x[3 * i], ytmp = x[3 * i + 1], ztmp = x[3 * i + 2];
  double G[ORACLE_MAX_SF], dE_dG[ORACLE_MAX_SF];
  memset(G, 0, sizeof G);
  memset(dG, 0, sizeof(double) * (size_t) jnum * nsf * 3);

  for (int jj = 0; jj < jnum; jj++) {
    int j = jlist[jj] & NEIGHMASK;
    double xij[3] = {xtmp - x[3 * j], ytmp - x[3 * j + 1], ztmp - x[3 * j + 2]};
    double r2ij = xij[0] * xij[0] + xij[1] * xij[1] + xij[2] * xij[2];
    double rijinv = 1.0 / sqrt(xij[0] * xij[0] + xij[1] * xij[1] + xij[2] * xij[2]);
    double rij_unit[3] = {rijinv * xij[0], rijinv * xij[1], rijinv * xij[2]};
    double rij = sqrt(r2ij);
    double dr_dj[3];
    ni_dr_dij(1, rij, xij, dr_dj);
    double *dGj = dG + (size_t) jj * nsf * 3;
    {                                                                 /* annp_symmetry_pair :686-708 */
      double rij_m = rij * CFLENGTH;
      double Rc = p->sym_coerad[2];
      if (rij_m < Rc) {
        for (int m = 0; m < npsf; m++) {
          double fc, dfc;
          double eta = p->sym_coerad[m * 3 + 0];
          ni_fc(rij_m, Rc, &fc, &dfc);
          double term1 = exp(-eta * rij_m * rij_m);
          double term2 = term1 * (-fc * 2.0 * eta * rij_m + dfc);
          G[m] += term1 * fc;
          for (int n = 0; n < 3; n++) dGj[m * 3 + n] += term2 * dr_dj[n];
        }
      }
    }
    for (int kk = jj + 1; kk < jnum; kk++) {
      int k = jlist[kk];                                              /* not masked in the reference, :143 */
      double xik[3] = {xtmp - x[3 * k], ytmp - x[3 * k + 1], ztmp - x[3 * k + 2]};
      double xjk[3] = {x[3 * j] - x[3 * k], x[3 * j + 1] - x[3 * k + 1], x[3 * j + 2] - x[3 * k + 2]};
      double r2ik = xik[0] * xik[0] + xik[1] * xik[1] + xik[2] * xik[2];
      double r2jk = xjk[0] * xjk[0] + xjk[1] * xjk[1] + xjk[2] * xjk[2];
      double rikinv = 1.0 / sqrt(xik[0] * xik[0] + xik[1] * xik[1] + xik[2] * xik[2]);
      double rik_unit[3] = {rikinv * xik[0], rikinv * xik[1], rikinv * xik[2]};
      double cos_theta = rij_unit[0] * rik_unit[0] + rij_unit[1] * rik_unit[1] + rij_unit[2] * rik_unit[2];
      double rik = sqrt(r2ik);
      double rjk = sqrt(r2jk);
      {                                                               /* annp_symmetry_trip :710-767 */
        double fcij, fcik, fcjk, dfcij, dfcik, dfcjk;
        double dct_dj[3], dct_dk[3], dr_dk[3], dr_djk[3];
        ni_dr_dij(1, rik, xik, dr_dk);
        ni_dr_dij(0, rjk, xjk, dr_djk);
        ni_dct_djk(rij, rik, xij, xik, cos_theta, dct_dj, dct_dk);
        double rij_m = rij * CFLENGTH, rik_m = rik * CFLENGTH, rjk_m = rjk * CFLENGTH;
        double r2sum = rij_m * rij_m + rik_m * rik_m + rjk_m * rjk_m;
        double term2_drj[3], term2_drk[3], term3_drj[3], term3_drk[3];
        double Rc = p->sym_coeang[3];
        if (rij_m < Rc && rik_m < Rc && rjk_m < Rc) {
          ni_fc(rij_m, Rc, &fcij, &dfcij);
          ni_fc(rik_m, Rc, &fcik, &dfcik);
          ni_fc(rjk_m, Rc, &fcjk, &dfcjk);
          double term_fc = fcij * fcik * fcjk;
          for (int m = 0; m < 3; m++) {
            term2_drj[m] = 2.0 * (rij_m * dr_dj[m] + rik_m * dr_djk[m]);   /* r_ik where r_jk is meant: kept */
            term2_drk[m] = 2.0 * (rik_m * dr_dk[m] - rik_m * dr_djk[m]);
            term3_drj[m] = fcik * (dfcij * dr_dj[m] * fcjk + fcij * dfcjk * dr_djk[m]);
            term3_drk[m] = fcij * (dfcik * dr_dk[m] * fcjk - fcik * dfcjk * dr_djk[m]);
          }
          double *dGk = dG + (size_t) kk * nsf * 3;
          for (int n = 0; n < ntsf; n++) {
            double eta = p->sym_coeang[n * 4 + 0], lambda = p->sym_coeang[n * 4 + 1], zeta = p->sym_coeang[n * 4 + 2];
            double flag = (1 + lambda * cos_theta);
            if (flag <= 0) continue;
            double term_coe = pow(2, 1 - zeta);
            double term_cot = term_coe * pow(flag, zeta);
            double term_exp = exp(-eta * (r2sum));
            double tempG = term_cot * term_exp * term_fc;
            G[n + npsf] += tempG;
            double term1 = lambda * term_cot * term_exp * term_fc * zeta / flag / CFLENGTH;
            double term3 = term_cot * term_exp;
            double term2 = term3 * term_fc * eta;
            for (int m = 0; m < 3; m++) {
              dGj[(n + npsf) * 3 + m] += term1 * dct_dj[m] - term2 * term2_drj[m] + term3 * term3_drj[m];
              dGk[(n + npsf) * 3 + m] += term1 * dct_dk[m] - term2 * term2_drk[m] + term3 * term3_drk[m];
            }
          }
        }
      }
    }
  }
  for (int n = 0; n < nsf; n++) G[n] = (G[n] - p->sfnor_cov[n]) / sf_range[n];          /* :168-170 */
  if (G_out) memcpy(G_out, G, sizeof(double) * nsf);
  double e = ni_feed_forward(p, itype, G, dE_dG);
  for (int jj = 0; jj < jnum; jj++) {                                                    /* :180-190 */
    const double *dGj = dG + (size_t) jj * nsf * 3;
    for (int k = 0; k < 3; k++) {
      double Fj = 0.0;
      for (int n = 0; n < nsf; n++) Fj += (-1.0) * dE_dG[n] * dGj[n * 3 + k] / sf_range[n];
      Fj_out[jj * 3 + k] = Fj;
    }
  }
  return e;
}

/* Whole PairANNP::compute of the Ni copy (pair_annp.cpp:74-212), first call, newton_pair = 1.
 * Same conventions as annp_oracle_compute: f, eatom, eng, virial6, vatom are accumulated into. */
int annp_oracle_compute_ni(const oracle_params_t *p, int nlocal, int nghost, const double *x, const int *type, int inum,
                           const int *ilist, const int *numneigh, const int64_t *offsets, const int *neigh, double *f,
                           double *eng, double *eatom, double *virial6, double *vatom, double *G_dump, int nthreads) {
  (void) nlocal; (void) nghost;
  if (p->nsf > ORACLE_MAX_SF || p->nnod > ORACLE_MAX_NOD || p->ntl - 1 > ORACLE_MAX_LAYERS) return -1;
  if (!p->sym_coerad || !p->sym_coeang) return -2;
  const int nsf = p->nsf;
  double sf_range[ORACLE_MAX_SF];
  for (int i = 0; i < nsf; i++) sf_range[i] = p->sfnor_avg[i] - p->sfnor_cov[i];         /* sf_max - sf_min, :99-101 */
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
      E_all[c] = ni_atom(p, sf_range, x, type, ilist[ii], neigh + offsets[ii], numneigh[ii],
                         dG_all + (size_t) tid * (maxj + 1) * nsf * 3, Fj_all + (size_t) c * (maxj + 1) * 3,
                         G_dump ? G_dump + (size_t) ii * nsf : NULL);
    }
    for (int c = 0; c < cnt; c++) {                                                      /* serial tally, :172-204 */
      int ii = base + c, i = ilist[ii];
      const int *jlist = neigh + offsets[ii];
      const double *Fj = Fj_all + (size_t) c * (maxj + 1) * 3;
      etot += E_all[c];
      if (eatom) eatom[i] += E_all[c];
      double Fi[3] = {0.0, 0.0, 0.0};
      for (int jj = 0; jj < numneigh[ii]; jj++) {
        int j = jlist[jj] & NEIGHMASK;
        for (int k = 0; k < 3; k++) { Fi[k] += Fj[jj * 3 + k] * CFFORCE; f[3 * j + k] += Fj[jj * 3 + k] * CFFORCE; }
        if (virial6 || vatom) {                                                          /* tally WITHOUT CFFORCE, :191-198 */
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

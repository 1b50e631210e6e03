/*
 * anna_adp_oracle.c -- CPU restatement of the reference ANNA-ADP pair style (physically informed NN potential:
 * a Chebyshev descriptor feeds a small network whose two outputs d2, q2 parametrise an angular-dependent potential).
 *
 * TEST INFRASTRUCTURE ONLY (see annp_oracle.c): the checker for the CUDA path, never linked into
 * libannp_b200.so.  Parity status: PINNED against the unmodified reference source compiled into
 * oracle/_ref/ref_anna_adp (tests/test_oracle.py) and the golden vectors generated from it.
 *
 * Paths are relative to /root/reference/anna-gpu-lammps/bcc_fe/src/.  Operation order follows the reference
 * (-ffp-contract=off).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "oracle_params.h"

/* pair_anna_adp.cpp:639-648 */
static void anna_Tx(double x, int n, double *Tx) {
  for (int i = 0; i < n; i++) {
    if (i == 0) Tx[i] = 1;
    else if (i == 1) Tx[i] = x;
    else Tx[i] = 2 * x * Tx[i - 1] - Tx[i - 2];
  }
}

/* pair_anna_adp.cpp:696-720: activations 3 and 4 are 1.7 tanh(0.3 x) in this copy */
static void anna_actf(int flag_act, int nr, const double *wxb, double *h) {
  const double coeff_a = 1.7, coeff_b = 0.3;
  for (int i = 0; i < nr; i++) {
    if (flag_act == 0) h[i] = wxb[i];
    if (flag_act == 1) h[i] = tanh(wxb[i]);
    if (flag_act == 2) h[i] = 1.0 / (1.0 + exp(wxb[i]));
    if (flag_act == 3 || flag_act == 4) h[i] = coeff_a * tanh(coeff_b * wxb[i]);
  }
}

/* pair_anna_adp.cpp:722-751: forward pass only; the last layer has nout rows */
static void anna_feed_forward(const oracle_params_t *p, int itype, const double *G, double *lparams) {
  const int nsf = p->nsf, nnod = p->nnod, nl = p->ntl - 1;
  double h[ORACLE_MAX_LAYERS][ORACLE_MAX_NOD];
  memset(h, 0, sizeof h);
  for (int l = 0; l < nl; l++) {
    const double *W = p->weights + ((size_t) itype * nl + l) * nnod * nsf;
    const double *b = p->bias + ((size_t) itype * nl + l) * nnod;
    int nr = nnod, nc = nnod;
    const double *in = G;
    if (l == 0) nc = nsf;
    else { if (l == nl - 1) nr = p->nout; in = h[l - 1]; }
    double wxb[ORACLE_MAX_NOD];
    for (int i = 0; i < nr; i++) {
      double a = 0.0;
      for (int j = 0; j < nc; j++) a += W[i * nsf + j] * in[j];
      a += b[i];
      wxb[i] = a;
    }
    anna_actf(p->flagact[l], nr, wxb, h[l]);
  }
  for (int i = 0; i < p->nout; i++) lparams[i] = h[nl - 1][i];
}

/* One centre atom (pair_anna_adp.cpp:104-272).  all_xij [jnum][4] scratch; Fj_out [jnum][3] = (fx, fy, fz) of the
 * reference's force loop (f[j] += , f[i] -= ), zero for entries outside Rc.  Returns E_i. */
static double anna_atom(const oracle_params_t *p, const double *x, const int *type, int i, const int *jlist, int jnum,
                        double *all_xij, double *Fj_out, double *G_out, double *lp_out) {
  const int nsf = p->nsf, npsf = p->npsf, ntsf = p->ntsf, nt1 = p->ntypes + 1;
  const int ritype = type[i], itype = p->map[ritype];
  const double Rc = p->cut, coeff_b = MY_PI / Rc;
  const double *gp = p->gparams;
  const double A0 = gp[0], yy = gp[1], gamma = gp[2], C0 = gp[3], c1F = gp[4], c2F = gp[5], V0 = gp[6], b1 = gp[7];
  const double b2 = gp[8], delta = gp[9], r0 = gp[10], r1 = gp[11], hc = gp[12], d1 = gp[13], q1 = gp[14], d3 = gp[15], q3 = gp[16];
  const double E_base = p->e_base;
  const double xtmp = x[3 * i], ytmp = x[3 * i + 1], ztmp = x[3 * i + 2];
  double G[ORACLE_MAX_SF], Tx[ORACLE_MAX_SF], lparams[ORACLE_MAX_NOD];
  memset(G, 0, sizeof G);
  memset(lparams, 0, sizeof lparams);
  memset(Fj_out, 0, sizeof(double) * 3 * (size_t) jnum);

  for (int jj = 0; jj < jnum; jj++) {
    int j = jlist[jj] & NEIGHMASK;
    double xij[3] = {xtmp - x[3 * j], ytmp - x[3 * j + 1], ztmp - x[3 * j + 2]};
    double rsqij = xij[0] * xij[0] + xij[1] * xij[1] + xij[2] * xij[2];
    all_xij[jj * 4 + 0] = xij[0]; all_xij[jj * 4 + 1] = xij[1]; all_xij[jj * 4 + 2] = xij[2];
    all_xij[jj * 4 + 3] = sqrt(rsqij);
    if (rsqij > p->cutsq[ritype * nt1 + type[j]] || rsqij < 1.0e-12) continue;          /* :134 */
    double rijinv = 1.0 / sqrt(xij[0] * xij[0] + xij[1] * xij[1] + xij[2] * xij[2]);
    double rij_unit[3] = {rijinv * xij[0], rijinv * xij[1], rijinv * xij[2]};
    double rij = all_xij[jj * 4 + 3];
    double fcij = 0.5 * (cos(coeff_b * rij) + 1.0);
    {                                                                                    /* symmetry_pair :653-666 */
      double xx = 2 * rij / Rc - 1;
      anna_Tx(xx, npsf, Tx);
      for (int m = 0; m < npsf; m++) G[m] += Tx[m] * fcij;
    }
    for (int kk = jj + 1; kk < jnum; kk++) {
      int k = jlist[kk];                                                                 /* not masked, :141 */
      double xik[3] = {xtmp - x[3 * k], ytmp - x[3 * k + 1], ztmp - x[3 * k + 2]};
      double rsqik = xik[0] * xik[0] + xik[1] * xik[1] + xik[2] * xik[2];
      if (rsqik > p->cutsq[ritype * nt1 + type[k]] || rsqik < 1.0e-12) continue;        /* :149 */
      double rikinv = 1.0 / sqrt(xik[0] * xik[0] + xik[1] * xik[1] + xik[2] * xik[2]);
      double rik_unit[3] = {rikinv * xik[0], rikinv * xik[1], rikinv * xik[2]};
      double cos_theta = rij_unit[0] * rik_unit[0] + rij_unit[1] * rik_unit[1] + rij_unit[2] * rik_unit[2];
      double rik = sqrt(rsqik);
      double fcik = 0.5 * (cos(coeff_b * rik) + 1.0);
      double xa = 0.5 * (cos_theta + 1);                                                 /* symmetry_trip :668-681 */
      anna_Tx(xa, ntsf, Tx);
      for (int n = 0; n < ntsf; n++) G[n + npsf] += Tx[n] * fcij * fcik;
    }
  }
  if (G_out) memcpy(G_out, G, sizeof(double) * nsf);
  anna_feed_forward(p, itype, G, lparams);
  double d2 = lparams[0], q2 = lparams[1];
  if (lp_out) { lp_out[0] = d2; lp_out[1] = q2; }

  double mu_i[3] = {0, 0, 0}, lambda_i[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  double rho_i = 0.0;
  double coeff_repul = V0 / (b2 - b1);
  double adp_repul_eng = 0.0;
  for (int jj = 0; jj < jnum; jj++) {                                                    /* :177-199 */
    const double *a = all_xij + jj * 4;
    if (a[3] > Rc || a[3] < 1.0e-12) continue;
    double stpf_x = (a[3] - Rc) / hc;
    double adp_stpf = pow(stpf_x, 4) / (1 + pow(stpf_x, 4));
    double adp_u = adp_stpf * (d1 * exp(-d2 * a[3]) + d3);
    double adp_w = adp_stpf * (q1 * exp(-q2 * a[3]) + q3);
    mu_i[0] += adp_u * a[0];
    mu_i[1] += adp_u * a[1];
    mu_i[2] += adp_u * a[2];
    for (int row = 0; row < 3; row++)
      for (int col = 0; col < 3; col++) lambda_i[row][col] += adp_w * a[row] * a[col];
    double rho_z = a[3] - r0;
    double exp_z = exp(-gamma * rho_z);
    rho_i += adp_stpf * (A0 * pow(rho_z, yy) * exp_z * (1 + exp_z) + C0);
    double repul_z = a[3] / r1;
    adp_repul_eng += adp_stpf * (coeff_repul * (b2 / pow(repul_z, b1) - b1 / pow(repul_z, b2)) + delta);
  }
  double v_i = lambda_i[0][0] + lambda_i[1][1] + lambda_i[2][2];
  double sum_mu_i = 0.0, sum_lambda_i = 0.0;
  for (int row = 0; row < 3; row++) {
    sum_mu_i += mu_i[row] * mu_i[row];
    for (int col = 0; col < 3; col++) sum_lambda_i += pow(lambda_i[row][col], 2);
  }
  double f_v = -1.0 / 3.0 * v_i;
  double rep_coeff = V0 / (b2 - b1);
  double adp_angular_eng = 0.5 * sum_mu_i + 0.5 * sum_lambda_i - 1.0 / 6.0 * v_i * v_i;
  double adp_embed_eng = c1F * sqrt(rho_i) + c2F * pow(rho_i, 2);
  double evdwl = 0.5 * (adp_repul_eng) + adp_embed_eng + adp_angular_eng + E_base;      /* :213 */

  for (int jj = 0; jj < jnum; jj++) {                                                    /* :216-272 */
    const double *a = all_xij + jj * 4;
    if (a[3] > Rc || a[3] < 1.0e-12) continue;
    double xi = a[0], yi = a[1], zi = a[2], rij = a[3];
    double stpf_x = (rij - Rc) / hc;
    double adp_stpf_t1 = 1 + pow(stpf_x, 4);
    double adp_stpf = pow(stpf_x, 4) / adp_stpf_t1;
    double d_adp_stpf = 4 * pow(stpf_x, 3) / pow(adp_stpf_t1, 2) / hc;
    double rho_z = rij - r0;
    double exp_z = exp(-gamma * rho_z);
    double z_yy = A0 * pow(rho_z, yy);
    double ga_zyy = z_yy * gamma;
    double d_adp_rho = exp_z * (1.0 + exp_z) * (z_yy * (d_adp_stpf + adp_stpf * yy / rho_z) - ga_zyy) + C0 * d_adp_stpf - ga_zyy * exp_z * exp_z;
    double d_embed_eng = (0.5 * c1F * pow(rho_i, -0.5) + 2.0 * c2F * rho_i) * d_adp_rho;
    double repul_z = rij / r1;
    double zb1 = pow(repul_z, b1);
    double zb2 = pow(repul_z, b2);
    double drep_t = b2 * b1 / r1;
    double rep_t1 = rep_coeff * (b2 / zb1 - b1 / zb2) + delta;
    double d_repul_eng = d_adp_stpf * rep_t1 + adp_stpf * rep_coeff * (drep_t / repul_z * (-1.0 / zb1 + 1.0 / zb2));
    double adp_u_term = d1 * exp(-d2 * rij);
    double adp_w_term = q1 * exp(-q2 * rij);
    double adp_u = adp_stpf * (adp_u_term + d3);
    double adp_w = 2.0 * adp_stpf * (adp_w_term + q3);
    double d_adp_u = d_adp_stpf * (adp_u_term + d3) + adp_stpf * (-d2 * adp_u_term);
    double d_adp_w = d_adp_stpf * (adp_w_term + q3) + adp_stpf * (-q2 * adp_w_term);
    double d_angular_lamb1 = d_adp_w * (lambda_i[0][0] * xi * xi + lambda_i[1][1] * yi * yi + lambda_i[2][2] * zi * zi);
    double d_angular_lamb2 = d_adp_w * (lambda_i[0][1] * xi * yi + lambda_i[0][2] * xi * zi + lambda_i[1][2] * yi * zi) * 2.0 + d_angular_lamb1;
    double df_term1 = 0.5 * d_repul_eng + d_embed_eng + d_adp_u * (mu_i[0] * xi + mu_i[1] * yi + mu_i[2] * zi) + d_angular_lamb2;
    double df_term3 = f_v * (d_adp_w * rij + adp_w);
    double fx = df_term1 * xi / rij + adp_w * (yi * lambda_i[0][1] + zi * lambda_i[0][2] + xi * lambda_i[0][0]) + mu_i[0] * adp_u + xi * df_term3;
    double fy = df_term1 * yi / rij + adp_w * (yi * lambda_i[1][1] + zi * lambda_i[1][2] + xi * lambda_i[0][1]) + mu_i[1] * adp_u + yi * df_term3;
    double fz = df_term1 * zi / rij + adp_w * (yi * lambda_i[1][2] + zi * lambda_i[2][2] + xi * lambda_i[0][2]) + mu_i[2] * adp_u + zi * df_term3;
    Fj_out[jj * 3 + 0] = fx; Fj_out[jj * 3 + 1] = fy; Fj_out[jj * 3 + 2] = fz;
  }
  return evdwl;
}

/* Whole PairANNA_ADP::compute (pair_anna_adp.cpp:73-286), newton_pair = 1.  Conventions as annp_oracle_compute;
 * G_dump [inum][nsf] raw descriptors, lp_dump [inum][2] = (d2, q2) are optional test hooks.
 * The reference interleaves f[i] -= and f[j] += per pair; the tally below keeps that order. */
int anna_oracle_compute(const oracle_params_t *p, int nlocal, int nghost, const double *x, const int *type, int inum,
                        const int *ilist, const int *numneigh, const int64_t *offsets, const int *neigh, double *f,
                        double *eng, double *eatom, double *virial6, double *vatom, double *G_dump, double *lp_dump,
                        int nthreads) {
  (void) nlocal; (void) nghost;
  if (p->nsf > ORACLE_MAX_SF || p->nnod > ORACLE_MAX_NOD || p->ntl - 1 > ORACLE_MAX_LAYERS) return -1;
  if (!p->gparams || p->ngp < 17 || p->nout < 2) return -2;
  int maxj = 0;
  for (int ii = 0; ii < inum; ii++) if (numneigh[ii] > maxj) maxj = numneigh[ii];
  if (nthreads < 1) nthreads = 1;
  const int chunk = 256 * nthreads;
  double *Fj_all = (double *) malloc(sizeof(double) * (size_t) chunk * (maxj + 1) * 3);
  double *E_all = (double *) malloc(sizeof(double) * chunk);
  double *xij_all = (double *) malloc(sizeof(double) * (size_t) nthreads * (maxj + 1) * 4);
  if (!Fj_all || !E_all || !xij_all) { free(Fj_all); free(E_all); free(xij_all); return -3; }
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
      E_all[c] = anna_atom(p, x, type, ilist[ii], neigh + offsets[ii], numneigh[ii], xij_all + (size_t) tid * (maxj + 1) * 4,
                           Fj_all + (size_t) c * (maxj + 1) * 3, G_dump ? G_dump + (size_t) ii * p->nsf : NULL,
                           lp_dump ? lp_dump + (size_t) ii * 2 : NULL);
    }
    for (int c = 0; c < cnt; c++) {
      int ii = base + c, i = ilist[ii];
      const int *jlist = neigh + offsets[ii];
      const double *Fj = Fj_all + (size_t) c * (maxj + 1) * 3;
      for (int jj = 0; jj < numneigh[ii]; jj++) {
        int j = jlist[jj] & NEIGHMASK;
        double fx = Fj[jj * 3], fy = Fj[jj * 3 + 1], fz = Fj[jj * 3 + 2];
        if (fx == 0.0 && fy == 0.0 && fz == 0.0) continue;      /* outside Rc (adding +-0 changes nothing) */
        f[3 * i] -= fx; f[3 * i + 1] -= fy; f[3 * i + 2] -= fz;
        f[3 * j] += fx; f[3 * j + 1] += fy; f[3 * j + 2] += fz;
        if (virial6 || vatom) {                                 /* ev_tally_xyz(i,j,..,-fx,-fy,-fz, xij) :264-266 */
          double del[3] = {x[3 * i] - x[3 * j], x[3 * i + 1] - x[3 * j + 1], x[3 * i + 2] - x[3 * j + 2]};
          double v[6] = {del[0] * -fx, del[1] * -fy, del[2] * -fz, del[0] * -fy, del[0] * -fz, del[1] * -fz};
          if (virial6) for (int k = 0; k < 6; k++) virial6[k] += v[k];
          if (vatom) for (int k = 0; k < 6; k++) { vatom[6 * i + k] += 0.5 * v[k]; vatom[6 * j + k] += 0.5 * v[k]; }
        }
      }
      etot += E_all[c];
      if (eatom) eatom[i] += E_all[c];
    }
  }
  if (eng) *eng += etot;
  free(Fj_all); free(E_all); free(xij_all);
  return 0;
}

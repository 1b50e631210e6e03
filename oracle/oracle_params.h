/* oracle_params.h -- parameter block shared by the CPU restatements under oracle/ (TEST INFRASTRUCTURE ONLY;
 * mirrored field by field in oracle/restatement.py). */
#ifndef ORACLE_PARAMS_H
#define ORACLE_PARAMS_H
#define ORACLE_MAX_SF 64
#define ORACLE_MAX_NOD 64
#define ORACLE_MAX_LAYERS 8
#define NEIGHMASK 0x1FFFFFFF
#define MY_PI 3.14159265358979323846

typedef struct {
  int ntypes;               /* LAMMPS atom types, 1-based                              */
  int nelements;            /* elements in the potential file                          */
  int ntl, nhl, nnod;       /* total layers (incl. input), hidden layers, nodes/layer  */
  int nsf, npsf, ntsf;      /* descriptor sizes: total, radial, angular                */
  int flagsym;              /* 0 = Chebyshev                                           */
  int flagact[ORACLE_MAX_LAYERS];
  double cut, e_scale, e_shift, e_atom;
  const double *sfnor_cov;  /* [nsf]                                                   */
  const double *sfnor_avg;  /* [nsf]                                                   */
  const int *map;           /* [ntypes+1] type -> element                              */
  const double *cutsq;      /* [(ntypes+1)*(ntypes+1)]                                 */
  const double *weights;    /* [nelements][ntl-1][nnod][nsf]  (weight_all, padded)     */
  const double *bias;       /* [nelements][ntl-1][nnod]       (bias_all[..][0][..])    */
  /* Ni copy (annp_oracle_compute_ni): sfnor_cov = sf_min row, sfnor_avg = sf_max row of the file  */
  const double *sym_coerad; /* [npsf][3] eta, rs, Rc      (ni/src/pair_annp.cpp:510-545)        */
  const double *sym_coeang; /* [ntsf][4] eta, lambda, zeta, Rc                                  */
  /* ANNA-ADP copy (anna_oracle_compute): nout network outputs, ngp global ADP parameters        */
  int nout, ngp;
  const double *gparams;    /* [ngp]                                                            */
  double e_base;
} oracle_params_t;
#endif

// Shim of LAMMPS' Pair base class for the oracle harness (TEST INFRASTRUCTURE, not product).
// Implements the members and the tally semantics the reference pair styles rely on:
//   ev_init / ev_setup flag decoding, ev_tally_xyz (newton-aware pair virial, half/half per-atom
//   split) and virial_fdotr_compute (sum over local+ghost of x (x) f), following the documented
//   behaviour of LAMMPS stable_2Aug2023 pair.cpp.  Written from the documented semantics.
#ifndef SHIM_PAIR_H
#define SHIM_PAIR_H
#include "pointers.h"
#include "atom.h"
#include "force.h"
#include "memory.h"
#include "neigh_list.h"
#include <vector>

#define PairStyle(key, Class)

namespace LAMMPS_NS {

enum { ENERGY_NONE = 0, ENERGY_GLOBAL = 1, ENERGY_ATOM = 2 };
enum { VIRIAL_NONE = 0, VIRIAL_PAIR = 1, VIRIAL_FDOTR = 2, VIRIAL_ATOM = 4, VIRIAL_CENTROID = 8 };

class Pair : protected Pointers {
 public:
  double eng_vdwl = 0.0, eng_coul = 0.0;
  double virial[6] = {0, 0, 0, 0, 0, 0};
  double *eatom = nullptr, **vatom = nullptr, **cvatom = nullptr;
  double cutforce = 0.0;
  double **cutsq = nullptr;
  int **setflag = nullptr;
  int comm_forward = 0, comm_reverse = 0, comm_reverse_off = 0;
  int single_enable = 1, respa_enable = 0, one_coeff = 0, manybody_flag = 0, restartinfo = 1;
  int no_virial_fdotr = 0, ghostneigh = 0, unit_convert_flag = 0;
  int nextra = 0; double *pvector = nullptr;
  int evflag = 0, eflag_either = 0, eflag_global = 0, eflag_atom = 0;
  int vflag_either = 0, vflag_global = 0, vflag_atom = 0, cvflag_atom = 0;
  int vflag_fdotr = 0;
  int maxeatom = 0, maxvatom = 0;
  int allocated = 0, copymode = 0, suffix_flag = 0;
  int instance_me = 0;
  int nparams = 0, maxparam = 0;
  int *map = nullptr;
  NeighList *list = nullptr;

  explicit Pair(LAMMPS *p) : Pointers(p) {}
  ~Pair() override {
    if (copymode) return;
    memory->destroy(eatom);
    memory->destroy(vatom);
    delete[] map;
  }

  virtual void compute(int, int) = 0;
  virtual void settings(int, char **) = 0;
  virtual void coeff(int, char **) = 0;
  virtual void init_style() {}
  virtual double init_one(int, int) { return 0.0; }
  virtual double memory_usage() {
    return (double) maxeatom * sizeof(double) + (double) maxvatom * 6 * sizeof(double);
  }
  virtual int pack_forward_comm(int, int *, double *, int, int *) { return 0; }
  virtual void unpack_forward_comm(int, int, double *) {}
  virtual int pack_reverse_comm(int, int, double *) { return 0; }
  virtual void unpack_reverse_comm(int, int *, double *) {}
  virtual void *extract(const char *, int &) { return nullptr; }

  // what Pair::init() does for the styles used here: cutsq[i][j] = init_one(i,j)^2
  void init_cutsq() {
    int n = atom->ntypes;
    for (int i = 1; i <= n; i++)
      for (int j = i; j <= n; j++) {
        double c = init_one(i, j);
        cutsq[i][j] = cutsq[j][i] = c * c;
        if (c > cutforce) cutforce = c;
      }
  }

  void ev_init(int eflag, int vflag, int alloc = 1) {
    if (eflag || vflag) ev_setup(eflag, vflag, alloc);
    else ev_unset();
  }
  void ev_unset() {
    evflag = 0;
    eflag_either = eflag_global = eflag_atom = 0;
    vflag_either = vflag_global = vflag_atom = cvflag_atom = 0;
    vflag_fdotr = 0;
  }
  void ev_setup(int eflag, int vflag, int alloc = 1) {
    evflag = 1;
    eflag_either = eflag;
    eflag_global = eflag & ENERGY_GLOBAL;
    eflag_atom = eflag & ENERGY_ATOM;
    vflag_global = vflag & (VIRIAL_PAIR | VIRIAL_FDOTR);
    vflag_atom = vflag & VIRIAL_ATOM;
    vflag_either = vflag_global || vflag_atom;
    int nall = atom->nlocal + atom->nghost;
    if (eflag_atom && nall > maxeatom) {
      maxeatom = nall;
      if (alloc) { memory->destroy(eatom); memory->create(eatom, maxeatom, "pair:eatom"); }
    }
    if (vflag_atom && nall > maxvatom) {
      maxvatom = nall;
      if (alloc) { memory->destroy(vatom); memory->create(vatom, maxvatom, 6, "pair:vatom"); }
    }
    if (eflag_global) eng_vdwl = eng_coul = 0.0;
    if (vflag_global) for (int i = 0; i < 6; i++) virial[i] = 0.0;
    if (eflag_atom && alloc) for (int i = 0; i < nall; i++) eatom[i] = 0.0;
    if (vflag_atom && alloc)
      for (int i = 0; i < nall; i++) for (int k = 0; k < 6; k++) vatom[i][k] = 0.0;
    // global virial through F dot r when allowed; then per-pair global tally is switched off
    if (vflag_global == VIRIAL_FDOTR && no_virial_fdotr == 0) {
      vflag_fdotr = 1;
      vflag_global = 0;
      if (vflag_atom == 0) vflag_either = 0;
      if (vflag_either == 0 && eflag_either == 0) evflag = 0;
    } else {
      vflag_fdotr = 0;
    }
  }

  void ev_tally_xyz(int i, int j, int nlocal, int newton_pair, double evdwl, double ecoul,
                    double fx, double fy, double fz, double delx, double dely, double delz) {
    if (eflag_either) {
      if (eflag_global) {
        if (newton_pair) { eng_vdwl += evdwl; eng_coul += ecoul; }
        else {
          if (i < nlocal) { eng_vdwl += 0.5 * evdwl; eng_coul += 0.5 * ecoul; }
          if (j < nlocal) { eng_vdwl += 0.5 * evdwl; eng_coul += 0.5 * ecoul; }
        }
      }
      if (eflag_atom) {
        double h = 0.5 * (evdwl + ecoul);
        if (newton_pair || i < nlocal) eatom[i] += h;
        if (newton_pair || j < nlocal) eatom[j] += h;
      }
    }
    if (vflag_either) {
      double v[6] = {delx * fx, dely * fy, delz * fz, delx * fy, delx * fz, dely * fz};
      if (vflag_global) {
        if (newton_pair) for (int k = 0; k < 6; k++) virial[k] += v[k];
        else {
          if (i < nlocal) for (int k = 0; k < 6; k++) virial[k] += 0.5 * v[k];
          if (j < nlocal) for (int k = 0; k < 6; k++) virial[k] += 0.5 * v[k];
        }
      }
      if (vflag_atom) {
        if (newton_pair || i < nlocal) for (int k = 0; k < 6; k++) vatom[i][k] += 0.5 * v[k];
        if (newton_pair || j < nlocal) for (int k = 0; k < 6; k++) vatom[j][k] += 0.5 * v[k];
      }
    }
  }
  void ev_tally(int i, int j, int nlocal, int newton_pair, double evdwl, double ecoul, double fpair,
                double delx, double dely, double delz) {
    ev_tally_xyz(i, j, nlocal, newton_pair, evdwl, ecoul, delx * fpair, dely * fpair, delz * fpair,
                 delx, dely, delz);
  }

  void virial_fdotr_compute() {
    double **x = atom->x, **f = atom->f;
    int nall = atom->nlocal + atom->nghost;   // newton on: ghosts included
    if (!force->newton_pair) nall = atom->nlocal;
    for (int i = 0; i < nall; i++) {
      virial[0] += x[i][0] * f[i][0];
      virial[1] += x[i][1] * f[i][1];
      virial[2] += x[i][2] * f[i][2];
      virial[3] += f[i][1] * x[i][0];
      virial[4] += f[i][2] * x[i][0];
      virial[5] += f[i][2] * x[i][1];
    }
    vflag_fdotr = 0;
  }
};

}    // namespace LAMMPS_NS
#endif

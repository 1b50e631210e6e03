#ifndef SHIM_ATOM_H
#define SHIM_ATOM_H
#include "pointers.h"
namespace LAMMPS_NS {
class Atom {
 public:
  double **x = nullptr, **f = nullptr;
  int *type = nullptr;
  tagint *tag = nullptr;
  int nlocal = 0, nghost = 0, ntypes = 0, nmax = 0;
  int tag_enable = 1;
  int **nspecial = nullptr;
  tagint **special = nullptr;
};
}
#endif

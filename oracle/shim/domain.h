#ifndef SHIM_DOMAIN_H
#define SHIM_DOMAIN_H
#include "pointers.h"
namespace LAMMPS_NS {
class Domain {
 public:
  int triclinic = 0;
  double sublo[3] = {0, 0, 0}, subhi[3] = {0, 0, 0}, sublo_lamda[3] = {0, 0, 0}, subhi_lamda[3] = {1, 1, 1};
  void bbox(double *, double *, double *lo, double *hi) { for (int i = 0; i < 3; i++) { lo[i] = sublo[i]; hi[i] = subhi[i]; } }
};
}
#endif

#ifndef SHIM_COMM_H
#define SHIM_COMM_H
#include "pointers.h"
namespace LAMMPS_NS {
class Pair;
class Comm {
 public:
  int me = 0, nprocs = 1;
  // single-process stand-in; a harness may install a callback to emulate the ghost copy
  void (*forward_hook)(Pair *) = nullptr;
  void (*reverse_hook)(Pair *) = nullptr;
  void forward_comm(Pair *p) { if (forward_hook) forward_hook(p); }
  void reverse_comm(Pair *p) { if (reverse_hook) reverse_hook(p); }
};
}
#endif

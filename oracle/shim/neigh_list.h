#ifndef SHIM_NEIGH_LIST_H
#define SHIM_NEIGH_LIST_H
#include "pointers.h"
namespace LAMMPS_NS {
class NeighList {
 public:
  int inum = 0, gnum = 0;
  int *ilist = nullptr, *numneigh = nullptr, **firstneigh = nullptr;
};
}
#endif

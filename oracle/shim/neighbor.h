#ifndef SHIM_NEIGHBOR_H
#define SHIM_NEIGHBOR_H
#include "pointers.h"
#include "neigh_request.h"
namespace LAMMPS_NS {
class Neighbor {
 public:
  double skin = 2.0;
  int ago = 0, oneatom = 2000;
  int last_request_flags = -1;
  NeighRequest *add_request(Pair *, int flags = 0) { last_request_flags = flags; return &req; }
 private:
  NeighRequest req;
};
}
#endif

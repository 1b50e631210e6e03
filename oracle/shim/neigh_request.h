#ifndef SHIM_NEIGH_REQUEST_H
#define SHIM_NEIGH_REQUEST_H
#include "pointers.h"
namespace LAMMPS_NS {
namespace NeighConst { enum { REQ_DEFAULT = 0, REQ_FULL = 1 << 0, REQ_GHOST = 1 << 1 }; }
class NeighRequest { public: int half = 1, full = 0; };
}
#endif

#ifndef SHIM_FORCE_H
#define SHIM_FORCE_H
#include "pointers.h"
namespace LAMMPS_NS {
class Pair;
class Force { public: int newton_pair = 1; int newton = 1; Pair *pair = nullptr; };
}
#endif

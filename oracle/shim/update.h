#ifndef SHIM_UPDATE_H
#define SHIM_UPDATE_H
#include "pointers.h"
namespace LAMMPS_NS { class Update { public: bigint ntimestep = 0; }; }
#endif

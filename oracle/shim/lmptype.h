// Minimal stand-in for LAMMPS' lmptype.h, written for the oracle harness only.
// TEST INFRASTRUCTURE: lets the unmodified reference pair styles compile without a LAMMPS tree.
#ifndef SHIM_LMPTYPE_H
#define SHIM_LMPTYPE_H
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
namespace LAMMPS_NS {
typedef int tagint;
typedef int64_t bigint;
}
#define NEIGHMASK 0x1FFFFFFF
#define FLERR __FILE__, __LINE__
#endif

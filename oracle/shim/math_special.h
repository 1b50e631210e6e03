#ifndef SHIM_MATH_SPECIAL_H
#define SHIM_MATH_SPECIAL_H
namespace LAMMPS_NS { namespace MathSpecial {
static inline double square(double x) { return x * x; }
static inline double cube(double x) { return x * x * x; }
} }
#endif

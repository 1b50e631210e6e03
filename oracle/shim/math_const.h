#ifndef SHIM_MATH_CONST_H
#define SHIM_MATH_CONST_H
namespace LAMMPS_NS { namespace MathConst {
static constexpr double MY_PI = 3.14159265358979323846;
static constexpr double MY_2PI = 6.28318530717958647692;
static constexpr double MY_PI2 = 1.57079632679489661923;
} }
#endif

#ifndef SHIM_MATH_EXTRA_H
#define SHIM_MATH_EXTRA_H
namespace MathExtra {
inline double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
inline void scale3(double s, const double *v, double *ans) { ans[0] = s * v[0]; ans[1] = s * v[1]; ans[2] = s * v[2]; }
inline void scale3(double s, double *v) { v[0] *= s; v[1] *= s; v[2] *= s; }
inline double len3(const double *v) { return __builtin_sqrt(dot3(v, v)); }
inline double lensq3(const double *v) { return dot3(v, v); }
inline void sub3(const double *a, const double *b, double *ans) { ans[0] = a[0] - b[0]; ans[1] = a[1] - b[1]; ans[2] = a[2] - b[2]; }
inline void add3(const double *a, const double *b, double *ans) { ans[0] = a[0] + b[0]; ans[1] = a[1] + b[1]; ans[2] = a[2] + b[2]; }
}
#endif

#ifndef SHIM_MEMORY_H
#define SHIM_MEMORY_H
#include "pointers.h"
#include <cstdlib>
namespace LAMMPS_NS {
class Memory {
 public:
  // contiguous 2-D array with row pointers, same shape contract as LAMMPS Memory::create
  template <typename T> T **create(T **&a, int n1, int n2, const char *) {
    T *data = (T *) calloc((size_t) n1 * n2, sizeof(T));
    a = (T **) malloc(sizeof(T *) * n1);
    for (int i = 0; i < n1; i++) a[i] = data + (size_t) i * n2;
    return a;
  }
  template <typename T> T *create(T *&a, int n1, const char *) {
    a = (T *) calloc((size_t) n1, sizeof(T));
    return a;
  }
  template <typename T> void destroy(T **&a) {
    if (!a) return;
    free(a[0]); free(a); a = nullptr;
  }
  template <typename T> void destroy(T *&a) { free(a); a = nullptr; }
};
}
#endif

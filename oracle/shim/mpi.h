// Single-process MPI stub (oracle harness only): MPI_Bcast is the one call the reference
// pair styles make (pair_annp.cpp:526-584) and with one rank it is a no-op.
#ifndef SHIM_MPI_H
#define SHIM_MPI_H
#include <stddef.h>
typedef int MPI_Comm;
typedef int MPI_Datatype;
#define MPI_COMM_WORLD 0
#define MPI_INT 1
#define MPI_DOUBLE 2
#define MPI_CHAR 3
typedef int MPI_Op;
#define MPI_MIN 1
#define MPI_MAX 2
#define MPI_SUM 3
static inline int MPI_Bcast(void *, int, MPI_Datatype, int, MPI_Comm) { return 0; }
// one rank: the reduction of a value is the value (used by our own pair styles to agree on an init code)
static inline int MPI_Allreduce(const void *send, void *recv, int n, MPI_Datatype t, MPI_Op, MPI_Comm) {
  const size_t w = (t == MPI_DOUBLE) ? sizeof(double) : (t == MPI_CHAR ? 1 : sizeof(int));
  if (send != recv) __builtin_memcpy(recv, send, w * (size_t) n);
  return 0;
}
#endif

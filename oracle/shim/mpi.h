// Single-process MPI stub (oracle harness only): MPI_Bcast is the one call the reference
// pair styles make (pair_annp.cpp:526-584) and with one rank it is a no-op.
#ifndef SHIM_MPI_H
#define SHIM_MPI_H
typedef int MPI_Comm;
typedef int MPI_Datatype;
#define MPI_COMM_WORLD 0
#define MPI_INT 1
#define MPI_DOUBLE 2
#define MPI_CHAR 3
static inline int MPI_Bcast(void *, int, MPI_Datatype, int, MPI_Comm) { return 0; }
#endif

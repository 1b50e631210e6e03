// Shim of LAMMPS' Pointers/LAMMPS classes: just the members the reference pair styles touch.
#ifndef SHIM_POINTERS_H
#define SHIM_POINTERS_H
#include "lmptype.h"
#include "mpi.h"
namespace LAMMPS_NS {
class Memory; class Error; class Atom; class Force; class Comm; class Neighbor; class Update;
class Domain; class Modify;
class LAMMPS {
 public:
  Memory *memory = nullptr; Error *error = nullptr; Atom *atom = nullptr; Force *force = nullptr;
  Comm *comm = nullptr; Neighbor *neighbor = nullptr; Update *update = nullptr;
  Domain *domain = nullptr; Modify *modify = nullptr;
  FILE *screen = nullptr; MPI_Comm world = MPI_COMM_WORLD;
};
class Pointers {
 public:
  explicit Pointers(LAMMPS *p)
      : lmp(p), memory(p->memory), error(p->error), atom(p->atom), force(p->force), comm(p->comm),
        neighbor(p->neighbor), update(p->update), domain(p->domain), modify(p->modify),
        screen(p->screen), world(p->world) {}
  virtual ~Pointers() {}
 protected:
  LAMMPS *lmp;
  Memory *&memory; Error *&error; Atom *&atom; Force *&force; Comm *&comm; Neighbor *&neighbor;
  Update *&update; Domain *&domain; Modify *&modify; FILE *&screen; MPI_Comm &world;
};
}
#endif

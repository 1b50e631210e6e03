// Oracle harness driver (TEST INFRASTRUCTURE — never linked into the product path).
//
// Instantiates a LAMMPS pair style compiled against oracle/shim (the UNMODIFIED reference source
// from /root/reference, or our own LAMMPS-facing style) and performs what LAMMPS does around
// Pair::compute for one force evaluation:
//   pair_style <name>; pair_coeff * * <file> <elem...>; Pair::init (init_style + cutsq=init_one^2);
//   compute(eflag, vflag)   (reference call stack: SURVEY.md section 3 A-D)
// Input  (binary, little endian): see read_input();  Output: see write_output().
//
// build:  g++ -O2 -std=c++17 -Ishim -DDRIVER_PAIR_HEADER='"pair_annp.h"' -DDRIVER_PAIR=PairANNP \
//             -I/root/reference/annp-gpu-lammps/fe_v2/src ref_driver.cpp <pair source> -o _ref/ref_annp_fe
#include DRIVER_PAIR_HEADER
#include "atom.h"
#include "comm.h"
#include "domain.h"
#include "error.h"
#include "force.h"
#include "memory.h"
#include "neigh_list.h"
#include "neighbor.h"
#include "update.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace LAMMPS_NS;

namespace {
struct Input {
  int nlocal, nghost, ntypes, inum, eflag, vflag, ncalls, reserved;
  std::vector<double> x;
  std::vector<int> type, ilist, numneigh, neigh, ghost_owner;
};

void rd(FILE *fp, void *p, size_t n) {
  if (fread(p, 1, n, fp) != n) { fprintf(stderr, "ref_driver: short read\n"); exit(2); }
}

Input read_input(const char *path) {
  Input in;
  FILE *fp = fopen(path, "rb");
  if (!fp) { fprintf(stderr, "ref_driver: cannot open %s\n", path); exit(2); }
  int hdr[8];
  rd(fp, hdr, sizeof hdr);
  in.nlocal = hdr[0]; in.nghost = hdr[1]; in.ntypes = hdr[2]; in.inum = hdr[3];
  in.eflag = hdr[4]; in.vflag = hdr[5]; in.ncalls = hdr[6]; in.reserved = hdr[7];
  int nall = in.nlocal + in.nghost;
  in.x.resize((size_t) nall * 3); rd(fp, in.x.data(), in.x.size() * 8);
  in.type.resize(nall); rd(fp, in.type.data(), (size_t) nall * 4);
  in.ilist.resize(in.inum); rd(fp, in.ilist.data(), (size_t) in.inum * 4);
  in.numneigh.resize(in.inum); rd(fp, in.numneigh.data(), (size_t) in.inum * 4);
  size_t tot = 0;
  for (int v : in.numneigh) tot += v;
  in.neigh.resize(tot); rd(fp, in.neigh.data(), tot * 4);
  in.ghost_owner.resize(in.nghost); rd(fp, in.ghost_owner.data(), (size_t) in.nghost * 4);
  fclose(fp);
  return in;
}

// forward_comm emulation for styles that exchange per-atom scalars (anna_adp): ghost <- owner
const Input *g_in = nullptr;
void forward_hook(Pair *p) {
  int n = p->comm_forward;
  if (n <= 0 || g_in->nghost == 0) return;
  std::vector<double> buf((size_t) n * g_in->nghost);
  std::vector<int> lst(g_in->ghost_owner);
  int pbc[3] = {0, 0, 0};
  p->pack_forward_comm(g_in->nghost, lst.data(), buf.data(), 0, pbc);
  p->unpack_forward_comm(g_in->nghost, g_in->nlocal, buf.data());
}
// reverse_comm emulation (Comm::reverse_comm(Pair *)): ghost -> owner, summed
void reverse_hook(Pair *p) {
  int n = p->comm_reverse;
  if (n <= 0 || g_in->nghost == 0) return;
  std::vector<double> buf((size_t) n * g_in->nghost);
  std::vector<int> lst(g_in->ghost_owner);
  p->pack_reverse_comm(g_in->nghost, g_in->nlocal, buf.data());
  p->unpack_reverse_comm(g_in->nghost, lst.data(), buf.data());
}
}    // namespace

int main(int argc, char **argv) {
  if (argc < 5) {
    fprintf(stderr, "usage: %s <input.bin> <output.bin> <potential file> <elem per type...>\n", argv[0]);
    return 2;
  }
  Input in = read_input(argv[1]);
  g_in = &in;
  const int nall = in.nlocal + in.nghost;

  LAMMPS lmp;
  Memory memory; Error error; Atom atom; Force force; Comm comm; Neighbor neighbor; Update update; Domain domain;
  lmp.memory = &memory; lmp.error = &error; lmp.atom = &atom; lmp.force = &force; lmp.comm = &comm;
  lmp.neighbor = &neighbor; lmp.update = &update; lmp.domain = &domain; lmp.screen = nullptr;
  comm.forward_hook = forward_hook;
  comm.reverse_hook = reverse_hook;
#ifdef NEWTON_PAIR
  force.newton_pair = NEWTON_PAIR;
#endif
  if (const char *nw = getenv("ANNP_DRIVER_NEWTON")) force.newton_pair = force.newton = atoi(nw);    // `newton on|off` of the deck

  atom.nlocal = in.nlocal; atom.nghost = in.nghost; atom.ntypes = in.ntypes; atom.nmax = nall;
  std::vector<double *> xrow(nall), frow(nall);
  std::vector<double> f((size_t) nall * 3, 0.0);
  for (int i = 0; i < nall; i++) { xrow[i] = &in.x[(size_t) i * 3]; frow[i] = &f[(size_t) i * 3]; }
  atom.x = xrow.data(); atom.f = frow.data(); atom.type = in.type.data();
  std::vector<int> tag(nall);
  for (int i = 0; i < nall; i++) tag[i] = (i < in.nlocal ? i : in.ghost_owner[i - in.nlocal]) + 1;
  atom.tag = tag.data();

  NeighList list;
  std::vector<int> numneigh_by_atom(nall, 0);
  std::vector<int *> firstneigh(nall, nullptr);
  {
    size_t off = 0;
    for (int ii = 0; ii < in.inum; ii++) {
      int i = in.ilist[ii];
      numneigh_by_atom[i] = in.numneigh[ii];
      firstneigh[i] = in.neigh.data() + off;
      off += in.numneigh[ii];
    }
  }
  list.inum = in.inum; list.ilist = in.ilist.data();
  list.numneigh = numneigh_by_atom.data(); list.firstneigh = firstneigh.data();

  int rc = 0;
  try {
    DRIVER_PAIR pair(&lmp);
    pair.list = &list;
    pair.settings(0, nullptr);
    std::vector<char *> cargs;
    char star[] = "*";
    cargs.push_back(star); cargs.push_back(star);
    for (int a = 3; a < argc; a++) cargs.push_back(argv[a]);
    pair.coeff((int) cargs.size(), cargs.data());
    // the only pair style of the run (Force::pair), unless the harness asks for the sub-style situation of pair hybrid
    if (!getenv("ANNP_DRIVER_HYBRID")) force.pair = &pair;
    pair.init_style();
    pair.init_cutsq();

    double secs = 0.0;
    std::vector<double> per_call;
    for (int call = 0; call < (in.ncalls > 0 ? in.ncalls : 1); call++) {
      std::fill(f.begin(), f.end(), 0.0);       // Verlet::force_clear
      neighbor.ago = call;                      // the list is new on the first call only (no atom moves here)
      auto t0 = std::chrono::steady_clock::now();
      pair.compute(in.eflag, in.vflag);
      auto t1 = std::chrono::steady_clock::now();
      per_call.push_back(std::chrono::duration<double>(t1 - t0).count());
      secs += per_call.back();
    }

    FILE *fp = fopen(argv[2], "wb");
    if (!fp) { fprintf(stderr, "ref_driver: cannot write %s\n", argv[2]); return 2; }
    int hdr[4] = {nall, pair.eatom ? 1 : 0, pair.vatom ? 1 : 0, in.ncalls > 0 ? in.ncalls : 1};
    fwrite(hdr, sizeof hdr, 1, fp);
    fwrite(&pair.eng_vdwl, 8, 1, fp);
    fwrite(pair.virial, 8, 6, fp);
    fwrite(&secs, 8, 1, fp);
    fwrite(f.data(), 8, f.size(), fp);
    if (pair.eatom) fwrite(pair.eatom, 8, nall, fp);
    if (pair.vatom) for (int i = 0; i < nall; i++) fwrite(pair.vatom[i], 8, 6, fp);
    fwrite(per_call.data(), 8, per_call.size(), fp);      // wall seconds of every compute() call
    fclose(fp);
  } catch (std::exception &e) {
    fprintf(stderr, "%s\n", e.what());
    rc = 1;
  }
  return rc;
}

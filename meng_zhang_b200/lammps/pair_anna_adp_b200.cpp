/* ----------------------------------------------------------------------
   pair_style anna_adp/gpu on libannp_b200.so -- host side inside LAMMPS.

   Mirrors PairANNAADPGPU of the reference (anna-gpu-lammps/bcc_fe/src/pair_anna_adp_gpu.cpp):
     constructor 66-71, destructor 76-78, memory_usage 81-84, compute 89-157, init_style 162-260.
   Differences, all on purpose:
     * anna_adp_gpu_init / _compute / _compute_force / _clear / _bytes become anna_b200_init and the shared
       annp_b200_* entry points of include/annp_b200.h
     * no forward communication of rho / mu / lambda / d2 / q2: the numbers follow the reference CPU style, which
       centres everything on the local atom; `newton on` and `newton off` decks both work (see the header)
     * forces / energies are ADDED to LAMMPS' arrays; no `fix gpu` / `package gpu`; no CPU fallback
------------------------------------------------------------------------- */

#include "pair_anna_adp_b200.h"

#include "annp_b200.h"

#include "atom.h"
#include "comm.h"
#include "error.h"
#include "force.h"
#include "memory.h"
#include "neigh_list.h"
#include "neigh_request.h"
#include "neighbor.h"
#include "suffix.h"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace LAMMPS_NS;

/* ---------------------------------------------------------------------- */

PairANNAADPB200::PairANNAADPB200(LAMMPS *lmp) : PairANNA_ADP(lmp), handle(nullptr), fstage(nullptr), device_neigh(0)
{
  respa_enable = 0;
  suffix_flag |= Suffix::GPU;
  comm_forward = 0;      // the CPU base class reserves one slot it never uses; nothing is forwarded here
  comm_reverse = 3;      // ghost forces, only used under `newton off`
}

/* ---------------------------------------------------------------------- */

PairANNAADPB200::~PairANNAADPB200()
{
  hb.clear();
  annp_b200_clear(handle);
  handle = nullptr;
}

/* ---------------------------------------------------------------------- */

double PairANNAADPB200::memory_usage()
{
  return Pair::memory_usage() + hb.bytes() + annp_b200_bytes(handle);
}

/* ----------------------------------------------------------------------
   compute force and energy   (reference: pair_anna_adp_gpu.cpp:89-157, one library call instead of two phases)
------------------------------------------------------------------------- */

void PairANNAADPB200::compute(int eflag, int vflag)
{
  ev_init(eflag, vflag);
  const int nlocal = atom->nlocal, nghost = atom->nghost, nall = nlocal + nghost;
  const bool rebuilt = neighbor->ago == 0;
  int rc;

  if (nall > 0) hb.x.track(atom->x[0], sizeof(double) * 3 * (size_t) atom->nmax);

  if (rebuilt) {
    if (device_neigh) {
      double lo[3], hi[3];
      ANNP_B200_NS::bounds_of(nall > 0 ? atom->x[0] : nullptr, nall, lo, hi);
      rc = annp_b200_neigh_build_host(handle, nlocal, nall, nall > 0 ? atom->x[0] : nullptr, lo, hi, cutmax + neighbor->skin);
    } else {
      rc = annp_b200_neigh(handle, list->inum, nall, list->ilist, list->numneigh, list->firstneigh);
    }
    if (rc == ANNP_B200_ENOMEM) error->one(FLERR, "Insufficient memory on accelerator");
    if (rc) error->one(FLERR, std::string("anna_adp/gpu: ") + annp_b200_last_error(handle));
  }
  if (nall == 0) return;

  // only pair style of a newton-on run: the device writes straight into LAMMPS' page-locked force array (see
  // annp_b200_host.h); otherwise the forces are staged (pair hybrid: added; newton off: folded by the style itself)
  const bool direct = force->pair == this && force->newton_pair && !ANNP_B200_NS::force_add_requested();
  double *f0 = atom->f[0];
  double *fdst = f0;
  if (direct) hb.f.track(f0, sizeof(double) * 3 * (size_t) atom->nmax);
  else fdst = hb.fbuf.reserve(3 * (size_t) nall);
  double *ebuf = eflag_atom ? hb.ebuf.reserve((size_t) nall) : nullptr;
  double *vbuf = vflag_atom ? hb.vbuf.reserve(6 * (size_t) nall) : nullptr;
  if (!fdst || (eflag_atom && !ebuf) || (vflag_atom && !vbuf)) error->one(FLERR, "Out of host memory in pair anna_adp/gpu");

  double eng = 0.0, vir[6] = {0, 0, 0, 0, 0, 0};
  const int want_pair_virial = vflag_global && !vflag_fdotr;
  rc = annp_b200_compute(handle, nlocal, nghost, atom->x[0], rebuilt ? atom->type : nullptr, eflag_either, vflag_either || vflag_fdotr,
                         fdst, eflag_global ? &eng : nullptr, ebuf, want_pair_virial ? vir : nullptr, vbuf);
  if (rc == ANNP_B200_ENOMEM) error->one(FLERR, "Insufficient memory on accelerator");
  if (rc) error->one(FLERR, std::string("anna_adp/gpu: ") + annp_b200_last_error(handle));

  if (force->newton_pair) {
    // local and ghost rows: LAMMPS' reverse_comm carries the ghost part home
    if (!direct) ANNP_B200_NS::add_into(f0, fdst, 3 * (size_t) nall);
  } else {
    // `newton off`: LAMMPS will not reverse-communicate forces, so the style does it for its own ghost rows
    fstage = fdst;
    comm->reverse_comm(this);
    fstage = nullptr;
    ANNP_B200_NS::add_into(f0, fdst, 3 * (size_t) nlocal);
  }
  if (eflag_global) eng_vdwl += eng;
  if (eflag_atom) ANNP_B200_NS::add_into(eatom, ebuf, (size_t) nall);
  if (want_pair_virial) for (int k = 0; k < 6; k++) virial[k] += vir[k];
  if (vflag_atom) ANNP_B200_NS::add_into(vatom[0], vbuf, 6 * (size_t) nall);

  if (vflag_fdotr) virial_fdotr_compute();
}

/* ----------------------------------------------------------------------
   ghost forces -> owners under `newton off` (Comm::reverse_comm(Pair *))
------------------------------------------------------------------------- */

int PairANNAADPB200::pack_reverse_comm(int n, int first, double *buf)
{
  int m = 0;
  for (int i = first; i < first + n; i++) {
    buf[m++] = fstage[3 * (size_t) i];
    buf[m++] = fstage[3 * (size_t) i + 1];
    buf[m++] = fstage[3 * (size_t) i + 2];
  }
  return m;
}

void PairANNAADPB200::unpack_reverse_comm(int n, int *list, double *buf)
{
  int m = 0;
  for (int i = 0; i < n; i++) {
    const int j = list[i];
    fstage[3 * (size_t) j] += buf[m++];
    fstage[3 * (size_t) j + 1] += buf[m++];
    fstage[3 * (size_t) j + 2] += buf[m++];
  }
}

/* ----------------------------------------------------------------------
   init specific to this pair style   (reference: pair_anna_adp_gpu.cpp:162-260)
------------------------------------------------------------------------- */

void PairANNAADPB200::init_style()
{
  if (atom->tag_enable == 0) error->all(FLERR, "Pair style anna_adp/gpu requires atom IDs");

  const ANNAPARA &p = params[0];
  const int ntl = p.ntl, nnod = p.nnod, nsf = p.nsf, nout = p.nout, nelements = p.nelements, ntypes = atom->ntypes;
  if (ntl - 1 > ANNP_B200_MAX_LAYERS) error->all(FLERR, "anna_adp/gpu: too many network layers");

  for (int i = 1; i <= ntypes; i++)
    for (int j = i; j <= ntypes; j++) {
      double cut = 0.0;
      if (setflag[i][j] != 0 || (setflag[i][i] != 0 && setflag[j][j] != 0)) { cut = init_one(i, j); cut *= cut; }
      cutsq[i][j] = cutsq[j][i] = cut;
    }

  // weights/biases flattened row-major per layer; the last layer has nout rows (reference lines 196-222)
  std::vector<double> w, b;
  for (int e = 0; e < nelements; e++)
    for (int l = 0; l < ntl - 1; l++) {
      const int nrow = (l == ntl - 2) ? nout : nnod, ncol = (l == 0) ? nsf : nnod;
      for (int j = 0; j < nrow; j++)
        for (int k = 0; k < ncol; k++) w.push_back(p.all_anna[e].weight_all[l][j][k]);
      for (int j = 0; j < nrow; j++) b.push_back(p.all_anna[e].bias_all[l][0][j]);
    }

  std::vector<double> cs((size_t) (ntypes + 1) * (ntypes + 1), 0.0);
  for (int i = 1; i <= ntypes; i++) for (int j = 1; j <= ntypes; j++) cs[(size_t) i * (ntypes + 1) + j] = cutsq[i][j];
  std::vector<int> mp(ntypes + 1, 0);
  for (int i = 1; i <= ntypes; i++) mp[i] = map[i] < 0 ? 0 : map[i];

  anna_b200_params P;
  memset(&P, 0, sizeof(P));
  P.abi_version = ANNP_B200_ABI_VERSION;
  P.ntypes = ntypes; P.nelements = nelements;
  P.ntl = ntl; P.nhl = p.nhl; P.nnod = nnod; P.nout = nout; P.nsf = nsf; P.npsf = p.npsf; P.ntsf = p.ntsf; P.ngp = p.ngp;
  P.flagsym = p.flagsym;
  for (int l = 0; l < ntl - 1; l++) P.flagact[l] = p.flagact[l];
  P.e_base = p.e_base; P.cut = p.cut;
  P.cutsq = cs.data(); P.map = mp.data(); P.weights = w.data(); P.bias = b.data(); P.gparams = p.gparams;

  annp_b200_clear(handle);
  handle = nullptr;
  const int ndev = annp_b200_device_count();
  char msg[512] = "";
  const int device = ndev > 0 ? comm->me % ndev : 0;      // one rank per GPU
  int rc = anna_b200_init(&P, device, &handle, msg, (int) sizeof(msg));
  // reduce the code over all ranks before aborting, as GPU_EXTRA::check_flag does for the reference (codes are <= 0)
  int rc_all = rc;
  MPI_Allreduce(&rc, &rc_all, 1, MPI_INT, MPI_MIN, world);
  if (rc_all == ANNP_B200_ENOMEM) error->all(FLERR, "Insufficient memory on accelerator");
  if (rc_all != 0) error->all(FLERR, std::string("anna_adp/gpu initialisation failed") + (rc != 0 ? std::string(": ") + msg : std::string(" on another rank")));
  // The input deck keeps the reference's zero-argument pair_style line; ANNP_B200_SCATTER=gather selects the ordered
  // FP64 gather instead of the default fixed-point force accumulation (include/annp_b200.h: annp_b200_set_scatter)
  if (const char *sc = getenv("ANNP_B200_SCATTER")) {
    const int mode = !strcmp(sc, "gather") ? ANNP_B200_SCATTER_GATHER : (!strcmp(sc, "fixed") ? ANNP_B200_SCATTER_FIXED : -1);
    if (mode < 0 || annp_b200_set_scatter(handle, mode) != 0) error->all(FLERR, "ANNP_B200_SCATTER must be 'fixed' or 'gather'");
  }

  device_neigh = ANNP_B200_NS::device_neigh_requested() ? 1 : 0;      // `package gpu N neigh yes` of the reference
  if (!device_neigh) neighbor->add_request(this, NeighConst::REQ_FULL);
}

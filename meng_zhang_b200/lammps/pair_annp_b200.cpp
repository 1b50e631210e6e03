/* ----------------------------------------------------------------------
   pair_style annp/gpu on libannp_b200.so -- host side inside LAMMPS.

   Mirrors PairANNPGPU of the reference (annp-gpu-lammps/fe_v2/src/pair_annp_gpu.cpp):
     constructor 61-66, destructor 71-73, memory_usage 76-79, compute 84-131, init_style 136-243.
   Differences, all on purpose:
     * the five annp_gpu_* library calls become the C ABI of include/annp_b200.h
     * forces / energies are ADDED to LAMMPS' arrays (the reference assigns f[][], which breaks
       pair hybrid/overlay, lal_annp.cpp:336-347)
     * no `fix gpu` / `package gpu`: the rank picks device (rank mod #devices)
     * no CPU fallback: a missing device is a fatal error at init_style
------------------------------------------------------------------------- */

#include "pair_annp_b200.h"

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
#include <type_traits>
#include <utility>
#include <vector>

using namespace LAMMPS_NS;

/* ----------------------------------------------------------------------
   The three `annp` copies of the reference share the class name PairANNP and the style name; a LAMMPS tree holds
   ONE of them.  This file compiles against whichever pair_annp.h is installed: the Ni copy's parameter struct
   carries sf_min / sf_max / sym_coerad / sym_coeang (ni/src/pair_annp.h), the Fe copies' carries
   sfnor_cov / sfnor_avg (fe_v2/src/pair_annp.h).  The helpers below pick the members that exist.
------------------------------------------------------------------------- */
namespace {

template <class P, class = void> struct is_ni_copy : std::false_type {};
template <class P> struct is_ni_copy<P, std::void_t<decltype(std::declval<P>().sym_coerad)>> : std::true_type {};

struct DescriptorArrays {
  std::vector<double> scal, avg, corad, coang;
  int variant = ANNP_B200_VARIANT_FE;
};

// Fe / fe_v2 copy: sfnor_scal = 1/sqrt(cov - avg^2), 0 when degenerate (pair_annp_gpu.cpp:211-220)
template <class P> typename std::enable_if<!is_ni_copy<P>::value>::type fill_descriptor(const P &p, DescriptorArrays &d)
{
  const int nsf = p.nsf;
  d.scal.resize(nsf); d.avg.resize(nsf);
  for (int i = 0; i < nsf; i++) {
    d.avg[i] = p.sfnor_avg[i];
    const double t = sqrt(p.sfnor_cov[i] - d.avg[i] * d.avg[i]);
    d.scal[i] = (t <= 1.0e-10) ? 0.0 : 1.0 / t;
  }
  d.variant = ANNP_B200_VARIANT_FE;
}

// Ni copy: (G - sf_min) / (sf_max - sf_min) and the Behler-Parrinello coefficient tables
// (ni/src/pair_annp.cpp:99-101,168-170; ni/src/pair_annp_gpu.cpp:31-40 host_cofsymrad / host_cofsymang).
// The Ni CPU style overwrites sf_max with the range on every compute() call; init_style runs before the first one,
// so sf_max still holds the file's values here.
template <class P> typename std::enable_if<is_ni_copy<P>::value>::type fill_descriptor(const P &p, DescriptorArrays &d)
{
  const int nsf = p.nsf;
  d.scal.resize(nsf); d.avg.resize(nsf);
  for (int i = 0; i < nsf; i++) { d.avg[i] = p.sf_min[i]; d.scal[i] = 1.0 / (p.sf_max[i] - p.sf_min[i]); }
  d.corad.resize((size_t) p.npsf * 3);
  d.coang.resize((size_t) p.ntsf * 4);
  for (int m = 0; m < p.npsf; m++) for (int k = 0; k < 3; k++) d.corad[(size_t) m * 3 + k] = p.sym_coerad[m][k];
  for (int n = 0; n < p.ntsf; n++) for (int k = 0; k < 4; k++) d.coang[(size_t) n * 4 + k] = p.sym_coeang[n][k];
  d.variant = ANNP_B200_VARIANT_NI;
}

}    // namespace

/* ---------------------------------------------------------------------- */

PairANNPB200::PairANNPB200(LAMMPS *lmp) : PairANNP(lmp), handle(nullptr), device_neigh(0)
{
  respa_enable = 0;
  suffix_flag |= Suffix::GPU;
}

/* ---------------------------------------------------------------------- */

PairANNPB200::~PairANNPB200()
{
  hb.clear();
  annp_b200_clear(handle);
  handle = nullptr;
}

/* ---------------------------------------------------------------------- */

double PairANNPB200::memory_usage()
{
  return Pair::memory_usage() + hb.bytes() + annp_b200_bytes(handle);
}

/* ----------------------------------------------------------------------
   compute force and energy   (reference: pair_annp_gpu.cpp:84-131)
------------------------------------------------------------------------- */

void PairANNPB200::compute(int eflag, int vflag)
{
  ev_init(eflag, vflag);
  const int nlocal = atom->nlocal, nghost = atom->nghost, nall = nlocal + nghost;
  const bool rebuilt = neighbor->ago == 0;
  int rc;

  // LAMMPS' position array, page-locked in place (re-registered when Memory::grow moved it)
  if (nall > 0) hb.x.track(atom->x[0], sizeof(double) * 3 * (size_t) atom->nmax);

  if (rebuilt) {
    if (device_neigh) {     // annp_gpu_compute_n of the reference: the list is built on the device from the positions
      double lo[3], hi[3];
      ANNP_B200_NS::bounds_of(nall > 0 ? atom->x[0] : nullptr, nall, lo, hi);
      rc = annp_b200_neigh_build_host(handle, nlocal, nall, nall > 0 ? atom->x[0] : nullptr, lo, hi, cutmax + neighbor->skin);
    } else {                // reset_nbors of the reference: hand LAMMPS' full list to the device
      rc = annp_b200_neigh(handle, list->inum, nall, list->ilist, list->numneigh, list->firstneigh);
    }
    if (rc == ANNP_B200_ENOMEM) error->one(FLERR, "Insufficient memory on accelerator");
    if (rc) error->one(FLERR, std::string("annp/gpu: ") + annp_b200_last_error(handle));
  }
  if (nall == 0) return;

  // Only pair style of the run: LAMMPS has zeroed f, the device writes the forces straight into the page-locked array
  // (the reference assigns f too, lal_annp.cpp:345-347).  As a sub-style of pair hybrid / overlay they are staged and added.
  const bool sole = force->pair == this && !ANNP_B200_NS::force_add_requested();
  double *f0 = atom->f[0];
  double *fdst = f0;
  if (sole) hb.f.track(f0, sizeof(double) * 3 * (size_t) atom->nmax);
  else fdst = hb.fbuf.reserve(3 * (size_t) nall);
  double *ebuf = eflag_atom ? hb.ebuf.reserve((size_t) nall) : nullptr;
  double *vbuf = vflag_atom ? hb.vbuf.reserve(6 * (size_t) nall) : nullptr;
  if (!fdst || (eflag_atom && !ebuf) || (vflag_atom && !vbuf)) error->one(FLERR, "Out of host memory in pair annp/gpu");

  double eng = 0.0, vir[6] = {0, 0, 0, 0, 0, 0};
  const int want_pair_virial = vflag_global && !vflag_fdotr;
  // types travel with the list: atoms keep their slots (and types) until the next re-neighbouring
  rc = annp_b200_compute(handle, nlocal, nghost, atom->x[0], rebuilt ? atom->type : nullptr, eflag_either, vflag_either || vflag_fdotr,
                         fdst, eflag_global ? &eng : nullptr, ebuf, want_pair_virial ? vir : nullptr, vbuf);
  if (rc == ANNP_B200_ENOMEM) error->one(FLERR, "Insufficient memory on accelerator");
  if (rc) error->one(FLERR, std::string("annp/gpu: ") + annp_b200_last_error(handle));

  // local and ghost rows: LAMMPS' reverse_comm carries the ghost part home (newton_pair on)
  if (!sole) ANNP_B200_NS::add_into(f0, fdst, 3 * (size_t) nall);
  if (eflag_global) eng_vdwl += eng;
  if (eflag_atom) ANNP_B200_NS::add_into(eatom, ebuf, (size_t) nall);
  if (want_pair_virial) for (int k = 0; k < 6; k++) virial[k] += vir[k];
  if (vflag_atom) ANNP_B200_NS::add_into(vatom[0], vbuf, 6 * (size_t) nall);

  if (vflag_fdotr) virial_fdotr_compute();
}

/* ----------------------------------------------------------------------
   init specific to this pair style   (reference: pair_annp_gpu.cpp:136-243)
------------------------------------------------------------------------- */

void PairANNPB200::init_style()
{
  if (atom->tag_enable == 0) error->all(FLERR, "Pair style annp/gpu requires atom IDs");
  if (force->newton_pair == 0) error->all(FLERR, "Pair style annp/gpu requires newton pair on");

  const Param_ANNP &p = params[0];
  const int ntl = p.ntl, nnod = p.nnod, nsf = p.nsf, nelements = p.nelements, ntypes = atom->ntypes;
  if (ntl - 1 > ANNP_B200_MAX_LAYERS) error->all(FLERR, "annp/gpu: too many network layers");

  // cutsq exactly as the reference fills it (lines 175-188)
  for (int i = 1; i <= ntypes; i++)
    for (int j = i; j <= ntypes; j++) {
      double cut = 0.0;
      if (setflag[i][j] != 0 || (setflag[i][i] != 0 && setflag[j][j] != 0)) { cut = init_one(i, j); cut *= cut; }
      cutsq[i][j] = cutsq[j][i] = cut;
    }

  // weights/biases flattened row-major per layer, [k + j*ncol] (lines 190-209)
  const size_t wpe = annp_b200_weights_per_element(ntl, nnod, nsf), bpe = annp_b200_bias_per_element(ntl, nnod);
  std::vector<double> w(wpe * nelements, 0.0), b(bpe * nelements, 0.0);
  for (int e = 0; e < nelements; e++) {
    size_t wo = wpe * e, bo = bpe * e;
    for (int l = 0; l < ntl - 1; l++) {
      const int nrow = (l == ntl - 2) ? 1 : nnod, ncol = (l == 0) ? nsf : nnod;
      for (int j = 0; j < nrow; j++)
        for (int k = 0; k < ncol; k++) w[wo + k + (size_t) j * ncol] = p.all_annp[e].weight_all[l][j][k];
      for (int j = 0; j < nrow; j++) b[bo + j] = p.all_annp[e].bias_all[l][0][j];
      wo += (size_t) nrow * ncol;
      bo += nrow;
    }
  }

  // descriptor normalisation (and, for the Ni copy, the symmetry-function coefficient tables)
  DescriptorArrays desc;
  fill_descriptor(p, desc);

  std::vector<double> cs((size_t) (ntypes + 1) * (ntypes + 1), 0.0);
  for (int i = 1; i <= ntypes; i++) for (int j = 1; j <= ntypes; j++) cs[(size_t) i * (ntypes + 1) + j] = cutsq[i][j];
  std::vector<int> mp(ntypes + 1, 0);
  for (int i = 1; i <= ntypes; i++) mp[i] = map[i] < 0 ? 0 : map[i];

  annp_b200_params P;
  memset(&P, 0, sizeof(P));
  P.abi_version = ANNP_B200_ABI_VERSION;
  P.ntypes = ntypes; P.nelements = nelements;
  P.ntl = ntl; P.nhl = p.nhl; P.nnod = nnod; P.nsf = nsf; P.npsf = p.npsf; P.ntsf = p.ntsf;
  P.flagsym = p.flagsym;
  for (int l = 0; l < ntl - 1; l++) P.flagact[l] = p.flagact[l];
  P.e_scale = p.e_scale; P.e_shift = p.e_shift; P.e_atom = p.e_atom; P.cut = p.cut;
  P.sfnor_scal = desc.scal.data(); P.sfnor_avg = desc.avg.data(); P.cutsq = cs.data(); P.map = mp.data();
  P.weights = w.data(); P.bias = b.data();
  P.variant = desc.variant;
  if (desc.variant == ANNP_B200_VARIANT_NI) { P.sym_coerad = desc.corad.data(); P.sym_coeang = desc.coang.data(); }

  annp_b200_clear(handle);
  handle = nullptr;
  const int ndev = annp_b200_device_count();
  char msg[512] = "";
  const int device = ndev > 0 ? comm->me % ndev : 0;      // one rank per GPU
  int rc = annp_b200_init(&P, device, atom->nlocal + atom->nghost, 0, &handle, msg, (int) sizeof(msg));
  // the reference funnels init codes through GPU_EXTRA::check_flag, which reduces them over all ranks before aborting
  // (line 235): a failure on one rank (out of memory, no device) must stop every rank, or the others hang in the next
  // collective.  Codes are <= 0, so the minimum over the ranks is the worst one.
  int rc_all = rc;
  MPI_Allreduce(&rc, &rc_all, 1, MPI_INT, MPI_MIN, world);
  if (rc_all == ANNP_B200_ENOMEM) error->all(FLERR, "Insufficient memory on accelerator");
  if (rc_all != 0) error->all(FLERR, std::string("annp/gpu initialisation failed") + (rc != 0 ? std::string(": ") + msg : std::string(" on another rank")));
  // The input deck keeps the reference's zero-argument pair_style line; ANNP_B200_SCATTER=gather selects the ordered
  // FP64 gather instead of the default fixed-point force accumulation (include/annp_b200.h: annp_b200_set_scatter)
  if (const char *sc = getenv("ANNP_B200_SCATTER")) {
    const int mode = !strcmp(sc, "gather") ? ANNP_B200_SCATTER_GATHER : (!strcmp(sc, "fixed") ? ANNP_B200_SCATTER_FIXED : -1);
    if (mode < 0 || annp_b200_set_scatter(handle, mode) != 0) error->all(FLERR, "ANNP_B200_SCATTER must be 'fixed' or 'gather'");
  }

  // ANNP_B200_NEIGH=device: the reference's GPU_NEIGH mode (`package gpu N neigh yes`, annp_gpu_compute_n): the library
  // builds the full list on the device from the positions, LAMMPS builds none for this style
  device_neigh = ANNP_B200_NS::device_neigh_requested() ? 1 : 0;
  if (!device_neigh) neighbor->add_request(this, NeighConst::REQ_FULL);
}

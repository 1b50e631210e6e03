/*
 * annp_b200.h -- C ABI of libannp_b200.so: the B200 (sm_100a) implementation of the ANNP
 * neural-network-potential force evaluation behind LAMMPS `pair_style annp/gpu`.
 *
 * This is the drop-in boundary.  Each entry point names the reference interface it replaces
 * (paths relative to the reference repo, annp-gpu-lammps/fe_v2/):
 *
 *   annp_b200_init            <- annp_gpu_init      src/pair_annp_gpu.cpp:31-39, lib/lal_annp_ext.cpp:25-92
 *   annp_b200_neigh           <- ANNP::reset_nbors  lib/lal_annp.cpp:301-304 (ago == 0 branch of compute)
 *   annp_b200_compute         <- annp_gpu_compute   src/pair_annp_gpu.cpp:51-56, lib/lal_annp_ext.cpp:109-119
 *   annp_b200_clear           <- annp_gpu_clear     src/pair_annp_gpu.cpp:41,   lib/lal_annp_ext.cpp:94-96
 *   annp_b200_bytes           <- annp_gpu_bytes     src/pair_annp_gpu.cpp:58,   lib/lal_annp_ext.cpp:121-123
 *   annp_b200_read_potential  <- PairANNP::read_file src/pair_annp.cpp:332-585 (potential file format)
 *   annp_b200_neigh_build / annp_b200_compute_device / annp_b200_halo_* / annp_b200_nve_* : the device-resident mode
 *       (replaces the per-step H2D x / D2H f staging of lib/lal_annp.cpp:310-312,336-347 and the
 *        GPU_NEIGH path annp_gpu_compute_n, lib/lal_annp_ext.cpp:98-108)
 *
 * Conventions
 *   - plain C: pointers and sizes only.  All pointers are caller-owned; init copies what it needs.
 *   - handle based (the reference keeps one static singleton per process, lal_annp_ext.cpp:20).
 *   - return value: 0 on success, negative on failure.  -1..-8 keep the meaning of the reference's
 *     init codes (lib/lal_annp.h:28-33): -3 out of device memory, -4 library built without CUDA
 *     support / no usable device, -5 device lacks FP64.  Further codes below.  The message of the
 *     last failure on a handle is available from annp_b200_last_error().
 *   - there is NO CPU fallback: every compute entry point fails with ANNP_B200_ENODEVICE when no
 *     sm_100 device is present.
 *   - not re-entrant per handle; different handles may be used from different threads.
 */
#ifndef ANNP_B200_H
#define ANNP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ANNP_B200_ABI_VERSION 3

#define ANNP_B200_MAX_SF 64        /* descriptor components                      */
#define ANNP_B200_MAX_NOD 32       /* nodes per hidden layer                     */
#define ANNP_B200_MAX_LAYERS 6     /* weight layers (ntl - 1)                    */
#define ANNP_B200_MAX_ELEMENTS 4
#define ANNP_B200_MAX_NEIGH 384    /* in-cutoff neighbours of one atom (smem tile) */
/* Chebyshev descriptor shapes (Fe copies, ANNA-ADP): any 1 <= npsf <= 16, 1 <= ntsf <= 24 (the reference's CPU style reads any
 * `TL HL nodes nsf npsf ntsf` line, fe_v2/src/pair_annp.cpp:367-389).  (9,19), (8,20), (4,6) have exact kernel instantiations;
 * every other shape runs zero-padded on (8,24) or (16,24).  ntsf stops at 24 because the angular polynomial is evaluated in
 * monomial form, which is well conditioned up to degree 23.  All five activations of each copy, nelements <= 4. */
#define ANNP_B200_MAX_NPSF 16
#define ANNP_B200_MAX_NTSF 24

/* error codes */
#define ANNP_B200_OK 0
#define ANNP_B200_ENOFIX (-1)      /* kept for parity with the reference's -1    */
#define ANNP_B200_ENOMEM (-3)
#define ANNP_B200_ENODEVICE (-4)
#define ANNP_B200_ENOFP64 (-5)
#define ANNP_B200_ESPLIT (-8)
#define ANNP_B200_EINVAL (-20)     /* bad argument / unsupported parameter set   */
#define ANNP_B200_ECUDA (-21)      /* CUDA runtime error, see last_error         */
#define ANNP_B200_ESTATE (-22)     /* call order (compute before neigh, ...)     */
#define ANNP_B200_EOVERFLOW (-23)  /* an atom has more than MAX_NEIGH neighbours inside the cutoff */
#define ANNP_B200_EIO (-24)        /* potential file unreadable / malformed      */
#define ANNP_B200_ECOMM (-25)      /* NCCL unavailable or a NCCL call failed, see last_error */

typedef struct annp_b200_handle_s *annp_b200_handle;

/* descriptor families: flagsym of the reference (src/pair_annp.cpp:415-417) */
#define ANNP_B200_SYM_CHEBYSHEV 0

/* Which copy of the reference pair style the numbers are to follow.  The three `annp` copies share the
 * file format and style name and differ in the source the user installs; the potential file does not say
 * which one it is for (the Ni file still carries the keyword "Chebyshev"), so the host states it.
 *   VARIANT_FE  annp-gpu-lammps/fe, fe_v2: Chebyshev radial/angular descriptor, (G - avg)/sigma normalisation,
 *               E = e_scale*out + e_shift + e_atom, activation 4 = 1.7159 tanh(2x/3) + 0.1 x
 *   VARIANT_NI  annp-gpu-lammps/ni: Behler-Parrinello G2/G3(narrow) in Bohr (x1.889726), (G - min)/(max - min),
 *               E = raw network output, forces x51.422515, activations 3 and 4 = tanh
 *               (ni/src/pair_annp.cpp:74-212, 686-767, 786-807, 858-860) */
#define ANNP_B200_VARIANT_FE 0
#define ANNP_B200_VARIANT_NI 1
/*   VARIANT_ANNA_ADP  anna-gpu-lammps/bcc_fe (`pair_style anna_adp/gpu`): handles of this kind are created by
 *               anna_b200_init below; every other entry point (neigh, compute, halo, ...) is shared */
#define ANNP_B200_VARIANT_ANNA_ADP 2
/* or'ed into params.variant: run the generic Behler-Parrinello kernel even when the coefficient table has the product
 * structure the fast kernel is specialised for (used by the tests to keep both kernels covered) */
#define ANNP_B200_VARIANT_FLAG_GENERIC 0x100
/* or'ed into params.variant: use the lane-per-neighbour fast kernel even when the tile fits the pair-compaction kernel */
#define ANNP_B200_VARIANT_FLAG_NOPAIR 0x200

/*
 * Flat parameter block == the argument list of annp_gpu_init (src/pair_annp_gpu.cpp:31-39).
 *   sfnor_scal[n] = 1/sqrt(cov_n - avg_n^2), 0 if <= 1e-10      (src/pair_annp_gpu.cpp:211-220)
 *   cutsq  : (ntypes+1) x (ntypes+1), row-major, 1-based types   (host_cutsq)
 *   map    : ntypes+1 entries, LAMMPS type -> element index      (host_map)
 *   weights: per element, per layer l = 0..ntl-2, row-major [row * ncol + col] exactly as the
 *            reference flattens them (src/pair_annp_gpu.cpp:190-209); the blocks of one element are
 *            concatenated in layer order: nnod*nsf, (ntl-3) x nnod*nnod, nnod; elements follow
 *            each other.  bias likewise: (ntl-2) x nnod, then 1.
 */
typedef struct annp_b200_params {
  int abi_version;            /* ANNP_B200_ABI_VERSION */
  int ntypes;
  int nelements;
  int ntl, nhl, nnod, nsf, npsf, ntsf;
  int flagsym;
  int flagact[ANNP_B200_MAX_LAYERS];
  double e_scale, e_shift, e_atom;
  double cut;                 /* descriptor cutoff Rc of the potential file */
  const double *sfnor_scal;   /* [nsf] */
  const double *sfnor_avg;    /* [nsf] */
  const double *cutsq;        /* [(ntypes+1)^2] */
  const int *map;             /* [ntypes+1] */
  const double *weights;      /* see above */
  const double *bias;
  /* ---- ABI 2 ---- */
  int variant;                /* ANNP_B200_VARIANT_*                                                     */
  /* VARIANT_NI only (host_cofsymrad / host_cofsymang of ni/src/pair_annp_gpu.cpp:31-40):
   *   sym_coerad [npsf][3] = eta, rs (unused by the reference), Rc      sym_coeang [ntsf][4] = eta, lambda, zeta, Rc
   *   and sfnor_scal = 1/(sf_max - sf_min), sfnor_avg = sf_min */
  const double *sym_coerad;
  const double *sym_coeang;
} annp_b200_params;

/* number of doubles in params.weights / params.bias for one element */
size_t annp_b200_weights_per_element(int ntl, int nnod, int nsf);
size_t annp_b200_bias_per_element(int ntl, int nnod);

/* ---- potential file (host only, needs no GPU) --------------------------------------------- */

typedef struct annp_b200_potential {
  int nelements;
  int ntl, nhl, nnod, nsf, npsf, ntsf;
  int flagsym;
  int flagact[ANNP_B200_MAX_LAYERS];
  double cut, e_scale, e_shift, e_atom;
  int id_elem[ANNP_B200_MAX_ELEMENTS];
  double mass[ANNP_B200_MAX_ELEMENTS];
  char elements[ANNP_B200_MAX_ELEMENTS][16];
  double sfnor_cov[ANNP_B200_MAX_SF];
  double sfnor_avg[ANNP_B200_MAX_SF];
  /* weight_all[elem][layer][row][col] with rows padded to nsf columns, as PairANNP stores it
   * (src/pair_annp.cpp:441-447); bias_all[elem][layer][node].  malloc'd by the reader. */
  double *weight_all;   /* [nelements][ntl-1][nnod][nsf] */
  double *bias_all;     /* [nelements][ntl-1][nnod]      */
  /* ---- ABI 2: trailing "#coefficent of symmetry funciton" blocks of the Ni files (ni/src/pair_annp.cpp:510-545);
   * for those files sfnor_cov / sfnor_avg hold the sf_min / sf_max rows */
  int has_sym_coeff;
  double sym_coerad[ANNP_B200_MAX_SF][3];
  double sym_coeang[ANNP_B200_MAX_SF][4];
} annp_b200_potential;

/* Parse a `.ann` file the way PairANNP::read_file does (line-index addressing, tab-then-digit-or-
 * minus tokenisation, CRLF tolerated, "ta"->activation 4 ...).  elements_coeff are the element
 * names given on the pair_coeff line (used to attach `#<Elem>` weight blocks).  On success the
 * caller releases the arrays with annp_b200_free_potential. */
int annp_b200_read_potential(const char *filename, int nelements_coeff, const char *const *elements_coeff,
                             annp_b200_potential *out, char *err, int errlen);
void annp_b200_free_potential(annp_b200_potential *pot);

/* ---- life cycle ---------------------------------------------------------------------------- */

/* device < 0: use the current CUDA device.  nall_hint/max_nbors_hint size the first allocation
 * (they are grown on demand, as the reference does in lal_annp.cpp:562-580). */
int annp_b200_init(const annp_b200_params *params, int device, int nall_hint, int max_nbors_hint,
                   annp_b200_handle *out, char *err, int errlen);
void annp_b200_clear(annp_b200_handle h);
double annp_b200_bytes(annp_b200_handle h);            /* device bytes currently held */
const char *annp_b200_last_error(annp_b200_handle h);

/* ---- ANNA-ADP (anna-gpu-lammps/bcc_fe, `pair_style anna_adp/gpu`) ---------------------------------------------
 *
 *   anna_b200_read_potential <- PairANNA_ADP::read_file   src/pair_anna_adp.cpp:392-637 (`.anna` file format)
 *   anna_b200_init           <- anna_adp_gpu_init         src/pair_anna_adp_gpu.cpp:31-40, lib/lal_anna_adp_ext.cpp:25-95
 *   annp_b200_neigh + annp_b200_compute on the returned handle
 *                            <- anna_adp_gpu_compute followed by anna_adp_gpu_compute_force
 *                               (src/pair_anna_adp_gpu.cpp:42-63, 93-157).  The reference GPU style runs with newton
 *                               off and forward-communicates rho, mu[3], lambda[6], d2, q2 of every atom to the ghosts
 *                               between its two phases; this library follows the reference CPU style instead
 *                               (newton on, src/pair_anna_adp.cpp:73-286): everything is centred on the local atom and
 *                               ghost forces go home through the usual reverse communication, so no per-atom
 *                               intermediate ever leaves the device.
 * The numbers follow the CPU style: raw (unnormalised) Chebyshev descriptor -> network (activations 3, 4 =
 * 1.7 tanh(0.3 x)) -> (d2, q2) -> rho, mu, lambda sums with the smooth step psi -> energy and i-centred forces with
 * d2, q2 held fixed. */
#define ANNA_B200_MAX_GPARAMS 32
typedef struct anna_b200_potential {
  int nelements;
  int ntl, nhl, nnod, nout, nsf, npsf, ntsf;
  int flagsym;
  int flagact[ANNP_B200_MAX_LAYERS];
  double cut, e_base, e_scal;
  int ngp;
  double gparams[ANNA_B200_MAX_GPARAMS];  /* A0, yy, gamma, C0, c1F, c2F, V0, b1, b2, delta, r0, r1, hc, d1, q1, d3, q3 */
  int id_elem[ANNP_B200_MAX_ELEMENTS];
  double mass[ANNP_B200_MAX_ELEMENTS];
  char elements[ANNP_B200_MAX_ELEMENTS][16];
  double *weight_all;   /* [nelements][ntl-1][nnod][nsf], rows padded to nsf; last layer has nout rows */
  double *bias_all;     /* [nelements][ntl-1][nnod] */
} anna_b200_potential;
int anna_b200_read_potential(const char *filename, int nelements_coeff, const char *const *elements_coeff,
                             anna_b200_potential *out, char *err, int errlen);
void anna_b200_free_potential(anna_b200_potential *pot);

/* Flat parameter block == the argument list of anna_adp_gpu_init.  weights per element: layer 0 [nnod][nsf], hidden
 * [nnod][nnod], last [nout][nnod], row-major, concatenated; bias likewise (nnod ... nnod, nout). */
typedef struct anna_b200_params {
  int abi_version;
  int ntypes, nelements;
  int ntl, nhl, nnod, nout, nsf, npsf, ntsf, ngp;
  int flagsym;
  int flagact[ANNP_B200_MAX_LAYERS];
  double e_base, cut;
  const double *cutsq;        /* [(ntypes+1)*(ntypes+1)] */
  const int *map;             /* [ntypes+1] */
  const double *weights;
  const double *bias;
  const double *gparams;      /* [ngp], ngp >= 17 */
} anna_b200_params;
int anna_b200_init(const anna_b200_params *params, int device, annp_b200_handle *out, char *err, int errlen);

/* ---- LAMMPS host-list mode (gpu_mode == GPU_FORCE) ------------------------------------------ */

/* Upload a LAMMPS full neighbour list (call when neighbor->ago == 0).  numneigh and firstneigh are
 * indexed by atom index i = ilist[ii] as in LAMMPS' NeighList; entries are masked with NEIGHMASK. */
int annp_b200_neigh(annp_b200_handle h, int inum, int nall, const int *ilist, const int *numneigh,
                    const int *const *firstneigh);
/* Same list as flat CSR rows in ilist order (offsets has inum+1 entries). */
int annp_b200_neigh_csr(annp_b200_handle h, int inum, int nall, const int *ilist, const int64_t *offsets,
                        const int *neigh);

/* Device-neighbour mode of the host-driven style (the reference's GPU_NEIGH path annp_gpu_compute_n,
 * fe/lib/lal_annp.cpp:376-498: `package gpu ... neigh yes`): build the full list on the device from HOST positions
 * x [nall][3] (atoms 0..nlocal-1 are centres); bbox_lo / bbox_hi bound all nall positions (sub-domain +- ghost cutoff).
 * Replaces annp_b200_neigh at neighbor->ago == 0; annp_b200_compute follows as usual. */
int annp_b200_neigh_build_host(annp_b200_handle h, int nlocal, int nall, const double *x, const double *bbox_lo,
                               const double *bbox_hi, double cutneigh);

/* Page-locking of caller-owned host arrays.  annp_b200_compute copies x / type in and f / eatom / vatom out with
 * cudaMemcpyAsync: from pageable memory the driver stages every copy through its own bounce buffer (a few GB/s and a
 * synchronous hand-over); from page-locked memory it is one DMA at PCIe speed.  LAMMPS allocates atom->x / atom->f with
 * malloc, so the pair style registers them once per (re)allocation (atom->nmax growth) - host_register on a range that is
 * already registered is a no-op - and allocates its own staging arrays with host_alloc.  (The reference's LAL layer keeps
 * page-locked staging arrays of its own: UCL_H_Vec in lal_atom.h / lal_answer.h of LAMMPS' lib/gpu.) */
int annp_b200_host_register(void *ptr, size_t bytes);
int annp_b200_host_unregister(void *ptr);
void *annp_b200_host_alloc(size_t bytes);
void annp_b200_host_free(void *ptr);

/* One Pair::compute.  x is [nall][3], type [nall] (1-based); outputs may be NULL when not requested.
 *   type    may be NULL on the calls between two neighbour-list builds: the types uploaded by the last call that
 *           passed them are reused (atoms keep their slots until the next re-neighbouring; the reference uploads x and
 *           type together every step, lal_annp.cpp:310-312)
 *   f       [nall][3]  ASSIGNED (not accumulated), like the reference (lal_annp.cpp:336-347); ghost
 *                      rows carry the contributions LAMMPS reverse-communicates to the owners
 *   eng     sum of atomic energies of the inum centre atoms (eng_vdwl)
 *   eatom   [inum...]  eatom[ilist[ii]] = E_i      (assigned for centre atoms)
 *   virial6 per-pair tally sum (xi-xj) (x) (-Fj), LAMMPS order xx,yy,zz,xy,xz,yz; equals
 *           virial_fdotr_compute over local+ghost
 *   vatom   [nall][6]  half/half per-atom virial as ev_tally_xyz (assigned) */
int annp_b200_compute(annp_b200_handle h, int nlocal, int nghost, const double *x, const int *type,
                      int eflag, int vflag, double *f, double *eng, double *eatom, double *virial6,
                      double *vatom);

/* ---- device-resident mode ------------------------------------------------------------------- */

/* All pointers below are DEVICE pointers (cudaMalloc'd by the caller, e.g. torch tensors);
 * stream is a cudaStream_t passed as void* (NULL = default stream).  Nothing is synchronised:
 * results are valid after the stream reaches this point. */

/* Build the full neighbour list on the device from positions (cell binning).  Atoms 0..nlocal-1 are
 * centres; partners are all nall atoms within cutneigh.  bbox_lo/hi (host pointers, 3 doubles) bound
 * all nall positions.  Rows are sorted by neighbour index, so the result is deterministic. */
int annp_b200_neigh_build(annp_b200_handle h, int nlocal, int nall, const double *d_x, const double *bbox_lo,
                          const double *bbox_hi, double cutneigh, void *stream);

/* d_f [nall][3] assigned; d_eatom [nlocal] assigned or NULL; d_eng_virial: 7 doubles
 * (eng, v_xx, v_yy, v_zz, v_xy, v_xz, v_yz) assigned or NULL; d_vatom [nall][6] or NULL. */
int annp_b200_compute_device(annp_b200_handle h, int nlocal, int nghost, const double *d_x, const int *d_type,
                             int eflag, int vflag, double *d_f, double *d_eatom, double *d_eng_virial,
                             double *d_vatom, void *stream);

/* Halo ("ghost") exchange for the device-resident mode: LAMMPS' forward / reverse communication
 * (Comm::forward_comm / reverse_comm around Pair::compute) on device buffers.
 *   set_halo registers the send list: entry m sends local atom send_index[m] displaced by
 *   send_shift[m][3] (periodic image shift as seen by the receiver).  Entries are ordered by
 *   destination rank, so the packed buffer can be handed to one grouped NCCL send/recv
 *   (all_to_all_single); on one rank the "receiver" is this rank's own ghost block.  Device
 *   pointers, which must stay alive until the next set_halo.
 *     halo_pack        sendbuf[m] = x[send_index[m]] + send_shift[m]                    (forward)
 *     halo_unpack_add  f[send_index[m]] += recvbuf[m], summed per atom in ascending m   (reverse;
 *                      deterministic, no atomics).  recvbuf holds the ghost forces returned by
 *                      the receivers in send order. */
int annp_b200_set_halo(annp_b200_handle h, int nlocal, int nsend, const int *d_send_index,
                       const double *d_send_shift, void *stream);
int annp_b200_halo_pack(annp_b200_handle h, const double *d_x, double *d_sendbuf, void *stream);
int annp_b200_halo_unpack_add(annp_b200_handle h, const double *d_recvbuf, double *d_f, void *stream);

/* ---- ghost map on the device, halo exchange over NCCL ---------------------------------------------------------------------
 * What LAMMPS' Comm::borders / forward_comm / reverse_comm do on host arrays over MPI (and the reference's library stages
 * through the host around them, fe_v2/lib/lal_annp.cpp:310-312, 336-347), for runs that keep LAMMPS' brick decomposition
 * with one rank per GPU.
 *
 * Ghost map (once per re-neighbouring).  The caller describes up to 26 SLOTS: slot k takes every local atom that lies within
 * cutghost of the faces named by slot_dir[k] = (sx, sy, sz), s = +1: x >= hi - cutghost, -1: x < lo + cutghost, 0: any;
 * slots are listed in the order their atoms are to appear in the send list (by destination rank, then direction).
 *   send_lists_count  classifies the nlocal atoms at d_x (device) and returns the 26 slot counts - the only numbers of a
 *                     re-neighbouring that travel to the host; the caller sizes the send list from them
 *   send_lists_fill   writes send_index[m] / send_shift[m][3] (device) for all slots, entries of a slot in ascending atom
 *                     index: the arguments of annp_b200_set_halo.  slot_shift[k][3] is the periodic image shift of slot k.
 * Deterministic (two-level prefix sums, no atomics on the output order). */
int annp_b200_send_lists_count(annp_b200_handle h, int nlocal, const double *d_x, const double *lo, const double *hi,
                               double cutghost, int nslots, const int *slot_dir, int *slot_counts, void *stream);
int annp_b200_send_lists_fill(annp_b200_handle h, const double *slot_shift, int *d_send_index, double *d_send_shift,
                              void *stream);

/* NCCL communicator of the handle: one rank (of the communicator) calls comm_unique_id and hands the 128 bytes to every
 * rank by whatever it has (MPI_Bcast inside LAMMPS, torch.distributed in the stand-alone driver); every rank then calls
 * comm_init.  nranks == 1 needs no id and no NCCL.  NCCL itself is taken from the process at run time (libnccl.so.2;
 * ANNP_B200_NCCL_LIB names another file), so the library has no link-time dependency on it. */
int annp_b200_comm_unique_id(char *id128);
int annp_b200_comm_init(annp_b200_handle h, int nranks, int rank, const char *id128);
void annp_b200_comm_destroy(annp_b200_handle h);
/* per-rank atom counts of one exchange, after annp_b200_set_halo: send_counts[r] entries of the send list (which is ordered
 * by destination) go to rank r, recv_counts[r] ghosts arrive from rank r (the ghost block is ordered by source rank) */
int annp_b200_set_halo_peers(annp_b200_handle h, int nranks, const int *send_counts, const int *recv_counts);
/* Comm::forward_comm: ghost block of d_x [nall][3] <- owners' positions (+ periodic shift): pack kernel and ONE grouped
 * ncclSend / ncclRecv straight into the ghost rows (on one rank: the pack kernel writes them).
 * Comm::reverse_comm: ghost rows of d_f go home and are added to the owners' rows in a fixed order (halo_unpack_add).
 * Plain stream work, no host synchronisation: a whole MD step may be captured in a CUDA graph. */
int annp_b200_halo_forward(annp_b200_handle h, double *d_x, void *stream);
int annp_b200_halo_reverse(annp_b200_handle h, double *d_f, void *stream);
/* Peer scatter: the reverse halo FUSED into the force kernel (ranks of one node, fixed-point scatter mode).  Every rank
 * exports its force accumulators (peer_export: cudaIpc handle, 64 bytes), the caller gathers the handles of all ranks and
 * hands them to peer_open together with, for every ghost atom g of this rank, the rank that owns it (d_ghost_rank[g]) and
 * the owner's local index (d_ghost_index[g] = the owner's send_index entry; device arrays that stay alive until the next
 * peer_open).  From then on the force kernel adds the force on a ghost straight into its OWNER's accumulator with
 * system-scope 64-bit integer atomics over NVLink, annp_b200_halo_reverse becomes a no-op and annp_b200_compute_device
 * places one barrier (a one-element all-reduce on the handle's communicator) between the kernels and the conversion of the
 * accumulators.  Integer addition commutes: forces are bit-identical to the exchange path.  Call both after
 * annp_b200_set_halo_peers at every re-neighbouring (nall = nlocal + ghosts; the export re-allocates the accumulators if
 * they must grow); annp_b200_halo_forward must precede every annp_b200_compute_device (it clears the accumulators before
 * the peers can reach them). */
int annp_b200_peer_export(annp_b200_handle h, int nall, char *ipc64);
int annp_b200_peer_open(annp_b200_handle h, int nranks, const char *ipc64_all, int nghost, const int *d_ghost_rank,
                        const int *d_ghost_index, void *stream);
int annp_b200_peer_close(annp_b200_handle h);

/* sum of d_buf[0..n) over the ranks of the communicator, in place (thermo scalars, the 12 Nose-Hoover tensors) */
int annp_b200_allreduce_sum(annp_b200_handle h, double *d_buf, int n, void *stream);

/* d_out[0] = max over the nlocal atoms of |x - xref|^2 (device pointers): the quantity `neigh_modify check yes` compares
 * with (skin/2)^2 (Neighbor::check_distance); one fused kernel, exact and order independent */
int annp_b200_max_displacement_sq(annp_b200_handle h, int nlocal, const double *d_x, const double *d_xref, double *d_out,
                                  void *stream);

/* velocity-Verlet halves for the stand-alone MD loop (metal units; ftm2v = 1/(1.0364269e-4)):
 *   initial: v += dtf * f / m ; x += dt * v        final: v += dtf * f / m
 * d_ke (1 double, may be NULL) receives sum 0.5 m v^2 in mass*velocity^2 units on `final`. */
int annp_b200_nve_initial(annp_b200_handle h, int nlocal, double dt, double mass, double *d_x, double *d_v,
                          const double *d_f, void *stream);
int annp_b200_nve_final(annp_b200_handle h, int nlocal, double dt, double mass, double *d_v, const double *d_f,
                        double *d_ke, void *stream);

/* ---- Nose-Hoover chains on the device: `fix nvt` / `fix npt` of the reference's decks (in.st_test:30-37) ---------------
 * LAMMPS' FixNH operator splitting (MTK equations, tchain = pchain = 3, one sub-cycle), orthogonal box, independent
 * x / y / z barostat coupling, metal units.  The extended variables live in device memory and are advanced by
 * single-thread kernels, so a step never synchronises with the host.  One MD step of a rank:
 *     annp_b200_nh_initial     chain half step + per-atom: thermostat/barostat velocity scaling, kick, box dilation,
 *                              drift (positions and, for npt, the periodic image shifts of the halo send list)
 *     [halo, annp_b200_compute_device with vflag = 1, reverse halo]
 *     annp_b200_nh_final_kick  per-atom kick + barostat scaling; red12 = this rank's [m sum v(x)v (6), pair virial (6)]
 *     [sum red12 over ranks: ncclAllReduce / torch.distributed.all_reduce; nothing to do on one rank]
 *     annp_b200_nh_final_scale chain half step from the summed tensors + per-atom thermostat scaling
 * annp_b200_nh_reduce + annp_b200_nh_setup initialise temperature, pressure and chain masses before the first step
 * (FixNH::setup).  All d_* pointers are device pointers; stream is a cudaStream_t. */
typedef struct annp_b200_nh_s *annp_b200_nh;
typedef struct annp_b200_nh_config {
  int tstat, pstat;              /* fix nvt: 1, 0     fix npt: 1, 1                                        */
  double t_start, t_stop, t_damp;/* K, K, ps                                                               */
  int p_flag[3];                 /* coupled dimensions (the deck: y only)                                  */
  double p_start[3], p_stop[3], p_damp[3];   /* bar, bar, ps                                               */
  int tchain, pchain, mtk;       /* LAMMPS defaults 3, 3, 1                                                */
  double dt, mass;               /* ps, g/mol (single species)                                             */
  double natoms_total;           /* atoms of the whole system (all ranks)                                  */
  double tdof;                   /* temperature degrees of freedom; <= 0: 3 natoms - 3                     */
  long long nsteps_ramp;         /* run length over which start -> stop ramps (0: constant targets)        */
} annp_b200_nh_config;
typedef struct annp_b200_nh_state {
  long long step;
  double t_current, t_target;
  double p_current[3];           /* bar */
  double boxlo[3], boxhi[3];
  double omega_dot[3];
  double ke_tensor[6], virial[6];/* eV */
  double eta[8], eta_dot[8], etap[8], etap_dot[8];
  double extended_energy;        /* FixNH::compute_scalar: PE + KE + this is the conserved quantity (eV)   */
} annp_b200_nh_state;
int annp_b200_nh_create(const annp_b200_nh_config *cfg, const double *boxlo, const double *boxhi, int device,
                        annp_b200_nh *out, char *err, int errlen);
void annp_b200_nh_destroy(annp_b200_nh nh);
int annp_b200_nh_reduce(annp_b200_nh nh, int nlocal, const double *d_v, const double *d_eng_virial, double *d_red12, void *stream);
int annp_b200_nh_setup(annp_b200_nh nh, const double *d_red12, void *stream);
int annp_b200_nh_initial(annp_b200_nh nh, int nlocal, double *d_x, double *d_v, const double *d_f, int nsend,
                         double *d_send_shift, void *stream);
int annp_b200_nh_final_kick(annp_b200_nh nh, int nlocal, double *d_v, const double *d_f, const double *d_eng_virial,
                            double *d_red12, void *stream);
int annp_b200_nh_final_scale(annp_b200_nh nh, int nlocal, double *d_v, const double *d_red12, void *stream);
/* new edges for dimensions the barostat does not couple (which[d] != 0): LAMMPS' shrink-wrapped `boundary m / s` faces
 * are reset to the atoms' extent at every re-neighbouring and enter the pressure through the volume */
int annp_b200_nh_set_box(annp_b200_nh nh, const double *lo, const double *hi, const int *which, void *stream);
int annp_b200_nh_get_state(annp_b200_nh nh, annp_b200_nh_state *out, void *stream);   /* synchronises the stream */

/* measured FP64 FMA throughput of this device in TFLOP/s (pure DFMA loop, best of reps): the
 * denominator of the force kernel's roofline (MEASURED_PEAKS.json has no FP64 entry) */
double annp_b200_fp64_peak_tflops(annp_b200_handle h, int reps);

/* ---- introspection (tests, bench) ----------------------------------------------------------- */

typedef struct annp_b200_stats {
  int inum, nall;
  int max_neigh_list;        /* longest list row                           */
  int max_neigh_cut;         /* most in-cutoff neighbours of any atom (last compute) */
  double avg_neigh_cut;      /* mean in-cutoff neighbours (last compute)   */
  double sum_triplets;       /* sum over atoms of N(N-1)/2 (last compute)  */
  long long kernel_launches; /* kernels launched by this handle so far     */
  float last_force_kernel_ms;/* CUDA-event time of the descriptor+force kernel of the last
                                compute call issued with timing enabled, else 0 */
  double force_kernel_ms_total; /* sum of the CUDA-event times of the force kernel over the (up to 256)  */
  int force_kernel_samples;     /* ... timed launches since annp_b200_set_timing(h, 1)                    */
  /* ---- ABI 3 ---- */
  long long overflow_pass_atoms;/* atoms redone by the overflow pass (in-cutoff neighbours outgrew the first-pass shared-memory
                                   tile between two list builds) since the previous get_stats / host-mode compute          */
  int tile_capacity;            /* neighbour slots of the first-pass tile                                                   */
  double stage_cycles[8];       /* libraries built with -DANNP_STAGE_CLOCKS: SM cycles summed over warps since the previous
                                   call, per stage: 0 filter, 1 radial, 2 forward angular, 3 reduction + MLP, 4 backward
                                   angular, 5 force assembly + scatter, 6 scheduler; zeros otherwise                        */
} annp_b200_stats;
/* Synchronises the device.  Also the error check of the device-resident mode: returns ANNP_B200_EOVERFLOW if, in any
 * annp_b200_compute_device since the previous call, an atom had more than ANNP_B200_MAX_NEIGH in-cutoff neighbours or a
 * pair force left the fixed-point range (the flags are sticky on the device until read here).  Atoms that merely outgrow
 * the first-pass tile are NOT an error: the overflow pass redoes them inside the same step. */
int annp_b200_get_stats(annp_b200_handle h, annp_b200_stats *out);
int annp_b200_set_timing(annp_b200_handle h, int enabled);

/* How the forces on the NEIGHBOURS of a centre atom (f[j] += Fj, fe_v2/src/pair_annp.cpp:198-200) reach f.  Both are
 * deterministic (no floating-point atomics):
 *   ANNP_B200_SCATTER_FIXED   every pair force is rounded to a multiple of 2^-43 eV/A (1.1e-13) and added to a 64-bit
 *                             integer accumulator per atom with integer atomics - exact and order independent, no
 *                             per-list-entry buffer in HBM.  Default.  A pair force that is NaN or >= 2^18 eV/A makes the
 *                             compute call fail with ANNP_B200_EOVERFLOW.
 *   ANNP_B200_SCATTER_GATHER  pair forces are stored per list entry (32 B each) and summed per atom in list order by a
 *                             gather over a reverse map (plain FP64 sums).                                              */
#define ANNP_B200_SCATTER_GATHER 0
#define ANNP_B200_SCATTER_FIXED 1
int annp_b200_set_scatter(annp_b200_handle h, int mode);

/* centred descriptors G [inum][nsf] and dE/dG [inum][nsf] of the last compute (device->host copy;
 * test hook replacing the reference's printf debugging). Either pointer may be NULL. */
int annp_b200_debug_descriptors(annp_b200_handle h, double *G, double *dE_dG);

/* The neighbour list the handle currently holds (uploaded or built on the device), as CSR rows in ilist order: offsets
 * [inum+1] and neigh [total], either may be NULL.  Returns the number of entries (>= 0) or a negative error code; call with
 * NULLs first to size neigh.  Test / bench hook (device->host copy, synchronises the device). */
long long annp_b200_debug_neighbors(annp_b200_handle h, int64_t *offsets, int *neigh);

/* Host arithmetic, no device needed: the two basis-conversion matrices of the angular passes, each [ntsf][ntsf] row-major
 * (ntsf <= 24).  With y = (z+1)/2, z = cos(theta) and psi_{4b+i}(z) = T_{4b}(z) z^i:
 *     T_n(y) = sum_k cheb2mono[k*ntsf + n] z^k = sum_j blk2cheb[j*ntsf + n] psi_j(z)
 * (the reference evaluates T_n(y) by recurrence, fe_v2/src/pair_annp.cpp:596-611,658-695).  Exposed for tests. */
int annp_b200_basis_matrices(int ntsf, double *cheb2mono, double *blk2cheb);

int annp_b200_abi_version(void);
/* number of CUDA devices visible to the library (0 = none, no error) */
int annp_b200_device_count(void);

#ifdef __cplusplus
}
#endif
#endif /* ANNP_B200_H */

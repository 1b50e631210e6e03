// annp_capi.cu -- the C ABI of libannp_b200.so (include/annp_b200.h): handle, device buffers and the
// per-step launch sequence.  No CPU fallback: without an sm_100 device every entry point that needs
// the GPU returns ANNP_B200_ENODEVICE.
#include "annp_handle.cuh"

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

// ---- kernels' host entry points (annp_force.cu / annp_aux.cu / annp_neigh.cu)
size_t annp_force_smem_bytes(const DevParams &hp, int capacity);
bool annp_force_kernel_shape(int npsf, int ntsf, int variant, int *npsf_k, int *ntsf_k);
cudaError_t annp_force_launch(const ForceArgs &args, const DevParams &hp, int num_sms, cudaStream_t stream, int max_blocks);
cudaError_t annp_bp_force_launch(const ForceArgs &args, const DevParams &hp, int num_sms, cudaStream_t stream, int max_blocks);
int annp_bp_layout(const DevParams &hp);
void aux_pack_xq(const double *x, const int *type, double4 *xq, int nall, cudaStream_t s);
void aux_build_reverse(const int *nbr, long long total, int nall, long long *rev_off, int *rev_pos, int *cnt, int *tmp,
                       long long *tile_sum, cudaStream_t s);
void aux_centre_of(const int *ilist, int inum, int nall, int *centre_of, cudaStream_t s);
void aux_gather_force(const double4 *fpair, const double4 *fself, const int *centre_of, const long long *rev_off,
                      const int *rev_pos, double *f, int nall, cudaStream_t s);
void aux_finish_force(const long long *facc, const double4 *fself, const int *centre_of, double *f, int nall, cudaStream_t s);
void aux_gather_vatom(const double *vpair, const int *centre_of, const long long *row_off, const long long *rev_off,
                      const int *rev_pos, double *vatom, int nall, cudaStream_t s);
void aux_scatter_eatom(const double4 *fself, const int *ilist, int inum, double *eatom, cudaStream_t s);
int aux_reduce_blocks(int inum);
void aux_reduce_ev(const double4 *fself, const double *vir_c, int inum, double *partial, double *out7, cudaStream_t s);
void aux_halo_pack(int nsend, const int *idx, const double *shift, const double *x, double *out, cudaStream_t s);
void aux_build_ghost_csr(const int *owner, int nghost, int nlocal, long long *goff, int *glist, int *cnt, int *tmp,
                         long long *tile_sum, cudaStream_t s);
void aux_halo_unpack_add(int nlocal, const long long *goff, const int *glist, const double *src, double *f, cudaStream_t s);
void aux_nve_initial(int n, double dt, double dtfm, double *x, double *v, const double *f, cudaStream_t s);
void aux_max_disp2(int n, const double *x, const double *xref, double *out, cudaStream_t s);
void aux_iota(int *p, int n, cudaStream_t s);
extern "C" int annp_peer_barrier(annp_b200_handle h, cudaStream_t s);
void aux_nve_final(int n, double dtfm, double *v, const double *f, cudaStream_t s);
int aux_ke_blocks(int n);
void aux_kinetic(int n, const double *v, double half_mass, double *partial, double *out, cudaStream_t s);
double aux_fp64_peak_tflops(int num_sms, int reps, cudaStream_t s);

struct NeighScratch {
  int *cell_of, *cell_cnt, *cell_atoms, *row_cnt;
  long long *cell_off, *tile_sum;
};
void neigh_grid(const double *lo, const double *hi, double cutneigh, int *n_out, double *inv_out, long long *ncells);
void neigh_count(const double *d_x, int nlocal, int nall, const double *lo, const int *n, const double *inv, double cutneigh,
                 NeighScratch sc, long long *d_row_off, int *d_maxrow, cudaStream_t s);
void neigh_fill(const double *d_x, int nlocal, const double *lo, const int *n, const double *inv, double cutneigh,
                NeighScratch sc, const long long *d_row_off, int *d_rows_tmp, int *d_rows, cudaStream_t s);

namespace {

__global__ void k_count_cut(const DevParams *prm, const double4 *__restrict__ xq, const int *__restrict__ ilist,
                            const long long *__restrict__ row_off, const int *__restrict__ nbr, int inum, int *__restrict__ maxn) {
  const int lane = threadIdx.x & 31;
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nw = (gridDim.x * blockDim.x) >> 5;
  const int nt1 = prm->ntypes + 1;
  int best = 0;
  for (int ii = wid; ii < inum; ii += nw) {
    const double4 xi = xq[ilist[ii]];
    const int ti = (int) xi.w;
    int n = 0;
    for (long long p = row_off[ii] + lane; p < row_off[ii + 1]; p += 32) {
      const double4 xj = xq[nbr[p] & ANNP_NEIGHMASK];
      const double dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z;
      const double rsq = dx * dx + dy * dy + dz * dz;
      if (prm->variant == ANNP_B200_VARIANT_NI) n += (sqrt(rsq) * 1.889726 < fmax(prm->rad_rc, prm->ang_rc));   // filter of annp_bp_force_kernel
      else n += !(rsq > prm->cutsq[ti * nt1 + (int) xj.w] || rsq < 1.0e-12);
    }
    for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
    best = max(best, n);
  }
  if (lane == 0) atomicMax(maxn, best);
}

// a caller-supplied list must only name atoms 0..nall-1: bad[0] counts ilist entries and masked neighbour indices outside
// that range, bad[1] rows whose offsets decrease (the force kernels index xq / facc / centre_of with these numbers)
__global__ void k_validate_list(const int *__restrict__ ilist, const long long *__restrict__ row_off, const int *__restrict__ nbr,
                                int inum, long long total, int nall, int *__restrict__ bad) {
  const long long stride = (long long) gridDim.x * blockDim.x;
  int nb = 0, nr = 0;
  for (long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const int j = nbr[t] & ANNP_NEIGHMASK;
    nb += (j >= nall);
  }
  for (long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x; t < inum; t += stride) {
    const int i = ilist[t];
    nb += (i < 0 || i >= nall);
    nr += (row_off[t + 1] < row_off[t]);
  }
  if (nb) atomicAdd(bad, nb);
  if (nr) atomicAdd(bad + 1, nr);
}

}    // namespace

namespace {

int round_capacity(int n) {
  int c = ((n + 15) / 16) * 16;
  if (c < 32) c = 32;
  return c;
}

int finish_list(annp_b200_handle h, cudaStream_t s) {
  // centre index for the list now in d_ilist / d_row_off / d_nbr; the reverse map is built on first use (ensure_reverse)
  const int nall = h->nall_list;
  if (h->total >= (1LL << 31)) return fail(h, ANNP_B200_EINVAL, "neighbour list has 2^31 or more entries");
  CK(h->d_centre_of.reserve(sizeof(int) * (size_t) std::max(nall, 1)));
  aux_centre_of(h->d_ilist.as<int>(), h->inum, nall, h->d_centre_of.as<int>(), s);
  h->launches += 1;
  CK(cudaGetLastError());
  h->have_list = true;
  h->have_reverse = false;
  h->need_calibrate = true;
  return ANNP_B200_OK;
}

// reverse map (for every atom the list positions that name it, ascending): needed by the ordered gather and by vatom
int ensure_reverse(annp_b200_handle h, cudaStream_t s) {
  if (h->have_reverse) return ANNP_B200_OK;
  const int nall = h->nall_list;
  CK(h->d_rev_off.reserve(sizeof(long long) * ((size_t) nall + 1)));
  CK(h->d_rev_pos.reserve(sizeof(int) * (size_t) std::max<long long>(h->total, 1)));
  CK(h->d_scratch_cnt.reserve(sizeof(int) * (size_t) std::max(nall, 1)));
  CK(h->d_scratch_tmp.reserve(sizeof(int) * (size_t) std::max<long long>(h->total, 1)));
  CK(h->d_tile_sum.reserve(sizeof(long long) * ((size_t) nall / 1024 + 2)));
  aux_build_reverse(h->d_nbr.as<int>(), h->total, nall, h->d_rev_off.as<long long>(), h->d_rev_pos.as<int>(),
                    h->d_scratch_cnt.as<int>(), h->d_scratch_tmp.as<int>(), h->d_tile_sum.as<long long>(), s);
  h->launches += 8;
  CK(cudaGetLastError());
  h->have_reverse = true;
  return ANNP_B200_OK;
}

// the per-step device sequence; everything asynchronous on `s`
int step_device(annp_b200_handle h, int nlocal, int nghost, const double *d_x, const int *d_type, int eflag, int vflag,
                double *d_f, double *d_eatom, double *d_engvir, double *d_vatom, cudaStream_t s, bool allow_sync) {
  const int nall = nlocal + nghost;
  if (!h->have_list) return fail(h, ANNP_B200_ESTATE, "compute called before a neighbour list was provided");
  if (nall != h->nall_list) return fail(h, ANNP_B200_ESTATE, "nall differs from the neighbour list's nall");
  const bool want_vir = (vflag != 0) && d_engvir != nullptr;
  const bool want_vatom = d_vatom != nullptr;
  const int inum = h->inum;

  CK(h->d_xq.reserve(sizeof(double4) * (size_t) std::max(nall, 1)));
  const bool fixed = h->scatter_fixed != 0;
  if (fixed) CK(h->d_facc.reserve(sizeof(long long) * 3 * (size_t) std::max(nall, 1)));
  else CK(h->d_fpair.reserve(sizeof(double4) * (size_t) std::max<long long>(h->total, 1)));
  if (!fixed || want_vatom) { int rc = ensure_reverse(h, s); if (rc) return rc; }
  CK(h->d_fself.reserve(sizeof(double4) * (size_t) std::max(inum, 1)));
  CK(h->d_ovf_list.reserve(sizeof(int) * (size_t) std::max(inum, 1)));
  if (!h->d_counters.p) {
    CK(h->d_counters.reserve(sizeof(DevCounters)));
    CK(cudaMemsetAsync(h->d_counters.p, 0, sizeof(DevCounters), s));     // the sticky flags start clear
  }
  CK(h->d_partial.reserve(sizeof(double) * 7 * (size_t) aux_reduce_blocks(inum)));
  if (want_vir || want_vatom) CK(h->d_vir_c.reserve(sizeof(double) * 6 * (size_t) std::max(inum, 1)));
  if (want_vatom) CK(h->d_vpair.reserve(sizeof(double) * 6 * (size_t) std::max<long long>(h->total, 1)));
  if (h->debug_desc) {
    CK(h->d_Gdbg.reserve(sizeof(double) * (size_t) h->hp.nsf * std::max(inum, 1)));
    CK(h->d_dEdbg.reserve(sizeof(double) * (size_t) h->hp.nsf * std::max(inum, 1)));
  }

  aux_pack_xq(d_x, d_type, h->d_xq.as<double4>(), nall, s);
  h->launches += 1;

  if (h->need_calibrate) {
    // size the shared-memory neighbour tile from the actual in-cutoff maximum (once per list)
    if (allow_sync) {
      CK(h->d_small.reserve(64));
      CK(cudaMemsetAsync(h->d_small.p, 0, sizeof(int), s));
      if (inum > 0) {
        k_count_cut<<<std::min(148 * 16, (inum + 7) / 8), 256, 0, s>>>((const DevParams *) h->d_params.p, h->d_xq.as<double4>(), h->d_ilist.as<int>(),
                                                                     h->d_row_off.as<long long>(), h->d_nbr.as<int>(), inum, h->d_small.as<int>());
        h->launches += 1;
      }
      int maxn = 0;
      CK(cudaMemcpyAsync(&maxn, h->d_small.p, sizeof(int), cudaMemcpyDeviceToHost, s));
      CK(cudaStreamSynchronize(s));
      if (maxn > ANNP_B200_MAX_NEIGH) return fail(h, ANNP_B200_EOVERFLOW, "an atom has more in-cutoff neighbours than ANNP_B200_MAX_NEIGH");
      h->capacity = std::max(h->capacity, round_capacity(std::min(maxn + 12, ANNP_B200_MAX_NEIGH)));
    } else {
      h->capacity = std::max(h->capacity, round_capacity(std::min(h->max_row, ANNP_B200_MAX_NEIGH)));
    }
    h->need_calibrate = false;
  }

  // reset the schedulers and the per-step statistics; the sticky flags (overflow, bad_force) stay until they are reported
  CK(cudaMemsetAsync(h->d_counters.p, 0, offsetof(DevCounters, overflow), s));

  ForceArgs a;
  a.prm = (const DevParams *) h->d_params.p;
  a.xq = h->d_xq.as<double4>();
  a.ilist = h->d_ilist.as<int>();
  a.row_off = h->d_row_off.as<long long>();
  a.nbr = h->d_nbr.as<int>();
  a.fpair = fixed ? nullptr : h->d_fpair.as<double4>();
  a.facc = fixed ? h->d_facc.as<long long>() : nullptr;
  if (fixed && !(h->peer_on && h->peer_zeroed)) CK(cudaMemsetAsync(h->d_facc.p, 0, sizeof(long long) * 3 * (size_t) std::max(nall, 1), s));
  h->peer_zeroed = false;
  a.fself = h->d_fself.as<double4>();
  a.vir_c = (want_vir || want_vatom) ? h->d_vir_c.as<double>() : nullptr;
  a.vpair = want_vatom ? h->d_vpair.as<double>() : nullptr;
  a.G_dbg = h->debug_desc ? h->d_Gdbg.as<double>() : nullptr;
  a.dEdG_dbg = h->debug_desc ? h->d_dEdbg.as<double>() : nullptr;
  a.cnt = h->d_counters.as<DevCounters>();
  a.inum = inum;
  a.capacity = h->capacity;
  for (int k = 0; k < 17; k++) a.gp[k] = h->hp.gparams[k];
  // peer scatter: ghosts' forces go to their owners' accumulators (this rank's were zeroed before its forward exchange)
  const bool peer = h->peer_on && fixed && nghost == h->g_nghost_recv;
  a.peer_facc = peer ? h->d_peer_table.as<long long *>() : nullptr;
  a.ghost_rank = peer ? h->peer_ghost_rank : nullptr;
  a.ghost_index = peer ? h->peer_ghost_index : nullptr;
  a.peer_nlocal = nlocal;
  a.work_list = nullptr;
  a.work_count = nullptr;
  a.work_ctr = &a.cnt->work;
  a.ovf_list = h->d_ovf_list.as<int>();
  // Overflow pass: atoms whose in-cutoff neighbours outgrew the first-pass tile (they moved inside the skin since the tile
  // was sized) are redone with a tile of the list's longest row - a list row cannot overflow that.  The pass is a second
  // launch of the same kernel over the device-side list the first pass wrote (one block per SM; it exits at once when the
  // list is empty), so a device-resident step is complete without a host round trip.
  const int cap_full = round_capacity(std::min(std::max(h->max_row, 1), ANNP_B200_MAX_NEIGH));
  const bool need_pass2 = cap_full > h->capacity;

  if (inum > 0) {
    const bool timed = h->timing && h->ev_count < annp_b200_handle_s::kEvRing;
    if (timed) { CK(cudaEventRecord(h->ev0[h->ev_count], s)); }
    const bool ni = h->hp.variant == ANNP_B200_VARIANT_NI;
    cudaError_t e = ni ? annp_bp_force_launch(a, h->hp, h->num_sms, s, 0) : annp_force_launch(a, h->hp, h->num_sms, s, 0);
    if (e != cudaSuccess) return cuda_fail(h, e, "annp_force_launch");
    if (timed) { CK(cudaEventRecord(h->ev1[h->ev_count], s)); h->ev_count++; }
    h->launches += 1;
    if (need_pass2) {
      ForceArgs b = a;
      b.capacity = cap_full;
      b.work_list = a.ovf_list;
      b.work_count = &a.cnt->ovf_count;
      b.work_ctr = &a.cnt->work2;
      b.ovf_list = nullptr;                     // beyond the largest tile: sticky overflow flag, reported as an error
      e = ni ? annp_bp_force_launch(b, h->hp, h->num_sms, s, h->num_sms) : annp_force_launch(b, h->hp, h->num_sms, s, h->num_sms);
      if (e != cudaSuccess) return cuda_fail(h, e, "annp_force_launch (overflow pass)");
      h->launches += 1;
    }
  }
  if (peer) { int rc = annp_peer_barrier(h, s); if (rc) return rc; }      // every rank's kernel has added its ghost forces
  if (d_f && fixed) {
    aux_finish_force(h->d_facc.as<long long>(), h->d_fself.as<double4>(), h->d_centre_of.as<int>(), d_f, nall, s);
    h->launches += 1;
  } else if (d_f) {
    aux_gather_force(h->d_fpair.as<double4>(), h->d_fself.as<double4>(), h->d_centre_of.as<int>(), h->d_rev_off.as<long long>(),
                     h->d_rev_pos.as<int>(), d_f, nall, s);
    h->launches += 1;
  }
  if (d_engvir && (eflag || vflag)) {
    aux_reduce_ev(h->d_fself.as<double4>(), want_vir ? h->d_vir_c.as<double>() : nullptr, inum, h->d_partial.as<double>(), d_engvir, s);
    h->launches += 2;
  }
  if (d_eatom) { aux_scatter_eatom(h->d_fself.as<double4>(), h->d_ilist.as<int>(), inum, d_eatom, s); h->launches += 1; }
  if (want_vatom) {
    aux_gather_vatom(h->d_vpair.as<double>(), h->d_centre_of.as<int>(), h->d_row_off.as<long long>(), h->d_rev_off.as<long long>(),
                     h->d_rev_pos.as<int>(), d_vatom, nall, s);
    h->launches += 1;
  }
  CK(cudaGetLastError());
  return ANNP_B200_OK;
}

int fetch_counters(annp_b200_handle h, cudaStream_t s) {
  CK(cudaMemcpyAsync(&h->last_cnt, h->d_counters.p, sizeof(DevCounters), cudaMemcpyDeviceToHost, s));
  // the sticky part is handed to the caller (who reports it) and starts clear again
  CK(cudaMemsetAsync((char *) h->d_counters.p + offsetof(DevCounters, overflow), 0, sizeof(DevCounters) - offsetof(DevCounters, overflow), s));
  CK(cudaStreamSynchronize(s));
  // atoms took the overflow pass: size the first-pass tile for them from now on
  if (h->last_cnt.ovf_total > 0 && h->last_cnt.max_neigh + (h->last_cnt.max_neigh & 1) > h->capacity)
    h->capacity = round_capacity(std::min(h->last_cnt.max_neigh + 12, ANNP_B200_MAX_NEIGH));
  for (int k = 0; k < h->ev_count; k++) {
    float ms = 0.f;
    cudaEventSynchronize(h->ev1[k]);
    if (cudaEventElapsedTime(&ms, h->ev0[k], h->ev1[k]) == cudaSuccess) {
      h->last_force_ms = ms;
      h->force_ms_total += ms;
      h->force_samples++;
    }
  }
  h->ev_count = 0;
  return ANNP_B200_OK;
}

}    // namespace

// ================================================================================================
extern "C" {

int annp_b200_abi_version(void) { return ANNP_B200_ABI_VERSION; }

int annp_b200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

// device checks + handle with stream and event ring (shared by annp_b200_init / anna_b200_init)
static int open_handle(int device, annp_b200_handle *out, char *err, int errlen) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    set_err(err, errlen, "no CUDA device: libannp_b200 has no CPU fallback");
    return ANNP_B200_ENODEVICE;
  }
  if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) device = 0; }
  if (device >= ndev) { set_err(err, errlen, "device index out of range"); return ANNP_B200_ENODEVICE; }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { set_err(err, errlen, "cannot query device"); return ANNP_B200_ENODEVICE; }
  if (prop.major < 10) { set_err(err, errlen, "device is not sm_100 class; this library is built for sm_100a only"); return ANNP_B200_ENODEVICE; }
  if (cudaSetDevice(device) != cudaSuccess) { set_err(err, errlen, "cudaSetDevice failed"); return ANNP_B200_ENODEVICE; }
  annp_b200_handle h = new annp_b200_handle_s();
  h->device = device;
  h->num_sms = prop.multiProcessorCount;
  memset(&h->hp, 0, sizeof(h->hp));
  memset(&h->last_cnt, 0, sizeof(h->last_cnt));
  cudaError_t e;
  if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) {
    set_err(err, errlen, std::string("cudaStreamCreate: ") + cudaGetErrorString(e));
    annp_b200_clear(h);
    return ANNP_B200_ECUDA;
  }
  for (int k = 0; k < annp_b200_handle_s::kEvRing; k++) {
    if ((e = cudaEventCreate(&h->ev0[k])) != cudaSuccess || (e = cudaEventCreate(&h->ev1[k])) != cudaSuccess) {
      set_err(err, errlen, std::string("cudaEventCreate: ") + cudaGetErrorString(e));
      annp_b200_clear(h);
      return ANNP_B200_ECUDA;
    }
  }
  *out = h;
  return ANNP_B200_OK;
}

// network shape, type map, cutoffs -> hp; nout = rows of the last layer
static int fill_network(annp_b200_handle h, int ntypes, int nelements, int ntl, int nnod, int nout, int nsf, int npsf, int ntsf,
                        const int *flagact, const double *cutsq, const int *map, double cut, char *err, int errlen) {
  DevParams &hp = h->hp;
  const int nl = ntl - 1;
  hp.ntypes = ntypes; hp.nelements = nelements; hp.nlayers = nl; hp.nnod = nnod; hp.nout = nout;
  hp.nsf = nsf; hp.npsf = npsf; hp.ntsf = ntsf;
  for (int l = 0; l < nl; l++) hp.flagact[l] = flagact[l];
  const int nt1 = ntypes + 1;
  for (int t = 0; t < nt1; t++) hp.map[t] = (t == 0) ? 0 : map[t];
  for (int t = 1; t < nt1; t++)
    if (hp.map[t] < 0 || hp.map[t] >= nelements) { set_err(err, errlen, "type->element map entry out of range"); return ANNP_B200_EINVAL; }
  for (int a = 0; a < nt1 * nt1; a++) {
    hp.cutsq[a] = cutsq[a];
    hp.rcinv[a] = cutsq[a] > 0.0 ? 1.0 / sqrt(cutsq[a]) : 0.0;
  }
  hp.cut = cut; hp.two_over_cut = 2.0 / cut;
  int wo = 0, bo = 0;
  for (int l = 0; l < nl; l++) {
    hp.w_off[l] = wo; hp.b_off[l] = bo;
    const int nr = (l == nl - 1) ? nout : nnod, nc = (l == 0) ? nsf : nnod;
    wo += nr * nc; bo += nr;
  }
  hp.w_per_elem = wo; hp.b_per_elem = bo;
  return ANNP_B200_OK;
}

static bool valid_shape(int ntypes, int nelements, int ntl, int nnod, int nsf, int npsf, int ntsf) {
  const int nl = ntl - 1;
  return !(ntypes < 1 || ntypes > ANNP_MAX_TYPES || nelements < 1 || nelements > ANNP_B200_MAX_ELEMENTS || nl < 1 ||
           nl > ANNP_B200_MAX_LAYERS || nnod < 1 || nnod > ANNP_B200_MAX_NOD || nsf != npsf + ntsf || nsf > ANNP_B200_MAX_SF ||
           npsf < 1 || ntsf < 1);
}

// The kernel shape (npsf_k, ntsf_k) may be larger than the potential's (npsf, ntsf): descriptor component n < npsf keeps
// its place, angular component t moves to npsf_k + t, and the added components get zero scale / zero first-layer weights,
// so they contribute exact zeros to the energy and to every derivative.
static int padded_index(int n, int npsf, int npsf_k) { return n < npsf ? n : npsf_k + (n - npsf); }

static std::vector<double> pad_weights(const double *w, int nelements, int nl, int nnod, int nout, int nsf, int npsf, int npsf_k, int nsf_k) {
  std::vector<double> out;
  const double *src = w;
  for (int e = 0; e < nelements; e++)
    for (int l = 0; l < nl; l++) {
      const int nr = (l == nl - 1) ? nout : nnod, nc = (l == 0) ? nsf : nnod, nck = (l == 0) ? nsf_k : nnod;
      const size_t base = out.size();
      out.resize(base + (size_t) nr * nck, 0.0);
      for (int r = 0; r < nr; r++)
        for (int c = 0; c < nc; c++) out[base + (size_t) r * nck + (l == 0 ? padded_index(c, npsf, npsf_k) : c)] = src[(size_t) r * nc + c];
      src += (size_t) nr * nc;
    }
  return out;
}

// weights, biases, (Chebyshev variants) the monomial conversion matrix, then the parameter block itself
static int upload_params(annp_b200_handle h, const double *weights, const double *bias, bool chebyshev, char *err, int errlen) {
  DevParams &hp = h->hp;
  auto bail = [&](cudaError_t e, const char *w) {
    int rc = cuda_fail(h, e, w);
    set_err(err, errlen, h->err);
    return rc;
  };
  cudaError_t e;
  const size_t wbytes = sizeof(double) * (size_t) hp.w_per_elem * hp.nelements, bbytes = sizeof(double) * (size_t) hp.b_per_elem * hp.nelements;
  if ((e = h->d_weights.reserve(wbytes, 1.0)) != cudaSuccess) return bail(e, "cudaMalloc weights");
  if ((e = h->d_bias.reserve(bbytes, 1.0)) != cudaSuccess) return bail(e, "cudaMalloc bias");
  if ((e = h->d_params.reserve(sizeof(DevParams), 1.0)) != cudaSuccess) return bail(e, "cudaMalloc params");
  if ((e = cudaMemcpy(h->d_weights.p, weights, wbytes, cudaMemcpyHostToDevice)) != cudaSuccess) return bail(e, "upload weights");
  if ((e = cudaMemcpy(h->d_bias.p, bias, bbytes, cudaMemcpyHostToDevice)) != cudaSuccess) return bail(e, "upload bias");
  hp.weights = h->d_weights.as<double>();
  hp.bias = h->d_bias.as<double>();
  if (chebyshev) {
    // the two basis-conversion matrices of the angular passes (host arithmetic: annp_b200_basis_matrices)
    const int nt = hp.ntsf;
    std::vector<double> Md((size_t) nt * nt), Xd((size_t) nt * nt);
    if (annp_b200_basis_matrices(nt, Md.data(), Xd.data()) != ANNP_B200_OK)
      return bail(cudaErrorInvalidValue, "ntsf > 24 is not supported by the monomial conversion");
    if ((e = h->d_cheb2mono.reserve(sizeof(double) * Md.size(), 1.0)) != cudaSuccess) return bail(e, "cudaMalloc cheb2mono");
    if ((e = cudaMemcpy(h->d_cheb2mono.p, Md.data(), sizeof(double) * Md.size(), cudaMemcpyHostToDevice)) != cudaSuccess) return bail(e, "upload cheb2mono");
    hp.cheb2mono = h->d_cheb2mono.as<double>();
    if ((e = h->d_blk2cheb.reserve(sizeof(double) * Xd.size(), 1.0)) != cudaSuccess) return bail(e, "cudaMalloc blk2cheb");
    if ((e = cudaMemcpy(h->d_blk2cheb.p, Xd.data(), sizeof(double) * Xd.size(), cudaMemcpyHostToDevice)) != cudaSuccess) return bail(e, "upload blk2cheb");
    hp.blk2cheb = h->d_blk2cheb.as<double>();
  }
  if ((e = cudaMemcpy(h->d_params.p, &hp, sizeof(DevParams), cudaMemcpyHostToDevice)) != cudaSuccess) return bail(e, "upload params");
  return ANNP_B200_OK;
}

int annp_b200_init(const annp_b200_params *p, int device, int nall_hint, int max_nbors_hint, annp_b200_handle *out,
                   char *err, int errlen) {
  (void) nall_hint; (void) max_nbors_hint;
  if (!p || !out) { set_err(err, errlen, "null argument"); return ANNP_B200_EINVAL; }
  *out = nullptr;
  if (p->abi_version != ANNP_B200_ABI_VERSION) { set_err(err, errlen, "ABI version mismatch"); return ANNP_B200_EINVAL; }
  const int variant = p->variant & 0xff;
  const bool ni = variant == ANNP_B200_VARIANT_NI;
  if (variant != ANNP_B200_VARIANT_FE && !ni) { set_err(err, errlen, "unknown variant (ANNA-ADP handles are created by anna_b200_init)"); return ANNP_B200_EINVAL; }
  // the Ni files still say "Chebyshev" on their keyword line; the Ni copy of the style ignores flagsym altogether
  if (!ni && p->flagsym != ANNP_B200_SYM_CHEBYSHEV) { set_err(err, errlen, "only the Chebyshev descriptor (flagsym 0) is implemented for the Fe copy"); return ANNP_B200_EINVAL; }
  if (!valid_shape(p->ntypes, p->nelements, p->ntl, p->nnod, p->nsf, p->npsf, p->ntsf)) {
    set_err(err, errlen, "parameter block outside the supported range");
    return ANNP_B200_EINVAL;
  }
  int npsf_k = p->npsf, ntsf_k = p->ntsf;
  if (!ni && !annp_force_kernel_shape(p->npsf, p->ntsf, variant, &npsf_k, &ntsf_k)) {
    set_err(err, errlen, "Chebyshev descriptor beyond the kernel's range: npsf <= 16 and ntsf <= 24 (the monomial form of the angular polynomial is well conditioned up to degree 23)");
    return ANNP_B200_EINVAL;
  }
  const int nsf_k = npsf_k + ntsf_k;
  if (!p->sfnor_scal || !p->sfnor_avg || !p->cutsq || !p->map || !p->weights || !p->bias) { set_err(err, errlen, "null parameter array"); return ANNP_B200_EINVAL; }
  if (ni && (!p->sym_coerad || !p->sym_coeang)) { set_err(err, errlen, "the Ni variant needs sym_coerad / sym_coeang"); return ANNP_B200_EINVAL; }

  annp_b200_handle h = nullptr;
  int rc = open_handle(device, &h, err, errlen);
  if (rc) return rc;
  rc = fill_network(h, p->ntypes, p->nelements, p->ntl, p->nnod, 1, nsf_k, npsf_k, ntsf_k, p->flagact, p->cutsq, p->map, p->cut, err, errlen);
  if (rc) { annp_b200_clear(h); return rc; }
  h->npsf_file = p->npsf; h->ntsf_file = p->ntsf;
  DevParams &hp = h->hp;
  hp.variant = variant;
  hp.e_scale = p->e_scale; hp.e_shift = p->e_shift; hp.e_atom = p->e_atom;
  for (int n = 0; n < p->nsf; n++) { const int k = padded_index(n, p->npsf, npsf_k); hp.sf_scale[k] = p->sfnor_scal[n]; hp.sf_avg[k] = p->sfnor_avg[n]; }
  if (ni) {
    for (int m = 0; m < p->npsf; m++) hp.rad_eta[m] = p->sym_coerad[m * 3 + 0];
    hp.rad_rc = p->sym_coerad[2];                          // Rc of the first row, as the reference (ni/src/pair_annp.cpp:690)
    for (int n = 0; n < p->ntsf; n++) {
      hp.ang_eta[n] = p->sym_coeang[n * 4 + 0]; hp.ang_lambda[n] = p->sym_coeang[n * 4 + 1]; hp.ang_zeta[n] = p->sym_coeang[n * 4 + 2];
    }
    hp.ang_rc = p->sym_coeang[3];                          // ni/src/pair_annp.cpp:731
    hp.bp_layout = (p->variant & ANNP_B200_VARIANT_FLAG_GENERIC) ? 0 : annp_bp_layout(hp);
    if (hp.bp_layout == 1 && (p->variant & ANNP_B200_VARIANT_FLAG_NOPAIR)) hp.bp_layout = 2;
  }
  if (nsf_k != p->nsf) {
    const std::vector<double> wk = pad_weights(p->weights, p->nelements, p->ntl - 1, p->nnod, 1, p->nsf, p->npsf, npsf_k, nsf_k);
    rc = upload_params(h, wk.data(), p->bias, !ni, err, errlen);
  } else {
    rc = upload_params(h, p->weights, p->bias, !ni, err, errlen);
  }
  if (rc) { annp_b200_clear(h); return rc; }
  h->scatter_fixed = 1;
  *out = h;
  return ANNP_B200_OK;
}

int anna_b200_init(const anna_b200_params *p, int device, annp_b200_handle *out, char *err, int errlen) {
  if (!p || !out) { set_err(err, errlen, "null argument"); return ANNP_B200_EINVAL; }
  *out = nullptr;
  if (p->abi_version != ANNP_B200_ABI_VERSION) { set_err(err, errlen, "ABI version mismatch"); return ANNP_B200_EINVAL; }
  if (p->flagsym != ANNP_B200_SYM_CHEBYSHEV) { set_err(err, errlen, "only the Chebyshev descriptor (flagsym 0) is implemented"); return ANNP_B200_EINVAL; }
  if (!valid_shape(p->ntypes, p->nelements, p->ntl, p->nnod, p->nsf, p->npsf, p->ntsf) || p->nout != 2 || p->nout > p->nnod ||
      p->ngp < 17 || p->ngp > ANNA_B200_MAX_GPARAMS) {
    set_err(err, errlen, "parameter block outside the supported range (ANNA-ADP needs nout = 2 and >= 17 global parameters)");
    return ANNP_B200_EINVAL;
  }
  int npsf_k = p->npsf, ntsf_k = p->ntsf;
  if (!annp_force_kernel_shape(p->npsf, p->ntsf, ANNP_B200_VARIANT_ANNA_ADP, &npsf_k, &ntsf_k)) {
    set_err(err, errlen, "Chebyshev descriptor beyond the kernel's range: npsf <= 16 and ntsf <= 24");
    return ANNP_B200_EINVAL;
  }
  const int nsf_k = npsf_k + ntsf_k;
  if (!p->cutsq || !p->map || !p->weights || !p->bias || !p->gparams) { set_err(err, errlen, "null parameter array"); return ANNP_B200_EINVAL; }
  annp_b200_handle h = nullptr;
  int rc = open_handle(device, &h, err, errlen);
  if (rc) return rc;
  rc = fill_network(h, p->ntypes, p->nelements, p->ntl, p->nnod, p->nout, nsf_k, npsf_k, ntsf_k, p->flagact, p->cutsq, p->map, p->cut, err, errlen);
  if (rc) { annp_b200_clear(h); return rc; }
  h->npsf_file = p->npsf; h->ntsf_file = p->ntsf;
  DevParams &hp = h->hp;
  hp.variant = ANNP_B200_VARIANT_ANNA_ADP;
  hp.e_base = p->e_base;
  for (int k = 0; k < 17; k++) hp.gparams[k] = p->gparams[k];
  for (int n = 0; n < p->nsf; n++) { const int k = padded_index(n, p->npsf, npsf_k); hp.sf_scale[k] = 1.0; hp.sf_avg[k] = 0.0; }   // raw descriptor (pair_anna_adp.cpp:124-166)
  // the forward pass needs the block-basis conversion matrix
  if (nsf_k != p->nsf) {
    const std::vector<double> wk = pad_weights(p->weights, p->nelements, p->ntl - 1, p->nnod, p->nout, p->nsf, p->npsf, npsf_k, nsf_k);
    rc = upload_params(h, wk.data(), p->bias, true, err, errlen);
  } else
    rc = upload_params(h, p->weights, p->bias, true, err, errlen);
  if (rc) { annp_b200_clear(h); return rc; }
  h->scatter_fixed = 1;
  *out = h;
  return ANNP_B200_OK;
}

void annp_b200_clear(annp_b200_handle h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  DevBuf *bufs[] = {&h->d_cheb2mono, &h->d_blk2cheb, &h->d_params, &h->d_weights, &h->d_bias, &h->d_ilist, &h->d_row_off, &h->d_nbr, &h->d_rev_off, &h->d_rev_pos,
                    &h->d_centre_of, &h->d_scratch_cnt, &h->d_scratch_tmp, &h->d_tile_sum, &h->d_cell_of, &h->d_cell_cnt,
                    &h->d_cell_off, &h->d_cell_atoms, &h->d_row_cnt, &h->d_small, &h->d_xq, &h->d_fpair, &h->d_facc, &h->d_fself, &h->d_vir_c,
                    &h->d_vpair, &h->d_partial, &h->d_counters, &h->d_engvir, &h->d_Gdbg, &h->d_dEdbg, &h->d_x, &h->d_type, &h->d_f,
                    &h->d_eatom, &h->d_vatom, &h->d_goff, &h->d_glist, &h->d_ke_partial};
  for (DevBuf *b : bufs) b->release();
  annp_b200_peer_close(h);
  annp_b200_comm_destroy(h);
  DevBuf *more[] = {&h->d_ovf_list, &h->d_sl_tile_cnt, &h->d_sl_tile_off, &h->d_sl_tile_sum, &h->d_sendbuf, &h->d_recvbuf, &h->d_peer_table, &h->d_peer_sync};
  for (DevBuf *b : more) b->release();
  if (h->pin_list) cudaFreeHost(h->pin_list);
  for (int k = 0; k < annp_b200_handle_s::kEvRing; k++) {
    if (h->ev0[k]) cudaEventDestroy(h->ev0[k]);
    if (h->ev1[k]) cudaEventDestroy(h->ev1[k]);
  }
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

double annp_b200_bytes(annp_b200_handle h) {
  if (!h) return 0.0;
  const DevBuf *bufs[] = {&h->d_cheb2mono, &h->d_blk2cheb, &h->d_params, &h->d_weights, &h->d_bias, &h->d_ilist, &h->d_row_off, &h->d_nbr, &h->d_rev_off, &h->d_rev_pos,
                          &h->d_centre_of, &h->d_scratch_cnt, &h->d_scratch_tmp, &h->d_tile_sum, &h->d_cell_of, &h->d_cell_cnt,
                          &h->d_cell_off, &h->d_cell_atoms, &h->d_row_cnt, &h->d_small, &h->d_xq, &h->d_fpair, &h->d_facc, &h->d_fself, &h->d_vir_c,
                          &h->d_vpair, &h->d_partial, &h->d_counters, &h->d_engvir, &h->d_Gdbg, &h->d_dEdbg, &h->d_x, &h->d_type, &h->d_f,
                          &h->d_eatom, &h->d_vatom, &h->d_goff, &h->d_glist, &h->d_ke_partial};
  double b = 0.0;
  for (const DevBuf *d : bufs) b += (double) d->cap;
  const DevBuf *more[] = {&h->d_ovf_list, &h->d_sl_tile_cnt, &h->d_sl_tile_off, &h->d_sl_tile_sum, &h->d_sendbuf, &h->d_recvbuf};
  for (const DevBuf *d : more) b += (double) d->cap;
  return b;
}

const char *annp_b200_last_error(annp_b200_handle h) { return h ? h->err.c_str() : "null handle"; }

int annp_b200_neigh_csr(annp_b200_handle h, int inum, int nall, const int *ilist, const int64_t *offsets, const int *neigh) {
  if (!h) return ANNP_B200_EINVAL;
  if (inum < 0 || nall < 0 || (inum > 0 && (!ilist || !offsets))) return fail(h, ANNP_B200_EINVAL, "bad neighbour list arguments");
  CK(cudaSetDevice(h->device));
  cudaStream_t s = h->stream;
  const long long total = inum > 0 ? (long long) offsets[inum] : 0;
  h->inum = inum; h->nall_list = nall; h->total = total;
  int maxrow = 0;
  for (int ii = 0; ii < inum; ii++) maxrow = std::max(maxrow, (int) (offsets[ii + 1] - offsets[ii]));
  h->max_row = maxrow;
  CK(h->d_ilist.reserve(sizeof(int) * (size_t) std::max(inum, 1)));
  CK(h->d_row_off.reserve(sizeof(long long) * ((size_t) inum + 1)));
  CK(h->d_nbr.reserve(sizeof(int) * (size_t) std::max<long long>(total, 1)));
  if (inum > 0) {
    CK(cudaMemcpyAsync(h->d_ilist.p, ilist, sizeof(int) * (size_t) inum, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(h->d_row_off.p, offsets, sizeof(long long) * ((size_t) inum + 1), cudaMemcpyHostToDevice, s));
    if (total > 0) CK(cudaMemcpyAsync(h->d_nbr.p, neigh, sizeof(int) * (size_t) total, cudaMemcpyHostToDevice, s));
  } else {
    const long long zero = 0;
    CK(cudaMemcpyAsync(h->d_row_off.p, &zero, sizeof(long long), cudaMemcpyHostToDevice, s));
  }
  int bad[2] = {0, 0};
  if (inum > 0) {
    if (offsets[0] != 0) return fail(h, ANNP_B200_EINVAL, "neighbour list offsets must start at 0");
    CK(h->d_small.reserve(64));
    CK(cudaMemsetAsync(h->d_small.p, 0, 2 * sizeof(int), s));
    k_validate_list<<<std::min<long long>(148 * 8, (std::max<long long>(total, inum) + 255) / 256), 256, 0, s>>>(
        h->d_ilist.as<int>(), h->d_row_off.as<long long>(), h->d_nbr.as<int>(), inum, total, nall, h->d_small.as<int>());
    h->launches += 1;
    CK(cudaMemcpyAsync(bad, h->d_small.p, 2 * sizeof(int), cudaMemcpyDeviceToHost, s));
  }
  int rc = finish_list(h, s);
  if (rc) return rc;
  CK(cudaStreamSynchronize(s));
  if (bad[0] || bad[1]) {
    h->have_list = false;
    return fail(h, ANNP_B200_EINVAL, bad[1] ? "neighbour list offsets decrease" : "neighbour list names atoms outside 0..nall-1");
  }
  return ANNP_B200_OK;
}

int annp_b200_neigh(annp_b200_handle h, int inum, int nall, const int *ilist, const int *numneigh, const int *const *firstneigh) {
  if (!h) return ANNP_B200_EINVAL;
  if (inum < 0 || (inum > 0 && (!ilist || !numneigh || !firstneigh))) return fail(h, ANNP_B200_EINVAL, "bad neighbour list arguments");
  CK(cudaSetDevice(h->device));
  // flatten LAMMPS' paged rows (NeighList::firstneigh) into CSR in ilist order, straight into a page-locked staging buffer
  // (kept across calls) so the upload is one DMA at PCIe speed instead of a pageable copy
  std::vector<int64_t> off((size_t) inum + 1, 0);
  for (int ii = 0; ii < inum; ii++) {
    const int i = ilist[ii];
    if (i < 0 || i >= nall || numneigh[i] < 0) return fail(h, ANNP_B200_EINVAL, "neighbour list names atoms outside 0..nall-1");
    off[ii + 1] = off[ii] + numneigh[i];
  }
  const size_t bytes = sizeof(int) * (size_t) std::max<int64_t>(off[inum], 1);
  if (bytes > h->pin_list_cap) {
    if (h->pin_list) cudaFreeHost(h->pin_list);
    h->pin_list = nullptr; h->pin_list_cap = 0;
    const size_t want = bytes + bytes / 8;
    if (cudaHostAlloc(&h->pin_list, want, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return fail(h, ANNP_B200_ENOMEM, "cannot page-lock the neighbour list staging buffer"); }
    h->pin_list_cap = want;
  }
  int *flat = (int *) h->pin_list;
  // half a gigabyte of row copies at 524 288 atoms: a few host threads (LAMMPS runs one rank per GPU, the other cores idle)
  const int nthreads = (int) std::max<int64_t>(1, std::min<int64_t>({8, (int64_t) std::thread::hardware_concurrency() / 2, off[inum] / (1 << 22)}));
  auto copy_rows = [&](int lo, int hi) {
    for (int ii = lo; ii < hi; ii++) {
      const int i = ilist[ii];
      if (numneigh[i] > 0) memcpy(flat + off[ii], firstneigh[i], sizeof(int) * (size_t) numneigh[i]);
    }
  };
  if (nthreads <= 1) {
    copy_rows(0, inum);
  } else {
    std::vector<std::thread> pool;
    for (int t = 0; t < nthreads; t++)
      pool.emplace_back(copy_rows, (int) ((int64_t) inum * t / nthreads), (int) ((int64_t) inum * (t + 1) / nthreads));
    for (std::thread &t : pool) t.join();
  }
  return annp_b200_neigh_csr(h, inum, nall, ilist, off.data(), flat);
}

int annp_b200_host_register(void *ptr, size_t bytes) {
  if (!ptr || bytes == 0) return ANNP_B200_EINVAL;
  const cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterDefault);
  if (e == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return ANNP_B200_OK; }
  if (e != cudaSuccess) { cudaGetLastError(); return e == cudaErrorMemoryAllocation ? ANNP_B200_ENOMEM : ANNP_B200_ECUDA; }
  return ANNP_B200_OK;
}

int annp_b200_host_unregister(void *ptr) {
  if (!ptr) return ANNP_B200_OK;
  const cudaError_t e = cudaHostUnregister(ptr);
  if (e != cudaSuccess) { cudaGetLastError(); return e == cudaErrorHostMemoryNotRegistered ? ANNP_B200_OK : ANNP_B200_ECUDA; }
  return ANNP_B200_OK;
}

void *annp_b200_host_alloc(size_t bytes) {
  void *p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return p;
}

void annp_b200_host_free(void *ptr) {
  if (ptr && cudaFreeHost(ptr) != cudaSuccess) cudaGetLastError();
}

int annp_b200_compute(annp_b200_handle h, int nlocal, int nghost, const double *x, const int *type, int eflag, int vflag,
                      double *f, double *eng, double *eatom, double *virial6, double *vatom) {
  if (!h) return ANNP_B200_EINVAL;
  const int nall = nlocal + nghost;
  if (nall < 0 || (nall > 0 && !x)) return fail(h, ANNP_B200_EINVAL, "bad position/type arguments");
  if (nall > 0 && !type && !(h->types_valid && h->d_type.cap >= sizeof(int) * (size_t) nall))
    return fail(h, ANNP_B200_ESTATE, "type == NULL needs an earlier call that passed the types of these atoms");
  CK(cudaSetDevice(h->device));
  cudaStream_t s = h->stream;
  CK(h->d_x.reserve(sizeof(double) * 3 * (size_t) std::max(nall, 1)));
  if (type || !h->d_type.p) { h->types_valid = false; CK(h->d_type.reserve(sizeof(int) * (size_t) std::max(nall, 1))); }
  CK(h->d_f.reserve(sizeof(double) * 3 * (size_t) std::max(nall, 1)));
  CK(h->d_engvir.reserve(sizeof(double) * 8));
  if (eatom) CK(h->d_eatom.reserve(sizeof(double) * (size_t) std::max(nall, 1)));
  if (vatom) CK(h->d_vatom.reserve(sizeof(double) * 6 * (size_t) std::max(nall, 1)));
  if (nall > 0) {
    // x / type / f may be pageable (the copy is then staged by the driver) or page-locked (annp_b200_host_register /
    // annp_b200_host_alloc: direct DMA at PCIe speed)
    CK(cudaMemcpyAsync(h->d_x.p, x, sizeof(double) * 3 * (size_t) nall, cudaMemcpyHostToDevice, s));
    if (type) {
      CK(cudaMemcpyAsync(h->d_type.p, type, sizeof(int) * (size_t) nall, cudaMemcpyHostToDevice, s));
      h->types_valid = true;
    }
  }
  if (eatom) CK(cudaMemsetAsync(h->d_eatom.p, 0, sizeof(double) * (size_t) std::max(nall, 1), s));
  const bool want_ev = (eng && eflag) || (virial6 && vflag);
  int rc = step_device(h, nlocal, nghost, h->d_x.as<double>(), h->d_type.as<int>(), eflag, (virial6 || vatom) ? (vflag ? vflag : 1) : 0,
                       f ? h->d_f.as<double>() : nullptr, eatom ? h->d_eatom.as<double>() : nullptr,
                       want_ev ? h->d_engvir.as<double>() : nullptr, vatom ? h->d_vatom.as<double>() : nullptr, s, true);
  if (rc) return rc;
  // results are queued behind the kernels; ONE synchronisation at the end covers them and the counters
  double ev[7] = {0, 0, 0, 0, 0, 0, 0};
  if (want_ev) CK(cudaMemcpyAsync(ev, h->d_engvir.p, sizeof(double) * 7, cudaMemcpyDeviceToHost, s));
  if (f && nall > 0) CK(cudaMemcpyAsync(f, h->d_f.p, sizeof(double) * 3 * (size_t) nall, cudaMemcpyDeviceToHost, s));
  if (eatom && nall > 0) CK(cudaMemcpyAsync(eatom, h->d_eatom.p, sizeof(double) * (size_t) nall, cudaMemcpyDeviceToHost, s));
  if (vatom && nall > 0) CK(cudaMemcpyAsync(vatom, h->d_vatom.p, sizeof(double) * 6 * (size_t) nall, cudaMemcpyDeviceToHost, s));
  rc = fetch_counters(h, s);          // synchronises the stream
  if (rc) return rc;
  if (h->last_cnt.bad_force)
    return fail(h, ANNP_B200_EOVERFLOW, "a pair force is NaN or beyond 2^18 eV/A (fixed-point force accumulation): the configuration is unphysical");
  if (h->last_cnt.overflow)
    return fail(h, ANNP_B200_EOVERFLOW, "an atom has more in-cutoff neighbours than ANNP_B200_MAX_NEIGH");
  if (eng && eflag) *eng = ev[0];
  if (virial6 && vflag) for (int k = 0; k < 6; k++) virial6[k] = ev[1 + k];
  return ANNP_B200_OK;
}

int annp_b200_compute_device(annp_b200_handle h, int nlocal, int nghost, const double *d_x, const int *d_type, int eflag, int vflag,
                             double *d_f, double *d_eatom, double *d_eng_virial, double *d_vatom, void *stream) {
  if (!h) return ANNP_B200_EINVAL;
  CK(cudaSetDevice(h->device));
  return step_device(h, nlocal, nghost, d_x, d_type, eflag, vflag, d_f, d_eatom, d_eng_virial, d_vatom, (cudaStream_t) stream,
                     /*allow_sync=*/h->need_calibrate);
}

int annp_b200_neigh_build(annp_b200_handle h, int nlocal, int nall, const double *d_x, const double *bbox_lo, const double *bbox_hi,
                          double cutneigh, void *stream) {
  if (!h) return ANNP_B200_EINVAL;
  if (nlocal < 0 || nall < nlocal || !d_x || !bbox_lo || !bbox_hi || !(cutneigh > 0.0)) return fail(h, ANNP_B200_EINVAL, "bad neigh_build arguments");
  CK(cudaSetDevice(h->device));
  cudaStream_t s = (cudaStream_t) stream;
  int n[3];
  double inv[3];
  long long ncells = 0;
  neigh_grid(bbox_lo, bbox_hi, cutneigh, n, inv, &ncells);
  if (ncells > (1LL << 30)) return fail(h, ANNP_B200_EINVAL, "bounding box too large for the cell grid");
  CK(h->d_cell_of.reserve(sizeof(int) * (size_t) std::max(nall, 1)));
  CK(h->d_cell_cnt.reserve(sizeof(int) * (size_t) ncells));
  CK(h->d_cell_off.reserve(sizeof(long long) * ((size_t) ncells + 1)));
  CK(h->d_cell_atoms.reserve(sizeof(int) * (size_t) std::max(nall, 1)));
  CK(h->d_row_cnt.reserve(sizeof(int) * (size_t) std::max(nlocal, 1)));
  CK(h->d_tile_sum.reserve(sizeof(long long) * ((size_t) std::max<long long>(std::max<long long>(ncells, nall), 1) / 1024 + 2)));
  CK(h->d_row_off.reserve(sizeof(long long) * ((size_t) nlocal + 1)));
  CK(h->d_ilist.reserve(sizeof(int) * (size_t) std::max(nlocal, 1)));
  CK(h->d_small.reserve(64));
  NeighScratch sc;
  sc.cell_of = h->d_cell_of.as<int>(); sc.cell_cnt = h->d_cell_cnt.as<int>(); sc.cell_atoms = h->d_cell_atoms.as<int>();
  sc.row_cnt = h->d_row_cnt.as<int>(); sc.cell_off = h->d_cell_off.as<long long>(); sc.tile_sum = h->d_tile_sum.as<long long>();
  neigh_count(d_x, nlocal, nall, bbox_lo, n, inv, cutneigh, sc, h->d_row_off.as<long long>(), h->d_small.as<int>(), s);
  h->launches += 12;
  long long total = 0;
  int maxrow = 0;
  CK(cudaMemcpyAsync(&total, h->d_row_off.as<long long>() + nlocal, sizeof(long long), cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(&maxrow, h->d_small.p, sizeof(int), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  h->inum = nlocal; h->nall_list = nall; h->total = total; h->max_row = maxrow;
  CK(h->d_nbr.reserve(sizeof(int) * (size_t) std::max<long long>(total, 1)));
  CK(h->d_scratch_tmp.reserve(sizeof(int) * (size_t) std::max<long long>(total, 1)));
  neigh_fill(d_x, nlocal, bbox_lo, n, inv, cutneigh, sc, h->d_row_off.as<long long>(), h->d_scratch_tmp.as<int>(), h->d_nbr.as<int>(), s);
  h->launches += 2;
  aux_iota(h->d_ilist.as<int>(), nlocal, s);      // ilist = 0..nlocal-1
  h->launches += 1;
  return finish_list(h, s);
}

int annp_b200_neigh_build_host(annp_b200_handle h, int nlocal, int nall, const double *x, const double *bbox_lo, const double *bbox_hi,
                               double cutneigh) {
  if (!h) return ANNP_B200_EINVAL;
  if (nlocal < 0 || nall < nlocal || (nall > 0 && !x)) return fail(h, ANNP_B200_EINVAL, "bad neigh_build arguments");
  CK(cudaSetDevice(h->device));
  CK(h->d_x.reserve(sizeof(double) * 3 * (size_t) std::max(nall, 1)));
  if (nall > 0) CK(cudaMemcpyAsync(h->d_x.p, x, sizeof(double) * 3 * (size_t) nall, cudaMemcpyHostToDevice, h->stream));
  const int rc = annp_b200_neigh_build(h, nlocal, nall, h->d_x.as<double>(), bbox_lo, bbox_hi, cutneigh, h->stream);
  if (rc) return rc;
  CK(cudaStreamSynchronize(h->stream));
  return ANNP_B200_OK;
}

int annp_b200_set_halo(annp_b200_handle h, int nlocal, int nsend, const int *d_index, const double *d_shift, void *stream) {
  if (!h) return ANNP_B200_EINVAL;
  if (nlocal < 0 || nsend < 0 || (nsend > 0 && (!d_index || !d_shift))) return fail(h, ANNP_B200_EINVAL, "bad halo arguments");
  CK(cudaSetDevice(h->device));
  cudaStream_t s = (cudaStream_t) stream;
  h->g_nlocal = nlocal; h->g_nghost = nsend; h->g_owner = d_index; h->g_shift = d_shift;
  CK(h->d_goff.reserve(sizeof(long long) * ((size_t) nlocal + 1)));
  CK(h->d_glist.reserve(sizeof(int) * (size_t) std::max(nsend, 1)));
  CK(h->d_scratch_cnt.reserve(sizeof(int) * (size_t) std::max(nlocal, 1)));
  CK(h->d_scratch_tmp.reserve(sizeof(int) * (size_t) std::max(nsend, 1)));
  CK(h->d_tile_sum.reserve(sizeof(long long) * ((size_t) nlocal / 1024 + 2)));
  aux_build_ghost_csr(d_index, nsend, nlocal, h->d_goff.as<long long>(), h->d_glist.as<int>(), h->d_scratch_cnt.as<int>(),
                      h->d_scratch_tmp.as<int>(), h->d_tile_sum.as<long long>(), s);
  h->launches += 7;
  CK(cudaGetLastError());
  return ANNP_B200_OK;
}

int annp_b200_halo_pack(annp_b200_handle h, const double *d_x, double *d_sendbuf, void *stream) {
  if (!h || !d_x || (h->g_nghost > 0 && !d_sendbuf)) return ANNP_B200_EINVAL;
  aux_halo_pack(h->g_nghost, h->g_owner, h->g_shift, d_x, d_sendbuf, (cudaStream_t) stream);
  h->launches += 1;
  return ANNP_B200_OK;
}

int annp_b200_halo_unpack_add(annp_b200_handle h, const double *d_recvbuf, double *d_f, void *stream) {
  if (!h || !d_f || (h->g_nghost > 0 && !d_recvbuf)) return ANNP_B200_EINVAL;
  aux_halo_unpack_add(h->g_nlocal, h->d_goff.as<long long>(), h->d_glist.as<int>(), d_recvbuf, d_f, (cudaStream_t) stream);
  h->launches += 1;
  return ANNP_B200_OK;
}

// LAMMPS metal units: ftm2v = 1 / 1.0364269e-4, mvv2e = 1.0364269e-4   (update.cpp, units metal)
static const double kFtm2v = 1.0 / 1.0364269e-4;
static const double kMvv2e = 1.0364269e-4;

int annp_b200_nve_initial(annp_b200_handle h, int nlocal, double dt, double mass, double *d_x, double *d_v, const double *d_f, void *stream) {
  if (!h || !d_x || !d_v || !d_f || !(mass > 0.0)) return ANNP_B200_EINVAL;
  aux_nve_initial(nlocal, dt, 0.5 * dt * kFtm2v / mass, d_x, d_v, d_f, (cudaStream_t) stream);
  h->launches += 1;
  return ANNP_B200_OK;
}

int annp_b200_nve_final(annp_b200_handle h, int nlocal, double dt, double mass, double *d_v, const double *d_f, double *d_ke, void *stream) {
  if (!h || !d_v || !d_f || !(mass > 0.0)) return ANNP_B200_EINVAL;
  aux_nve_final(nlocal, 0.5 * dt * kFtm2v / mass, d_v, d_f, (cudaStream_t) stream);
  h->launches += 1;
  if (d_ke) {
    CK(h->d_ke_partial.reserve(sizeof(double) * (size_t) aux_ke_blocks(nlocal)));
    aux_kinetic(nlocal, d_v, 0.5 * mass * kMvv2e, h->d_ke_partial.as<double>(), d_ke, (cudaStream_t) stream);
    h->launches += 2;
  }
  return ANNP_B200_OK;
}

int annp_b200_max_displacement_sq(annp_b200_handle h, int nlocal, const double *d_x, const double *d_xref, double *d_out, void *stream) {
  if (!h || nlocal < 0 || !d_out || (nlocal > 0 && (!d_x || !d_xref))) return ANNP_B200_EINVAL;
  aux_max_disp2(nlocal, d_x, d_xref, d_out, (cudaStream_t) stream);
  h->launches += 1;
  return ANNP_B200_OK;
}

double annp_b200_fp64_peak_tflops(annp_b200_handle h, int reps) {
  if (!h) return -1.0;
  cudaSetDevice(h->device);
  return aux_fp64_peak_tflops(h->num_sms, reps < 1 ? 1 : reps, h->stream);
}

int annp_b200_get_stats(annp_b200_handle h, annp_b200_stats *out) {
  if (!h || !out) return ANNP_B200_EINVAL;
  CK(cudaSetDevice(h->device));
  CK(cudaDeviceSynchronize());
  if (h->d_counters.p) {
    int rc = fetch_counters(h, h->stream);
    if (rc) return rc;
  }
  out->inum = h->inum; out->nall = h->nall_list; out->max_neigh_list = h->max_row;
  out->max_neigh_cut = h->last_cnt.max_neigh;
  out->avg_neigh_cut = h->inum > 0 ? (double) h->last_cnt.sum_neigh / h->inum : 0.0;
  out->sum_triplets = (double) h->last_cnt.sum_trip;
  out->kernel_launches = h->launches;
  out->last_force_kernel_ms = h->last_force_ms;
  out->force_kernel_ms_total = h->force_ms_total;
  out->force_kernel_samples = h->force_samples;
  out->overflow_pass_atoms = (long long) h->last_cnt.ovf_total;
  out->tile_capacity = h->capacity;
  for (int k = 0; k < 8; k++) out->stage_cycles[k] = (double) h->last_cnt.stage_clk[k];
  // sticky flags of every device-mode step since the last call (fetch_counters cleared them on the device)
  if (h->last_cnt.overflow) return fail(h, ANNP_B200_EOVERFLOW, "an atom had more in-cutoff neighbours than ANNP_B200_MAX_NEIGH in a device-mode step since the last check: results from that step on are invalid");
  if (h->last_cnt.bad_force) return fail(h, ANNP_B200_EOVERFLOW, "a pair force was NaN or beyond 2^18 eV/A in a device-mode step since the last check (fixed-point force accumulation)");
  return ANNP_B200_OK;
}

int annp_b200_set_scatter(annp_b200_handle h, int mode) {
  if (!h || (mode != ANNP_B200_SCATTER_GATHER && mode != ANNP_B200_SCATTER_FIXED)) return ANNP_B200_EINVAL;
  h->scatter_fixed = mode == ANNP_B200_SCATTER_FIXED;
  return ANNP_B200_OK;
}

int annp_b200_set_timing(annp_b200_handle h, int enabled) {
  if (!h) return ANNP_B200_EINVAL;
  h->timing = enabled != 0;
  h->force_ms_total = 0.0;
  h->force_samples = 0;
  h->ev_count = 0;
  return ANNP_B200_OK;
}

long long annp_b200_debug_neighbors(annp_b200_handle h, int64_t *offsets, int *neigh) {
  if (!h) return ANNP_B200_EINVAL;
  if (!h->have_list) return fail(h, ANNP_B200_ESTATE, "no neighbour list");
  if (cudaSetDevice(h->device) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) return fail(h, ANNP_B200_ECUDA, "device synchronisation failed");
  static_assert(sizeof(int64_t) == sizeof(long long), "row offsets are 64-bit");
  if (offsets && cudaMemcpy(offsets, h->d_row_off.p, sizeof(long long) * ((size_t) h->inum + 1), cudaMemcpyDeviceToHost) != cudaSuccess)
    return fail(h, ANNP_B200_ECUDA, "copying the row offsets failed");
  if (neigh && h->total > 0 && cudaMemcpy(neigh, h->d_nbr.p, sizeof(int) * (size_t) h->total, cudaMemcpyDeviceToHost) != cudaSuccess)
    return fail(h, ANNP_B200_ECUDA, "copying the neighbour rows failed");
  return h->total;
}

int annp_b200_debug_descriptors(annp_b200_handle h, double *G, double *dE_dG) {
  if (!h) return ANNP_B200_EINVAL;
  if (!h->debug_desc) { h->debug_desc = true; return ANNP_B200_OK; }   // first call arms the capture
  if (!h->d_Gdbg.p) return fail(h, ANNP_B200_ESTATE, "no compute since the capture was armed");
  CK(cudaSetDevice(h->device));
  CK(cudaDeviceSynchronize());
  const int nsf_k = h->hp.nsf, nsf = h->npsf_file + h->ntsf_file;
  std::vector<double> tmp((size_t) nsf_k * std::max(h->inum, 1));
  for (int pass = 0; pass < 2; pass++) {
    double *dst = pass ? dE_dG : G;
    if (!dst) continue;
    CK(cudaMemcpy(tmp.data(), pass ? h->d_dEdbg.p : h->d_Gdbg.p, sizeof(double) * (size_t) nsf_k * h->inum, cudaMemcpyDeviceToHost));
    for (int ii = 0; ii < h->inum; ii++)
      for (int n = 0; n < nsf; n++) dst[(size_t) ii * nsf + n] = tmp[(size_t) ii * nsf_k + padded_index(n, h->npsf_file, h->hp.npsf)];
  }
  return ANNP_B200_OK;
}

}    // extern "C"

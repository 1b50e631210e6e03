// annp_handle.cuh -- the handle behind the C ABI (include/annp_b200.h), shared by annp_capi.cu (life cycle, per-step launch
// sequence) and annp_halo.cu (ghost map, NCCL exchange).  Private to the library.
#pragma once
#include "annp_device.cuh"

#include <cstdio>
#include <string>
#include <vector>

// growable device allocation
struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes, double slack = 1.1) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = (size_t) ((double) bytes * slack) + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) { want = bytes; e = cudaMalloc(&p, want); }
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

struct annp_b200_handle_s {
  int device = 0;
  int num_sms = 0;
  cudaStream_t stream = nullptr;      // host-mode stream
  DevParams hp;                       // host copy (weights/bias pointers are device pointers)
  DevBuf d_params, d_weights, d_bias, d_cheb2mono, d_blk2cheb;
  // neighbour list
  bool have_list = false;
  int inum = 0, nall_list = 0, max_row = 0;
  long long total = 0;
  DevBuf d_ilist, d_row_off, d_nbr, d_rev_off, d_rev_pos, d_centre_of, d_scratch_cnt, d_scratch_tmp, d_tile_sum;
  // device neighbour build scratch
  DevBuf d_cell_of, d_cell_cnt, d_cell_off, d_cell_atoms, d_row_cnt, d_small;
  // per-step
  DevBuf d_xq, d_fpair, d_facc, d_fself, d_vir_c, d_vpair, d_partial, d_counters, d_engvir, d_Gdbg, d_dEdbg, d_ovf_list;
  // descriptor shape of the potential; hp.npsf / hp.ntsf / hp.nsf are the (possibly padded) shape of the kernel instantiation
  int npsf_file = 0, ntsf_file = 0;
  // how neighbour forces reach f: 1 = fixed-point integer atomics into d_facc (default),
  // 0 = per-entry pair forces (d_fpair) summed by an ordered gather over the reverse map (annp_b200_set_scatter)
  int scatter_fixed = 0;
  bool have_reverse = false;
  // host-mode staging
  DevBuf d_x, d_type, d_f, d_eatom, d_vatom;
  // ghosts
  int g_nlocal = 0, g_nghost = 0;
  const int *g_owner = nullptr;
  const double *g_shift = nullptr;
  DevBuf d_goff, d_glist, d_ke_partial;
  // ghost map built on the device (annp_halo.cu: annp_b200_send_lists_count / _fill)
  DevBuf d_sl_tile_cnt, d_sl_tile_off, d_sl_tile_sum;
  int sl_nlocal = 0, sl_nslots = 0, sl_ntiles = 0;
  const double *sl_x = nullptr;
  double sl_lo[3] = {0, 0, 0}, sl_hi[3] = {0, 0, 0}, sl_cut = 0.0;
  int sl_dir[26][3] = {};
  // halo exchange over NCCL (annp_halo.cu: annp_b200_comm_* / annp_b200_halo_forward / _reverse)
  void *nccl_comm = nullptr;          // ncclComm_t of this rank, created by annp_b200_comm_init
  int comm_rank = 0, comm_size = 1;
  std::vector<int> peer_send, peer_recv;   // atoms sent to / received from every rank per exchange
  DevBuf d_sendbuf, d_recvbuf;
  // peer scatter (annp_halo.cu: annp_b200_peer_*): ghost contributions go straight to the owner rank's facc over NVLink
  bool peer_on = false;
  std::vector<void *> peer_ptr;            // IPC mappings of the other ranks' facc (null for this rank)
  std::vector<std::string> peer_key;       // the 64 handle bytes each mapping was opened from
  DevBuf d_peer_table, d_peer_sync;        // device array of the nranks base pointers; one double for the step barrier
  const int *peer_ghost_rank = nullptr, *peer_ghost_index = nullptr;
  bool peer_zeroed = false;                // the accumulators were zeroed by annp_b200_halo_forward of this step already
  int g_nghost_recv = 0;                   // ghosts this rank holds (sum of the receive counts of set_halo_peers)
  int capacity = 0;
  bool need_calibrate = true;
  bool types_valid = false;           // host mode: d_type holds the types of the current atoms (annp_b200_compute with type == NULL)
  void *pin_list = nullptr;           // pinned staging of the flattened host neighbour list (annp_b200_neigh)
  size_t pin_list_cap = 0;
  bool timing = false;
  bool debug_desc = false;
  static constexpr int kEvRing = 256;
  cudaEvent_t ev0[kEvRing] = {}, ev1[kEvRing] = {};
  int ev_count = 0;                   // timed launches recorded and not yet collected
  float last_force_ms = 0.f;
  double force_ms_total = 0.0;
  int force_samples = 0;
  long long launches = 0;
  DevCounters last_cnt;
  std::string err;
};

static inline int fail(annp_b200_handle h, int code, const std::string &msg) {
  if (h) h->err = msg;
  return code;
}
static inline int cuda_fail(annp_b200_handle h, cudaError_t e, const char *where) {
  return fail(h, e == cudaErrorMemoryAllocation ? ANNP_B200_ENOMEM : ANNP_B200_ECUDA,
              std::string(where) + ": " + cudaGetErrorString(e));
}
#define CK(call)                                                         \
  do {                                                                   \
    cudaError_t e__ = (call);                                            \
    if (e__ != cudaSuccess) return cuda_fail(h, e__, #call);             \
  } while (0)

static inline void set_err(char *err, int errlen, const std::string &msg) {
  if (err && errlen > 0) snprintf(err, (size_t) errlen, "%s", msg.c_str());
}

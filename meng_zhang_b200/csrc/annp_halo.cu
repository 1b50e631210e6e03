// annp_halo.cu -- ghost map and halo exchange of the device-resident mode, below the C ABI.
//
// LAMMPS' Comm::borders builds, at every re-neighbouring, the list of local atoms each neighbouring sub-domain needs as
// ghosts; Comm::forward_comm / reverse_comm then move positions out and forces back every step, staged through host
// buffers and MPI (the reference: fe_v2/lib/lal_annp.cpp:310-312, 336-347 around LAMMPS' own comm).  Here
//   * annp_b200_send_lists_count / _fill build that list on the device: every local atom is classified against the 26
//     directions of the brick decomposition, entries are compacted in (destination slot, atom index) order by a
//     two-level prefix sum - deterministic, and only the 26 slot counts travel to the host;
//   * annp_b200_comm_init creates an NCCL communicator for the handle (the unique id travels over whatever the caller
//     has: MPI_Bcast inside LAMMPS, torch.distributed in the stand-alone driver);
//   * annp_b200_halo_forward = pack kernel + ONE grouped ncclSend/ncclRecv over NVLink straight into the ghost block of x,
//     annp_b200_halo_reverse = grouped send/recv of the ghost forces + ordered add on the owners.  Both are plain stream
//     work, so a whole multi-GPU MD step can be captured in a CUDA graph.
// NCCL is taken from the process (dlopen of libnccl.so.2: inside a torch process that is the copy torch already loaded),
// so the library itself has no link-time dependency on it and loads on boxes without NCCL.
#include "annp_handle.cuh"

#include <dlfcn.h>

#include <algorithm>
#include <cstring>

void aux_exclusive_scan(const int *cnt, long long *off, int n, long long *tile_sum, cudaStream_t s);
void aux_halo_pack(int nsend, const int *idx, const double *shift, const double *x, double *out, cudaStream_t s);
void aux_halo_unpack_add(int nlocal, const long long *goff, const int *glist, const double *src, double *f, cudaStream_t s);

namespace {

constexpr int kSlTile = 256;          // atoms per block of the ghost-map kernels
constexpr int kSlWarps = kSlTile / 32;

struct SendGeom {
  double lo[3], hi[3], cut;
  int nslots;
  signed char dir[26][3];
};

// bit d of `lo_bits` / `hi_bits`: the atom is within `cut` of the lower / upper face of the brick along d
__device__ __forceinline__ void face_bits(const double *__restrict__ x, int i, int n, const SendGeom &g, unsigned &lo_bits, unsigned &hi_bits) {
  lo_bits = hi_bits = 0u;
  if (i >= n) return;
#pragma unroll
  for (int d = 0; d < 3; d++) {
    const double v = x[3 * (size_t) i + d];
    if (v < g.lo[d] + g.cut) lo_bits |= 1u << d;
    if (v >= g.hi[d] - g.cut) hi_bits |= 1u << d;
  }
}
__device__ __forceinline__ bool in_slot(const SendGeom &g, int k, unsigned lo_bits, unsigned hi_bits, bool valid) {
  bool in = valid;
#pragma unroll
  for (int d = 0; d < 3; d++) {
    const int s = g.dir[k][d];
    in = in && (s == 0 || (s > 0 ? (hi_bits >> d) & 1u : (lo_bits >> d) & 1u));
  }
  return in;
}

// tile_cnt[k * ntiles + tile] = atoms of the tile that go to slot k
__global__ void __launch_bounds__(kSlTile) k_send_count(const double *__restrict__ x, int n, SendGeom g, int ntiles, int *__restrict__ tile_cnt) {
  __shared__ int cnt[26];
  if (threadIdx.x < 26) cnt[threadIdx.x] = 0;
  __syncthreads();
  const int i = blockIdx.x * kSlTile + threadIdx.x;
  unsigned lb, hb;
  face_bits(x, i, n, g, lb, hb);
  for (int k = 0; k < g.nslots; k++) {
    const unsigned m = __ballot_sync(0xffffffffu, in_slot(g, k, lb, hb, i < n));
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(&cnt[k], __popc(m));
  }
  __syncthreads();
  if (threadIdx.x < g.nslots) tile_cnt[threadIdx.x * ntiles + blockIdx.x] = cnt[threadIdx.x];
}

// entry (slot k, atom i) lands at tile_off[k * ntiles + tile] + (atoms of the tile before i that go to k)
__global__ void __launch_bounds__(kSlTile) k_send_fill(const double *__restrict__ x, int n, SendGeom g, int ntiles,
                                                       const long long *__restrict__ tile_off, const double *__restrict__ slot_shift,
                                                       int *__restrict__ send_index, double *__restrict__ send_shift) {
  __shared__ int wcnt[26][kSlWarps];
  const int i = blockIdx.x * kSlTile + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned lb, hb;
  face_bits(x, i, n, g, lb, hb);
  unsigned mine = 0u;                 // slots of this atom
  int rank_in_warp[26];
#pragma unroll
  for (int k = 0; k < 26; k++) {
    rank_in_warp[k] = 0;
    if (k < g.nslots) {
      const bool in = in_slot(g, k, lb, hb, i < n);
      const unsigned m = __ballot_sync(0xffffffffu, in);
      if (in) { mine |= 1u << k; rank_in_warp[k] = __popc(m & ((1u << lane) - 1u)); }
      if (lane == 0) wcnt[k][warp] = __popc(m);
    }
  }
  __syncthreads();
  if (!mine) return;
#pragma unroll
  for (int k = 0; k < 26; k++) {
    if (!((mine >> k) & 1u)) continue;
    int base = 0;
    for (int w = 0; w < warp; w++) base += wcnt[k][w];
    const long long pos = tile_off[(size_t) k * ntiles + blockIdx.x] + base + rank_in_warp[k];
    send_index[pos] = i;
    send_shift[3 * pos + 0] = slot_shift[3 * k + 0];
    send_shift[3 * pos + 1] = slot_shift[3 * k + 1];
    send_shift[3 * pos + 2] = slot_shift[3 * k + 2];
  }
}

__global__ void k_slot_counts(const long long *__restrict__ tile_off, int ntiles, int nslots, long long *__restrict__ out) {
  const int k = threadIdx.x;
  if (k <= nslots) out[k] = tile_off[(size_t) k * ntiles];          // k == nslots: the total (tile_off has nslots*ntiles + 1 entries)
}

// ---- NCCL, resolved at run time -------------------------------------------------------------------------------------
// minimal declarations (nccl.h is not needed to build the library); values as in NCCL 2.x
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { kNcclInt8 = 0, kNcclFloat64 = 8 };
enum { kNcclSum = 0, kNcclMax = 2 };

struct NcclApi {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};

NcclApi &nccl() {
  static NcclApi api;
  if (api.lib) return api;
  const char *names[] = {getenv("ANNP_B200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  for (const char *nm : names) {
    if (!nm || !*nm) continue;
    api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (api.lib) break;
  }
  if (!api.lib) return api;
  bool all = true;
  auto sym = [&](const char *n) { void *p = dlsym(api.lib, n); all = all && p; return p; };
  api.GetUniqueId = (decltype(api.GetUniqueId)) sym("ncclGetUniqueId");
  api.CommInitRank = (decltype(api.CommInitRank)) sym("ncclCommInitRank");
  api.CommDestroy = (decltype(api.CommDestroy)) sym("ncclCommDestroy");
  api.GroupStart = (decltype(api.GroupStart)) sym("ncclGroupStart");
  api.GroupEnd = (decltype(api.GroupEnd)) sym("ncclGroupEnd");
  api.Send = (decltype(api.Send)) sym("ncclSend");
  api.Recv = (decltype(api.Recv)) sym("ncclRecv");
  api.AllReduce = (decltype(api.AllReduce)) sym("ncclAllReduce");
  api.GetErrorString = (decltype(api.GetErrorString)) sym("ncclGetErrorString");
  api.ok = all;
  return api;
}

int nccl_fail(annp_b200_handle h, ncclResult_t r, const char *where) {
  const NcclApi &n = nccl();
  return fail(h, ANNP_B200_ECOMM, std::string(where) + ": " + (n.GetErrorString ? n.GetErrorString(r) : "NCCL error"));
}
#define NK(call)                                                 \
  do {                                                           \
    ncclResult_t r__ = (call);                                   \
    if (r__ != 0) return nccl_fail(h, r__, #call);               \
  } while (0)

// grouped exchange of 3 doubles per atom: rank r gets sendbuf[soff[r] .. + scnt[r]) and delivers rcnt[r] atoms at recvbuf + roff[r];
// the part a rank keeps for itself (periodic images inside one brick) is a device copy
int exchange3(annp_b200_handle h, const double *sendbuf, const std::vector<int> &scnt, double *recvbuf, const std::vector<int> &rcnt,
              cudaStream_t s) {
  const NcclApi &n = nccl();
  const int P = h->comm_size, me = h->comm_rank;
  size_t soff = 0, roff = 0;
  bool grouped = false;
  for (int r = 0; r < P; r++) {
    const size_t sc = 3 * (size_t) scnt[r], rc = 3 * (size_t) rcnt[r];
    if (r == me) {
      if (sc != rc) return fail(h, ANNP_B200_EINVAL, "halo: self send and receive counts differ");
      if (sc) CK(cudaMemcpyAsync(recvbuf + roff, sendbuf + soff, sizeof(double) * sc, cudaMemcpyDeviceToDevice, s));
    } else if (sc || rc) {
      if (!h->nccl_comm) return fail(h, ANNP_B200_ESTATE, "halo exchange with other ranks needs annp_b200_comm_init");
      if (!grouped) { NK(n.GroupStart()); grouped = true; }
      if (sc) NK(n.Send(sendbuf + soff, sc, kNcclFloat64, r, (ncclComm_t) h->nccl_comm, s));
      if (rc) NK(n.Recv(recvbuf + roff, rc, kNcclFloat64, r, (ncclComm_t) h->nccl_comm, s));
    }
    soff += sc;
    roff += rc;
  }
  if (grouped) NK(n.GroupEnd());
  return ANNP_B200_OK;
}

}    // namespace

extern "C" {

int annp_b200_send_lists_count(annp_b200_handle h, int nlocal, const double *d_x, const double *lo, const double *hi, double cutghost,
                               int nslots, const int *slot_dir, int *slot_counts, void *stream) {
  if (!h) return ANNP_B200_EINVAL;
  if (nlocal < 0 || nslots < 0 || nslots > 26 || (nlocal > 0 && !d_x) || !lo || !hi || !(cutghost > 0.0) || (nslots > 0 && (!slot_dir || !slot_counts)))
    return fail(h, ANNP_B200_EINVAL, "bad send-list arguments");
  CK(cudaSetDevice(h->device));
  cudaStream_t s = (cudaStream_t) stream;
  SendGeom g;
  for (int d = 0; d < 3; d++) { g.lo[d] = lo[d]; g.hi[d] = hi[d]; h->sl_lo[d] = lo[d]; h->sl_hi[d] = hi[d]; }
  g.cut = cutghost; g.nslots = nslots;
  for (int k = 0; k < 26; k++)
    for (int d = 0; d < 3; d++) {
      const int v = k < nslots ? slot_dir[3 * k + d] : 0;
      if (v < -1 || v > 1) return fail(h, ANNP_B200_EINVAL, "slot directions must be -1, 0 or 1");
      g.dir[k][d] = (signed char) v;
      h->sl_dir[k][d] = v;
    }
  const int ntiles = std::max(1, (nlocal + kSlTile - 1) / kSlTile);
  h->sl_nlocal = nlocal; h->sl_nslots = nslots; h->sl_ntiles = ntiles; h->sl_x = d_x; h->sl_cut = cutghost;
  const int ncnt = std::max(nslots, 1) * ntiles;
  CK(h->d_sl_tile_cnt.reserve(sizeof(int) * (size_t) ncnt));
  CK(h->d_sl_tile_off.reserve(sizeof(long long) * ((size_t) ncnt + 1 + 32)));
  CK(h->d_sl_tile_sum.reserve(sizeof(long long) * ((size_t) ncnt / 1024 + 2)));
  CK(cudaMemsetAsync(h->d_sl_tile_cnt.p, 0, sizeof(int) * (size_t) ncnt, s));
  if (nslots > 0 && nlocal > 0) k_send_count<<<ntiles, kSlTile, 0, s>>>(d_x, nlocal, g, ntiles, h->d_sl_tile_cnt.as<int>());
  aux_exclusive_scan(h->d_sl_tile_cnt.as<int>(), h->d_sl_tile_off.as<long long>(), ncnt, h->d_sl_tile_sum.as<long long>(), s);
  h->launches += 4;
  // the one size read of a re-neighbouring: where each slot starts (27 numbers)
  long long *d_starts = h->d_sl_tile_off.as<long long>() + ncnt + 1;
  long long starts[27];
  if (nslots > 0) {
    // slot k starts at tile_off[k * ntiles]; the total sits at tile_off[ncnt] = entry nslots * ntiles when nslots >= 1
    k_slot_counts<<<1, 32, 0, s>>>(h->d_sl_tile_off.as<long long>(), ntiles, nslots, d_starts);
    h->launches += 1;
    CK(cudaMemcpyAsync(starts, d_starts, sizeof(long long) * (size_t) (nslots + 1), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    for (int k = 0; k < nslots; k++) slot_counts[k] = (int) (starts[k + 1] - starts[k]);
  }
  CK(cudaGetLastError());
  return ANNP_B200_OK;
}

int annp_b200_send_lists_fill(annp_b200_handle h, const double *slot_shift, int *d_send_index, double *d_send_shift, void *stream) {
  if (!h) return ANNP_B200_EINVAL;
  if (h->sl_nslots > 0 && (!slot_shift || !d_send_index || !d_send_shift)) return fail(h, ANNP_B200_EINVAL, "bad send-list arguments");
  if (h->sl_nslots == 0 || h->sl_nlocal == 0) return ANNP_B200_OK;
  CK(cudaSetDevice(h->device));
  cudaStream_t s = (cudaStream_t) stream;
  SendGeom g;
  for (int d = 0; d < 3; d++) { g.lo[d] = h->sl_lo[d]; g.hi[d] = h->sl_hi[d]; }
  g.cut = h->sl_cut; g.nslots = h->sl_nslots;
  for (int k = 0; k < 26; k++) for (int d = 0; d < 3; d++) g.dir[k][d] = (signed char) h->sl_dir[k][d];
  // per-slot shifts (26 x 3 doubles) staged in the handle's small scratch buffer
  CK(h->d_small.reserve(sizeof(double) * 26 * 3 + 64));
  CK(cudaMemcpyAsync(h->d_small.p, slot_shift, sizeof(double) * 3 * (size_t) h->sl_nslots, cudaMemcpyHostToDevice, s));
  k_send_fill<<<h->sl_ntiles, kSlTile, 0, s>>>(h->sl_x, h->sl_nlocal, g, h->sl_ntiles, h->d_sl_tile_off.as<long long>(),
                                              h->d_small.as<double>(), d_send_index, d_send_shift);
  h->launches += 1;
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(s));        // slot_shift is the caller's (pageable) array: the copy must have left it
  return ANNP_B200_OK;
}

int annp_b200_comm_unique_id(char *id128) {
  if (!id128) return ANNP_B200_EINVAL;
  NcclApi &n = nccl();
  if (!n.ok) return ANNP_B200_ECOMM;
  ncclUniqueId id;
  if (n.GetUniqueId(&id) != 0) return ANNP_B200_ECOMM;
  memcpy(id128, id.internal, 128);
  return ANNP_B200_OK;
}

int annp_b200_comm_init(annp_b200_handle h, int nranks, int rank, const char *id128) {
  if (!h) return ANNP_B200_EINVAL;
  if (nranks < 1 || rank < 0 || rank >= nranks) return fail(h, ANNP_B200_EINVAL, "bad communicator shape");
  CK(cudaSetDevice(h->device));
  h->comm_size = nranks; h->comm_rank = rank;
  h->peer_send.assign((size_t) nranks, 0);
  h->peer_recv.assign((size_t) nranks, 0);
  if (nranks == 1) return ANNP_B200_OK;             // nothing to talk to: the exchange is a device copy
  if (!id128) return fail(h, ANNP_B200_EINVAL, "unique id missing");
  NcclApi &n = nccl();
  if (!n.ok) return fail(h, ANNP_B200_ECOMM, "libnccl.so.2 not found in this process (set ANNP_B200_NCCL_LIB)");
  if (h->nccl_comm) { n.CommDestroy((ncclComm_t) h->nccl_comm); h->nccl_comm = nullptr; }
  ncclUniqueId id;
  memcpy(id.internal, id128, 128);
  ncclComm_t c = nullptr;
  NK(n.CommInitRank(&c, nranks, id, rank));
  h->nccl_comm = c;
  return ANNP_B200_OK;
}

void annp_b200_comm_destroy(annp_b200_handle h) {
  if (!h || !h->nccl_comm) return;
  cudaSetDevice(h->device);
  nccl().CommDestroy((ncclComm_t) h->nccl_comm);
  h->nccl_comm = nullptr;
}

int annp_b200_set_halo_peers(annp_b200_handle h, int nranks, const int *send_counts, const int *recv_counts) {
  if (!h || !send_counts || !recv_counts) return ANNP_B200_EINVAL;
  if (nranks != h->comm_size) return fail(h, ANNP_B200_ESTATE, "set_halo_peers: rank count differs from annp_b200_comm_init");
  long long ns = 0, nr = 0;
  for (int r = 0; r < nranks; r++) {
    if (send_counts[r] < 0 || recv_counts[r] < 0) return fail(h, ANNP_B200_EINVAL, "negative halo count");
    ns += send_counts[r]; nr += recv_counts[r];
  }
  if (ns != h->g_nghost) return fail(h, ANNP_B200_ESTATE, "set_halo_peers: send counts do not add up to the send list of annp_b200_set_halo");
  h->peer_send.assign(send_counts, send_counts + nranks);
  h->peer_recv.assign(recv_counts, recv_counts + nranks);
  h->g_nghost_recv = (int) nr;
  CK(cudaSetDevice(h->device));
  CK(h->d_sendbuf.reserve(sizeof(double) * 3 * (size_t) std::max<long long>(ns, 1)));
  CK(h->d_recvbuf.reserve(sizeof(double) * 3 * (size_t) std::max<long long>(ns, 1)));
  (void) nr;
  return ANNP_B200_OK;
}

int annp_b200_halo_forward(annp_b200_handle h, double *d_x, void *stream) {
  if (!h || !d_x) return ANNP_B200_EINVAL;
  cudaStream_t s = (cudaStream_t) stream;
  if (h->peer_on) {
    // peer scatter: this rank's accumulators must be clear before any peer's force kernel of this step can run, and a
    // peer's kernel runs only after it has received this rank's positions from the exchange below
    CK(cudaMemsetAsync(h->d_facc.p, 0, sizeof(long long) * 3 * (size_t) std::max(h->g_nlocal + h->g_nghost_recv, 1), s));
    h->peer_zeroed = true;
  }
  double *ghost = d_x + 3 * (size_t) h->g_nlocal;
  if (h->comm_size == 1) {                          // the receiver is this rank's own ghost block
    if (h->g_nghost > 0) { aux_halo_pack(h->g_nghost, h->g_owner, h->g_shift, d_x, ghost, s); h->launches += 1; }
    return ANNP_B200_OK;
  }
  if ((int) h->peer_send.size() != h->comm_size) return fail(h, ANNP_B200_ESTATE, "halo_forward before set_halo_peers");
  if (h->g_nghost == 0 && std::all_of(h->peer_recv.begin(), h->peer_recv.end(), [](int c) { return c == 0; })) return ANNP_B200_OK;
  aux_halo_pack(h->g_nghost, h->g_owner, h->g_shift, d_x, h->d_sendbuf.as<double>(), s);
  h->launches += 1;
  return exchange3(h, h->d_sendbuf.as<double>(), h->peer_send, ghost, h->peer_recv, s);
}

int annp_b200_halo_reverse(annp_b200_handle h, double *d_f, void *stream) {
  if (!h || !d_f) return ANNP_B200_EINVAL;
  if (h->peer_on && h->scatter_fixed) return ANNP_B200_OK;      // peer scatter: the owners already hold the ghost contributions
  cudaStream_t s = (cudaStream_t) stream;
  if (h->comm_size > 1 && (int) h->peer_send.size() != h->comm_size) return fail(h, ANNP_B200_ESTATE, "halo_reverse before set_halo_peers");
  if (h->g_nghost == 0 && std::all_of(h->peer_recv.begin(), h->peer_recv.end(), [](int c) { return c == 0; })) return ANNP_B200_OK;
  const double *ghost = d_f + 3 * (size_t) h->g_nlocal;
  const double *src = ghost;
  if (h->comm_size > 1) {                           // ghost forces travel back along the same lists: send what was received
    int rc = exchange3(h, ghost, h->peer_recv, h->d_recvbuf.as<double>(), h->peer_send, s);
    if (rc) return rc;
    src = h->d_recvbuf.as<double>();
  }
  aux_halo_unpack_add(h->g_nlocal, h->d_goff.as<long long>(), h->d_glist.as<int>(), src, d_f, s);
  h->launches += 1;
  return ANNP_B200_OK;
}

// ---- peer scatter: the reverse halo fused into the force kernel -------------------------------------------------------
// Every rank exports its fixed-point accumulator array (cudaIpc), maps the others' and tells the force kernel, for every ghost,
// which rank owns it and under which local index.  The kernel then adds the force on a ghost straight into the owner's
// accumulator with system-scope 64-bit integer atomics over NVLink; the per-step reverse exchange (grouped send/recv of the
// ghost forces + ordered add) is replaced by one barrier (a one-element all-reduce) between the kernels and the conversion
// of the accumulators.  Integer addition commutes, so the result is bit-identical to the exchange path and reproducible.
// Ordering: a rank zeroes its accumulators BEFORE its forward exchange; a peer's kernel starts after that peer received this
// rank's positions, i.e. after the zeroing; the barrier orders every peer's kernel before this rank reads its accumulators.
int annp_b200_peer_export(annp_b200_handle h, int nall, char *ipc64) {
  if (!h || !ipc64 || nall < 0) return ANNP_B200_EINVAL;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  CK(cudaSetDevice(h->device));
  CK(h->d_facc.reserve(sizeof(long long) * 3 * (size_t) std::max(nall, 1)));
  cudaIpcMemHandle_t mh;
  CK(cudaIpcGetMemHandle(&mh, h->d_facc.p));
  memcpy(ipc64, &mh, 64);
  return ANNP_B200_OK;
}

int annp_b200_peer_close(annp_b200_handle h) {
  if (!h) return ANNP_B200_EINVAL;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  for (void *&p : h->peer_ptr)
    if (p) { cudaIpcCloseMemHandle(p); p = nullptr; }
  h->peer_ptr.clear();
  h->peer_key.clear();
  h->peer_on = false;
  return ANNP_B200_OK;
}

int annp_b200_peer_open(annp_b200_handle h, int nranks, const char *ipc64_all, int nghost, const int *d_ghost_rank,
                        const int *d_ghost_index, void *stream) {
  if (!h || !ipc64_all || nghost < 0 || (nghost > 0 && (!d_ghost_rank || !d_ghost_index))) return ANNP_B200_EINVAL;
  if (nranks != h->comm_size) return fail(h, ANNP_B200_ESTATE, "peer_open: rank count differs from annp_b200_comm_init");
  if (nranks > 1 && !h->nccl_comm) return fail(h, ANNP_B200_ESTATE, "peer scatter needs annp_b200_comm_init (step barrier)");
  CK(cudaSetDevice(h->device));
  h->peer_ptr.resize((size_t) nranks, nullptr);
  h->peer_key.resize((size_t) nranks);
  std::vector<long long *> table((size_t) nranks, nullptr);
  for (int r = 0; r < nranks; r++) {
    if (r == h->comm_rank) { table[r] = h->d_facc.as<long long>(); continue; }
    const std::string key(ipc64_all + 64 * (size_t) r, 64);
    if (!h->peer_ptr[r] || h->peer_key[r] != key) {        // the peer re-allocated its array (or first use): (re)map it
      if (h->peer_ptr[r]) { CK(cudaDeviceSynchronize()); CK(cudaIpcCloseMemHandle(h->peer_ptr[r])); h->peer_ptr[r] = nullptr; }
      cudaIpcMemHandle_t mh;
      memcpy(&mh, key.data(), 64);
      void *p = nullptr;
      CK(cudaIpcOpenMemHandle(&p, mh, cudaIpcMemLazyEnablePeerAccess));
      h->peer_ptr[r] = p;
      h->peer_key[r] = key;
    }
    table[r] = (long long *) h->peer_ptr[r];
  }
  CK(h->d_peer_table.reserve(sizeof(long long *) * (size_t) nranks));
  CK(h->d_peer_sync.reserve(sizeof(double)));
  cudaStream_t s = (cudaStream_t) stream;
  CK(cudaMemcpyAsync(h->d_peer_table.p, table.data(), sizeof(long long *) * (size_t) nranks, cudaMemcpyHostToDevice, s));
  CK(cudaMemsetAsync(h->d_peer_sync.p, 0, sizeof(double), s));
  CK(cudaStreamSynchronize(s));          // `table` is a host temporary
  h->peer_ghost_rank = d_ghost_rank;
  h->peer_ghost_index = d_ghost_index;
  h->peer_on = nranks > 1;
  return ANNP_B200_OK;
}

// the barrier of the peer-scatter step: every rank's force kernel has finished adding (called by the per-step sequence)
int annp_peer_barrier(annp_b200_handle h, cudaStream_t s) {
  NK(nccl().AllReduce(h->d_peer_sync.p, h->d_peer_sync.p, 1, kNcclFloat64, kNcclSum, (ncclComm_t) h->nccl_comm, s));
  return ANNP_B200_OK;
}

int annp_b200_allreduce_sum(annp_b200_handle h, double *d_buf, int n, void *stream) {
  if (!h || (n > 0 && !d_buf)) return ANNP_B200_EINVAL;
  if (h->comm_size == 1 || n <= 0) return ANNP_B200_OK;
  if (!h->nccl_comm) return fail(h, ANNP_B200_ESTATE, "allreduce needs annp_b200_comm_init");
  NK(nccl().AllReduce(d_buf, d_buf, (size_t) n, kNcclFloat64, kNcclSum, (ncclComm_t) h->nccl_comm, (cudaStream_t) stream));
  return ANNP_B200_OK;
}

}    // extern "C"

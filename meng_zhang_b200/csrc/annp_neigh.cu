// annp_neigh.cu -- device-side full neighbour list (replaces the reference's GPU_NEIGH path,
// annp_gpu_compute_n: lib/lal_annp_ext.cpp:98-108, and the host list upload of reset_nbors).
//
// Cell binning with cell edge >= cutneigh: count pass -> exclusive scan -> fill pass -> per-row rank
// sort, so rows come out sorted by neighbour index and the list is bit-reproducible even though the
// cell fill uses integer atomics.  Centres are atoms 0..nlocal-1, partners all nall atoms.
#include "annp_device.cuh"

void aux_exclusive_scan(const int *cnt, long long *off, int n, long long *tile_sum, cudaStream_t s);

namespace {

struct Grid {
  double lo[3];
  double inv_cell[3];
  int n[3];
};

__device__ __forceinline__ int cell_coord(double x, double lo, double inv, int n) {
  int c = (int) floor((x - lo) * inv);
  return min(max(c, 0), n - 1);
}

__global__ void k_cell_count(const double *__restrict__ x, int nall, Grid g, int *__restrict__ cell_of, int *__restrict__ cnt) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nall; i += gridDim.x * blockDim.x) {
    const int cx = cell_coord(x[3 * (size_t) i], g.lo[0], g.inv_cell[0], g.n[0]);
    const int cy = cell_coord(x[3 * (size_t) i + 1], g.lo[1], g.inv_cell[1], g.n[1]);
    const int cz = cell_coord(x[3 * (size_t) i + 2], g.lo[2], g.inv_cell[2], g.n[2]);
    const int c = (cz * g.n[1] + cy) * g.n[0] + cx;
    cell_of[i] = c;
    atomicAdd(&cnt[c], 1);
  }
}

__global__ void k_cell_fill(const int *__restrict__ cell_of, int nall, const long long *__restrict__ cell_off,
                            int *__restrict__ cursor, int *__restrict__ cell_atoms) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nall; i += gridDim.x * blockDim.x) {
    const int c = cell_of[i];
    cell_atoms[cell_off[c] + atomicAdd(&cursor[c], 1)] = i;
  }
}

// one warp per centre atom; MODE 0 counts, MODE 1 writes (unsorted) rows
template <int MODE>
__global__ void k_rows(const double *__restrict__ x, int nlocal, Grid g, const long long *__restrict__ cell_off,
                       const int *__restrict__ cell_atoms, double cutsq, int *__restrict__ row_cnt,
                       const long long *__restrict__ row_off, int *__restrict__ rows) {
  const int lane = threadIdx.x & 31;
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nw = (gridDim.x * blockDim.x) >> 5;
  for (int i = wid; i < nlocal; i += nw) {
    const double xi = x[3 * (size_t) i], yi = x[3 * (size_t) i + 1], zi = x[3 * (size_t) i + 2];
    const int cx = cell_coord(xi, g.lo[0], g.inv_cell[0], g.n[0]);
    const int cy = cell_coord(yi, g.lo[1], g.inv_cell[1], g.n[1]);
    const int cz = cell_coord(zi, g.lo[2], g.inv_cell[2], g.n[2]);
    int count = 0;
    const long long out0 = (MODE == 1) ? row_off[i] : 0;
    for (int dz = -1; dz <= 1; dz++) {
      const int z = cz + dz;
      if (z < 0 || z >= g.n[2]) continue;
      for (int dy = -1; dy <= 1; dy++) {
        const int y = cy + dy;
        if (y < 0 || y >= g.n[1]) continue;
        // the three x-cells of this (y,z) line are contiguous in cell order
        const int x0 = max(cx - 1, 0), x1 = min(cx + 1, g.n[0] - 1);
        const long long b = cell_off[(z * g.n[1] + y) * g.n[0] + x0];
        const long long e = cell_off[(z * g.n[1] + y) * g.n[0] + x1 + 1];
        for (long long m0 = b; m0 < e; m0 += 32) {
          const long long m = m0 + lane;
          bool hit = false;
          int j = -1;
          if (m < e) {
            j = cell_atoms[m];
            const double dx = xi - x[3 * (size_t) j], dyy = yi - x[3 * (size_t) j + 1], dzz = zi - x[3 * (size_t) j + 2];
            hit = (j != i) && (dx * dx + dyy * dyy + dzz * dzz < cutsq);
          }
          const unsigned mask = __ballot_sync(0xffffffffu, hit);
          if (MODE == 1 && hit) rows[out0 + count + __popc(mask & ((1u << lane) - 1u))] = j;
          count += __popc(mask);
        }
      }
    }
    if (MODE == 0 && lane == 0) row_cnt[i] = count;
  }
}

__global__ void k_row_sort(const long long *__restrict__ off, const int *__restrict__ in, int *__restrict__ out, int nseg) {
  const int lane = threadIdx.x & 31;
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nw = (gridDim.x * blockDim.x) >> 5;
  for (int s = wid; s < nseg; s += nw) {
    const long long b = off[s];
    const int n = (int) (off[s + 1] - b);
    for (int e = lane; e < n; e += 32) {
      const int v = in[b + e];
      int rank = 0;
      for (int q = 0; q < n; q++) rank += (in[b + q] < v);
      out[b + rank] = v;
    }
  }
}

__global__ void k_max_int(const int *__restrict__ v, int n, int *__restrict__ out) {
  int m = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) m = max(m, v[i]);
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

inline int grid_for(long long n, int threads, int cap = 148 * 16) {
  long long g = (n + threads - 1) / threads;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int) g;
}

}    // namespace

// Phase A: bin atoms, count row lengths, scan -> row_off (device) ; returns nothing, caller reads the
// total from row_off[nlocal] (one synchronising 8-byte copy) to size the row arrays.
// scratch: cell_of [nall], cell_cnt [ncells] , cell_off [ncells+1], cell_atoms [nall], tile_sum
struct NeighScratch {
  int *cell_of, *cell_cnt, *cell_atoms, *row_cnt;
  long long *cell_off, *tile_sum;
};

void neigh_grid(const double *lo, const double *hi, double cutneigh, int *n_out, double *inv_out, long long *ncells) {
  long long tot = 1;
  for (int d = 0; d < 3; d++) {
    double len = hi[d] - lo[d];
    if (len < 1e-9) len = 1e-9;
    int n = (int) (len / cutneigh);
    if (n < 1) n = 1;
    n_out[d] = n;
    inv_out[d] = (double) n / len;
    tot *= n;
  }
  *ncells = tot;
}

void neigh_count(const double *d_x, int nlocal, int nall, const double *lo, const int *n, const double *inv, double cutneigh,
                 NeighScratch sc, long long *d_row_off, int *d_maxrow, cudaStream_t s) {
  Grid g;
  for (int d = 0; d < 3; d++) { g.lo[d] = lo[d]; g.inv_cell[d] = inv[d]; g.n[d] = n[d]; }
  const long long ncells = (long long) n[0] * n[1] * n[2];
  cudaMemsetAsync(sc.cell_cnt, 0, sizeof(int) * (size_t) ncells, s);
  k_cell_count<<<grid_for(nall, 256), 256, 0, s>>>(d_x, nall, g, sc.cell_of, sc.cell_cnt);
  aux_exclusive_scan(sc.cell_cnt, sc.cell_off, (int) ncells, sc.tile_sum, s);
  cudaMemsetAsync(sc.cell_cnt, 0, sizeof(int) * (size_t) ncells, s);
  k_cell_fill<<<grid_for(nall, 256), 256, 0, s>>>(sc.cell_of, nall, sc.cell_off, sc.cell_cnt, sc.cell_atoms);
  k_rows<0><<<grid_for((long long) nlocal * 32, 256, 148 * 32), 256, 0, s>>>(d_x, nlocal, g, sc.cell_off, sc.cell_atoms, cutneigh * cutneigh,
                                                                               sc.row_cnt, nullptr, nullptr);
  aux_exclusive_scan(sc.row_cnt, d_row_off, nlocal, sc.tile_sum, s);
  cudaMemsetAsync(d_maxrow, 0, sizeof(int), s);
  k_max_int<<<grid_for(nlocal, 256), 256, 0, s>>>(sc.row_cnt, nlocal, d_maxrow);
}

void neigh_fill(const double *d_x, int nlocal, const double *lo, const int *n, const double *inv, double cutneigh,
                NeighScratch sc, const long long *d_row_off, int *d_rows_tmp, int *d_rows, cudaStream_t s) {
  Grid g;
  for (int d = 0; d < 3; d++) { g.lo[d] = lo[d]; g.inv_cell[d] = inv[d]; g.n[d] = n[d]; }
  k_rows<1><<<grid_for((long long) nlocal * 32, 256, 148 * 32), 256, 0, s>>>(d_x, nlocal, g, sc.cell_off, sc.cell_atoms, cutneigh * cutneigh,
                                                                               nullptr, d_row_off, d_rows_tmp);
  k_row_sort<<<grid_for((long long) nlocal * 32, 256, 148 * 32), 256, 0, s>>>(d_row_off, d_rows_tmp, d_rows, nlocal);
}

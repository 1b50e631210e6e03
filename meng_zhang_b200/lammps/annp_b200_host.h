/* -*- c++ -*- ----------------------------------------------------------
   Host-side buffers shared by the two LAMMPS pair styles on libannp_b200.so
   (pair_annp_b200.cpp, pair_anna_adp_b200.cpp).

   The reference's GPU styles hand LAMMPS' own pageable arrays to the LAL layer, which copies them
   into page-locked staging vectors on the host every step (ANNP::compute -> atom->cast_x_data /
   add_x_data, fe_v2/lib/lal_annp.cpp:310-312; forces come back through a page-locked answer
   vector and a scalar host loop, :336-347).  On a B200 the force kernel of 524 288 atoms takes
   24 ms and 16 MB cross PCIe in each direction, so those host copies would be a tenth of the
   step.  Here LAMMPS' arrays are page-locked IN PLACE (cudaHostRegister through the C ABI) and
   the device reads / writes them directly:
     * atom->x  [nmax][3]  registered, re-registered whenever LAMMPS reallocates it (atom->nmax growth)
     * atom->f  [nmax][3]  registered; when this style is the only pair style of the run the
       device writes the forces straight into it (LAMMPS has zeroed f before Pair::compute, and
       the reference assigns f as well, lal_annp.cpp:345-347).  Under pair hybrid / overlay the
       forces go to a page-locked staging array and are ADDED by a vectorisable loop.
     * atom->type is uploaded only at re-neighbouring (atoms keep their slots in between).
   Registration failures are not fatal: the copy is then staged by the driver as before.
------------------------------------------------------------------------- */

#ifndef LMP_ANNP_B200_HOST_H
#define LMP_ANNP_B200_HOST_H

#include "annp_b200.h"

#include <cstddef>
#include <cstdlib>
#include <cstring>

namespace ANNP_B200_NS {

// one caller-owned array that is kept page-locked while it stays where it is
struct RegisteredArray {
  void *ptr = nullptr;
  size_t bytes = 0;
  bool locked = false;
  // page-lock [p, p + n) unless exactly this range is registered already; the previous range is released first
  // (LAMMPS' Memory::grow may have moved the array: the old range is stale either way)
  void track(void *p, size_t n)
  {
    if (p == ptr && n <= bytes) return;
    release();
    ptr = p;
    bytes = n;
    locked = p && n && pagelock_enabled() && annp_b200_host_register(p, n) == ANNP_B200_OK;
  }
  // ANNP_B200_PAGELOCK=0 leaves LAMMPS' arrays pageable (every copy is then staged by the CUDA driver)
  static bool pagelock_enabled()
  {
    const char *s = getenv("ANNP_B200_PAGELOCK");
    return !(s && s[0] == '0');
  }
  void release()
  {
    if (locked) annp_b200_host_unregister(ptr);
    ptr = nullptr;
    bytes = 0;
    locked = false;
  }
};

// page-locked staging array owned by the style (falls back to malloc if page-locking fails)
struct StagingArray {
  double *p = nullptr;
  size_t n = 0;
  bool pinned = false;
  double *reserve(size_t count)
  {
    if (count <= n) return p;
    release();
    const size_t want = count + count / 8 + 16;
    p = (double *) annp_b200_host_alloc(sizeof(double) * want);
    pinned = p != nullptr;
    if (!p) p = (double *) malloc(sizeof(double) * want);
    n = p ? want : 0;
    return p;
  }
  void release()
  {
    if (p && pinned) annp_b200_host_free(p);
    else free(p);
    p = nullptr;
    n = 0;
    pinned = false;
  }
};

// dst[i] += src[i]: contiguous, no aliasing -> the compiler vectorises it
inline void add_into(double *__restrict__ dst, const double *__restrict__ src, size_t n)
{
  for (size_t i = 0; i < n; i++) dst[i] += src[i];
}

// axis-aligned bounds of x[0..n) (device-neighbour mode: the cell grid must cover every atom the rank holds)
inline void bounds_of(const double *x, int n, double *lo, double *hi)
{
  for (int d = 0; d < 3; d++) { lo[d] = 0.0; hi[d] = 0.0; }
  if (n <= 0) return;
  for (int d = 0; d < 3; d++) lo[d] = hi[d] = x[d];
  for (int i = 1; i < n; i++)
    for (int d = 0; d < 3; d++) {
      const double v = x[3 * (size_t) i + d];
      if (v < lo[d]) lo[d] = v;
      if (v > hi[d]) hi[d] = v;
    }
  for (int d = 0; d < 3; d++) { lo[d] -= 1.0e-6; hi[d] += 1.0e-6; }
}

struct HostBuffers {
  RegisteredArray x, f;
  StagingArray fbuf, ebuf, vbuf;
  ~HostBuffers() { clear(); }
  void clear()
  {
    x.release(); f.release();
    fbuf.release(); ebuf.release(); vbuf.release();
  }
  double bytes() const { return (double) (fbuf.n + ebuf.n + vbuf.n) * sizeof(double); }
};

// ANNP_B200_FORCE_ADD=1 : always stage the forces and ADD them to atom->f, even as the only pair style (decks in which
// something else has written forces before Pair::compute, e.g. a fix with a pre_force method)
inline bool force_add_requested()
{
  const char *s = getenv("ANNP_B200_FORCE_ADD");
  return s && s[0] == '1';
}

// ANNP_B200_NEIGH=device : build the neighbour list on the GPU (the reference's `package gpu ... neigh yes`)
inline bool device_neigh_requested()
{
  const char *s = getenv("ANNP_B200_NEIGH");
  return s && (!strcmp(s, "device") || !strcmp(s, "yes"));
}

}    // namespace ANNP_B200_NS

#endif

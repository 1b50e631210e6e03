// annp_force_generic.cu -- the padded instantiations of the fused ANNP force kernel: any potential with npsf <= 16 radial
// and ntsf <= 24 angular Chebyshev components that has no exact instantiation in annp_force.cu runs here, its descriptor
// padded with zero-scale / zero-weight components (exact zeros in energy and forces).  Kept in its own translation unit so
// the two files compile in parallel.
#include "annp_force_kernel.cuh"

template <bool FIXED>
static annp_force_kernel_t pick_generic(int npsf, int ntsf, int variant) {
  const bool anna = variant == ANNP_B200_VARIANT_ANNA_ADP;
  if (ntsf != 24) return nullptr;
  if (npsf == 8) return anna ? annp_force_kernel<8, 24, 1, FIXED> : annp_force_kernel<8, 24, 0, FIXED>;
  if (npsf == 16) return anna ? annp_force_kernel<16, 24, 1, FIXED> : annp_force_kernel<16, 24, 0, FIXED>;
  return nullptr;
}

annp_force_kernel_t annp_force_generic_kernel(int npsf, int ntsf, int variant, bool fixed) {
  return fixed ? pick_generic<true>(npsf, ntsf, variant) : pick_generic<false>(npsf, ntsf, variant);
}

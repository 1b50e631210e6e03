// annp_force.cu -- instantiations of the fused ANNP force kernel (annp_force_kernel.cuh) for the descriptor shapes of the
// potentials the reference ships, and the shape dispatcher.
//
// The reference's CPU style takes any `TL HL nodes nsf npsf ntsf` line (fe_v2/src/pair_annp.cpp:367-389).  The kernel keeps
// its per-lane descriptor state in registers, so (npsf, ntsf) are template parameters:
//   * exact instantiations for the shipped files: (9, 19) fe / fe_v2 / anna, (8, 20), (4, 6) (unit-test size);
//   * every other shape runs on a PADDED instantiation (annp_force_generic.cu): the descriptor is extended with
//     components whose normalisation scale and first-layer weights are zero, which contribute exact zeros to the energy
//     and to every derivative (annp_capi.cu: pad_shape).  (8, 24) covers nsf <= 32, (16, 24) the rest.
#include "annp_force_kernel.cuh"

annp_force_kernel_t annp_force_generic_kernel(int npsf, int ntsf, int variant, bool fixed);

template <bool FIXED>
static annp_force_kernel_t pick_exact(int npsf, int ntsf, int variant) {
  if (variant == ANNP_B200_VARIANT_ANNA_ADP) {
    if (npsf == 9 && ntsf == 19) return annp_force_kernel<9, 19, 1, FIXED>;  // fe_adp_potential_2310.anna
    if (npsf == 4 && ntsf == 6) return annp_force_kernel<4, 6, 1, FIXED>;
    return nullptr;
  }
  if (npsf == 9 && ntsf == 19) return annp_force_kernel<9, 19, 0, FIXED>;     // fe / fe_v2 potential
  if (npsf == 8 && ntsf == 20) return annp_force_kernel<8, 20, 0, FIXED>;
  if (npsf == 4 && ntsf == 6) return annp_force_kernel<4, 6, 0, FIXED>;       // small set used by unit tests
  return nullptr;
}

static annp_force_kernel_t pick_kernel(int npsf, int ntsf, int variant, bool fixed) {
  annp_force_kernel_t k = fixed ? pick_exact<true>(npsf, ntsf, variant) : pick_exact<false>(npsf, ntsf, variant);
  return k ? k : annp_force_generic_kernel(npsf, ntsf, variant, fixed);
}

// The kernel shape (npsf_k >= npsf, ntsf_k >= ntsf) a potential of shape (npsf, ntsf) runs on; false if none covers it.
bool annp_force_kernel_shape(int npsf, int ntsf, int variant, int *npsf_k, int *ntsf_k) {
  if (pick_exact<true>(npsf, ntsf, variant)) { *npsf_k = npsf; *ntsf_k = ntsf; return true; }
  if (npsf < 1 || ntsf < 1 || ntsf > 24 || npsf > 16) return false;
  *ntsf_k = 24;
  *npsf_k = npsf <= 8 ? 8 : 16;
  return true;
}

size_t annp_force_smem_bytes(const DevParams &hp, int capacity) {
  return annp_force_smem_bytes_shape(hp.npsf, hp.ntsf, hp.nlayers, hp.nnod, capacity);
}

// Launch on `stream`.  max_blocks <= 0: one full wave of resident blocks (or fewer when there is less work).
cudaError_t annp_force_launch(const ForceArgs &args, const DevParams &hp, int num_sms, cudaStream_t stream, int max_blocks) {
  annp_force_kernel_t k = pick_kernel(hp.npsf, hp.ntsf, hp.variant, args.facc != nullptr);
  if (!k) return cudaErrorInvalidValue;
  return annp_force_launch_kernel(k, args, hp, num_sms, stream, max_blocks);
}

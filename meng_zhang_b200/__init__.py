"""meng_zhang_b200 -- B200-native ANNP / ANNA-ADP neural-network-potential force evaluation.

The product is the CUDA library `lib/libannp_b200.so` (sources in `csrc/`, C ABI in `include/annp_b200.h`) and the
LAMMPS-facing C++ pair styles in `lammps/`.  Python here is host glue only:
  capi           ctypes prototypes of the C ABI
  pair           PairANNPGPU: mirror of the reference `annp/gpu` interface (settings/coeff/init_style/compute), Fe and Ni copies
  pair_anna      PairANNAADPGPU: the same for `anna_adp/gpu`
  lattice        synthetic configurations, ghost shells and host neighbour lists
  structures     screw-dislocation and symmetric-tilt-boundary generators (the reference's two programs)
  md             device-resident MD driver: domain decomposition over NCCL, NVE / Nose-Hoover NVT / NPT, cg minimiser
  lammps_compat  LAMMPS behaviours a replay of the reference's runs needs (velocity generator, shrink-wrapped box)
  deck           runs the reference's LAMMPS input decks verbatim on `md`
"""
__all__ = ["capi", "pair", "pair_anna", "lattice", "structures", "md", "lammps_compat", "deck"]

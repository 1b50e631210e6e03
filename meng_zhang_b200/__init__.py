"""meng_zhang_b200 -- B200-native ANNP neural-network-potential force evaluation.

The product is the CUDA library `lib/libannp_b200.so` (sources in `csrc/`, C ABI in
`include/annp_b200.h`).  Python here is host glue only:
  capi      ctypes prototypes of the C ABI
  pair      PairANNPGPU: mirror of the reference pair-style interface (settings/coeff/init_style/compute)
  lattice   synthetic configurations, ghost shells and host neighbour lists
  md        device-resident MD driver (NVE, domain decomposition over NCCL) used by bench.py
"""
__all__ = ["capi", "pair", "lattice"]

/* -*- c++ -*- ----------------------------------------------------------
   pair_style annp/gpu served by libannp_b200.so (NVIDIA B200, sm_100a).

   Drop-in for the reference's src/pair_annp_gpu.{h,cpp} (annp-gpu-lammps/fe_v2): copy this
   header and pair_annp_b200.cpp next to the reference's own pair_annp.{h,cpp} in the LAMMPS src
   tree INSTEAD of pair_annp_gpu.{h,cpp}, drop include/annp_b200.h beside them and link
   libannp_b200.so.  Nothing of LAMMPS' GPU package (lib/gpu, Geryon, fix gpu, `package gpu`) is
   needed any more.  Input decks are unchanged:

       pair_style  annp/gpu
       pair_coeff  * * fe_annp_potential_2.ann Fe

   Environment (the deck keeps the reference's zero-argument pair_style line):
       ANNP_B200_NEIGH=device    build the neighbour list on the GPU (the reference's `package gpu N neigh yes`)
       ANNP_B200_SCATTER=gather  ordered FP64 gather instead of the fixed-point force accumulation
------------------------------------------------------------------------- */

#ifdef PAIR_CLASS
// clang-format off
PairStyle(annp/gpu, PairANNPB200);
// clang-format on
#else

#ifndef LMP_PAIR_ANNP_B200_H
#define LMP_PAIR_ANNP_B200_H

#include "pair_annp.h"      // the reference's CPU style: file parsing, coeff(), init_one()
#include "annp_b200_host.h"

struct annp_b200_handle_s;

namespace LAMMPS_NS {

class PairANNPB200 : public PairANNP {
 public:
  PairANNPB200(class LAMMPS *);
  ~PairANNPB200() override;
  void compute(int, int) override;
  void init_style() override;
  double memory_usage() override;

 protected:
  annp_b200_handle_s *handle;
  ANNP_B200_NS::HostBuffers hb;   // page-locked views of atom->x / atom->f and the style's own staging arrays
  int device_neigh;               // ANNP_B200_NEIGH=device: neighbour list built on the GPU (`package gpu ... neigh yes`)
};

}    // namespace LAMMPS_NS

#endif
#endif

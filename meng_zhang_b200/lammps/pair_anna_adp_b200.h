/* -*- c++ -*- ----------------------------------------------------------
   pair_style anna_adp/gpu served by libannp_b200.so (NVIDIA B200, sm_100a).

   Drop-in for the reference's src/pair_anna_adp_gpu.{h,cpp} (anna-gpu-lammps/bcc_fe): copy this header and
   pair_anna_adp_b200.cpp next to the reference's own pair_anna_adp.{h,cpp} INSTEAD of pair_anna_adp_gpu.{h,cpp},
   drop include/annp_b200.h beside them and link libannp_b200.so.  Input decks keep

       pair_style  anna_adp/gpu
       pair_coeff  * * fe_adp_potential_2310.anna Fe

   The numbers follow the reference CPU style (pair_anna_adp.cpp:73-286): everything is centred on the local atom, so
   the 12 forward communications per step of the reference GPU style (pair_anna_adp_gpu.cpp:135-153) disappear and
   only the forces on ghost atoms have to go home.  Both settings of `newton` work: with `newton on` LAMMPS' own reverse
   communication carries them; with `newton off` - what the reference's GPU decks say (bcc_fe/README.md:39-40) - the style
   reverse-communicates its ghost forces itself (comm->reverse_comm(this), three doubles per ghost).
------------------------------------------------------------------------- */

#ifdef PAIR_CLASS
// clang-format off
PairStyle(anna_adp/gpu, PairANNAADPB200);
// clang-format on
#else

#ifndef LMP_PAIR_ANNA_ADP_B200_H
#define LMP_PAIR_ANNA_ADP_B200_H

#include "pair_anna_adp.h"      // the reference's CPU style: file parsing, coeff(), init_one()
#include "annp_b200_host.h"

struct annp_b200_handle_s;

namespace LAMMPS_NS {

class PairANNAADPB200 : public PairANNA_ADP {
 public:
  PairANNAADPB200(class LAMMPS *);
  ~PairANNAADPB200() override;
  void compute(int, int) override;
  void init_style() override;
  double memory_usage() override;
  int pack_reverse_comm(int, int, double *) override;
  void unpack_reverse_comm(int, int *, double *) override;

 protected:
  annp_b200_handle_s *handle;
  ANNP_B200_NS::HostBuffers hb;   // page-locked views of atom->x / atom->f and the style's own staging arrays
  double *fstage;                 // forces of the current compute() while the style's own reverse_comm runs (newton off)
  int device_neigh;               // ANNP_B200_NEIGH=device: neighbour list built on the GPU
};

}    // namespace LAMMPS_NS

#endif
#endif

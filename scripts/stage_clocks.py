"""Per-stage time split of the Fe force kernel: run the bench-size force evaluation with the
-DANNP_STAGE_CLOCKS build of the library (clock64 stamps between the stages, summed over warps) and print each stage's
share of the warp-cycles.

    make -C meng_zhang_b200/csrc OUT=../lib_clk EXTRA=-DANNP_STAGE_CLOCKS
    ANNP_B200_LIB=$PWD/meng_zhang_b200/lib_clk/libannp_b200.so python scripts/stage_clocks.py --cells 64
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from meng_zhang_b200 import capi, lattice as L, potentials  # noqa: E402
from meng_zhang_b200.md import DomainMD  # noqa: E402
from meng_zhang_b200.pair import PairANNPGPU  # noqa: E402
from meng_zhang_b200.pair_anna import PairANNAADPGPU  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cells", type=int, default=64)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--style", default="fe", choices=["fe", "anna"])
a = ap.parse_args()
if a.style == "anna":
    pot, cls = potentials.write_anna_fe_potential("/tmp/annp_b200_clk.anna"), PairANNAADPGPU
else:
    pot, cls = potentials.write_fe_potential("/tmp/annp_b200_clk.ann"), PairANNPGPU
pair = cls(ntypes=1)
pair.settings([])
pair.coeff(["*", "*", pot, "Fe"])
pair.init_style()
x, box = L.bcc(a.cells, a.cells, a.cells)
md = DomainMD(pair, L.perturb(x, 0.05, 1), box)
md.set_velocities(300.0, 1)
md.reneighbor()
md.compute(eflag=True)
torch.cuda.synchronize()
pair.stats()                                   # clears the accumulated stage clocks
capi.lib().annp_b200_set_timing(pair.handle, 1)
for _ in range(a.steps):
    md.compute()
st = pair.stats()
cyc = np.array(st.stage_cycles[:])
names = ["1a filter (list row walk, cutoff test, compaction)", "1b per-neighbour geometry, fc, radial sums", "2 forward angular sums",
         "3 reduction, basis conversion, MLP forward + backward", "4 backward angular moments", "5 force assembly + scatter", "scheduler (atom counter)", "-"]
tot = cyc.sum()
out = {"lib": capi.LIB_PATH, "style": a.style, "atoms": md.nlocal, "steps": a.steps, "kernel_ms": st.force_kernel_ms_total / max(st.force_kernel_samples, 1),
       "neighbors_in_cutoff": st.avg_neigh_cut,
       "stage_share": {n: float(c / tot) for n, c in zip(names, cyc) if n != "-"} if tot > 0 else None,
       "warp_cycles_per_atom": float(tot / a.steps / md.nlocal) if tot > 0 else None,
       "note": "clock64 stamps add ~14 instructions per atom; shares are of summed per-warp cycles (4 warps per scheduler interleave, so a "
               "stage's share is its share of the warp's residency, not of the pipe)"}
print(json.dumps(out))

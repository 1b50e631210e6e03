"""Small end-to-end case touching every kernel of libannp_b200 (for `compute-sanitizer --tool memcheck`, development aid):
Fe / Ni (three kernels) / ANNA-ADP force evaluations with per-atom energy and virial, device neighbour build, halo,
NVE and Nose-Hoover NPT steps."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import util  # noqa: E402
from meng_zhang_b200 import capi, lattice as L  # noqa: E402
from meng_zhang_b200.md import DomainMD  # noqa: E402
from meng_zhang_b200.pair import PairANNPGPU  # noqa: E402
from meng_zhang_b200.pair_anna import PairANNAADPGPU  # noqa: E402


def pair_of(cls, pot, elem, **kw):
    p = cls(ntypes=1, **kw)
    p.settings([])
    p.coeff(["*", "*", pot, elem])
    p.init_style()
    return p


fe = util.write_fe_potential("/tmp/san_fe.ann")
ni = util.write_ni_potential("/tmp/san_ni.ann")
an = util.write_anna_fe_potential("/tmp/san_an.anna")
for name, prefix, mk in (("bcc4_perturbed", "annp_fe", lambda: pair_of(PairANNPGPU, fe, "Fe")),
                         ("cluster_ragged", "annp_fe", lambda: pair_of(PairANNPGPU, fe, "Fe")),
                         ("fcc3_perturbed", "annp_ni", lambda: pair_of(PairANNPGPU, ni, "Ni")),
                         ("fcc3_perturbed", "annp_ni", lambda: pair_of(PairANNPGPU, ni, "Ni", variant=capi.VARIANT_NI | capi.VARIANT_FLAG_NOPAIR)),
                         ("cluster_ragged", "annp_ni", lambda: pair_of(PairANNPGPU, ni, "Ni", variant=capi.VARIANT_NI | capi.VARIANT_FLAG_GENERIC)),
                         ("bcc4_perturbed", "anna_adp", lambda: pair_of(PairANNAADPGPU, an, "Fe")),
                         ("cluster_ragged", "anna_adp", lambda: pair_of(PairANNAADPGPU, an, "Fe"))):
    cfg, elems, ref = util.load_case(name, prefix)
    pair = mk()
    f = pair.compute(3, 1 + 4, cfg, ago=0)
    print(prefix, name, "max|dF|", float(np.abs(f - ref["f"]).max()))
    pair.clear()

pair = pair_of(PairANNPGPU, fe, "Fe")
x, box = L.bcc(4, 4, 4)
md = DomainMD(pair, L.perturb(x, 0.05, 1), box)
md.set_velocities(300.0, 1)
md.reneighbor()
md.compute(eflag=True)
for _ in range(3):
    md.step(eflag=True)
md.fix_nh(300.0, 300.0, 0.1, p_flag=(0, 1, 0), p_start=(0.0,) * 3, p_stop=(0.0,) * 3, p_damp=(1.0,) * 3)
for _ in range(3):
    md.step_nh(eflag=True)
md.sync_box_from_nh()
md.reneighbor()
md.step_nh()
print("md ok", md.nh_state().t_current)
pair.clear()

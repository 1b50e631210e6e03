"""Quick on-GPU sanity run of the CUDA path against the golden fixtures (development aid)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from meng_zhang_b200.pair import PairANNPGPU  # noqa: E402
import util  # noqa: E402

pot = util.write_fe_potential("/tmp/fe.ann")
for name in util.FE_CASES:
    cfg, elems, ref = util.load_case(name)
    pair = PairANNPGPU(ntypes=len(elems))
    pair.settings([])
    pair.coeff(["*", "*", pot] + elems)
    pair.init_style()
    t = time.time()
    f = pair.compute(3, 1 + 4, cfg, ago=0)
    dt = time.time() - t
    st = pair.stats()
    print(f"{name}: dE {pair.eng_vdwl - ref['eng_vdwl']:.3e} (E {ref['eng_vdwl']:.6f})  "
          f"eatom {np.abs(pair.eatom - ref['eatom']).max():.3e}  f {np.abs(f - ref['f']).max():.3e} (fmax {np.abs(ref['f']).max():.3f})  "
          f"vir {np.abs(pair.virial - ref['virial_pair']).max():.3e}  vatom {np.abs(pair.vatom - ref['vatom']).max():.3e}  "
          f"maxN {st.max_neigh_cut} avgN {st.avg_neigh_cut:.2f} t {dt * 1e3:.1f} ms")
    f2 = pair.compute(3, 1 + 4, cfg, ago=1)
    print("   repeat bitwise identical:", np.array_equal(f, f2))
    pair.clear()

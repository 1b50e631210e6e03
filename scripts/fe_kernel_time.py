"""Fe force kernel: parity against the oracle on a small cell + CUDA-event kernel time at a chosen size (development
aid for A/B runs of kernel builds: ANNP_B200_LIB=<other .so> python scripts/fe_kernel_time.py --cells 40)."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import util  # noqa: E402
from meng_zhang_b200 import capi, lattice as L  # noqa: E402
from meng_zhang_b200.pair import PairANNPGPU, read_potential  # noqa: E402
from test_gpu_parity import build_large_config  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cells", type=int, default=40)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--scatter", default="fixed", choices=["fixed", "gather"])
a = ap.parse_args()
pot = util.write_fe_potential("/tmp/annp_b200_fkt_fe.ann")
pair = PairANNPGPU(ntypes=1)
pair.settings([])
pair.coeff(["*", "*", pot, "Fe"])
pair.init_style()
pair.set_scatter(capi.SCATTER_FIXED if a.scatter == "fixed" else capi.SCATTER_GATHER)
out = {"scatter": a.scatter, "lib": os.path.basename(capi.LIB_PATH)}
from oracle import restatement  # noqa: E402
x, box = L.bcc(4, 4, 4)
cfg = L.build_config(L.perturb(x, 0.08, 99), box, 6.5)
f = pair.compute(3, 1, cfg, ago=0)
ref = restatement.compute(read_potential(pot, ["Fe"]), cfg, nthreads=4)
out["max_dF"] = float(np.abs(f - ref["f"]).max())
out["max_dEi"] = float(np.abs(pair.eatom - ref["eatom"]).max())
x, box = L.bcc(a.cells, a.cells, a.cells)
cfg = build_large_config(L.perturb(x, 0.05, 1), box, (True, True, True), cutoff=6.5)
pair.compute(1, 0, cfg, ago=0)
for _ in range(2):
    pair.compute(1, 0, cfg, ago=1)
lib = capi.lib()
lib.annp_b200_set_timing(pair.handle, 1)
import time  # noqa: E402
t0 = time.perf_counter()
for _ in range(a.steps):
    pair.compute(1, 0, cfg, ago=1)
out["host_call_ms"] = (time.perf_counter() - t0) / a.steps * 1e3
st = pair.stats()
ms = st.force_kernel_ms_total / max(st.force_kernel_samples, 1)
out.update({"atoms": cfg.nlocal, "neighbors_in_cutoff": st.avg_neigh_cut, "force_kernel_ms": ms,
            "kernel_atom_steps_per_s": cfg.nlocal / (ms * 1e-3), "E_per_atom": pair.eng_vdwl / cfg.nlocal})
print(json.dumps(out), flush=True)

"""Timing of the Ni (BASELINE config 2) and ANNA-ADP force paths on one GPU (development aid, not bench.py).

    python scripts/bench_variants.py [--ni-cells 20] [--anna-cells 40] [--steps 10]

Prints one JSON line per variant: host-call atom-steps/s (annp_b200_compute, pinned-less numpy buffers), the force
kernel's CUDA-event time and the time of one Pair::compute of the LAMMPS-facing C++ class (page-locked in place).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import util  # noqa: E402
from meng_zhang_b200 import capi, lattice as L  # noqa: E402
from meng_zhang_b200.pair import PairANNPGPU  # noqa: E402
from meng_zhang_b200.pair_anna import PairANNAADPGPU  # noqa: E402
from test_gpu_parity import build_large_config  # noqa: E402


def plugin_call_ms(binary, cfg, pot, elem, steps):
    """Pair::compute of the LAMMPS-facing C++ class (oracle/_ref/plugin_*: shim driver hosting PairANNPB200 / PairANNAADPB200,
    x / f malloc'd and page-locked in place) timed inside the driver; None when the binary is absent."""
    from oracle import run_ref
    if not run_ref.available(binary):
        return None
    out = run_ref.run_reference(binary, cfg, pot, [elem], eflag=1, vflag=0, ncalls=steps + 3)
    return float(out["per_call_seconds"][3:].mean() * 1e3)


def run(kind, pair, cfg, steps, plugin=None):
    lib = capi.lib()
    t = time.perf_counter()
    pair.compute(1, 0, cfg, ago=0)
    t_first = time.perf_counter() - t
    for _ in range(3):
        pair.compute(1, 0, cfg, ago=1)
    lib.annp_b200_set_timing(pair.handle, 1)
    t = time.perf_counter()
    for _ in range(steps):
        pair.compute(1, 0, cfg, ago=1)
    dt = (time.perf_counter() - t) / steps
    st = pair.stats()
    lib.annp_b200_set_timing(pair.handle, 0)
    print(json.dumps({"variant": kind, "atoms": cfg.nlocal, "nall": cfg.nall, "list_neighbors": float(cfg.numneigh.mean()),
                      "neighbors_in_cutoff": st.avg_neigh_cut, "max_in_cutoff": st.max_neigh_cut,
                      "host_call_ms": dt * 1e3, "host_call_atom_steps_per_s": cfg.nlocal / dt,
                      "force_kernel_ms": st.force_kernel_ms_total / max(st.force_kernel_samples, 1),
                      "kernel_atom_steps_per_s": cfg.nlocal / (st.force_kernel_ms_total / max(st.force_kernel_samples, 1) * 1e-3),
                      "first_call_with_list_upload_ms": t_first * 1e3, "E_per_atom": pair.eng_vdwl / cfg.nlocal,
                      "plugin_call_ms": plugin_call_ms(*plugin, steps) if plugin else None}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ni-cells", type=int, default=20)
    ap.add_argument("--anna-cells", type=int, default=40)
    ap.add_argument("--steps", type=int, default=10)
    a = ap.parse_args()
    if a.ni_cells > 0:
        pot = util.write_ni_potential("/tmp/annp_b200_bv_ni.ann")
        x, box = L.fcc(a.ni_cells, a.ni_cells, a.ni_cells)
        cfg = build_large_config(L.perturb(x, 0.05, 1), box, (True, True, True), cutoff=6.5)
        pair = PairANNPGPU(ntypes=1)
        pair.settings([])
        pair.coeff(["*", "*", pot, "Ni"])
        pair.init_style()
        run("ni", pair, cfg, a.steps, plugin=("plugin_annp_ni_b200", cfg, pot, "Ni"))
        pair.clear()
    if a.anna_cells > 0:
        pot = util.write_anna_fe_potential("/tmp/annp_b200_bv_anna.anna")
        x, box = L.bcc(a.anna_cells, a.anna_cells, a.anna_cells)
        cfg = build_large_config(L.perturb(x, 0.05, 1), box, (True, True, True), cutoff=5.055)
        pair = PairANNAADPGPU(ntypes=1)
        pair.settings([])
        pair.coeff(["*", "*", pot, "Fe"])
        pair.init_style()
        run("anna_adp", pair, cfg, a.steps, plugin=("plugin_anna_adp_b200", cfg, pot, "Fe"))
        pair.clear()


if __name__ == "__main__":
    main()

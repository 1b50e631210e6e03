"""The other BASELINE configurations, end to end on the device (bench.py measures the headline one, configs[4]):

    config 2   fcc Ni ANNP (Ni copy of the style), 32 000 atoms, NVT 300 K
    config 3   bcc Fe ANNP, 40x40x80 cells = 256 000 atoms, NPT 300 K (the deck's `y 0 0 1` coupling)
    config 4   bcc Fe screw dislocation (structures.screw_dislocation((22, 38, 50)) ~ 5.1e5 atoms), free x / y surfaces,
               periodic z, rim atoms (type 2) held fixed, NVE, decomposed over the ranks it is launched on
    config 5b  bcc Fe bicrystal with two symmetric tilt grain boundaries (structures.stgb, the reference's stgb.cpp geometry),
               519 480 atoms = one weak-scaling cell, two LAMMPS atom types both mapped to Fe, periodic, NVE  (--only 5b)
    anna       bcc Fe ANNA-ADP, 40^3 cells = 128 000 atoms, NVT 300 K

    python scripts/bench_configs.py [--steps 100] [--only 2,3,4,anna]      (ANNP_BENCH_GRAPH=1: replay the step as a CUDA graph)
    python -m torch.distributed.run --nproc-per-node 8 ... scripts/bench_configs.py --only 4

One JSON line per configuration on rank 0: atom-steps/s (CUDA events, max over ranks), ms per step, temperature /
pressure at the end as a sanity value, and the same `roofline` / `clocks` objects as bench.py for the configuration's force
kernel (rank 0's launch): algorithmic flops of SURVEY 8d (Fe 278 T + 168 N + 1560; Ni 350 T' + 100 N + 4(27*24 + 24^2 + 24)
with T' the contributing triplets; ANNA 86 T + 300 N + 432) over the CUDA-event kernel time against the measured FP64 peak,
plus the algorithmic HBM bytes (list 4 B per entry, position 32 B, force 24 B, energy 8 B per atom) against the measured
copy bandwidth of MEASURED_PEAKS.json.  Development aid + evidence for profiles/, not the driver's bench.
"""
import argparse
import json
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (ClockSampler)
from meng_zhang_b200 import capi, lattice as L, potentials as util, structures as S  # noqa: E402
from meng_zhang_b200.md import DomainMD, decompose, rank_coords  # noqa: E402
from meng_zhang_b200.pair import PairANNPGPU  # noqa: E402
from meng_zhang_b200.pair_anna import PairANNAADPGPU  # noqa: E402


def make_pair(kind, dev):
    tmp = tempfile.gettempdir()
    if kind == "ni":
        pot, elem, cls = util.write_ni_potential(os.path.join(tmp, f"bc_ni_{dev}.ann")), "Ni", PairANNPGPU
    elif kind == "anna":
        pot, elem, cls = util.write_anna_fe_potential(os.path.join(tmp, f"bc_anna_{dev}.anna")), "Fe", PairANNAADPGPU
    else:
        pot, elem, cls = util.write_fe_potential(os.path.join(tmp, f"bc_fe_{dev}.ann")), "Fe", PairANNPGPU
    pair = cls(ntypes=2 if kind == "fe2" else 1, device=dev)
    pair.settings([])
    pair.coeff(["*", "*", pot] + [elem] * pair.ntypes)
    pair.init_style()
    return pair


def run_case(name, kind, x_all, box, mass, periodic, ensemble, steps, rank, world, local, dev, types=None, frozen=None):
    grid = decompose(world)
    coords = rank_coords(rank, grid)
    lo = np.array([box[d] * coords[d] / grid[d] for d in range(3)])
    hi = np.array([box[d] * (coords[d] + 1) / grid[d] for d in range(3)])
    x_all = L.wrap(x_all, box, periodic)
    inside = np.ones(len(x_all), dtype=bool)
    for d in range(3):      # free surfaces: the outermost bricks own whatever lies beyond the nominal box face
        lo_d = -np.inf if (coords[d] == 0 and not periodic[d]) else lo[d]
        hi_d = np.inf if (coords[d] == grid[d] - 1 and not periodic[d]) else hi[d]
        inside &= (x_all[:, d] >= lo_d) & (x_all[:, d] < hi_d)
    pair = make_pair(kind, local)
    md = DomainMD(pair, x_all[inside], box, grid=grid, rank=rank, device=dev, mass=mass, dt=0.001, periodic=periodic,
                  type_local=None if types is None else types[inside], frozen_local=None if frozen is None else frozen[inside])
    md.set_velocities(300.0, 4928459)
    md.reneighbor()
    if ensemble == "nve":
        md.compute(eflag=True)
        stepper = md.step
    else:
        p_flag = (0, 1, 0) if ensemble == "npt" else (0, 0, 0)
        md.fix_nh(300.0, 300.0, 0.1, p_flag=p_flag, p_start=(0.0,) * 3, p_stop=(0.0,) * 3, p_damp=(1.0,) * 3)
        stepper = md.step_nh
    for _ in range(5):
        stepper()
    graph = bool(os.environ.get("ANNP_BENCH_GRAPH"))
    if graph:       # the step is ~10 short kernels (+ the grouped NCCL exchange): replay it as one CUDA graph per rank
        md.capture_step(nh=ensemble != "nve")
        stepper = lambda: md.replay(1)
        for _ in range(5):
            stepper()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    Lb = capi.lib()
    pair.stats()
    if not graph:
        Lb.annp_b200_set_timing(pair.handle, 1)      # CUDA-event ring around the force kernel (eager launches only)
    sampler = bench.ClockSampler(local)
    if rank == 0:
        sampler.start()
        sampler.mark_begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        stepper()
    e1.record()
    torch.cuda.synchronize(dev)
    if rank == 0:
        sampler.mark_end()
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t) / steps
    st = pair.stats()
    Lb.annp_b200_set_timing(pair.handle, 0)
    extra = {}
    if rank == 0:
        kern_ms = st.force_kernel_ms_total / max(st.force_kernel_samples, 1)
        n_loc, nbar = md.nlocal, st.avg_neigh_cut
        if kind == "ni":
            flop = 350.0 * st.sum_triplets + 100.0 * nbar * n_loc + 4.0 * (27 * 24 + 24 * 24 + 24) * n_loc
            formula = "350 T' + 100 N + 4 (27*24 + 24^2 + 24), T' = contributing triplets counted by the kernel"
        elif kind == "anna":
            flop = 86.0 * st.sum_triplets + 300.0 * nbar * n_loc + 432.0 * n_loc
            formula = "86 T + 300 N + 432, T = N(N-1)/2"
        else:
            flop = 278.0 * st.sum_triplets + 168.0 * nbar * n_loc + 1560.0 * n_loc
            formula = "278 T + 168 N + 1560, T = N(N-1)/2"
        list_entries = Lb.annp_b200_debug_neighbors(pair.handle, None, None)
        hbm_bytes = 4.0 * list_entries + 32.0 * (md.nlocal + md.nghost) + (24.0 + 8.0) * n_loc
        peak_tf = Lb.annp_b200_fp64_peak_tflops(pair.handle, 3)
        try:
            hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
            hbm_src = "MEASURED_PEAKS.json hbm_gbs"
        except (OSError, KeyError, ValueError):
            hbm_peak, hbm_src = 6468.6, "fallback of B200_PROFILING.md (MEASURED_PEAKS.json absent)"
        if kern_ms > 0:
            tf = flop / (kern_ms * 1e-3) / 1e12
            gbs = hbm_bytes / (kern_ms * 1e-3) / 1e9
            extra["roofline"] = {"bound": "fp64", "kernel": {"ni": "annp_bp_pair_kernel<3,4>", "anna": "annp_force_kernel<9,19,1>"}.get(kind, "annp_force_kernel<9,19,0>"),
                                 "achieved": tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": tf / peak_tf if peak_tf > 0 else None,
                                 "traffic": None, "kernel_ms": kern_ms, "kernel_share_of_step": kern_ms / ms,
                                 "flop_per_atom_step": flop / n_loc, "flop_formula": formula,
                                 "peak_source": "annp_b200_fp64_peak_tflops (DFMA loop on this GPU)",
                                 "hbm": {"achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                                         "algorithmic_bytes_per_atom_step": hbm_bytes / n_loc, "peak_source": hbm_src}}
        extra["clocks"] = clocks
    if ensemble != "nve":
        s = md.nh_state()
        extra.update({"T": s.t_current, "p_bar": list(s.p_current[:]), "box": [s.boxhi[d] - s.boxlo[d] for d in range(3)]})
    nat = torch.tensor([float(md.nlocal)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(nat)
    if rank == 0:
        print(json.dumps({"config": name, "atoms": int(nat), "n_gpus": world, "grid": "x".join(map(str, grid)), "ensemble": ensemble, "cuda_graph": graph,
                          "steps": steps, "ms_per_step": ms, "atom_steps_per_s": float(nat) / (ms * 1e-3),
                          "ns_per_day": 86400.0 / (ms * 1e-3) * 1e-6, "neighbors_in_cutoff": st.avg_neigh_cut,
                          "list_neighbors_max": st.max_neigh_list, **extra}), flush=True)
    md.close()
    pair.clear()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--only", default="2,3,4,anna")
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    only = a.only.split(",")
    pbc = (True, True, True)
    if "2" in only:
        x, box = L.fcc(20, 20, 20)
        run_case("2: fcc Ni ANNP 32 000 atoms NVT 300 K", "ni", L.perturb(x, 0.02, 1), box, 58.6934, pbc, "nvt", a.steps, rank, world, local, dev)
    if "3" in only:
        x, box = L.bcc(40, 40, 80)
        run_case("3: bcc Fe ANNP 256 000 atoms NPT 300 K (y 0 0 1)", "fe", L.perturb(x, 0.02, 2), box, 55.845, pbc, "npt", a.steps, rank, world, local, dev)
    if "4" in only:
        xs, box, types, core = S.screw_dislocation((22, 38, 50))
        run_case("4: bcc Fe screw dislocation, free x/y, periodic z, fixed rim, NVE", "fe", xs, box, 55.845, (False, False, True), "nve",
                 a.steps, rank, world, local, dev, frozen=(types == 2))
    if "5b" in only:      # BASELINE configs[4] variant: the weak-scaling cell filled with the reference's STGB bicrystal
        u = S.stgb_unit_lengths()
        xs, box, types = S.stgb(length_box=(13 * u[0], 37 * u[1], 45 * u[2]))     # 26 x 37 x 45 units = 181.8 x 183.0 x 181.7 A
        run_case(f"5b: bcc Fe symmetric tilt grain boundaries (stgb.cpp geometry), {len(xs)} atoms, two atom types -> Fe, PBC, NVE", "fe2",
                 xs, box, 55.845, pbc, "nve", a.steps, rank, world, local, dev, types=types)
    if "anna" in only:
        x, box = L.bcc(40, 40, 40)
        run_case("anna: bcc Fe ANNA-ADP 128 000 atoms NVT 300 K", "anna", L.perturb(x, 0.02, 3), box, 55.845, pbc, "nvt", a.steps, rank, world, local, dev)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

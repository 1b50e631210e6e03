#!/usr/bin/env python
"""bench.py -- atom-steps/s of the ANNP force path on bcc Fe (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--cells C] [--impl ours|reference]

Workload (config.workload): BASELINE.json configs[4], the weak-scaling cell - 64x64x64 bcc-Fe cells
= 524 288 atoms PER GPU (a=2.8553 A, thermally perturbed lattice, 300 K velocities, PBC, 2 A skin),
potential fe_annp_potential_2 (28 SF, 28-10-10-1).  One step = one velocity-Verlet MD step: forward
halo, force evaluation (pack, fused descriptor/MLP/force kernel with fixed-point force scatter, conversion),
reverse halo, integration.  N ranks = LAMMPS-style brick decomposition 1x1x1 / 2x1x1 / 2x2x1 / 2x2x2, one rank per
GPU, halo exchange on device buffers over NCCL.

value   device-resident MD (positions never leave HBM), CUDA-event timed, max over ranks
e2e     the same force evaluation through the reference-facing host call annp_b200_compute (host x in
        pinned memory -> device, forces/energy back to the host every step), i.e. what
        PairANNPGPU::compute costs inside a host-driven LAMMPS
--impl reference : the reference's own CPU pair style (oracle/_ref/ref_annp_fe, unmodified source) on the
        host cores, one process per core on spatial chunks of a bounded sample of the same lattice.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

A_FE = 2.8553
RC = 6.5
SKIN = 2.0
FLOP_TRIPLET, FLOP_PAIR, FLOP_MLP = 278.0, 168.0, 1560.0     # F_alg = 278 T + 168 N + 1560
NCU_TRAFFIC_BYTES = 5.563e8                                  # profiles/r1c_force_kernel.md: 521.7 MB read + 34.7 MB written per launch
# FP64 flops the kernel EXECUTES per atom-step at this workload, from the same capture: (2 x 7.324e9 DFMA + 0.934e9 DMUL +
# 0.743e9 DADD warp instructions) x 30.87 active threads / 524 288 atoms.  The algorithmic count prices the
# straightforward recompute formulation at 278 flops per triplet; the kernel needs ~150, so `frac` (algorithmic, the
# contract's definition) can exceed 1 while the pipe itself is `executed.frac` busy with useful flops.
NCU_EXECUTED_FLOP_PER_ATOM_STEP = 9.612e5
PUBLISHED_ATOM_STEPS_PER_S = 152880 * 1000 / 1789.44         # BASELINE.md section 1 (the reference's own 2-GPU log)


def lattice_block(cells, origin_cells, amp, seed):
    """bcc block of `cells`^3 unit cells whose first cell sits at origin_cells (in cells)."""
    from meng_zhang_b200 import lattice as L
    x, _ = L.bcc(cells, cells, cells, A_FE)
    x += np.asarray(origin_cells, dtype=np.float64) * A_FE
    rng = np.random.default_rng(seed)
    x += rng.uniform(-amp, amp, size=x.shape)
    return x


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe).

    The sampler runs from before the warm-up; mark() brackets the timed region and only samples whose
    arrival time falls inside it are reported."""

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index
        self.t0 = self.t1 = None

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,timestamp"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        # nvidia-smi block-buffers its output when it writes to a pipe, so lines ARRIVE in bursts: each sample is placed
        # by nvidia-smi's own timestamp column (same host clock as time.time()); arrival time is only the fallback
        import datetime
        for line in self.proc.stdout:
            cols = [c.strip() for c in line.split(",")]
            t = time.time()
            try:
                t = datetime.datetime.strptime(cols[7], "%Y/%m/%d %H:%M:%S.%f").timestamp()
            except (IndexError, ValueError):
                pass
            self.rows.append((t, cols))

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], None, [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, r in self.rows:
            if len(r) < 7 or self.t0 is None or not (self.t0 <= t <= (self.t1 or t) + 0.06):
                continue
            try:
                sm.append(float(r[0]))
                smax = float(r[1])
                power.append(float(r[2]))
            except ValueError:
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "power_w_max": max(power) if power else None, "samples": len(sm)}


def run_ours(args):
    import ctypes as C
    import torch
    import torch.distributed as dist
    import util
    from meng_zhang_b200 import capi
    from meng_zhang_b200.md import DomainMD, decompose
    from meng_zhang_b200.pair import PairANNPGPU

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    grid = decompose(world)
    cells = args.cells
    from meng_zhang_b200.md import rank_coords
    coords = rank_coords(rank, grid)
    box = np.array([grid[d] * cells * A_FE for d in range(3)])
    x_local = lattice_block(cells, [coords[d] * cells for d in range(3)], 0.05, 1000 + rank)
    nlocal = len(x_local)

    pot_file = util.write_fe_potential(os.path.join(tempfile.gettempdir(), f"annp_b200_bench_fe_{rank}.ann"))
    pair = PairANNPGPU(ntypes=1, device=local_rank, skin=SKIN)
    pair.settings([])
    pair.coeff(["*", "*", pot_file, "Fe"])
    pair.init_style()
    L = capi.lib()

    md = DomainMD(pair, x_local, box, grid=grid, rank=rank, device=dev, skin=SKIN, mass=55.845, dt=0.001)
    md.set_velocities(300.0, 4928459)
    md.reneighbor()
    md.compute(eflag=True)
    torch.cuda.synchronize(dev)
    st0 = pair.stats()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---------------- device-resident MD: `value`
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        md.step()
    barrier()
    L.annp_b200_set_timing(pair.handle, 1)
    launches0 = pair.stats().kernel_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.mark_begin()
    e0.record()
    for _ in range(args.steps):
        md.step()
    e1.record()
    barrier()
    sampler.mark_end()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    st = pair.stats()
    L.annp_b200_set_timing(pair.handle, 0)
    launches = st.kernel_launches - launches0
    tmax = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_total = float(tmax)
    natoms_total = nlocal * world
    value = natoms_total * args.steps / (ms_total * 1e-3)

    # force-kernel roofline (this rank): algorithmic flops from the actual neighbour histogram
    kern_ms = st.force_kernel_ms_total / max(st.force_kernel_samples, 1)
    flop_per_launch = FLOP_TRIPLET * st.sum_triplets + FLOP_PAIR * st.avg_neigh_cut * nlocal + FLOP_MLP * nlocal
    achieved_tf = flop_per_launch / (kern_ms * 1e-3) / 1e12 if kern_ms > 0 else 0.0
    peak_tf = L.annp_b200_fp64_peak_tflops(pair.handle, 5)

    # ---------------- e2e through the host-buffer C ABI (PairANNPGPU::compute path)
    nall = md.nlocal + md.nghost
    hx = torch.empty((nall, 3), dtype=torch.float64).pin_memory()
    hx.copy_(md.x)
    htype = torch.empty(nall, dtype=torch.int32).pin_memory()
    htype.copy_(md.type)
    hf = torch.empty((nall, 3), dtype=torch.float64).pin_memory()
    eng = C.c_double(0.0)
    dp = lambda t: C.cast(C.c_void_p(t.data_ptr()), capi.c_double_p)
    ip = lambda t: C.cast(C.c_void_p(t.data_ptr()), capi.c_int_p)

    def host_step():
        rc = L.annp_b200_compute(pair.handle, md.nlocal, md.nghost, dp(hx), ip(htype), 1, 0, dp(hf), C.byref(eng), None, None, None)
        if rc != 0:
            raise RuntimeError(L.annp_b200_last_error(pair.handle).decode())

    for _ in range(min(args.warmup, 3)):
        host_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        host_step()
    torch.cuda.synchronize(dev)
    t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = natoms_total * args.steps / float(t_e2e)
    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_base = cpu_baseline_reference(sample_cells=10, repeats=3)
    published = None
    if rank == 0 and world == 1 and not args.no_published_deck:
        published = run_published_deck(pot_file, dev)

    if rank == 0:
        out = {
            "metric": "atom-steps/sec (bcc Fe ANNP)", "value": value, "unit": "atom-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak",
            # the only speed the reference publishes for this metric, 8.55e4 atom-steps/s (whole job:
            # its 152 880-atom deck on the authors' 2 GPUs).  Same metric, different box size and hardware; the
            # like-for-like rerun of that deck is the `published_deck` object below.
            "vs_baseline": value / PUBLISHED_ATOM_STEPS_PER_S,
            "vs_baseline_note": "value / 8.55e4 atom-steps/s (annp/gpu fe_v2, 152 880 atoms, 2 GPUs, zip:log_relaxing_new.lammps:1168); see published_deck for the same deck",
            "dtype": "f64", "data": "synthetic",
            "ns_per_day": 86400.0 * (args.steps / (ms_total * 1e-3)) * 1e-6,
            "config": {"workload": f"bcc Fe ANNP weak-scaling cell: {cells}^3 bcc cells = {nlocal} atoms per GPU, "
                                   f"{natoms_total} atoms total, NVE dt=1 fs, 300 K, skin 2 A, PBC (BASELINE configs[4])",
                       "potential": "fe_annp_potential_2 (28 SF = 9 radial + 19 angular, 28-10-10-1, Rc 6.5 A)",
                       "atoms_per_gpu": nlocal, "ghosts_per_gpu": md.nghost, "decomposition": "x".join(map(str, grid)),
                       "neighbors_in_cutoff": st.avg_neigh_cut, "list_neighbors": st0.max_neigh_list,
                       "l2_policy": "inputs larger than L2 (every step streams the 0.49 GB neighbour list of 524288 x 234 entries; L2 is 126 MB)"},
            "roofline": {"bound": "fp64", "kernel": "annp_force_kernel<9,19>", "achieved": achieved_tf, "peak": peak_tf,
                         "unit": "TFLOP/s", "frac": achieved_tf / peak_tf if peak_tf > 0 else None,
                         # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch at exactly this workload
                         # (profiles/r1c_force_kernel.md); algorithmic bytes are ~0.94 kB/atom = 0.49 GB
                         "traffic": NCU_TRAFFIC_BYTES if (cells == 64 and world == 1) else None,
                         "kernel_ms": kern_ms, "kernel_share_of_step": kern_ms / (ms_total / args.steps),
                         "flop_per_atom_step": flop_per_launch / nlocal,
                         "executed": ({"flop_per_atom_step": NCU_EXECUTED_FLOP_PER_ATOM_STEP,
                                       "tflops": NCU_EXECUTED_FLOP_PER_ATOM_STEP * nlocal / (kern_ms * 1e-3) / 1e12,
                                       "frac": NCU_EXECUTED_FLOP_PER_ATOM_STEP * nlocal / (kern_ms * 1e-3) / 1e12 / peak_tf,
                                       "note": "FP64 flops actually issued (ncu opcode counts of profiles/r1c_force_kernel.md): the algorithmic "
                                               "count credits 278 flops per triplet, the kernel executes ~150"}
                                      if (cells == 64 and kern_ms > 0 and peak_tf > 0) else None),
                         "peak_source": "measured on this GPU by annp_b200_fp64_peak_tflops (pure DFMA loop); "
                                        "MEASURED_PEAKS.json has no FP64 entry"},
            "e2e": {"value": e2e_value, "unit": "atom-steps/s", "h2d_bytes_per_step": nall * (24 + 4), "d2h_bytes_per_step": nall * 24 + 8,
                    "path": "annp_b200_compute (host x/type -> device, f/energy -> host, every step)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "cpu_baseline": cpu_base,
            "published_deck": published,
        }
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def run_published_deck(pot_file, device, steps=1000):
    """The one run the reference publishes a speed for: its own 152 880-atom bcc-Fe slab fe_st.dat,
    `boundary m p m`, one cg minimiser iteration, `velocity all create 300 4928459`, `fix npt temp 300 300 0.1 y 0 0 1`,
    dt 1 fs, thermo every step, 1 000 steps - loop time 1 789.44 s on the authors' 2 GPUs = 8.55e4 atom-steps/s
    (log_relaxing_new.lammps:1168-1176).  scripts/replay_published_deck.py replays that deck on ONE B200, device
    resident, and the thermo columns land on the log's to its printed precision; the timed region is LAMMPS' "Loop
    time" region (the 1 000 steps incl. per-step thermo, without setup and minimisation)."""
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import replay_published_deck as R
    ours, ref, mini, log, rebuilds = R.replay(steps=steps, device_index=device.index)
    secs = mini["loop_seconds"]
    natoms = 152880
    col = {str(c): i for i, c in enumerate(log["columns"])}
    return {"deck": "fe_st.dat 152 880 atoms, boundary m p m, cg minimiser step, velocity create 300 4928459, fix npt temp 300 300 0.1 y 0 0 1, "
                    "thermo 1, 1000 steps (zip:in.st_test)",
            "steps": steps, "seconds": secs, "atom_steps_per_s": natoms * steps / secs, "ns_per_day": 86400.0 * steps / secs * 1e-6,
            "published_atom_steps_per_s": PUBLISHED_ATOM_STEPS_PER_S, "published_seconds": 1789.44,
            "published_setup": "2 MPI ranks x 2 GPUs (RTX A5000 class), LAMMPS 29Sep2021, annp/gpu fe_v2",
            "ratio_vs_published": natoms * steps / secs / PUBLISHED_ATOM_STEPS_PER_S, "rebuilds": rebuilds,
            "T_final": float(ours[-1, col["Temp"]]), "T_final_log": float(ref[-1, col["Temp"]]),
            "Ly_final": float(ours[-1, col["Ly"]]), "Ly_final_log": float(ref[-1, col["Ly"]]),
            "Press_final_bar": float(ours[-1, col["Press"]]), "Press_final_log_bar": float(ref[-1, col["Press"]]),
            "max_rel_dT_over_run": float(np.abs(ours[:, col["Temp"]] / ref[:, col["Temp"]] - 1.0).max())}


def split_config(cfg, nparts):
    """Cut a Config into `nparts` rank-like pieces: chunk of centre atoms + every atom their rows touch."""
    from meng_zhang_b200 import lattice as L
    order = np.lexsort((cfg.x[: cfg.nlocal, 2], cfg.x[: cfg.nlocal, 1], cfg.x[: cfg.nlocal, 0]))
    off = cfg.offsets
    parts = []
    for chunk in np.array_split(order, nparts):
        if len(chunk) == 0:
            continue
        rows = [cfg.neigh[off[i]:off[i + 1]] for i in chunk]
        touched = np.unique(np.concatenate(rows + [chunk.astype(np.int32)]))
        ghosts = np.setdiff1d(touched, chunk, assume_unique=False)
        new_index = np.full(cfg.nall, -1, dtype=np.int64)
        new_index[chunk] = np.arange(len(chunk))
        new_index[ghosts] = len(chunk) + np.arange(len(ghosts))
        x = np.concatenate([cfg.x[chunk], cfg.x[ghosts]])
        typ = np.concatenate([cfg.type[chunk], cfg.type[ghosts]])
        neigh = new_index[np.concatenate(rows)].astype(np.int32)
        numneigh = np.array([len(r) for r in rows], dtype=np.int32)
        parts.append(L.Config(nlocal=len(chunk), nghost=len(ghosts), x=np.ascontiguousarray(x), type=typ.astype(np.int32),
                              ghost_owner=np.zeros(len(ghosts), dtype=np.int32), ilist=np.arange(len(chunk), dtype=np.int32),
                              numneigh=numneigh, neigh=neigh, box=cfg.box))
    return parts


def cpu_baseline_reference(sample_cells=10, repeats=1, cores=None):
    """The reference CPU pair style (unmodified source, oracle/_ref/ref_annp_fe) on all host cores:
    `mpirun -np P` is emulated by P independent processes on spatial chunks (compute() has no
    communication inside).  Returns the cpu_baseline object of the bench line."""
    import util
    from meng_zhang_b200 import lattice as L
    from oracle import run_ref, restatement
    from concurrent.futures import ThreadPoolExecutor
    cores = cores or os.cpu_count() or 1
    x = lattice_block(sample_cells, [0, 0, 0], 0.05, 1000)
    box = np.array([sample_cells * A_FE] * 3)
    cfg = L.build_config(x, box, RC, SKIN)
    natoms = cfg.nlocal
    pot_file = util.write_fe_potential(os.path.join(tempfile.gettempdir(), "annp_b200_bench_fe_cpu.ann"))
    if run_ref.available("annp_fe"):
        parts = split_config(cfg, cores)
        t0 = time.perf_counter()
        for _ in range(repeats):
            with ThreadPoolExecutor(max_workers=len(parts)) as ex:
                list(ex.map(lambda c: run_ref.run_reference("annp_fe", c, pot_file, ["Fe"], eflag=1, vflag=0), parts))
        dt = (time.perf_counter() - t0) / repeats
        kind = "reference"
        how = f"unmodified fe_v2/src/pair_annp.cpp, {len(parts)} processes x 1 thread on spatial chunks"
    else:
        from meng_zhang_b200.pair import read_potential
        pot = read_potential(pot_file, ["Fe"])
        t0 = time.perf_counter()
        for _ in range(repeats):
            restatement.compute(pot, cfg, nthreads=cores)
        dt = (time.perf_counter() - t0) / repeats
        kind = "port"
        how = f"oracle/annp_oracle.c with {cores} OpenMP threads"
    return {"value": natoms / dt, "unit": "atom-steps/s", "cores": cores, "kind": kind,
            "sample": f"{sample_cells}^3 bcc cells = {natoms} atoms of the same lattice (mean of {repeats} force evaluation(s), each incl. process start-up), {how}",
            "seconds": dt}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t_all = time.perf_counter()
    vals = []
    for _ in range(args.warmup if args.warmup < 2 else 1):
        cpu_baseline_reference(sample_cells=args.ref_cells)
    steps = max(1, args.steps)
    base = None
    for _ in range(steps):
        base = cpu_baseline_reference(sample_cells=args.ref_cells)
        vals.append(base["seconds"])
        if time.perf_counter() - t_all > 240:
            break
    natoms = 2 * args.ref_cells ** 3
    v = natoms * len(vals) / sum(vals)
    base["value"] = v
    out = {"impl": "reference", "metric": "atom-steps/sec (bcc Fe ANNP)", "value": v, "unit": "atom-steps/s",
           "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": len(vals), "warmup": 1, "ms_per_step": 1e3 * sum(vals) / len(vals),
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": f"bcc Fe ANNP weak-scaling cell (BASELINE configs[4]); reference CPU pair style timed on a bounded "
                                  f"sample: {args.ref_cells}^3 cells = {natoms} atoms of the same lattice per step"},
           "cpu_baseline": base,
           "e2e": {"value": v, "unit": "atom-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--cells", type=int, default=64, help="bcc cells per edge per GPU (64 -> 524288 atoms)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-cells", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-published-deck", action="store_true", help="skip the reference's own published deck (N=1 only)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

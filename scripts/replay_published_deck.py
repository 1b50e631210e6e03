"""Replay the reference's one published run without LAMMPS and compare with its own thermo log, step by step.

    python scripts/replay_published_deck.py [--steps 1000] [--out profiles/r1b_published_deck_replay.json]

Deck (performance test.zip: in.st_test): read_data fe_st.dat (152 880 atoms), boundary m p m, pair_style annp/gpu,
minimize 1e-6 1e-6 1000 10000 (the log shows ONE cg iteration = one line-search step of alpha = dmax/|f|_max),
velocity all create 300 4928459, fix npt temp 300 300 0.1 y 0 0 1, thermo 1, run 1000.
Everything LAMMPS itself contributes is restated in meng_zhang_b200/lammps_compat.py (Park-Miller velocity generator,
shrink-wrapped box) and csrc/annp_nh.cu (FixNH); the forces are the CUDA path.  The log comes from the reference's
mixed-precision GPU build (its forces differ from its own FP64 CPU style by ~2e-5 eV/A), so agreement is expected to
~6-7 digits at the start and to thermodynamic accuracy once the trajectories decorrelate.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import util  # noqa: E402
from meng_zhang_b200.lammps_compat import velocity_create  # noqa: E402
from meng_zhang_b200.md import DomainMD  # noqa: E402
from meng_zhang_b200.pair import PairANNPGPU  # noqa: E402


def replay(steps=1000, device_index=None):
    """Returns (ours[steps+1,12], log[steps+1,12], minimiser summary, log arrays, rebuilds).  Under torch.distributed
    (one rank per GPU, as the reference's `processors 2 1 1`) the slab is brick-decomposed; every rank returns the same
    (all-reduced) thermo table."""
    import torch.distributed as dist
    from meng_zhang_b200.md import decompose, rank_coords
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    if world > 1:
        device_index = int(os.environ.get("LOCAL_RANK", 0))
        torch.cuda.set_device(device_index)
        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=torch.device("cuda", device_index))
    grid = decompose(world)
    z = np.load(os.path.join(util.GOLDEN, "fe_st.npz"))
    log = np.load(os.path.join(util.GOLDEN, "fe_st_log.npz"))
    box = z["box"]
    boxlen = box[:, 1] - box[:, 0]
    x_all = z["x"] - box[:, 0]                                      # atoms in ID order
    n_all = len(x_all)
    # brick ownership as md._migrate assigns it (outermost bricks own what lies beyond a free face)
    cell = np.floor(x_all / (boxlen / np.array(grid))).astype(np.int64)
    cell = np.clip(cell, 0, np.array(grid) - 1)
    owner = cell[:, 0] + cell[:, 1] * grid[0] + cell[:, 2] * grid[0] * grid[1]
    gid = np.nonzero(owner == rank)[0]
    x0 = x_all[gid]
    n = len(x0)
    pair = PairANNPGPU(ntypes=1, device=-1 if device_index is None else device_index)
    pair.settings([])
    pair.coeff(["*", "*", util.write_fe_potential(f"/tmp/annp_b200_replay_fe_{rank}.ann"), "Fe"])
    pair.init_style()
    kw = dict(mass=55.845, dt=0.001, periodic=(False, True, False), shrink_wrap=(True, False, True), grid=grid, rank=rank,
              device=None if device_index is None else torch.device("cuda", device_index), gid_local=gid)

    def gsum(v):
        t = torch.tensor([float(v)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t)
        return float(t)

    def gmax(v):
        t = torch.tensor([float(v)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    # ---- minimize 1.0e-6 1.0e-6 1000 10000 (min_style cg): DomainMD.minimize restates MinCG + linemin_quadratic.  On this
    # input it stops after ONE iteration on the energy tolerance with two line-search evaluations, as the log reports;
    # the log's "alpha" is the trial step dmax / max|f|, the atoms end at the secant-projected alpha0.
    md = DomainMD(pair, x0, boxlen, **kw)
    md.reneighbor()
    assert np.array_equal(md.gid.cpu().numpy(), gid)                # nobody migrates at the first build
    mini = md.minimize(1.0e-6, 1.0e-6, 1000, 10000)
    mini["n_gpus"] = world
    x1 = md.x[: md.nlocal].clone()
    md = DomainMD(pair, x1.cpu().numpy(), boxlen, **kw)
    md.v = torch.as_tensor(velocity_create(n_all, 55.845, 300.0, 4928459)[gid], dtype=torch.float64, device=md.dev)
    md.reneighbor()
    md.fix_nh(300.0, 300.0, 0.1, p_flag=(0, 1, 0), p_start=(0.0,) * 3, p_stop=(0.0,) * 3, p_damp=(1.0,) * 3)
    pe0 = gsum(md.engvir[0])
    st = md.nh_state()
    nk = 1.6021765e6
    b0 = [st.boxhi[d] - st.boxlo[d] for d in range(3)]
    vol0 = b0[0] * b0[1] * b0[2]
    p0 = [(st.ke_tensor[d] + st.virial[d]) / vol0 * nk for d in range(3)]
    rows = [[0, st.t_current, pe0, 0.5 * sum(st.ke_tensor[:3]), *b0, sum(p0) / 3.0, vol0, *p0]]
    torch.cuda.synchronize(md.dev)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for s, pe, ke, ext, T, p, b in md.run_nh(steps, check_every=5, thermo_every=1):      # LAMMPS' "Loop time" region
        rows.append([s, T, pe, ke, *b, sum(p) / 3.0, b[0] * b[1] * b[2], *p])
    ev1.record()
    torch.cuda.synchronize(md.dev)
    mini["loop_seconds"] = gmax(ev0.elapsed_time(ev1) * 1e-3)
    ours = np.array(rows)
    ref = log["thermo_new"][: len(ours)]
    pair.clear()
    return ours, ref, mini, log, md.rebuilds


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    ours, ref, mini, log, rebuilds = replay(a.steps)
    cols = [str(c) for c in log["columns"]]
    rel = lambda k, sl: float(np.abs(ours[sl, k] / ref[sl, k] - 1.0).max())
    absd = lambda k, sl: float(np.abs(ours[sl, k] - ref[sl, k]).max())
    res = {"steps": int(a.steps), "rebuilds": rebuilds, "minimizer": mini,
           "minimizer_log": {"energy_initial_final": log["min_energy_initial_final_new"].tolist(), "fnorm_initial_final": log["min_fnorm_initial_final_new"].tolist(),
                             "fmax_initial_final": log["min_fmax_initial_final_new"].tolist(), "alpha_maxmove": log["min_alpha_maxmove_new"].tolist()},
           "step0": dict(zip(cols, ours[0].tolist())), "step0_log": dict(zip(cols, ref[0].tolist())),
           "last": dict(zip(cols, ours[-1].tolist())), "last_log": dict(zip(cols, ref[-1].tolist()))}
    for name, sl in (("steps_0_20", slice(0, 21)), ("steps_0_100", slice(0, 101)), ("steps_0_end", slice(0, None))):
        res[name] = {"max_rel_dT": rel(1, sl), "max_rel_dKinEng": rel(3, sl), "max_abs_dLx": absd(4, sl), "max_abs_dLy": absd(5, sl), "max_abs_dLz": absd(6, sl),
                     "max_abs_dPress_bar": absd(7, sl), "max_abs_dPyy_bar": absd(10, sl), "max_rel_dVolume": rel(8, sl),
                     "max_abs_dPotEng_eV": absd(2, sl)}
    if int(os.environ.get("RANK", 0)) != 0:
        return
    res["loop_seconds"] = mini["loop_seconds"]
    res["atom_steps_per_s"] = 152880 * a.steps / mini["loop_seconds"]
    print(json.dumps(res, indent=1))
    if a.out:
        with open(a.out, "w") as fp:
            json.dump(res, fp, indent=1)
        np.savez_compressed(os.path.splitext(a.out)[0] + "_thermo.npz", ours=ours, log=ref, columns=np.array(cols))


if __name__ == "__main__":
    main()

"""NVE energy drift of BASELINE config 1 (bcc Fe ANNP, 10x10x10 cells = 2 000 atoms, dt = 1 fs, 300 K start):
the reference's CPU algorithm against the CUDA path, from the SAME initial positions and velocities.

    python scripts/nve_drift.py --impl oracle --steps 10000 --out profiles/nve_drift_reference.json   (CPU, hours)
    python scripts/nve_drift.py --impl gpu    --steps 10000 --out profiles/nve_drift_gpu.json         (GPU box)
    python scripts/nve_drift.py --compare profiles/nve_drift_reference.json profiles/nve_drift_gpu.json

`--impl oracle` integrates with forces from oracle/annp_oracle.c, which is bit-identical to the unmodified reference
pair style (tests/test_oracle.py); running the reference binary itself would cost ~24 s per step (its O(nall)
allocation per atom).  Both sides use velocity Verlet in LAMMPS metal units and the deck's
`neigh_modify every 5 delay 5 check yes` re-neighbouring rule with a 2 A skin.  The oracle is only the checker here:
the GPU leg never touches it.
Drift metric: |E_tot(t) - E_tot(0)| / N, and a least-squares slope of E_tot/N in eV/atom/ps.
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

KB = 8.617343e-5
MVV2E = 1.0364269e-4
FTM2V = 1.0 / 1.0364269e-4
MASS = 55.845
DT = 0.001
SKIN = 2.0
RC = 6.5


def initial_state(cells, temperature, seed):
    from meng_zhang_b200 import lattice as L
    x, box = L.bcc(cells, cells, cells)
    rng = np.random.default_rng(seed)
    sigma = (KB * temperature / (MASS * MVV2E)) ** 0.5
    v = rng.standard_normal(x.shape) * sigma
    v -= v.mean(axis=0, keepdims=True)
    return x, box, v


def summarize(trace, natoms):
    t = np.array([r[0] for r in trace]) * DT                     # ps
    e = np.array([r[1] + r[2] for r in trace]) / natoms
    slope = float(np.polyfit(t, e, 1)[0]) if len(t) > 2 else 0.0
    return {"max_abs_dE_per_atom": float(np.abs(e - e[0]).max()), "final_dE_per_atom": float(e[-1] - e[0]),
            "slope_eV_per_atom_per_ps": slope, "rms_fluctuation_per_atom": float(np.std(e - np.polyval(np.polyfit(t, e, 1), t)))}


def run_oracle(args):
    import util
    from meng_zhang_b200 import lattice as L
    from meng_zhang_b200.pair import read_potential
    from oracle import restatement
    pot = read_potential(util.write_fe_potential("/tmp/annp_b200_drift_fe.ann"), ["Fe"])
    x, box, v = initial_state(args.cells, args.temperature, args.seed)
    n = len(x)
    dtf = 0.5 * DT * FTM2V / MASS
    trace, rebuilds = [], 0
    t0 = time.time()

    def rebuild(xw):
        cfg = L.build_config(xw, box, RC, SKIN)
        return cfg, cfg.x[:n].copy()

    def force(cfg, xl):
        xa = np.concatenate([xl, xl[cfg.ghost_owner] + cfg.ghost_shift])
        c = L.Config(**{**cfg.__dict__, "x": np.ascontiguousarray(xa)})
        o = restatement.compute(pot, c, vflag=False, nthreads=args.threads)
        return c.fold(o["f"]), o["eng_vdwl"]

    cfg, x = rebuild(x)
    x_ref = x.copy()
    f, pe = force(cfg, x)
    trace.append((0, pe, 0.5 * MASS * MVV2E * float((v * v).sum())))
    for step in range(1, args.steps + 1):
        v += dtf * f
        x += DT * v
        if step % 5 == 0 and float(((x - x_ref) ** 2).sum(axis=1).max()) > (0.5 * SKIN) ** 2:
            cfg, x = rebuild(x)
            x_ref = x.copy()
            rebuilds += 1
        f, pe = force(cfg, x)
        v += dtf * f
        if step % args.every == 0 or step == args.steps:
            trace.append((step, pe, 0.5 * MASS * MVV2E * float((v * v).sum())))
            if args.out:
                dump(args, trace, n, rebuilds, time.time() - t0, "oracle (annp_oracle.c, bit-identical to fe_v2/src/pair_annp.cpp)")
    dump(args, trace, n, rebuilds, time.time() - t0, "oracle (annp_oracle.c, bit-identical to fe_v2/src/pair_annp.cpp)")


def run_gpu(args):
    import torch
    import util
    from meng_zhang_b200.md import DomainMD
    from meng_zhang_b200.pair import PairANNPGPU
    pot_file = util.write_fe_potential("/tmp/annp_b200_drift_fe.ann")
    x, box, v = initial_state(args.cells, args.temperature, args.seed)
    pair = PairANNPGPU(ntypes=1, skin=SKIN)
    pair.settings([])
    pair.coeff(["*", "*", pot_file, "Fe"])
    pair.init_style()
    md = DomainMD(pair, x, box, skin=SKIN, mass=MASS, dt=DT)
    md.v = torch.as_tensor(v, dtype=torch.float64, device=md.dev)
    md.reneighbor()
    md.compute(eflag=True)
    torch.cuda.synchronize()
    pe0 = float(md.engvir[0])
    trace = [(0, pe0, 0.5 * MASS * MVV2E * float((md.v * md.v).sum()))]
    t0 = time.time()
    out = md.run(args.steps, check_every=5, thermo_every=args.every)
    torch.cuda.synchronize()
    trace += [(s, pe, ke) for s, pe, ke in out]
    dump(args, trace, len(x), md.rebuilds, time.time() - t0, "libannp_b200.so, device-resident MD (DomainMD.run)")


def dump(args, trace, natoms, rebuilds, secs, impl):
    res = {"impl": impl, "natoms": natoms, "steps": trace[-1][0], "dt_ps": DT, "temperature0": args.temperature, "seed": args.seed,
           "rebuilds": rebuilds, "wall_seconds": secs, "summary": summarize(trace, natoms),
           "trace_step_pe_ke": [[int(s), float(pe), float(ke)] for s, pe, ke in trace]}
    if args.out:
        tmp = args.out + ".tmp"
        with open(tmp, "w") as fp:
            json.dump(res, fp)
        os.replace(tmp, args.out)
    return res


def compare(a_path, b_path):
    a, b = json.load(open(a_path)), json.load(open(b_path))
    ta = {r[0]: r for r in a["trace_step_pe_ke"]}
    tb = {r[0]: r for r in b["trace_step_pe_ke"]}
    common = sorted(set(ta) & set(tb))
    n = a["natoms"]
    de = [abs((ta[s][1] + ta[s][2]) - (tb[s][1] + tb[s][2])) / n for s in common]
    print(json.dumps({"reference": a["summary"], "gpu": b["summary"], "steps_compared": len(common),
                      "max_abs_diff_total_energy_per_atom": max(de), "first_100_steps_max_diff": max(de[: max(1, 100 // max(1, common[1] - common[0]))]),
                      "gpu_drift_no_worse": abs(b["summary"]["slope_eV_per_atom_per_ps"]) <= abs(a["summary"]["slope_eV_per_atom_per_ps"]) * 1.05 + 1e-9}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", choices=["oracle", "gpu"], default="gpu")
    ap.add_argument("--steps", type=int, default=10000)
    ap.add_argument("--every", type=int, default=10)
    ap.add_argument("--cells", type=int, default=10)
    ap.add_argument("--temperature", type=float, default=300.0)
    ap.add_argument("--seed", type=int, default=4928459)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 4)
    ap.add_argument("--out", default=None)
    ap.add_argument("--compare", nargs=2, default=None)
    a = ap.parse_args()
    if a.compare:
        compare(*a.compare)
    elif a.impl == "oracle":
        run_oracle(a)
    else:
        run_gpu(a)


if __name__ == "__main__":
    main()

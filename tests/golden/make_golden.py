"""Generate the committed golden fixtures from the UNMODIFIED reference (run in the build container).

    python tests/golden/make_golden.py

Needs /root/reference and the binaries of `make -C oracle` (oracle/_ref/ref_*).  Writes
  meng_zhang_b200/data/fe_potential.json   parameters of annp-gpu-lammps/fe_v2/fe_annp_potential_2.ann as parsed by
                                   the reference-compatible reader (numbers round-trip exactly; the test
                                   suite re-creates a `.ann` file from it with pair.write_potential)
  tests/golden/annp_fe_*.npz       inputs (x, type, ghosts, neighbour rows) and the reference's outputs
                                   (eng_vdwl, eatom, f, virial by pair tally and by f.r, vatom)
  meng_zhang_b200/data/{ni,anna}_potential.json, tests/golden/annp_ni_*.npz, anna_adp_*.npz
                                   the same for the Ni copy of the style (ref_annp_ni) and for ANNA-ADP (ref_anna_adp)
  tests/golden/annp_general_*.npz  synthetic potentials (other shapes, activations, element counts) + the reference's answers
  tests/golden/fe_st.npz           the 152 880-atom slab of `performance test.zip` + its logged thermo values
The GPU box has no /root/reference: tests read only these files.
"""
import io
import json
import os
import sys
import zipfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from meng_zhang_b200 import lattice as L  # noqa: E402
from meng_zhang_b200.pair import read_potential  # noqa: E402
from oracle.run_ref import run_reference  # noqa: E402

REF = "/root/reference"
POT = f"{REF}/annp-gpu-lammps/fe_v2/fe_annp_potential_2.ann"
OUT = os.path.dirname(os.path.abspath(__file__))
DATA = os.path.join(ROOT, "meng_zhang_b200", "data")      # the parsed potentials ship with the package


def dump_potential():
    pot = read_potential(POT, ["Fe"])
    d = {k: getattr(pot, k) for k in ("nelements", "ntl", "nhl", "nnod", "nsf", "npsf", "ntsf", "flagsym", "flagact",
                                      "cut", "e_scale", "e_shift", "e_atom", "id_elem", "mass", "elements")}
    d["sfnor_cov"] = pot.sfnor_cov.tolist()
    d["sfnor_avg"] = pot.sfnor_avg.tolist()
    d["weight_all"] = pot.weight_all.tolist()
    d["bias_all"] = pot.bias_all.tolist()
    d["source"] = "annp-gpu-lammps/fe_v2/fe_annp_potential_2.ann (MPL-2.0), parsed numbers only"
    with open(os.path.join(DATA, "fe_potential.json"), "w") as fp:
        json.dump(d, fp)


def cases():
    # name -> Config (+ element list)
    out = {}
    x, box = L.bcc(4, 4, 4)
    out["bcc4_perfect"] = (L.build_config(x, box, 6.5), ["Fe"])
    out["bcc4_perturbed"] = (L.build_config(L.perturb(x, 0.05, 12345), box, 6.5, shuffle_rows=7), ["Fe"])
    x3, box3 = L.bcc(3, 3, 4)   # box edge 8.57 A: barely above the 8.5 A ghost cutoff, many self images
    out["bcc334_hot"] = (L.build_config(L.perturb(x3, 0.15, 99), box3, 6.5, shuffle_rows=3), ["Fe"])
    # free cluster: ragged rows, atoms with few or zero neighbours inside Rc, one isolated atom
    xc, _ = L.bcc(3, 3, 3)
    xc = L.perturb(xc, 0.08, 5)
    xc = np.concatenate([xc, [[40.0, 40.0, 40.0]], [[46.0, 40.0, 40.0]], [[46.0, 44.5, 40.0]]])
    out["cluster_ragged"] = (L.build_config(xc, np.array([100.0, 100.0, 100.0]), 6.5, periodic=(False, False, False)), ["Fe"])
    # two LAMMPS types mapped to the same element (as the STGB generator's two grains)
    rng = np.random.default_rng(11)
    types = rng.integers(1, 3, size=128).astype(np.int32)
    out["bcc4_two_types"] = (L.build_config(L.perturb(x, 0.05, 777), box, 6.5, types=types), ["Fe", "Fe"])
    # BASELINE config 1 geometry: 10x10x10 cells = 2000 atoms
    x10, box10 = L.bcc(10, 10, 10)
    out["bcc10_perturbed"] = (L.build_config(L.perturb(x10, 0.05, 2024), box10, 6.5), ["Fe"])
    return out


def dump_cases():
    for name, (cfg, elems) in cases().items():
        r1 = run_reference("annp_fe", cfg, POT, elems, eflag=3, vflag=1 + 4)   # pair tally + per-atom virial
        r2 = run_reference("annp_fe", cfg, POT, elems, eflag=3, vflag=2)       # virial through f.r
        assert np.array_equal(r1["f"], r2["f"])
        np.savez_compressed(
            os.path.join(OUT, f"annp_fe_{name}.npz"),
            nlocal=cfg.nlocal, nghost=cfg.nghost, x=cfg.x, type=cfg.type, ghost_owner=cfg.ghost_owner,
            ilist=cfg.ilist, numneigh=cfg.numneigh, neigh=cfg.neigh, box=cfg.box, elements=np.array(elems),
            eng_vdwl=r1["eng_vdwl"], eatom=r1["eatom"], f=r1["f"], virial_pair=r1["virial"], virial_fdotr=r2["virial"],
            vatom=r1["vatom"], ref_seconds=r1["seconds"])
        print(f"{name}: nlocal {cfg.nlocal} nghost {cfg.nghost} E/atom {r1['eng_vdwl'] / cfg.nlocal:.10f} "
              f"fmax {np.abs(cfg.fold(r1['f'])).max():.6e} ({r1['seconds']:.1f} s)")


def dump_fe_st():
    z = zipfile.ZipFile(f"{REF}/annp-gpu-lammps/fe_v2/performance test.zip")
    txt = z.read("performance comparsion/fe_st.dat").decode()
    lines = txt.splitlines()
    natoms = int(lines[1].split()[0])
    box = np.array([[float(v) for v in lines[3 + d].split()[:2]] for d in range(3)])
    start = next(i for i, l in enumerate(lines) if l.startswith("Atoms")) + 2
    arr = np.loadtxt(io.StringIO("\n".join(lines[start:start + natoms])))
    assert arr.shape == (natoms, 5)
    order = np.argsort(arr[:, 0].astype(np.int64))
    x = arr[order, 2:5]
    # thermo values of the reference's own 2-GPU annp/gpu run (log_relaxing_new.lammps:109,117-120;
    # log_relaxing_old.lammps:111,119-122); boundary m p m, 217.55887 neighbours/atom at 8.5 A
    # The logged run is the reference's GPU build (LAL precision not logged; it differs from the reference's
    # own FP64 CPU algorithm by 4.9e-9 in E and 1.9e-5 eV/A in max|F|).  The FP64 answer for this input is
    # produced with the restatement (bit-identical to the reference CPU source on every small case; the
    # reference binary itself needs ~12 h here because of its O(nall) allocation per atom).
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_gpu_parity import build_large_config
    from oracle import restatement
    cfg = build_large_config(x - box[:, 0], box[:, 1] - box[:, 0], (False, True, False))
    o = restatement.compute(read_potential(POT, ["Fe"]), cfg, nthreads=os.cpu_count() or 4)
    ff = cfg.fold(o["f"])
    idx = np.arange(0, len(ff), 997)
    np.savez_compressed(os.path.join(OUT, "fe_st.npz"), x=x.astype(np.float64), box=box,
                        e_cpu_fp64=o["eng_vdwl"], fnorm_cpu_fp64=np.linalg.norm(ff), fmax_cpu_fp64=np.abs(ff).max(),
                        virial_cpu_fp64=o["virial"], f_sample_idx=idx, f_sample_cpu_fp64=ff[idx],
                        e_pair_new=-684876292.365723, e_pair_old=-684876292.28418,
                        fnorm_new=39.623051, fnorm_old=39.623117, fmax_new=0.93490135, fmax_old=0.93490485,
                        press_new=-40423.638, press_old=-40426.438, volume=1773495.9, neighs_per_atom=217.55887)
    print("fe_st:", natoms, box.tolist())


NI_POT = f"{REF}/annp-gpu-lammps/ni/ni_annp_potential_2.ann"
ANNA_POT = f"{REF}/anna-gpu-lammps/bcc_fe/fe_adp_potential_2310.anna"


def dump_ni_anna_potentials():
    """Parsed numbers of the Ni `.ann` and the ANNA-ADP `.anna` file (tests re-create the files with the writers)."""
    from meng_zhang_b200.pair_anna import read_anna_potential
    pot = read_potential(NI_POT, ["Ni"])
    d = {k: getattr(pot, k) for k in ("nelements", "ntl", "nhl", "nnod", "nsf", "npsf", "ntsf", "flagsym", "flagact",
                                      "cut", "e_scale", "e_shift", "e_atom", "id_elem", "mass", "elements")}
    for k in ("sfnor_cov", "sfnor_avg", "weight_all", "bias_all", "sym_coerad", "sym_coeang"):
        d[k] = getattr(pot, k).tolist()
    d["source"] = "annp-gpu-lammps/ni/ni_annp_potential_2.ann (MPL-2.0), parsed numbers only"
    with open(os.path.join(DATA, "ni_potential.json"), "w") as fp:
        json.dump(d, fp)
    pa = read_anna_potential(ANNA_POT, ["Fe"])
    d = {k: getattr(pa, k) for k in ("nelements", "ntl", "nhl", "nnod", "nout", "nsf", "npsf", "ntsf", "flagsym", "flagact",
                                     "cut", "e_base", "e_scal", "ngp", "id_elem", "mass", "elements")}
    for k in ("gparams", "weight_all", "bias_all"):
        d[k] = getattr(pa, k).tolist()
    d["source"] = "anna-gpu-lammps/bcc_fe/fe_adp_potential_2310.anna (MPL-2.0), parsed numbers only"
    with open(os.path.join(DATA, "anna_potential.json"), "w") as fp:
        json.dump(d, fp)


def ni_cases():
    """Ni copy: thermally perturbed cells only - in a perfect lattice 1 + lambda cos(theta) = 0 sits on the reference's
    `flag <= 0` branch and the answer depends on the last bit of cos(theta)."""
    out = {}
    x, box = L.fcc(3, 3, 3)
    out["fcc3_perturbed"] = (L.build_config(L.perturb(x, 0.05, 12345), box, 6.5, shuffle_rows=7), ["Ni"])
    x2, box2 = L.fcc(3, 3, 4)
    out["fcc334_hot"] = (L.build_config(L.perturb(x2, 0.15, 99), box2, 6.5, shuffle_rows=3), ["Ni"])
    xc, _ = L.fcc(2, 2, 2)
    xc = L.perturb(xc, 0.08, 5)
    xc = np.concatenate([xc, [[40.0, 40.0, 40.0]], [[46.0, 40.0, 40.0]], [[42.5, 40.0, 40.0]], [[41.2, 42.0, 40.3]]])
    out["cluster_ragged"] = (L.build_config(xc, np.array([100.0, 100.0, 100.0]), 6.5, periodic=(False, False, False)), ["Ni"])
    rng = np.random.default_rng(11)
    types = rng.integers(1, 3, size=len(x)).astype(np.int32)
    out["fcc3_two_types"] = (L.build_config(L.perturb(x, 0.05, 777), box, 6.5, types=types), ["Ni", "Ni"])
    return out


def anna_cases():
    out = {}
    rc = 5.055
    x, box = L.bcc(4, 4, 4)
    out["bcc4_perfect"] = (L.build_config(x, box, rc), ["Fe"])
    out["bcc4_perturbed"] = (L.build_config(L.perturb(x, 0.05, 12345), box, rc, shuffle_rows=7), ["Fe"])
    x3, box3 = L.bcc(3, 3, 4)
    out["bcc334_hot"] = (L.build_config(L.perturb(x3, 0.15, 99), box3, rc, shuffle_rows=3), ["Fe"])
    xc, _ = L.bcc(3, 3, 3)
    xc = L.perturb(xc, 0.08, 5)
    xc = np.concatenate([xc, [[40.0, 40.0, 40.0]], [[44.5, 40.0, 40.0]], [[44.5, 43.5, 40.0]]])
    out["cluster_ragged"] = (L.build_config(xc, np.array([100.0, 100.0, 100.0]), rc, periodic=(False, False, False)), ["Fe"])
    rng = np.random.default_rng(11)
    types = rng.integers(1, 3, size=128).astype(np.int32)
    out["bcc4_two_types"] = (L.build_config(L.perturb(x, 0.05, 777), box, rc, types=types), ["Fe", "Fe"])
    x8, box8 = L.bcc(8, 8, 8)
    out["bcc8_perturbed"] = (L.build_config(L.perturb(x8, 0.05, 2024), box8, rc), ["Fe"])
    return out


def dump_kind(kind, prefix, potfile, cases_fn):
    for name, (cfg, elems) in cases_fn().items():
        r1 = run_reference(kind, cfg, potfile, elems, eflag=3, vflag=1 + 4)
        r2 = run_reference(kind, cfg, potfile, elems, eflag=3, vflag=2)
        assert np.array_equal(r1["f"], r2["f"])
        np.savez_compressed(
            os.path.join(OUT, f"{prefix}_{name}.npz"),
            nlocal=cfg.nlocal, nghost=cfg.nghost, x=cfg.x, type=cfg.type, ghost_owner=cfg.ghost_owner,
            ilist=cfg.ilist, numneigh=cfg.numneigh, neigh=cfg.neigh, box=cfg.box, elements=np.array(elems),
            eng_vdwl=r1["eng_vdwl"], eatom=r1["eatom"], f=r1["f"], virial_pair=r1["virial"], virial_fdotr=r2["virial"],
            vatom=r1["vatom"], ref_seconds=r1["seconds"])
        print(f"{prefix} {name}: nlocal {cfg.nlocal} nghost {cfg.nghost} E/atom {r1['eng_vdwl'] / cfg.nlocal:.10f} "
              f"fmax {np.abs(cfg.fold(r1['f'])).max():.6e} ({r1['seconds']:.1f} s)")


def dump_general():
    """Potentials the shipped files do not exercise, through the UNMODIFIED reference: activations 1, 2, 3
    and a non-linear output layer, a five-layer network, descriptor shapes (8,20), (6,11), (12,22), (16,24), two element
    blocks.  The potential files are written by pair.write_potential from the synthetic numbers of util.general_potential;
    the golden file stores those numbers next to the reference's answers."""
    import tempfile
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import util
    from meng_zhang_b200.pair import write_potential
    x, box = L.bcc(4, 4, 4)
    for name in util.GENERAL_CASES:
        pot, elems = util.general_potential(name)
        if len(elems) > 1:
            types = np.random.default_rng(11).integers(1, len(elems) + 1, size=len(x)).astype(np.int32)
            cfg = L.build_config(L.perturb(x, 0.05, 4711), box, pot.cut, types=types, shuffle_rows=5)
        else:
            cfg = L.build_config(L.perturb(x, 0.05, 4711), box, pot.cut, shuffle_rows=5)
        with tempfile.TemporaryDirectory() as td:
            pf = os.path.join(td, f"{name}.ann")
            write_potential(pf, pot, comment=f"synthetic potential {name}")
            r1 = run_reference("annp_fe", cfg, pf, elems, eflag=3, vflag=1 + 4)
            r2 = run_reference("annp_fe", cfg, pf, elems, eflag=3, vflag=2)
        assert np.array_equal(r1["f"], r2["f"])
        np.savez_compressed(
            os.path.join(OUT, f"annp_general_{name}.npz"),
            nlocal=cfg.nlocal, nghost=cfg.nghost, x=cfg.x, type=cfg.type, ghost_owner=cfg.ghost_owner,
            ilist=cfg.ilist, numneigh=cfg.numneigh, neigh=cfg.neigh, box=cfg.box, elements=np.array(elems),
            eng_vdwl=r1["eng_vdwl"], eatom=r1["eatom"], f=r1["f"], virial_pair=r1["virial"], virial_fdotr=r2["virial"],
            vatom=r1["vatom"], ref_seconds=r1["seconds"],
            pot_nelements=pot.nelements, pot_ntl=pot.ntl, pot_nnod=pot.nnod, pot_npsf=pot.npsf, pot_ntsf=pot.ntsf,
            pot_flagact=np.array(pot.flagact), pot_cut=pot.cut, pot_e_scale=pot.e_scale, pot_e_shift=pot.e_shift, pot_e_atom=pot.e_atom,
            pot_elements=np.array(pot.elements), pot_sfnor_cov=pot.sfnor_cov, pot_sfnor_avg=pot.sfnor_avg,
            pot_weight_all=pot.weight_all, pot_bias_all=pot.bias_all)
        ec = r1["eatom"][: cfg.nlocal] - pot.e_shift - pot.e_atom
        print(f"general {name}: E/atom {r1['eng_vdwl'] / cfg.nlocal:.10f} cohesive part {ec.min():.4f}..{ec.max():.4f} "
              f"fmax {np.abs(cfg.fold(r1['f'])).max():.6e} ({r1['seconds']:.1f} s)")


def dump_fe_st_log():
    """Thermo table of the reference's own published run (numbers only): 1 001 rows of
    Step Temp PotEng KinEng Lx Ly Lz Press Volume Pxx Pyy Pzz from log_relaxing_{new,old}.lammps, plus the minimiser
    summary (energies, force norms, line-search alpha) that precedes it."""
    z = zipfile.ZipFile(f"{REF}/annp-gpu-lammps/fe_v2/performance test.zip")
    out = {}
    for name in ("new", "old"):
        t = z.read(f"performance comparsion/log_relaxing_{name}.lammps").decode().splitlines()
        i0 = [i for i, l in enumerate(t) if l.startswith("Step Temp PotEng")][0]
        rows = []
        for l in t[i0 + 1:]:
            p = l.split()
            if len(p) != 12:
                break
            rows.append([float(v) for v in p])
        out[f"thermo_{name}"] = np.array(rows)
        assert out[f"thermo_{name}"].shape == (1001, 12)
    out["columns"] = np.array("Step Temp PotEng KinEng Lx Ly Lz Press Volume Pxx Pyy Pzz".split())
    out["min_energy_initial_final_new"] = np.array([-684876292.365723, -684876369.462402])      # log_relaxing_new.lammps:117
    out["min_fnorm_initial_final_new"] = np.array([39.623051, 19.978295])                        # :118
    out["min_fmax_initial_final_new"] = np.array([0.93490135, 0.52800152])                       # :119
    out["min_alpha_maxmove_new"] = np.array([0.10696316, 0.056476709])                           # :120
    # the input deck itself (1 kB of LAMMPS commands; run verbatim by meng_zhang_b200.deck in the GPU tests)
    with open(os.path.join(OUT, "in.st_test"), "wb") as fp:
        fp.write(z.read("performance comparsion/in.st_test"))
    np.savez_compressed(os.path.join(OUT, "fe_st_log.npz"), **out)
    print("fe_st_log:", out["thermo_new"][0], out["thermo_new"][-1][:6])


def dump_structures():
    """Atoms written by the reference's own structure generators (oracle/_ref/gen_screw, gen_stgb)."""
    import subprocess
    import tempfile
    from meng_zhang_b200 import structures as S
    gen = os.path.join(ROOT, "oracle", "_ref")

    def run(prog, *args):
        with tempfile.TemporaryDirectory() as td:
            out = os.path.join(td, "o.txt")
            subprocess.run([os.path.join(gen, prog), out, *map(str, args)], check=True, capture_output=True, cwd=td)
            with open(out) as fp:
                n = int(fp.readline())
                box = np.array(fp.readline().split(), dtype=float)
                arr = np.loadtxt(fp)
            assert len(arr) == n
            return arr, box

    arr, box = run("gen_screw")
    # the three atoms the reference asks its user for: two neighbours along x, the third on the vertex above them
    x, b, t = S.screw_block(reference_rules=True)
    core, cols = S.screw_core(x, b)
    ids = []
    for c in cols:
        k = np.argmin(np.linalg.norm(arr[:, 2:4] - c, axis=1))
        assert np.linalg.norm(arr[k, 2:4] - c) < 1e-5
        ids.append(int(arr[k, 0]))
    arr_s, _ = run("gen_screw", *ids)
    np.savez_compressed(os.path.join(OUT, "gen_screw.npz"), x=arr[:, 2:5], type=arr[:, 1].astype(np.int32), box=box,
                        ids=np.array(ids), core=core, x_screw=arr_s[:, 2:5])
    arr2, box2 = run("gen_stgb")
    np.savez_compressed(os.path.join(OUT, "gen_stgb.npz"), x=arr2[:, 2:5], type=arr2[:, 1].astype(np.int32), box=box2)
    print("structures:", len(arr), len(arr2), ids, core)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "log":
        dump_fe_st_log()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "general":     # only the round-2 additions
        dump_general()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "structures":
        dump_structures()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "ni_anna":      # only the round-1b additions
        dump_ni_anna_potentials()
        dump_kind("annp_ni", "annp_ni", NI_POT, ni_cases)
        dump_kind("anna_adp", "anna_adp", ANNA_POT, anna_cases)
        sys.exit(0)
    dump_potential()
    dump_ni_anna_potentials()
    dump_kind("annp_ni", "annp_ni", NI_POT, ni_cases)
    dump_kind("anna_adp", "anna_adp", ANNA_POT, anna_cases)
    dump_structures()
    dump_fe_st_log()
    dump_fe_st()
    dump_cases()
    dump_general()

"""Parity of the CUDA path for the other two copies of the reference style: the Ni `annp` copy (Behler-Parrinello
descriptor, BASELINE config 2) and ANNA-ADP (`pair_style anna_adp/gpu`).  Run with -m gpu.

Golden vectors: tests/golden/annp_ni_*.npz, anna_adp_*.npz, produced by the UNMODIFIED reference sources
(ref_annp_ni / ref_anna_adp, see make_golden.py).  Tolerances asserted (north_star: energy <= 1e-6 relative,
forces <= 1e-5 eV/A):
    Ni    per-atom energy <= 1e-10 (raw network output, ~0.76), forces <= 1e-9 eV/A, virial <= 1e-9
    ANNA  per-atom energy <= 1e-9 eV,  forces <= 1e-9 eV/A, virial <= 1e-8 eV
"""
import numpy as np
import pytest

import util
from meng_zhang_b200 import capi, lattice as L
from meng_zhang_b200.pair import PairANNPGPU, read_potential
from meng_zhang_b200.pair_anna import PairANNAADPGPU, read_anna_potential

pytestmark = pytest.mark.gpu


def make_ni(pot_file, elems=("Ni",), kernel="pair"):
    # variant decided from the file (coefficient blocks -> Ni copy); the flags force one of the three kernels:
    # "pair" (default for tiles <= 32: pair compaction), "lane" (lane per neighbour), "generic" (table driven, pow/exp)
    flag = {"pair": None, "lane": capi.VARIANT_NI | capi.VARIANT_FLAG_NOPAIR, "generic": capi.VARIANT_NI | capi.VARIANT_FLAG_GENERIC}[kernel]
    pair = PairANNPGPU(ntypes=len(elems), variant=flag)
    pair.settings([])
    pair.coeff(["*", "*", pot_file] + list(elems))
    pair.init_style()
    return pair


def make_anna(pot_file, elems=("Fe",)):
    pair = PairANNAADPGPU(ntypes=len(elems))
    pair.settings([])
    pair.coeff(["*", "*", pot_file] + list(elems))
    pair.init_style()
    return pair


@pytest.mark.parametrize("kernel", ["pair", "lane", "generic"])
@pytest.mark.parametrize("name", util.NI_CASES)
def test_ni_golden_case(name, kernel, ni_pot_file):
    cfg, elems, ref = util.load_case(name, "annp_ni")
    pair = make_ni(ni_pot_file, elems, kernel)
    f = pair.compute(3, 1 + 4, cfg, ago=0)
    assert np.abs(pair.eatom - ref["eatom"]).max() <= 1e-10
    assert abs(pair.eng_vdwl - ref["eng_vdwl"]) <= 1e-9
    assert np.abs(f - ref["f"]).max() <= 1e-9
    assert np.abs(pair.virial - ref["virial_pair"]).max() <= 1e-9
    assert np.abs(pair.vatom - ref["vatom"]).max() <= 1e-9
    f2 = pair.compute(3, 1 + 4, cfg, ago=1)
    assert np.array_equal(f, f2)                       # deterministic
    assert np.array_equal(f, pair.compute(0, 0, cfg, ago=1))
    # the ordered-gather scatter gives the same forces to the fixed-point resolution, and the same golden parity
    pair.set_scatter(capi.SCATTER_GATHER)
    fg = pair.compute(3, 1 + 4, cfg, ago=1)
    assert np.abs(fg - f).max() <= 2e-11 and np.abs(fg - ref["f"]).max() <= 1e-9
    assert np.array_equal(fg, pair.compute(0, 0, cfg, ago=1))
    pair.clear()


@pytest.mark.parametrize("name", util.ANNA_CASES)
def test_anna_golden_case(name, anna_pot_file):
    cfg, elems, ref = util.load_case(name, "anna_adp")
    pair = make_anna(anna_pot_file, elems)
    f = pair.compute(3, 1 + 4, cfg, ago=0)
    assert np.abs(pair.eatom - ref["eatom"]).max() <= 1e-9
    assert abs(pair.eng_vdwl - ref["eng_vdwl"]) <= 1e-12 * abs(ref["eng_vdwl"]) + 1e-9
    assert np.abs(f - ref["f"]).max() <= 1e-9
    assert np.abs(pair.virial - ref["virial_pair"]).max() <= 1e-8
    assert np.abs(pair.virial - ref["virial_fdotr"]).max() <= 1e-8
    assert np.abs(pair.vatom - ref["vatom"]).max() <= 1e-9
    f2 = pair.compute(3, 1 + 4, cfg, ago=1)
    assert np.array_equal(f, f2)
    assert np.array_equal(f, pair.compute(0, 0, cfg, ago=1))
    pair.set_scatter(capi.SCATTER_GATHER)
    fg = pair.compute(3, 1 + 4, cfg, ago=1)
    assert np.abs(fg - f).max() <= 2e-11 and np.abs(fg - ref["f"]).max() <= 1e-9
    pair.clear()


def test_ni_against_live_oracle_32000_atom_geometry_sample(ni_pot_file):
    """BASELINE config 2 geometry (fcc Ni, a = 3.52) at a size the oracle finishes in seconds: 6^3 cells = 864 atoms,
    thermal displacements."""
    from oracle import restatement
    x, box = L.fcc(6, 6, 6)
    cfg = L.build_config(L.perturb(x, 0.07, 2718), box, 6.5, shuffle_rows=5)
    pair = make_ni(ni_pot_file)
    f = pair.compute(3, 1, cfg, ago=0)
    ref = restatement.compute_ni(read_potential(ni_pot_file, ["Ni"]), cfg, nthreads=4, dump_G=True)
    assert np.abs(f - ref["f"]).max() <= 1e-9
    assert np.abs(pair.eatom - ref["eatom"]).max() <= 1e-10
    assert np.abs(pair.virial - ref["virial"]).max() <= 1e-8
    G, _ = pair.descriptors(cfg)
    assert np.abs(G - ref["G"]).max() <= 1e-11
    assert np.abs(f.sum(axis=0)).max() < 1e-9          # momentum conservation (each pair force is applied +/-)
    pair.clear()


def test_ni_forces_under_decomposition_follow_the_reference_on_each_ranks_own_list(ni_pot_file):
    """The Ni copy's forces depend on the ORDER of a neighbour row (r_ik where r_jk is meant, ni/src/pair_annp.cpp:734-735),
    so a decomposed run need not reproduce the single-domain forces (profiles/r1c_multi_gpu_check: 5e-3 eV/A) - the
    only well-defined oracle for a rank is the reference evaluated on THAT rank's atoms and list.  Cut a box into 2 and
    into 8 rank-like pieces (centre chunk + every atom its rows touch, renumbered, rows in the rank's own order) and pin
    every piece to the restatement of the reference (bit-identical to the reference binary, tests/test_oracle.py); the
    energies, which do not depend on the order, also add up to the single-domain energy."""
    import bench
    from oracle import restatement
    pot = read_potential(ni_pot_file, ["Ni"])
    x, box = L.fcc(5, 5, 5)
    cfg = L.build_config(L.perturb(x, 0.07, 99), box, 6.5, shuffle_rows=3)
    whole = make_ni(ni_pot_file)
    whole.compute(3, 0, cfg, ago=0)
    e_whole = whole.eng_vdwl
    whole.clear()
    for nparts in (2, 8):
        e_sum = 0.0
        for part in bench.split_config(cfg, nparts):
            rng = np.random.default_rng(part.nlocal)
            # a rank's list is in ITS bin order: reshuffle every row
            off = part.offsets
            neigh = part.neigh.copy()
            for i in range(part.nlocal):
                rng.shuffle(neigh[off[i]:off[i + 1]])
            part = L.Config(nlocal=part.nlocal, nghost=part.nghost, x=part.x, type=part.type, ghost_owner=part.ghost_owner, ilist=part.ilist,
                            numneigh=part.numneigh, neigh=neigh, box=part.box)
            pair = make_ni(ni_pot_file)
            f = pair.compute(3, 1, part, ago=0)
            ref = restatement.compute_ni(pot, part, nthreads=2)
            assert np.abs(f - ref["f"]).max() <= 1e-9
            assert np.abs(pair.eatom - ref["eatom"]).max() <= 1e-10
            e_sum += pair.eng_vdwl
            pair.clear()
        assert abs(e_sum - e_whole) <= 1e-9 * max(1.0, abs(e_whole))


def test_anna_against_live_oracle_and_network_outputs(anna_pot_file):
    from oracle import restatement
    x, box = L.bcc(5, 4, 3)
    cfg = L.build_config(L.perturb(x, 0.12, 2718), box, 5.055, shuffle_rows=5)
    pair = make_anna(anna_pot_file)
    f = pair.compute(3, 1, cfg, ago=0)
    ref = restatement.compute_anna(read_anna_potential(anna_pot_file, ["Fe"]), cfg, nthreads=4, dump_G=True)
    assert np.abs(f - ref["f"]).max() <= 1e-9
    assert np.abs(pair.eatom - ref["eatom"]).max() <= 1e-9
    G, lp = pair.descriptors(cfg)                       # for ANNA the second array carries (d2, q2) in columns 0, 1
    assert np.abs(G - ref["G"]).max() <= 1e-10
    assert np.abs(lp[:, :2] - ref["lparams"]).max() <= 1e-12
    pair.clear()


def test_variants_empty_and_isolated_atoms(ni_pot_file, anna_pot_file):
    from oracle import restatement
    one = L.build_config(np.array([[5.0, 5.0, 5.0]]), np.array([50.0, 50, 50]), 6.5, periodic=(False, False, False))
    pair = make_ni(ni_pot_file)
    f = pair.compute(3, 1, one, ago=0)
    ref = restatement.compute_ni(read_potential(ni_pot_file, ["Ni"]), one)
    assert np.all(f == 0.0) and abs(pair.eng_vdwl - ref["eng_vdwl"]) <= 1e-12
    pair.clear()
    # ANNA: an isolated atom has rho = 0 -> sqrt(0) in the energy, 0 * inf never formed because there is no neighbour
    pair = make_anna(anna_pot_file)
    f = pair.compute(3, 1, one, ago=0)
    ref = restatement.compute_anna(read_anna_potential(anna_pot_file, ["Fe"]), one)
    assert np.all(f == 0.0) and abs(pair.eng_vdwl - ref["eng_vdwl"]) <= 1e-9
    for pts in ([[5, 5, 5], [7.4, 5, 5]], [[5, 5, 5], [7.4, 5, 5], [6.0, 7.2, 5.3]]):
        c = L.build_config(np.array(pts, dtype=float), np.array([50.0, 50, 50]), 5.055, periodic=(False, False, False))
        f = pair.compute(3, 1, c, ago=0)
        ref = restatement.compute_anna(read_anna_potential(anna_pot_file, ["Fe"]), c)
        assert np.abs(f - ref["f"]).max() <= 1e-9 and np.abs(pair.eatom - ref["eatom"]).max() <= 1e-9
    pair.clear()


def test_explicit_variant_overrides_and_rejects(fe_pot_file):
    """A Fe file has no coefficient blocks: asking for the Ni copy on it is an error, not a silent fallback."""
    from meng_zhang_b200.pair import LammpsError
    pair = PairANNPGPU(ntypes=1, variant=capi.VARIANT_NI)
    pair.settings([])
    pair.coeff(["*", "*", fe_pot_file, "Fe"])
    with pytest.raises(LammpsError):
        pair.init_style()


@pytest.mark.parametrize("kind,prefix,name", [("plugin_annp_ni_b200", "annp_ni", "fcc3_perturbed"), ("plugin_annp_ni_b200", "annp_ni", "fcc3_two_types"),
                                              ("plugin_anna_adp_b200", "anna_adp", "bcc4_perturbed"), ("plugin_anna_adp_b200", "anna_adp", "cluster_ragged"),
                                              ("plugin_anna_adp_b200", "anna_adp", "bcc4_two_types")])
def test_lammps_pair_styles_through_the_shim_driver(kind, prefix, name, ni_pot_file, anna_pot_file):
    """The C++ classes LAMMPS would compile - PairANNPB200 built against the Ni copy of pair_annp.h, and
    PairANNAADPB200 (anna_adp/gpu) - run by the driver that runs the reference's own classes."""
    from oracle import run_ref
    if not run_ref.available(kind):
        pytest.skip(f"{kind} not built")
    cfg, elems, ref = util.load_case(name, prefix)
    pot = ni_pot_file if prefix == "annp_ni" else anna_pot_file
    for vflag, vkey in ((1 + 4, "virial_pair"), (2, "virial_fdotr")):
        out = run_ref.run_reference(kind, cfg, pot, elems, eflag=3, vflag=vflag)
        assert abs(out["eng_vdwl"] - ref["eng_vdwl"]) <= 1e-12 * abs(ref["eng_vdwl"]) + 1e-9
        assert np.abs(out["eatom"] - ref["eatom"]).max() <= 1e-9
        assert np.abs(out["f"] - ref["f"]).max() <= 1e-9
        assert np.abs(out["virial"] - ref[vkey]).max() <= 1e-8
        if vflag & 4:
            assert np.abs(out["vatom"] - ref["vatom"]).max() <= 1e-9


@pytest.mark.parametrize("kind,prefix,name", [("plugin_annp_ni_b200", "annp_ni", "fcc3_perturbed"), ("plugin_anna_adp_b200", "anna_adp", "bcc4_perturbed")])
def test_lammps_pair_styles_host_paths_of_the_other_two_styles(kind, prefix, name, ni_pot_file, anna_pot_file):
    """Repeated calls on one list (page-locked arrays, types not re-sent), the pair-hybrid situation (forces staged and
    added) and the device-neighbour mode for the Ni build of PairANNPB200 and for PairANNAADPB200.  The device-built list
    holds the same atoms per row in index order instead of the golden's shuffled order: ANNA-ADP reproduces the forces to
    rounding; the Ni copy's forces depend on the row order by construction (ni/src/pair_annp.cpp:734-735), so there only the
    energies - which do not - are compared."""
    from oracle import run_ref
    if not run_ref.available(kind):
        pytest.skip(f"{kind} not built")
    cfg, elems, ref = util.load_case(name, prefix)
    pot = ni_pot_file if prefix == "annp_ni" else anna_pot_file
    one = run_ref.run_reference(kind, cfg, pot, elems, eflag=3, vflag=1 + 4)
    many = run_ref.run_reference(kind, cfg, pot, elems, eflag=3, vflag=1 + 4, ncalls=3)
    assert np.array_equal(many["f"], one["f"]) and np.array_equal(many["eatom"], one["eatom"])
    hyb = run_ref.run_reference(kind, cfg, pot, elems, eflag=3, vflag=1 + 4, ncalls=2, env_extra={"ANNP_DRIVER_HYBRID": "1"})
    assert np.array_equal(hyb["f"], one["f"]) and hyb["eng_vdwl"] == one["eng_vdwl"]
    dev = run_ref.run_reference(kind, cfg, pot, elems, eflag=3, vflag=1 + 4, ncalls=2, env_extra={"ANNP_B200_NEIGH": "device"})
    assert np.abs(dev["eatom"] - ref["eatom"]).max() <= 1e-9
    if prefix != "annp_ni":
        assert np.abs(dev["f"] - ref["f"]).max() <= 1e-9
        assert np.abs(dev["virial"] - ref["virial_pair"]).max() <= 1e-8


def test_ni_device_md_uses_the_descriptor_cutoff_for_its_list(ni_pot_file):
    """The Ni file's `Cut` (6.5 A) only sizes LAMMPS' list; nothing beyond Rc = 7.3699 Bohr = 3.9 A contributes.  The
    device-resident driver builds its list with the tighter radius: same forces to rounding (the surviving row entries
    keep their order, which is what the Ni copy's forces depend on; only the lane partition of the ordered gather sums
    changes with the row length), a third of the list entries."""
    import torch
    from meng_zhang_b200.md import DomainMD
    x, box = L.fcc(5, 5, 5)
    xp = L.perturb(x, 0.06, 17)
    out = {}
    for key, cut in (("tight", None), ("file", 6.5)):
        pair = make_ni(ni_pot_file)
        md = DomainMD(pair, xp, box, mass=58.6934, list_cutoff=cut)
        md.reneighbor()
        md.compute(eflag=True)
        torch.cuda.synchronize()
        out[key] = (md.f[: md.nlocal].cpu().numpy().copy(), float(md.engvir[0]), pair.stats().max_neigh_list, md.cut)
        pair.clear()
    assert abs(out["tight"][3] - 7.3699319 / 1.889726) < 1e-12 and out["file"][3] == 6.5
    assert np.abs(out["tight"][0] - out["file"][0]).max() < 1e-13 and abs(out["tight"][1] - out["file"][1]) < 1e-11
    assert out["tight"][2] < 0.45 * out["file"][2]
    # and against the host path with LAMMPS' own list radius (different ghost numbering -> different row order ->
    # the reference's order dependence shows up in the forces, not in the energy)
    cfg = L.build_config(xp, box, 6.5)
    pair = make_ni(ni_pot_file)
    fh = cfg.fold(pair.compute(3, 0, cfg, ago=0))
    assert abs(pair.eng_vdwl - out["tight"][1]) < 1e-10
    assert np.abs(fh - out["tight"][0]).max() < 5e-2
    pair.clear()


@pytest.mark.parametrize("name", ["bcc4_perturbed", "bcc334_hot"])
def test_anna_plugin_accepts_newton_off_decks(name, anna_pot_file):
    """The reference's anna_adp/gpu decks say `newton off` (bcc_fe/README.md:39-40).  PairANNAADPB200 then returns its
    ghost forces itself (comm->reverse_comm(this)): the local rows equal the reference CPU style's folded forces."""
    from oracle import run_ref
    if not run_ref.available("plugin_anna_adp_b200"):
        pytest.skip("plugin_anna_adp_b200 not built")
    cfg, elems, ref = util.load_case(name, "anna_adp")
    out = run_ref.run_reference("plugin_anna_adp_b200", cfg, anna_pot_file, elems, eflag=3, vflag=1, newton=0)
    assert np.abs(out["f"][: cfg.nlocal] - cfg.fold(ref["f"])).max() <= 1e-9
    assert np.all(out["f"][cfg.nlocal:] == 0.0)                   # nothing is left on ghosts for LAMMPS to carry
    assert abs(out["eng_vdwl"] - ref["eng_vdwl"]) <= 1e-12 * abs(ref["eng_vdwl"]) + 1e-9
    assert np.abs(out["virial"] - ref["virial_pair"]).max() <= 1e-8

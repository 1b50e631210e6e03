"""The CPU oracle (oracle/annp_oracle.c) pinned against the reference.

Golden vectors under tests/golden/ were produced by the UNMODIFIED reference pair style
(annp-gpu-lammps/fe_v2/src/pair_annp.cpp compiled against oracle/shim, see make_golden.py).
The restatement keeps the reference's operation order, so the comparison is bit-exact.
"""
import os

import numpy as np
import pytest

import util
from meng_zhang_b200 import lattice as L
from meng_zhang_b200.pair import read_potential
from oracle import restatement, run_ref


@pytest.fixture(scope="module")
def pot(fe_pot_file):
    return read_potential(fe_pot_file, ["Fe"])


@pytest.mark.parametrize("name", [c for c in util.FE_CASES if c != "bcc10_perturbed"])
def test_restatement_matches_reference_golden_bit_exact(name, pot):
    cfg, elems, ref = util.load_case(name)
    out = restatement.compute(pot, cfg, ntypes=len(elems), vatom=True, nthreads=2)
    assert out["eng_vdwl"] == ref["eng_vdwl"]
    assert np.array_equal(out["eatom"], ref["eatom"])
    assert np.array_equal(out["f"], ref["f"])
    assert np.array_equal(out["virial"], ref["virial_pair"])
    assert np.array_equal(out["vatom"], ref["vatom"])


def test_restatement_2000_atoms_threads_do_not_change_bits(pot):
    """BASELINE config 1 geometry; the OpenMP evaluation tallies in ilist order like the serial code."""
    cfg, elems, ref = util.load_case("bcc10_perturbed")
    out = restatement.compute(pot, cfg, nthreads=os.cpu_count() or 4)
    assert out["eng_vdwl"] == ref["eng_vdwl"]
    assert np.array_equal(out["f"], ref["f"])
    assert np.array_equal(out["virial"], ref["virial_pair"])


def test_golden_known_answers_of_the_survey_probe():
    """Numbers the survey measured with the reference source: perfect bcc Fe."""
    cfg, _, ref = util.load_case("bcc4_perfect")
    assert abs(ref["eng_vdwl"] / cfg.nlocal - (-4479.8817655598)) < 1e-9
    assert np.abs(cfg.fold(ref["f"])).max() < 1e-12
    assert np.allclose(ref["virial_pair"][:3], -37.90672569, atol=1e-7)
    assert np.allclose(ref["virial_fdotr"], ref["virial_pair"], atol=1e-10)


@pytest.mark.skipif(not run_ref.available("annp_fe"), reason="oracle/_ref/ref_annp_fe not built")
def test_restatement_matches_live_reference_on_a_fresh_configuration(pot, fe_pot_file):
    x, box = L.bcc(3, 4, 5)
    cfg = L.build_config(L.perturb(x, 0.1, 31337), box, 6.5, shuffle_rows=1)
    ref = run_ref.run_reference("annp_fe", cfg, fe_pot_file, ["Fe"], eflag=3, vflag=1 + 4)
    out = restatement.compute(pot, cfg, vatom=True)
    assert out["eng_vdwl"] == ref["eng_vdwl"]
    assert np.array_equal(out["f"], ref["f"])
    assert np.array_equal(out["vatom"], ref["vatom"])


def test_forces_are_the_energy_gradient(pot):
    """Finite-difference self-consistency of the oracle (iii)."""
    x, box = L.bcc(3, 3, 3)
    x = L.perturb(x, 0.08, 5)
    big = np.array([60.0, 60.0, 60.0])
    cfg = L.build_config(x + 20.0, big, 6.5, periodic=(False, False, False))
    base = restatement.compute(pot, cfg)
    h = 1e-4
    for atom, k in [(0, 0), (13, 1), (40, 2)]:
        e = []
        for s in (+1, -1):
            xs = cfg.x.copy()
            xs[atom, k] += s * h
            c2 = L.Config(**{**cfg.__dict__, "x": xs})
            e.append(restatement.compute(pot, c2)["eng_vdwl"])
        fd = -(e[0] - e[1]) / (2 * h)
        assert abs(fd - base["f"][atom, k]) < 5e-6, (atom, k, fd, base["f"][atom, k])


def test_momentum_conservation_and_energy_sum(pot):
    cfg, _, ref = util.load_case("bcc4_perturbed")
    assert np.abs(ref["f"].sum(axis=0)).max() < 1e-11
    assert abs(ref["eatom"][: cfg.nlocal].sum() - ref["eng_vdwl"]) < 1e-7
    assert np.all(ref["eatom"][cfg.nlocal:] == 0.0)


# ---------------------------------------------------------------------------------------------- Ni copy, ANNA-ADP
@pytest.fixture(scope="module")
def ni_pot(ni_pot_file):
    return read_potential(ni_pot_file, ["Ni"])


@pytest.fixture(scope="module")
def anna_pot(anna_pot_file):
    from meng_zhang_b200.pair_anna import read_anna_potential
    return read_anna_potential(anna_pot_file, ["Fe"])


@pytest.mark.parametrize("name", util.NI_CASES)
def test_ni_restatement_matches_reference_golden_bit_exact(name, ni_pot):
    """oracle/annp_ni_oracle.c against ref_annp_ni (unmodified annp-gpu-lammps/ni/src/pair_annp.cpp)."""
    cfg, elems, ref = util.load_case(name, "annp_ni")
    out = restatement.compute_ni(ni_pot, cfg, ntypes=len(elems), vatom=True, nthreads=2)
    assert out["eng_vdwl"] == ref["eng_vdwl"]
    assert np.array_equal(out["eatom"], ref["eatom"])
    assert np.array_equal(out["f"], ref["f"])
    assert np.array_equal(out["virial"], ref["virial_pair"])
    assert np.array_equal(out["vatom"], ref["vatom"])


@pytest.mark.parametrize("name", util.ANNA_CASES)
def test_anna_restatement_matches_reference_golden_bit_exact(name, anna_pot):
    """oracle/anna_adp_oracle.c against ref_anna_adp (unmodified anna-gpu-lammps/bcc_fe/src/pair_anna_adp.cpp)."""
    cfg, elems, ref = util.load_case(name, "anna_adp")
    out = restatement.compute_anna(anna_pot, cfg, ntypes=len(elems), vatom=True, nthreads=2)
    assert out["eng_vdwl"] == ref["eng_vdwl"]
    assert np.array_equal(out["eatom"], ref["eatom"])
    assert np.array_equal(out["f"], ref["f"])
    assert np.array_equal(out["virial"], ref["virial_pair"])
    assert np.array_equal(out["vatom"], ref["vatom"])


def test_ni_anna_survey_known_answers():
    """SURVEY.md 8c probe values from the reference source: ANNA perfect bcc 4^3 E/atom."""
    cfg, _, ref = util.load_case("bcc4_perfect", "anna_adp")
    assert abs(ref["eng_vdwl"] / cfg.nlocal - (-4479.6483797674)) < 1e-9
    assert np.abs(cfg.fold(ref["f"])).max() < 1e-12


@pytest.mark.skipif(not run_ref.available("annp_ni"), reason="oracle/_ref/ref_annp_ni not built")
def test_ni_restatement_matches_live_reference(ni_pot, ni_pot_file):
    x, box = L.fcc(3, 2, 4)
    cfg = L.build_config(L.perturb(x, 0.1, 31337), box, 6.5, shuffle_rows=1)
    ref = run_ref.run_reference("annp_ni", cfg, ni_pot_file, ["Ni"], eflag=3, vflag=1 + 4)
    out = restatement.compute_ni(ni_pot, cfg, vatom=True)
    assert out["eng_vdwl"] == ref["eng_vdwl"]
    assert np.array_equal(out["f"], ref["f"])
    assert np.array_equal(out["vatom"], ref["vatom"])


@pytest.mark.skipif(not run_ref.available("anna_adp"), reason="oracle/_ref/ref_anna_adp not built")
def test_anna_restatement_matches_live_reference(anna_pot, anna_pot_file):
    x, box = L.bcc(3, 4, 5)
    cfg = L.build_config(L.perturb(x, 0.1, 31337), box, 5.055, shuffle_rows=1)
    ref = run_ref.run_reference("anna_adp", cfg, anna_pot_file, ["Fe"], eflag=3, vflag=1 + 4)
    out = restatement.compute_anna(anna_pot, cfg, vatom=True)
    assert out["eng_vdwl"] == ref["eng_vdwl"]
    assert np.array_equal(out["f"], ref["f"])
    assert np.array_equal(out["vatom"], ref["vatom"])


def test_ni_forces_are_not_the_exact_energy_gradient_but_close(ni_pot):
    """The Ni copy differentiates r_ij^2 + r_ik^2 + r_jk^2 with r_ik in place of r_jk (ni/src/pair_annp.cpp:734-735),
    so its forces are NOT the gradient of its energy; the oracle keeps that.  Document the size of the defect."""
    x, box = L.fcc(2, 2, 2)
    x = L.perturb(x, 0.08, 5)
    cfg = L.build_config(x + 20.0, np.array([60.0, 60.0, 60.0]), 6.5, periodic=(False, False, False))
    base = restatement.compute_ni(ni_pot, cfg)
    h = 1e-4
    worst = 0.0
    for atom, k in [(0, 0), (13, 1), (25, 2)]:
        e = []
        for s in (+1, -1):
            xs = cfg.x.copy()
            xs[atom, k] += s * h
            e.append(restatement.compute_ni(ni_pot, L.Config(**{**cfg.__dict__, "x": xs}))["eng_vdwl"])
        fd = -(e[0] - e[1]) / (2 * h) * 51.422515 / 1.889726     # network units per Angstrom -> CFFORCE per Bohr^-1
        worst = max(worst, abs(fd - base["f"][atom, k]))
    assert 1e-3 < worst < 5e-2      # ~1e-2 eV/A: a property of the reference's formula, not rounding


def test_ni_forces_depend_on_neighbour_row_order_energies_do_not(ni_pot):
    """Reference property kept by the oracle and the CUDA path: the first/second member of a pair is decided by the
    position in the neighbour row and enters the Ni copy's derivative asymmetrically (ni/src/pair_annp.cpp:734-735),
    so re-ordering the rows changes the forces at the 1e-2 eV/A level while G and E stay put (to rounding)."""
    x, box = L.fcc(3, 3, 3)
    xp = L.perturb(x, 0.05, 12345)
    a = restatement.compute_ni(ni_pot, L.build_config(xp, box, 6.5), dump_G=True)
    b = restatement.compute_ni(ni_pot, L.build_config(xp, box, 6.5, shuffle_rows=3), dump_G=True)
    assert np.abs(a["G"] - b["G"]).max() < 1e-12 and abs(a["eng_vdwl"] - b["eng_vdwl"]) < 1e-10
    d = np.abs(a["f"] - b["f"]).max()
    assert 1e-4 < d < 5e-2


@pytest.mark.parametrize("name", util.GENERAL_CASES)
def test_restatement_matches_reference_on_general_potentials(name, tmp_path, built):
    """Other descriptor shapes, activations 1/2/3, a non-linear output layer, five layers, two element blocks: golden
    answers of the unmodified reference on synthetic potentials (make_golden.py general).  The file is re-created with the
    writer and parsed by the library's reader, which files every weight block under element 0 as the reference does
    (fe_v2/src/pair_annp.cpp:455: the element index is reset on every line), so the two-element case runs with the LAST
    block as element 0 and an all-zero network for element 1 on both sides."""
    from meng_zhang_b200.pair import write_potential
    cfg, elems, ref, pot = util.load_general_case(name)
    pf = str(tmp_path / f"{name}.ann")
    write_potential(pf, pot)
    parsed = read_potential(pf, elems)
    assert (parsed.npsf, parsed.ntsf, parsed.nnod, parsed.ntl, parsed.flagact) == (pot.npsf, pot.ntsf, pot.nnod, pot.ntl, pot.flagact)
    if len(elems) > 1:
        assert np.array_equal(parsed.weight_all[0], pot.weight_all[-1]) and not parsed.weight_all[1].any()
    out = restatement.compute(parsed, cfg, ntypes=len(elems), type_map=[0] + list(range(len(elems))), vatom=True, nthreads=2)
    assert out["eng_vdwl"] == ref["eng_vdwl"]
    assert np.array_equal(out["eatom"], ref["eatom"])
    assert np.array_equal(out["f"], ref["f"])
    assert np.array_equal(out["virial"], ref["virial_pair"])
    assert np.array_equal(out["vatom"], ref["vatom"])

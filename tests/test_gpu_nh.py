"""Device-resident Nose-Hoover integrators (csrc/annp_nh.cu: fix nvt / fix npt of the reference decks).  -m gpu.

LAMMPS is not in this image, so the integrator is pinned two ways:
  * step by step against tests/nh_host.py, a numpy twin written line by line after FixNH's published algorithm, fed with
    the forces and virial the GPU produced (positions / velocities / box / chain variables to 1e-11)
  * physically: the extended energy PE + KE + E_chain is conserved, the temperature settles at the target, and under
    `npt y 0 0 1` (in.st_test:30-37) the coupled stress relaxes to the target while the uncoupled box edges stay put.
"""
import numpy as np
import pytest
import torch

import nh_host
import util
from meng_zhang_b200 import lattice as L
from meng_zhang_b200.md import DomainMD
from meng_zhang_b200.pair import PairANNPGPU

pytestmark = pytest.mark.gpu
MASS, DT = 55.845, 0.001


def make_md(fe_pot_file, cells, seed=7):
    pair = PairANNPGPU(ntypes=1)
    pair.settings([])
    pair.coeff(["*", "*", fe_pot_file, "Fe"])
    pair.init_style()
    x, box = L.bcc(cells, cells, cells)
    md = DomainMD(pair, x, box, mass=MASS, dt=DT)
    md.set_velocities(300.0, seed)
    md.reneighbor()
    return pair, md


@pytest.mark.parametrize("p_flag", [(0, 0, 0), (0, 1, 0), (1, 1, 1)], ids=["nvt", "npt_y", "npt_xyz"])
def test_step_by_step_against_the_host_twin(p_flag, fe_pot_file):
    pair, md = make_md(fe_pot_file, 4)
    kw = dict(p_flag=p_flag, p_start=(0.0, 0.0, 0.0), p_stop=(0.0, 0.0, 0.0), p_damp=(1.0, 1.0, 1.0))
    md.fix_nh(300.0, 300.0, 0.1, **kw)
    n = md.nlocal
    x_h, v_h = md.x[:n].cpu().numpy().copy(), md.v.cpu().numpy().copy()
    f_prev = md.f[:n].cpu().numpy().copy()
    twin = nh_host.HostNH(n, md.box, DT, MASS, 300.0, 300.0, 0.1, **kw)
    twin.setup(v_h, md.engvir[1:7].cpu().numpy())
    st = md.nh_state()
    assert abs(st.t_current - twin.t_current) < 1e-10
    for step in range(30):
        md.step_nh(eflag=True)
        x_h, v_h = twin.initial(x_h, v_h, f_prev)
        f_dev = md.f[:n].cpu().numpy().copy()
        v_h = twin.final(v_h, f_dev, md.engvir[1:7].cpu().numpy())
        f_prev = f_dev
        st = md.nh_state()
        assert np.abs(md.x[:n].cpu().numpy() - x_h).max() < 1e-11, step
        assert np.abs(md.v.cpu().numpy() - v_h).max() < 1e-11, step
        assert np.abs(np.array(st.boxhi[:]) - twin.boxhi).max() < 1e-11 and np.abs(np.array(st.boxlo[:]) - twin.boxlo).max() < 1e-11
        assert abs(st.t_current - twin.t_current) < 1e-9
        assert np.abs(np.array(st.eta_dot[:3]) - twin.eta_dot[:3]).max() < 1e-12
        assert abs(st.extended_energy - twin.extended_energy()) < 1e-10
        if any(p_flag):
            assert np.abs(np.array(st.omega_dot[:]) - twin.omega_dot).max() < 1e-13
            assert np.abs(np.array(st.p_current[:]) - twin.p_current).max() < 1e-5        # bar
    # ghosts follow the dilating box: periodic images stay exactly one (current) box edge away
    if any(p_flag):
        box_now = np.array(st.boxhi[:]) - np.array(st.boxlo[:])
        sh = md.send_shift.cpu().numpy()
        k = np.round(sh / box_now)
        assert np.abs(sh - k * box_now).max() < 1e-9 and np.abs(k).max() == 1
    pair.clear()


def test_nvt_controls_temperature_and_conserves_extended_energy(fe_pot_file):
    pair, md = make_md(fe_pot_file, 10, seed=4928459)           # BASELINE config 1 geometry, config 2's thermostat
    md.fix_nh(300.0, 300.0, 0.1)
    out = md.run_nh(3000, thermo_every=20)
    n = md.nlocal
    cons = np.array([(pe + ke + ext) / n for _, pe, ke, ext, *_ in out])
    temp = np.array([o[4] for o in out])
    assert np.abs(cons - cons[0]).max() < 2e-5                  # eV/atom over 3 ps (NVE drift of the same system: 1.6e-5)
    late = temp[len(temp) // 2:]
    assert abs(late.mean() - 300.0) < 10.0                      # 2 000 atoms: sigma_T ~ 300 sqrt(2/6000) = 5.5 K
    assert 2.0 < late.std() < 25.0                              # canonical sigma 5.5 K plus the chain's ringing at t_damp = 0.1 ps
    pair.clear()


def test_npt_y_relaxes_the_coupled_stress_only(fe_pot_file):
    pair, md = make_md(fe_pot_file, 8, seed=11)
    box0 = md.box.copy()
    md.fix_nh(300.0, 300.0, 0.1, p_flag=(0, 1, 0), p_start=(0, 0, 0), p_stop=(0, 0, 0), p_damp=(1.0, 1.0, 1.0))
    p0 = np.array(md.nh_state().p_current[:])
    out = md.run_nh(4000, thermo_every=20)
    n = md.nlocal
    cons = np.array([(pe + ke + ext) / n for _, pe, ke, ext, *_ in out])
    assert np.abs(cons - cons[0]).max() < 5e-5
    boxes = np.array([o[6] for o in out])
    pyy = np.array([o[5][1] for o in out])
    assert np.all(boxes[:, 0] == box0[0]) and np.all(boxes[:, 2] == box0[2])     # uncoupled edges never move
    assert np.ptp(boxes[:, 1]) > 1e-3                                            # the coupled one breathes
    # at a = 2.8553 A this potential is under ~ -4e4 bar of tension (the reference's own log: -40 423 bar at step 0,
    # log_relaxing_new.lammps:120); the barostat brings <p_yy> to the target by contracting the coupled edge
    assert p0[1] < -2.0e4
    late = pyy[len(pyy) // 2:]
    assert abs(late.mean()) < 0.25 * abs(p0[1]) + 1500.0
    assert boxes[-1, 1] < box0[1]
    pair.clear()


def test_replay_of_the_references_published_run_matches_its_lammps_log():
    """End to end against the only run the reference publishes (performance test.zip): fe_st.dat, boundary m p m, one cg
    minimiser iteration, `velocity all create 300 4928459`, `fix npt temp 300 300 0.1 y 0 0 1`, thermo 1.  LAMMPS' own
    parts (velocity generator, shrink-wrapped box, FixNH, neighbour trigger) are restated here; the log's thermo columns
    are reproduced to the printed precision over the first 120 steps, including the first re-neighbouring (Lx, Lz jump)."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))
    import replay_published_deck as R
    ours, ref, mini, log, rebuilds = R.replay(steps=120)
    # minimiser summary (log_relaxing_new.lammps:117-120); the reference's GPU build is mixed precision (2e-5 eV/A)
    assert mini["iterations"] == 1 and mini["evaluations"] == 2 and mini["stopping_criterion"] == "energy tolerance"
    assert abs(mini["fnorm_final"] - float(log["min_fnorm_initial_final_new"][1])) < 2e-4
    assert abs(mini["fmax_final"] - float(log["min_fmax_initial_final_new"][1])) < 5e-5
    assert abs(mini["alpha_final"] - float(log["min_alpha_maxmove_new"][0])) < 5e-6
    assert abs(mini["max_atom_move"] - float(log["min_alpha_maxmove_new"][1])) < 5e-6
    # energies: the reference's GPU build carries ~5e-9 relative error (3.4 eV at step 0, 7.1 eV after the step)
    for ours_e, log_e in zip((mini["energy_initial"], mini["energy_final"]), log["min_energy_initial_final_new"]):
        assert abs(ours_e / float(log_e) - 1.0) < 2e-8
    col = {c: i for i, c in enumerate(str(c) for c in log["columns"])}
    T, Tl = ours[:, col["Temp"]], ref[:, col["Temp"]]
    assert np.abs(T[:21] / Tl[:21] - 1.0).max() < 2e-6 and np.abs(T / Tl - 1.0).max() < 1e-5
    assert np.abs(ours[:, col["KinEng"]] / ref[:, col["KinEng"]] - 1.0).max() < 1e-5
    for c in ("Lx", "Ly", "Lz"):
        assert np.abs(ours[:, col[c]] - ref[:, col[c]]).max() < 2e-5, c
    assert ref[-1, col["Lx"]] != ref[0, col["Lx"]] and rebuilds >= 1          # the window contains a shrink-wrap update
    assert np.abs(ours[:, col["Volume"]] / ref[:, col["Volume"]] - 1.0).max() < 1e-6
    for c in ("Press", "Pxx", "Pyy", "Pzz"):
        assert np.abs(ours[:, col[c]] - ref[:, col[c]]).max() < 5.0, c         # bar; their virial is mixed precision


def test_cg_minimiser_relaxes_a_perturbed_crystal(fe_pot_file):
    """DomainMD.minimize on a thermally displaced cell: monotone in energy, stops on the force tolerance near the
    perfect-lattice energy of the same box."""
    pair, md = make_md(fe_pot_file, 4)                       # the perfect lattice of the same box (own handle)
    md.compute(eflag=True)
    e_perfect = float(md.engvir[0])
    pair2, _ = make_md(fe_pot_file, 4)
    x, box = L.bcc(4, 4, 4)
    md_p = DomainMD(pair2, L.perturb(x, 0.08, 3), box, mass=MASS, dt=DT)
    md_p.reneighbor()
    st = md_p.minimize(0.0, 1.0e-6, 400, 2000)
    assert st["stopping_criterion"] == "force tolerance" and st["fnorm_final"] < 1e-6 < st["fnorm_initial"]
    assert st["energy_final"] < st["energy_initial"]
    assert abs(st["energy_final"] - e_perfect) < 1e-7
    pair.clear()
    pair2.clear()


def test_the_references_input_deck_runs_verbatim(tmp_path, fe_pot_file):
    """meng_zhang_b200.deck takes the reference's own LAMMPS deck (in.st_test) as it is - data file, potential file and
    deck in one directory, like a LAMMPS run - and its thermo output follows the reference's LAMMPS log."""
    import io
    import shutil
    from meng_zhang_b200.deck import Deck
    from meng_zhang_b200.structures import write_lammps_data
    z = np.load(f"{util.GOLDEN}/fe_st.npz")
    log = np.load(f"{util.GOLDEN}/fe_st_log.npz")
    write_lammps_data(str(tmp_path / "fe_st.dat"), z["x"], z["box"][:, 1], np.ones(len(z["x"]), dtype=np.int32), ntypes=1)
    shutil.copy(fe_pot_file, tmp_path / "fe_annp_potential_2.ann")
    deck = open(f"{util.GOLDEN}/in.st_test", newline="").read()
    assert "run\t\t\t1000" in deck
    deck = deck.replace("run\t\t\t1000", "run\t\t\t60")           # the only edit: 60 of the 1000 steps
    with open(tmp_path / "in.st_test", "w", newline="") as fp:
        fp.write(deck)
    buf = io.StringIO()
    d = Deck(out=buf)
    d.run_file(str(tmp_path / "in.st_test"))
    out = buf.getvalue()
    assert "Stopping criterion = energy tolerance" in out and "Iterations, force evaluations = 1 2" in out
    assert "Step" in out and "Loop time of" in out
    cols = ["step", "temp", "pe", "ke", "lx", "ly", "lz", "press", "vol", "pxx", "pyy", "pzz"]      # the deck's thermo_style custom
    assert d.thermo_cols == cols and len(d.rows) == 61
    ref = log["thermo_new"][:61]
    ours = np.array([[r[c] for c in cols] for r in d.rows])
    assert np.array_equal(ours[:, 0], ref[:, 0])
    assert np.abs(ours[:, 1] / ref[:, 1] - 1.0).max() < 5e-6                  # Temp
    assert np.abs(ours[:, 5] - ref[:, 5]).max() < 2e-5                        # Ly
    assert np.abs(ours[:, 7] - ref[:, 7]).max() < 5.0                         # Press (bar)
    d.pair.clear()

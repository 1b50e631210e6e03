"""Host-side logic: potential reader/writer, pair-style argument handling, synthetic configurations."""
import numpy as np
import pytest

import util
from meng_zhang_b200 import lattice as L
from meng_zhang_b200.pair import LammpsError, PairANNPGPU, read_potential, write_potential


def test_reader_reproduces_the_reference_file_contract(fe_pot_file):
    pot = read_potential(fe_pot_file, ["Fe"])
    g = util.load_potential_json()
    assert (pot.ntl, pot.nhl, pot.nnod, pot.nsf, pot.npsf, pot.ntsf) == (4, 2, 10, 28, 9, 19)
    assert pot.cut == 6.5 and pot.flagsym == 0
    assert pot.flagact == [4, 4, 0]          # "tanh" is matched by "ta" -> activation 4 (pair_annp.cpp:423)
    assert pot.e_scale == 0.80684104305538540 and pot.e_shift == -1019.0781365280557 and pot.e_atom == -3460.0
    assert pot.elements == ["Fe"] and pot.mass == [55.847]
    assert np.array_equal(pot.weight_all, g.weight_all) and np.array_equal(pot.bias_all, g.bias_all)
    assert np.array_equal(pot.sfnor_cov, g.sfnor_cov) and np.array_equal(pot.sfnor_avg, g.sfnor_avg)
    # rows are padded to nsf columns like c_3d_matrix(n_lay, n_nod, n_sf) (pair_annp.cpp:445)
    assert pot.weight_all.shape == (1, 3, 10, 28) and np.all(pot.weight_all[0, 1, :, 10:] == 0)
    assert np.all(pot.weight_all[0, 2, 1:, :] == 0)


def test_reader_tokenisation_quirks(tmp_path):
    """Numbers are taken at column 0 and after TAB+digit/'-' only; CRLF is tolerated."""
    pot = util.load_potential_json()
    p = tmp_path / "q.ann"
    write_potential(str(p), pot)
    txt = open(p, newline="").read()
    assert "\r\n" in txt
    lines = txt.split("\r\n")
    # a value written as ".5" after a tab is skipped by the reference scanner, shifting the row left
    row = lines[12].split("\t")
    row[1] = ".5"
    lines[12] = "\t".join(row)
    open(p, "w", newline="").write("\r\n".join(lines))
    q = read_potential(str(p), ["Fe"])
    assert q.sfnor_cov[0] == pot.sfnor_cov[0]
    assert q.sfnor_cov[1] == pot.sfnor_cov[2]          # shifted
    # LF-only files parse identically
    open(p, "w", newline="").write("\n".join(txt.split("\r\n")))
    r = read_potential(str(p), ["Fe"])
    assert np.array_equal(r.weight_all, pot.weight_all) and r.flagact == pot.flagact


def test_reader_errors(tmp_path):
    with pytest.raises(LammpsError, match="Cannot open neural network potential file"):
        read_potential(str(tmp_path / "missing.ann"), ["Fe"])
    p = tmp_path / "short.ann"
    p.write_text("#a\r\n#b\r\n")
    with pytest.raises(LammpsError):
        read_potential(str(p), ["Fe"])


def test_sf_scale_matches_reference_formula():
    pot = util.load_potential_json()
    s = pot.sf_scale()
    assert np.allclose(s, 1.0 / np.sqrt(pot.sfnor_cov - pot.sfnor_avg ** 2), rtol=0, atol=0)
    pot.sfnor_cov[3] = pot.sfnor_avg[3] ** 2     # degenerate -> 0 with a warning in the reference
    assert pot.sf_scale()[3] == 0.0


def test_pair_style_argument_errors(fe_pot_file):
    pair = PairANNPGPU(ntypes=1)
    with pytest.raises(LammpsError, match="Illegal pair_style command"):
        pair.settings(["1.0"])
    pair.settings([])
    with pytest.raises(LammpsError, match="Incorrect args for pair coefficients"):
        pair.coeff(["*", "*", fe_pot_file])                 # missing element
    with pytest.raises(LammpsError, match="Incorrect args for pair coefficients"):
        pair.coeff(["1", "*", fe_pot_file, "Fe"])
    pair2 = PairANNPGPU(ntypes=2)
    with pytest.raises(LammpsError, match="Incorrect args for pair coefficients"):
        pair2.coeff(["*", "*", fe_pot_file, "Fe", "Ni"])      # two elements, file has one
    pair2.coeff(["*", "*", fe_pot_file, "Fe", "Fe"])
    assert list(pair2.map) == [-1, 0, 0] and pair2.cutmax == 6.5
    assert pair2.init_one(1, 2) == 6.5
    off = PairANNPGPU(ntypes=1, newton_pair=0)
    off.coeff(["*", "*", fe_pot_file, "Fe"])
    with pytest.raises(LammpsError, match="requires newton pair on"):
        off.init_style()


def test_flat_weights_layout():
    pot = util.load_potential_json()
    w, b = pot.flat_weights()
    assert w.size == 390 and b.size == 21
    assert w[28 * 3 + 5] == pot.weight_all[0, 0, 3, 5]        # k + j*ncol (pair_annp_gpu.cpp:203)
    assert w[280 + 10 * 2 + 7] == pot.weight_all[0, 1, 2, 7]
    assert w[380 + 4] == pot.weight_all[0, 2, 0, 4]
    assert b[20] == pot.bias_all[0, 2, 0]


def test_bcc_neighbour_shells_and_ghosts():
    x, box = L.bcc(4, 4, 4)
    cfg = L.build_config(x, box, 6.5)
    assert cfg.nlocal == 128
    assert np.all(cfg.numneigh == 228)                       # 8.5 A list
    off = cfg.offsets
    d = np.linalg.norm(cfg.x[cfg.neigh[off[0]:off[1]]] - cfg.x[0], axis=1)
    assert (d <= 6.5).sum() == 112                           # 112 in-cutoff neighbours
    assert np.allclose(cfg.x[cfg.nlocal:], cfg.x[cfg.ghost_owner] + cfg.ghost_shift)
    f = np.random.default_rng(0).normal(size=(cfg.nall, 3))
    folded = cfg.fold(f)
    assert np.allclose(folded.sum(axis=0), f.sum(axis=0))


def test_free_boundaries_have_no_ghosts():
    x, _ = L.bcc(2, 2, 2)
    cfg = L.build_config(x + 10, np.array([50.0, 50, 50]), 6.5, periodic=(False, False, False))
    assert cfg.nghost == 0 and cfg.numneigh.max() == 15


# ---------------------------------------------------------------------------------------------------------------
# the Nose-Hoover host twin (tests/nh_host.py) on an analytic force field: the transcription of FixNH used to pin
# csrc/annp_nh.cu must itself be a sound integrator (conserved extended energy, canonical temperature, relaxing stress)
def _lj_forces(x, boxlen, eps=0.0104, sig=3.4, rc=7.5):
    """Shifted-force Lennard-Jones (argon-like, metal units) with minimum image; returns f, pe, virial diag (W_aa)."""
    d = x[:, None, :] - x[None, :, :]
    d -= np.round(d / boxlen) * boxlen
    r2 = (d * d).sum(axis=2)
    np.fill_diagonal(r2, np.inf)
    r = np.sqrt(r2)
    m = r < rc
    sr6 = np.where(m, (sig * sig / r2) ** 3, 0.0)
    fr = np.where(m, 24.0 * eps * (2.0 * sr6 * sr6 - sr6) / r2, 0.0)            # -(dU/dr)/r
    src6 = (sig / rc) ** 6
    frc = 24.0 * eps * (2.0 * src6 * src6 - src6) / rc                          # -(dU/dr) at rc
    fr = np.where(m, fr - frc / r, 0.0)
    u = np.where(m, 4.0 * eps * (sr6 * sr6 - sr6) - 4.0 * eps * (src6 * src6 - src6) + frc * (r - rc), 0.0)
    fvec = fr[:, :, None] * d
    f = fvec.sum(axis=1)
    vir = 0.5 * np.array([(d[:, :, a] * fvec[:, :, a]).sum() for a in range(3)] + [0.0, 0.0, 0.0])
    return f, 0.5 * float(u.sum()), vir


@pytest.mark.parametrize("p_flag", [(0, 0, 0), (0, 1, 0)], ids=["nvt", "npt_y"])
def test_nose_hoover_host_twin_conserves_the_extended_energy(p_flag):
    import nh_host
    mass, dt, T = 39.948, 0.002, 40.0
    x, box = L.fcc(3, 3, 3, a=5.31)                                               # 108 atoms, solid argon
    n = len(x)
    rng = np.random.default_rng(5)
    sigma = (nh_host.BOLTZ * T / (mass * nh_host.MVV2E)) ** 0.5
    v = rng.standard_normal((n, 3)) * sigma
    v -= v.mean(axis=0)
    nh = nh_host.HostNH(n, box, dt, mass, T, T, 0.2, p_flag=p_flag, p_start=(0, 0, 0), p_stop=(0, 0, 0), p_damp=(2, 2, 2))
    f, pe, vir = _lj_forces(x, nh.boxhi - nh.boxlo)
    nh.setup(v, vir)
    temps, cons, ly = [], [], []
    for step in range(4000):
        x, v = nh.initial(x, v, f)
        f, pe, vir = _lj_forces(x, nh.boxhi - nh.boxlo)
        v = nh.final(v, f, vir)
        cons.append(pe + 0.5 * nh.mvv[:3].sum() + nh.extended_energy())
        temps.append(nh.t_current)
        ly.append(nh.boxhi[1] - nh.boxlo[1])
    cons = np.array(cons)
    assert np.abs(cons - cons[0]).max() / n < 5e-6               # eV per atom over 8 ps (KE per atom is 5e-3 eV)
    late = np.array(temps[1500:])
    assert abs(late.mean() - T) < 4.0
    if any(p_flag):
        assert nh.boxhi[0] - nh.boxlo[0] == box[0] and nh.boxhi[2] - nh.boxlo[2] == box[2]    # uncoupled edges fixed
        assert np.ptp(ly) > 1e-3                                                                # coupled edge breathes
        assert abs(np.mean(nh.p_current[1])) < 3000.0                                           # bar, near the 0 target


@pytest.mark.parametrize("ntsf", [6, 19, 20, 24])
def test_angular_basis_conversion_matrices(ntsf):
    """The kernel evaluates the reference's T_n((cos theta + 1)/2) (fe_v2/src/pair_annp.cpp:596-611, 671) in two other
    bases of z = cos theta: monomials (backward Horner) and psi_{4b+i} = T_{4b}(z) z^i (forward accumulation).  Both
    conversion matrices are host arithmetic (annp_b200_basis_matrices): check the identities against the recurrence the
    reference uses, and the conditioning that makes the block basis as accurate as the recurrence."""
    import ctypes as C
    from meng_zhang_b200 import capi
    from numpy.polynomial import chebyshev as Ch
    lib = capi.lib()
    c2m = np.zeros((ntsf, ntsf))
    b2c = np.zeros((ntsf, ntsf))
    assert lib.annp_b200_basis_matrices(ntsf, c2m.ctypes.data_as(capi.c_double_p), b2c.ctypes.data_as(capi.c_double_p)) == 0
    z = np.linspace(-1.0, 1.0, 401).astype(np.longdouble)
    y = (z + 1) / 2

    def cheb_rec(x, n):          # the reference's recurrence (annp_Tx)
        t0, t1 = np.ones_like(x), x.copy()
        if n == 0:
            return t0
        for _ in range(n - 1):
            t0, t1 = t1, 2 * x * t1 - t0
        return t1

    T_ref = np.stack([cheb_rec(y, n) for n in range(ntsf)])                       # [n][sample]
    mono = np.stack([z ** k for k in range(ntsf)])                                # [k][sample]
    psi = np.stack([cheb_rec(z, 4 * (j // 4)) * z ** (j % 4) for j in range(ntsf)])
    assert np.abs(c2m.astype(np.longdouble).T @ mono - T_ref).max() < 1e-13
    assert np.abs(b2c.astype(np.longdouble).T @ psi - T_ref).max() < 1e-15
    # conditioning: |columns| of the block-basis conversion sum to a few tens, those of the monomial one to ~1e4
    assert np.abs(b2c).sum(axis=0).max() < 40.0
    if ntsf >= 19:
        assert np.abs(c2m).sum(axis=0).max() > 1e3
    # entries are dyadic rationals: the first rows are plain Chebyshev structure (T_0 = 1, T_1 = (z+1)/2)
    assert c2m[0, 0] == 1.0 and c2m[0, 1] == 0.5 and c2m[1, 1] == 0.5
    assert lib.annp_b200_basis_matrices(25, c2m.ctypes.data_as(capi.c_double_p), b2c.ctypes.data_as(capi.c_double_p)) == capi.EINVAL


@pytest.mark.parametrize("fixed_point", [True, False])
def test_kernel_arithmetic_restated_in_numpy_matches_the_oracle(fixed_point, fe_pot_file):
    """The CUDA kernel's reformulation (block-basis forward sums, reverse-mode MLP, four-moment backward pass with
    S = u.V, monomial Horner, fixed-point scatter), restated in numpy (tests/kernel_math_host.py), gives the reference's
    forces and energies: the algebra is checked on the CPU, the GPU tests then only have to show that the kernel
    implements it."""
    import kernel_math_host as K
    from meng_zhang_b200 import capi
    from oracle import restatement
    pot = read_potential(fe_pot_file, ["Fe"])
    c2m, b2c = np.zeros((pot.ntsf, pot.ntsf)), np.zeros((pot.ntsf, pot.ntsf))
    assert capi.lib().annp_b200_basis_matrices(pot.ntsf, c2m.ctypes.data_as(capi.c_double_p), b2c.ctypes.data_as(capi.c_double_p)) == 0
    x, box = L.bcc(3, 3, 3)
    cfg = L.build_config(L.perturb(x, 0.08, 31), box, 6.5)
    ref = restatement.compute(pot, cfg, nthreads=2)
    f, eatom = K.compute(pot, cfg, c2m, b2c, fixed_point=fixed_point)
    assert np.abs(eatom[: cfg.nlocal] - ref["eatom"][: cfg.nlocal]).max() <= 5e-12
    assert np.abs(f - ref["f"]).max() <= (3e-12 if fixed_point else 3e-13)       # observed 5.2e-13 / 5.4e-14, as on the GPU
    if fixed_point:       # every pair force moved by at most half a unit of 2^-43 eV/A: <= 112 * 5.7e-14 per atom
        f_plain, _ = K.compute(pot, cfg, c2m, b2c, fixed_point=False)
        assert 0.0 < np.abs(f - f_plain).max() <= 120 * 0.5 * 2.0 ** -43


def test_deck_formulas_are_parsed_not_evaluated_as_python():
    """`variable equal` / $(...) formulas: arithmetic, variables, dt, sqrt/exp/ln - and nothing else (no attribute access,
    no other calls, bounded exponents)."""
    from meng_zhang_b200.deck import Deck, DeckError
    d = Deck.__new__(Deck)
    d.vars, d.dt = {"a": "2.5", "n": "4"}, 0.001
    assert d.evaluate("1+2*3") == 7.0
    assert d.evaluate("v_a^2/dt") == 6250.0
    assert d.evaluate("-sqrt(16)+exp(0)*v_n") == 0.0
    assert abs(d.evaluate("ln(PI)") - 1.1447298858494002) < 1e-15
    for bad in ("().__class__", "9^9^9^9", "__import__('os')", "a.b", "sqrt(1,2)", "[1][0]", "1 if 1 else 2", "1/0", "v_missing"):
        with pytest.raises(DeckError):
            d.evaluate(bad)

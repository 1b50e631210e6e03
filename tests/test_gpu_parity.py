"""Parity of the CUDA path (through the C ABI) against the reference's numbers.  Run with -m gpu.

Tolerances (north_star: per-atom energy <= 1e-6 relative, forces <= 1e-5 eV/A).  The kernel re-associates
sums, so agreement is to rounding, far inside that bar; the thresholds below are what is asserted:
    per-atom energy  |dE_i| <= 1e-9 eV   (2e-13 relative to |E_i| ~ 4480 eV, ~1e-9 of the cohesive part)
    forces           |dF|   <= 1e-9 eV/A
    virial           |dV|   <= 1e-8 eV (pair tally) ; per-atom virial <= 1e-9
"""
import ctypes as C

import numpy as np
import pytest

import util
from meng_zhang_b200 import capi, lattice as L
from meng_zhang_b200.pair import PairANNPGPU, read_potential

pytestmark = pytest.mark.gpu

TOL_E, TOL_F, TOL_V = 1e-9, 1e-9, 1e-8


def make_pair(pot_file, elems=("Fe",)):
    pair = PairANNPGPU(ntypes=len(elems))
    pair.settings([])
    pair.coeff(["*", "*", pot_file] + list(elems))
    pair.init_style()
    return pair


@pytest.mark.parametrize("name", util.FE_CASES)
def test_golden_case(name, fe_pot_file):
    cfg, elems, ref = util.load_case(name)
    pair = make_pair(fe_pot_file, elems)
    f = pair.compute(3, 1 + 4, cfg, ago=0)
    assert np.abs(pair.eatom - ref["eatom"]).max() <= TOL_E
    assert abs(pair.eng_vdwl - ref["eng_vdwl"]) <= 1e-12 * abs(ref["eng_vdwl"]) + TOL_E
    assert np.abs(f - ref["f"]).max() <= TOL_F
    assert np.abs(pair.virial - ref["virial_pair"]).max() <= TOL_V
    assert np.abs(pair.virial - ref["virial_fdotr"]).max() <= TOL_V
    assert np.abs(pair.vatom - ref["vatom"]).max() <= TOL_F
    # bit-reproducible: deterministic reductions, no floating-point atomics
    f2 = pair.compute(3, 1 + 4, cfg, ago=1)
    assert np.array_equal(f, f2) and np.array_equal(pair.vatom, pair.vatom)
    # energy/virial are optional outputs: forces must not depend on the flags
    f3 = pair.compute(0, 0, cfg, ago=1)
    assert np.array_equal(f, f3)
    pair.clear()


@pytest.mark.parametrize("name", ["bcc4_perturbed", "cluster_ragged"])
def test_scatter_modes_agree(name, fe_pot_file):
    """Fixed-point integer atomics (default) and the ordered FP64 gather give the same forces to the fixed-point
    resolution (2^-43 eV/A per pair force), both match the reference, both are bit-reproducible."""
    cfg, elems, ref = util.load_case(name)
    pair = make_pair(fe_pot_file, elems)
    f_fixed = pair.compute(3, 1 + 4, cfg, ago=0)
    va_fixed = pair.vatom.copy()
    assert np.array_equal(f_fixed, pair.compute(1, 0, cfg, ago=1))
    pair.set_scatter(capi.SCATTER_GATHER)
    f_gather = pair.compute(3, 1 + 4, cfg, ago=1)
    va_gather = pair.vatom.copy()
    assert np.array_equal(f_gather, pair.compute(1, 0, cfg, ago=1))
    assert np.abs(f_fixed - f_gather).max() <= 2e-11
    assert np.abs(f_fixed - ref["f"]).max() <= TOL_F and np.abs(f_gather - ref["f"]).max() <= TOL_F
    assert np.array_equal(va_fixed, va_gather)
    pair.set_scatter(capi.SCATTER_FIXED)
    assert np.array_equal(f_fixed, pair.compute(1, 0, cfg, ago=1))
    # the sum of all forces (ghost forces folded) vanishes to the rounding of the centre sums
    assert np.abs(cfg.fold(f_fixed).sum(axis=0)).max() <= 1e-10
    pair.clear()


def test_fixed_point_scatter_refuses_non_finite_forces(fe_pot_file):
    """A NaN position gives NaN pair forces: the fixed-point accumulation must fail the call (ANNP_B200_EOVERFLOW), not
    add garbage; the handle stays usable afterwards."""
    cfg, elems, ref = util.load_case("bcc4_perturbed")
    pair = make_pair(fe_pot_file, elems)
    x_bad = cfg.x.copy()
    x_bad[3, 1] = np.nan
    with pytest.raises(capi.AnnpError) as ei:
        pair.compute(1, 0, cfg, ago=0, x=x_bad)
    assert ei.value.code == capi.EOVERFLOW and "fixed-point" in str(ei.value)
    f = pair.compute(1, 0, cfg, ago=1)
    assert np.abs(f - ref["f"]).max() <= TOL_F
    pair.clear()


def test_against_live_oracle_random_configuration(fe_pot_file):
    from oracle import restatement
    x, box = L.bcc(5, 4, 3)
    cfg = L.build_config(L.perturb(x, 0.12, 2718), box, 6.5, shuffle_rows=5)
    pair = make_pair(fe_pot_file)
    f = pair.compute(3, 1, cfg, ago=0)
    ref = restatement.compute(read_potential(fe_pot_file, ["Fe"]), cfg, nthreads=4, dump_G=True)
    assert np.abs(f - ref["f"]).max() <= TOL_F
    assert np.abs(pair.eatom - ref["eatom"]).max() <= TOL_E
    G, dE = pair.descriptors(cfg)
    assert np.abs(G - ref["G"]).max() <= 1e-10
    pair.clear()


def test_reference_run_log_golden_thermo(fe_pot_file):
    """The reference's own 2-GPU annp/gpu run (performance test.zip): 152 880-atom slab, boundary m p m.
    Its LAL precision is not logged; old and new kernels differ by 1.2e-10 relative in E, 1.7e-6 in |F|."""
    z = np.load(f"{util.GOLDEN}/fe_st.npz")
    x, box = z["x"], z["box"]
    lo, hi = box[:, 0], box[:, 1]
    cfg = build_large_config(x - lo, hi - lo, periodic=(False, True, False))
    pair = make_pair(fe_pot_file)
    f = pair.compute(1, 1, cfg, ago=0)
    st = pair.stats()
    assert abs(cfg.numneigh.sum() / cfg.nlocal - float(z["neighs_per_atom"])) < 1e-4   # 217.55887
    ff = cfg.fold(f)
    # (1) the reference's FP64 CPU algorithm on this input (restatement, bit-identical to the reference source)
    assert abs(pair.eng_vdwl - float(z["e_cpu_fp64"])) <= 1e-12 * abs(float(z["e_cpu_fp64"]))
    assert np.abs(ff[z["f_sample_idx"]] - z["f_sample_cpu_fp64"]).max() <= TOL_F
    assert abs(np.linalg.norm(ff) - float(z["fnorm_cpu_fp64"])) <= 1e-9
    assert np.abs(pair.virial - z["virial_cpu_fp64"]).max() <= 1e-6
    # (2) the thermo output logged by the reference's GPU run: reduced precision on their side
    #     (their log vs their own FP64 CPU algorithm: dE/E = 4.9e-9, d|F|max = 1.9e-5 eV/A, dP = 2.3 bar)
    e_ref = float(z["e_pair_new"])
    assert abs(pair.eng_vdwl - e_ref) <= 1e-8 * abs(e_ref)
    assert abs(np.linalg.norm(ff) - float(z["fnorm_new"])) <= 2e-4
    assert abs(np.abs(ff).max() - float(z["fmax_new"])) <= 5e-5
    # pressure of the minimiser's step 0 (T = 0): P = virial / (3V) * nktv2p(metal) = 1.6021765e6
    p = pair.virial[:3].sum() / (3.0 * float(z["volume"])) * 1.6021765e6
    assert abs(p - float(z["press_new"])) <= 5.0
    assert st.max_neigh_cut <= capi.MAX_NEIGH
    pair.clear()


def build_large_config(x, box, periodic, cutoff=6.5, skin=2.0):
    """Config for ~1e5 atoms without Python-level per-atom loops (cKDTree sparse distance matrix)."""
    from scipy.spatial import cKDTree
    x = L.wrap(np.asarray(x, dtype=np.float64), box, periodic)
    cut = cutoff + skin
    gx, gowner, gshift = L.make_ghosts(x, box, cut, periodic)
    xa = np.ascontiguousarray(np.concatenate([x, gx]))
    nlocal = len(x)
    pairs = cKDTree(xa).query_pairs(cut, output_type="ndarray")
    i = np.concatenate([pairs[:, 0], pairs[:, 1]])
    j = np.concatenate([pairs[:, 1], pairs[:, 0]])
    keep = i < nlocal
    i, j = i[keep], j[keep]
    order = np.lexsort((j, i))
    i, j = i[order], j[order]
    numneigh = np.bincount(i, minlength=nlocal).astype(np.int32)
    return L.Config(nlocal=nlocal, nghost=len(gx), x=xa, type=np.ones(len(xa), dtype=np.int32), ghost_owner=gowner,
                    ilist=np.arange(nlocal, dtype=np.int32), numneigh=numneigh, neigh=j.astype(np.int32), box=box, ghost_shift=gshift)


def test_device_neighbour_build_equals_host_list(fe_pot_file):
    import torch
    x, box = L.bcc(6, 5, 4)
    cfg = L.build_config(L.perturb(x, 0.1, 77), box, 6.5)
    pair = make_pair(fe_pot_file)
    f_host = pair.compute(3, 1, cfg, ago=0)
    e_host = pair.eng_vdwl
    lib = capi.lib()
    dx = torch.as_tensor(cfg.x, device="cuda")
    dt = torch.as_tensor(cfg.type, device="cuda")
    lo = (cfg.x.min(axis=0) - 1e-3).copy()
    hi = (cfg.x.max(axis=0) + 1e-3).copy()
    rc = lib.annp_b200_neigh_build(pair.handle, cfg.nlocal, cfg.nall, C.c_void_p(dx.data_ptr()), lo.ctypes.data_as(capi.c_double_p),
                                   hi.ctypes.data_as(capi.c_double_p), 8.5, None)
    assert rc == 0, lib.annp_b200_last_error(pair.handle)
    st = pair.stats()
    assert st.max_neigh_list == cfg.numneigh.max() and st.inum == cfg.nlocal
    df = torch.zeros((cfg.nall, 3), dtype=torch.float64, device="cuda")
    ev = torch.zeros(8, dtype=torch.float64, device="cuda")
    rc = lib.annp_b200_compute_device(pair.handle, cfg.nlocal, cfg.nghost, C.c_void_p(dx.data_ptr()), C.c_void_p(dt.data_ptr()), 1, 1,
                                      C.c_void_p(df.data_ptr()), None, C.c_void_p(ev.data_ptr()), None, None)
    assert rc == 0, lib.annp_b200_last_error(pair.handle)
    torch.cuda.synchronize()
    # host rows are sorted by index exactly like the device build -> identical summation order -> identical bits
    assert np.array_equal(df.cpu().numpy(), f_host)
    assert float(ev[0]) == e_host
    pair.clear()


def test_invariances_at_scale(fe_pot_file):
    """Size-independent properties on 16 000 atoms (too large for the CPU oracle in a unit test):
    momentum conservation, translation invariance, permutation invariance of the centre order."""
    x, box = L.bcc(20, 20, 20)
    xp = L.perturb(x, 0.06, 99)
    cfg = build_large_config(xp, box, (True, True, True))
    pair = make_pair(fe_pot_file)
    f = pair.compute(3, 1, cfg, ago=0)
    ff = cfg.fold(f)
    assert np.abs(f.sum(axis=0)).max() < 1e-9
    assert abs(pair.eatom[: cfg.nlocal].sum() - pair.eng_vdwl) < 1e-6
    e0 = pair.eng_vdwl
    shifted = L.Config(**{**cfg.__dict__, "x": cfg.x + np.array([0.123, -0.456, 0.789])})
    f_s = pair.compute(3, 1, shifted, ago=1)
    assert np.abs(f_s - f).max() < 1e-10 and abs(pair.eng_vdwl - e0) < 1e-9 * abs(e0)
    perm = np.random.default_rng(3).permutation(cfg.nlocal).astype(np.int32)
    off = cfg.offsets
    rows = [cfg.neigh[off[i]:off[i + 1]] for i in perm]
    pc = L.Config(**{**cfg.__dict__, "ilist": perm, "numneigh": cfg.numneigh[perm], "neigh": np.concatenate(rows)})
    f_p = pair.compute(3, 1, pc, ago=0)
    assert np.abs(cfg.fold(f_p) - ff).max() < 1e-10
    assert np.abs(pair.eatom - pair.eatom).max() == 0.0
    pair.clear()


def test_tile_overflow_regrows_capacity(fe_pot_file):
    """Atoms moving inside the skin between list builds can raise the in-cutoff count: the library must
    grow its shared-memory tile and redo the step, not truncate."""
    x, box = L.bcc(4, 4, 4)
    cfg = L.build_config(x, box, 6.5)
    pair = make_pair(fe_pot_file)
    f0 = pair.compute(3, 0, cfg, ago=0)
    assert pair.stats().max_neigh_cut == 112
    squeezed = L.Config(**{**cfg.__dict__, "x": cfg.x * 0.90})      # 10 % compression: the next shell (24 atoms at 6.99 A) enters Rc
    pair.compute(3, 0, squeezed, ago=1)                              # same list (ago > 0)
    st = pair.stats()
    assert st.max_neigh_cut > 112 + 12
    from oracle import restatement
    ref = restatement.compute(read_potential(fe_pot_file, ["Fe"]), squeezed, nthreads=4)
    f1 = pair.compute(3, 0, squeezed, ago=1)
    assert np.abs(f1 - ref["f"]).max() <= TOL_F
    pair.clear()


def test_empty_and_tiny_inputs(fe_pot_file):
    pair = make_pair(fe_pot_file)
    # one isolated atom: no neighbours, energy is the network at G = -s*avg
    one = L.build_config(np.array([[5.0, 5.0, 5.0]]), np.array([50.0, 50, 50]), 6.5, periodic=(False, False, False))
    f = pair.compute(3, 1, one, ago=0)
    from oracle import restatement
    ref = restatement.compute(read_potential(fe_pot_file, ["Fe"]), one)
    assert np.all(f == 0.0) and abs(pair.eng_vdwl - ref["eng_vdwl"]) <= TOL_E
    # dimer and trimer
    for pts in ([[5, 5, 5], [7.4, 5, 5]], [[5, 5, 5], [7.4, 5, 5], [6.0, 7.2, 5.3]]):
        c = L.build_config(np.array(pts, dtype=float), np.array([50.0, 50, 50]), 6.5, periodic=(False, False, False))
        f = pair.compute(3, 1, c, ago=0)
        ref = restatement.compute(read_potential(fe_pot_file, ["Fe"]), c)
        assert np.abs(f - ref["f"]).max() <= TOL_F and np.abs(pair.eatom - ref["eatom"]).max() <= TOL_E
    # no atoms at all
    empty = L.Config(nlocal=0, nghost=0, x=np.zeros((0, 3)), type=np.zeros(0, dtype=np.int32), ghost_owner=np.zeros(0, dtype=np.int32),
                     ilist=np.zeros(0, dtype=np.int32), numneigh=np.zeros(0, dtype=np.int32), neigh=np.zeros(0, dtype=np.int32), box=np.ones(3))
    f = pair.compute(3, 1, empty, ago=0)
    assert f.shape == (0, 3) and pair.eng_vdwl == 0.0
    pair.clear()


def test_compute_before_neighbour_list_is_an_error(fe_pot_file):
    pair = make_pair(fe_pot_file)
    cfg, _, _ = util.load_case("bcc4_perfect")
    with pytest.raises(capi.AnnpError) as ei:
        pair.compute(1, 0, cfg, ago=5)
    assert ei.value.code == capi.ESTATE
    pair.clear()


@pytest.mark.parametrize("name", ["bcc4_perturbed", "cluster_ragged", "bcc4_two_types"])
def test_lammps_pair_style_through_the_shim_driver(name, fe_pot_file):
    """The C++ class LAMMPS would compile (PairANNPB200, registered as annp/gpu) run by the same driver that
    runs the reference's PairANNP: settings -> coeff -> init_style -> compute, LAMMPS flag semantics."""
    from oracle import run_ref
    if not run_ref.available("plugin_annp_b200"):
        pytest.skip("plugin binary not built")
    cfg, elems, ref = util.load_case(name)
    for vflag, vkey in ((1 + 4, "virial_pair"), (2, "virial_fdotr")):
        out = run_ref.run_reference("plugin_annp_b200", cfg, fe_pot_file, elems, eflag=3, vflag=vflag)
        assert abs(out["eng_vdwl"] - ref["eng_vdwl"]) <= 1e-12 * abs(ref["eng_vdwl"]) + TOL_E
        assert np.abs(out["eatom"] - ref["eatom"]).max() <= TOL_E
        assert np.abs(out["f"] - ref["f"]).max() <= TOL_F
        assert np.abs(out["virial"] - ref[vkey]).max() <= TOL_V
        if vflag & 4:
            assert np.abs(out["vatom"] - ref["vatom"]).max() <= TOL_F
    # ANNP_B200_SCATTER (read by the pair class at init_style) switches the force accumulation; anything else is an error
    import os
    fixed = run_ref.run_reference("plugin_annp_b200", cfg, fe_pot_file, elems, eflag=1, vflag=0)["f"]
    os.environ["ANNP_B200_SCATTER"] = "gather"
    try:
        gathered = run_ref.run_reference("plugin_annp_b200", cfg, fe_pot_file, elems, eflag=1, vflag=0)["f"]
        os.environ["ANNP_B200_SCATTER"] = "sideways"
        with pytest.raises(RuntimeError):
            run_ref.run_reference("plugin_annp_b200", cfg, fe_pot_file, elems, eflag=1, vflag=0)
    finally:
        del os.environ["ANNP_B200_SCATTER"]
    assert np.abs(gathered - ref["f"]).max() <= TOL_F and np.abs(gathered - fixed).max() <= 2e-11


@pytest.mark.parametrize("name", ["bcc4_perturbed", "cluster_ragged"])
def test_lammps_pair_style_host_paths(name, fe_pot_file):
    """The paths of PairANNPB200::compute the single-call test above does not reach:
    * repeated calls on one list (neighbor->ago > 0: no list upload, types not re-sent, LAMMPS' x / f page-locked in place
      and the forces written straight into atom->f);
    * the sub-style situation of pair hybrid (Force::pair is another object): forces staged and ADDED;
    * ANNP_B200_NEIGH=device, the reference's `package gpu N neigh yes` (annp_gpu_compute_n): the list is built on the
      device from the positions, in a different order than LAMMPS' - same atoms in every row, so the same forces to
      rounding."""
    from oracle import run_ref
    if not run_ref.available("plugin_annp_b200"):
        pytest.skip("plugin binary not built")
    cfg, elems, ref = util.load_case(name)
    one = run_ref.run_reference("plugin_annp_b200", cfg, fe_pot_file, elems, eflag=3, vflag=1 + 4)
    many = run_ref.run_reference("plugin_annp_b200", cfg, fe_pot_file, elems, eflag=3, vflag=1 + 4, ncalls=4)
    assert len(many["per_call_seconds"]) == 4
    assert np.array_equal(many["f"], one["f"]) and np.array_equal(many["eatom"], one["eatom"]) and np.array_equal(many["vatom"], one["vatom"])
    hyb = run_ref.run_reference("plugin_annp_b200", cfg, fe_pot_file, elems, eflag=3, vflag=1 + 4, ncalls=3, env_extra={"ANNP_DRIVER_HYBRID": "1"})
    assert np.array_equal(hyb["f"], one["f"]) and hyb["eng_vdwl"] == one["eng_vdwl"]
    dev = run_ref.run_reference("plugin_annp_b200", cfg, fe_pot_file, elems, eflag=3, vflag=1 + 4, ncalls=2, env_extra={"ANNP_B200_NEIGH": "device"})
    assert np.abs(dev["eatom"] - ref["eatom"]).max() <= TOL_E
    assert np.abs(dev["f"] - ref["f"]).max() <= TOL_F
    assert np.abs(dev["virial"] - ref["virial_pair"]).max() <= TOL_V
    assert np.abs(dev["vatom"] - ref["vatom"]).max() <= TOL_F


def test_bad_neighbour_lists_are_refused(fe_pot_file):
    """A list that names atoms outside 0..nall-1 or whose offsets decrease must fail with EINVAL instead of indexing
    device memory out of bounds."""
    cfg, elems, ref = util.load_case("bcc4_perturbed")
    pair = make_pair(fe_pot_file, elems)
    Lb = capi.lib()
    ilist = np.ascontiguousarray(cfg.ilist, dtype=np.int32)
    off = np.ascontiguousarray(cfg.offsets, dtype=np.int64)
    neigh = np.ascontiguousarray(cfg.neigh, dtype=np.int32)
    ip = lambda a: a.ctypes.data_as(capi.c_int_p)
    call = lambda il, o, ng, nall=cfg.nall: Lb.annp_b200_neigh_csr(pair.handle, len(il), nall, ip(il), o.ctypes.data_as(capi.c_int64_p), ip(ng))
    bad = neigh.copy(); bad[17] = cfg.nall + 3
    assert call(ilist, off, bad) == capi.EINVAL
    assert call(ilist, off, neigh, nall=cfg.nall // 2) == capi.EINVAL       # nall disagrees with the list
    bad_il = ilist.copy(); bad_il[5] = -1
    assert call(bad_il, off, neigh) == capi.EINVAL
    bad_off = off.copy(); bad_off[3] = bad_off[2] - 1
    assert call(ilist, bad_off, neigh) == capi.EINVAL
    with pytest.raises(capi.AnnpError):                                      # no list is installed after a refusal
        pair.compute(1, 0, cfg, ago=1)
    assert call(ilist, off, neigh) == 0
    assert np.abs(pair.compute(1, 0, cfg, ago=1) - ref["f"]).max() <= TOL_F
    pair.clear()


# ---- general potentials: everything the reference's CPU style reads must run on the GPU ------------------

@pytest.mark.parametrize("name", util.GENERAL_CASES)
def test_general_potentials_against_the_reference(name, tmp_path):
    """Descriptor shapes without an exact kernel instantiation (zero-padded onto <8,24> / <16,24>, the latter with the
    two-round descriptor reduction), the exact <8,20> instantiation, activations 1, 2 (with the reference's sign), 3, a
    non-linear output layer, five layers, 32 nodes per layer, and a two-element file: golden answers of the UNMODIFIED
    reference (tests/golden/make_golden.py general)."""
    from meng_zhang_b200.pair import write_potential
    cfg, elems, ref, pot = util.load_general_case(name)
    pf = str(tmp_path / f"{name}.ann")
    write_potential(pf, pot)
    pair = make_pair(pf, elems)
    f = pair.compute(3, 1 + 4, cfg, ago=0)
    assert np.abs(pair.eatom - ref["eatom"]).max() <= TOL_E
    assert np.abs(f - ref["f"]).max() <= TOL_F
    assert np.abs(pair.virial - ref["virial_pair"]).max() <= TOL_V
    assert np.abs(pair.vatom - ref["vatom"]).max() <= TOL_F
    assert np.array_equal(f, pair.compute(0, 0, cfg, ago=1))
    # the ordered-gather scatter runs the other template instantiation of the same shape
    pair.set_scatter(capi.SCATTER_GATHER)
    assert np.abs(pair.compute(1, 0, cfg, ago=1) - ref["f"]).max() <= TOL_F
    # descriptors come back in the potential's own layout whatever the kernel's padded shape is
    G, dE = pair.descriptors(cfg)
    assert G.shape == (cfg.nlocal, pot.nsf)
    from oracle import restatement
    parsed = read_potential(pf, elems)
    o = restatement.compute(parsed, cfg, ntypes=len(elems), type_map=[0] + list(range(len(elems))), dump_G=True)
    assert np.abs(G - o["G"]).max() <= 1e-9 * max(1.0, np.abs(o["G"]).max())
    pair.clear()


def test_two_elements_with_two_different_networks(tmp_path):
    """The reference's FILE FORMAT cannot carry a second network (see test above), but its compute() selects weights per
    element (fe_v2/src/pair_annp.cpp:114,181: all_annp[itype]).  Give the C ABI two different non-zero networks and a
    non-trivial type -> element map (types 1,3 -> element 1, type 2 -> element 0) and compare with the pinned
    restatement of that compute()."""
    from meng_zhang_b200.pair import write_potential
    from oracle import restatement
    pot, _ = util.general_potential("two_elements")
    assert np.abs(pot.weight_all[0] - pot.weight_all[1]).max() > 0.1
    x, box = L.bcc(4, 4, 4)
    types = np.random.default_rng(5).integers(1, 4, size=len(x)).astype(np.int32)
    cfg = L.build_config(L.perturb(x, 0.05, 99), box, pot.cut, types=types, shuffle_rows=2)
    pf = str(tmp_path / "two.ann")
    write_potential(pf, pot)
    pair = PairANNPGPU(ntypes=3)
    pair.settings([])
    pair.coeff(["*", "*", pf, "Cr", "Fe", "Cr"])
    assert list(pair.map[1:]) == [0, 1, 0]           # pair_coeff order: Cr is element 0 of the style
    pair.params = pot                                 # both networks, bypassing the single-block reader
    pair.map[1:] = [1, 0, 1]
    pair.init_style()
    f = pair.compute(3, 1, cfg, ago=0)
    o = restatement.compute(pot, cfg, ntypes=3, type_map=[0, 1, 0, 1], nthreads=2)
    assert np.abs(pair.eatom - o["eatom"]).max() <= TOL_E
    assert np.abs(f - o["f"]).max() <= TOL_F
    assert np.abs(pair.virial - o["virial"]).max() <= TOL_V
    # and the two networks really are different: swapping the map changes the answer
    e0 = pair.eatom.copy()
    pair.map[1:] = [0, 1, 0]
    pair.init_style()
    pair.compute(3, 0, cfg, ago=0)
    assert np.abs(pair.eatom - e0).max() > 1e-3
    pair.clear()


def test_shapes_beyond_the_kernel_range_are_refused(tmp_path):
    from meng_zhang_b200.pair import write_potential
    pot, _ = util.general_potential("shape_16_24")
    pot.ntsf, pot.nsf = 25, 41
    pot.sfnor_cov = np.append(pot.sfnor_cov, pot.sfnor_cov[-1]); pot.sfnor_avg = np.append(pot.sfnor_avg, pot.sfnor_avg[-1])
    w = np.zeros((1, pot.ntl - 1, pot.nnod, 41)); w[..., :40] = pot.weight_all
    pot.weight_all = w
    pf = str(tmp_path / "big.ann")
    write_potential(pf, pot)
    pair = PairANNPGPU(ntypes=1)
    pair.settings([])
    pair.coeff(["*", "*", pf, "Fe"])
    with pytest.raises(capi.AnnpError) as ei:
        pair.init_style()
    assert ei.value.code == capi.EINVAL


def test_atoms_that_outgrow_the_tile_take_the_overflow_pass(fe_pot_file):
    """Device-resident mode sizes the first-pass shared-memory tile once per list (in-cutoff maximum + 12).  Atoms that
    gain more neighbours than that inside the skin are redone by the overflow pass of the same step with a tile of the
    list's longest row: compress the box by 8 % after the tile was sized (the 24-atom shell at 6.99 A moves inside Rc) and the
    forces still match an evaluation that sized its tile for the compressed box."""
    import torch
    x, box = L.bcc(6, 6, 6)
    x = L.perturb(x, 0.05, 3)
    cfg = L.build_config(x, box, 6.5)
    scale = 0.92
    cfg2 = L.Config(nlocal=cfg.nlocal, nghost=cfg.nghost, x=cfg.x * scale, type=cfg.type, ghost_owner=cfg.ghost_owner, ilist=cfg.ilist,
                    numneigh=cfg.numneigh, neigh=cfg.neigh, box=cfg.box * scale)
    ref_pair = make_pair(fe_pot_file)
    f_ref = ref_pair.compute(1, 0, cfg2, ago=0)                 # host mode: tile sized for the compressed positions
    e_ref = ref_pair.eng_vdwl
    assert ref_pair.stats().overflow_pass_atoms == 0
    pair = make_pair(fe_pot_file)
    pair.upload_neighbors(cfg)
    Lb = capi.lib()
    dev = torch.device("cuda")
    typ = torch.as_tensor(cfg.type, dtype=torch.int32, device=dev)
    f = torch.zeros((cfg.nall, 3), dtype=torch.float64, device=dev)
    ev = torch.zeros(8, dtype=torch.float64, device=dev)
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def step(xs):
        xd = torch.as_tensor(xs, dtype=torch.float64, device=dev)
        rc = Lb.annp_b200_compute_device(pair.handle, cfg.nlocal, cfg.nghost, C.c_void_p(xd.data_ptr()), C.c_void_p(typ.data_ptr()), 1, 0,
                                         C.c_void_p(f.data_ptr()), None, C.c_void_p(ev.data_ptr()), None, s)
        assert rc == 0
        torch.cuda.synchronize()
        return f.cpu().numpy().copy(), float(ev[0])

    step(cfg.x)                                                  # sizes the tile: 112-113 neighbours + 12
    st = pair.stats()
    cap0 = st.tile_capacity
    assert st.overflow_pass_atoms == 0 and cap0 < 136
    f2, e2 = step(cfg2.x)                                        # ~136 neighbours per atom now
    st = pair.stats()                                            # no error: the step is complete
    assert st.overflow_pass_atoms > 0 and st.max_neigh_cut > cap0
    assert np.abs(f2 - f_ref).max() <= 1e-11 and abs(e2 - e_ref) <= 1e-9 * abs(e_ref)
    assert st.tile_capacity > cap0                               # the next steps size the first pass for them
    f3, e3 = step(cfg2.x)
    assert pair.stats().overflow_pass_atoms == 0 and np.array_equal(f3, f2)
    pair.clear(); ref_pair.clear()

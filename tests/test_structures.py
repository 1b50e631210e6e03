"""Structure generators (meng_zhang_b200/structures.py) against the reference programs' own output.

tests/golden/gen_screw.npz and gen_stgb.npz hold the atoms produced by the UNMODIFIED reference generators
(screw-dislocation-bcc-fe/screw_dislocation_bcc_fe.cpp, symmetry_tilt_grain_boundary/stgb.cpp + stgb_b.cpp, compiled by
oracle/Makefile into oracle/_ref/gen_screw / gen_stgb; regenerate with `python tests/golden/make_golden.py structures`).
"""
import os
import subprocess

import numpy as np
import pytest
from scipy.spatial import cKDTree

import util
from meng_zhang_b200 import structures as S

GEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")


def same_atoms(a, ta, b, tb, tol=1e-9):
    assert len(a) == len(b)
    d, i = cKDTree(b).query(a)
    assert d.max() <= tol and len(np.unique(i)) == len(a)
    assert np.array_equal(ta, tb[i])


def test_screw_block_equals_reference_program():
    z = np.load(os.path.join(util.GOLDEN, "gen_screw.npz"))
    x, box, types = S.screw_block(reference_rules=True)
    assert np.allclose(box, z["box"], rtol=0, atol=1e-12) and len(x) == 5115          # 5 115 atoms
    same_atoms(x, types, z["x"], z["type"])


def test_screw_displacement_field_equals_reference_function():
    """screw_dislocation() of the reference (interactive: three atom ids from stdin) against apply_screw."""
    z = np.load(os.path.join(util.GOLDEN, "gen_screw.npz"))
    x, box, types = S.screw_block(reference_rules=True)
    core, _ = S.screw_core(x, box)
    assert np.allclose(core, z["core"], atol=1e-9)
    xs = S.apply_screw(x, core)
    same_atoms(xs, types, z["x_screw"], z["type"], tol=1e-9)
    # Burgers circuit: going once around the core the z displacement changes by b = sqrt(3) a / 2
    u = xs[:, 2] - x[:, 2]
    assert abs((u.max() - u.min()) - np.sqrt(3) * S.A_FE / 2) < 0.05
    assert u.min() >= 0.0


def test_stgb_equals_reference_program():
    z = np.load(os.path.join(util.GOLDEN, "gen_stgb.npz"))
    x, box, types = S.stgb(reference_rules=True)
    assert np.allclose(box, z["box"], rtol=0, atol=1e-12) and len(x) == 10240
    same_atoms(x, types, z["x"], z["type"])


@pytest.mark.skipif(not os.access(os.path.join(GEN, "gen_stgb"), os.X_OK), reason="oracle/_ref/gen_stgb not built")
def test_live_reference_generators(tmp_path):
    for prog, fn in (("gen_screw", lambda: S.screw_block(reference_rules=True)), ("gen_stgb", lambda: S.stgb(reference_rules=True))):
        out = tmp_path / f"{prog}.txt"
        subprocess.run([os.path.join(GEN, prog), str(out)], check=True, capture_output=True, cwd=tmp_path)
        with open(out) as fp:
            n = int(fp.readline())
            box = np.array(fp.readline().split(), dtype=float)
            arr = np.loadtxt(fp)
        x, b, t = fn()
        assert n == len(x) and np.allclose(b, box, atol=1e-12)
        same_atoms(x, t, arr[:, 2:5], arr[:, 1].astype(np.int32))


def test_md_ready_structures_are_periodic_and_overlap_free():
    # STGB: two grains, periodic in all three directions after wrapping, no pair closer than 0.5 A, both GBs present
    x, box, types = S.stgb()
    assert set(np.unique(types)) == {1, 2}
    assert np.all(x >= 0.0) and np.all(x < box)
    d = cKDTree(x, boxsize=box).query(x, k=2)[0][:, 1]
    assert d.min() > 2.4 and len(x) == 9280                      # closest pair = bcc nearest-neighbour distance
    assert cKDTree(S.stgb(overlap=0.5)[0], boxsize=box).query(S.stgb(overlap=0.5)[0], k=2)[0][:, 1].min() > 0.5
    bulk = np.abs(d - np.sqrt(3) / 2 * S.A_FE) < 1e-6          # nearest-neighbour distance of perfect bcc
    assert bulk.mean() > 0.8                                     # everything but the two boundary regions
    # a larger box scales the atom count with the volume (the reference is capped by its +-31 cell replication)
    u = S.stgb_unit_lengths()
    x2, box2, _ = S.stgb(length_box=(10 * u[0], 12 * u[1], 10 * u[2]))
    assert abs(len(x2) / (2 * np.prod(box2) / S.A_FE ** 3) - 1.0) < 0.03
    # screw dislocation, config-4 shape at reduced thickness: periodic along z, one Burgers vector of mismatch
    xs, boxs, ts, core = S.screw_dislocation(num_lattice=(22, 38, 2))
    assert np.all(xs[:, 2] >= 0.0) and np.all(xs[:, 2] < boxs[2])
    assert len(xs) == 4 * len(S.screw_block(num_lattice=(22, 38, 0.5))[0])
    dz = cKDTree(xs + np.array([1.0, 1.0, 0.0]), boxsize=[1e6, 1e6, boxs[2]]).query(xs + np.array([1.0, 1.0, 0.0]), k=2)[0][:, 1]
    assert dz.min() > 2.0                                        # no overlapping atoms, also across the periodic z face
    # full config 4 would be (22, 38, 50): 100 b thick, ~5.1e5 atoms
    assert abs(S.screw_box((22, 38, 50))[2] / (np.sqrt(3) / 2 * S.A_FE) - 100) < 1e-9

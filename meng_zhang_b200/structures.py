"""Defect structures of the BASELINE configurations (host code, numpy): oriented bcc blocks, the screw dislocation
of screw-dislocation-bcc-fe/screw_dislocation_bcc_fe.cpp (config 4) and the symmetric tilt grain boundary of
symmetry_tilt_grain_boundary/stgb.cpp + stgb_b.cpp (config 5 variant).

The reference programs replicate a two-atom basis +-31 cells along each cubic axis, rotate every atom by an Euler
matrix and clip to the box (screw...cpp:103-170, stgb_b.cpp:104-170), which caps the model at a ~177 A cube and costs
O(62^3) whatever the box.  Here the lattice points inside the box are enumerated directly in the rotated frame, so
the size is a parameter.  `reference_rules=True` reproduces the reference's conventions
bit-for-geometry (closed clip intervals, its lattice origin, its +-31 replication limit, no overlap removal) and is
what tests/test_structures.py compares with the reference programs' own output; the default produces boxes that are
ready for periodic MD (half-open intervals, overlapping atoms at the boundary planes removed).
"""
from __future__ import annotations

import numpy as np

A_FE = 2.8553          # screw_dislocation_bcc_fe.cpp:21, stgb.cpp:19
BCC_BASIS = np.array([[0.0, 0.0, 0.0], [0.5, 0.5, 0.5]])


def unit_rows(orient) -> np.ndarray:
    """Rows = box axes expressed in cubic axes, normalised (Box::get_length_unitorient, stgb.cpp:26-31)."""
    o = np.asarray(orient, dtype=np.float64)
    return o / np.linalg.norm(o, axis=1, keepdims=True)


def euler_matrix(unit_orient) -> np.ndarray:
    """The rotation the reference programs apply to the cubic-frame coordinates (get_euler_angle + rotation_euler,
    screw...cpp:52-99; CRY_BOX::get_euler_angle, stgb_b.cpp:36-70): z-x-z Euler angles that carry the cubic axes onto
    the box axes.  It is always a proper rotation, so for the orientations shipped it equals the unit-row matrix up to
    mirror operations of the cubic lattice (x=[11-2],y=[1-10],z=[-1-1-1] -> the y axis comes out reversed; the
    left-handed STGB triple -> one cubic axis mirrored).  Using the same matrix keeps atom positions identical."""
    o = np.asarray(unit_orient, dtype=np.float64)
    z1 = np.sqrt(o[2, 0] ** 2 + o[2, 1] ** 2)
    if z1 > np.finfo(np.float64).eps:
        x11 = np.array([o[2, 1], -o[2, 0], 0.0])
        psi = np.arctan2(x11 @ o[1], x11 @ o[0])
        the = np.arctan2(z1, o[2, 2])
        fai = -np.arctan2(x11[1], x11[0])
    else:
        psi, the, fai = 0.0, (0.0 if o[2, 2] > 0.0 else np.pi), -np.arctan2(o[0, 1], o[0, 0])
    c, sn = np.cos, np.sin
    return np.array([[c(psi) * c(fai) - c(the) * sn(fai) * sn(psi), c(psi) * sn(fai) + c(the) * c(fai) * sn(psi), sn(psi) * sn(the)],
                     [-sn(psi) * c(fai) - c(the) * sn(fai) * c(psi), -sn(psi) * sn(fai) + c(the) * c(fai) * c(psi), c(psi) * sn(the)],
                     [sn(the) * sn(fai), -sn(the) * c(fai), c(the)]])


def oriented_lattice(E: np.ndarray, lo, hi, a: float, start, sub, add, closed: bool, nmax: int | None = None,
                     basis: np.ndarray = BCC_BASIS, eps: float = 1e-9) -> np.ndarray:
    """All points p = E ((start + a basis + a n) - sub) + add inside the box, n integer (|n_i| <= nmax if given: the
    reference's replication limit).  The arithmetic is done in the reference's order (lattice point, subtract, rotate
    with a left-to-right sum, add) so that with eps = 0 even atoms lying on a clip plane are classified identically.

    closed: lo <= p <= hi; otherwise lo <= p < hi.  eps widens the lower and (closed) upper bounds."""
    E = np.asarray(E, dtype=np.float64)
    lo, hi = np.asarray(lo, dtype=np.float64), np.asarray(hi, dtype=np.float64)
    start, sub, add = (np.asarray(v, dtype=np.float64) for v in (start, sub, add))
    corners = np.array([[x, y, z] for x in (lo[0], hi[0]) for y in (lo[1], hi[1]) for z in (lo[2], hi[2])])
    q = ((corners - add) @ E + sub - start) / a               # cubic cell coordinates of the box corners (E^-1 = E^T)
    nlo = np.floor(q.min(axis=0) - 1.0).astype(np.int64)
    nhi = np.ceil(q.max(axis=0) + 1.0).astype(np.int64)
    if nmax is not None:
        nlo, nhi = np.maximum(nlo, -nmax), np.minimum(nhi, nmax)
    out = []
    j, k = np.meshgrid(np.arange(nlo[1], nhi[1] + 1), np.arange(nlo[2], nhi[2] + 1), indexing="ij")
    jk = np.stack([j.ravel(), k.ravel()], axis=1).astype(np.float64)
    for i in range(int(nlo[0]), int(nhi[0]) + 1):             # slab by slab bounds the temporary memory
        n = np.concatenate([np.full((len(jk), 1), float(i)), jk], axis=1)
        for b in basis:
            cub = ((start + a * b) + n * a) - sub
            p = np.empty_like(cub)
            for d in range(3):
                p[:, d] = ((E[d, 0] * cub[:, 0] + E[d, 1] * cub[:, 1]) + E[d, 2] * cub[:, 2]) + add[d]
            if closed:
                m = np.all((p >= lo - eps) & (p <= hi + eps), axis=1)
            else:
                m = np.all((p >= lo - eps) & (p < hi - eps), axis=1)
            if m.any():
                out.append(p[m])
    return np.concatenate(out) if out else np.zeros((0, 3))


# ------------------------------------------------------------------------------------------------ screw dislocation
SCREW_ORIENT = [[1, 1, -2], [1, -1, 0], [-1, -1, -1]]           # screw_dislocation_bcc_fe.cpp:28


def screw_box(num_lattice=(22, 38, 0.5), a: float = A_FE, orient=SCREW_ORIENT):
    """Box edges = num_lattice[i] * |orient[i]| * a (cpp:29-37).  The reference's z entry counts [111] vectors of
    length sqrt(3) a; one Burgers vector is half of that, so nz = 0.5 is one b and nz = 50 is 100 b."""
    o = np.asarray(orient, dtype=np.float64)
    return np.asarray(num_lattice, dtype=np.float64) * np.linalg.norm(o, axis=1) * a


def screw_block(num_lattice=(22, 38, 0.5), a: float = A_FE, orient=SCREW_ORIENT, reference_rules: bool = False):
    """Perfect bcc block in the dislocation frame x=[11-2], y=[1-10], z=[-1-1-1] (building_matrix, cpp:103-170).

    Returns (x[n,3], box[3], type[n]); type 2 marks the outer shell `dis > Lx/2 - 10` of cpp:160-168 (the rim the
    deck holds fixed).  reference_rules: closed clip and the +-31-cell replication limit of the reference."""
    E = euler_matrix(unit_rows(orient))
    box = screw_box(num_lattice, a, orient)
    c = box / 2.0
    # basic_atom1 sits at the box centre; every atom is shifted by -L/2, rotated, shifted back (cpp:113-150)
    x = oriented_lattice(E, np.zeros(3), box, a, start=c, sub=c, add=c, closed=reference_rules,
                         nmax=31 if reference_rules else None, eps=0.0 if reference_rules else 1e-9)
    dis = np.sqrt((x[:, 0] - box[0] / 2) ** 2 + (x[:, 1] - box[1] / 2) ** 2)
    types = np.where(dis > box[0] / 2.0 - 10, 2, 1).astype(np.int32)
    return x, box, types


def screw_core(x: np.ndarray, box: np.ndarray, a: float = A_FE):
    """Dislocation line position: between three neighbouring [111] atomic columns nearest to the box centre, chosen as
    the reference asks its user to (README: two atoms parallel to x, the third on the vertex):
    core = (mid x of the pair, y0 + (y2 - y0)/3)  (screw_dislocation, cpp:224-226)."""
    # [111] columns project onto a triangular lattice in the x-y plane: spacing sqrt(6) a/3 along x, sqrt(2) a/2 in y
    _, first = np.unique(np.round(x[:, :2], 6), axis=0, return_index=True)
    cols = x[first, :2]                                        # one representative atom per column, unrounded
    d = np.linalg.norm(cols - box[:2] / 2, axis=1)
    c0 = cols[np.argmin(d)]
    same_row = cols[np.abs(cols[:, 1] - c0[1]) < 1e-4]
    right = same_row[same_row[:, 0] > c0[0] + 1e-4]
    c1 = right[np.argmin(right[:, 0])]
    xm = 0.5 * (c0[0] + c1[0])
    above = cols[(cols[:, 1] > c0[1] + 1e-4)]
    c2 = above[np.argmin(np.abs(above[:, 0] - xm) + np.abs(above[:, 1] - c0[1]))]
    return np.array([xm, c0[1] + (c2[1] - c0[1]) / 3.0]), (c0, c1, c2)


def apply_screw(x: np.ndarray, core_xy, a: float = A_FE) -> np.ndarray:
    """Elastic displacement field of a screw dislocation, u_z = b/(2 pi) theta with theta in [0, 2 pi) measured from
    the atom towards the core (cpp:228-236): b = sqrt(3) a / 2."""
    out = x.copy()
    rx, ry = -x[:, 0] + core_xy[0], -x[:, 1] + core_xy[1]
    th = np.arctan2(ry, rx)
    th = np.where(ry >= 0.0, th, 2.0 * np.pi + th)
    out[:, 2] += np.sqrt(3.0) * a / 2.0 / (2.0 * np.pi) * th
    return out


def screw_dislocation(num_lattice=(22, 38, 50), a: float = A_FE):
    """BASELINE config 4: ~5e5-atom bcc Fe cylinder-in-a-box with one 1/2[111] screw dislocation along z.
    Periodic along z (the displacement is independent of z), free surfaces in x and y.
    Returns (x, box, type, core_xy)."""
    x, box, types = screw_block(num_lattice, a)
    core, _ = screw_core(x, box, a)
    xs = apply_screw(x, core, a)
    xs[:, 2] -= np.floor(xs[:, 2] / box[2]) * box[2]        # wrap the shifted atoms back into the periodic z range
    return xs, box, types, core


# ------------------------------------------------------------------------------------------ symmetric tilt boundary
STGB_ORIENT = [[-1, 1, -2], [1, -1, -1], [1, 1, 0]]            # stgb.cpp:21
STGB_LENGTH = (34.97014031, 49.45524671, 32.30403188)          # stgb.cpp:22 = (5 sqrt6, 10 sqrt3, 8 sqrt2) a


def stgb_unit_lengths(a: float = A_FE, orient=STGB_ORIENT):
    return np.linalg.norm(np.asarray(orient, dtype=np.float64), axis=1) * a * np.array([1.0, 1.0, 1.0])


def stgb(length_box=STGB_LENGTH, a: float = A_FE, orient=STGB_ORIENT, reference_rules: bool = False, overlap: float = 1.5):
    """Bicrystal with two symmetric tilt boundaries (build_crystal + symm_crystal, stgb_b.cpp:104-188).

    Grain 1 (type 1) fills -1 <= x <= Lx + 1, grain 2 (type 2) is its mirror image at x = Lx; the box is 2 Lx long.
    reference_rules=True returns exactly the reference program's atoms (closed clips; atoms of the two grains overlap
    at the boundary planes and the README tells the user to delete them by hand).  Otherwise the structure is made
    ready for periodic MD: y, z half-open, x wrapped into [0, 2 Lx), and of any two atoms closer than `overlap` the one
    with the higher index is removed.  The mirror plane of the shipped size falls between two (112) layers, so the
    two grains interleave within +-1 A of each boundary with pairs 1.10 A apart; 0.5 A (the literal reading of
    "overlap") keeps those, the default 1.5 A leaves a boundary whose closest pair is the bcc nearest-neighbour
    distance (9 280 instead of 10 240 atoms at the default size).
    Returns (x[n,3], box[3], type[n])."""
    E = euler_matrix(unit_rows(orient))
    L = np.asarray(length_box, dtype=np.float64)
    lo = np.array([-1.0, 0.0, 0.0])
    hi = np.array([L[0] + 1.0, L[1], L[2]])
    # atom1 at the cubic origin; -L/2 in the cubic frame, rotation, +L/2 (stgb_b.cpp:118-160)
    if reference_rules:
        g1 = oriented_lattice(E, lo, hi, a, start=np.zeros(3), sub=L / 2, add=L / 2, closed=True, nmax=31, eps=0.0)
    else:
        # closed in x (the mirror needs both faces), half-open in the periodic directions
        g1 = oriented_lattice(E, lo, hi, a, start=np.zeros(3), sub=L / 2, add=L / 2, closed=True)
        g1 = g1[(g1[:, 1] < L[1] - 1e-6) & (g1[:, 2] < L[2] - 1e-6)]
    g2 = g1.copy()
    g2[:, 0] = 2.0 * L[0] - g1[:, 0]
    x = np.concatenate([g1, g2])
    types = np.concatenate([np.ones(len(g1), dtype=np.int32), np.full(len(g2), 2, dtype=np.int32)])
    box = np.array([2.0 * L[0], L[1], L[2]])
    if reference_rules:
        return x, box, types
    x[:, 0] -= np.floor(x[:, 0] / box[0]) * box[0]
    keep = _remove_overlaps(x, box, overlap)
    return x[keep], box, types[keep]


def _remove_overlaps(x: np.ndarray, box: np.ndarray, dmin: float) -> np.ndarray:
    from scipy.spatial import cKDTree
    xw = x - np.floor(x / box) * box
    xw = np.minimum(xw, np.nextafter(box, 0.0))
    pairs = cKDTree(xw, boxsize=box).query_pairs(dmin, output_type="ndarray")
    keep = np.ones(len(x), dtype=bool)
    if len(pairs):
        keep[np.unique(pairs.max(axis=1))] = False
    return keep


def write_lammps_data(path: str, x: np.ndarray, box: np.ndarray, types: np.ndarray, ntypes: int = 2,
                      title: str = "#BCC Fe model") -> None:
    """`Atoms # atomic` data file as both reference programs write it (cpp:194-212, stgb_b.cpp:190-205)."""
    with open(path, "w") as fp:
        fp.write(f"{title}\n{len(x)} atoms\n{ntypes} atom types\n")
        for d, n in enumerate("xyz"):
            fp.write(f"0 {box[d]:.10g} {n}lo {n}hi\n")
        fp.write("\nAtoms # atomic\n\n")
        for i, (p, t) in enumerate(zip(x, types), start=1):
            fp.write(f"{i} {int(t)} {p[0]:.10g} {p[1]:.10g} {p[2]:.10g}\n")

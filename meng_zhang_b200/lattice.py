"""Synthetic configurations for the ANNP hot path (SURVEY.md 8d "concrete synthetic inputs").

Pure numpy host code: lattices, thermal-like perturbation, periodic ghost shells and the
full neighbour list a LAMMPS `full/bin/atomonly` build would hand to `Pair::compute`
(reference consumer: annp-gpu-lammps/fe_v2/src/pair_annp.cpp:89-92,134-136).
"""
from __future__ import annotations

import dataclasses

import numpy as np

A_FE = 2.8553    # screw-dislocation-bcc-fe/screw_dislocation_bcc_fe.cpp:21, stgb.cpp:19
A_NI = 3.52      # not given by the reference
MASS_FE = 55.845  # in.st_test:21


def bcc(nx: int, ny: int, nz: int, a: float = A_FE):
    """bcc supercell, atom order (i, j, k, basis). Returns (x[n,3], box[3])."""
    basis = np.array([[0.0, 0.0, 0.0], [0.5, 0.5, 0.5]])
    return _cubic(nx, ny, nz, a, basis)


def fcc(nx: int, ny: int, nz: int, a: float = A_NI):
    basis = np.array([[0.0, 0.0, 0.0], [0.5, 0.5, 0.0], [0.5, 0.0, 0.5], [0.0, 0.5, 0.5]])
    return _cubic(nx, ny, nz, a, basis)


def _cubic(nx, ny, nz, a, basis):
    i, j, k = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    cells = np.stack([i, j, k], axis=-1).reshape(-1, 1, 3).astype(np.float64)
    x = (cells + basis[None, :, :]).reshape(-1, 3) * a
    return np.ascontiguousarray(x), np.array([nx * a, ny * a, nz * a])


def perturb(x: np.ndarray, amp: float, seed: int) -> np.ndarray:
    """Uniform +-amp displacement per coordinate from a fixed seed (documented own RNG)."""
    rng = np.random.default_rng(seed)
    return x + rng.uniform(-amp, amp, size=x.shape)


def wrap(x: np.ndarray, box: np.ndarray, periodic=(True, True, True)) -> np.ndarray:
    x = x.copy()
    for d in range(3):
        if periodic[d]:
            x[:, d] -= np.floor(x[:, d] / box[d]) * box[d]
    return x


@dataclasses.dataclass
class Config:
    """What the pair style sees on one rank: local + ghost atoms and a full neighbour list."""
    nlocal: int
    nghost: int
    x: np.ndarray            # [nall,3] f64, locals first
    type: np.ndarray         # [nall] i32, 1-based LAMMPS types
    ghost_owner: np.ndarray  # [nghost] i32 local index owning each ghost
    ilist: np.ndarray        # [inum] i32
    numneigh: np.ndarray     # [inum] i32 (per ilist entry)
    neigh: np.ndarray        # [sum numneigh] i32 flat rows, in ilist order
    box: np.ndarray
    ghost_shift: np.ndarray | None = None  # [nghost,3] f64 image shift (x_ghost = x_owner + shift)

    @property
    def nall(self):
        return self.nlocal + self.nghost

    @property
    def offsets(self):
        off = np.zeros(len(self.numneigh) + 1, dtype=np.int64)
        np.cumsum(self.numneigh, out=off[1:])
        return off

    def fold(self, f_all: np.ndarray) -> np.ndarray:
        """Reverse communication: add ghost contributions onto their owners."""
        out = f_all[: self.nlocal].copy()
        if self.nghost:
            np.add.at(out, self.ghost_owner, f_all[self.nlocal:])
        return out


def make_ghosts(x: np.ndarray, box: np.ndarray, cutghost: float, periodic=(True, True, True)):
    """Periodic images within cutghost of the box, as LAMMPS' ghost shell for one sub-domain."""
    n = len(x)
    rng = [range(-int(np.ceil(cutghost / box[d])), int(np.ceil(cutghost / box[d])) + 1) if periodic[d]
           else range(0, 1) for d in range(3)]
    gx, gowner, gshift = [], [], []
    idx = np.arange(n)
    for sx in rng[0]:
        for sy in rng[1]:
            for sz in rng[2]:
                if sx == 0 and sy == 0 and sz == 0:
                    continue
                shift = np.array([sx, sy, sz]) * box
                y = x + shift
                m = np.all((y >= -cutghost) & (y < box + cutghost), axis=1)
                if m.any():
                    gx.append(y[m])
                    gowner.append(idx[m])
                    gshift.append(np.broadcast_to(shift, (int(m.sum()), 3)))
    if gx:
        return (np.concatenate(gx), np.concatenate(gowner).astype(np.int32),
                np.concatenate(gshift).astype(np.float64))
    return np.zeros((0, 3)), np.zeros(0, dtype=np.int32), np.zeros((0, 3))


def full_neighbor_list(x_all: np.ndarray, nlocal: int, cutneigh: float):
    """Full list (i local; j local or ghost, j != i, r < cutneigh), rows sorted by j."""
    from scipy.spatial import cKDTree
    tree = cKDTree(x_all)
    rows = tree.query_ball_point(x_all[:nlocal], cutneigh, return_sorted=True)
    numneigh = np.empty(nlocal, dtype=np.int32)
    out = []
    for i, r in enumerate(rows):
        r = np.asarray(r, dtype=np.int32)
        r = r[r != i]
        numneigh[i] = len(r)
        out.append(r)
    neigh = np.concatenate(out) if out else np.zeros(0, dtype=np.int32)
    return numneigh, neigh.astype(np.int32)


def build_config(x: np.ndarray, box: np.ndarray, cutoff: float, skin: float = 2.0,
                 periodic=(True, True, True), types: np.ndarray | None = None,
                 shuffle_rows: int | None = None) -> Config:
    x = wrap(np.asarray(x, dtype=np.float64), box, periodic)
    nlocal = len(x)
    cut = cutoff + skin
    gx, gowner, gshift = make_ghosts(x, box, cut, periodic)
    x_all = np.ascontiguousarray(np.concatenate([x, gx]))
    t_local = np.ones(nlocal, dtype=np.int32) if types is None else np.asarray(types, dtype=np.int32)
    t_all = np.concatenate([t_local, t_local[gowner]]).astype(np.int32)
    numneigh, neigh = full_neighbor_list(x_all, nlocal, cut)
    if shuffle_rows is not None:   # LAMMPS rows are in bin order, not sorted: exercise that too
        rng = np.random.default_rng(shuffle_rows)
        off = np.concatenate([[0], np.cumsum(numneigh)])
        for i in range(nlocal):
            rng.shuffle(neigh[off[i]:off[i + 1]])
    return Config(nlocal=nlocal, nghost=len(gx), x=x_all, type=t_all, ghost_owner=gowner,
                  ilist=np.arange(nlocal, dtype=np.int32), numneigh=numneigh, neigh=neigh,
                  box=np.asarray(box, dtype=np.float64), ghost_shift=gshift)

"""Pieces of LAMMPS' own behaviour that the reference's published run depends on, restated so that the run can be
replayed without LAMMPS (which is not in this image): the `velocity ... create` generator and the shrink-wrapped
(`boundary m`) box.  Host code (numpy); written from the documented behaviour of LAMMPS stable_2Aug2023
(velocity.cpp, random_park.cpp, domain.cpp) - the reference repository does not contain these files.
"""
from __future__ import annotations

import numpy as np

BOLTZ, MVV2E = 8.617343e-5, 1.0364269e-4          # units metal
_IA, _IM, _IQ, _IR = 16807, 2147483647, 127773, 2836


def ranpark_uniform(seed: int, n: int) -> np.ndarray:
    """n draws of RanPark::uniform() (Park-Miller minimal standard with Schrage's trick, random_park.cpp)."""
    out = np.empty(n)
    am = 1.0 / _IM
    s = int(seed)
    for i in range(n):
        k = s // _IQ
        s = _IA * (s - k * _IQ) - _IR * k
        if s < 0:
            s += _IM
        out[i] = am * s
    return out


def velocity_create(natoms: int, mass: float, temperature: float, seed: int) -> np.ndarray:
    """`velocity all create T seed` with LAMMPS' defaults (dist uniform, loop all, mom yes, rot no): three uniform
    deviates minus 0.5 per atom in atom-ID order, scaled by 1/sqrt(mass); centre-of-mass velocity removed; rescaled to
    T with dof = 3N - 3.  Returns v[natoms,3] in ID order."""
    u = ranpark_uniform(seed, 3 * natoms).reshape(natoms, 3) - 0.5
    v = u / np.sqrt(mass)
    v -= v.mean(axis=0, keepdims=True)              # equal masses: vcm is the plain mean
    dof = 3.0 * natoms - 3.0
    t = mass * MVV2E * float((v * v).sum()) / (dof * BOLTZ)
    return v * np.sqrt(temperature / t)


def shrink_wrap(lo_min: float, hi_min: float, xmin: float, xmax: float, small: float | None = None):
    """`boundary m`: the face follows the atoms plus a buffer but never moves inside the data-file box
    (Domain::reset_box, boundary code 3).  The buffer is SMALL = 1e-4 times the data-file box length in that direction
    (Domain::set_initial_box); the reference's log shows exactly this: Lx = 184 + 0.0184 and Lz = 112.5 + 0.01125
    before its minimisation (Volume 1773495.9 = 184.0184 x 85.659 x 112.51125, log_relaxing_new.lammps:108)."""
    if small is None:
        small = 1.0e-4 * (hi_min - lo_min)
    return min(xmin - small, lo_min), max(xmax + small, hi_min)

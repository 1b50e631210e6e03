"""ctypes loader for oracle/annp_oracle.c (the CPU restatement).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_ref", "liboracle_annp.so")

MAX_LAYERS = 8


class OracleParams(C.Structure):
    _fields_ = [
        ("ntypes", C.c_int), ("nelements", C.c_int), ("ntl", C.c_int), ("nhl", C.c_int), ("nnod", C.c_int),
        ("nsf", C.c_int), ("npsf", C.c_int), ("ntsf", C.c_int), ("flagsym", C.c_int), ("flagact", C.c_int * MAX_LAYERS),
        ("cut", C.c_double), ("e_scale", C.c_double), ("e_shift", C.c_double), ("e_atom", C.c_double),
        ("sfnor_cov", C.POINTER(C.c_double)), ("sfnor_avg", C.POINTER(C.c_double)), ("map", C.POINTER(C.c_int)),
        ("cutsq", C.POINTER(C.c_double)), ("weights", C.POINTER(C.c_double)), ("bias", C.POINTER(C.c_double)),
    ]


_lib = None


def build():
    subprocess.run(["make", "-C", HERE, "restatement"], check=True, capture_output=True)


def lib():
    global _lib
    if _lib is None:
        if not os.path.isfile(SO):
            build()
        L = C.CDLL(SO)
        L.annp_oracle_compute.restype = C.c_int
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def compute(pot, cfg, ntypes=1, type_map=None, eflag=True, vflag=True, vatom=False, dump_G=False, nthreads=1):
    """PairANNP::compute restated on the CPU.  pot: AnnPotential-like (fields of the .ann file).

    Returns dict(eng_vdwl, eatom[nall], f[nall,3] unfolded, virial[6] (pair tally), vatom, G)."""
    L = lib()
    nall = cfg.nall
    keep = {}
    P = OracleParams()
    P.ntypes, P.nelements = ntypes, pot.nelements
    P.ntl, P.nhl, P.nnod, P.nsf, P.npsf, P.ntsf, P.flagsym = pot.ntl, pot.nhl, pot.nnod, pot.nsf, pot.npsf, pot.ntsf, pot.flagsym
    for i, a in enumerate(pot.flagact):
        P.flagact[i] = a
    P.cut, P.e_scale, P.e_shift, P.e_atom = pot.cut, pot.e_scale, pot.e_shift, pot.e_atom
    keep["cov"] = np.ascontiguousarray(pot.sfnor_cov, dtype=np.float64)
    keep["avg"] = np.ascontiguousarray(pot.sfnor_avg, dtype=np.float64)
    mp = np.zeros(ntypes + 1, dtype=np.int32) if type_map is None else np.ascontiguousarray(type_map, dtype=np.int32)
    keep["map"] = mp
    keep["cutsq"] = np.full((ntypes + 1) * (ntypes + 1), pot.cut * pot.cut)
    keep["w"] = np.ascontiguousarray(pot.weight_all, dtype=np.float64)
    keep["b"] = np.ascontiguousarray(pot.bias_all, dtype=np.float64)
    P.sfnor_cov, P.sfnor_avg, P.map = _dp(keep["cov"]), _dp(keep["avg"]), _ip(mp)
    P.cutsq, P.weights, P.bias = _dp(keep["cutsq"]), _dp(keep["w"]), _dp(keep["b"])

    x = np.ascontiguousarray(cfg.x, dtype=np.float64)
    typ = np.ascontiguousarray(cfg.type, dtype=np.int32)
    ilist = np.ascontiguousarray(cfg.ilist, dtype=np.int32)
    numneigh = np.ascontiguousarray(cfg.numneigh, dtype=np.int32)
    off = np.ascontiguousarray(cfg.offsets, dtype=np.int64)
    neigh = np.ascontiguousarray(cfg.neigh, dtype=np.int32)
    f = np.zeros((nall, 3))
    eng = C.c_double(0.0)
    eatom = np.zeros(nall) if eflag else None
    vir = np.zeros(6) if vflag else None
    va = np.zeros((nall, 6)) if vatom else None
    G = np.zeros((len(ilist), pot.nsf)) if dump_G else None
    rc = L.annp_oracle_compute(C.byref(P), C.c_int(cfg.nlocal), C.c_int(cfg.nghost), _dp(x), _ip(typ),
                               C.c_int(len(ilist)), _ip(ilist), _ip(numneigh), off.ctypes.data_as(C.POINTER(C.c_int64)),
                               _ip(neigh), _dp(f), C.byref(eng), _dp(eatom), _dp(vir), _dp(va), _dp(G), C.c_int(nthreads))
    if rc != 0:
        raise RuntimeError(f"annp_oracle_compute failed: {rc}")
    return {"eng_vdwl": eng.value, "eatom": eatom, "f": f, "virial": vir, "vatom": va, "G": G}

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
        ("sym_coerad", C.POINTER(C.c_double)), ("sym_coeang", C.POINTER(C.c_double)),
        ("nout", C.c_int), ("ngp", C.c_int), ("gparams", C.POINTER(C.c_double)), ("e_base", C.c_double),
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
        L.annp_oracle_compute_ni.restype = C.c_int
        L.anna_oracle_compute.restype = C.c_int
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _marshal(pot, cfg, ntypes, type_map, kind):
    keep = {}
    P = OracleParams()
    P.ntypes, P.nelements = ntypes, pot.nelements
    P.ntl, P.nhl, P.nnod, P.nsf, P.npsf, P.ntsf, P.flagsym = pot.ntl, pot.nhl, pot.nnod, pot.nsf, pot.npsf, pot.ntsf, pot.flagsym
    for i, a in enumerate(pot.flagact):
        P.flagact[i] = a
    P.cut = pot.cut
    if kind == "anna":
        P.nout, P.ngp, P.e_base = pot.nout, pot.ngp, pot.e_base
        keep["gp"] = np.ascontiguousarray(pot.gparams, dtype=np.float64)
        P.gparams = _dp(keep["gp"])
    else:
        P.e_scale, P.e_shift, P.e_atom = pot.e_scale, pot.e_shift, pot.e_atom
        keep["cov"] = np.ascontiguousarray(pot.sfnor_cov, dtype=np.float64)
        keep["avg"] = np.ascontiguousarray(pot.sfnor_avg, dtype=np.float64)
        P.sfnor_cov, P.sfnor_avg = _dp(keep["cov"]), _dp(keep["avg"])
    if kind == "ni":
        keep["rad"] = np.ascontiguousarray(pot.sym_coerad, dtype=np.float64)
        keep["ang"] = np.ascontiguousarray(pot.sym_coeang, dtype=np.float64)
        P.sym_coerad, P.sym_coeang = _dp(keep["rad"]), _dp(keep["ang"])
    mp = np.zeros(ntypes + 1, dtype=np.int32) if type_map is None else np.ascontiguousarray(type_map, dtype=np.int32)
    keep["map"] = mp
    keep["cutsq"] = np.full((ntypes + 1) * (ntypes + 1), pot.cut * pot.cut)
    keep["w"] = np.ascontiguousarray(pot.weight_all, dtype=np.float64)
    keep["b"] = np.ascontiguousarray(pot.bias_all, dtype=np.float64)
    P.map, P.cutsq, P.weights, P.bias = _ip(mp), _dp(keep["cutsq"]), _dp(keep["w"]), _dp(keep["b"])
    keep["x"] = np.ascontiguousarray(cfg.x, dtype=np.float64)
    keep["type"] = np.ascontiguousarray(cfg.type, dtype=np.int32)
    keep["ilist"] = np.ascontiguousarray(cfg.ilist, dtype=np.int32)
    keep["numneigh"] = np.ascontiguousarray(cfg.numneigh, dtype=np.int32)
    keep["off"] = np.ascontiguousarray(cfg.offsets, dtype=np.int64)
    keep["neigh"] = np.ascontiguousarray(cfg.neigh, dtype=np.int32)
    return P, keep


def _run(kind, pot, cfg, ntypes, type_map, eflag, vflag, vatom, dump_G, nthreads):
    L = lib()
    nall = cfg.nall
    P, k = _marshal(pot, cfg, ntypes, type_map, kind)
    f = np.zeros((nall, 3))
    eng = C.c_double(0.0)
    eatom = np.zeros(nall) if eflag else None
    vir = np.zeros(6) if vflag else None
    va = np.zeros((nall, 6)) if vatom else None
    G = np.zeros((len(k["ilist"]), pot.nsf)) if dump_G else None
    args = [C.byref(P), C.c_int(cfg.nlocal), C.c_int(cfg.nghost), _dp(k["x"]), _ip(k["type"]), C.c_int(len(k["ilist"])),
            _ip(k["ilist"]), _ip(k["numneigh"]), k["off"].ctypes.data_as(C.POINTER(C.c_int64)), _ip(k["neigh"]), _dp(f),
            C.byref(eng), _dp(eatom), _dp(vir), _dp(va), _dp(G)]
    out = {}
    if kind == "anna":
        lp = np.zeros((len(k["ilist"]), 2)) if dump_G else None
        rc = L.anna_oracle_compute(*args, _dp(lp), C.c_int(nthreads))
        out["lparams"] = lp
    elif kind == "ni":
        rc = L.annp_oracle_compute_ni(*args, C.c_int(nthreads))
    else:
        rc = L.annp_oracle_compute(*args, C.c_int(nthreads))
    if rc != 0:
        raise RuntimeError(f"oracle ({kind}) failed: {rc}")
    out.update({"eng_vdwl": eng.value, "eatom": eatom, "f": f, "virial": vir, "vatom": va, "G": G})
    return out


def compute(pot, cfg, ntypes=1, type_map=None, eflag=True, vflag=True, vatom=False, dump_G=False, nthreads=1):
    """PairANNP::compute (Fe / fe_v2 copy) restated on the CPU.  pot: AnnPotential-like (fields of the .ann file).

    Returns dict(eng_vdwl, eatom[nall], f[nall,3] unfolded, virial[6] (pair tally), vatom, G)."""
    return _run("fe", pot, cfg, ntypes, type_map, eflag, vflag, vatom, dump_G, nthreads)


def compute_ni(pot, cfg, ntypes=1, type_map=None, eflag=True, vflag=True, vatom=False, dump_G=False, nthreads=1):
    """First PairANNP::compute call of the Ni copy (annp_ni_oracle.c)."""
    return _run("ni", pot, cfg, ntypes, type_map, eflag, vflag, vatom, dump_G, nthreads)


def compute_anna(pot, cfg, ntypes=1, type_map=None, eflag=True, vflag=True, vatom=False, dump_G=False, nthreads=1):
    """PairANNA_ADP::compute (anna_adp_oracle.c).  pot: AnnaPotential-like."""
    return _run("anna", pot, cfg, ntypes, type_map, eflag, vflag, vatom, dump_G, nthreads)

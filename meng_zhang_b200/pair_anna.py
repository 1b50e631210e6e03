"""Host-side mirror of the reference ANNA-ADP pair style (`pair_style anna_adp/gpu`).

`PairANNAADPGPU` follows `PairANNAADPGPU` / `PairANNA_ADP` of the reference
(anna-gpu-lammps/bcc_fe/src/pair_anna_adp_gpu.{h,cpp}, pair_anna_adp.{h,cpp}): settings -> coeff -> init_style ->
compute, same argument meaning and error messages.  All arithmetic happens in libannp_b200.so; the numbers follow the
reference CPU style (newton on), see include/annp_b200.h.
"""
from __future__ import annotations

import ctypes as C
import dataclasses

import numpy as np

from . import capi
from .capi import AnnpError
from .pair import LammpsError, PairANNPGPU, _dp, _ip


@dataclasses.dataclass
class AnnaPotential:
    """Contents of a `.anna` file as PairANNA_ADP::read_file stores them (pair_anna_adp.cpp:392-637)."""
    nelements: int
    ntl: int
    nhl: int
    nnod: int
    nout: int
    nsf: int
    npsf: int
    ntsf: int
    flagsym: int
    flagact: list
    cut: float
    e_base: float
    e_scal: float
    ngp: int
    gparams: np.ndarray        # [ngp] A0, yy, gamma, C0, c1F, c2F, V0, b1, b2, delta, r0, r1, hc, d1, q1, d3, q3
    id_elem: list
    mass: list
    elements: list
    weight_all: np.ndarray     # [nelements][ntl-1][nnod][nsf]
    bias_all: np.ndarray       # [nelements][ntl-1][nnod]

    def flat_weights(self):
        """Flatten like pair_anna_adp_gpu.cpp init_style: per layer row-major, layers concatenated; last layer nout rows."""
        nl = self.ntl - 1
        w, b = [], []
        for e in range(self.nelements):
            for l in range(nl):
                nrow = self.nout if l == nl - 1 else self.nnod
                ncol = self.nsf if l == 0 else self.nnod
                w.append(self.weight_all[e, l, :nrow, :ncol].reshape(-1))
                b.append(self.bias_all[e, l, :nrow])
        return np.ascontiguousarray(np.concatenate(w)), np.ascontiguousarray(np.concatenate(b))


def read_anna_potential(filename: str, elements_coeff=("Fe",)) -> AnnaPotential:
    L = capi.lib()
    pot = capi.AnnaPotential()
    err = C.create_string_buffer(256)
    names = (C.c_char_p * len(elements_coeff))(*[e.encode() for e in elements_coeff])
    rc = L.anna_b200_read_potential(filename.encode(), len(elements_coeff), names, C.byref(pot), err, 256)
    if rc != 0:
        raise LammpsError(err.value.decode())
    try:
        nl = pot.ntl - 1
        w = np.ctypeslib.as_array(pot.weight_all, shape=(pot.nelements, nl, pot.nnod, pot.nsf)).copy()
        b = np.ctypeslib.as_array(pot.bias_all, shape=(pot.nelements, nl, pot.nnod)).copy()
        return AnnaPotential(
            nelements=pot.nelements, ntl=pot.ntl, nhl=pot.nhl, nnod=pot.nnod, nout=pot.nout, nsf=pot.nsf, npsf=pot.npsf,
            ntsf=pot.ntsf, flagsym=pot.flagsym, flagact=[pot.flagact[i] for i in range(nl)], cut=pot.cut,
            e_base=pot.e_base, e_scal=pot.e_scal, ngp=pot.ngp, gparams=np.array(pot.gparams[:pot.ngp]),
            id_elem=[pot.id_elem[i] for i in range(pot.nelements)], mass=[pot.mass[i] for i in range(pot.nelements)],
            elements=[pot.elements[i].value.decode() for i in range(pot.nelements)], weight_all=w, bias_all=b)
    finally:
        L.anna_b200_free_potential(C.byref(pot))


def write_anna_potential(path: str, pot: AnnaPotential, comment: str = "written by meng_zhang_b200") -> None:
    """Write the `.anna` text format (CRLF, tab separated, fixed line positions)."""
    nl = pot.ntl - 1
    act_names = {0: "linear", 1: "hyp", 2: "sig", 3: "mod", 4: "tanh"}
    sym = {0: "Chebyshev", 1: "Behler", 2: "Customized"}[pot.flagsym]
    fmt = lambda v: repr(float(v))
    # e_scal is only picked up after TAB + digit (pair_anna_adp.cpp:466-471): keep it non-negative and digit-led
    lines = [f"#Sourse: {comment}", "#Date: -", "#contact information: -", "",
             "#element parameters_(nelement #n element mass)", str(pot.nelements)]
    for e in range(pot.nelements):
        lines.append(f"{pot.id_elem[e]}\t{pot.elements[e]}\t{fmt(pot.mass[e])}")
    lines += ["", "#artificial neural network parameters_(TL HL Nodes_HL Num_out Num_SF Num_PSF Num_TSF Cut) ",
              f"{pot.ntl}\t{pot.nhl}\t{pot.nnod}\t{pot.nout}\t{pot.nsf}\t{pot.npsf}\t{pot.ntsf}\t{fmt(pot.cut)} ", "",
              "#types of symmetry function and activation function",
              "\t".join([sym] + [act_names[a] for a in pot.flagact]), "",
              "#energy base_(e_base e_scale)", f"{fmt(pot.e_base)}\t{fmt(pot.e_scal)}", "",
              "#adp parameters (A0, yy, gamma, C0, c1F, c2F, V0, b1, b2, delta, r0, r1, hc, d1, q1, d3, q3)",
              str(pot.ngp), "\t".join(fmt(v) for v in pot.gparams), "",
              "#weight_bias_matrix_(#1.....#TL)"]
    for e in range(pot.nelements):
        lines.append(f"#{pot.elements[e]}")
        for l in range(nl):
            nrow = pot.nout if l == nl - 1 else pot.nnod
            ncol = pot.nsf if l == 0 else pot.nnod
            lines.append(f"#{l + 1}_(weight)")
            for r in range(nrow):
                lines.append("\t".join(fmt(v) for v in pot.weight_all[e, l, r, :ncol]))
            lines.append(f"#{l + 1}_(bias)")
            lines.append("\t".join(fmt(v) for v in pot.bias_all[e, l, :nrow]))
            lines.append("")
    with open(path, "w", newline="") as fp:
        fp.write("\r\n".join(lines) + "\r\n")


class PairANNAADPGPU(PairANNPGPU):
    """`pair_style anna_adp/gpu` served by libannp_b200.so (same call sequence as PairANNPGPU)."""

    def coeff(self, args):
        n = self.ntypes
        if not self.allocated:
            self.setflag = np.zeros((n + 1, n + 1), dtype=np.int32)
            self.cutsq = np.zeros((n + 1, n + 1))
            self.map = np.full(n + 1, -1, dtype=np.int32)
            self.allocated = True
        if len(args) != 3 + n or args[0] != "*" or args[1] != "*":
            raise LammpsError("Incorrect args for pair coefficients")              # pair_anna_adp.cpp:323-326
        elements = []
        for i in range(3, len(args)):
            if args[i] == "":
                continue
            if args[i] in elements:
                j = elements.index(args[i])
            else:
                j = len(elements)
                elements.append(args[i])
            self.map[i - 2] = j
        self.elements_coeff = elements
        self.params = read_anna_potential(args[2], elements)
        if len(elements) != self.params.nelements:
            raise LammpsError("Incorrect args for pair coefficients")
        self.cutmax = max(0.0, self.params.cut)
        count = 0
        for i in range(1, n + 1):
            for j in range(i, n + 1):
                if self.map[i] >= 0 and self.map[j] >= 0:
                    self.setflag[i, j] = 1
                    count += 1
        if count == 0:
            raise LammpsError("Incorrect args for pair coefficients")

    def init_style(self):
        # the reference CPU style needs newton on (pair_anna_adp.cpp:371-372); its GPU style demands newton off because of
        # its two-phase layout (pair_anna_adp_gpu.cpp:166-167) - this implementation follows the CPU style
        if self.newton_pair == 0:
            raise LammpsError("Pair style Neural Network Potential requires newton pair on")
        if self.params is None:
            raise LammpsError("All pair coeffs are not set")
        p, n = self.params, self.ntypes
        for i in range(1, n + 1):
            for j in range(i, n + 1):
                if self.setflag[i, j] != 0 or (self.setflag[i, i] != 0 and self.setflag[j, j] != 0):
                    self.cutsq[i, j] = self.cutsq[j, i] = self.cutmax * self.cutmax
                else:
                    self.cutsq[i, j] = self.cutsq[j, i] = 0.0
        w, b = p.flat_weights()
        cutsq = np.ascontiguousarray(self.cutsq.reshape(-1))
        mp = np.ascontiguousarray(np.where(self.map < 0, 0, self.map).astype(np.int32))
        gp = np.ascontiguousarray(p.gparams, dtype=np.float64)
        P = capi.AnnaParams()
        P.abi_version = capi.ABI_VERSION
        P.ntypes, P.nelements = n, p.nelements
        P.ntl, P.nhl, P.nnod, P.nout, P.nsf, P.npsf, P.ntsf, P.ngp = p.ntl, p.nhl, p.nnod, p.nout, p.nsf, p.npsf, p.ntsf, p.ngp
        P.flagsym = p.flagsym
        for i, a in enumerate(p.flagact):
            P.flagact[i] = a
        P.e_base, P.cut = p.e_base, p.cut
        P.cutsq, P.map, P.weights, P.bias, P.gparams = _dp(cutsq), _ip(mp), _dp(w), _dp(b), _dp(gp)
        self.clear()
        err = C.create_string_buffer(512)
        h = C.c_void_p(None)
        rc = capi.lib().anna_b200_init(C.byref(P), self.device, C.byref(h), err, 512)
        if rc != 0:
            raise AnnpError(rc, err.value.decode())
        self.handle = h
        self.cell_size = self.cutmax + self.skin

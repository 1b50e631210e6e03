"""Host-side mirror of the reference pair-style interface for the ANNP hot path.

`PairANNPGPU` follows `PairANNPGPU` / `PairANNP` of the reference
(annp-gpu-lammps/fe_v2/src/pair_annp_gpu.{h,cpp}, pair_annp.{h,cpp}): same call sequence
(settings -> coeff -> init_style -> compute), same argument meaning and the same error messages,
so parity tests read like LAMMPS input decks.  All arithmetic happens in libannp_b200.so (CUDA,
sm_100a); this file only marshals arrays.  The C++ twin that LAMMPS itself would compile is
meng_zhang_b200/lammps/pair_annp_b200.{h,cpp}.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import math

import numpy as np

from . import capi
from .capi import AnnpError


class LammpsError(RuntimeError):
    """error->all / error->one of the reference."""


@dataclasses.dataclass
class AnnPotential:
    """Contents of a `.ann` file as PairANNP::read_file stores them (pair_annp.cpp:332-518)."""
    nelements: int
    ntl: int
    nhl: int
    nnod: int
    nsf: int
    npsf: int
    ntsf: int
    flagsym: int
    flagact: list
    cut: float
    e_scale: float
    e_shift: float
    e_atom: float
    id_elem: list
    mass: list
    elements: list
    sfnor_cov: np.ndarray      # [nsf]
    sfnor_avg: np.ndarray      # [nsf]
    weight_all: np.ndarray     # [nelements][ntl-1][nnod][nsf]  rows padded with zeros
    bias_all: np.ndarray       # [nelements][ntl-1][nnod]
    # Ni copy only (ni/src/pair_annp.cpp:510-545): sfnor_cov / sfnor_avg then hold the sf_min / sf_max rows
    sym_coerad: np.ndarray | None = None   # [npsf][3] eta, rs, Rc
    sym_coeang: np.ndarray | None = None   # [ntsf][4] eta, lambda, zeta, Rc

    @property
    def variant(self) -> int:
        """The file itself does not name the copy it is for; only Ni files carry the coefficient blocks."""
        return capi.VARIANT_NI if self.sym_coerad is not None else capi.VARIANT_FE

    def sf_scale(self) -> np.ndarray:
        """s_n = 1/sqrt(cov - avg^2), 0 if <= 1e-10 (pair_annp.cpp:98-108, pair_annp_gpu.cpp:211-220)."""
        out = np.zeros(self.nsf)
        for i in range(self.nsf):
            t = math.sqrt(self.sfnor_cov[i] - self.sfnor_avg[i] * self.sfnor_avg[i]) \
                if self.sfnor_cov[i] - self.sfnor_avg[i] * self.sfnor_avg[i] >= 0 else float("nan")
            out[i] = 0.0 if not (t > 1.0e-10) else 1.0 / t
        return out

    def flat_weights(self):
        """Flatten like pair_annp_gpu.cpp:190-209: per layer row-major [row*ncol+col], layers concatenated."""
        nl = self.ntl - 1
        w, b = [], []
        for e in range(self.nelements):
            for l in range(nl):
                nrow = 1 if l == nl - 1 else self.nnod
                ncol = self.nsf if l == 0 else self.nnod
                w.append(self.weight_all[e, l, :nrow, :ncol].reshape(-1))
                b.append(self.bias_all[e, l, :nrow])
        return np.ascontiguousarray(np.concatenate(w)), np.ascontiguousarray(np.concatenate(b))


def read_potential(filename: str, elements_coeff=("Fe",)) -> AnnPotential:
    """Parse a `.ann` file with the library's reader (csrc/annp_potential.cpp)."""
    L = capi.lib()
    pot = capi.Potential()
    err = C.create_string_buffer(256)
    names = (C.c_char_p * len(elements_coeff))(*[e.encode() for e in elements_coeff])
    rc = L.annp_b200_read_potential(filename.encode(), len(elements_coeff), names, C.byref(pot), err, 256)
    if rc != 0:
        raise LammpsError(err.value.decode())
    try:
        nl = pot.ntl - 1
        w = np.ctypeslib.as_array(pot.weight_all, shape=(pot.nelements, nl, pot.nnod, pot.nsf)).copy()
        b = np.ctypeslib.as_array(pot.bias_all, shape=(pot.nelements, nl, pot.nnod)).copy()
        return AnnPotential(
            nelements=pot.nelements, ntl=pot.ntl, nhl=pot.nhl, nnod=pot.nnod, nsf=pot.nsf, npsf=pot.npsf, ntsf=pot.ntsf,
            flagsym=pot.flagsym, flagact=[pot.flagact[i] for i in range(nl)], cut=pot.cut,
            e_scale=pot.e_scale, e_shift=pot.e_shift, e_atom=pot.e_atom,
            id_elem=[pot.id_elem[i] for i in range(pot.nelements)], mass=[pot.mass[i] for i in range(pot.nelements)],
            elements=[pot.elements[i].value.decode() for i in range(pot.nelements)],
            sfnor_cov=np.array(pot.sfnor_cov[:pot.nsf]), sfnor_avg=np.array(pot.sfnor_avg[:pot.nsf]),
            weight_all=w, bias_all=b,
            sym_coerad=np.array([list(pot.sym_coerad[i]) for i in range(pot.npsf)]) if pot.has_sym_coeff else None,
            sym_coeang=np.array([list(pot.sym_coeang[i]) for i in range(pot.ntsf)]) if pot.has_sym_coeff else None)
    finally:
        L.annp_b200_free_potential(C.byref(pot))


def write_potential(path: str, pot: AnnPotential, comment: str = "written by meng_zhang_b200") -> None:
    """Write the `.ann` text format (CRLF, tab separated, fixed line positions) read by read_file."""
    nl = pot.ntl - 1
    act_names = {0: "linear", 1: "hyp", 2: "sig", 3: "mod", 4: "tanh"}   # only the scanned 2-char keys
    sym = {0: "Chebyshev", 1: "Behler", 2: "Customized"}[pot.flagsym]
    fmt = lambda v: repr(float(v))
    ni = pot.sym_coerad is not None
    lines = [f"#Sourse: {comment}", "#Date: -", "#contact information: -", "",
             "#element parameters_(nelement #n element mass)", str(pot.nelements)]
    for e in range(pot.nelements):
        lines.append(f"{pot.id_elem[e]}\t{pot.elements[e]}\t{fmt(pot.mass[e])}")
    lines += ["", "#artificial neural network parameters_(TL HL Nodes_HL Num_SF Num_PSF Num_TSF Cut) ",
              f"{pot.ntl}\t{pot.nhl}\t{pot.nnod}\t{pot.nsf}\t{pot.npsf}\t{pot.ntsf}\t{fmt(pot.cut)} ", "",
              "#symmetry function normization_(sf_min sf_max)" if ni else "#symmetry function normization_(sfval_cov sfval_avg)",
              "\t".join(fmt(v) for v in pot.sfnor_cov), "\t".join(fmt(v) for v in pot.sfnor_avg), "",
              "#types of symmetry function and activation function",
              "\t".join([sym] + [act_names[a] for a in pot.flagact]), "",
              "#energy scale_(E_scale E_shift E_atom)", fmt(pot.e_scale), fmt(pot.e_shift), fmt(pot.e_atom), "",
              "#weight_bias_matrix_(#1.....#TL)"]
    for e in range(pot.nelements):
        for l in range(nl):
            nrow = 1 if l == nl - 1 else pot.nnod
            ncol = pot.nsf if l == 0 else pot.nnod
            lines.append(f"#{pot.elements[e]}")
            lines.append(f"#{l + 1}_(weight)")
            for r in range(nrow):
                lines.append("\t".join(fmt(v) for v in pot.weight_all[e, l, r, :ncol]))
            lines.append(f"#{l + 1}_(bias)")
            lines.append("\t".join(fmt(v) for v in pot.bias_all[e, l, :nrow]))
            lines.append("")
    if ni:
        # trailing blocks of the Ni files; values are only picked up after TAB + digit/'-', the element names are not
        lines += ["#coefficent of symmetry funciton", f"#rad\t{pot.npsf}"]
        for row in pot.sym_coerad:
            lines.append(pot.elements[0] + "\t" + "\t".join(fmt(v) for v in row))
        lines.append(f"#angl\t{pot.ntsf}")
        for row in pot.sym_coeang:
            lines.append(pot.elements[0] + "\t" + pot.elements[0] + "\t" + "\t".join(fmt(v) for v in row))
    with open(path, "w", newline="") as fp:
        fp.write("\r\n".join(lines) + "\r\n")


def _dp(a: np.ndarray):
    return a.ctypes.data_as(capi.c_double_p)


def _ip(a: np.ndarray):
    return a.ctypes.data_as(capi.c_int_p)


class PairANNPGPU:
    """`pair_style annp/gpu` served by libannp_b200.so.

    Usage (same order as a LAMMPS deck):
        pair = PairANNPGPU(ntypes=1)
        pair.settings([])                                   # pair_style annp/gpu
        pair.coeff(["*", "*", "fe_annp_potential_2.ann", "Fe"])
        pair.init_style()
        f = pair.compute(eflag, vflag, cfg, ago=0)          # cfg: lattice.Config
    After compute: pair.eng_vdwl, pair.eatom, pair.virial (6), pair.vatom as in LAMMPS' Pair.
    """

    def __init__(self, ntypes: int = 1, device: int = -1, newton_pair: int = 1, skin: float = 2.0, variant: int | None = None):
        """variant: which copy of the reference style to follow (capi.VARIANT_FE / VARIANT_NI); None = decide from the
        potential file (Ni files carry symmetry-function coefficient blocks, Fe files do not)."""
        self.variant = variant
        self.ntypes = ntypes
        self.device = device
        self.newton_pair = newton_pair
        self.skin = skin
        self.allocated = False
        self.params: AnnPotential | None = None
        self.map = None
        self.setflag = None
        self.cutsq = None
        self.cutmax = 0.0
        self.handle = C.c_void_p(None)
        self.eng_vdwl = 0.0
        self.eatom = None
        self.virial = np.zeros(6)
        self.vatom = None
        self._keep = None

    # ---- PairANNP::settings (pair_annp.cpp:249-252)
    def settings(self, args):
        if len(args) != 0:
            raise LammpsError("Illegal pair_style command")

    # ---- PairANNP::coeff (pair_annp.cpp:257-304)
    def coeff(self, args):
        n = self.ntypes
        if not self.allocated:
            self.setflag = np.zeros((n + 1, n + 1), dtype=np.int32)
            self.cutsq = np.zeros((n + 1, n + 1))
            self.map = np.full(n + 1, -1, dtype=np.int32)
            self.allocated = True
        if len(args) != 3 + n:
            raise LammpsError("Incorrect args for pair coefficients")
        if args[0] != "*" or args[1] != "*":
            raise LammpsError("Incorrect args for pair coefficients")
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
        try:
            self.params = read_potential(args[2], elements)
        except LammpsError:
            raise
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

    # ---- PairANNP::init_one (pair_annp.cpp:323-327)
    def init_one(self, i, j):
        if self.setflag[i, j] == 0:
            raise LammpsError("All pair coeffs are not set")
        return self.cutmax

    # ---- PairANNPGPU::init_style (pair_annp_gpu.cpp:136-243)
    def init_style(self):
        if self.newton_pair == 0:
            raise LammpsError("Pair style annp/gpu requires newton pair on")
        if self.params is None:
            raise LammpsError("All pair coeffs are not set")
        p, n = self.params, self.ntypes
        for i in range(1, n + 1):
            for j in range(i, n + 1):
                if self.setflag[i, j] != 0 or (self.setflag[i, i] != 0 and self.setflag[j, j] != 0):
                    cut = self.cutmax
                    self.cutsq[i, j] = self.cutsq[j, i] = cut * cut
                else:
                    self.cutsq[i, j] = self.cutsq[j, i] = 0.0
        w, b = p.flat_weights()
        variant = p.variant if self.variant is None else self.variant
        if (variant & 0xff) == capi.VARIANT_NI:
            if p.sym_coerad is None:
                raise LammpsError("potential file has no symmetry-function coefficient blocks (needed by the Ni copy)")
            # (G - sf_min) / (sf_max - sf_min), ni/src/pair_annp.cpp:99-101,168-170
            scal = np.ascontiguousarray(1.0 / (p.sfnor_avg - p.sfnor_cov))
            avg = np.ascontiguousarray(p.sfnor_cov, dtype=np.float64)
            corad = np.ascontiguousarray(p.sym_coerad, dtype=np.float64)
            coang = np.ascontiguousarray(p.sym_coeang, dtype=np.float64)
        else:
            scal = np.ascontiguousarray(p.sf_scale())
            avg = np.ascontiguousarray(p.sfnor_avg, dtype=np.float64)
            corad = coang = None
        cutsq = np.ascontiguousarray(self.cutsq.reshape(-1))
        mp = np.ascontiguousarray(np.where(self.map < 0, 0, self.map).astype(np.int32))
        P = capi.Params()
        P.abi_version = capi.ABI_VERSION
        P.ntypes, P.nelements = n, p.nelements
        P.ntl, P.nhl, P.nnod, P.nsf, P.npsf, P.ntsf = p.ntl, p.nhl, p.nnod, p.nsf, p.npsf, p.ntsf
        P.flagsym = p.flagsym
        for i, a in enumerate(p.flagact):
            P.flagact[i] = a
        P.e_scale, P.e_shift, P.e_atom, P.cut = p.e_scale, p.e_shift, p.e_atom, p.cut
        P.sfnor_scal, P.sfnor_avg, P.cutsq, P.map = _dp(scal), _dp(avg), _dp(cutsq), _ip(mp)
        P.weights, P.bias = _dp(w), _dp(b)
        P.variant = variant
        if corad is not None:
            P.sym_coerad, P.sym_coeang = _dp(corad), _dp(coang)
        self.clear()
        err = C.create_string_buffer(512)
        h = C.c_void_p(None)
        rc = capi.lib().annp_b200_init(C.byref(P), self.device, 0, 0, C.byref(h), err, 512)
        if rc != 0:
            # GPU_EXTRA::check_flag (pair_annp_gpu.cpp:235) aborts all ranks on a non-zero init code
            raise AnnpError(rc, err.value.decode())
        self.handle = h
        self.cell_size = self.cutmax + self.skin           # pair_annp_gpu.cpp:222

    def interaction_cutoff(self) -> float:
        """Largest distance at which a neighbour can contribute.  Fe copies / ANNA: the pair cutoff.  Ni copy: the
        descriptor cutoff Rc[Bohr]/1.889726 = 3.9 A, well inside the 6.5 A list cutoff LAMMPS derives from the file's
        `Cut` field - a driver that builds its own list (md.py) can use this tighter radius; entries beyond it are
        filtered by the kernel anyway and the order of the remaining ones is unchanged, so results are identical."""
        p = self.params
        if getattr(p, "sym_coerad", None) is not None and (self.variant is None or (self.variant & 0xff) == capi.VARIANT_NI):
            return float(max(p.sym_coerad[0][2], p.sym_coeang[0][3]) / 1.889726)
        return float(self.cutmax)

    def _check(self, rc):
        if rc != 0:
            msg = capi.lib().annp_b200_last_error(self.handle).decode()
            if rc == capi.ENOMEM:
                raise LammpsError("Insufficient memory on accelerator")   # pair_annp_gpu.cpp:125-126
            raise AnnpError(rc, msg)

    def upload_neighbors(self, cfg):
        """neighbor->ago == 0: hand the full list to the device (ANNP::reset_nbors)."""
        off = np.ascontiguousarray(cfg.offsets, dtype=np.int64)
        ilist = np.ascontiguousarray(cfg.ilist, dtype=np.int32)
        neigh = np.ascontiguousarray(cfg.neigh, dtype=np.int32)
        self._check(capi.lib().annp_b200_neigh_csr(self.handle, len(ilist), cfg.nall, _ip(ilist),
                                                   off.ctypes.data_as(capi.c_int64_p), _ip(neigh)))

    # ---- PairANNPGPU::compute (pair_annp_gpu.cpp:84-131)
    def compute(self, eflag, vflag, cfg, ago=0, x=None):
        """eflag/vflag use LAMMPS bits: eflag 1 global | 2 per-atom; vflag 1|2 global, 4 per-atom.
        Returns f[nall,3] (ghost rows = what reverse_comm would send home)."""
        if not self.handle:
            raise LammpsError("init_style was not called")
        if ago == 0:
            self.upload_neighbors(cfg)
        nall = cfg.nall
        xx = np.ascontiguousarray(cfg.x if x is None else x, dtype=np.float64)
        typ = np.ascontiguousarray(cfg.type, dtype=np.int32)
        f = np.zeros((nall, 3))
        eng = C.c_double(0.0)
        eflag_atom, vflag_atom = bool(eflag & 2), bool(vflag & 4)
        vflag_global = bool(vflag & 3)
        eatom = np.zeros(nall) if eflag_atom else None
        vatom = np.zeros((nall, 6)) if vflag_atom else None
        vir = np.zeros(6)
        rc = capi.lib().annp_b200_compute(
            self.handle, cfg.nlocal, cfg.nghost, _dp(xx), _ip(typ), int(eflag != 0), int(vflag != 0),
            _dp(f), C.byref(eng) if eflag else None, _dp(eatom) if eflag_atom else None,
            _dp(vir) if vflag_global else None, _dp(vatom) if vflag_atom else None)
        self._check(rc)
        if eflag:
            self.eng_vdwl = eng.value
        self.eatom, self.vatom = eatom, vatom
        if vflag_global:
            self.virial = vir
        return f

    def descriptors(self, cfg):
        """Centred descriptors G[inum,nsf] and dOut/dG of the last compute (debug hook)."""
        L = capi.lib()
        self._check(L.annp_b200_debug_descriptors(self.handle, None, None))   # arm
        self.compute(1, 0, cfg, ago=1)
        nsf = self.params.nsf
        G = np.zeros((len(cfg.ilist), nsf))
        dE = np.zeros((len(cfg.ilist), nsf))
        self._check(L.annp_b200_debug_descriptors(self.handle, _dp(G), _dp(dE)))
        return G, dE

    def set_scatter(self, mode: int) -> None:
        """capi.SCATTER_FIXED (default for the Chebyshev / ANNA-ADP styles) or capi.SCATTER_GATHER, see include/annp_b200.h."""
        self._check(capi.lib().annp_b200_set_scatter(self.handle, int(mode)))

    def stats(self):
        st = capi.Stats()
        self._check(capi.lib().annp_b200_get_stats(self.handle, C.byref(st)))
        return st

    def memory_usage(self):
        return capi.lib().annp_b200_bytes(self.handle)

    def clear(self):
        if self.handle:
            capi.lib().annp_b200_clear(self.handle)
            self.handle = C.c_void_p(None)

    def __del__(self):
        try:
            self.clear()
        except Exception:
            pass

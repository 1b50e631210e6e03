"""Shared helpers for the test-suite: golden fixtures -> Config / potential file."""
from __future__ import annotations

import json
import os

import numpy as np

from meng_zhang_b200 import lattice as L
from meng_zhang_b200.pair import AnnPotential, write_potential
from meng_zhang_b200.pair_anna import AnnaPotential, write_anna_potential

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FE_CASES = ["bcc4_perfect", "bcc4_perturbed", "bcc334_hot", "cluster_ragged", "bcc4_two_types", "bcc10_perturbed"]
NI_CASES = ["fcc3_perturbed", "fcc334_hot", "cluster_ragged", "fcc3_two_types"]
ANNA_CASES = ["bcc4_perfect", "bcc4_perturbed", "bcc334_hot", "cluster_ragged", "bcc4_two_types", "bcc8_perturbed"]


# the shipped potentials live in the package (meng_zhang_b200/data): bench.py and smoke() use them too
from meng_zhang_b200.potentials import (load_anna_potential_json, load_potential_json, write_anna_fe_potential,  # noqa: E402,F401
                                        write_fe_potential, write_ni_potential)


def load_case(name, prefix="annp_fe"):
    z = np.load(os.path.join(GOLDEN, f"{prefix}_{name}.npz"))
    cfg = L.Config(nlocal=int(z["nlocal"]), nghost=int(z["nghost"]), x=z["x"], type=z["type"],
                   ghost_owner=z["ghost_owner"], ilist=z["ilist"], numneigh=z["numneigh"], neigh=z["neigh"], box=z["box"])
    ref = {k: z[k] for k in ("eng_vdwl", "eatom", "f", "virial_pair", "virial_fdotr", "vatom")}
    ref["eng_vdwl"] = float(ref["eng_vdwl"])
    return cfg, [str(e) for e in z["elements"]], ref


# ---- general potentials: shapes, activations and element counts the shipped files do not exercise ----
GENERAL_CASES = ["act_hyp_sig", "act_mod_deep", "shape_8_20", "shape_6_11", "shape_12_22", "shape_16_24", "two_elements"]


def general_potential(name) -> tuple:
    """(AnnPotential, elements on the pair_coeff line, cutoff) of a synthetic potential.  Only used by
    tests/golden/make_golden.py; the tests read the numbers back from the golden file (load_general_case)."""
    fe = load_potential_json()
    rng = np.random.default_rng(sum(ord(c) * (i + 1) for i, c in enumerate(name)))

    def net(nelements, nl, nnod, nsf, gain=1.0):
        w = np.zeros((nelements, nl, nnod, nsf))
        b = np.zeros((nelements, nl, nnod))
        for e in range(nelements):
            for l in range(nl):
                nrow = 1 if l == nl - 1 else nnod
                ncol = nsf if l == 0 else nnod
                w[e, l, :nrow, :ncol] = rng.normal(0.0, gain / np.sqrt(ncol), size=(nrow, ncol))
                b[e, l, :nrow] = rng.normal(0.0, 0.1, size=nrow)
        return w, b

    def norm(npsf, ntsf):
        """normalisation rows of the Fe file for the same Chebyshev orders (last entry repeated beyond the file's)"""
        idx = [min(m, fe.npsf - 1) for m in range(npsf)] + [fe.npsf + min(n, fe.ntsf - 1) for n in range(ntsf)]
        return fe.sfnor_cov[idx].copy(), fe.sfnor_avg[idx].copy()

    def pot(npsf, ntsf, nnod, flagact, cut, nelements=1, elements=("Fe",), w=None, b=None):
        nl = len(flagact)
        cov, avg = norm(npsf, ntsf)
        if w is None:
            w, b = net(nelements, nl, nnod, npsf + ntsf)
        return AnnPotential(nelements=nelements, ntl=nl + 1, nhl=nl - 1, nnod=nnod, nsf=npsf + ntsf, npsf=npsf, ntsf=ntsf, flagsym=0,
                            flagact=list(flagact), cut=cut, e_scale=fe.e_scale, e_shift=fe.e_shift, e_atom=fe.e_atom,
                            id_elem=list(range(1, nelements + 1)), mass=[55.845, 51.9961][:nelements], elements=list(elements),
                            sfnor_cov=cov, sfnor_avg=avg, weight_all=w, bias_all=b)

    if name == "act_hyp_sig":        # the Fe file's own network with activations 1 (tanh) and 2 (the reference's 1/(1+exp(+x)))
        return pot(9, 19, 10, [1, 2, 0], 6.5, w=fe.weight_all.copy(), b=fe.bias_all.copy()), ["Fe"]
    if name == "act_mod_deep":       # five layers, activation 3 (1.7159 tanh(2x/3)), 2, 1 and a NON-linear output layer (4)
        return pot(9, 19, 8, [3, 2, 1, 4], 6.5), ["Fe"]
    if name == "shape_8_20":         # exact instantiation <8,20>
        return pot(8, 20, 10, [4, 4, 0], 6.5), ["Fe"]
    if name == "shape_6_11":         # padded to <8,24>
        return pot(6, 11, 7, [4, 4, 0], 5.9), ["Fe"]
    if name == "shape_12_22":        # nsf = 34 > 32: padded to <16,24>, two-round descriptor reduction
        return pot(12, 22, 12, [4, 1, 0], 6.5), ["Fe"]
    if name == "shape_16_24":        # the largest supported shape, 32 nodes per layer
        return pot(16, 24, 32, [1, 4, 0], 6.2), ["Fe"]
    if name == "two_elements":       # two element blocks: the reference's reader files every block under element 0 (the
        return pot(9, 19, 10, [4, 4, 0], 6.5, nelements=2, elements=("Fe", "Cr")), ["Fe", "Cr"]   # last one wins), element 1 stays zero
    raise KeyError(name)


def load_general_case(name):
    """Config, pair_coeff elements, the potential AS THE REFERENCE READ IT is irrelevant here: the file is re-created from
    the stored numbers with the writer and handed to both sides.  Returns (cfg, elems, ref, AnnPotential)."""
    z = np.load(os.path.join(GOLDEN, f"annp_general_{name}.npz"))
    cfg = L.Config(nlocal=int(z["nlocal"]), nghost=int(z["nghost"]), x=z["x"], type=z["type"],
                   ghost_owner=z["ghost_owner"], ilist=z["ilist"], numneigh=z["numneigh"], neigh=z["neigh"], box=z["box"])
    ref = {k: z[k] for k in ("eng_vdwl", "eatom", "f", "virial_pair", "virial_fdotr", "vatom")}
    ref["eng_vdwl"] = float(ref["eng_vdwl"])
    nel = int(z["pot_nelements"])
    pot = AnnPotential(nelements=nel, ntl=int(z["pot_ntl"]), nhl=int(z["pot_ntl"]) - 2, nnod=int(z["pot_nnod"]),
                       nsf=int(z["pot_npsf"]) + int(z["pot_ntsf"]), npsf=int(z["pot_npsf"]), ntsf=int(z["pot_ntsf"]), flagsym=0,
                       flagact=[int(a) for a in z["pot_flagact"]], cut=float(z["pot_cut"]), e_scale=float(z["pot_e_scale"]),
                       e_shift=float(z["pot_e_shift"]), e_atom=float(z["pot_e_atom"]), id_elem=list(range(1, nel + 1)),
                       mass=[55.845, 51.9961][:nel], elements=[str(e) for e in z["pot_elements"]], sfnor_cov=z["pot_sfnor_cov"],
                       sfnor_avg=z["pot_sfnor_avg"], weight_all=z["pot_weight_all"], bias_all=z["pot_bias_all"])
    return cfg, [str(e) for e in z["elements"]], ref, pot

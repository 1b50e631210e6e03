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


def load_potential_json(name="fe_potential.json") -> AnnPotential:
    with open(os.path.join(GOLDEN, name)) as fp:
        d = json.load(fp)
    d.pop("source", None)
    for k in ("sfnor_cov", "sfnor_avg", "weight_all", "bias_all", "sym_coerad", "sym_coeang"):
        if k in d:
            d[k] = np.array(d[k])
    return AnnPotential(**d)


def load_anna_potential_json(name="anna_potential.json") -> AnnaPotential:
    with open(os.path.join(GOLDEN, name)) as fp:
        d = json.load(fp)
    d.pop("source", None)
    for k in ("gparams", "weight_all", "bias_all"):
        d[k] = np.array(d[k])
    return AnnaPotential(**d)


def write_ni_potential(path) -> str:
    write_potential(str(path), load_potential_json("ni_potential.json"),
                    comment="ANN potential for Ni re-written from tests/golden/ni_potential.json")
    return str(path)


def write_anna_fe_potential(path) -> str:
    write_anna_potential(str(path), load_anna_potential_json(),
                         comment="ANNA-ADP potential for Fe re-written from tests/golden/anna_potential.json")
    return str(path)


def write_fe_potential(path) -> str:
    write_potential(str(path), load_potential_json(), comment="ANN potential for Fe re-written from tests/golden/fe_potential.json")
    return str(path)


def load_case(name, prefix="annp_fe"):
    z = np.load(os.path.join(GOLDEN, f"{prefix}_{name}.npz"))
    cfg = L.Config(nlocal=int(z["nlocal"]), nghost=int(z["nghost"]), x=z["x"], type=z["type"],
                   ghost_owner=z["ghost_owner"], ilist=z["ilist"], numneigh=z["numneigh"], neigh=z["neigh"], box=z["box"])
    ref = {k: z[k] for k in ("eng_vdwl", "eatom", "f", "virial_pair", "virial_fdotr", "vatom")}
    ref["eng_vdwl"] = float(ref["eng_vdwl"])
    return cfg, [str(e) for e in z["elements"]], ref

"""Shared helpers for the test-suite: golden fixtures -> Config / potential file."""
from __future__ import annotations

import json
import os

import numpy as np

from meng_zhang_b200 import lattice as L
from meng_zhang_b200.pair import AnnPotential, write_potential

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FE_CASES = ["bcc4_perfect", "bcc4_perturbed", "bcc334_hot", "cluster_ragged", "bcc4_two_types", "bcc10_perturbed"]


def load_potential_json(name="fe_potential.json") -> AnnPotential:
    with open(os.path.join(GOLDEN, name)) as fp:
        d = json.load(fp)
    d.pop("source", None)
    d["sfnor_cov"] = np.array(d["sfnor_cov"])
    d["sfnor_avg"] = np.array(d["sfnor_avg"])
    d["weight_all"] = np.array(d["weight_all"])
    d["bias_all"] = np.array(d["bias_all"])
    return AnnPotential(**d)


def write_fe_potential(path) -> str:
    write_potential(str(path), load_potential_json(), comment="ANN potential for Fe re-written from tests/golden/fe_potential.json")
    return str(path)


def load_case(name):
    z = np.load(os.path.join(GOLDEN, f"annp_fe_{name}.npz"))
    cfg = L.Config(nlocal=int(z["nlocal"]), nghost=int(z["nghost"]), x=z["x"], type=z["type"],
                   ghost_owner=z["ghost_owner"], ilist=z["ilist"], numneigh=z["numneigh"], neigh=z["neigh"], box=z["box"])
    ref = {k: z[k] for k in ("eng_vdwl", "eatom", "f", "virial_pair", "virial_fdotr", "vatom")}
    ref["eng_vdwl"] = float(ref["eng_vdwl"])
    return cfg, [str(e) for e in z["elements"]], ref

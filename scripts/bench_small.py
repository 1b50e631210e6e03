"""BASELINE config 1 (bcc Fe ANNP, 10x10x10 cells = 2 000 atoms, NVE): eager launches vs CUDA-graph replay.
Development aid; prints one JSON line."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import util  # noqa: E402
from meng_zhang_b200 import lattice as L  # noqa: E402
from meng_zhang_b200.md import DomainMD  # noqa: E402
from meng_zhang_b200.pair import PairANNPGPU  # noqa: E402

pair = PairANNPGPU(ntypes=1)
pair.settings([])
pair.coeff(["*", "*", util.write_fe_potential("/tmp/bs_fe.ann"), "Fe"])
pair.init_style()
x, box = L.bcc(10, 10, 10)
md = DomainMD(pair, x, box)
md.set_velocities(300.0, 4928459)
md.reneighbor()
md.compute(eflag=True)
N = 2000


def timed(fn, n):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t = time.perf_counter()
    e0.record()
    fn(n)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (time.perf_counter() - t) * 1e3 / n


def eager(n):
    for _ in range(n):
        md.step()


eager(50)
ms_eager, wall_eager = timed(eager, N)
x_eager = md.x[: md.nlocal].clone()
md.capture_step()
md.replay(50)
ms_graph, wall_graph = timed(md.replay, N)
md.step(eflag=True)
pe, ke = md.thermo()
print(json.dumps({"config": "1: bcc Fe ANNP 2 000 atoms NVE", "atoms": md.nlocal, "eager_ms_per_step": ms_eager, "graph_ms_per_step": ms_graph,
                  "eager_atom_steps_per_s": md.nlocal / (ms_eager * 1e-3), "graph_atom_steps_per_s": md.nlocal / (ms_graph * 1e-3),
                  "graph_ns_per_day": 86400.0 / (ms_graph * 1e-3) * 1e-6, "steps_total": md.nsteps, "etot_per_atom": (pe + ke) / md.nlocal}))

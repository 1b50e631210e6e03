"""Run under torch.distributed.run with N ranks: decomposed device-resident MD vs the same box on one GPU.

Checks (a) forces of the decomposed box equal the single-domain forces to rounding, (b) after 20 NVE steps
the positions agree (the halo exchange keeps ghosts consistent), (c) total energy is the all-reduced sum."""
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import util  # noqa: E402
from meng_zhang_b200 import lattice as L  # noqa: E402
from meng_zhang_b200.md import DomainMD, decompose, rank_coords  # noqa: E402
from meng_zhang_b200.pair import PairANNPGPU  # noqa: E402


def make_pair(dev):
    pot = util.write_fe_potential(os.path.join(tempfile.gettempdir(), f"annp_b200_multi_{dev}.ann"))
    pair = PairANNPGPU(ntypes=1, device=dev)
    pair.settings([])
    pair.coeff(["*", "*", pot, "Fe"])
    pair.init_style()
    return pair


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    grid = decompose(world)
    cells = (8 * grid[0], 8 * grid[1], 8 * grid[2])
    x_all, box = L.bcc(*cells)
    x_all = L.wrap(L.perturb(x_all, 0.05, 5), box)
    rng = np.random.default_rng(9)
    v_all = rng.normal(size=x_all.shape) * 2.0
    v_all -= v_all.mean(axis=0)
    coords = rank_coords(rank, grid)
    lo = np.array([box[d] * coords[d] / grid[d] for d in range(3)])
    hi = np.array([box[d] * (coords[d] + 1) / grid[d] for d in range(3)])
    mine = np.all((x_all >= lo) & (x_all < hi), axis=1)
    gid = torch.as_tensor(np.nonzero(mine)[0], device=dev)

    pair = make_pair(local)
    md = DomainMD(pair, x_all[mine], box, grid=grid, rank=rank, device=dev)
    md.v = torch.as_tensor(v_all[mine], device=dev)
    md.reneighbor()
    md.compute(eflag=True)
    f0 = md.f[: md.nlocal].clone()
    pe_dec, _ = md.thermo()
    for _ in range(20):
        md.step()
    x20 = md.x[: md.nlocal].clone()

    # gather on rank 0
    n_all = len(x_all)
    fg = torch.zeros((n_all, 3), dtype=torch.float64, device=dev)
    xg = torch.zeros((n_all, 3), dtype=torch.float64, device=dev)
    fg[gid] = f0
    xg[gid] = x20
    dist.all_reduce(fg)
    dist.all_reduce(xg)
    ok = True
    if rank == 0:
        pair1 = make_pair(local)
        md1 = DomainMD(pair1, x_all, box, grid=(1, 1, 1), rank=0, device=dev)
        md1.v = torch.as_tensor(v_all, device=dev)
        md1.reneighbor()
        md1.compute(eflag=True)
        f1 = md1.f[: md1.nlocal].clone()
        pe1 = float(md1.engvir[0])
        for _ in range(20):
            md1.step()
        df = float((fg - f1).abs().max())
        dx = float((xg - md1.x[: md1.nlocal]).abs().max())
        de = abs(pe_dec - pe1) / abs(pe1)
        print(f"multi_gpu_check world={world} grid={grid} atoms={n_all}: max|dF| {df:.3e}  max|dx| after 20 steps {dx:.3e}  rel dE {de:.3e}")
        ok = df < 1e-11 and dx < 1e-11 and de < 1e-13
        print("MULTI_GPU_CHECK", "PASS" if ok else "FAIL")
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()

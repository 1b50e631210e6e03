"""Domain decomposition + halo bookkeeping on CPU (gloo, world_size 2 and 4).

The device kernels (halo_pack / halo_unpack_add / force) need a GPU; here the same send lists, split sizes
and all_to_all exchanges that meng_zhang_b200/md.py drives over NCCL are exercised with numpy standing in
for the pack/unpack kernels and the CPU oracle standing in for the force kernel.  Checked:
  * every rank ends up with exactly the ghost shell LAMMPS would give it
  * forward + force + reverse reproduces the single-domain forces to rounding (invariance to P)
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import util
from meng_zhang_b200 import lattice as L
from meng_zhang_b200.md import build_send_lists, coords_rank, decompose, rank_coords


def test_decompose_and_rank_mapping():
    assert decompose(1) == (1, 1, 1) and decompose(2) == (2, 1, 1) and decompose(4) == (2, 2, 1) and decompose(8) == (2, 2, 2)
    for n in (1, 2, 4, 8, 6):
        g = decompose(n)
        assert g[0] * g[1] * g[2] == n
        for r in range(n):
            assert coords_rank(rank_coords(r, g), g) == r


def test_send_lists_single_rank_equal_periodic_ghosts():
    x, box = L.bcc(4, 4, 4)
    x = L.wrap(L.perturb(x, 0.05, 3), box)
    idx, shift, counts = build_send_lists(x, np.zeros(3), box, box, (1, 1, 1), (0, 0, 0), 8.5)
    gx, gowner, _ = L.make_ghosts(x, box, 8.5)
    mine = np.round(x[idx] + shift, 9)
    theirs = np.round(gx, 9)
    assert len(mine) == len(theirs) == counts[0]
    assert {tuple(r) for r in mine} == {tuple(r) for r in theirs}


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, cells, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import restatement
        from meng_zhang_b200.pair import read_potential
        pot = read_potential(util.write_fe_potential(os.path.join(out_dir, f"fe_{rank}.ann")), ["Fe"])
        grid = decompose(world)
        x_all, box = L.bcc(*cells)
        x_all = L.wrap(L.perturb(x_all, 0.05, 21), box)
        coords = rank_coords(rank, grid)
        lo = np.array([box[d] * coords[d] / grid[d] for d in range(3)])
        hi = np.array([box[d] * (coords[d] + 1) / grid[d] for d in range(3)])
        mine = np.all((x_all >= lo) & (x_all < hi), axis=1)
        gid = np.nonzero(mine)[0]
        xl = x_all[mine]
        nlocal = len(xl)
        idx, shift, send_counts = build_send_lists(xl, lo, hi, box, grid, coords, 8.5)
        sc = torch.tensor(send_counts, dtype=torch.int64)
        rc = torch.empty_like(sc)
        dist.all_to_all_single(rc, sc)
        recv_counts = [int(c) for c in rc]
        # forward: pack -> exchange -> ghost block
        sendbuf = torch.from_numpy(np.ascontiguousarray(xl[idx] + shift))
        ghosts = torch.empty((sum(recv_counts), 3), dtype=torch.float64)
        dist.all_to_all_single(ghosts, sendbuf, recv_counts, [int(c) for c in send_counts])
        x = np.concatenate([xl, ghosts.numpy()])
        # the ghost shell is exactly the set of periodic images within 8.5 A of the brick
        gx, _, _ = L.make_ghosts(x_all, box, 8.5 + max(box))     # all images around the box
        allimg = np.concatenate([x_all, gx])
        inside = np.all((allimg >= lo - 8.5) & (allimg < hi + 8.5), axis=1) & ~np.all((allimg >= lo) & (allimg < hi), axis=1)
        assert {tuple(r) for r in np.round(allimg[inside], 8)} == {tuple(r) for r in np.round(ghosts.numpy(), 8)}
        # force evaluation on the sub-domain (CPU oracle stands in for the kernel)
        numneigh, neigh = L.full_neighbor_list(x, nlocal, 8.5)
        cfg = L.Config(nlocal=nlocal, nghost=len(ghosts), x=x, type=np.ones(len(x), dtype=np.int32), ghost_owner=np.zeros(len(ghosts), dtype=np.int32),
                       ilist=np.arange(nlocal, dtype=np.int32), numneigh=numneigh, neigh=neigh, box=box)
        o = restatement.compute(pot, cfg)
        f = o["f"]
        # reverse: ghost forces -> owners, ordered accumulate
        back = torch.empty((len(idx), 3), dtype=torch.float64)
        dist.all_to_all_single(back, torch.from_numpy(np.ascontiguousarray(f[nlocal:])), [int(c) for c in send_counts], recv_counts)
        floc = f[:nlocal].copy()
        np.add.at(floc, idx, back.numpy())
        e = torch.tensor([o["eng_vdwl"]], dtype=torch.float64)
        dist.all_reduce(e)
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), gid=gid, f=floc, e=e.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,cells", [(2, (8, 4, 4)), (4, (8, 8, 4))])
def test_decomposed_forces_equal_single_domain(world, cells, tmp_path, fe_pot_file):
    from oracle import restatement
    from meng_zhang_b200.pair import read_potential
    mp.spawn(_worker, args=(world, _free_port(), cells, str(tmp_path)), nprocs=world, join=True)
    x_all, box = L.bcc(*cells)
    x_all = L.wrap(L.perturb(x_all, 0.05, 21), box)
    cfg = L.build_config(x_all, box, 6.5)
    ref = restatement.compute(read_potential(fe_pot_file, ["Fe"]), cfg, nthreads=4)
    fref = cfg.fold(ref["f"])
    got = np.zeros_like(fref)
    seen = np.zeros(len(fref), dtype=int)
    for r in range(world):
        z = np.load(tmp_path / f"rank{r}.npz")
        got[z["gid"]] = z["f"]
        seen[z["gid"]] += 1
        assert abs(float(z["e"][0]) - ref["eng_vdwl"]) < 1e-9 * abs(ref["eng_vdwl"])
    assert np.all(seen == 1)
    assert np.abs(got - fref).max() < 1e-11


# ------------------------------------------------------------------------------------------------ atom migration
def _migrate_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from meng_zhang_b200.md import DomainMD
        grid = decompose(world)
        box = np.array([20.0, 16.0, 12.0])
        rng = np.random.default_rng(100)
        n_all = 4000
        x_all = rng.uniform(0.0, 1.0, size=(n_all, 3)) * box
        coords = rank_coords(rank, grid)
        lo = np.array([box[d] * coords[d] / grid[d] for d in range(3)])
        hi = np.array([box[d] * (coords[d] + 1) / grid[d] for d in range(3)])
        mine = np.all((x_all >= lo) & (x_all < hi), axis=1)
        gid = np.nonzero(mine)[0]
        # every atom moves by up to +-6 A (some leave through periodic faces, some cross to non-adjacent bricks)
        disp = np.random.default_rng(7).uniform(-6.0, 6.0, size=(n_all, 3))
        md = DomainMD.__new__(DomainMD)                       # bookkeeping only: no GPU, no pair style
        md.dev, md.world, md.rank, md.group, md.grid = torch.device("cpu"), world, rank, None, grid
        md.box, md.box_origin, md.periodic = box, np.zeros(3), (True, True, True)
        md.nlocal = int(mine.sum())
        md.v = torch.from_numpy(np.stack([gid, 2.0 * gid, 3.0 * gid], axis=1).astype(np.float64))     # recognisable payload
        md._type_local = torch.from_numpy((1 + gid % 2).astype(np.int32))
        fz = np.nonzero(gid % 5 == 0)[0]
        md.frozen_idx = torch.from_numpy(fz.astype(np.int64))
        md.gid = torch.from_numpy(gid.astype(np.int64))
        xl = torch.from_numpy(x_all[mine] + disp[mine])
        fl = torch.from_numpy(np.stack([-1.0 * gid, 0.5 * gid, 7.0 + gid], axis=1).astype(np.float64))
        xn, fn = md._migrate(xl, fl)
        g = md.gid.numpy()
        assert np.all((xn.numpy() >= lo - 1e-12) & (xn.numpy() < hi + 1e-12)), "an atom is outside its new owner's brick"
        want = x_all[g] + disp[g]
        want -= np.floor(want / box) * box
        assert np.abs(xn.numpy() - want).max() < 1e-12
        assert np.array_equal(md.v.numpy(), np.stack([g, 2.0 * g, 3.0 * g], axis=1))
        assert np.array_equal(fn.numpy(), np.stack([-1.0 * g, 0.5 * g, 7.0 + g], axis=1))
        assert np.array_equal(md._type_local.numpy(), 1 + g % 2)
        frozen = np.zeros(md.nlocal, dtype=bool)
        if md.frozen_idx is not None:
            frozen[md.frozen_idx.numpy()] = True
        assert np.array_equal(frozen, g % 5 == 0)
        np.save(os.path.join(out_dir, f"gid{rank}.npy"), g)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_atom_migration_keeps_every_atom_once_with_its_payload(world, tmp_path):
    mp.spawn(_migrate_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    allg = np.concatenate([np.load(tmp_path / f"gid{r}.npy") for r in range(world)])
    assert len(allg) == 4000 and np.array_equal(np.sort(allg), np.arange(4000))


def test_send_slots_describe_the_host_send_lists():
    """send_slots (what the device ghost map is told) and build_send_lists (the numpy twin) agree: applying the slots'
    face tests in slot order reproduces the twin's index / shift / count arrays entry by entry."""
    from meng_zhang_b200.md import send_slots
    rng = np.random.default_rng(3)
    box = np.array([40.0, 36.0, 44.0])
    for grid, periodic in (((1, 1, 1), (True, True, True)), ((2, 2, 1), (True, False, True)), ((2, 2, 2), (True, True, True))):
        nr = grid[0] * grid[1] * grid[2]
        for rank in range(nr):
            coords = rank_coords(rank, grid)
            lo = np.array([box[d] * coords[d] / grid[d] for d in range(3)])
            hi = np.array([box[d] * (coords[d] + 1) / grid[d] for d in range(3)])
            x = lo + rng.random((500, 3)) * (hi - lo)
            idx, shift, counts = build_send_lists(x, lo, hi, box, grid, coords, 8.5, periodic)
            dirs, shifts, dests = send_slots(lo, hi, box, grid, coords, 8.5, periodic)
            got_i, got_s, got_c = [], [], np.zeros(nr, dtype=np.int64)
            for k in range(len(dests)):
                m = np.ones(len(x), dtype=bool)
                for d in range(3):
                    if dirs[k, d] == 1:
                        m &= x[:, d] >= hi[d] - 8.5
                    elif dirs[k, d] == -1:
                        m &= x[:, d] < lo[d] + 8.5
                sel = np.nonzero(m)[0]
                got_i.append(sel)
                got_s.append(np.broadcast_to(shifts[k], (len(sel), 3)))
                got_c[dests[k]] += len(sel)
            assert np.array_equal(np.concatenate(got_i), idx)
            assert np.array_equal(np.concatenate(got_s), shift)
            assert np.array_equal(got_c, counts)

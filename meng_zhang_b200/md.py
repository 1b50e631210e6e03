"""Device-resident MD driver for the ANNP path (LAMMPS is not available in this image).

Plays the role of LAMMPS' `Verlet::run` around `Pair::compute` for one rank per GPU
(SURVEY.md section 3 C, section 8e): spatial brick decomposition of a periodic orthogonal box,
ghost shell of `cutoff + skin`, forward communication of ghost positions and reverse communication of
ghost forces every step, velocity-Verlet NVE.  Everything stays in HBM:

    nve_initial -> halo_pack -> [NCCL all_to_all_single] -> compute_device -> [NCCL] -> halo_unpack_add -> nve_final

The ghost map (which local atoms each neighbouring brick needs) is built on the device at every re-neighbouring
(annp_b200_send_lists_*), and the exchange itself runs below the C ABI: one grouped ncclSend/ncclRecv over NVLink per
direction on the handle's own NCCL communicator (annp_b200_halo_forward / _reverse); with one rank the "exchange" is the
pack kernel writing straight into the ghost block.  PyTorch is used for device allocations, streams and to hand the NCCL
unique id to the ranks - all arithmetic and all per-step communication is in libannp_b200.so, so a step is plain stream
work and can be replayed as a CUDA graph on any number of ranks.
"""
from __future__ import annotations

import ctypes as C
import itertools

import numpy as np
import torch

from . import capi

KB = 8.617343e-5          # eV/K        (LAMMPS units metal: force->boltz)
MVV2E = 1.0364269e-4      # (g/mol)(A/ps)^2 -> eV


def decompose(nranks: int):
    """1/2/4/8 -> 1x1x1 / 2x1x1 / 2x2x1 / 2x2x2; otherwise the most cubic factorisation."""
    best = None
    for px in range(1, nranks + 1):
        if nranks % px:
            continue
        for py in range(1, nranks // px + 1):
            if (nranks // px) % py:
                continue
            pz = nranks // px // py
            key = (max(px, py, pz) - min(px, py, pz), -px, -py)
            if best is None or key < best[0]:
                best = (key, (px, py, pz))
    return best[1]


def rank_coords(rank, grid):
    px, py, pz = grid
    return (rank % px, (rank // px) % py, rank // (px * py))


def coords_rank(c, grid):
    px, py, pz = grid
    return (c[0] % px) + (c[1] % py) * px + (c[2] % pz) * px * py


def send_slots(lo, hi, box, grid, coords, cutghost: float, periodic=(True, True, True)):
    """The (up to 26) slots of the send list in list order: by destination rank, then by direction.  Returns
    (dirs[int32 nslots,3], shifts[nslots,3], dests[nslots])."""
    for d in range(3):
        if hi[d] - lo[d] < cutghost:
            raise ValueError("sub-domain thinner than the ghost cutoff: multi-hop halo not supported")
    slots = []
    for order, s in enumerate(itertools.product((-1, 0, 1), repeat=3)):
        if s == (0, 0, 0):
            continue
        if any(s[d] != 0 and not periodic[d] and not (0 <= coords[d] + s[d] < grid[d]) for d in range(3)):
            continue        # leaves the box through a free surface: no receiver
        shift = np.zeros(3)
        for d in range(3):
            c = coords[d] + s[d]
            if c >= grid[d]:
                shift[d] = -box[d]
            elif c < 0:
                shift[d] = box[d]
        slots.append((coords_rank([coords[d] + s[d] for d in range(3)], grid), order, s, shift))
    slots.sort(key=lambda t: (t[0], t[1]))
    dirs = np.array([t[2] for t in slots], dtype=np.int32).reshape(-1, 3)
    shifts = np.array([t[3] for t in slots], dtype=np.float64).reshape(-1, 3)
    dests = np.array([t[0] for t in slots], dtype=np.int64)
    return np.ascontiguousarray(dirs), np.ascontiguousarray(shifts), dests


def build_send_lists(x_local: np.ndarray, lo, hi, box, grid, coords, cutghost: float, periodic=(True, True, True)):
    """Host twin (numpy) of the device ghost map annp_b200_send_lists_count / _fill: the CPU tests check the
    decomposition bookkeeping with it and the GPU tests check the kernels against it entry by entry.

    Send entries for the 26 directions, grouped by destination rank.  A direction that leaves the box through a
    non-periodic face (`boundary m`/`f`/`s` of the decks: free surface) has no receiver and is skipped.

    Returns (index[int32 nsend], shift[nsend,3], send_counts[nranks]) with entries ordered by
    destination rank and, inside one destination, by direction."""
    nranks = grid[0] * grid[1] * grid[2]
    for d in range(3):
        if hi[d] - lo[d] < cutghost:
            raise ValueError("sub-domain thinner than the ghost cutoff: multi-hop halo not supported")
    per_dest = [[] for _ in range(nranks)]
    for s in itertools.product((-1, 0, 1), repeat=3):
        if s == (0, 0, 0):
            continue
        mask = np.ones(len(x_local), dtype=bool)
        shift = np.zeros(3)
        if any(s[d] != 0 and not periodic[d] and not (0 <= coords[d] + s[d] < grid[d]) for d in range(3)):
            continue
        for d in range(3):
            if s[d] == 1:
                mask &= x_local[:, d] >= hi[d] - cutghost
            elif s[d] == -1:
                mask &= x_local[:, d] < lo[d] + cutghost
            c = coords[d] + s[d]
            if c >= grid[d]:
                shift[d] = -box[d]
            elif c < 0:
                shift[d] = box[d]
        idx = np.nonzero(mask)[0].astype(np.int32)
        if len(idx):
            dest = coords_rank([coords[d] + s[d] for d in range(3)], grid)
            per_dest[dest].append((idx, shift))
    index, shifts, counts = [], [], np.zeros(nranks, dtype=np.int64)
    for dest in range(nranks):
        for idx, shift in per_dest[dest]:
            index.append(idx)
            shifts.append(np.broadcast_to(shift, (len(idx), 3)))
            counts[dest] += len(idx)
    if index:
        return np.concatenate(index), np.ascontiguousarray(np.concatenate(shifts)), counts
    return np.zeros(0, dtype=np.int32), np.zeros((0, 3)), counts


class DomainMD:
    """One rank's sub-domain, device resident.  `pair` is an initialised PairANNPGPU."""

    def __init__(self, pair, x_local, box, grid=(1, 1, 1), rank=0, device=None, type_local=None,
                 skin=2.0, mass=55.845, dt=0.001, group=None, periodic=(True, True, True), frozen_local=None, gid_local=None, list_cutoff=None,
                 shrink_wrap=(False, False, False)):
        """periodic: per-axis `boundary p` (True) or free surface (False).  frozen_local: boolean mask of atoms held
        fixed (force and velocity zeroed every step: `fix setforce 0 0 0` on the rim of the dislocation cylinder)."""
        self.periodic = tuple(bool(p) for p in periodic)
        # LAMMPS `boundary m` on a free-surface axis: the box edge follows the atoms' extent at every re-neighbouring
        # (never inside the initial box) and enters the pressure through the volume
        self.shrink_wrap = tuple(bool(w) and not self.periodic[d] for d, w in enumerate(shrink_wrap))
        self.pair = pair
        self.L = capi.lib()
        self.h = pair.handle
        self.grid, self.rank = tuple(grid), rank
        self.world = grid[0] * grid[1] * grid[2]
        self.group = group
        self.dev = torch.device(device if device is not None else "cuda")
        self.box = np.asarray(box, dtype=np.float64)
        self.coords = rank_coords(rank, self.grid)
        self.lo = np.array([self.box[d] * self.coords[d] / self.grid[d] for d in range(3)])
        self.hi = np.array([self.box[d] * (self.coords[d] + 1) / self.grid[d] for d in range(3)])
        # list / ghost radius: what can actually interact (for the Ni copy tighter than the file's list cutoff)
        self.cut = pair.interaction_cutoff() if hasattr(pair, "interaction_cutoff") else pair.cutmax
        if list_cutoff is not None:
            self.cut = float(list_cutoff)
        self.skin, self.mass, self.dt = skin, mass, dt
        self.nlocal = len(x_local)
        self.nghost = 0
        self._x_local0 = torch.as_tensor(np.ascontiguousarray(x_local), dtype=torch.float64, device=self.dev)
        self._type_local = torch.ones(self.nlocal, dtype=torch.int32, device=self.dev) if type_local is None else \
            torch.as_tensor(np.ascontiguousarray(type_local), dtype=torch.int32, device=self.dev)
        self.v = torch.zeros((self.nlocal, 3), dtype=torch.float64, device=self.dev)
        self.engvir = torch.zeros(8, dtype=torch.float64, device=self.dev)
        self.ke = torch.zeros(1, dtype=torch.float64, device=self.dev)
        self.x = self.f = self.type = None
        self.nsteps = 0
        self.box_origin = np.zeros(3)
        self.box_min_lo, self.box_min_hi = np.zeros(3), self.box.copy()      # data-file box: the shrink-wrap minimum
        self.wrap_lo, self.wrap_hi = np.zeros(3), self.box.copy()
        self.nh = None
        # global atom ids (travel with the atoms when they migrate); default: rank-concatenated numbering
        self.gid = None if gid_local is None else torch.as_tensor(np.ascontiguousarray(gid_local), dtype=torch.int64, device=self.dev)
        self.migrated = 0
        self.frozen_idx = None
        if frozen_local is not None and np.any(frozen_local):
            self.frozen_idx = torch.as_tensor(np.nonzero(np.asarray(frozen_local))[0], dtype=torch.int64, device=self.dev)
        self.disp2 = torch.zeros(1, dtype=torch.float64, device=self.dev)
        # peer scatter (annp_b200_peer_*): ghost forces go straight to their owners' accumulators over NVLink, fusing the
        # reverse exchange into the force kernel.  Bit-identical to the exchange path, but MEASURED SLOWER at the bench size
        # (8 GPUs: 24.69 vs 24.33 ms per step - 22 M remote atomics per step lose against one 3.5 MB bulk transfer), so it
        # is opt-in: ANNP_B200_PEER=1.  Switched off for the run if any rank cannot map its peers' memory.
        import os
        self.peer = self.world > 1 and os.environ.get("ANNP_B200_PEER", "0") == "1"
        self._init_comm()

    def close(self):
        """Drop the captured CUDA graph (it references the handle's NCCL communicator, which must not be destroyed under
        it) and wait for the device: call before pair.clear() when capture_step was used."""
        self._graph = None
        torch.cuda.synchronize(self.dev)

    def _open_peers(self, nall):
        """Peer scatter set-up of a re-neighbouring: exchange the IPC handles of the accumulator arrays and tell every ghost
        its owner (rank, local index = the owner's send-list entry)."""
        import sys
        import torch.distributed as dist
        # mappings of the previous list are dropped on every rank BEFORE any rank may re-allocate its (exported) array
        self.L.annp_b200_peer_close(self.h)
        dist.barrier(group=self.group)
        buf = C.create_string_buffer(64)
        rc = self.L.annp_b200_peer_export(self.h, nall, buf)
        mine = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).to(self.dev)
        handles = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(handles, mine, group=self.group)
        all64 = b"".join(bytes(t.cpu().numpy().tobytes()) for t in handles)
        self._peer_gidx = torch.empty(max(self.nghost, 1), dtype=torch.int32, device=self.dev)[: self.nghost]
        dist.all_to_all_single(self._peer_gidx, self.send_index.contiguous(), self.recv_counts, self.send_counts, group=self.group)
        self._peer_grank = torch.repeat_interleave(torch.arange(self.world, dtype=torch.int32, device=self.dev),
                                                   torch.as_tensor(self.recv_counts, device=self.dev)).contiguous()
        if rc == 0:
            rc = self.L.annp_b200_peer_open(self.h, self.world, all64, self.nghost, C.c_void_p(self._peer_grank.data_ptr()),
                                            C.c_void_p(self._peer_gidx.data_ptr()), self._stream())
        ok = torch.tensor([1 if rc == 0 else 0], dtype=torch.int32, device=self.dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
        if int(ok) == 0:            # some rank cannot map peer memory (no IPC in this container, ...): everybody falls back
            if rc != 0:
                print(f"annp_b200: peer scatter unavailable on rank {self.rank} ({self.L.annp_b200_last_error(self.h).decode()}); "
                      "using the NCCL reverse exchange", file=sys.stderr)
            self.L.annp_b200_peer_close(self.h)
            self.peer = False

    def _init_comm(self):
        """NCCL communicator of the handle (annp_b200_comm_init): rank 0 draws the unique id, torch.distributed only carries
        its 128 bytes to the other ranks (inside LAMMPS an MPI_Bcast would)."""
        import contextlib
        import os
        import sys

        @contextlib.contextmanager
        def stdout_to_stderr():
            # NCCL announces itself on stdout when NCCL_DEBUG is VERSION / WARN; stdout belongs to the caller's own output
            # (bench.py prints exactly one JSON line), so NCCL is initialised with fd 1 pointing at stderr
            sys.stdout.flush()
            saved = os.dup(1)
            os.dup2(2, 1)
            try:
                yield
            finally:
                os.dup2(saved, 1)
                os.close(saved)

        ident = None
        if self.world > 1:
            import torch.distributed as dist
            buf = C.create_string_buffer(128)
            if self.rank == 0:
                with stdout_to_stderr():
                    rc = self.L.annp_b200_comm_unique_id(buf)
                if rc != 0:
                    raise capi.AnnpError(rc, "NCCL unavailable: annp_b200_comm_unique_id failed")
            t = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).to(self.dev)
            ranks = dist.get_process_group_ranks(self.group) if self.group is not None else None
            dist.broadcast(t, src=ranks[0] if ranks else 0, group=self.group)
            ident = bytes(t.cpu().numpy().tobytes())
        with stdout_to_stderr():
            rc = self.L.annp_b200_comm_init(self.h, self.world, self.rank, ident)
        self._ck(rc)

    # ------------------------------------------------------------------ helpers
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)

    def _ck(self, rc):
        if rc != 0:
            raise capi.AnnpError(rc, self.L.annp_b200_last_error(self.h).decode())

    def set_velocities(self, temperature: float, seed: int):
        """Gaussian velocities at `temperature`, zero net momentum (own RNG; the deck's
        `velocity all create 300 4928459` uses LAMMPS' generator, in.st_test:29)."""
        g = torch.Generator(device="cpu").manual_seed(seed + 7919 * self.rank)
        sigma = (KB * temperature / (self.mass * MVV2E)) ** 0.5
        v = torch.randn((self.nlocal, 3), generator=g, dtype=torch.float64) * sigma
        v -= v.mean(dim=0, keepdim=True)
        self.v = v.to(self.dev)
        if self.frozen_idx is not None:
            self.v.index_fill_(0, self.frozen_idx, 0.0)

    # ------------------------------------------------------------------ re-neighbouring (ago == 0)
    def _migrate(self, xl, fl):
        """Atom migration of LAMMPS' re-neighbouring (Comm::exchange): wrap positions into the periodic box and hand
        every atom to the rank whose brick now contains it, together with its velocity, last force (the next half kick
        needs it), type, frozen flag and global id.  One all_to_all of the per-destination counts and one of the payload;
        arrival order is (source rank, source order), so the result is deterministic."""
        origin = torch.as_tensor(self.box_origin, dtype=torch.float64, device=self.dev)
        boxd = torch.as_tensor(self.box, dtype=torch.float64, device=self.dev)
        per = torch.as_tensor(self.periodic, dtype=torch.bool, device=self.dev)
        xl = torch.where(per, xl - torch.floor((xl - origin) / boxd) * boxd, xl)
        if self.gid is None:
            counts = torch.zeros(self.world, dtype=torch.int64, device=self.dev)
            counts[self.rank] = self.nlocal
            if self.world > 1:
                import torch.distributed as dist
                dist.all_reduce(counts, group=self.group)
            start = int(counts[: self.rank].sum())
            self.gid = torch.arange(start, start + self.nlocal, dtype=torch.int64, device=self.dev)
        if self.world == 1:
            return xl, fl
        import torch.distributed as dist
        grid = torch.as_tensor(self.grid, dtype=torch.float64, device=self.dev)
        cell = torch.floor((xl - origin) / (boxd / grid)).clamp_(min=0).to(torch.int64)
        cell = torch.minimum(cell, torch.as_tensor(self.grid, dtype=torch.int64, device=self.dev) - 1)
        dest = cell[:, 0] + cell[:, 1] * self.grid[0] + cell[:, 2] * self.grid[0] * self.grid[1]
        order = torch.argsort(dest, stable=True)
        send_counts = torch.bincount(dest, minlength=self.world)
        frozen = torch.zeros(self.nlocal, dtype=torch.float64, device=self.dev)
        if self.frozen_idx is not None:
            frozen[self.frozen_idx] = 1.0
        payload = torch.cat([xl, self.v, fl, self._type_local.to(torch.float64)[:, None], frozen[:, None],
                             self.gid.to(torch.float64)[:, None]], dim=1)[order].contiguous()
        recv_counts = torch.empty_like(send_counts)
        dist.all_to_all_single(recv_counts, send_counts, group=self.group)
        sc, rc = [int(c) for c in send_counts.cpu()], [int(c) for c in recv_counts.cpu()]
        self.migrated = self.nlocal - sc[self.rank]
        got = torch.empty((sum(rc), payload.shape[1]), dtype=torch.float64, device=self.dev)
        dist.all_to_all_single(got, payload, rc, sc, group=self.group)
        self.nlocal = got.shape[0]
        self.v = got[:, 3:6].contiguous()
        self._type_local = got[:, 9].to(torch.int32).contiguous()
        fz = torch.nonzero(got[:, 10] > 0.5).flatten()
        self.frozen_idx = fz if fz.numel() else None
        self.gid = got[:, 11].to(torch.int64).contiguous()
        return got[:, 0:3].contiguous(), got[:, 6:9].contiguous()

    def _shrink_wrap_box(self, xl):
        from .lammps_compat import shrink_wrap
        ext = torch.stack([xl.amin(dim=0), xl.amax(dim=0)]) if xl.shape[0] else torch.zeros((2, 3), dtype=torch.float64, device=self.dev)
        ext[0] = -ext[0]
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(ext, op=dist.ReduceOp.MAX, group=self.group)
        ext = ext.cpu().numpy()
        which = np.zeros(3, dtype=np.int32)
        for d in range(3):
            if self.shrink_wrap[d]:
                self.wrap_lo[d], self.wrap_hi[d] = shrink_wrap(self.box_min_lo[d], self.box_min_hi[d], -ext[0, d], ext[1, d])
                which[d] = 1
        if self.nh is not None:
            lo, hi = self.wrap_lo.copy(), self.wrap_hi.copy()
            self._nh_ck(self.L.annp_b200_nh_set_box(self.nh, lo.ctypes.data_as(capi.c_double_p), hi.ctypes.data_as(capi.c_double_p),
                                                    which.ctypes.data_as(capi.c_int_p), self._stream()))

    def reneighbor(self, migrate: bool = True):
        """Migrate atoms, rebuild send lists, ghosts and the device neighbour list from the current local positions.
        migrate=False keeps every atom on its rank, unwrapped and in place (used inside a line search, where the
        minimiser's direction vectors must stay aligned with the atoms)."""
        self._graph = None          # a captured step refers to the buffers and the list replaced below
        xl = (self.x[: self.nlocal] if self.x is not None else self._x_local0)
        fl = self.f[: self.nlocal] if self.f is not None else torch.zeros_like(xl)
        if migrate or self.gid is None:
            xl, fl = self._migrate(xl, fl)
        cutghost = self.cut + self.skin
        if any(self.shrink_wrap):
            self._shrink_wrap_box(xl)
        # ghost map on the device: classify the local atoms against the (up to 26) slots; only the slot counts come back
        xl = xl.contiguous()
        dirs, shifts, dests = send_slots(self.lo, self.hi, self.box, self.grid, self.coords, cutghost, self.periodic)
        nslots = len(dests)
        slot_counts = np.zeros(max(nslots, 1), dtype=np.int32)
        lo_h, hi_h = np.ascontiguousarray(self.lo, dtype=np.float64), np.ascontiguousarray(self.hi, dtype=np.float64)
        self._ck(self.L.annp_b200_send_lists_count(self.h, self.nlocal, C.c_void_p(xl.data_ptr()), lo_h.ctypes.data_as(capi.c_double_p),
                                                   hi_h.ctypes.data_as(capi.c_double_p), float(cutghost), nslots,
                                                   dirs.ctypes.data_as(capi.c_int_p), slot_counts.ctypes.data_as(capi.c_int_p), self._stream()))
        send_counts = np.zeros(self.world, dtype=np.int64)
        np.add.at(send_counts, dests, slot_counts[:nslots])
        self.send_counts = [int(c) for c in send_counts]
        if self.world > 1:
            import torch.distributed as dist
            sc = torch.tensor(self.send_counts, dtype=torch.int64, device=self.dev)
            rc = torch.empty_like(sc)
            dist.all_to_all_single(rc, sc, group=self.group)
            self.recv_counts = [int(c) for c in rc.cpu()]
        else:
            self.recv_counts = list(self.send_counts)
        self.nsend = int(send_counts.sum())
        self.nghost = int(sum(self.recv_counts))
        nall = self.nlocal + self.nghost
        self.send_index = torch.empty(max(self.nsend, 1), dtype=torch.int32, device=self.dev)[: self.nsend]
        self.send_shift = torch.empty((max(self.nsend, 1), 3), dtype=torch.float64, device=self.dev)[: self.nsend]
        self._ck(self.L.annp_b200_send_lists_fill(self.h, shifts.ctypes.data_as(capi.c_double_p), C.c_void_p(self.send_index.data_ptr()),
                                                  C.c_void_p(self.send_shift.data_ptr()), self._stream()))
        x = torch.empty((nall, 3), dtype=torch.float64, device=self.dev)
        x[: self.nlocal] = xl
        self.x = x
        self.f = torch.zeros((nall, 3), dtype=torch.float64, device=self.dev)
        self.f[: self.nlocal] = fl           # forces of the last evaluation: the next half kick uses them
        self._ck(self.L.annp_b200_set_halo(self.h, self.nlocal, self.nsend, C.c_void_p(self.send_index.data_ptr()),
                                           C.c_void_p(self.send_shift.data_ptr()), self._stream()))
        sc_h, rc_h = np.array(self.send_counts, dtype=np.int32), np.array(self.recv_counts, dtype=np.int32)
        self._ck(self.L.annp_b200_set_halo_peers(self.h, self.world, sc_h.ctypes.data_as(capi.c_int_p), rc_h.ctypes.data_as(capi.c_int_p)))
        if self.peer:
            self._open_peers(nall)
        # ghost types travel once per re-neighbouring
        tl = self._type_local
        if self.world > 1:
            import torch.distributed as dist
            tsend = tl[self.send_index.long()].contiguous()
            tghost = torch.empty(self.nghost, dtype=torch.int32, device=self.dev)
            dist.all_to_all_single(tghost, tsend, self.recv_counts, self.send_counts, group=self.group)
        else:
            tghost = tl[self.send_index.long()]
        self.type = torch.cat([tl, tghost]).contiguous()
        self.forward_comm()
        # bounding box of everything the list sees (free surfaces may lie anywhere outside the nominal brick)
        lo = (self.x.amin(dim=0).cpu().numpy() - 1e-6).astype(np.float64) if nall else (self.lo - cutghost)
        hi = (self.x.amax(dim=0).cpu().numpy() + 1e-6).astype(np.float64) if nall else (self.hi + cutghost)
        self._ck(self.L.annp_b200_neigh_build(self.h, self.nlocal, nall, C.c_void_p(self.x.data_ptr()),
                                              lo.ctypes.data_as(capi.c_double_p), hi.ctypes.data_as(capi.c_double_p),
                                              float(cutghost), self._stream()))

    # ------------------------------------------------------------------ communication
    def forward_comm(self):
        """Ghost positions <- owners (Comm::forward_comm): pack kernel + one grouped NCCL send/recv inside the library."""
        self._ck(self.L.annp_b200_halo_forward(self.h, C.c_void_p(self.x.data_ptr()), self._stream()))

    def reverse_comm(self):
        """Ghost forces -> owners, deterministic ordered accumulation (Comm::reverse_comm)."""
        self._ck(self.L.annp_b200_halo_reverse(self.h, C.c_void_p(self.f.data_ptr()), self._stream()))

    # ------------------------------------------------------------------ force evaluation and integration
    def compute(self, eflag=False, vflag=False):
        self.forward_comm()
        ev = C.c_void_p(self.engvir.data_ptr()) if (eflag or vflag) else None
        self._ck(self.L.annp_b200_compute_device(self.h, self.nlocal, self.nghost, C.c_void_p(self.x.data_ptr()),
                                                 C.c_void_p(self.type.data_ptr()), int(eflag), int(vflag),
                                                 C.c_void_p(self.f.data_ptr()), None, ev, None, self._stream()))
        self.reverse_comm()
        if self.frozen_idx is not None:
            self.f.index_fill_(0, self.frozen_idx, 0.0)

    def step(self, eflag=False):
        """One velocity-Verlet step (FixNVE::initial_integrate, force, final_integrate)."""
        s = self._stream()
        self._ck(self.L.annp_b200_nve_initial(self.h, self.nlocal, self.dt, self.mass, C.c_void_p(self.x.data_ptr()),
                                              C.c_void_p(self.v.data_ptr()), C.c_void_p(self.f.data_ptr()), s))
        self.compute(eflag=eflag)
        ke = C.c_void_p(self.ke.data_ptr()) if eflag else None
        self._ck(self.L.annp_b200_nve_final(self.h, self.nlocal, self.dt, self.mass, C.c_void_p(self.v.data_ptr()),
                                            C.c_void_p(self.f.data_ptr()), ke, s))
        self.nsteps += 1

    # ------------------------------------------------------------------ `minimize etol ftol maxiter maxeval` (min_style cg)
    def minimize(self, etol: float, ftol: float, maxiter: int, maxeval: int, dmax: float = 0.1):
        """LAMMPS' default minimiser, restated: Polak-Ribiere conjugate gradients (MinCG::iterate) with the quadratic
        line search of MinLineSearch::linemin_quadratic (trial step alpha = min(1, dmax / max|h|), secant projection of
        the directional derivative when the quadratic model holds, else backtracking with slope 0.4), metal units
        (thermo norm no).  Vectors live on the device; only dot products come to the host.  The neighbour list is
        re-checked at every evaluation as LAMMPS does during minimisation (`every 1 delay 0 check yes`).
        Returns a dict like LAMMPS' "Minimization stats"."""
        ALPHA_MAX, ALPHA_REDUCE, BACKTRACK_SLOPE, QUADRATIC_TOL, EMACH, EPS_QUAD, EPS_ENERGY = 1.0, 0.5, 0.4, 0.1, 1.0e-8, 1.0e-28, 1.0e-8
        if self.x is None:
            self.reneighbor()
        n = self.nlocal
        self._min_xref = self.x[:n].clone()

        def gsum(t):
            t = t.reshape(1).clone()
            if self.world > 1:
                import torch.distributed as dist
                dist.all_reduce(t, group=self.group)
            return float(t)

        def gmax(t):
            t = t.reshape(1).clone()
            if self.world > 1:
                import torch.distributed as dist
                dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
            return float(t)

        def energy_force():
            moved = self._moved_sq(self._min_xref)
            self.check_device_flags()
            if moved > (0.5 * self.skin) ** 2:
                self.reneighbor(migrate=False)
                self._min_xref = self.x[:n].clone()
            self.compute(eflag=True)
            return gsum(self.engvir[0])

        stats = {"evaluations": 0, "iterations": 0}
        ecurrent = energy_force()
        einitial = ecurrent
        f = lambda: self.f[:n]
        h = f().clone()
        g = f().clone()
        gg = gsum((g * g).sum())
        stats["fnorm_initial"], stats["fmax_initial"] = gg ** 0.5, gmax(f().abs().max())
        stop, alpha_final, enext_to_last = "max iterations", 0.0, ecurrent

        def linemin(eoriginal):
            nonlocal ecurrent
            fdothall = gsum((f() * h).sum())
            if fdothall <= 0.0:
                return "search direction is not downhill", 0.0
            hmaxall = gmax(h.abs().max())
            if hmaxall == 0.0:
                return "forces are zero", 0.0
            alphamax = min(ALPHA_MAX, dmax / hmaxall)
            x0 = self.x[:n].clone()

            def alpha_step(a):
                self.x[:n] = x0 + a * h
                stats["evaluations"] += 1
                return energy_force()

            alpha, fhprev, engprev, alphaprev = alphamax, fdothall, eoriginal, 0.0
            while True:
                ecurrent = alpha_step(alpha)
                fh = gsum((f() * h).sum())
                delfh = fh - fhprev
                if abs(fh) < EPS_QUAD or abs(delfh) < EPS_QUAD:
                    self.x[:n] = x0
                    ecurrent = energy_force()
                    return "linesearch alpha is zero (quadratic factors)", 0.0
                relerr = abs(1.0 - (0.5 * (alpha - alphaprev) * (fh + fhprev) + ecurrent) / engprev)
                alpha0 = alpha - (alpha - alphaprev) * fh / delfh
                if relerr <= QUADRATIC_TOL and 0.0 < alpha0 < ALPHA_MAX:
                    ecurrent = alpha_step(alpha0)
                    if ecurrent - eoriginal < EMACH:
                        return None, alpha            # LAMMPS reports the trial alpha, the atoms sit at alpha0
                de_ideal = -BACKTRACK_SLOPE * alpha * fdothall
                de = ecurrent - eoriginal
                if de <= de_ideal:
                    return None, alpha
                fhprev, engprev, alphaprev = fh, ecurrent, alpha
                alpha *= ALPHA_REDUCE
                if alpha <= 0.0 or de_ideal >= -EMACH:
                    self.x[:n] = x0
                    ecurrent = energy_force()
                    return "linesearch alpha is zero", 0.0

        ndof = 3.0 * gsum(torch.tensor(float(n), dtype=torch.float64, device=self.dev))
        for it in range(maxiter):
            stats["iterations"] += 1
            eprevious = ecurrent
            enext_to_last = eprevious
            fail, alpha_final = linemin(ecurrent)
            if fail:
                stop = fail
                break
            if stats["evaluations"] >= maxeval:
                stop = "max force evaluations"
                break
            if abs(ecurrent - eprevious) < etol * 0.5 * (abs(ecurrent) + abs(eprevious) + EPS_ENERGY):
                stop = "energy tolerance"
                break
            ff, fg = gsum((f() * f()).sum()), gsum((f() * g).sum())
            if ftol > 0.0 and ff < ftol * ftol:
                stop = "force tolerance"
                break
            beta = max(0.0, (ff - fg) / gg)
            if (stats["iterations"] + 1) % int(min(2 ** 31 - 1, ndof)) == 0:
                beta = 0.0
            gg = ff
            g = f().clone()
            h = g + beta * h
            if gsum((g * h).sum()) <= 0.0:
                h = g.clone()
        ffin = f()
        stats.update({"stopping_criterion": stop, "energy_initial": einitial, "energy_next_to_last": enext_to_last, "energy_final": ecurrent,
                      "fnorm_final": gsum((ffin * ffin).sum()) ** 0.5, "fmax_final": gmax(ffin.abs().max()), "alpha_final": alpha_final,
                      "max_atom_move": alpha_final * gmax(ffin.abs().max())})
        return stats

    # ------------------------------------------------------------------ CUDA-graph replay of the step (launch-bound sizes)
    def capture_step(self, nh: bool = False, eflag: bool = False):
        """Capture one MD step (NVE, or Nose-Hoover when nh=True) into a CUDA graph.  A step of a few thousand atoms is
        ~10 short kernels and is bound by launch latency, not by the GPU; replaying the captured graph issues them
        with one call.  Valid until the next reneighbor() (buffers and list are re-created there).  With several ranks the
        halo exchange (grouped ncclSend / ncclRecv issued by the library) and the Nose-Hoover all-reduce are captured too:
        every rank replays its own graph and NCCL pairs the sends and receives as in eager mode."""
        fn = (lambda: self.step_nh(eflag=eflag)) if nh else (lambda: self.step(eflag=eflag))
        side = torch.cuda.Stream(device=self.dev)
        side.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(side):
            for _ in range(2):              # warm-up on the capture stream: every buffer reaches its final size
                fn()
        torch.cuda.current_stream(self.dev).wait_stream(side)
        torch.cuda.synchronize(self.dev)
        self.pair.stats()                   # collects counters: raises if the tile overflowed during warm-up
        g = torch.cuda.CUDAGraph()
        n0 = self.nsteps
        with torch.cuda.graph(g, stream=side):
            fn()                            # recorded, not executed
        self.nsteps = n0
        self._graph = g
        return g

    def replay(self, nsteps: int = 1):
        for _ in range(nsteps):
            self._graph.replay()
        self.nsteps += nsteps

    # ------------------------------------------------------------------ fix nvt / fix npt (device-resident Nose-Hoover)
    def fix_nh(self, t_start, t_stop, t_damp, p_flag=(0, 0, 0), p_start=(0.0, 0.0, 0.0), p_stop=(0.0, 0.0, 0.0),
               p_damp=(1.0, 1.0, 1.0), tchain=3, pchain=3, nsteps_ramp=0):
        """`fix nvt temp t_start t_stop t_damp` (p_flag all 0) or `fix npt ... x|y|z p_start p_stop p_damp`
        (in.st_test:30-37 couples y only: p_flag=(0,1,0)).  Call after reneighbor() + compute(vflag=True)."""
        cfg = capi.NhConfig()
        cfg.tstat, cfg.pstat = 1, int(any(p_flag))
        cfg.t_start, cfg.t_stop, cfg.t_damp = t_start, t_stop, t_damp
        for d in range(3):
            cfg.p_flag[d], cfg.p_start[d], cfg.p_stop[d], cfg.p_damp[d] = int(p_flag[d]), p_start[d], p_stop[d], p_damp[d]
        cfg.tchain, cfg.pchain, cfg.mtk = tchain, pchain, 1
        cfg.dt, cfg.mass = self.dt, self.mass
        ntot = torch.tensor([float(self.nlocal)], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(ntot, group=self.group)
        cfg.natoms_total = float(ntot)
        cfg.tdof = 0.0
        cfg.nsteps_ramp = nsteps_ramp
        lo, hi = np.zeros(3), self.box.astype(np.float64).copy()
        for d in range(3):
            if self.shrink_wrap[d]:
                lo[d], hi[d] = self.wrap_lo[d], self.wrap_hi[d]
        err = C.create_string_buffer(256)
        h = C.c_void_p(None)
        rc = self.L.annp_b200_nh_create(C.byref(cfg), lo.ctypes.data_as(capi.c_double_p), hi.ctypes.data_as(capi.c_double_p),
                                        self.dev.index if self.dev.index is not None else -1, C.byref(h), err, 256)
        if rc != 0:
            raise capi.AnnpError(rc, err.value.decode())
        self.nh = h
        self.nh_pstat = bool(cfg.pstat)
        self.red12 = torch.zeros(12, dtype=torch.float64, device=self.dev)
        if self.x is None:
            self.reneighbor()
        self.compute(eflag=True, vflag=True)
        s = self._stream()
        self._nh_ck(self.L.annp_b200_nh_reduce(self.nh, self.nlocal, C.c_void_p(self.v.data_ptr()), C.c_void_p(self.engvir.data_ptr()),
                                               C.c_void_p(self.red12.data_ptr()), s))
        self._allreduce_red12()
        self._nh_ck(self.L.annp_b200_nh_setup(self.nh, C.c_void_p(self.red12.data_ptr()), s))

    def _nh_ck(self, rc):
        if rc != 0:
            raise capi.AnnpError(rc, "Nose-Hoover integrator call failed")

    def _allreduce_red12(self):
        if self.world > 1:
            self._ck(self.L.annp_b200_allreduce_sum(self.h, C.c_void_p(self.red12.data_ptr()), 12, self._stream()))

    def _nh_initial(self):
        s = self._stream()
        shift = C.c_void_p(self.send_shift.data_ptr()) if self.nsend > 0 else None
        self._nh_ck(self.L.annp_b200_nh_initial(self.nh, self.nlocal, C.c_void_p(self.x.data_ptr()), C.c_void_p(self.v.data_ptr()),
                                                C.c_void_p(self.f.data_ptr()), self.nsend, shift, s))

    def _nh_final(self, eflag):
        s = self._stream()
        self.compute(eflag=eflag, vflag=True)
        self._nh_ck(self.L.annp_b200_nh_final_kick(self.nh, self.nlocal, C.c_void_p(self.v.data_ptr()), C.c_void_p(self.f.data_ptr()),
                                                   C.c_void_p(self.engvir.data_ptr()), C.c_void_p(self.red12.data_ptr()), s))
        self._allreduce_red12()
        self._nh_ck(self.L.annp_b200_nh_final_scale(self.nh, self.nlocal, C.c_void_p(self.v.data_ptr()), C.c_void_p(self.red12.data_ptr()), s))
        self.nsteps += 1

    def step_nh(self, eflag=False):
        """One step of FixNH::initial_integrate, force, FixNH::final_integrate."""
        self._nh_initial()
        self._nh_final(eflag)

    def nh_state(self):
        st = capi.NhState()
        self._nh_ck(self.L.annp_b200_nh_get_state(self.nh, C.byref(st), self._stream()))
        return st

    def _nh_box_corners(self):
        st = self.nh_state()
        return np.array(st.boxlo[:]), np.array(st.boxhi[:])

    def sync_box_from_nh(self):
        """After npt steps the box lives in the thermostat state: refresh the host copy (needed to re-neighbour)."""
        st = self.nh_state()
        lo, hi = np.array(st.boxlo[:]), np.array(st.boxhi[:])
        self.box_origin = lo
        self.box = hi - lo
        self.lo = lo + np.array([self.box[d] * self.coords[d] / self.grid[d] for d in range(3)])
        self.hi = lo + np.array([self.box[d] * (self.coords[d] + 1) / self.grid[d] for d in range(3)])
        return st

    def run_nh(self, nsteps: int, check_every: int = 5, thermo_every: int = 0):
        """`run nsteps` under fix nvt / npt with the deck's `neigh_modify every 5 delay 5 check yes` rule, in the order of
        LAMMPS' Verlet::run: initial_integrate, neighbour decision on the new positions, force, final_integrate.
        Returns [(step, pe, ke, extended_energy, T, (pxx, pyy, pzz), box[3])] at the thermo steps; the pressures are
        LAMMPS' thermo values, (m v(x)v + virial) / V from the velocities at the END of the step."""
        x_ref = self.x[: self.nlocal].clone()
        box_ref = self._nh_box_corners()
        out, self.rebuilds = [], 0
        nktv2p = 1.6021765e6
        for n in range(1, nsteps + 1):
            want = thermo_every > 0 and (n % thermo_every == 0 or n == nsteps)
            self._nh_initial()
            if check_every > 0 and n % check_every == 0:
                moved = self._moved_sq(x_ref)
                self.check_device_flags()
                trigger = 0.5 * self.skin
                if self.nh_pstat:
                    # Neighbor::check_distance with a changing box: the skin is reduced by the displacement of the two
                    # box corners since the last build, delta = (skin - (d_lo + d_hi)) / 2
                    st = self.nh_state()
                    lo_now, hi_now = np.array(st.boxlo[:]), np.array(st.boxhi[:])
                    trigger = 0.5 * (self.skin - (np.linalg.norm(lo_now - box_ref[0]) + np.linalg.norm(hi_now - box_ref[1])))
                if trigger <= 0.0 or float(moved) > trigger ** 2:
                    if self.nh_pstat:
                        self.sync_box_from_nh()
                    self.reneighbor()
                    x_ref = self.x[: self.nlocal].clone()
                    box_ref = self._nh_box_corners()
                    self.rebuilds += 1
            self._nh_final(want)
            if want:
                st = self.nh_state()
                pe = self.engvir[:1].clone()
                if self.world > 1:
                    import torch.distributed as dist
                    dist.all_reduce(pe, group=self.group)
                ke = 0.5 * (st.ke_tensor[0] + st.ke_tensor[1] + st.ke_tensor[2])
                box = tuple(st.boxhi[d] - st.boxlo[d] for d in range(3))
                vol = box[0] * box[1] * box[2]
                p = tuple((st.ke_tensor[d] + st.virial[d]) / vol * nktv2p for d in range(3))
                out.append((self.nsteps, float(pe), ke, st.extended_energy, st.t_current, p, box))
        return out

    def thermo(self):
        """(pe, ke) of the whole system in eV after a step(eflag=True)."""
        t = torch.stack([self.engvir[0], self.ke[0]])
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, group=self.group)
        pe, ke = t.cpu().tolist()
        return pe, ke

    def max_displacement_since(self, x_ref: torch.Tensor) -> float:
        return self._moved_sq(x_ref) ** 0.5

    def _moved_sq(self, x_ref: torch.Tensor) -> float:
        """max |x - x_ref|^2 over the atoms of ALL ranks (Neighbor::check_distance): one fused kernel, the maximum over
        ranks, and the one host read the re-neighbouring decision needs."""
        self._ck(self.L.annp_b200_max_displacement_sq(self.h, self.nlocal, C.c_void_p(self.x.data_ptr()), C.c_void_p(x_ref.data_ptr()),
                                                      C.c_void_p(self.disp2.data_ptr()), self._stream()))
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(self.disp2, op=dist.ReduceOp.MAX, group=self.group)
        return float(self.disp2)

    def check_device_flags(self):
        """Raises if a device-resident step since the last call flagged an error on the device (an atom beyond the largest
        neighbour tile, a pair force outside the fixed-point range): annp_b200_get_stats reads and clears the sticky flags."""
        self.pair.stats()

    def run(self, nsteps: int, check_every: int = 5, thermo_every: int = 0, log=None):
        """`run nsteps` with `neigh_modify every 5 delay 5 check yes` semantics (in.st_test:10-11):
        every `check_every` steps re-neighbour if any atom moved more than skin/2 since the last build."""
        if self.x is None:
            self.reneighbor()
            self.compute(eflag=True)
        x_ref = self.x[: self.nlocal].clone()
        rebuilds = 0
        out = []
        for n in range(1, nsteps + 1):
            want_e = thermo_every > 0 and (n % thermo_every == 0 or n == nsteps)
            s = self._stream()
            self._ck(self.L.annp_b200_nve_initial(self.h, self.nlocal, self.dt, self.mass, C.c_void_p(self.x.data_ptr()),
                                                  C.c_void_p(self.v.data_ptr()), C.c_void_p(self.f.data_ptr()), s))
            if check_every > 0 and n % check_every == 0:
                moved = self._moved_sq(x_ref)
                self.check_device_flags()        # the host is synchronised here anyway: read the device's sticky error flags
                if moved > (0.5 * self.skin) ** 2:
                    self.reneighbor()
                    x_ref = self.x[: self.nlocal].clone()
                    rebuilds += 1
            self.compute(eflag=want_e)
            ke = C.c_void_p(self.ke.data_ptr()) if want_e else None
            self._ck(self.L.annp_b200_nve_final(self.h, self.nlocal, self.dt, self.mass, C.c_void_p(self.v.data_ptr()),
                                                C.c_void_p(self.f.data_ptr()), ke, s))
            self.nsteps += 1
            if want_e:
                pe, ke_v = self.thermo()
                out.append((self.nsteps, pe, ke_v))
                if log:
                    log(self.nsteps, pe, ke_v)
        self.rebuilds = rebuilds
        return out

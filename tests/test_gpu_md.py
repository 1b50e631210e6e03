"""Device-resident MD driver on one GPU: halo pack/unpack, NVE energy conservation, smoke()."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_smoke_entry_point():
    import __graft_entry__
    __graft_entry__.smoke()


def test_device_md_matches_host_path_and_conserves_energy(fe_pot_file):
    import torch
    from meng_zhang_b200 import lattice as L
    from meng_zhang_b200.md import DomainMD
    from meng_zhang_b200.pair import PairANNPGPU
    pair = PairANNPGPU(ntypes=1)
    pair.settings([])
    pair.coeff(["*", "*", fe_pot_file, "Fe"])
    pair.init_style()
    x, box = L.bcc(6, 6, 6)
    x = L.perturb(x, 0.03, 1)
    md = DomainMD(pair, x, box, dt=0.001)
    md.set_velocities(300.0, 12345)
    md.reneighbor()
    md.compute(eflag=True)
    torch.cuda.synchronize()
    # same configuration through the host path with a host-built list and explicit ghosts
    cfg = L.build_config(x, box, 6.5)
    pair2 = PairANNPGPU(ntypes=1)
    pair2.settings([])
    pair2.coeff(["*", "*", fe_pot_file, "Fe"])
    pair2.init_style()
    f_host = cfg.fold(pair2.compute(3, 0, cfg, ago=0))
    f_dev = md.f[: md.nlocal].cpu().numpy()
    assert md.nghost == cfg.nghost
    assert np.abs(f_dev - f_host).max() < 1e-11
    assert abs(float(md.engvir[0]) - pair2.eng_vdwl) < 1e-8
    # 200 steps of NVE at 300 K, dt = 1 fs: total energy drift per atom
    log = md.run(200, check_every=5, thermo_every=20)
    etot = np.array([pe + ke for _, pe, ke in log])
    fluct = np.abs(etot - etot[0]).max() / md.nlocal
    assert fluct < 2e-5, fluct                     # eV/atom: velocity-Verlet fluctuation (~(w dt)^2/8 of KE = 0.039 eV)
    secular = abs(etot[len(etot) // 2:].mean() - etot[: len(etot) // 2].mean()) / md.nlocal
    assert secular < 5e-6, secular
    ke_mean = np.mean([ke for _, _, ke in log])
    assert 0.2 < ke_mean / (1.5 * md.nlocal * 8.617343e-5 * 300.0) < 1.0   # equipartition: T settles near 150 K
    pair.clear()
    pair2.clear()


@pytest.mark.parametrize("nh", [False, True], ids=["nve", "npt"])
def test_cuda_graph_replay_equals_eager_steps(nh, fe_pot_file):
    """The captured step replays the same kernels on the same buffers: bit-identical trajectories."""
    import torch
    from meng_zhang_b200 import lattice as L
    from meng_zhang_b200.md import DomainMD
    from meng_zhang_b200.pair import PairANNPGPU

    def fresh():
        pair = PairANNPGPU(ntypes=1)
        pair.settings([])
        pair.coeff(["*", "*", fe_pot_file, "Fe"])
        pair.init_style()
        x, box = L.bcc(5, 5, 5)
        md = DomainMD(pair, L.perturb(x, 0.03, 1), box, dt=0.001)
        md.set_velocities(300.0, 12345)
        md.reneighbor()
        md.compute(eflag=True)
        if nh:
            md.fix_nh(300.0, 300.0, 0.1, p_flag=(0, 1, 0), p_start=(0.0,) * 3, p_stop=(0.0,) * 3, p_damp=(1.0,) * 3)
        return pair, md

    pa, a = fresh()
    for _ in range(22):                     # 2 warm-up + 20 replayed steps on the other side (capturing does not execute)
        a.step_nh() if nh else a.step()
    pb, b = fresh()
    b.capture_step(nh=nh)
    b.replay(20)
    torch.cuda.synchronize()
    assert torch.equal(a.x[: a.nlocal], b.x[: b.nlocal]) and torch.equal(a.v, b.v)
    if nh:
        sa, sb = a.nh_state(), b.nh_state()
        assert sa.step == sb.step == 22 and sa.t_current == sb.t_current and list(sa.boxhi[:]) == list(sb.boxhi[:])
    pa.clear()
    pb.clear()


def test_device_ghost_map_equals_the_host_twin(fe_pot_file):
    """annp_b200_send_lists_count / _fill against md.build_send_lists (numpy): same entries in the same order for one
    periodic rank, for a rank of a 2x2x2 grid and for a rank next to a free surface; ragged atom counts (not a multiple
    of the 256-atom tile), an empty rank."""
    import ctypes as C
    import torch
    from meng_zhang_b200 import capi
    from meng_zhang_b200.md import build_send_lists, rank_coords, send_slots
    from meng_zhang_b200.pair import PairANNPGPU
    pair = PairANNPGPU(ntypes=1)
    pair.settings([])
    pair.coeff(["*", "*", fe_pot_file, "Fe"])
    pair.init_style()
    Lb = capi.lib()
    rng = np.random.default_rng(12)
    box = np.array([60.0, 52.0, 47.0])
    cases = [((1, 1, 1), 0, (True, True, True), 20011), ((2, 2, 2), 5, (True, True, True), 7777), ((2, 2, 1), 1, (True, False, True), 3000),
             ((2, 1, 1), 1, (False, True, True), 255), ((1, 1, 1), 0, (True, True, True), 0)]
    for grid, rank, periodic, n in cases:
        coords = rank_coords(rank, grid)
        lo = np.array([box[d] * coords[d] / grid[d] for d in range(3)])
        hi = np.array([box[d] * (coords[d] + 1) / grid[d] for d in range(3)])
        x = lo + rng.random((n, 3)) * (hi - lo)
        if n > 10:
            x[3] = lo                         # exactly on the lower corner: '<' on the low side, '>=' on the high side
            x[4] = hi - 8.5
        idx, shift, counts = build_send_lists(x, lo, hi, box, grid, coords, 8.5, periodic)
        dirs, shifts, dests = send_slots(lo, hi, box, grid, coords, 8.5, periodic)
        xd = torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float64, device="cuda")
        slot_counts = np.zeros(max(len(dests), 1), dtype=np.int32)
        rc = Lb.annp_b200_send_lists_count(pair.handle, n, C.c_void_p(xd.data_ptr()), lo.ctypes.data_as(capi.c_double_p),
                                           hi.ctypes.data_as(capi.c_double_p), 8.5, len(dests), dirs.ctypes.data_as(capi.c_int_p),
                                           slot_counts.ctypes.data_as(capi.c_int_p), None)
        assert rc == 0
        got_counts = np.zeros(grid[0] * grid[1] * grid[2], dtype=np.int64)
        np.add.at(got_counts, dests, slot_counts[: len(dests)])
        assert np.array_equal(got_counts, counts)
        nsend = int(got_counts.sum())
        di = torch.full((max(nsend, 1),), -1, dtype=torch.int32, device="cuda")
        ds = torch.full((max(nsend, 1), 3), np.nan, dtype=torch.float64, device="cuda")
        rc = Lb.annp_b200_send_lists_fill(pair.handle, shifts.ctypes.data_as(capi.c_double_p), C.c_void_p(di.data_ptr()), C.c_void_p(ds.data_ptr()), None)
        assert rc == 0
        torch.cuda.synchronize()
        assert np.array_equal(di.cpu().numpy()[:nsend], idx)
        assert np.array_equal(ds.cpu().numpy()[:nsend], shift.reshape(-1, 3))
    pair.clear()


def test_md_step_through_the_library_halo_is_graph_replayable(fe_pot_file):
    """One rank: forward / reverse halo, force and integration are all library calls on one stream; the captured graph
    replays bit-identically to eager stepping and the displacement check reads the same number as torch arithmetic."""
    import torch
    from meng_zhang_b200 import lattice as L
    from meng_zhang_b200.md import DomainMD
    from meng_zhang_b200.pair import PairANNPGPU

    def make():
        pair = PairANNPGPU(ntypes=1)
        pair.settings([])
        pair.coeff(["*", "*", fe_pot_file, "Fe"])
        pair.init_style()
        x, box = L.bcc(6, 6, 6)
        md = DomainMD(pair, L.perturb(x, 0.05, 1), box)
        md.set_velocities(300.0, 77)
        md.reneighbor()
        md.compute(eflag=True)
        return pair, md

    p1, m1 = make()
    p2, m2 = make()
    x0 = m1.x[: m1.nlocal].clone()
    for _ in range(6):
        m1.step()
    m2.capture_step()          # two warm-up steps on the side stream
    m2.replay(4)
    torch.cuda.synchronize()
    assert torch.equal(m1.x, m2.x) and torch.equal(m1.v, m2.v)
    want = float((m1.x[: m1.nlocal] - x0).square().sum(dim=1).max())
    assert m1._moved_sq(x0) == want
    p1.clear(); p2.clear()

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

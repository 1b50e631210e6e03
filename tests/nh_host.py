"""Host (numpy) twin of csrc/annp_nh.cu: LAMMPS FixNH (nvt / npt, orthogonal box, aniso coupling, MTK, one sub-cycle)
written line by line after the published algorithm.  TEST INFRASTRUCTURE: tests/test_gpu_nh.py drives it with the
forces the GPU produced and compares positions, velocities, box and chain variables step by step."""
import math

import numpy as np

BOLTZ, NKTV2P, MVV2E = 8.617343e-5, 1.6021765e6, 1.0364269e-4
FTM2V = 1.0 / MVV2E


class HostNH:
    def __init__(self, natoms, box, dt, mass, t_start, t_stop, t_damp, p_flag=(0, 0, 0), p_start=(0, 0, 0), p_stop=(0, 0, 0),
                 p_damp=(1, 1, 1), tchain=3, pchain=3, nsteps_ramp=0):
        self.natoms, self.dt, self.mass = float(natoms), dt, mass
        self.dthalf, self.dt4, self.dt8 = 0.5 * dt, 0.25 * dt, 0.125 * dt
        self.t_start, self.t_stop, self.t_freq = t_start, t_stop, 1.0 / t_damp
        self.p_flag = [int(f) for f in p_flag]
        self.pstat = any(self.p_flag)
        self.pdim = sum(self.p_flag)
        self.p_start, self.p_stop = list(p_start), list(p_stop)
        self.p_freq = [1.0 / p_damp[i] if self.p_flag[i] else 0.0 for i in range(3)]
        self.p_freq_max = max(self.p_freq)
        self.mt, self.mp = tchain, (pchain if self.pstat else 0)
        self.tdof = 3.0 * natoms - 3.0
        self.nsteps_ramp = nsteps_ramp
        self.boxlo, self.boxhi = np.zeros(3), np.array(box, dtype=float)
        self.fixed = 0.5 * (self.boxlo + self.boxhi)
        self.vol0 = float(np.prod(self.boxhi - self.boxlo))
        self.eta, self.eta_dot, self.eta_dd, self.eta_mass = np.zeros(tchain), np.zeros(tchain + 1), np.zeros(tchain), np.zeros(tchain)
        n = max(self.mp, 1)
        self.etap, self.etap_dot, self.etap_dd, self.etap_mass = np.zeros(n), np.zeros(n + 1), np.zeros(n), np.zeros(n)
        self.omega, self.omega_dot, self.omega_mass = np.zeros(3), np.zeros(3), np.zeros(3)
        self.p_target, self.p_current, self.p_hydro = np.zeros(3), np.zeros(3), 0.0
        self.step = 0
        self.mvv, self.virial = np.zeros(6), np.zeros(6)

    # ---- helpers
    def volume(self):
        return float(np.prod(self.boxhi - self.boxlo))

    def targets(self):
        delta = min(1.0, self.step / self.nsteps_ramp) if self.nsteps_ramp > 0 else 0.0
        self.t_target = self.t_start + delta * (self.t_stop - self.t_start)
        self.ke_target = self.tdof * BOLTZ * self.t_target
        self.p_hydro = 0.0
        for i in range(3):
            if self.p_flag[i]:
                self.p_target[i] = self.p_start[i] + delta * (self.p_stop[i] - self.p_start[i])
                self.p_hydro += self.p_target[i]
        if self.pdim:
            self.p_hydro /= self.pdim

    def load(self, v, virial):
        m = self.mass * MVV2E
        self.mvv = m * np.array([(v[:, 0] ** 2).sum(), (v[:, 1] ** 2).sum(), (v[:, 2] ** 2).sum(),
                                 (v[:, 0] * v[:, 1]).sum(), (v[:, 0] * v[:, 2]).sum(), (v[:, 1] * v[:, 2]).sum()])
        self.virial = np.array(virial, dtype=float)

    def temperature(self):
        self.t_current = self.mvv[:3].sum() / (self.tdof * BOLTZ)

    def pressure(self):
        self.p_current = (self.mvv[:3] + self.virial[:3]) * NKTV2P / self.volume()

    def setup(self, v, virial):
        self.load(v, virial)
        self.step = 0
        self.targets()
        self.temperature()
        tf2 = self.t_freq ** 2
        self.eta_mass[0] = self.tdof * BOLTZ * self.t_target / tf2
        self.eta_mass[1:] = BOLTZ * self.t_target / tf2
        for k in range(1, self.mt):
            self.eta_dd[k] = (self.eta_mass[k - 1] * self.eta_dot[k - 1] ** 2 - BOLTZ * self.t_target) / self.eta_mass[k]
        if self.pstat:
            self.pressure()
            kt = BOLTZ * self.t_target
            for i in range(3):
                if self.p_flag[i]:
                    self.omega_mass[i] = (self.natoms + 1) * kt / self.p_freq[i] ** 2
            self.etap_mass[:] = kt / self.p_freq_max ** 2
            for k in range(1, self.mp):
                self.etap_dd[k] = (self.etap_mass[k - 1] * self.etap_dot[k - 1] ** 2 - kt) / self.etap_mass[k]

    def nhc_temp(self):
        m, e = self.mt, math.exp
        ke = self.tdof * BOLTZ * self.t_current
        tf2 = self.t_freq ** 2
        self.eta_mass[0] = self.tdof * BOLTZ * self.t_target / tf2
        self.eta_mass[1:] = BOLTZ * self.t_target / tf2
        self.eta_dd[0] = (ke - self.ke_target) / self.eta_mass[0]
        for k in range(m - 1, 0, -1):
            x = e(-self.dt8 * self.eta_dot[k + 1])
            self.eta_dot[k] *= x
            self.eta_dot[k] += self.eta_dd[k] * self.dt4
            self.eta_dot[k] *= x
        x = e(-self.dt8 * self.eta_dot[1])
        self.eta_dot[0] *= x
        self.eta_dot[0] += self.eta_dd[0] * self.dt4
        self.eta_dot[0] *= x
        factor = e(-self.dthalf * self.eta_dot[0])
        self.t_current *= factor * factor
        self.mvv *= factor * factor
        ke = self.tdof * BOLTZ * self.t_current
        self.eta_dd[0] = (ke - self.ke_target) / self.eta_mass[0]
        self.eta[:m] += self.dthalf * self.eta_dot[:m]
        self.eta_dot[0] *= x
        self.eta_dot[0] += self.eta_dd[0] * self.dt4
        self.eta_dot[0] *= x
        for k in range(1, m):
            x = e(-self.dt8 * self.eta_dot[k + 1])
            self.eta_dot[k] *= x
            self.eta_dd[k] = (self.eta_mass[k - 1] * self.eta_dot[k - 1] ** 2 - BOLTZ * self.t_target) / self.eta_mass[k]
            self.eta_dot[k] += self.eta_dd[k] * self.dt4
            self.eta_dot[k] *= x
        return factor

    def nhc_press(self):
        m, e = self.mp, math.exp
        kt = BOLTZ * self.t_target
        for i in range(3):
            if self.p_flag[i]:
                self.omega_mass[i] = (self.natoms + 1) * kt / self.p_freq[i] ** 2
        if m == 0:
            return
        self.etap_mass[:] = kt / self.p_freq_max ** 2
        for k in range(1, m):
            self.etap_dd[k] = (self.etap_mass[k - 1] * self.etap_dot[k - 1] ** 2 - kt) / self.etap_mass[k]
        ke = sum(self.omega_mass[i] * self.omega_dot[i] ** 2 for i in range(3) if self.p_flag[i])
        lkt = self.pdim * kt
        self.etap_dd[0] = (ke - lkt) / self.etap_mass[0]
        for k in range(m - 1, 0, -1):
            x = e(-self.dt8 * self.etap_dot[k + 1])
            self.etap_dot[k] *= x
            self.etap_dot[k] += self.etap_dd[k] * self.dt4
            self.etap_dot[k] *= x
        x = e(-self.dt8 * self.etap_dot[1])
        self.etap_dot[0] *= x
        self.etap_dot[0] += self.etap_dd[0] * self.dt4
        self.etap_dot[0] *= x
        self.etap[:m] += self.dthalf * self.etap_dot[:m]
        fe = e(-self.dthalf * self.etap_dot[0])
        for i in range(3):
            if self.p_flag[i]:
                self.omega_dot[i] *= fe
        ke = sum(self.omega_mass[i] * self.omega_dot[i] ** 2 for i in range(3) if self.p_flag[i])
        self.etap_dd[0] = (ke - lkt) / self.etap_mass[0]
        self.etap_dot[0] *= x
        self.etap_dot[0] += self.etap_dd[0] * self.dt4
        self.etap_dot[0] *= x
        for k in range(1, m):
            x = e(-self.dt8 * self.etap_dot[k + 1])
            self.etap_dot[k] *= x
            self.etap_dd[k] = (self.etap_mass[k - 1] * self.etap_dot[k - 1] ** 2 - kt) / self.etap_mass[k]
            self.etap_dot[k] += self.etap_dd[k] * self.dt4
            self.etap_dot[k] *= x

    def omega_dot_update(self):
        vol = self.volume()
        t1 = sum(self.mvv[i] for i in range(3) if self.p_flag[i]) / (self.pdim * self.natoms)
        for i in range(3):
            if self.p_flag[i]:
                f = (self.p_current[i] - self.p_hydro) * vol / (self.omega_mass[i] * NKTV2P) + t1 / self.omega_mass[i]
                self.omega_dot[i] += f * self.dthalf
        t2 = sum(self.omega_dot[i] for i in range(3) if self.p_flag[i]) / (self.pdim * self.natoms)
        self.factor_v = np.array([math.exp(-self.dt4 * (self.omega_dot[i] + t2)) for i in range(3)])
        self.dil = np.array([math.exp(self.dthalf * self.omega_dot[i]) if self.p_flag[i] else 1.0 for i in range(3)])

    # ---- one step: forces through callbacks
    def initial(self, x, v, f):
        self.step += 1
        self.targets()
        if self.pstat and self.mp:
            self.nhc_press()
        fe = self.nhc_temp()
        fv, dil = np.ones(3), np.ones(3)
        if self.pstat:
            self.pressure()
            self.omega_dot_update()
            fv, dil = self.factor_v ** 2, self.dil
            for i in range(3):
                if self.p_flag[i]:
                    e2 = dil[i] * dil[i]
                    self.boxlo[i] = (self.boxlo[i] - self.fixed[i]) * e2 + self.fixed[i]
                    self.boxhi[i] = (self.boxhi[i] - self.fixed[i]) * e2 + self.fixed[i]
        dtfm = 0.5 * self.dt * FTM2V / self.mass
        v = v * fe
        v = v * fv
        v = v + dtfm * f
        x = self.fixed + dil * (x - self.fixed)
        x = x + self.dt * v
        x = self.fixed + dil * (x - self.fixed)
        return x, v

    def final(self, v, f, virial):
        dtfm = 0.5 * self.dt * FTM2V / self.mass
        fv = self.factor_v ** 2 if self.pstat else np.ones(3)
        v = (v + dtfm * f) * fv
        self.load(v, virial)
        self.temperature()
        if self.pstat:
            self.pressure()
            self.omega_dot_update()
        fe = self.nhc_temp()
        if self.pstat and self.mp:
            self.nhc_press()
        return v * fe

    def extended_energy(self):
        kt = BOLTZ * self.t_target
        e = self.ke_target * self.eta[0] + 0.5 * self.eta_mass[0] * self.eta_dot[0] ** 2
        for k in range(1, self.mt):
            e += kt * self.eta[k] + 0.5 * self.eta_mass[k] * self.eta_dot[k] ** 2
        if self.pstat:
            for i in range(3):
                if self.p_flag[i]:
                    e += 0.5 * self.omega_dot[i] ** 2 * self.omega_mass[i] + self.p_hydro * (self.volume() - self.vol0) / (self.pdim * NKTV2P)
            if self.mp:
                e += self.pdim * kt * self.etap[0] + 0.5 * self.etap_mass[0] * self.etap_dot[0] ** 2
                for k in range(1, self.mp):
                    e += kt * self.etap[k] + 0.5 * self.etap_mass[k] * self.etap_dot[k] ** 2
        return e

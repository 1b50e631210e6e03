// annp_nh.cu -- device-resident Nose-Hoover chain thermostat / barostat (LAMMPS `fix nvt`, `fix npt` with independent
// x / y / z coupling, orthogonal box) for the stand-alone MD loop: BASELINE configs 2 (NVT 300 K) and 3 (NPT, the deck
// couples y only: in.st_test:30-37).
//
// Equations and operator splitting follow LAMMPS' FixNH (MTK barostat with chains, nc_tchain = nc_pchain = 1):
//   initial_integrate: nhc_press_integrate, nhc_temp_integrate, pressure, nh_omega_dot, nh_v_press, nve_v,
//                      remap(dt/2), nve_x, remap(dt/2)
//   final_integrate:   nve_v, nh_v_press, temperature + pressure, nh_omega_dot, nhc_temp_integrate, nhc_press_integrate
// What is B200-specific: the chain variables live in ONE small device struct and are advanced by a single-thread kernel
// that reads the (deterministically reduced) kinetic and virial tensors from device memory, so a step never
// synchronises with the host; the per-atom work of each half step is one fused streaming kernel (scale, kick, drift,
// box dilation) reading its factors from that struct.  LAMMPS itself is not in this image: the module is pinned by a
// line-by-line host twin (tests/nh_host.py) and by conservation of the extended energy.
#include "annp_device.cuh"

#include <cstring>
#include <new>
#include <string>

namespace {

constexpr int kMaxChain = 8;
constexpr int kRedThreads = 256;
constexpr int kRedSlice = 8192;          // atoms per reduction block: fixed slices -> bit-reproducible sums

struct NhState {
  // ---- constants
  int tstat, pstat, mtchain, mpchain, mtk, pdim;
  int p_flag[3];
  double dt, dthalf, dt4, dt8, dto;
  double boltz, nktv2p, mvv2e, ftm2v;
  double t_start, t_stop, t_freq;
  double p_start[3], p_stop[3], p_freq[3], p_freq_max;
  double tdof, natoms, mass;
  double fixedpoint[3];
  long long nsteps_ramp;
  // ---- dynamic
  long long step;
  double t_target, ke_target, t_current;
  double mvv[6];                          // sum m v v (mvv2e applied), xx yy zz xy xz yz, whole system
  double virial[6];                       // pair virial of the last force evaluation, whole system
  double eta[kMaxChain], eta_dot[kMaxChain + 1], eta_dotdot[kMaxChain], eta_mass[kMaxChain];
  double etap[kMaxChain], etap_dot[kMaxChain + 1], etap_dotdot[kMaxChain], etap_mass[kMaxChain];
  double omega[3], omega_dot[3], omega_mass[3];
  double p_target[3], p_hydro, p_current[3];
  double mtk_term1, mtk_term2;
  double boxlo[3], boxhi[3], vol0;
  // ---- factors consumed by the per-atom kernels
  double factor_eta;                      // thermostat velocity scale of the half step just integrated
  double factor_v[3];                     // nh_v_press: exp(-dt4 (omega_dot + mtk_term2)), applied twice
  double dilation[3];                     // remap: exp(dto omega_dot), applied twice per step
};

__device__ double nh_volume(const NhState &s) {
  return (s.boxhi[0] - s.boxlo[0]) * (s.boxhi[1] - s.boxlo[1]) * (s.boxhi[2] - s.boxlo[2]);
}

__device__ void nh_targets(NhState &s) {
  double delta = s.nsteps_ramp > 0 ? (double) s.step / (double) s.nsteps_ramp : 0.0;
  if (delta > 1.0) delta = 1.0;
  s.t_target = s.t_start + delta * (s.t_stop - s.t_start);
  s.ke_target = s.tdof * s.boltz * s.t_target;
  s.p_hydro = 0.0;
  for (int i = 0; i < 3; i++)
    if (s.p_flag[i]) { s.p_target[i] = s.p_start[i] + delta * (s.p_stop[i] - s.p_start[i]); s.p_hydro += s.p_target[i]; }
  if (s.pdim > 0) s.p_hydro /= s.pdim;
}

__device__ void nh_temperature(NhState &s) { s.t_current = (s.mvv[0] + s.mvv[1] + s.mvv[2]) / (s.tdof * s.boltz); }

// pressure tensor diagonal: (kinetic + virial) / V * nktv2p          (compute pressure, pcouple none)
__device__ void nh_pressure(NhState &s) {
  const double inv = s.nktv2p / nh_volume(s);
  for (int i = 0; i < 3; i++) s.p_current[i] = (s.mvv[i] + s.virial[i]) * inv;
}

// FixNH::nhc_temp_integrate
__device__ void nh_temp_integrate(NhState &s) {
  const int m = s.mtchain;
  double kecurrent = s.tdof * s.boltz * s.t_current;
  s.eta_mass[0] = s.tdof * s.boltz * s.t_target / (s.t_freq * s.t_freq);
  for (int ich = 1; ich < m; ich++) s.eta_mass[ich] = s.boltz * s.t_target / (s.t_freq * s.t_freq);
  s.eta_dotdot[0] = s.eta_mass[0] > 0.0 ? (kecurrent - s.ke_target) / s.eta_mass[0] : 0.0;
  double expfac;
  for (int ich = m - 1; ich > 0; ich--) {
    expfac = exp(-s.dt8 * s.eta_dot[ich + 1]);
    s.eta_dot[ich] *= expfac;
    s.eta_dot[ich] += s.eta_dotdot[ich] * s.dt4;
    s.eta_dot[ich] *= expfac;
  }
  expfac = exp(-s.dt8 * s.eta_dot[1]);
  s.eta_dot[0] *= expfac;
  s.eta_dot[0] += s.eta_dotdot[0] * s.dt4;
  s.eta_dot[0] *= expfac;
  s.factor_eta = exp(-s.dthalf * s.eta_dot[0]);
  // nh_v_temp is applied by the per-atom kernel; the tensors follow analytically
  const double f2 = s.factor_eta * s.factor_eta;
  s.t_current *= f2;
  for (int k = 0; k < 6; k++) s.mvv[k] *= f2;
  kecurrent = s.tdof * s.boltz * s.t_current;
  s.eta_dotdot[0] = s.eta_mass[0] > 0.0 ? (kecurrent - s.ke_target) / s.eta_mass[0] : 0.0;
  for (int ich = 0; ich < m; ich++) s.eta[ich] += s.dthalf * s.eta_dot[ich];
  s.eta_dot[0] *= expfac;
  s.eta_dot[0] += s.eta_dotdot[0] * s.dt4;
  s.eta_dot[0] *= expfac;
  for (int ich = 1; ich < m; ich++) {
    expfac = exp(-s.dt8 * s.eta_dot[ich + 1]);
    s.eta_dot[ich] *= expfac;
    s.eta_dotdot[ich] = (s.eta_mass[ich - 1] * s.eta_dot[ich - 1] * s.eta_dot[ich - 1] - s.boltz * s.t_target) / s.eta_mass[ich];
    s.eta_dot[ich] += s.eta_dotdot[ich] * s.dt4;
    s.eta_dot[ich] *= expfac;
  }
}

// FixNH::nhc_press_integrate
__device__ void nh_press_integrate(NhState &s) {
  const int m = s.mpchain;
  const double kt = s.boltz * s.t_target;
  const double nkt = (s.natoms + 1.0) * kt;
  for (int i = 0; i < 3; i++)
    if (s.p_flag[i]) s.omega_mass[i] = nkt / (s.p_freq[i] * s.p_freq[i]);
  if (m == 0) return;
  for (int ich = 0; ich < m; ich++) s.etap_mass[ich] = kt / (s.p_freq_max * s.p_freq_max);
  for (int ich = 1; ich < m; ich++)
    s.etap_dotdot[ich] = (s.etap_mass[ich - 1] * s.etap_dot[ich - 1] * s.etap_dot[ich - 1] - kt) / s.etap_mass[ich];
  double kecurrent = 0.0;
  int pdof = 0;
  for (int i = 0; i < 3; i++)
    if (s.p_flag[i]) { kecurrent += s.omega_mass[i] * s.omega_dot[i] * s.omega_dot[i]; pdof++; }
  const double lkt_press = pdof * kt;                       // pstyle ANISO
  s.etap_dotdot[0] = (kecurrent - lkt_press) / s.etap_mass[0];
  double expfac;
  for (int ich = m - 1; ich > 0; ich--) {
    expfac = exp(-s.dt8 * s.etap_dot[ich + 1]);
    s.etap_dot[ich] *= expfac;
    s.etap_dot[ich] += s.etap_dotdot[ich] * s.dt4;
    s.etap_dot[ich] *= expfac;
  }
  expfac = exp(-s.dt8 * s.etap_dot[1]);
  s.etap_dot[0] *= expfac;
  s.etap_dot[0] += s.etap_dotdot[0] * s.dt4;
  s.etap_dot[0] *= expfac;
  for (int ich = 0; ich < m; ich++) s.etap[ich] += s.dthalf * s.etap_dot[ich];
  const double factor_etap = exp(-s.dthalf * s.etap_dot[0]);
  for (int i = 0; i < 3; i++)
    if (s.p_flag[i]) s.omega_dot[i] *= factor_etap;
  kecurrent = 0.0;
  for (int i = 0; i < 3; i++)
    if (s.p_flag[i]) kecurrent += s.omega_mass[i] * s.omega_dot[i] * s.omega_dot[i];
  s.etap_dotdot[0] = (kecurrent - lkt_press) / s.etap_mass[0];
  s.etap_dot[0] *= expfac;
  s.etap_dot[0] += s.etap_dotdot[0] * s.dt4;
  s.etap_dot[0] *= expfac;
  for (int ich = 1; ich < m; ich++) {
    expfac = exp(-s.dt8 * s.etap_dot[ich + 1]);
    s.etap_dot[ich] *= expfac;
    s.etap_dotdot[ich] = (s.etap_mass[ich - 1] * s.etap_dot[ich - 1] * s.etap_dot[ich - 1] - kt) / s.etap_mass[ich];
    s.etap_dot[ich] += s.etap_dotdot[ich] * s.dt4;
    s.etap_dot[ich] *= expfac;
  }
}

// FixNH::nh_omega_dot (orthogonal box, no deviatoric term) + the factors of nh_v_press and remap
__device__ void nh_omega_dot(NhState &s) {
  const double volume = nh_volume(s);
  s.mtk_term1 = 0.0;
  if (s.mtk && s.pdim > 0) {
    for (int i = 0; i < 3; i++)
      if (s.p_flag[i]) s.mtk_term1 += s.mvv[i];
    s.mtk_term1 /= s.pdim * s.natoms;
  }
  for (int i = 0; i < 3; i++)
    if (s.p_flag[i]) {
      const double f_omega = (s.p_current[i] - s.p_hydro) * volume / (s.omega_mass[i] * s.nktv2p) + s.mtk_term1 / s.omega_mass[i];
      s.omega_dot[i] += f_omega * s.dthalf;
    }
  s.mtk_term2 = 0.0;
  if (s.mtk && s.pdim > 0) {
    for (int i = 0; i < 3; i++)
      if (s.p_flag[i]) s.mtk_term2 += s.omega_dot[i];
    s.mtk_term2 /= s.pdim * s.natoms;
  }
  for (int i = 0; i < 3; i++) {
    s.factor_v[i] = exp(-s.dt4 * (s.omega_dot[i] + s.mtk_term2));
    s.dilation[i] = s.p_flag[i] ? exp(s.dto * s.omega_dot[i]) : 1.0;
  }
}

// the two box half-remaps of one step (FixNH::remap, orthogonal): lo/hi move about the fixed point
__device__ void nh_remap_box(NhState &s) {
  for (int i = 0; i < 3; i++)
    if (s.p_flag[i]) {
      const double e2 = s.dilation[i] * s.dilation[i];
      s.omega[i] += 2.0 * s.dto * s.omega_dot[i];
      s.boxlo[i] = (s.boxlo[i] - s.fixedpoint[i]) * e2 + s.fixedpoint[i];
      s.boxhi[i] = (s.boxhi[i] - s.fixedpoint[i]) * e2 + s.fixedpoint[i];
    }
}

__device__ void nh_load_red(NhState &s, const double *red12) {
  for (int k = 0; k < 6; k++) { s.mvv[k] = red12[k] * s.mvv2e; s.virial[k] = red12[6 + k]; }
}

// FixNH::setup: current temperature / pressure, masses, initial chain accelerations
__global__ void k_nh_setup(NhState *st, const double *red12) {
  NhState &s = *st;
  nh_load_red(s, red12);
  s.step = 0;
  nh_targets(s);
  nh_temperature(s);
  if (s.tstat) {
    s.eta_mass[0] = s.tdof * s.boltz * s.t_target / (s.t_freq * s.t_freq);
    for (int ich = 1; ich < s.mtchain; ich++) s.eta_mass[ich] = s.boltz * s.t_target / (s.t_freq * s.t_freq);
    for (int ich = 1; ich < s.mtchain; ich++)
      s.eta_dotdot[ich] = (s.eta_mass[ich - 1] * s.eta_dot[ich - 1] * s.eta_dot[ich - 1] - s.boltz * s.t_target) / s.eta_mass[ich];
  }
  if (s.pstat) {
    nh_pressure(s);
    const double kt = s.boltz * s.t_target, nkt = (s.natoms + 1.0) * kt;
    for (int i = 0; i < 3; i++)
      if (s.p_flag[i]) s.omega_mass[i] = nkt / (s.p_freq[i] * s.p_freq[i]);
    for (int ich = 0; ich < s.mpchain; ich++) s.etap_mass[ich] = kt / (s.p_freq_max * s.p_freq_max);
    for (int ich = 1; ich < s.mpchain; ich++)
      s.etap_dotdot[ich] = (s.etap_mass[ich - 1] * s.etap_dot[ich - 1] * s.etap_dot[ich - 1] - kt) / s.etap_mass[ich];
  }
  s.factor_eta = 1.0;
  for (int i = 0; i < 3; i++) { s.factor_v[i] = 1.0; s.dilation[i] = 1.0; }
}

// scalar part of initial_integrate; the tensors in the state are current (kept by k_nh_end)
__global__ void k_nh_begin(NhState *st) {
  NhState &s = *st;
  s.step += 1;
  s.factor_eta = 1.0;
  for (int i = 0; i < 3; i++) { s.factor_v[i] = 1.0; s.dilation[i] = 1.0; }
  nh_targets(s);
  if (s.pstat && s.mpchain) nh_press_integrate(s);
  if (s.tstat) nh_temp_integrate(s);
  if (s.pstat) {
    nh_pressure(s);
    nh_omega_dot(s);
    nh_remap_box(s);
  }
}

// scalar part of final_integrate after nve_v + nh_v_press: fresh tensors from the reduction
__global__ void k_nh_end(NhState *st, const double *red12) {
  NhState &s = *st;
  nh_load_red(s, red12);
  nh_temperature(s);
  s.factor_eta = 1.0;
  if (s.pstat) {
    nh_pressure(s);
    nh_omega_dot(s);       // also refreshes factor_v / dilation for the next begin (recomputed there anyway)
  }
  if (s.tstat) nh_temp_integrate(s);
  if (s.pstat && s.mpchain) nh_press_integrate(s);
}

// initial_integrate per atom: nh_v_temp, nh_v_press (twice dt4), nve_v, remap, nve_x, remap
__global__ void k_nh_initial(const NhState *__restrict__ st, int n, double *__restrict__ x, double *__restrict__ v,
                             const double *__restrict__ f) {
  const double dt = st->dt, dtfm = 0.5 * st->dt * st->ftm2v / st->mass;
  const double fe = st->tstat ? st->factor_eta : 1.0;
  double fv[3], e[3], fp[3];
  for (int d = 0; d < 3; d++) {
    fv[d] = st->pstat ? st->factor_v[d] * st->factor_v[d] : 1.0;
    e[d] = st->pstat ? st->dilation[d] : 1.0;
    fp[d] = st->fixedpoint[d];
  }
  for (long long t = blockIdx.x * (long long) blockDim.x + threadIdx.x; t < 3LL * n; t += (long long) gridDim.x * blockDim.x) {
    const int d = (int) (t % 3);
    double vv = v[t] * fe;
    vv *= fv[d];
    vv += dtfm * f[t];
    double xx = x[t];
    xx = fp[d] + e[d] * (xx - fp[d]);
    xx += dt * vv;
    xx = fp[d] + e[d] * (xx - fp[d]);
    v[t] = vv;
    x[t] = xx;
  }
}

// final_integrate per atom, first half: nve_v, nh_v_press; partial sums of v (x) v over fixed slices
__global__ void k_nh_final_kick(const NhState *__restrict__ st, int n, double *__restrict__ v, const double *__restrict__ f,
                                double *__restrict__ partial) {
  __shared__ double sh[kRedThreads];
  const double dtfm = 0.5 * st->dt * st->ftm2v / st->mass;
  double fv[3];
  for (int d = 0; d < 3; d++) fv[d] = st->pstat ? st->factor_v[d] * st->factor_v[d] : 1.0;
  const int a0 = blockIdx.x * kRedSlice, a1 = min(n, a0 + kRedSlice);
  double acc[6] = {0, 0, 0, 0, 0, 0};
  for (int i = a0 + threadIdx.x; i < a1; i += blockDim.x) {
    double w[3];
#pragma unroll
    for (int d = 0; d < 3; d++) {
      w[d] = (v[3 * (size_t) i + d] + dtfm * f[3 * (size_t) i + d]) * fv[d];
      v[3 * (size_t) i + d] = w[d];
    }
    acc[0] += w[0] * w[0]; acc[1] += w[1] * w[1]; acc[2] += w[2] * w[2];
    acc[3] += w[0] * w[1]; acc[4] += w[0] * w[2]; acc[5] += w[1] * w[2];
  }
  for (int k = 0; k < 6; k++) {
    sh[threadIdx.x] = acc[k];
    __syncthreads();
    for (int o = kRedThreads / 2; o > 0; o >>= 1) {
      if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) partial[(size_t) blockIdx.x * 6 + k] = sh[0];
    __syncthreads();
  }
}

// velocities only (setup): same partial sums without the kick
__global__ void k_nh_vv_partial(int n, const double *__restrict__ v, double *__restrict__ partial) {
  __shared__ double sh[kRedThreads];
  const int a0 = blockIdx.x * kRedSlice, a1 = min(n, a0 + kRedSlice);
  double acc[6] = {0, 0, 0, 0, 0, 0};
  for (int i = a0 + threadIdx.x; i < a1; i += blockDim.x) {
    const double w0 = v[3 * (size_t) i], w1 = v[3 * (size_t) i + 1], w2 = v[3 * (size_t) i + 2];
    acc[0] += w0 * w0; acc[1] += w1 * w1; acc[2] += w2 * w2; acc[3] += w0 * w1; acc[4] += w0 * w2; acc[5] += w1 * w2;
  }
  for (int k = 0; k < 6; k++) {
    sh[threadIdx.x] = acc[k];
    __syncthreads();
    for (int o = kRedThreads / 2; o > 0; o >>= 1) {
      if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) partial[(size_t) blockIdx.x * 6 + k] = sh[0];
    __syncthreads();
  }
}

// red12 = [mass * sum v v (6), pair virial (6)] of this rank; serial final sum -> fixed order
__global__ void k_nh_red_final(const double *__restrict__ partial, int nblocks, double mass, const double *__restrict__ engvir,
                               double *__restrict__ red12) {
  if (threadIdx.x < 6) {
    double s = 0.0;
    for (int b = 0; b < nblocks; b++) s += partial[(size_t) b * 6 + threadIdx.x];
    red12[threadIdx.x] = mass * s;
    red12[6 + threadIdx.x] = engvir ? engvir[1 + threadIdx.x] : 0.0;
  }
}

// final_integrate per atom, second half: nh_v_temp
__global__ void k_nh_scale_v(const NhState *__restrict__ st, int n, double *__restrict__ v) {
  const double fe = st->factor_eta;
  for (long long t = blockIdx.x * (long long) blockDim.x + threadIdx.x; t < 3LL * n; t += (long long) gridDim.x * blockDim.x) v[t] *= fe;
}

// box edges of dimensions the barostat does not couple (shrink-wrapped `boundary m/s` faces follow the atoms)
__global__ void k_nh_set_box(NhState *st, double lo0, double lo1, double lo2, double hi0, double hi1, double hi2, int w0, int w1, int w2) {
  const double lo[3] = {lo0, lo1, lo2}, hi[3] = {hi0, hi1, hi2};
  const int w[3] = {w0, w1, w2};
  for (int d = 0; d < 3; d++)
    if (w[d] && !st->p_flag[d]) { st->boxlo[d] = lo[d]; st->boxhi[d] = hi[d]; }
}

// periodic image shifts are multiples of the box edge: they dilate with the box
__global__ void k_nh_scale_shift(const NhState *__restrict__ st, int nsend, double *__restrict__ shift) {
  double e2[3];
  for (int d = 0; d < 3; d++) e2[d] = st->dilation[d] * st->dilation[d];
  for (long long t = blockIdx.x * (long long) blockDim.x + threadIdx.x; t < 3LL * nsend; t += (long long) gridDim.x * blockDim.x)
    shift[t] *= e2[t % 3];
}

int grid_for(long long n, int threads) {
  long long b = (n + threads - 1) / threads;
  if (b > 148LL * 16) b = 148LL * 16;
  if (b < 1) b = 1;
  return (int) b;
}

}    // namespace

struct annp_b200_nh_s {
  int device = 0;
  NhState host;               // configuration + last fetched copy
  NhState *d_state = nullptr;
  double *d_partial = nullptr;
  size_t partial_cap = 0;
  std::string err;
};

extern "C" {

int annp_b200_nh_create(const annp_b200_nh_config *c, const double *boxlo, const double *boxhi, int device, annp_b200_nh *out,
                        char *err, int errlen) {
  auto fail = [&](int code, const char *msg) {
    if (err && errlen > 0) snprintf(err, (size_t) errlen, "%s", msg);
    return code;
  };
  if (!c || !boxlo || !boxhi || !out) return fail(ANNP_B200_EINVAL, "null argument");
  *out = nullptr;
  if (c->tchain < 1 || c->tchain > kMaxChain || c->pchain < 0 || c->pchain > kMaxChain) return fail(ANNP_B200_EINVAL, "chain length out of range");
  if (!(c->dt > 0.0) || !(c->mass > 0.0) || !(c->natoms_total >= 1.0)) return fail(ANNP_B200_EINVAL, "dt, mass and natoms_total must be positive");
  if (c->tstat && !(c->t_damp > 0.0 && c->t_start > 0.0 && c->t_stop > 0.0)) return fail(ANNP_B200_EINVAL, "thermostat needs positive temperatures and damping time");
  if (c->pstat && !c->tstat) return fail(ANNP_B200_EINVAL, "the barostat is only implemented together with the thermostat (fix npt)");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail(ANNP_B200_ENODEVICE, "no CUDA device: libannp_b200 has no CPU fallback"); }
  if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) device = 0; }
  if (device >= ndev || cudaSetDevice(device) != cudaSuccess) return fail(ANNP_B200_ENODEVICE, "bad device");
  annp_b200_nh h = new (std::nothrow) annp_b200_nh_s();
  if (!h) return fail(ANNP_B200_ENOMEM, "out of host memory");
  h->device = device;
  NhState &s = h->host;
  memset(&s, 0, sizeof(s));
  s.tstat = c->tstat != 0; s.pstat = c->pstat != 0;
  s.mtchain = c->tchain; s.mpchain = s.pstat ? c->pchain : 0; s.mtk = c->mtk != 0;
  s.dt = c->dt; s.dthalf = 0.5 * c->dt; s.dt4 = 0.25 * c->dt; s.dt8 = 0.125 * c->dt; s.dto = s.dthalf;
  s.boltz = 8.617343e-5; s.nktv2p = 1.6021765e6; s.mvv2e = 1.0364269e-4; s.ftm2v = 1.0 / 1.0364269e-4;   // units metal
  s.t_start = c->t_start; s.t_stop = c->t_stop; s.t_freq = s.tstat ? 1.0 / c->t_damp : 0.0;
  s.pdim = 0;
  for (int i = 0; i < 3; i++) {
    s.p_flag[i] = s.pstat && c->p_flag[i];
    if (s.p_flag[i]) {
      if (!(c->p_damp[i] > 0.0)) { delete h; return fail(ANNP_B200_EINVAL, "barostat damping time must be positive"); }
      s.p_start[i] = c->p_start[i]; s.p_stop[i] = c->p_stop[i]; s.p_freq[i] = 1.0 / c->p_damp[i];
      if (s.p_freq[i] > s.p_freq_max) s.p_freq_max = s.p_freq[i];
      s.pdim++;
    }
  }
  if (s.pstat && s.pdim == 0) { delete h; return fail(ANNP_B200_EINVAL, "barostat without a coupled dimension"); }
  s.tdof = c->tdof > 0.0 ? c->tdof : 3.0 * c->natoms_total - 3.0;
  s.natoms = c->natoms_total; s.mass = c->mass; s.nsteps_ramp = c->nsteps_ramp;
  for (int i = 0; i < 3; i++) { s.boxlo[i] = boxlo[i]; s.boxhi[i] = boxhi[i]; s.fixedpoint[i] = 0.5 * (boxlo[i] + boxhi[i]); }
  s.vol0 = (boxhi[0] - boxlo[0]) * (boxhi[1] - boxlo[1]) * (boxhi[2] - boxlo[2]);
  s.factor_eta = 1.0;
  for (int i = 0; i < 3; i++) { s.factor_v[i] = 1.0; s.dilation[i] = 1.0; }
  if (cudaMalloc(&h->d_state, sizeof(NhState)) != cudaSuccess || cudaMemcpy(h->d_state, &s, sizeof(NhState), cudaMemcpyHostToDevice) != cudaSuccess) {
    cudaGetLastError();
    if (h->d_state) cudaFree(h->d_state);
    delete h;
    return fail(ANNP_B200_ENOMEM, "cannot allocate the thermostat state on the device");
  }
  *out = h;
  return ANNP_B200_OK;
}

void annp_b200_nh_destroy(annp_b200_nh h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->d_state) cudaFree(h->d_state);
  if (h->d_partial) cudaFree(h->d_partial);
  delete h;
}

static int nh_partial(annp_b200_nh h, int nlocal) {
  const size_t nb = (size_t) ((nlocal + kRedSlice - 1) / kRedSlice) + 1;
  if (nb * 6 * sizeof(double) > h->partial_cap) {
    if (h->d_partial) cudaFree(h->d_partial);
    h->partial_cap = 0;
    if (cudaMalloc(&h->d_partial, nb * 6 * sizeof(double) * 2) != cudaSuccess) { cudaGetLastError(); h->d_partial = nullptr; return ANNP_B200_ENOMEM; }
    h->partial_cap = nb * 6 * sizeof(double) * 2;
  }
  return ANNP_B200_OK;
}

int annp_b200_nh_reduce(annp_b200_nh h, int nlocal, const double *d_v, const double *d_engvir, double *d_red12, void *stream) {
  if (!h || nlocal < 0 || !d_v || !d_red12) return ANNP_B200_EINVAL;
  if (cudaSetDevice(h->device) != cudaSuccess) return ANNP_B200_ECUDA;
  if (nh_partial(h, nlocal)) return ANNP_B200_ENOMEM;
  cudaStream_t s = (cudaStream_t) stream;
  const int nb = (nlocal + kRedSlice - 1) / kRedSlice;
  if (nb > 0) k_nh_vv_partial<<<nb, kRedThreads, 0, s>>>(nlocal, d_v, h->d_partial);
  k_nh_red_final<<<1, 32, 0, s>>>(h->d_partial, nb, h->host.mass, d_engvir, d_red12);
  return cudaGetLastError() == cudaSuccess ? ANNP_B200_OK : ANNP_B200_ECUDA;
}

int annp_b200_nh_setup(annp_b200_nh h, const double *d_red12, void *stream) {
  if (!h || !d_red12) return ANNP_B200_EINVAL;
  if (cudaSetDevice(h->device) != cudaSuccess) return ANNP_B200_ECUDA;
  k_nh_setup<<<1, 1, 0, (cudaStream_t) stream>>>(h->d_state, d_red12);
  return cudaGetLastError() == cudaSuccess ? ANNP_B200_OK : ANNP_B200_ECUDA;
}

int annp_b200_nh_initial(annp_b200_nh h, int nlocal, double *d_x, double *d_v, const double *d_f, int nsend, double *d_send_shift,
                         void *stream) {
  if (!h || nlocal < 0 || !d_x || !d_v || !d_f) return ANNP_B200_EINVAL;
  if (cudaSetDevice(h->device) != cudaSuccess) return ANNP_B200_ECUDA;
  cudaStream_t s = (cudaStream_t) stream;
  k_nh_begin<<<1, 1, 0, s>>>(h->d_state);
  if (nlocal > 0) k_nh_initial<<<grid_for(3LL * nlocal, 256), 256, 0, s>>>(h->d_state, nlocal, d_x, d_v, d_f);
  if (h->host.pstat && nsend > 0 && d_send_shift) k_nh_scale_shift<<<grid_for(3LL * nsend, 256), 256, 0, s>>>(h->d_state, nsend, d_send_shift);
  return cudaGetLastError() == cudaSuccess ? ANNP_B200_OK : ANNP_B200_ECUDA;
}

int annp_b200_nh_final_kick(annp_b200_nh h, int nlocal, double *d_v, const double *d_f, const double *d_engvir, double *d_red12,
                            void *stream) {
  if (!h || nlocal < 0 || !d_v || !d_f || !d_red12) return ANNP_B200_EINVAL;
  if (cudaSetDevice(h->device) != cudaSuccess) return ANNP_B200_ECUDA;
  if (nh_partial(h, nlocal)) return ANNP_B200_ENOMEM;
  cudaStream_t s = (cudaStream_t) stream;
  const int nb = (nlocal + kRedSlice - 1) / kRedSlice;
  if (nb > 0) k_nh_final_kick<<<nb, kRedThreads, 0, s>>>(h->d_state, nlocal, d_v, d_f, h->d_partial);
  k_nh_red_final<<<1, 32, 0, s>>>(h->d_partial, nb, h->host.mass, d_engvir, d_red12);
  return cudaGetLastError() == cudaSuccess ? ANNP_B200_OK : ANNP_B200_ECUDA;
}

int annp_b200_nh_final_scale(annp_b200_nh h, int nlocal, double *d_v, const double *d_red12, void *stream) {
  if (!h || nlocal < 0 || !d_v || !d_red12) return ANNP_B200_EINVAL;
  if (cudaSetDevice(h->device) != cudaSuccess) return ANNP_B200_ECUDA;
  cudaStream_t s = (cudaStream_t) stream;
  k_nh_end<<<1, 1, 0, s>>>(h->d_state, d_red12);
  if (nlocal > 0 && h->host.tstat) k_nh_scale_v<<<grid_for(3LL * nlocal, 256), 256, 0, s>>>(h->d_state, nlocal, d_v);
  return cudaGetLastError() == cudaSuccess ? ANNP_B200_OK : ANNP_B200_ECUDA;
}

int annp_b200_nh_set_box(annp_b200_nh h, const double *lo, const double *hi, const int *which, void *stream) {
  if (!h || !lo || !hi || !which) return ANNP_B200_EINVAL;
  if (cudaSetDevice(h->device) != cudaSuccess) return ANNP_B200_ECUDA;
  k_nh_set_box<<<1, 1, 0, (cudaStream_t) stream>>>(h->d_state, lo[0], lo[1], lo[2], hi[0], hi[1], hi[2], which[0], which[1], which[2]);
  return cudaGetLastError() == cudaSuccess ? ANNP_B200_OK : ANNP_B200_ECUDA;
}

int annp_b200_nh_get_state(annp_b200_nh h, annp_b200_nh_state *out, void *stream) {
  if (!h || !out) return ANNP_B200_EINVAL;
  if (cudaSetDevice(h->device) != cudaSuccess) return ANNP_B200_ECUDA;
  cudaStream_t s = (cudaStream_t) stream;
  if (cudaMemcpyAsync(&h->host, h->d_state, sizeof(NhState), cudaMemcpyDeviceToHost, s) != cudaSuccess) return ANNP_B200_ECUDA;
  if (cudaStreamSynchronize(s) != cudaSuccess) return ANNP_B200_ECUDA;
  const NhState &st = h->host;
  memset(out, 0, sizeof(*out));
  out->step = st.step;
  out->t_current = st.t_current; out->t_target = st.t_target;
  for (int i = 0; i < 3; i++) {
    out->p_current[i] = st.p_current[i]; out->boxlo[i] = st.boxlo[i]; out->boxhi[i] = st.boxhi[i];
    out->omega_dot[i] = st.omega_dot[i];
  }
  for (int k = 0; k < 6; k++) { out->ke_tensor[k] = st.mvv[k]; out->virial[k] = st.virial[k]; }
  for (int k = 0; k < kMaxChain; k++) { out->eta[k] = st.eta[k]; out->eta_dot[k] = st.eta_dot[k]; out->etap[k] = st.etap[k]; out->etap_dot[k] = st.etap_dot[k]; }
  // FixNH::compute_scalar: energy of the extended variables (added to PE + KE it is conserved)
  const double kt = st.boltz * st.t_target;
  double e = 0.0;
  if (st.tstat) {
    e += st.ke_target * st.eta[0] + 0.5 * st.eta_mass[0] * st.eta_dot[0] * st.eta_dot[0];
    for (int ich = 1; ich < st.mtchain; ich++) e += kt * st.eta[ich] + 0.5 * st.eta_mass[ich] * st.eta_dot[ich] * st.eta_dot[ich];
  }
  if (st.pstat) {
    const double volume = (st.boxhi[0] - st.boxlo[0]) * (st.boxhi[1] - st.boxlo[1]) * (st.boxhi[2] - st.boxlo[2]);
    for (int i = 0; i < 3; i++)
      if (st.p_flag[i]) e += 0.5 * st.omega_dot[i] * st.omega_dot[i] * st.omega_mass[i] + st.p_hydro * (volume - st.vol0) / (st.pdim * st.nktv2p);
    if (st.mpchain) {
      const double lkt_press = st.pdim * kt;
      e += lkt_press * st.etap[0] + 0.5 * st.etap_mass[0] * st.etap_dot[0] * st.etap_dot[0];
      for (int ich = 1; ich < st.mpchain; ich++) e += kt * st.etap[ich] + 0.5 * st.etap_mass[ich] * st.etap_dot[ich] * st.etap_dot[ich];
    }
  }
  out->extended_energy = e;
  return ANNP_B200_OK;
}

}    // extern "C"

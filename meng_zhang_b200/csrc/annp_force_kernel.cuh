// annp_force_kernel.cuh -- the fused ANNP force kernel for sm_100a (template; instantiated by annp_force.cu for the shapes of
// the shipped potentials and by annp_force_generic.cu for every other (npsf, ntsf)).
//
// One warp owns one centre atom at a time (dynamic atom scheduler).  For that atom it
//   1. filters the LAMMPS list row to the in-cutoff neighbours: (a) four 32-entry chunks per trip with all their loads in
//      flight, ballot compaction keeps list order; (b) one lane per KEPT neighbour computes unit vector / fc / dfc / r
//      into shared memory and accumulates the radial Chebyshev sums
//                                                       (reference: pair_annp.cpp:134-153, 633-656)
//   2. walks every unordered neighbour pair (j,k) ONCE and accumulates the angular sums in registers, in the block basis
//      T_4b(z) z^i of z = cos(theta) (one accumulate per order); all nsf sums are then reduced across the warp together
//      (transposing butterfly) and converted to the reference's T_n((z+1)/2)
//                                                       (reference: pair_annp.cpp:156-176, 658-695)
//   3. runs the element's MLP forward and reverse-mode backprop -> E_i and dE/dG
//                                                       (reference: pair_annp.cpp:741-804)
//   4. walks the pairs a second time, evaluating  A(y) = sum_n c_n T_n(y) and A'(y)  by Horner's rule in
//      z = cos(theta) (2 FMA per order; c is converted to monomial coefficients once per atom, held in registers) and
//      accumulating, per neighbour, the four moments
//          V = sum_k P u_k,  Aa = sum_k A fc_k     (P = A'/2 fc_j fc_k;  S = sum_k P cos(theta) = u_j . V)
//      so dG/dx is never materialised and no floating-point atomics are used
//   5. turns the moments into the force on every neighbour, F_j = -e_scale dOut/dx_j
//      (reference: pair_annp.cpp:191-200), and scatters it: FIXED = true adds it to the neighbour's fixed-point
//      accumulator with 64-bit integer atomics (exact, order independent), FIXED = false writes it at the neighbour's
//      LIST position for a later ordered gather; reduces F_i = -sum F_j, the per-centre virial and the energy.
//
// Pair schedule of passes 2 and 4 ("row-pair circulant").  The N neighbours are padded to an even
// count Np = 2M with a zero-weight dummy.  A lane owns the ROW PAIR (2m, 2m+1) and steps e = 1..M+1;
// in step e it meets the single partner k = (2m + e) mod Np with both of its rows: triplet (2m, k)
// has circulant offset e, triplet (2m+1, k) offset e-1, and offsets 1..M (the last one only for
// rows < M) enumerate every unordered pair exactly once.  Consequences:
//   * two independent triplets per lane are in flight (latency hiding on the FP64 pipe)
//   * the partner's data is loaded once, and its accumulators are updated once, per TWO triplets
//   * rows are stored parity-split (even rows first), so the partners of consecutive lanes are
//     consecutive shared-memory words: conflict-free 128-bit accesses, and all lanes of a step
//     touch DISTINCT partners, which makes the plain read-modify-write race free and the
//     summation order fixed (bit-reproducible results)
//   * the step range is cut into Q segments so that M*Q work units fill the 32 lanes evenly; a pass stops at the last
//     step any of its lanes needs.
//
// All arithmetic is IEEE FP64 (the reference CPU pair style is the parity target); the kernel is bound
// by the FP64 FMA pipe, see DESIGN.md.
#pragma once
#include "annp_device.cuh"

#ifndef ANNP_KWARPS
#define ANNP_KWARPS 4
#endif
#ifndef ANNP_MINBLOCKS
#define ANNP_MINBLOCKS 4
#endif
// unroll factors of the two triplet step loops (kernel experiments: -DANNP_UNROLL_FWD=2 / -DANNP_UNROLL_BWD=2)
#ifndef ANNP_UNROLL_FWD
#define ANNP_UNROLL_FWD 2      // same-box A/B at the bench size: 24.16 -> 24.02 ms (3: 24.22, 4: 24.64; backward 2: 24.35)
#endif
#ifndef ANNP_UNROLL_BWD
#define ANNP_UNROLL_BWD 1
#endif
#define ANNP_PRAGMA_(x) _Pragma(#x)
#define ANNP_PRAGMA(x) ANNP_PRAGMA_(x)

namespace {

constexpr int kWarps = ANNP_KWARPS;            // warps per block; each warp is independent
constexpr double kPi = 3.14159265358979323846;

// One lane's work unit in a pass over the row pairs: row pair m, steps e = elo + t for t < nsteps.
// Triplet 1 = (row 2m, partner) is valid for t < c1; triplet 2 = (row 2m+1, partner) for t2lo <= t < c2.
struct Unit {
  int m, elo, c1, c2, t2lo, seg;
  bool active;
};

struct Sched {
  int M;        // row pairs = padded neighbours / 2 = largest circulant offset
  int Q;        // segments of the step range
  int Hs;       // steps per segment
  int npass;    // warp passes
};

__device__ __forceinline__ Sched make_sched(int Np) {
  Sched sc;
  sc.M = Np >> 1;
  const int steps = sc.M + 1;
  int bestQ = 1;
  int best = ((sc.M + 31) >> 5) * steps;
  if (sc.M >= 32) {
#pragma unroll
    for (int q = 2; q <= 8; q <<= 1) {
      const int cost = ((sc.M * q + 31) >> 5) * (((steps + q - 1) / q) | 1) + q;   // + q: per-pass overhead
      if (cost < best) { best = cost; bestQ = q; }
    }
  }
  sc.Q = bestQ;
  sc.Hs = (steps + bestQ - 1) / bestQ;
  // lanes of one pass can belong to two consecutive segments (never more: Q > 1 needs M >= 32).  With an
  // ODD segment length their step numbers differ in parity, so they address different parity halves of
  // the row arrays and can never meet at the same partner.
  if (bestQ > 1) sc.Hs |= 1;
  sc.npass = (sc.M * bestQ + 31) >> 5;
  return sc;
}

__device__ __forceinline__ Unit make_unit(const Sched &sc, int pass, int lane) {
  Unit u;
  const int v = pass * 32 + lane;
  const int M = sc.M;
  u.active = v < M * sc.Q;
  u.seg = v / M;      // (a multiply-high by a per-atom reciprocal instead of this division measured 0.5 % SLOWER)
  u.m = v - u.seg * M;
  if (!u.active) { u.seg = 0; u.m = 0; }
  u.elo = 1 + u.seg * sc.Hs;
  const int ehi = min(M + 1, u.elo + sc.Hs - 1);
  const int j1 = 2 * u.m, j2 = j1 + 1;
  const int e1max = (j1 < M) ? M : M - 1;          // offset M only for rows < M
  const int e2max = ((j2 < M) ? M : M - 1) + 1;    // triplet 2 has offset e - 1
  u.c1 = u.active ? (min(ehi, e1max) - u.elo + 1) : 0;
  u.c2 = u.active ? (min(ehi, e2max) - u.elo + 1) : 0;
  u.t2lo = (u.elo == 1) ? 1 : 0;
  return u;
}

// Steps a pass needs: no lane of the warp has a valid triplet at or beyond this step (warp uniform).
__device__ __forceinline__ int pass_end(const Unit &u) { return __reduce_max_sync(0xffffffffu, max(u.c1, u.c2)); }

// One triplet of the forward angular pass: S[4b+i] += w T_{4b}(z) z^i  (see stage 2 of the kernel).
template <int NTSF>
__device__ __forceinline__ void angular_accumulate(double (&S)[NTSF], const double z, const double w) {
  const double z2 = z * z, z3 = z2 * z;
  const double t4 = fma(8.0, fma(z2, z2, -z2), 1.0);     // T_4(z) = 8 z^4 - 8 z^2 + 1
  const double t4x2 = t4 + t4;
  double pm = w, p = w;
#pragma unroll
  for (int b = 0; b < NTSF; b += 4) {
    S[b] += p;
    if (b + 1 < NTSF) S[b + 1] = fma(p, z, S[b + 1]);
    if (b + 2 < NTSF) S[b + 2] = fma(p, z2, S[b + 2]);
    if (b + 3 < NTSF) S[b + 3] = fma(p, z3, S[b + 3]);
    if (b + 4 < NTSF) {
      const double pn = (b == 0) ? t4 * w : fma(t4x2, p, -pm);
      pm = p; p = pn;
    }
  }
}

// x^y for the ADP radial functions.  The reference calls pow() (pair_anna_adp.cpp:197,200); for the positive bases that
// occur (r - r0 with r0 < 0, r / r1) exp(y log x) agrees with it to ~1e-15 relative, costs a third of CUDA's pow() and lets
// powers of the SAME base share one logarithm.  Non-positive bases keep pow()'s semantics through one out-of-line copy
// (inlined copies of pow() made the ANNA instantiation stall on instruction fetch: ncu r1b, stall_no_instruction 2.5).
__device__ __noinline__ double anna_pow_slow(double x, double y) { return pow(x, y); }

// ANNA-ADP tail (MODE 1): the descriptor of the centre atom is in sG (raw sums).  Reference: pair_anna_adp.cpp:166-272.
//   network -> (d2, q2); per-neighbour sums rho, mu[3], lambda[3][3], E_rep with the smooth step psi = z^4/(1+z^4),
//   z = (r - Rc)/hc; E_i; then the i-centred pair forces with d2, q2 held fixed, written at the neighbours' list
//   positions like the ANNP forces.  One lane per neighbour, fixed butterfly sums -> deterministic.
//   The transcendental factors of a neighbour (three powers, three exponentials) are the same in the sum pass and in
//   the force pass: they are computed once and parked in the shared-memory slots the ANNP backward pass would use.
template <bool FIXED>
__device__ __forceinline__ void anna_adp_tail(const ForceArgs &a, const DevParams &P, const double *We, const double *Be,
                                              const double *sG, double *sH, const double2 *sA, const double2 *sB,
                                              double2 *sC, double2 *accA, double2 *accB, double *accC, const int *spos, int N,
                                              int Ch, long long p0, int ii, int lane) {
#define ROWPOS(r) ((((r) & 1) ? Ch : 0) + ((r) >> 1))
  annp_mlp_forward_warp(P, We, Be, sG, sH, lane);
  const double d2 = sH[(P.nlayers - 1) * P.nnod], q2 = sH[(P.nlayers - 1) * P.nnod + 1];
  if (a.G_dbg)
    for (int n = lane; n < P.nsf; n += 32) { a.G_dbg[(size_t) ii * P.nsf + n] = sG[n]; a.dEdG_dbg[(size_t) ii * P.nsf + n] = (n == 0) ? d2 : (n == 1 ? q2 : 0.0); }
  // A0, yy, gamma, C0, c1F, c2F, V0, b1, b2, delta, r0, r1, hc, d1, q1, d3, q3: kernel parameters (constant bank)
#define A0 a.gp[0]
#define yy a.gp[1]
#define gamma a.gp[2]
#define C0 a.gp[3]
#define c1F a.gp[4]
#define c2F a.gp[5]
#define V0 a.gp[6]
#define b1 a.gp[7]
#define b2 a.gp[8]
#define delta a.gp[9]
#define r0 a.gp[10]
#define r1 a.gp[11]
#define hc a.gp[12]
#define d1 a.gp[13]
#define q1 a.gp[14]
#define d3 a.gp[15]
#define q3 a.gp[16]
  const double Rc = P.cut, hcinv = 1.0 / hc, r1inv = 1.0 / r1;
  const double rep_coeff = V0 / (b2 - b1);
  double rho = 0, mx = 0, my = 0, mz = 0, lxx = 0, lyy = 0, lzz = 0, lxy = 0, lxz = 0, lyz = 0, erep = 0;
  for (int s = lane; s < N; s += 32) {
    const int ps = ROWPOS(s);
    const double2 A = sA[ps], B = sB[ps];
    const double r = sC[ps].y;
    if (r > Rc) continue;                                          // pair_anna_adp.cpp:178 (r >= 1e-6 by the filter)
    const double x = r * A.x, y = r * A.y, z = r * B.x;
    const double sx = (r - Rc) * hcinv, sx2 = sx * sx, sx4 = sx2 * sx2;
    const double stp = sx4 / (1.0 + sx4);
    const double ut = d1 * exp(-d2 * r), wt = q1 * exp(-q2 * r);
    const double u = stp * (ut + d3);
    const double w = stp * (wt + q3);
    mx = fma(u, x, mx); my = fma(u, y, my); mz = fma(u, z, mz);
    lxx = fma(w * x, x, lxx); lyy = fma(w * y, y, lyy); lzz = fma(w * z, z, lzz);
    lxy = fma(w * x, y, lxy); lxz = fma(w * x, z, lxz); lyz = fma(w * y, z, lyz);
    const double rz = r - r0, ez = exp(-gamma * rz);
    const double pz = r * r1inv;
    double zyy, izb1, izb2;                                        // A0 rz^yy, pz^-b1, pz^-b2
    if (rz > 0.0) {
      zyy = A0 * exp(yy * log(rz));
      const double lp = log(pz);                                   // r >= 1e-6 by the filter
      izb1 = exp(-b1 * lp); izb2 = exp(-b2 * lp);
    } else {
      zyy = A0 * anna_pow_slow(rz, yy);
      izb1 = 1.0 / anna_pow_slow(pz, b1); izb2 = 1.0 / anna_pow_slow(pz, b2);
    }
    rho += stp * (zyy * ez * (1.0 + ez) + C0);
    erep += stp * (rep_coeff * (b2 * izb1 - b1 * izb2) + delta);
    accA[ps] = make_double2(ut, wt);
    accB[ps] = make_double2(ez, zyy);
    accC[ps] = izb1;
    sC[ps].x = izb2;                                               // dfc of the Chebyshev cutoff is not used by ANNA-ADP
  }
  rho = warp_sum(rho); mx = warp_sum(mx); my = warp_sum(my); mz = warp_sum(mz);
  lxx = warp_sum(lxx); lyy = warp_sum(lyy); lzz = warp_sum(lzz);
  lxy = warp_sum(lxy); lxz = warp_sum(lxz); lyz = warp_sum(lyz); erep = warp_sum(erep);
  const double v_i = lxx + lyy + lzz;
  const double sum_mu = mx * mx + my * my + mz * mz;
  const double sum_lam = lxx * lxx + lyy * lyy + lzz * lzz + 2.0 * (lxy * lxy + lxz * lxz + lyz * lyz);
  const double f_v = -1.0 / 3.0 * v_i;
  const double e_ang = 0.5 * sum_mu + 0.5 * sum_lam - 1.0 / 6.0 * v_i * v_i;
  const double e_emb = c1F * sqrt(rho) + c2F * rho * rho;
  const double e_i = 0.5 * erep + e_emb + e_ang + P.e_base;           // pair_anna_adp.cpp:213
  const double demb = 0.5 * c1F / sqrt(rho) + 2.0 * c2F * rho;

  double fix = 0, fiy = 0, fiz = 0;
  double v0 = 0, v1 = 0, v2 = 0, v3 = 0, v4 = 0, v5 = 0;
  for (int s = lane; s < N; s += 32) {
    const int ps = ROWPOS(s);
    const double2 A = sA[ps], B = sB[ps], Cc = sC[ps];
    const double r = Cc.y;
    const int q = spos[s];
    double fx = 0.0, fy = 0.0, fz = 0.0;
    const double x = r * A.x, y = r * A.y, z = r * B.x;
    if (!(r > Rc)) {
      const double rinv = 1.0 / r;
      const double2 c1 = accA[ps], c2 = accB[ps];
      const double ut = c1.x, wt = c1.y, ez = c2.x, zyy = c2.y, izb1 = accC[ps], izb2 = Cc.x;
      const double sx = (r - Rc) * hcinv, sx2 = sx * sx, sx4 = sx2 * sx2;
      const double rt1 = 1.0 / (1.0 + sx4);
      const double stp = sx4 * rt1;
      const double dstp = 4.0 * sx2 * sx * (rt1 * rt1) * hcinv;
      const double rz = r - r0;
      const double gz = zyy * gamma;
      const double drho = ez * (1.0 + ez) * (zyy * (dstp + stp * yy / rz) - gz) + C0 * dstp - gz * ez * ez;
      const double d_emb = demb * drho;
      const double rep_t1 = rep_coeff * (b2 * izb1 - b1 * izb2) + delta;
      // d/dr [b2 pz^-b1 - b1 pz^-b2] = (b1 b2 / r) (pz^-b2 - pz^-b1)
      const double d_rep = dstp * rep_t1 + stp * rep_coeff * (b2 * b1 * rinv * (izb2 - izb1));
      const double au = stp * (ut + d3), aw = 2.0 * stp * (wt + q3);
      const double dau = dstp * (ut + d3) + stp * (-d2 * ut);
      const double daw = dstp * (wt + q3) + stp * (-q2 * wt);
      const double dl1 = daw * (lxx * x * x + lyy * y * y + lzz * z * z);
      const double dl2 = daw * (lxy * x * y + lxz * x * z + lyz * y * z) * 2.0 + dl1;
      const double df1 = 0.5 * d_rep + d_emb + dau * (mx * x + my * y + mz * z) + dl2;
      const double df3 = f_v * (daw * r + aw);
      fx = df1 * A.x + aw * (y * lxy + z * lxz + x * lxx) + mx * au + x * df3;        // x / r = u_x
      fy = df1 * A.y + aw * (y * lyy + z * lyz + x * lxy) + my * au + y * df3;
      fz = df1 * B.x + aw * (y * lyz + z * lzz + x * lxz) + mz * au + z * df3;
    }
    if constexpr (FIXED) {                                            // f[j] += (fx, fy, fz), f[i] -=
      if (!annp_fix_add(a, a.nbr[p0 + q] & ANNP_NEIGHMASK, fx, fy, fz)) atomicExch(&a.cnt->bad_force, 1);
    } else {
      a.fpair[p0 + q] = make_double4(fx, fy, fz, 0.0);
    }
    fix -= fx; fiy -= fy; fiz -= fz;
    if (a.vir_c || a.vpair) {                                         // ev_tally_xyz(i, j, .., -f, x_ij)
      const double w0 = -x * fx, w1 = -y * fy, w2 = -z * fz, w3 = -x * fy, w4 = -x * fz, w5 = -y * fz;
      v0 += w0; v1 += w1; v2 += w2; v3 += w3; v4 += w4; v5 += w5;
      if (a.vpair) {
        double *vp = a.vpair + (size_t) (p0 + q) * 6;
        vp[0] = w0; vp[1] = w1; vp[2] = w2; vp[3] = w3; vp[4] = w4; vp[5] = w5;
      }
    }
  }
  fix = warp_sum(fix); fiy = warp_sum(fiy); fiz = warp_sum(fiz);
  if (lane == 0) a.fself[ii] = make_double4(fix, fiy, fiz, e_i);
  if (a.vir_c) {
    v0 = warp_sum(v0); v1 = warp_sum(v1); v2 = warp_sum(v2);
    v3 = warp_sum(v3); v4 = warp_sum(v4); v5 = warp_sum(v5);
    if (lane == 0) {
      double *vc = a.vir_c + (size_t) ii * 6;
      vc[0] = v0; vc[1] = v1; vc[2] = v2; vc[3] = v3; vc[4] = v4; vc[5] = v5;
    }
  }
  __syncwarp();
#undef ROWPOS
#undef A0
#undef yy
#undef gamma
#undef C0
#undef c1F
#undef c2F
#undef V0
#undef b1
#undef b2
#undef delta
#undef r0
#undef r1
#undef hc
#undef d1
#undef q1
#undef d3
#undef q3
}

// resident blocks per SM the register budget is set for: the shipped shapes (ntsf <= 20) run 4 blocks of 128 registers; the
// padded generic shapes keep 24 + 25 angular coefficients per lane in registers and get 2 blocks of 255
constexpr int annp_min_blocks(int ntsf) { return ntsf > 20 ? 2 : ANNP_MINBLOCKS; }

#ifdef ANNP_STAGE_CLOCKS
#define ANNP_STAMP(k) do { const long long now__ = clock64(); stage_acc[k] += now__ - stage_t; stage_t = now__; } while (0)
#else
#define ANNP_STAMP(k) do { } while (0)
#endif

template <int NPSF, int NTSF, int MODE, bool FIXED>
__global__ void __launch_bounds__(kWarps * 32, annp_min_blocks(NTSF)) annp_force_kernel(const ForceArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const DevParams &P = *a.prm;
  const int C = a.capacity;
  const int nsf = NPSF + NTSF, nnod = P.nnod, nl = P.nlayers, nt1 = P.ntypes + 1;
  const int wtot = P.nelements * P.w_per_elem, btot = P.nelements * P.b_per_elem;

  // ---- block-shared parameters.  The network weights stay in global memory (3.7 KB, L1 resident, touched by <= 28
  // lanes for ~1 % of an atom's time): keeping them out of shared memory lets a FOURTH block fit on the SM
  // (59.9 -> 56.2 KB per block at a 128-slot tile), i.e. 4 instead of 3 warps per scheduler to hide FP64 latency.
  double *sScale = reinterpret_cast<double *>(smem_raw);
  double *sAvg = sScale + nsf;
  double *blk_end = sAvg + nsf;
  const double *__restrict__ sW = P.weights;
  const double *__restrict__ sBias = P.bias;
  const double *__restrict__ gC2M = P.cheb2mono;     // [NTSF][NTSF] Chebyshev -> monomial(z) matrix (L1/L2 resident)
  const double *__restrict__ gB2C = P.blk2cheb;      // [NTSF][NTSF] forward block basis -> Chebyshev
  (void) wtot; (void) btot;
  for (int t = threadIdx.x; t < nsf; t += blockDim.x) { sScale[t] = P.sf_scale[t]; sAvg[t] = P.sf_avg[t]; }

  // ---- per-warp region
  const size_t per_warp_doubles = (size_t) 11 * C + 2 * NTSF + 2 * NPSF + 2 * nsf + (size_t) 2 * nl * nnod + 2 * nnod;
  size_t warp_bytes = per_warp_doubles * sizeof(double) + (size_t) C * sizeof(int);
  warp_bytes = (warp_bytes + 15) & ~(size_t) 15;
  size_t blk_bytes = ((size_t) ((unsigned char *) blk_end - smem_raw) + 15) & ~(size_t) 15;
  unsigned char *wbase = smem_raw + blk_bytes + (size_t) warp * warp_bytes;
  double2 *sA = reinterpret_cast<double2 *>(wbase);   // ux, uy
  double2 *sB = sA + C;                               // uz, fc
  double2 *sC = sB + C;                               // dfc, r
  double2 *accA = sC + C;                             // Vx, Vy
  double2 *accB = accA + C;                           // Vz, Aa
  double *accC = reinterpret_cast<double *>(accB + C);   // scratch of the ANNA-ADP tail
  int *sj = reinterpret_cast<int *>(accC);            // Chebyshev ANNP, fixed-point scatter: atom index of the kept neighbour
  double2 *coefT = reinterpret_cast<double2 *>(accC + C);   // angular polynomial: NTSF monomial coefficients a_k (as doubles)
  double2 *coefR = coefT + NTSF;                      // radial  (d_m, e_m)
  double *sG = reinterpret_cast<double *>(coefR + NPSF);
  double *sdE = sG + nsf;
  double *sH = sdE + nsf;                             // [nl][nnod] activations
  double *sHd = sH + nl * nnod;                       // [nl][nnod] activation derivatives
  double *sDel = sHd + nl * nnod;                     // [2][nnod] backprop deltas
  int *spos = reinterpret_cast<int *>(sDel + 2 * nnod);
  __syncthreads();

  const double two_over_cut = P.two_over_cut;
  const int Ch = C >> 1;                  // rows are stored parity-split: even rows [0,Ch), odd rows [Ch,C)
#define ROWPOS(r) ((((r) & 1) ? Ch : 0) + ((r) >> 1))

  const unsigned long long nwork = annp_work_items(a);
  if (a.work_list && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&a.cnt->ovf_total, nwork);   // overflow pass: statistics
#ifdef ANNP_STAGE_CLOCKS
  long long stage_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long stage_t = clock64();
#endif
  for (;;) {
    unsigned long long item = 0;
    if (lane == 0) item = atomicAdd(a.work_ctr, 1ull);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= nwork) break;
    const int ii = annp_work_centre(a, item);
    const int i = a.ilist[ii];
    ANNP_STAMP(6);
    const double4 xi = a.xq[i];
    const int ti = (int) xi.w;
    const long long p0 = a.row_off[ii];
    const int L = (int) (a.row_off[ii + 1] - p0);

    // ------------------------------------------------------------------ 1. filter + radial sums
    // 1a (light, latency bound): walk the list row four 32-entry chunks at a time with all loads of the four chunks in
    // flight together, keep the in-cutoff entries (ballot compaction, list order) and park (dx, dy, dz, r^2, 1/Rc) in
    // the neighbour's shared-memory slot.
    int N = 0;
    for (int base = 0; base < L; base += 128) {
      int jn[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int q = base + 32 * u + lane;
        jn[u] = (q < L) ? (a.nbr[p0 + q] & ANNP_NEIGHMASK) : -1;
      }
      double4 xn[4];
#pragma unroll
      for (int u = 0; u < 4; u++) xn[u] = (jn[u] >= 0) ? a.xq[jn[u]] : make_double4(0.0, 0.0, 0.0, 0.0);
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int q = base + 32 * u + lane;
        const bool valid = jn[u] >= 0;
        const double dx = xi.x - xn[u].x, dy = xi.y - xn[u].y, dz = xi.z - xn[u].z;
        const double rsq = dx * dx + dy * dy + dz * dz;
        const int tj = (int) xn[u].w;
        const bool in = valid && !(rsq > P.cutsq[ti * nt1 + tj] || rsq < 1.0e-12);          // pair_annp.cpp:144
        const unsigned mask = __ballot_sync(0xffffffffu, in);
        const int slot = N + __popc(mask & ((1u << lane) - 1u));
        if (in && slot < C) {
          const int ps = ROWPOS(slot);
          sA[ps] = make_double2(dx, dy);
          sB[ps] = make_double2(dz, rsq);
          sC[ps] = make_double2(P.rcinv[ti * nt1 + tj], 0.0);
          spos[slot] = q;
          if constexpr (FIXED && MODE == 0) sj[slot] = jn[u];
        } else if (valid) {
          if constexpr (!FIXED) a.fpair[p0 + q] = make_double4(0.0, 0.0, 0.0, 0.0);
          if (a.vpair) {
            double *vp = a.vpair + (size_t) (p0 + q) * 6;
#pragma unroll
            for (int k = 0; k < 6; k++) vp[k] = 0.0;
          }
        }
        N += __popc(mask);
      }
    }
    __syncwarp();
    ANNP_STAMP(0);
    // 1b (FP64): one lane per KEPT neighbour: unit vector, cutoff function, radial Chebyshev sums
    double gr[NPSF];
#pragma unroll
    for (int m = 0; m < NPSF; m++) gr[m] = 0.0;
    for (int sl = lane; sl < min(N, C); sl += 32) {
      const int ps = ROWPOS(sl);
      const double2 dA = sA[ps], dB = sB[ps];
      const double rci = sC[ps].x;
      const double r = sqrt(dB.y);
      const double rinv = 1.0 / r;
      double sn, cs;
      sincospi(r * rci, &sn, &cs);
      const double fc = 0.5 * (cs + 1.0);          // pair_annp.cpp:590-594
      const double dfc = -0.5 * kPi * rci * sn;
      sA[ps] = make_double2(dA.x * rinv, dA.y * rinv);
      sB[ps] = make_double2(dB.x * rinv, fc);
      sC[ps] = make_double2(dfc, r);
      accA[ps] = make_double2(0.0, 0.0);
      accB[ps] = make_double2(0.0, 0.0);
      // radial Chebyshev sums, argument 2r/Rc - 1    (pair_annp.cpp:643-647)
      const double xr = r * two_over_cut - 1.0, xr2 = xr + xr;
      double t0 = 1.0, t1 = xr;
      gr[0] += fc;
      if (NPSF > 1) gr[1] = fma(t1, fc, gr[1]);
#pragma unroll
      for (int m = 2; m < NPSF; m++) {
        const double t = fma(xr2, t1, -t0);
        gr[m] = fma(t, fc, gr[m]);
        t0 = t1; t1 = t;
      }
    }
    if (lane == 0 && !a.work_list) {
      atomicMax(&a.cnt->max_neigh, N);
      atomicAdd(&a.cnt->sum_neigh, (unsigned long long) N);
      atomicAdd(&a.cnt->sum_trip, (unsigned long long) N * (unsigned long long) (N > 0 ? N - 1 : 0) / 2ull);
    }
    if (N + (N & 1) > C) {   // tile exceeded: hand the atom to the overflow pass (larger tile), emit zeros meanwhile
      if (lane == 0) { annp_note_overflow(a, ii); a.fself[ii] = make_double4(0.0, 0.0, 0.0, 0.0); }
      if constexpr (!FIXED) for (int q = lane; q < L; q += 32) a.fpair[p0 + q] = make_double4(0.0, 0.0, 0.0, 0.0);
      __syncwarp();
      continue;
    }
    if ((N & 1) && lane == 0) {   // zero-weight dummy row pads N to an even count
      const int ps = ROWPOS(N);
      sA[ps] = make_double2(1.0, 0.0);
      sB[ps] = make_double2(0.0, 0.0);
      sC[ps] = make_double2(0.0, 1.0);
      accA[ps] = make_double2(0.0, 0.0);
      accB[ps] = make_double2(0.0, 0.0);
    }
    __syncwarp();

    const Sched sch = make_sched(N + (N & 1));
    const int M = sch.M;
    ANNP_STAMP(1);

    // ------------------------------------------------------------------ 2. angular sums (forward)
    double S[NTSF];
#pragma unroll
    for (int n = 0; n < NTSF; n++) S[n] = 0.0;
    for (int pass = 0; pass < sch.npass; pass++) {
      const Unit un = make_unit(sch, pass, lane);
      const int end = pass_end(un);
      double2 A1 = sA[un.m], B1 = sB[un.m], A2 = sA[Ch + un.m], B2 = sB[Ch + un.m];
      // partner position: parity half of e, index (m + e/2) mod M
      int e = un.elo;
      int kc = un.m + (e >> 1);
      if (kc >= M) kc -= M;
      e &= 1;
      double2 Ak = sA[((e & 1) ? Ch : 0) + kc], Bk = sB[((e & 1) ? Ch : 0) + kc];
      int t = 0;
      // Block basis psi_{4b+i}(z) = T_{4b}(z) z^i in z = cos(theta): P_b = w T_{4b}(z) advances by the Chebyshev
      // recurrence in T_4(z) (one DFMA per FOUR orders) and every order is ONE accumulate, S[4b+i] += P_b z^i
      // (28 FP64 instructions per triplet for 19 orders; the order-by-order recurrence Z_n = 2y Z_{n-1} - Z_{n-2}
      // needs 38).  The sums are converted to the reference's T_n((z+1)/2) once per atom (blk2cheb); the basis is as
      // well conditioned as T_n itself (conversion rows sum to <= 25; parity stays at 1e-14 in G).
      auto step = [&]() {
        // next step: e + 1 flips the parity half; the index advances when e becomes even (prefetch)
        kc += e; if (kc == M) kc = 0;          // (e holds only the parity of the step number)
        e ^= 1;
        const int kpn = (e ? Ch : 0) + kc;
        const double2 Akn = sA[kpn], Bkn = sB[kpn];
        const double f1 = (t < un.c1) ? B1.y : 0.0;                       // zero for the triplets this lane must skip
        const double f2 = (t >= un.t2lo && t < un.c2) ? B2.y : 0.0;
        angular_accumulate<NTSF>(S, fma(A1.x, Ak.x, fma(A1.y, Ak.y, B1.x * Bk.x)), f1 * Bk.y);   // pair_annp.cpp:671-678
        angular_accumulate<NTSF>(S, fma(A2.x, Ak.x, fma(A2.y, Ak.y, B2.x * Bk.x)), f2 * Bk.y);
        Ak = Akn; Bk = Bkn;
      };
      ANNP_PRAGMA(unroll ANNP_UNROLL_FWD)
      for (; t < end; t++) step();
    }
    // warp reduction, scaling and centring (pair_annp.cpp:178-180).  All nsf <= 32 sums are reduced TOGETHER by a
    // transposing butterfly: in the round with lane mask m every lane keeps the half of its values that belongs to its
    // side of the mask and hands the other half over, so the number of live values halves each round (16+8+4+2+1 = 31
    // shuffles instead of 5 per sum) and lane n ends up with the total of sum n.  Fixed order -> deterministic.
    ANNP_STAMP(2);
    {
      constexpr int NSF = NPSF + NTSF;
      static_assert(NSF <= 64, "the transposing reduction handles at most 2 x 32 descriptor components");
#pragma unroll
      for (int rnd = 0; rnd < (NSF + 31) / 32; rnd++) {
        double w[32];
#pragma unroll
        for (int n = 0; n < 32; n++) {
          const int c = 32 * rnd + n;                             // component reduced into lane n of this round
          w[n] = (c < NPSF) ? gr[c < NPSF ? c : 0] : (c < NSF ? S[(c >= NPSF && c < NSF) ? c - NPSF : 0] : 0.0);
        }
        const double tot = warp_transpose_sum32(w, lane);       // total of component 32 rnd + lane
        const int c = 32 * rnd + lane;
        if (c < NPSF) sG[c] = sScale[c] * tot - sScale[c] * sAvg[c];
        else if (c < NSF) sdE[c] = tot;                           // block-basis sums, parked in sdE (free until stage 3)
      }
      __syncwarp();
      if (lane < NTSF) {                                        // block basis -> T_n((z+1)/2), lane n
        double v = 0.0;
#pragma unroll
        for (int j = 0; j < NTSF; j++) v = fma(__ldg(gB2C + j * NTSF + lane), sdE[NPSF + j], v);
        sG[NPSF + lane] = sScale[NPSF + lane] * v - sScale[NPSF + lane] * sAvg[NPSF + lane];
      }
    }
    __syncwarp();

    const int elem = P.map[ti];
    const double *We = sW + elem * P.w_per_elem;
    const double *Be = sBias + elem * P.b_per_elem;
    if constexpr (MODE == 1) {      // ANNA-ADP: forward network + ADP energy and forces, no descriptor derivatives
      anna_adp_tail<FIXED>(a, P, We, Be, sG, sH, sA, sB, sC, accA, accB, accC, spos, N, Ch, p0, ii, lane);
      ANNP_STAMP(5);
      continue;
    }

    // ------------------------------------------------------------------ 3. MLP forward + backprop
    const double out = annp_mlp_warp(P, We, Be, sG, sdE, sH, sHd, sDel, lane);
    const double e_i = P.e_scale * out + P.e_shift + P.e_atom;      // pair_annp.cpp:790-793
    if (a.G_dbg) for (int n = lane; n < nsf; n += 32) { a.G_dbg[(size_t) ii * nsf + n] = sG[n]; a.dEdG_dbg[(size_t) ii * nsf + n] = sdE[n]; }
    ANNP_STAMP(3);

    // Chebyshev-T coefficients c_n = s_n dOut/dG_n  ->  U-basis coefficients
    //   sum c_n T_n = sum d_n U_n,  d_0 = c_0 - c_2/2, d_n = (c_n - c_{n+2})/2
    //   d/dy sum c_n T_n = sum_{m} (m+1) c_{m+1} U_m        (angular e_m carries the reference's 1/2)
    // angular: p(z) = sum_k a_k z^k = sum_n c_n T_n((z+1)/2),  a = C2M c   (then A = p, A'/2 = dp/dz)
    double *aK = reinterpret_cast<double *>(coefT);
    if (lane < NTSF) {
      double acc = 0.0;
#pragma unroll
      for (int n = 0; n < NTSF; n++) acc = fma(__ldg(gC2M + lane * NTSF + n), sdE[NPSF + n] * sScale[NPSF + n], acc);
      aK[lane] = acc;
    }
    if (lane < NPSF) {
      const int n = lane;
      const double c0 = sdE[n] * sScale[n];
      const double c1 = (n + 1 < NPSF) ? sdE[n + 1] * sScale[n + 1] : 0.0;
      const double c2 = (n + 2 < NPSF) ? sdE[n + 2] * sScale[n + 2] : 0.0;
      const double dn = (n == 0) ? (c0 - 0.5 * c2) : 0.5 * (c0 - c2);
      coefR[n] = make_double2(dn, (double) (n + 1) * c1);
    }
    __syncwarp();

    // ------------------------------------------------------------------ 4. angular moments (backward)
    // the monomial coefficients live in registers for the whole backward pass (read two at a time)
    double akv[NTSF + 1];
#pragma unroll
    for (int i = 0; i < (NTSF + 1) / 2; i++) { const double2 c2 = coefT[i]; akv[2 * i] = c2.x; akv[2 * i + 1] = c2.y; }
    for (int pass = 0; pass < sch.npass; pass++) {
      const Unit un = make_unit(sch, pass, lane);
      const int end = pass_end(un);
      double2 A1 = sA[un.m], B1 = sB[un.m], A2 = sA[Ch + un.m], B2 = sB[Ch + un.m];
      double v1x = 0, v1y = 0, v1z = 0, a1 = 0;
      double v2x = 0, v2y = 0, v2z = 0, a2 = 0;
      int e = un.elo;
      int kc = un.m + (e >> 1);
      if (kc >= M) kc -= M;
      e &= 1;
      int kp = ((e & 1) ? Ch : 0) + kc;
      double2 Ak = sA[kp], Bk = sB[kp];
      int t = 0;
      auto step = [&]() {
        // this step's partner accumulators: only this lane touches them until the next __syncwarp
        double2 pa = accA[kp], pb = accB[kp];
        // next step's partner (read-only data, prefetched across the barrier)
        kc += e; if (kc == M) kc = 0;          // (e holds only the parity of the step number)
        e ^= 1;
        const int kpn = (e ? Ch : 0) + kc;
        const double2 Akn = sA[kpn], Bkn = sB[kpn];
        const bool ok1 = t < un.c1, ok2 = (t >= un.t2lo && t < un.c2);
        const double f1 = ok1 ? B1.y : 0.0, f2 = ok2 ? B2.y : 0.0;      // fc_j, zero for the triplets this lane must skip
        const double g1 = ok1 ? Bk.y : 0.0, g2 = ok2 ? Bk.y : 0.0;      // fc_k
        const double cta = fma(A1.x, Ak.x, fma(A1.y, Ak.y, B1.x * Bk.x));
        const double ctb = fma(A2.x, Ak.x, fma(A2.y, Ak.y, B2.x * Bk.x));
        // Horner with derivative in z = cos(theta):  d <- d z + b ; b <- b z + a_k   (A = b, A'(y)/2 = d)
        // (first step folded by hand: d = a_top, b = a_top z + a_{top-1})
        const double atop = akv[NTSF - 1], atop1 = akv[NTSF - 2];
        double Apa = atop, Apb = atop, Aa_ = fma(atop, cta, atop1), Ab_ = fma(atop, ctb, atop1);
#pragma unroll
        for (int k = NTSF - 3; k >= 0; k--) {
          const double ak = akv[k];
          Apa = fma(Apa, cta, Aa_);
          Aa_ = fma(Aa_, cta, ak);
          Apb = fma(Apb, ctb, Ab_);
          Ab_ = fma(Ab_, ctb, ak);
        }
        const double Pa = Apa * (f1 * Bk.y), Pb = Apb * (f2 * Bk.y);
        // row side (registers)
        // (S = sum_k P cos(theta_jk) is not accumulated: it equals u_j . V_j and is formed once per neighbour in stage 5)
        v1x = fma(Pa, Ak.x, v1x); v1y = fma(Pa, Ak.y, v1y); v1z = fma(Pa, Bk.x, v1z);
        a1 = fma(Aa_, g1, a1);
        v2x = fma(Pb, Ak.x, v2x); v2y = fma(Pb, Ak.y, v2y); v2z = fma(Pb, Bk.x, v2z);
        a2 = fma(Ab_, g2, a2);
        // partner side: ONE shared-memory read-modify-write for both triplets
        pa.x = fma(Pa, A1.x, pa.x); pa.y = fma(Pa, A1.y, pa.y); pb.x = fma(Pa, B1.x, pb.x);
        pa.x = fma(Pb, A2.x, pa.x); pa.y = fma(Pb, A2.y, pa.y); pb.x = fma(Pb, B2.x, pb.x);
        pb.y = fma(Aa_, f1, pb.y); pb.y = fma(Ab_, f2, pb.y);
        if (un.active) { accA[kp] = pa; accB[kp] = pb; }
        __syncwarp();
        kp = kpn; Ak = Akn; Bk = Bkn;
      };
      ANNP_PRAGMA(unroll ANNP_UNROLL_BWD)
      for (; t < end; t++) step();
      // flush the row side; the same row pair can sit in several lanes (segments) -> one segment at a time
      for (int g = 0; g < sch.Q; g++) {
        if (un.active && un.seg == g) {
          double2 ja = accA[un.m], jb = accB[un.m];
          ja.x += v1x; ja.y += v1y; jb.x += v1z; jb.y += a1;
          accA[un.m] = ja; accB[un.m] = jb;
          double2 oa = accA[Ch + un.m], ob = accB[Ch + un.m];
          oa.x += v2x; oa.y += v2y; ob.x += v2z; ob.y += a2;
          accA[Ch + un.m] = oa; accB[Ch + un.m] = ob;
        }
        __syncwarp();
      }
    }

    ANNP_STAMP(4);
    // ------------------------------------------------------------------ 5. forces on the neighbours
    double fix = 0, fiy = 0, fiz = 0;
    double v0 = 0, v1 = 0, v2 = 0, v3 = 0, v4 = 0, v5 = 0;
    const double mes = -P.e_scale;
    for (int s = lane; s < N; s += 32) {
      const int ps = ROWPOS(s);
      const double2 A = sA[ps], B = sB[ps], Cc = sC[ps];
      const double2 va = accA[ps], vb = accB[ps];
      const double aa = vb.y;                                          // Aa = sum_k A fc_k
      const double sdot = fma(va.x, A.x, fma(va.y, A.y, vb.x * B.x));  // S = sum_k P cos(theta_jk) = u_j . V_j
      const double ux = A.x, uy = A.y, uz = B.x, fc = B.y, dfc = Cc.x, r = Cc.y;
      const double rinv = 1.0 / r;
      // radial polynomial R(x) = sum c_m T_m(x) and R'(x) in the U basis
      const double x2 = 2.0 * (r * two_over_cut - 1.0);
      double u0 = 1.0, u1 = x2;
      const double2 q0 = coefR[0];
      double Rv = q0.x, Rp = q0.y;
      if (NPSF > 1) { const double2 q1 = coefR[1]; Rv = fma(q1.x, u1, Rv); Rp = fma(q1.y, u1, Rp); }
#pragma unroll
      for (int m = 2; m < NPSF; m++) {
        const double un_ = fma(x2, u1, -u0);
        const double2 qm = coefR[m];
        Rv = fma(qm.x, un_, Rv);
        if (m < NPSF - 1) Rp = fma(qm.y, un_, Rp);
        u0 = u1; u1 = un_;
      }
      // d out / d x_j = g u_j - V / r       with dr/dx_j = -u_j, dcos/dx_j = (cos u_j - u_k)/r
      const double g = -(Rp * two_over_cut * fc + Rv * dfc) - dfc * aa + sdot * rinv;
      const double gx = g * ux - va.x * rinv;
      const double gy = g * uy - va.y * rinv;
      const double gz = g * uz - vb.x * rinv;
      const double Fx = mes * gx, Fy = mes * gy, Fz = mes * gz;     // pair_annp.cpp:197
      const int q = spos[s];
      if constexpr (FIXED) {
        if (!annp_fix_add(a, sj[s], Fx, Fy, Fz)) atomicExch(&a.cnt->bad_force, 1);
      } else {
        a.fpair[p0 + q] = make_double4(Fx, Fy, Fz, 0.0);
      }
      fix -= Fx; fiy -= Fy; fiz -= Fz;
      if (a.vir_c || a.vpair) {
        // ev_tally_xyz(i, j, ..., -Fj, xi - xj)     (pair_annp.cpp:201-209)
        const double delx = r * ux, dely = r * uy, delz = r * uz;
        const double w0 = -delx * Fx, w1 = -dely * Fy, w2 = -delz * Fz;
        const double w3 = -delx * Fy, w4 = -delx * Fz, w5 = -dely * Fz;
        v0 += w0; v1 += w1; v2 += w2; v3 += w3; v4 += w4; v5 += w5;
        if (a.vpair) {
          double *vp = a.vpair + (size_t) (p0 + q) * 6;
          vp[0] = w0; vp[1] = w1; vp[2] = w2; vp[3] = w3; vp[4] = w4; vp[5] = w5;
        }
      }
    }
    fix = warp_sum(fix); fiy = warp_sum(fiy); fiz = warp_sum(fiz);
    if (lane == 0) a.fself[ii] = make_double4(fix, fiy, fiz, e_i);
    if (a.vir_c) {
      v0 = warp_sum(v0); v1 = warp_sum(v1); v2 = warp_sum(v2);
      v3 = warp_sum(v3); v4 = warp_sum(v4); v5 = warp_sum(v5);
      if (lane == 0) {
        double *vc = a.vir_c + (size_t) ii * 6;
        vc[0] = v0; vc[1] = v1; vc[2] = v2; vc[3] = v3; vc[4] = v4; vc[5] = v5;
      }
    }
    __syncwarp();
    ANNP_STAMP(5);
  }
#ifdef ANNP_STAGE_CLOCKS
  if (lane == 0)
    for (int k = 0; k < 8; k++) atomicAdd(&a.cnt->stage_clk[k], (unsigned long long) stage_acc[k]);
#endif
#undef ROWPOS
}

}    // namespace


// ---- host side shared by annp_force.cu / annp_force_generic.cu -------------------------------------------------------

typedef void (*annp_force_kernel_t)(const ForceArgs);

// dynamic shared memory of one block: two block-wide vectors + kWarps per-warp regions (layout at the top of the kernel)
static inline size_t annp_force_smem_bytes_shape(int npsf, int ntsf, int nlayers, int nnod, int capacity) {
  const int nsf = npsf + ntsf;
  size_t blk = (size_t) (2 * nsf) * sizeof(double);
  blk = (blk + 15) & ~(size_t) 15;
  size_t per_warp = ((size_t) 11 * capacity + 2 * ntsf + 2 * npsf + 2 * nsf + (size_t) 2 * nlayers * nnod + 2 * nnod) * sizeof(double) +
                    (size_t) capacity * sizeof(int);
  per_warp = (per_warp + 15) & ~(size_t) 15;
  return blk + kWarps * per_warp;
}

// Launch `k` on `stream`: one full wave of resident blocks, or fewer when there is less work.  max_blocks > 0 caps the grid
// (the overflow pass, whose item count is only known on the device, runs one block per SM).
static inline cudaError_t annp_force_launch_kernel(annp_force_kernel_t k, const ForceArgs &args, const DevParams &hp, int num_sms,
                                            cudaStream_t stream, int max_blocks) {
  const size_t smem = annp_force_smem_bytes_shape(hp.npsf, hp.ntsf, hp.nlayers, hp.nnod, args.capacity);
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
  if (e != cudaSuccess) return e;
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, kWarps * 32, smem);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) return cudaErrorInvalidConfiguration;
  long long want = ((long long) args.inum + kWarps - 1) / kWarps;
  long long blocks = (long long) per_sm * num_sms;
  if (blocks > want) blocks = want;
  if (max_blocks > 0 && blocks > max_blocks) blocks = max_blocks;
  if (blocks < 1) blocks = 1;
  k<<<(unsigned) blocks, kWarps * 32, smem, stream>>>(args);
  return cudaGetLastError();
}

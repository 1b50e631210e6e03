"""numpy twin of the ARITHMETIC the CUDA force kernel uses (meng_zhang_b200/csrc/annp_force.cu), for CPU tests.

The kernel does not follow the reference's operation order (fe_v2/src/pair_annp.cpp:74-222): it
  * accumulates the angular sums in the block basis psi_{4b+i}(z) = T_{4b}(z) z^i of z = cos(theta) and converts them to
    the reference's T_n((z+1)/2) with a matrix (annp_b200_basis_matrices),
  * replaces the forward-mode Jacobian by reverse-mode backprop,
  * never forms dG/dx: it evaluates A(z) = sum_n c_n T_n((z+1)/2) and dA/dz by Horner's rule in monomials of z and keeps,
    per neighbour, V = sum_k P u_k and Aa = sum_k A fc_k  (P = dA/dz fc_j fc_k; sum_k P cos(theta_jk) = u_j . V),
  * rounds every pair force to a multiple of 2^-43 eV/A and sums the integers.
This file restates exactly that in plain numpy so that the algebra can be checked against the oracle WITHOUT a GPU
(tests/test_host_logic.py).  Test infrastructure only; single element, activation flags 0 / 1 / 3 / 4 of the Fe copy.
"""
import numpy as np

FIX_BITS = 43


def _act(flag, z):
    ca, cb, cc = 1.7159, 0.666666666666667, 0.1
    if flag == 0:
        return z, np.ones_like(z)
    if flag == 1:
        h = np.tanh(z)
        return h, 1.0 - h * h
    t = np.tanh(cb * z)
    if flag == 3:
        return ca * t, ca * (1.0 - t * t) * cb
    return ca * t + cc * z, ca * (1.0 - t * t) * cb + cc          # 4 (what "tanh" in the files maps to)


def _mlp(pot, G):
    """raw output and d out / d G by reverse mode (annp_device.cuh: annp_mlp_warp)."""
    nl = pot.ntl - 1
    h, hs, ds = G, [], []
    for l in range(nl):
        nrow = 1 if l == nl - 1 else pot.nnod
        ncol = pot.nsf if l == 0 else pot.nnod
        z = pot.weight_all[0, l, :nrow, :ncol] @ h + pot.bias_all[0, l, :nrow]
        h, d = _act(pot.flagact[l], z)
        hs.append(h)
        ds.append(d)
    delta = ds[-1].copy()
    for l in range(nl - 1, 0, -1):
        nrow = 1 if l == nl - 1 else pot.nnod
        delta = (pot.weight_all[0, l, :nrow, :pot.nnod].T @ delta) * ds[l - 1]
    nrow0 = 1 if nl == 1 else pot.nnod
    return float(hs[-1][0]), pot.weight_all[0, 0, :nrow0, :pot.nsf].T @ delta


def _cheb(x, n):
    T = np.zeros((n,) + np.shape(x))
    T[0] = 1.0
    if n > 1:
        T[1] = x
    for m in range(2, n):
        T[m] = 2.0 * x * T[m - 1] - T[m - 2]
    return T


def compute(pot, cfg, cheb2mono, blk2cheb, fixed_point=True):
    """Forces f[nall,3] (ghost rows not folded) and per-atom energies, the kernel's way."""
    npsf, ntsf, Rc = pot.npsf, pot.ntsf, pot.cut
    s = pot.sf_scale()
    f = np.zeros((cfg.nall, 3))
    facc = np.zeros((cfg.nall, 3), dtype=object)       # exact integers
    eatom = np.zeros(cfg.nall)
    off = cfg.offsets
    for ii, i in enumerate(cfg.ilist):
        js = cfg.neigh[off[ii]:off[ii + 1]] & 0x1FFFFFFF
        d = cfg.x[i] - cfg.x[js]
        rsq = (d * d).sum(axis=1)
        keep = ~((rsq > Rc * Rc) | (rsq < 1.0e-12))
        js, d, r = js[keep], d[keep], np.sqrt(rsq[keep])
        N = len(js)
        u = d / r[:, None]
        fc = 0.5 * (np.cos(np.pi * r / Rc) + 1.0)
        dfc = -0.5 * np.pi / Rc * np.sin(np.pi * r / Rc)
        # ---- forward
        Grad = (_cheb(2.0 * r / Rc - 1.0, npsf) * fc).sum(axis=1) if N else np.zeros(npsf)
        jj, kk = np.triu_indices(N, 1)
        z = (u[jj] * u[kk]).sum(axis=1)
        w = fc[jj] * fc[kk]
        Tz = _cheb(z, 4 * ((ntsf + 3) // 4))
        S = np.array([(w * Tz[4 * (n // 4)] * z ** (n % 4)).sum() for n in range(ntsf)])
        Gang = blk2cheb.T @ S
        G = np.concatenate([Grad, Gang])
        G = s * G - s * pot.sfnor_avg
        out, dEdG = _mlp(pot, G)
        eatom[i] = pot.e_scale * out + pot.e_shift + pot.e_atom
        # ---- backward
        c = s * dEdG
        a = cheb2mono @ c[npsf:]                         # monomial coefficients of A(z)
        A = np.polynomial.polynomial.polyval(z, a)
        dA = np.polynomial.polynomial.polyval(z, a[1:] * np.arange(1, ntsf)) if ntsf > 1 else np.zeros_like(z)
        Pw = dA * w
        V = np.zeros((N, 3))
        Aa = np.zeros(N)
        np.add.at(V, jj, Pw[:, None] * u[kk])
        np.add.at(V, kk, Pw[:, None] * u[jj])
        np.add.at(Aa, jj, A * fc[kk])
        np.add.at(Aa, kk, A * fc[jj])
        x = 2.0 * r / Rc - 1.0
        cr = c[:npsf]
        Rv = np.polynomial.chebyshev.chebval(x, cr)
        Rp = np.polynomial.chebyshev.chebval(x, np.polynomial.chebyshev.chebder(cr)) if npsf > 1 else np.zeros_like(x)
        g = -(Rp * 2.0 / Rc * fc + Rv * dfc) - dfc * Aa + (u * V).sum(axis=1) / r
        grad = g[:, None] * u - V / r[:, None]           # d out / d x_j
        F = -pot.e_scale * grad                          # pair_annp.cpp:197
        if fixed_point:
            Fi = np.rint(F * 2.0 ** FIX_BITS)
            for t, j in enumerate(js):
                for k in range(3):
                    facc[j, k] += int(Fi[t, k])
        else:
            np.add.at(f, js, F)
        f[i] -= F.sum(axis=0)
    if fixed_point:
        f += np.array(facc, dtype=np.float64) * 2.0 ** -FIX_BITS
    return f, eatom

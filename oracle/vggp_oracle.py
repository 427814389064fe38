"""CPU oracle for the gridded-inducing-point ELBO hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module, and only as the checker / the timed CPU baseline.  The product package never imports it.

Pinning status: the *literal* half of this file (collapsed bound, `kuu_*`, `kuf_*`, `q_u`) is pinned against
fixtures produced by running the reference's own, unmodified model code (tests/golden/reference_models.npz,
made by oracle/make_golden.py on top of oracle/shim).  The reference itself ships no tests / golden vectors
(SURVEY.md §8c), so the pin is "reference code executed here", not "reference-published numbers"; the
gpytorch semantics the shim encodes are recalled, not verified (oracle/shim/README.md).  The B1 stencil is
pinned against the reference's `src/basis/bspline.py` executed as is (tests/golden/b1_stencil.npz).

Everything is pure torch on the CPU, float64 unless stated.  Reference citations are relative to
/root/reference/src.
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import torch

B1_ASVGP = 0      # B1-spline (hat) features, tridiagonal RKHS Kuu      (GriddedMatern12ASVGP, Matern12B1SplineASVGP)
B0_GRIDDED = 1    # cell-integrated Matern-1/2 features, Toeplitz Kuu     (Matern12GriddedGP, Matern12B0SplineGriddedGP)
SVGP_GRID = 2     # inducing points on a product grid, kernel Kuu / Kuf       (kronecker_structure.py:287-338 Matern12SVGP)
VFF_GRID = 3      # variational Fourier features, diagonal + rank-one Kuu    (kronecker_structure.py:347-514 Matern12VFFGP)


# ----------------------------------------------------------------------------------------------------------
# gpytorch parameter semantics (assumed: softplus constraints, raw init 0, noise lower bound 1e-4)
# ----------------------------------------------------------------------------------------------------------
NOISE_LOWER_BOUND = 1e-4


def softplus(raw: torch.Tensor) -> torch.Tensor:
    return torch.nn.functional.softplus(raw)


def constrain(raw_l: torch.Tensor, raw_s2: torch.Tensor, raw_noise: torch.Tensor):
    """raw (D,), (D,), () -> lengthscale (D,), outputscale (D,), noise ()."""
    return softplus(raw_l), softplus(raw_s2), softplus(raw_noise) + NOISE_LOWER_BOUND


# ----------------------------------------------------------------------------------------------------------
# meshes (models/sparse/gridded_kronecker_structure.py:707-720, 1278-1279): float32, torch.linspace
# ----------------------------------------------------------------------------------------------------------
def make_mesh(lo: float, hi: float, n_knots: int) -> torch.Tensor:
    return torch.linspace(lo, hi, n_knots)          # float32 on purpose (reference passes no dtype)


def make_padded_mesh(lo: float, hi: float, n_b0_splines: int, padding_factor: int) -> torch.Tensor:
    b0 = torch.linspace(lo, hi, n_b0_splines + 1)
    d = b0[1] - b0[0]
    left = torch.tensor([(b0[0] - (i * d)).item() for i in range(padding_factor, 0, -1)])
    right = torch.tensor([(b0[-1] + (i * d)).item() for i in range(1, padding_factor + 1)])
    return torch.cat((left, b0, right))


# ----------------------------------------------------------------------------------------------------------
# B1 stencil (basis/bspline.py:23-77, 92-94, 106-112)
# ----------------------------------------------------------------------------------------------------------
def b1_stencil(mesh: torch.Tensor, x: torch.Tensor):
    """Per-observation form of B1SplineBasis.__call__.

    Returns (c, w_lo, w_hi): the column of Phi(x) has w_lo at row c and w_hi at row c+1, zeros elsewhere;
    c == -1 marks an observation outside [mesh[0], mesh[-1]] (all-zero column; w_lo = w_hi = 0).
    x on an interior knot t_j belongs to the cell on its left (c = j-1, w_hi = 1); x == t_0 gives c = 0,
    w_lo = 1.  The denominator is the float32-rounded knot difference (0-dim float32 arithmetic in the
    reference), the numerator is computed in x's dtype.
    """
    K = mesh.numel()
    xd = x.contiguous()
    idx = torch.searchsorted(mesh, xd, right=False)            # first knot >= x
    inside = (xd >= mesh[0]) & (xd <= mesh[-1])
    c = (idx - 1).clamp(0, K - 2)
    t_lo = mesh[c]
    t_hi = mesh[c + 1]
    denom = (t_hi - t_lo).to(x.dtype)                          # float32 subtraction, then promoted
    w_hi = (xd - t_lo.to(x.dtype)) / denom
    w_lo = (t_hi.to(x.dtype) - xd) / denom
    zero = torch.zeros((), dtype=x.dtype)
    w_hi = torch.where(inside, w_hi, zero)
    w_lo = torch.where(inside, w_lo, zero)
    c = torch.where(inside, c, torch.full_like(c, -1))
    return c, w_lo, w_hi


def b1_features_dense(mesh: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """(K, N) dense feature matrix, numerically equal to B1SplineBasis(mesh)(x)."""
    K, N = mesh.numel(), x.numel()
    c, w_lo, w_hi = b1_stencil(mesh, x)
    phi = torch.zeros(K, N, dtype=x.dtype)
    cols = torch.arange(N)
    ok = c >= 0
    phi[c[ok], cols[ok]] += w_lo[ok]
    phi[c[ok] + 1, cols[ok]] += w_hi[ok]
    return phi


# ----------------------------------------------------------------------------------------------------------
# B0-integrated features (gridded_kronecker_structure.py:1325-1374)
# ----------------------------------------------------------------------------------------------------------
def b0_features_dense(mesh: torch.Tensor, x: torch.Tensor, l: torch.Tensor, s2: torch.Tensor) -> torch.Tensor:
    """Cov[int_cell_i f, f(x)] for the Matern-1/2 kernel; (K-1, N)."""
    m = mesh.numel() - 1
    k = torch.arange(m)
    indicator = -torch.sign(torch.searchsorted(mesh, x, right=False) - k[:, None] - 1)
    exp_1 = l * torch.exp(-torch.abs(x - mesh[:-1, None]) / l)
    exp_2 = l * torch.exp(-torch.abs(x - mesh[1:, None]) / l)
    out = indicator * (exp_1 - exp_2)
    inside = indicator == 0
    out = torch.where(inside, 2 * l - (exp_1 + exp_2), out)
    return out * s2


# ----------------------------------------------------------------------------------------------------------
# per-dimension Kuu factors
# ----------------------------------------------------------------------------------------------------------
def _sym_toeplitz(row: torch.Tensor) -> torch.Tensor:
    n = row.numel()
    idx = (torch.arange(n)[:, None] - torch.arange(n)[None, :]).abs()
    return row[idx]


def kuu_b0(mesh: torch.Tensor, l: torch.Tensor, s2: torch.Tensor) -> torch.Tensor:
    """gridded_kronecker_structure.py:1286-1323.  `(k +- 1) * delta` is rounded to float32 first (int64 tensor
    times 0-dim float32), then divided by the float64 lengthscale -- reproduced here by the same expressions."""
    m = mesh.numel() - 1
    delta = mesh[1] - mesh[0]                                   # 0-dim float32
    k = torch.arange(m)
    lv = l.reshape(1)
    first_row = (torch.exp((-(k - 1) * delta) / lv) + torch.exp((-(k + 1) * delta) / lv)
                 - 2 * torch.exp((-k * delta) / lv))
    diag0 = 2 * (torch.exp(-delta / lv) + (delta / lv) - 1)
    first_row = torch.cat([diag0.reshape(1), first_row[1:]])
    return _sym_toeplitz(first_row) * (lv ** 2 * s2)


def kuu_b1(mesh: torch.Tensor, l: torch.Tensor, s2: torch.Tensor, ref_quirks: bool = True) -> torch.Tensor:
    """gridded_kronecker_structure.py:731-780: (l*A + B/l + BC) / (2 s2), A/B tridiagonal Toeplitz with end
    corrections.  With ref_quirks the whole expression is evaluated in float32 exactly as the reference does
    (float32 matrices times 0-dim float64 scalars stay float32); the float32 factor is returned and the caller
    casts to float64 only AFTER torch.kron (:810-811), so the reference's Kuu is not an exact Kronecker product
    (each entry is fl32(K1[i,j]*K2[k,l])).  Without ref_quirks the float32 knot spacing is promoted once and
    everything is float64 -- the "intended semantics" the structured / CUDA path implements."""
    n = mesh.numel()
    delta = mesh[1] - mesh[0]                                   # 0-dim float32
    if not ref_quirks:
        delta = delta.to(torch.float64)
    dt = delta.dtype

    def tri(d0, d1, corr):
        row = torch.zeros(n, dtype=dt)
        row[0] = d0
        row[1] = d1
        bc = torch.zeros(n, dtype=dt)
        bc[0] = corr
        bc[-1] = corr
        return _sym_toeplitz(row) + torch.diag_embed(bc)

    A = tri(2 / 3 * delta, 1 / 6 * delta, -(1 / 3 * delta))
    B = tri(2 / delta, -1 / delta, -(1 / delta))
    bc = torch.zeros(n, dtype=dt)
    bc[0] = 1.0
    bc[-1] = 1.0
    BC = torch.diag_embed(bc)
    ls = l.reshape(())
    sc = s2.reshape(())
    return (A.mul(ls) + B.mul(1 / ls) + BC).mul(1 / (2 * sc))        # float32 under ref_quirks (cast after kron)


def kuu_svgp(z, l, s2):
    """Per-dimension factor of Matern12SVGP._Kuu (kronecker_structure.py:321-322): `self.kernel_d(self.Z).evaluate()` with
    kernel_d = ScaleKernel(MaternKernel(nu = 1/2, active_dims = [d])) (:30-31), i.e. s2 exp(-|z_i - z_j| / l) at the inducing
    locations `z` (float32 in the reference's parameter, promoted here)."""
    zz = z.to(torch.float64)
    return s2.reshape(()) * torch.exp(-(zz[:, None] - zz[None, :]).abs() / l.reshape(()))


def svgp_features_dense(z, x, l, s2):
    """Per-dimension factor of Matern12SVGP._Kuf (kronecker_structure.py:337-338): `self.kernel(full_Z, x).evaluate()` with the
    product kernel k1 * k2 (:32) on the Cartesian product of the per-dimension inducing locations factorises into
    prod_d s2_d exp(-|z_{i_d} - x_d| / l_d); this is the d-th factor, an (M_d, N) matrix."""
    zz = z.to(torch.float64)
    return s2.reshape(()) * torch.exp(-(zz[:, None] - x.to(torch.float64)[None, :]).abs() / l.reshape(()))


def vff_domain(mesh):
    """The library describes a VFF dimension by a float32 mesh of 2 M + 1 knots spanning the domain [a, b]: only its end points
    and its size (the number of features: M + 1 cosines, M sines) are used."""
    n = int(mesh.numel())
    assert n % 2 == 1 and n >= 3, "a VFF mesh has 2 * nfrequencies + 1 knots"
    return float(mesh[0]), float(mesh[-1]), (n - 1) // 2


def kuu_vff(mesh, l, s2, ref_quirks=True):
    """Matern12VFFGP._Kuu_along_dim (kronecker_structure.py:447-462): DiagLinearOperator(alpha).add_low_rank(beta) with
    alpha = (b - a) / 2 [2 / S(0), 1 / S(w_1..w_M), 1 / S(w_1..w_M)] (:403-420), S(w) = 2 s2 lambda / (lambda^2 + w^2), lambda = 1 / l
    (:373-392), beta = [1 / sqrt(s2) (M + 1 times), 0 (M times)] (:422-445).  ref_quirks: the reference's frequencies are a
    float32 tensor (fourier.py:13), which makes alpha, beta and hence Kuu float32 before the final cast (:462 'TODO: this is also
    wrong')."""
    a, b, Mf = vff_domain(mesh)
    wd = torch.float32 if ref_quirks else torch.float64
    omegas = ((2 * torch.pi) * torch.arange(Mf + 1, dtype=wd) / (b - a)).to(wd)       # fourier.py:13 (float32 there)
    lmbda = 1 / l.reshape(())
    num = 2 * s2.reshape(()) * lmbda                                    # float64
    den = (lmbda ** 2).to(wd) + omegas ** 2
    sd = (num.to(wd) / den) if ref_quirks else num / den
    S_inv = 1 / sd
    alpha = ((b - a) / 2) * torch.cat([2 * S_inv[0][None], S_inv[1:], S_inv[1:]])
    beta = torch.cat([torch.ones(Mf + 1, dtype=wd) / s2.reshape(()).sqrt().to(wd), torch.zeros(Mf, dtype=wd)])
    return (torch.diag_embed(alpha) + beta[:, None] * beta[None, :]).to(torch.float64)


def vff_features_dense(mesh, x, l, s2, ref_quirks=True):
    """Matern12VFFGP._Kuf_along_dim (kronecker_structure.py:464-480) = FourierBasisMatern12(nfrequencies, a, b, l)(x)
    (fourier.py:58-88): inside the domain a <= x < b the M + 1 cosines cos(w_k (x - a)) followed by the M sines sin(w_k (x - a));
    outside exp(-r / l) with r the distance to the nearer end for the cosine rows and zero for the sine rows.  They do not depend
    on s2.  ref_quirks: the products with the float32 frequencies are float32 (the 0-dim float64 coordinate is cast), as is the
    outside value once it multiplies torch.ones(M + 1)."""
    a, b, Mf = vff_domain(mesh)
    wd = torch.float32 if ref_quirks else torch.float64
    omegas = ((2 * torch.pi) * torch.arange(Mf + 1, dtype=wd) / (b - a)).to(wd)       # fourier.py:13 (float32 there)
    xx = x.to(torch.float64)
    inside = (xx >= a) & (xx < b)
    t = (xx - a).to(wd)
    cosr = torch.cos(omegas[:, None] * t[None, :])
    sinr = torch.sin(omegas[1:, None] * t[None, :])
    r = torch.minimum((xx - a).abs(), (xx - b).abs())
    out_val = torch.exp(-(1 / l.reshape(())) * r).to(wd)
    real = torch.where(inside[None, :], cosr, out_val[None, :].expand(Mf + 1, -1))
    imag = torch.where(inside[None, :], sinr, torch.zeros((), dtype=wd))
    return torch.cat([real, imag], dim=0).to(torch.float64)


def kuu_factor(family: int, mesh, l, s2, ref_quirks=True):
    if family == VFF_GRID:
        return kuu_vff(mesh, l, s2, ref_quirks)
    if family == SVGP_GRID:
        return kuu_svgp(mesh, l, s2)
    return kuu_b1(mesh, l, s2, ref_quirks) if family == B1_ASVGP else kuu_b0(mesh, l, s2)


def features_dense(family: int, mesh, x, l, s2, ref_quirks=False):
    if family == VFF_GRID:
        return vff_features_dense(mesh, x, l, s2, ref_quirks)
    if family == SVGP_GRID:
        return svgp_features_dense(mesh, x, l, s2)
    return b1_features_dense(mesh, x) if family == B1_ASVGP else b0_features_dense(mesh, x, l, s2)


def n_inducing(family: int, mesh) -> int:
    return mesh.numel() - 1 if family == B0_GRIDDED else mesh.numel()


def khatri_rao(feats: Sequence[torch.Tensor]) -> torch.Tensor:
    """Row-wise Khatri-Rao, first dimension slowest (gridded_kronecker_structure.py:827, 1406)."""
    out = feats[0]
    for f in feats[1:]:
        out = (out[:, None, :] * f[None, :, :]).reshape(-1, f.shape[-1])
    return out


def kron_all(mats: Sequence[torch.Tensor]) -> torch.Tensor:
    out = mats[0]
    for k in mats[1:]:
        out = torch.kron(out, k)
    return out


# ----------------------------------------------------------------------------------------------------------
# literal collapsed bound (kronecker_structure.py:249-278, univariate_structure.py:234-263)
# ----------------------------------------------------------------------------------------------------------
def _mvn_logprob_zero_mean(cov: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    L = torch.linalg.cholesky(cov)
    sol = torch.cholesky_solve(y.unsqueeze(-1), L)
    return -0.5 * ((y.unsqueeze(-1) * sol).sum() + 2.0 * torch.log(torch.diagonal(L)).sum()
                   + y.numel() * math.log(2 * math.pi))


def dense_Kuu_Kuf(family, meshes, X, l, s2, ref_quirks=True):
    D = len(meshes)
    Xc = X.reshape(X.shape[0], D)
    Ks = [kuu_factor(family, meshes[d], l[d], s2[d], ref_quirks) for d in range(D)]
    Fs = [features_dense(family, meshes[d], Xc[:, d], l[d], s2[d], ref_quirks) for d in range(D)]
    return kron_all(Ks).to(torch.float64), khatri_rao(Fs).to(torch.float64), Ks, Fs


def elbo_collapsed_literal(family, meshes, X, y, l, s2, noise, ref_quirks=True) -> torch.Tensor:
    """The reference's `_elbo`: dense N x N evidence covariance, Cholesky semantics."""
    Kuu, Kuf, _, _ = dense_Kuu_Kuf(family, meshes, X, l, s2, ref_quirks)
    N = y.numel()
    Lk = torch.linalg.cholesky(Kuu)
    approx_prior = Kuf.T @ torch.cholesky_solve(Kuf, Lk)
    # reference quirk (kronecker_structure.py:272): `noise * torch.eye(N)` is a float32 matrix, i.e. the noise on
    # the diagonal of the evidence covariance is rounded to float32 (2.7e-8 relative) before the float64 add
    eye = torch.eye(N) if ref_quirks else torch.eye(N, dtype=torch.float64)
    cov = approx_prior + noise * eye
    evidence = _mvn_logprob_zero_mean(cov, y)
    kff = torch.prod(s2)                                       # diagonal of the product Matern kernel
    trace_term = (N * kff - torch.trace(approx_prior)) / (2 * noise)
    return evidence - trace_term


def elbo_collapsed_woodbury(family, meshes, X, y, l, s2, noise, ref_quirks=True) -> torch.Tensor:
    """Same bound through M x M algebra."""
    Kuu, Kuf, _, _ = dense_Kuu_Kuf(family, meshes, X, l, s2, ref_quirks)
    N = y.numel()
    A = Kuf @ Kuf.T
    b = Kuf @ y
    Sigma = Kuu + A / noise
    Ls = torch.linalg.cholesky(Sigma)
    Lk = torch.linalg.cholesky(Kuu)
    logdet = 2 * torch.log(torch.diagonal(Ls)).sum() - 2 * torch.log(torch.diagonal(Lk)).sum() + N * torch.log(noise)
    quad = (y @ y - (b @ torch.cholesky_solve(b.unsqueeze(-1), Ls).squeeze(-1)) / noise) / noise
    evidence = -0.5 * (quad + logdet + N * math.log(2 * math.pi))
    kff = torch.prod(s2)
    trQ = torch.trace(torch.cholesky_solve(A, Lk))
    return evidence - (N * kff - trQ) / (2 * noise)


def optimal_q(family, meshes, X, y, l, s2, noise, ref_quirks=True):
    """q_u()/q_v(): m* = Kuu Sigma^-1 Kuf y / noise, S* = Kuu Sigma^-1 Kuu
    (gridded_kronecker_structure.py:903-916, 1409-1433)."""
    Kuu, Kuf, _, _ = dense_Kuu_Kuf(family, meshes, X, l, s2, ref_quirks)
    Sigma = Kuu + (Kuf @ Kuf.T) / noise
    Ls = torch.linalg.cholesky(Sigma)
    mean = (Kuu @ torch.cholesky_solve(Kuf, Ls) @ y) / noise
    cov = Kuu @ torch.cholesky_solve(Kuu, Ls)
    return mean, cov


# ----------------------------------------------------------------------------------------------------------
# uncollapsed bound, dense S (bridge between the literal bound and the structured path)
# ----------------------------------------------------------------------------------------------------------
def elbo_uncollapsed_dense(family, meshes, X, y, l, s2, noise, m, S, ref_quirks=True, scale=1.0):
    Kuu, Kuf, _, _ = dense_Kuu_Kuf(family, meshes, X, l, s2, ref_quirks)
    M = Kuu.shape[0]
    Lk = torch.linalg.cholesky(Kuu)
    alpha = torch.cholesky_solve(m.unsqueeze(-1), Lk).squeeze(-1)
    mu = Kuf.T @ alpha
    KinvKuf = torch.cholesky_solve(Kuf, Lk)
    kff = torch.prod(s2)
    var = kff - (Kuf * KinvKuf).sum(0) + (KinvKuf * (S @ KinvKuf)).sum(0)
    ell = (-0.5 * torch.log(2 * math.pi * noise) - ((y - mu) ** 2 + var) / (2 * noise)).sum()
    Ls = torch.linalg.cholesky(S)
    kl = 0.5 * (torch.trace(torch.cholesky_solve(S, Lk)) + m @ alpha - M
                + 2 * torch.log(torch.diagonal(Lk)).sum() - 2 * torch.log(torch.diagonal(Ls)).sum())
    return scale * ell - kl


# ----------------------------------------------------------------------------------------------------------
# structured uncollapsed bound: q(u) = N(m, kron_d L_d L_d^T)   (north-star path)
# ----------------------------------------------------------------------------------------------------------
def mode_product(T: torch.Tensor, A: torch.Tensor, d: int) -> torch.Tensor:
    """Apply A along mode d of the D-way tensor T."""
    return torch.movedim(torch.tensordot(A, T, dims=([1], [d])), 0, d)


def elbo_structured(family, meshes, X, y, l, s2, noise, m, Ls: Sequence[torch.Tensor],
                    ref_quirks=True, scale=1.0, work_dtype=torch.float64):
    """D-generic.  m: (M,) row-major over (M_1..M_D); Ls[d]: (M_d, M_d), lower triangle used.

    B1 family: 2^D-point gather per observation; B0 family: dense per-dimension features.
    `work_dtype` is the arithmetic type of the per-observation part (float32 reproduces the fp32 configs'
    arithmetic on the CPU for the timed baseline; parity tests use float64)."""
    D = len(meshes)
    N = y.numel()
    Xc = X.reshape(N, D)
    Ms = [n_inducing(family, meshes[d]) for d in range(D)]
    M = int(torch.tensor(Ms).prod())
    Ks = [kuu_factor(family, meshes[d], l[d], s2[d], ref_quirks).to(torch.float64) for d in range(D)]
    Cs = [torch.linalg.cholesky(K) for K in Ks]
    Ps = [torch.cholesky_inverse(C) for C in Cs]
    Lt = [torch.tril(L) for L in Ls]
    Rs = [P @ L for P, L in zip(Ps, Lt)]
    Qs = [R @ R.T for R in Rs]
    alpha = m.reshape(Ms)
    for d in range(D):
        alpha = mode_product(alpha, Ps[d], d)
    kff = torch.prod(s2)
    wd = work_dtype

    if family == B1_ASVGP:
        mu = torch.zeros(N, dtype=wd)
        ps, qs = [], []
        sten = [b1_stencil(meshes[d], Xc[:, d]) for d in range(D)]
        inside = torch.ones(N, dtype=torch.bool)
        for c, _, _ in sten:
            inside &= c >= 0
        a_flat = alpha.reshape(-1).to(wd)
        strides = [int(torch.tensor(Ms[d + 1:]).prod()) if d + 1 < D else 1 for d in range(D)]
        for corner in range(2 ** D):
            w = torch.ones(N, dtype=wd)
            idx = torch.zeros(N, dtype=torch.long)
            for d in range(D):
                hi = (corner >> (D - 1 - d)) & 1
                c, w_lo, w_hi = sten[d]
                w = w * (w_hi if hi else w_lo).to(wd)
                idx = idx + (c.clamp(min=0) + hi) * strides[d]
            mu = mu + torch.where(inside, w * a_flat[idx], torch.zeros((), dtype=wd))
        for d in range(D):
            c, w_lo, w_hi = sten[d]
            cc = c.clamp(min=0)
            w_lo, w_hi = w_lo.to(wd), w_hi.to(wd)
            for T, acc in ((Ps[d], ps), (Qs[d], qs)):
                dg = torch.diagonal(T).to(wd)
                od = torch.diagonal(T, 1).to(wd)
                acc.append(w_lo * w_lo * dg[cc] + 2 * w_lo * w_hi * od[cc] + w_hi * w_hi * dg[cc + 1])
        p = torch.stack(ps).prod(0)
        q = torch.stack(qs).prod(0)
    else:
        Fs = [features_dense(family, meshes[d], Xc[:, d], l[d], s2[d]).to(wd) for d in range(D)]
        t = alpha.to(wd)
        # contract modes one at a time: t[(i_1..i_D)] with Phi_d[i_d, n]
        t = torch.tensordot(Fs[0].T, t, dims=([1], [0]))                 # (N, M_2..M_D)
        for d in range(1, D):
            t = (t * Fs[d].T.reshape([N, Ms[d]] + [1] * (D - 1 - d))).sum(1)
        mu = t
        p = torch.ones(N, dtype=wd)
        q = torch.ones(N, dtype=wd)
        for d in range(D):
            p = p * (Fs[d] * (Ps[d].to(wd) @ Fs[d])).sum(0)
            q = q * (Fs[d] * (Qs[d].to(wd) @ Fs[d])).sum(0)
    var = kff.to(wd) - p + q
    yw = y.to(wd)
    nz = noise.to(wd)
    ell = (-0.5 * torch.log(2 * math.pi * nz) - ((yw - mu) ** 2 + var) / (2 * nz)).sum().to(torch.float64)

    tr = torch.ones((), dtype=torch.float64)
    logdets = torch.zeros((), dtype=torch.float64)
    for d in range(D):
        tr = tr * (Rs[d] * Lt[d]).sum()
        logdets = logdets + (M / Ms[d]) * (2 * torch.log(torch.diagonal(Cs[d])).sum()
                                           - 2 * torch.log(torch.diagonal(Lt[d]).abs()).sum())
    kl = 0.5 * (tr + (m * alpha.reshape(-1)).sum() - M + logdets)
    return scale * ell - kl


def kron_cov_from_factors(Ls: Sequence[torch.Tensor]) -> torch.Tensor:
    return kron_all([torch.tril(L) @ torch.tril(L).T for L in Ls])

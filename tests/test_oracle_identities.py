"""Identities that tie the structured (north-star) bound to the reference's collapsed bound.  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import vggp_oracle as O


def _data(N, D, seed=0, lo=0.0, hi=1.0):
    g = torch.Generator().manual_seed(seed)
    X = torch.rand(N, D, generator=g, dtype=torch.float64) * (hi - lo) + lo
    y = torch.sin(5 * X[:, 0]) + (torch.cos(7 * X[:, 1]) if D > 1 else 0) + 0.05 * torch.randn(N, generator=g, dtype=torch.float64)
    return X, y


def _nk(family, n):
    """knot count for a test mesh: a VFF mesh has 2 * nfrequencies + 1 knots"""
    return n + 1 - n % 2 if family == O.VFF_GRID else n


def _hyp(D, seed=1):
    g = torch.Generator().manual_seed(seed)
    rl = (torch.randn(D, generator=g, dtype=torch.float64) * 0.5).requires_grad_()
    rs = (torch.randn(D, generator=g, dtype=torch.float64) * 0.5).requires_grad_()
    rn = (torch.randn((), generator=g, dtype=torch.float64) * 0.5 - 1.0).requires_grad_()
    return rl, rs, rn


@pytest.mark.parametrize("family", [O.B1_ASVGP, O.B0_GRIDDED, O.SVGP_GRID, O.VFF_GRID])
@pytest.mark.parametrize("D", [1, 2])
def test_literal_equals_woodbury(family, D):
    X, y = _data(150, D)
    meshes = [O.make_mesh(0, 1, _nk(family, 7 + d)) for d in range(D)]
    rl, rs, rn = _hyp(D)
    l, s2, nz = O.constrain(rl, rs, rn)
    a = O.elbo_collapsed_literal(family, meshes, X, y, l, s2, nz, ref_quirks=False)
    b = O.elbo_collapsed_woodbury(family, meshes, X, y, l, s2, nz, ref_quirks=False)
    assert a.item() == pytest.approx(b.item(), rel=1e-11)
    ga = torch.autograd.grad(a, [rl, rs, rn], retain_graph=True)
    gb = torch.autograd.grad(b, [rl, rs, rn])
    for u, v in zip(ga, gb):
        assert torch.allclose(u, v, rtol=1e-8, atol=1e-9)


@pytest.mark.parametrize("family", [O.B1_ASVGP, O.B0_GRIDDED, O.SVGP_GRID, O.VFF_GRID])
@pytest.mark.parametrize("D", [1, 2])
def test_uncollapsed_at_optimum_equals_collapsed(family, D):
    """ELBO(m*, S*) == collapsed bound, and (envelope theorem) so are the hyper-parameter gradients."""
    X, y = _data(120, D)
    meshes = [O.make_mesh(0, 1, _nk(family, 6 + d)) for d in range(D)]
    rl, rs, rn = _hyp(D)
    l, s2, nz = O.constrain(rl, rs, rn)
    a = O.elbo_collapsed_literal(family, meshes, X, y, l, s2, nz, ref_quirks=False)
    with torch.no_grad():
        m, S = O.optimal_q(family, meshes, X, y, l, s2, nz, ref_quirks=False)
        S = (S + S.T) / 2
    b = O.elbo_uncollapsed_dense(family, meshes, X, y, l, s2, nz, m, S, ref_quirks=False)
    assert a.item() == pytest.approx(b.item(), rel=1e-11)
    ga = torch.autograd.grad(a, [rl, rs, rn], retain_graph=True)
    gb = torch.autograd.grad(b, [rl, rs, rn])
    for u, v in zip(ga, gb):
        assert torch.allclose(u, v, rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("family", [O.B1_ASVGP, O.B0_GRIDDED, O.SVGP_GRID, O.VFF_GRID])
@pytest.mark.parametrize("D,knots", [(1, [9]), (2, [9, 7]), (3, [5, 4, 6])])
def test_structured_equals_dense_uncollapsed(family, D, knots):
    """G4: Kronecker-factored q(u) through mode products == dense M x M algebra (values and all gradients)."""
    X, y = _data(200, D, seed=3, lo=-0.05, hi=1.05)          # some observations outside the mesh
    if D == 3:
        y = y + torch.sin(3 * X[:, 2])
    meshes = [O.make_mesh(0, 1, _nk(family, k)) for k in knots]
    Ms = [O.n_inducing(family, mh) for mh in meshes]
    M = int(np.prod(Ms))
    rl, rs, rn = _hyp(D, seed=5)
    g = torch.Generator().manual_seed(7)
    m = (torch.randn(M, generator=g, dtype=torch.float64) * 0.3).requires_grad_()
    Ls = [(torch.eye(k, dtype=torch.float64) + 0.1 * torch.tril(torch.randn(k, k, generator=g, dtype=torch.float64))).requires_grad_()
          for k in Ms]
    l, s2, nz = O.constrain(rl, rs, rn)
    a = O.elbo_structured(family, meshes, X, y, l, s2, nz, m, Ls, ref_quirks=False, scale=1.7)
    S = O.kron_cov_from_factors(Ls)
    b = O.elbo_uncollapsed_dense(family, meshes, X, y, l, s2, nz, m, S, ref_quirks=False, scale=1.7)
    assert a.item() == pytest.approx(b.item(), rel=1e-11)
    ga = torch.autograd.grad(a, [rl, rs, rn, m] + Ls, retain_graph=True)
    gb = torch.autograd.grad(b, [rl, rs, rn, m] + Ls)
    for u, v in zip(ga[:4], gb[:4]):
        assert torch.allclose(u, v, rtol=1e-7, atol=1e-8)
    for u, v in zip(ga[4:], gb[4:]):
        assert torch.allclose(torch.tril(u), torch.tril(v), rtol=1e-7, atol=1e-8)


def test_structured_1d_at_optimum_reproduces_reference_golden(golden_dir):
    """D = 1: S = L L^T is unrestricted, so the structured bound at (m*, chol S*) must reproduce the value the
    reference's own `_elbo()` returned (fixture G3, gridded_univariate_structure.Matern12GriddedGP)."""
    ref = np.load(os.path.join(golden_dir, "reference_models.npz"))
    x = torch.from_numpy(ref["g3.x"])
    y = torch.from_numpy(ref["g3.y"])
    meshes = [O.make_mesh(0., 2., 33)]
    for pset, ps in (("raw0", (0.0, 0.0, 0.0)), ("raw1", (-0.3, 0.5, -3.0))):
        rl = torch.tensor([ps[0]], dtype=torch.float64, requires_grad=True)
        rs = torch.tensor([ps[1]], dtype=torch.float64, requires_grad=True)
        rn = torch.tensor(ps[2], dtype=torch.float64, requires_grad=True)
        l, s2, nz = O.constrain(rl, rs, rn)
        key = f"G3_griddedgp1d.{pset}"
        m = torch.from_numpy(ref[key + ".q_mean"])
        S = torch.from_numpy(ref[key + ".q_cov"])
        L = torch.linalg.cholesky((S + S.T) / 2)
        e = O.elbo_structured(O.B0_GRIDDED, meshes, x, y, l, s2, nz, m, [L])
        assert e.item() == pytest.approx(float(ref[key + ".elbo"]), rel=2e-7)   # float32-eye quirk: 3e-8
        g = torch.autograd.grad(e, [rl, rs, rn])
        assert g[0].item() == pytest.approx(ref[key + ".grad.kernel.base_kernel.raw_lengthscale"][0], rel=1e-5, abs=1e-6)
        assert g[1].item() == pytest.approx(ref[key + ".grad.kernel.raw_outputscale"][0], rel=1e-5, abs=1e-6)
        assert g[2].item() == pytest.approx(ref[key + ".grad.likelihood.noise_covar.raw_noise"][0], rel=1e-5, abs=1e-6)

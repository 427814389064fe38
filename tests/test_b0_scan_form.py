"""The scan form of the B0 family (oracle/b0_scan.py) against the reference's dense features: same features, same
predictive mean and variance factors, hence (autograd) the same gradients with respect to the lengthscale,
outputscale, alpha, P_d and Q_d.  Design validation for DESIGN.md section 10; CPU only."""
import pytest
import torch

from oracle import b0_scan as S
from oracle import vggp_oracle as O


def problem(D, seed):
    g = torch.Generator().manual_seed(seed)
    meshes = [torch.linspace(0, 1, 12), (torch.cumsum(torch.rand(9, generator=g) + 0.3, 0)).to(torch.float32)][:D]
    N = 400
    X = torch.stack([(torch.rand(N, generator=g, dtype=torch.float64) * 1.4 - 0.2) * float(m[-1] - m[0]) + float(m[0])
                     for m in meshes], 1)
    for d in range(D):                                   # exact knot hits, first and last knot included
        X[: meshes[d].numel(), d] = meshes[d].to(torch.float64)
    l = (torch.rand(D, generator=g, dtype=torch.float64) * 0.3 + 0.05).requires_grad_(True)
    s2 = (torch.rand(D, generator=g, dtype=torch.float64) + 0.5).requires_grad_(True)
    Ms = [m.numel() - 1 for m in meshes]
    A = torch.randn(*Ms, generator=g, dtype=torch.float64).requires_grad_(True)
    Ps, Qs = [], []
    for n in Ms:
        B = torch.randn(n, n, generator=g, dtype=torch.float64)
        Ps.append((B @ B.T / n + torch.eye(n, dtype=torch.float64)).requires_grad_(True))
        C = torch.randn(n, n, generator=g, dtype=torch.float64)
        Qs.append((C @ C.T / n).requires_grad_(True))
    return meshes, X, l, s2, A, Ps, Qs


@pytest.mark.parametrize("D", [1, 2])
def test_scan_features_equal_reference_features(D):
    meshes, X, l, s2, *_ = problem(D, 1)
    for d in range(D):
        ref = O.b0_features_dense(meshes[d], X[:, d], l[d], s2[d])
        got = S.features_from_scan(meshes[d], X[:, d], l[d], s2[d])
        assert torch.allclose(got, ref, rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize("D", [1, 2])
def test_scan_mean_variance_and_gradients_equal_dense(D):
    meshes, X, l, s2, A, Ps, Qs = problem(D, 2)
    Fs = [O.b0_features_dense(meshes[d], X[:, d], l[d], s2[d]) for d in range(D)]
    if D == 1:
        mu_ref = Fs[0].T @ A
    else:
        mu_ref = torch.einsum("in,ij,jn->n", Fs[0], A, Fs[1])
    p_ref = torch.ones_like(mu_ref)
    q_ref = torch.ones_like(mu_ref)
    for d in range(D):
        p_ref = p_ref * (Fs[d] * (Ps[d] @ Fs[d])).sum(0)
        q_ref = q_ref * (Fs[d] * (Qs[d] @ Fs[d])).sum(0)
    mu, p, q = S.mean_p_q_scan(meshes, X, l, s2, A, Ps, Qs)
    assert torch.allclose(mu, mu_ref, rtol=1e-11, atol=1e-12)
    assert torch.allclose(p, p_ref, rtol=1e-11, atol=1e-12) and torch.allclose(q, q_ref, rtol=1e-11, atol=1e-12)
    w = torch.linspace(0.5, 1.5, mu.numel(), dtype=torch.float64)
    params = [l, s2, A] + Ps + Qs
    g_ref = torch.autograd.grad(((w * mu_ref) ** 2 - p_ref + q_ref).sum(), params, retain_graph=True)
    g = torch.autograd.grad(((w * mu) ** 2 - p + q).sum(), params)
    for a, b in zip(g, g_ref):
        assert torch.allclose(a, b, rtol=1e-9, atol=1e-10)

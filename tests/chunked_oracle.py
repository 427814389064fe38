"""Chunked evaluation of the oracle's structured bound (test infrastructure only).

O.elbo_structured keeps every per-observation temporary of its autograd graph alive, so one call cannot take the
BASELINE.json configurations (10^6 ... 2^26 observations).  The expected log-likelihood is a plain sum over observations
and the KL term does not depend on them, so with f(X_c) = scale * ELL(X_c) - KL

    ELBO(X) = sum_c f(X_c) - (n_chunks - 1) * f(empty)          (f(empty) = -KL)

and the same identity holds for every gradient.  The grid side (factors, Cholesky, inverse) is recomputed per chunk in
float64; nothing is approximated."""
import torch

from oracle import vggp_oracle as O


def elbo_and_grads_chunked(family, meshes, X, y, l, s2, noise, m, Ls, scale=1.0, chunk=1 << 20, ref_quirks=False):
    """Returns (elbo, [dl, ds2, dnoise, dm, dL_1..dL_D]) of O.elbo_structured over all of X (N, D) / y (N,), float64."""
    params = [l.clone().requires_grad_(True), s2.clone().requires_grad_(True), noise.clone().requires_grad_(True),
              m.clone().requires_grad_(True)] + [L.clone().requires_grad_(True) for L in Ls]
    N = y.numel()
    D = len(meshes)
    X = X.reshape(N, D)

    def f(Xc, yc):
        val = O.elbo_structured(family, meshes, Xc.to(torch.float64), yc.to(torch.float64), params[0], params[1], params[2],
                                params[3], params[4:], ref_quirks=ref_quirks, scale=scale)
        grads = torch.autograd.grad(val, params)
        return val.detach(), [g.detach() for g in grads]

    total, gtot, nchunks = None, None, 0
    for lo in range(0, max(N, 1), chunk):
        v, g = f(X[lo:lo + chunk], y[lo:lo + chunk])
        nchunks += 1
        if total is None:
            total, gtot = v, g
        else:
            total = total + v
            gtot = [a + b for a, b in zip(gtot, g)]
    if nchunks > 1:
        v0, g0 = f(X[:0], y[:0])
        total = total - (nchunks - 1) * v0
        gtot = [a - (nchunks - 1) * b for a, b in zip(gtot, g0)]
    return total, gtot

"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference code.

TEST INFRASTRUCTURE ONLY.  Runs in the build container (needs /root/reference); nothing on the GPU box
reads /root/reference, only the committed .npz/.json fixtures this script writes.

Two sources:
 (1) `src/basis/bspline.py` imports only torch -> imported as is (pure reference output);
 (2) `src/models/sparse/*.py` need gpytorch/linear_operator -> imported as is on top of oracle/shim
     (see oracle/shim/README.md for the assumptions the shim encodes).

Usage:  python oracle/make_golden.py            (writes tests/golden/*.npz)
"""
import os
import sys
import json
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
REF = os.environ.get("VGGP_REFERENCE", "/root/reference")


def _import_reference():
    sys.path.insert(0, os.path.join(HERE, "shim"))
    sys.path.insert(0, REF)
    from src.basis import bspline  # noqa
    from src.models.sparse import kronecker_structure as ks  # noqa
    from src.models.sparse import gridded_kronecker_structure as gks  # noqa
    from src.models.sparse import univariate_structure as us  # noqa
    from src.models.sparse import gridded_univariate_structure as gus  # noqa
    return bspline, ks, gks, us, gus


def latent_function_2d(x1, x2):
    # 5_gridded_kronecker_structure_models.ipynb cell 3
    return (np.sin(5 * x1) + np.cos(7 * x2) + 0.5 * np.sin(15 * x1) + 0.5 * np.cos(12 * x2)
            + 0.2 * np.sin(20 * x1) + 0.2 * np.cos(25 * x2))


def gen_2d(func, x1lims, x2lims, nobs):
    # src/utils/datagenerators.py:37-73 (evenly spaced branch)
    d1 = np.linspace(x1lims[0], x1lims[1], nobs)
    d2 = np.linspace(x2lims[0], x2lims[1], nobs)
    X1, X2 = np.meshgrid(d1, d2)
    X = np.vstack([X1.ravel(), X2.ravel()]).T
    return X, func(X[:, 0], X[:, 1])


def stencil_inputs(mesh: torch.Tensor, n_rand: int, seed: int, dtype):
    """x values exercising every edge: knots, knot neighbours (nextafter), outside, random."""
    g = torch.Generator().manual_seed(seed)
    lo, hi = float(mesh[0]), float(mesh[-1])
    span = hi - lo
    xr = (torch.rand(n_rand, generator=g, dtype=torch.float64) * 1.2 - 0.1) * span + lo
    knots = mesh.to(torch.float64)
    kn = knots.numpy()
    near = np.concatenate([np.nextafter(kn, -np.inf), np.nextafter(kn, np.inf)])
    if dtype == torch.float32:
        k32 = mesh.numpy().astype(np.float32)
        near = np.concatenate([np.nextafter(k32, np.float32(-np.inf)),
                               np.nextafter(k32, np.float32(np.inf))]).astype(np.float64)
    dec = torch.tensor([0.1 * i for i in range(-1, 13)], dtype=torch.float64) * span + lo
    x = torch.cat([xr, knots, torch.from_numpy(near), dec])
    return x.to(dtype)


def golden_stencils(bspline):
    """B1SplineBasis.__call__ (bspline.py:92-94) on the meshes the reference builds."""
    out = {}
    cases = {
        "lin11_01": torch.linspace(0, 1, 11),                       # non-uniform in float32
        "lin129_01": torch.linspace(0, 1, 129),                     # exactly uniform
        "lin16_02": torch.linspace(0, 2, 16),
        "lin21_m3_7": torch.linspace(-3.0, 7.0, 21),
    }
    # padded mesh exactly as gridded_kronecker_structure.py:707-720 (n_b0_splines=20, padding 2)
    b0 = torch.linspace(0.0, 1.0, 21)
    d = b0[1] - b0[0]
    left = torch.tensor([(b0[0] - (i * d)).item() for i in range(2, 0, -1)])
    right = torch.tensor([(b0[-1] + (i * d)).item() for i in range(1, 3)])
    cases["padded21_pad2"] = torch.cat((left, b0, right))
    for name, mesh in cases.items():
        basis = bspline.B1SplineBasis(mesh)
        for dt, tag in ((torch.float64, "f64"), (torch.float32, "f32")):
            x = stencil_inputs(mesh, 400, 1234, dt)
            phi = basis(x)                       # (K, N)
            assert phi.dtype == dt, (phi.dtype, dt)
            out[f"{name}.{tag}.mesh"] = mesh.numpy()
            out[f"{name}.{tag}.x"] = x.numpy()
            out[f"{name}.{tag}.phi"] = phi.numpy()
    np.savez_compressed(os.path.join(GOLD, "b1_stencil.npz"), **out)
    return len(cases)


def _grads(model, loss):
    names = [n for n, _ in model.named_parameters()]
    gs = torch.autograd.grad(loss, [p for _, p in model.named_parameters()])
    return {n: g.detach().reshape(-1).numpy() for n, g in zip(names, gs)}


def _set_params(model, spec):
    """spec: dict raw-parameter-name -> value (raw space)."""
    sd = dict(model.named_parameters())
    for k, v in spec.items():
        sd[k].data.fill_(v)


def golden_models(ks, gks, us, gus):
    out = {}
    meta = {}
    X, y = gen_2d(latent_function_2d, (0., 1.), (0., 1.), 25)
    Xt = torch.tensor(X).to(torch.float64)
    yt = torch.tensor(y).to(torch.float64)
    out["nb5.X"] = X
    out["nb5.y"] = y

    param_sets = {
        "raw0": {},
        "raw1": {"kernel_1.raw_outputscale": 0.3, "kernel_1.base_kernel.raw_lengthscale": -0.7,
                 "kernel_2.raw_outputscale": -0.2, "kernel_2.base_kernel.raw_lengthscale": 0.4,
                 "likelihood.noise_covar.raw_noise": -2.0},
    }

    def run2d(tag, ctor):
        for pname, spec in param_sets.items():
            model = ctor().to(torch.float64)
            _set_params(model, spec)
            elbo = model._elbo()
            g = _grads(model, elbo)
            key = f"{tag}.{pname}"
            out[key + ".elbo"] = np.array(elbo.item())
            for n, v in g.items():
                out[key + ".grad." + n] = v
            with torch.no_grad():
                out[key + ".Kuu"] = model._Kuu().numpy() if torch.is_tensor(model._Kuu()) else model._Kuu().to_dense().numpy()
                qd = model.q_u() if hasattr(model, "q_u") else (model.q_v() if hasattr(model, "q_v") else None)
                if qd is not None:
                    out[key + ".q_mean"] = qd.mean.numpy()
                    out[key + ".q_cov"] = qd.covariance_matrix.numpy()
            meta[key] = {"elbo": float(elbo.item())}

    # G1: Matern12GriddedGP (gridded_kronecker_structure.py:1255-1433), nb5 cells 24-26
    run2d("G1_griddedgp2d", lambda: gks.Matern12GriddedGP(Xt, yt, 11, (0, 1), (0, 1)))
    # G2q: GriddedMatern12ASVGP with the reference's float32-Kuu quirk (:685-969), pad 0 and pad 1
    run2d("G2q_asvgp2d_pad0", lambda: gks.GriddedMatern12ASVGP(Xt, yt, 10, 0, (0, 1), (0, 1)))
    run2d("G2q_asvgp2d_pad1", lambda: gks.GriddedMatern12ASVGP(Xt, yt, 10, 1, (0, 1), (0, 1)))
    # kronecker_structure.py twins (non-square grids are not expressible: nknots shared)
    run2d("K_b1asvgp2d", lambda: ks.Matern12B1SplineASVGP(Xt, yt, 9, (0, 1), (0, 1)))
    run2d("K_b0gridded2d", lambda: ks.Matern12B0SplineGriddedGP(Xt, yt, 9, (0, 1), (0, 1)))
    # product-grid SVGP (kronecker_structure.py:287-338): inducing points Z (m x 2), one column per dimension, the second
    # column non-uniform; Z is a parameter of the reference's class (its gradient is recorded, not used)
    Zs = torch.stack([torch.linspace(0, 1, 7), torch.linspace(0, 1, 7) ** 1.3], dim=1)
    out["svgp.Z"] = Zs.numpy()
    run2d("K_svgp2d", lambda: ks.Matern12SVGP(Xt, yt, Zs))
    # variational Fourier features (kronecker_structure.py:347-514): 4 frequencies per dimension; the second domain is smaller
    # than the data, so the outside-of-domain branch of the basis (fourier.py:64-75) is part of the fixture; limits exactly
    # representable in float32 (the library sees them as float32 knots)
    run2d("K_vff2d", lambda: ks.Matern12VFFGP(Xt, yt, 4, (-0.125, 1.125), (0.25, 0.75)))

    # 1-D: gridded_univariate_structure.Matern12GriddedGP (:709-844).  Its twin
    # univariate_structure.Matern12B0SplineGriddedGP (:721-825) builds a float32 Kuu (0-dim lengthscale) and then
    # mixes it with float64 Kuf inside gpytorch; what real gpytorch does with that mix is not verifiable offline,
    # so it is not used as a golden.
    g = torch.Generator().manual_seed(0)
    N1 = 600
    x1 = torch.rand(N1, generator=g, dtype=torch.float64) * 2.0
    y1 = torch.sin(x1) + torch.cos(x1) + 0.05 * torch.randn(N1, generator=g, dtype=torch.float64)
    out["g3.x"] = x1.numpy()
    out["g3.y"] = y1.numpy()
    param_sets_1d = {
        "raw0": {},
        "raw1": {"kernel.raw_outputscale": 0.5, "kernel.base_kernel.raw_lengthscale": -0.3,
                 "likelihood.noise_covar.raw_noise": -3.0},
    }
    for pname, spec in param_sets_1d.items():
        model = gus.Matern12GriddedGP(x1, y1, 32, (0., 2.)).to(torch.float64)
        _set_params(model, spec)
        elbo = model._elbo()
        gr = _grads(model, elbo)
        key = f"G3_griddedgp1d.{pname}"
        out[key + ".elbo"] = np.array(elbo.item())
        for n, v in gr.items():
            out[key + ".grad." + n] = v
        with torch.no_grad():
            qd = model.q_v()
            out[key + ".q_mean"] = qd.mean.numpy()
            out[key + ".q_cov"] = qd.covariance_matrix.numpy()
            out[key + ".Kuu"] = model._Kuu().numpy()
        meta[key] = {"elbo": float(elbo.item())}

    np.savez_compressed(os.path.join(GOLD, "reference_models.npz"), **out)
    with open(os.path.join(GOLD, "reference_models.json"), "w") as f:
        json.dump({"torch": torch.__version__, "cases": meta,
                   "note": "produced by oracle/make_golden.py from /root/reference through oracle/shim"}, f, indent=1)
    return meta


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.set_default_dtype(torch.float32)
    bspline, ks, gks, us, gus = _import_reference()
    n = golden_stencils(bspline)
    meta = golden_models(ks, gks, us, gus)
    print(f"wrote {n} stencil meshes, {len(meta)} model cases to {GOLD}")
    for k, v in meta.items():
        print(f"  {k:34s} ELBO = {v['elbo']:.10f}")


if __name__ == "__main__":
    main()

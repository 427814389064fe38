"""Oracle vs fixtures produced by the reference's own code (oracle/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import vggp_oracle as O

torch.set_default_dtype(torch.float32)


@pytest.fixture(scope="module")
def ref(golden_dir):
    return np.load(os.path.join(golden_dir, "reference_models.npz"))


@pytest.fixture(scope="module")
def sten(golden_dir):
    return np.load(os.path.join(golden_dir, "b1_stencil.npz"))


MESHES = ["lin11_01", "lin129_01", "lin16_02", "lin21_m3_7", "padded21_pad2"]


@pytest.mark.parametrize("name", MESHES)
@pytest.mark.parametrize("tag", ["f64", "f32"])
def test_b1_stencil_bit_exact_vs_reference(sten, name, tag):
    mesh = torch.from_numpy(sten[f"{name}.{tag}.mesh"])
    x = torch.from_numpy(sten[f"{name}.{tag}.x"])
    phi_ref = sten[f"{name}.{tag}.phi"]
    phi = O.b1_features_dense(mesh, x).numpy()
    assert phi.dtype == phi_ref.dtype
    # numerical equality (== treats -0.0 and 0.0 alike, as the reference emits both)
    assert np.array_equal(phi, phi_ref)
    # partition of unity inside the mesh, zero outside
    inside = (x >= mesh[0]) & (x <= mesh[-1])
    s = phi.sum(0)
    tol = 1e-6 if tag == "f32" else 1e-15
    assert np.allclose(s[inside.numpy()], 1.0, atol=tol)
    assert np.all(s[~inside.numpy()] == 0.0)


PSETS_2D = {
    "raw0": dict(l=[0.0, 0.0], s=[0.0, 0.0], n=0.0),
    "raw1": dict(l=[-0.7, 0.4], s=[0.3, -0.2], n=-2.0),
}

CASES_2D = {
    # tag -> (family, mesh builder)
    "G1_griddedgp2d": (O.B0_GRIDDED, lambda: [O.make_mesh(0, 1, 11)] * 2),
    "G2q_asvgp2d_pad0": (O.B1_ASVGP, lambda: [O.make_padded_mesh(0, 1, 10, 0)] * 2),
    "G2q_asvgp2d_pad1": (O.B1_ASVGP, lambda: [O.make_padded_mesh(0, 1, 10, 1)] * 2),
    "K_b1asvgp2d": (O.B1_ASVGP, lambda: [O.make_mesh(0, 1, 9)] * 2),
    "K_b0gridded2d": (O.B0_GRIDDED, lambda: [O.make_mesh(0, 1, 9)] * 2),
    # kronecker_structure.Matern12SVGP (:287-338): the "meshes" are the columns of its inducing-point parameter Z
    "K_svgp2d": (O.SVGP_GRID, None),
    # kronecker_structure.Matern12VFFGP (:347-514), 4 frequencies: a mesh of 9 knots spanning each domain
    "K_vff2d": (O.VFF_GRID, lambda: [torch.linspace(-0.125, 1.125, 9), torch.linspace(0.25, 0.75, 9)]),
}


def _params(ps, D):
    rl = torch.tensor(ps["l"][:D], dtype=torch.float64, requires_grad=True)
    rs = torch.tensor(ps["s"][:D], dtype=torch.float64, requires_grad=True)
    rn = torch.tensor(ps["n"], dtype=torch.float64, requires_grad=True)
    return rl, rs, rn


@pytest.mark.parametrize("tag", list(CASES_2D))
@pytest.mark.parametrize("pset", list(PSETS_2D))
def test_literal_elbo_matches_reference_2d(ref, tag, pset):
    family, mk = CASES_2D[tag]
    meshes = mk() if mk is not None else [torch.from_numpy(ref["svgp.Z"][:, d].copy()) for d in range(2)]
    X = torch.from_numpy(ref["nb5.X"])
    y = torch.from_numpy(ref["nb5.y"])
    rl, rs, rn = _params(PSETS_2D[pset], 2)
    l, s2, noise = O.constrain(rl, rs, rn)
    elbo = O.elbo_collapsed_literal(family, meshes, X, y, l, s2, noise, ref_quirks=True)
    key = f"{tag}.{pset}"
    assert abs(elbo.item() - float(ref[key + ".elbo"])) <= 1e-9 * abs(float(ref[key + ".elbo"]))
    gl, gs, gn = torch.autograd.grad(elbo, [rl, rs, rn])
    for d in range(2):
        g_ref_l = ref[f"{key}.grad.kernel_{d+1}.base_kernel.raw_lengthscale"][0]
        g_ref_s = ref[f"{key}.grad.kernel_{d+1}.raw_outputscale"][0]
        assert gl[d].item() == pytest.approx(g_ref_l, rel=1e-6, abs=1e-8)
        assert gs[d].item() == pytest.approx(g_ref_s, rel=1e-6, abs=1e-8)
    assert gn.item() == pytest.approx(ref[key + ".grad.likelihood.noise_covar.raw_noise"][0], rel=1e-6)
    # Kuu itself
    Kuu, _, _, _ = O.dense_Kuu_Kuf(family, meshes, X, l.detach(), s2.detach(), True)
    assert np.allclose(Kuu.numpy(), ref[key + ".Kuu"], rtol=1e-13, atol=0)


@pytest.mark.parametrize("tag", ["G1_griddedgp2d", "G2q_asvgp2d_pad1"])
def test_optimal_q_matches_reference(ref, tag):
    family, mk = CASES_2D[tag]
    meshes = mk()
    X = torch.from_numpy(ref["nb5.X"])
    y = torch.from_numpy(ref["nb5.y"])
    rl, rs, rn = _params(PSETS_2D["raw1"], 2)
    l, s2, noise = O.constrain(rl, rs, rn)
    with torch.no_grad():
        mean, cov = O.optimal_q(family, meshes, X, y, l, s2, noise, True)
    key = f"{tag}.raw1"
    q_cov = ref[key + ".q_cov"]
    assert np.allclose(mean.numpy(), ref[key + ".q_mean"], rtol=1e-7, atol=1e-9)
    covn = cov.numpy()
    if tag.startswith("G2q"):
        covn = (covn + covn.T) / 2          # the ASVGP class symmetrises (:915)
    assert np.allclose(covn, q_cov, rtol=1e-6, atol=1e-10)


@pytest.mark.parametrize("pset,ps", [("raw0", dict(l=[0.0], s=[0.0], n=0.0)), ("raw1", dict(l=[-0.3], s=[0.5], n=-3.0))])
def test_literal_elbo_matches_reference_1d(ref, pset, ps):
    x = torch.from_numpy(ref["g3.x"])
    y = torch.from_numpy(ref["g3.y"])
    meshes = [O.make_mesh(0., 2., 33)]
    rl, rs, rn = _params(ps, 1)
    l, s2, noise = O.constrain(rl, rs, rn)
    elbo = O.elbo_collapsed_literal(O.B0_GRIDDED, meshes, x, y, l, s2, noise)
    key = f"G3_griddedgp1d.{pset}"
    assert abs(elbo.item() - float(ref[key + ".elbo"])) <= 1e-9 * abs(float(ref[key + ".elbo"]))
    gl, gs, gn = torch.autograd.grad(elbo, [rl, rs, rn])
    assert gl[0].item() == pytest.approx(ref[key + ".grad.kernel.base_kernel.raw_lengthscale"][0], rel=1e-6)
    assert gs[0].item() == pytest.approx(ref[key + ".grad.kernel.raw_outputscale"][0], rel=1e-6)
    assert gn.item() == pytest.approx(ref[key + ".grad.likelihood.noise_covar.raw_noise"][0], rel=1e-6)

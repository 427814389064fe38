"""The Python host mirror (plan.GridPlan methods, the autograd bridge) driven against the emulated library on the CPU.

The product refuses CPU tensors (`GridPlan needs a CUDA device`), so this test builds a TEST-ONLY subclass whose
constructor repeats the device-independent part of GridPlan.__init__ with the emulated library (tests/emul_lib.py) and
a CPU `device`; every other method -- pack, bin, obs_fwd_bwd dispatch, step, predict, predict_metrics, the
torch.autograd.Function of gridded_elbo -- is the product's own code, executed unchanged.  It exists to catch
Python-level mistakes in host code that has not run on a GPU yet (binned layout, fused metrics); the numerical parity of
the kernels themselves is tests/test_full_emul.py."""
import ctypes as C
import importlib

import numpy as np
import pytest
import torch

import emul_lib
from oracle import vggp_oracle as O
from test_gpu_elbo import make_problem, oracle_value_and_grads, relerr

PKG = "variational-gridded-gaussian-processes_b200"


@pytest.fixture(scope="module")
def host(monkeypatch_module):
    got = emul_lib.load()
    if got is None:
        pytest.skip("g++ not available")
    lib, L = got
    plan_mod = importlib.import_module(PKG + ".plan")
    monkeypatch_module.setattr(plan_mod, "_stream_ptr", lambda device: None)

    def check(status):      # status checking of the product binding, against the emulated library
        if status != 0:
            raise RuntimeError(f"libvggp status {status}: {lib.vggp_last_error().decode()}")

    class EmuGridPlan(plan_mod.GridPlan):
        def __init__(self, family, meshes, obs_dtype):          # device-independent part of GridPlan.__init__
            self.lib = lib
            self.family = int(family)
            self.device = torch.device("cpu")
            self.obs_dtype = obs_dtype
            self.D = len(meshes)
            self.meshes = [m.detach().to("cpu", torch.float32).contiguous() for m in meshes]
            n_knots = (C.c_int * self.D)(*[int(m.numel()) for m in self.meshes])
            ptrs = (C.POINTER(C.c_float) * self.D)(*[C.cast(m.data_ptr(), C.POINTER(C.c_float)) for m in self.meshes])
            handle = C.c_void_p()
            assert lib.vggp_plan_create(C.byref(handle), self.family, self.D, n_knots, ptrs,
                                        plan_mod._OBS_CODE[obs_dtype], 0) == 0
            self.handle = handle
            dims, M, Dd = (C.c_int * 3)(), C.c_int64(), C.c_int()
            assert lib.vggp_plan_dims(handle, C.byref(Dd), dims, C.byref(M)) == 0
            self.m_per_dim = [int(dims[d]) for d in range(self.D)]
            self.M = int(M.value)
            self.L_sizes = [n * n for n in self.m_per_dim]
            self.L_total = sum(self.L_sizes)
            ne, so, ns, tot = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
            assert lib.vggp_gbuf_layout(handle, C.byref(ne), C.byref(so), C.byref(ns), C.byref(tot)) == 0
            self.gbuf_obs_elems, self.gbuf_scalar_offset = int(ne.value), int(so.value)
            self.gbuf_scalars, self.gbuf_bytes = int(ns.value), int(tot.value)
            self.gbuf = torch.zeros(self.gbuf_bytes, dtype=torch.uint8)
            self.last_out = None
            self._armed = False

        def workspace(self, which, dim=0):                      # host pointer -> tensor (the product wraps a device pointer)
            ptr, n = C.c_void_p(), C.c_int64()
            check(lib.vggp_workspace_ptr(self.handle, which, dim, C.byref(ptr), C.byref(n)))
            arr = np.ctypeslib.as_array(C.cast(ptr.value, C.POINTER(C.c_double)), shape=(int(n.value),)).copy()
            out = torch.from_numpy(arr)
            if which in (plan_mod._lib.WS_ALPHA, plan_mod._lib.WS_SCAL, plan_mod._lib.WS_QBAND):
                return out
            nd = self.m_per_dim[dim]
            return out.view(nd, nd)

        # no CUDA events / pinned memory on the CPU: the emulated library is synchronous, read the flag directly
        def arm_info_check(self):
            self._armed = True

        def poll_info(self, wait):
            if not self._armed:
                return None
            self._armed = False
            return self.read_info()

    monkeypatch_module.setattr(plan_mod._lib, "check", check)
    return plan_mod, EmuGridPlan, L


@pytest.fixture(scope="module")
def monkeypatch_module():
    mp = pytest.MonkeyPatch()
    yield mp
    mp.undo()


@pytest.mark.parametrize("layout", ["raw", "packed", "binned"])
@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-9), (torch.float32, 1e-3)])
def test_gridded_elbo_autograd_bridge_over_the_emulator(host, layout, dtype, tol):
    plan_mod, EmuGridPlan, L = host
    knots, N = (9, 7), 600
    meshes, X, y, l, s2, noise, m, Ls = make_problem(knots, N, seed=44)
    Xq, yq = X.to(dtype), y.to(dtype)
    elbo_ref, g_ref = oracle_value_and_grads(O.B1_ASVGP, meshes, Xq.double(), yq.double(), l, s2, noise, m, Ls, scale=1.0)
    plan = EmuGridPlan(L.B1_ASVGP, meshes, dtype)
    xs = [Xq[:, d].contiguous() for d in range(2)]
    if layout == "packed":
        obs, yy = plan.pack(xs, yq, sort_by_cell=True), None
        assert obs.n == N and obs.run_len % 4 == 0
    elif layout == "binned":
        obs, yy = plan.bin(xs, yq, run_cap=32), None
        assert obs.n == N and obs.n_tasks == (obs.n_runs + 31) // 32 and obs.streamed_bytes >= 12 * obs.n_inside * (dtype == torch.float32)
    else:
        obs, yy = xs, yq
    params = [t.clone().requires_grad_(True) for t in (l, s2, noise.reshape(1), m)] + [Lx.clone().requires_grad_(True) for Lx in Ls]
    elbo = plan_mod.gridded_elbo(plan, obs, yy, params[0], params[1], params[2], params[3], params[4:], 1.0, None)
    (-elbo).backward()
    assert abs(elbo.item() - elbo_ref.item()) <= tol * abs(elbo_ref.item())
    assert relerr(-params[0].grad, g_ref[0]) < 10 * tol and relerr(-params[1].grad, g_ref[1]) < 10 * tol
    assert relerr(-params[2].grad, g_ref[2]) < 10 * tol and relerr(-params[3].grad, g_ref[3]) < 10 * tol
    for d in range(2):
        assert relerr(torch.tril(-params[4 + d].grad), torch.tril(g_ref[4 + d])) < 10 * tol
    assert plan.read_info() == 0


def test_predict_and_fused_metrics_through_the_host_mirror(host):
    plan_mod, EmuGridPlan, L = host
    meshes, X, y, l, s2, noise, m, Ls = make_problem((9, 7), 500, seed=8)
    plan = EmuGridPlan(L.B1_ASVGP, meshes, torch.float64)
    plan.grid_forward(torch.cat([l, s2, noise.reshape(1)]), m.clone(), torch.cat([Lx.reshape(-1) for Lx in Ls]))
    xs = [X[:, d].contiguous() for d in range(2)]
    mean, var = plan.predict(xs)
    got = plan.predict_metrics(xs, y)
    e = y - mean
    assert abs(got["mse"].item() - float((e ** 2).mean())) < 1e-12
    assert abs(got["mae"].item() - float(e.abs().mean())) < 1e-12
    assert abs(got["rmse"].item() - float((e ** 2).mean().sqrt())) < 1e-12
    r2 = 1 - float((e ** 2).sum() / ((y - y.mean()) ** 2).sum())
    assert abs(got["r2"].item() - r2) < 1e-10
    c, wl, wh = plan.b1_stencil(0, xs[0])
    co, wlo, who = O.b1_stencil(meshes[0], xs[0])
    assert torch.equal(c.to(torch.int64), co) and torch.equal(wl, wlo) and torch.equal(wh, who)


def test_b0_family_binned_through_the_host_mirror(host):
    plan_mod, EmuGridPlan, L = host
    meshes, X, y, l, s2, noise, m, Ls = make_problem((8, 7), 400, seed=3, family=O.B0_GRIDDED, x_lo=-0.2, x_hi=1.2)
    plan = EmuGridPlan(L.B0_GRIDDED, meshes, torch.float64)
    theta = torch.cat([l, s2, noise.reshape(1)])
    Lcat = torch.cat([Lx.reshape(-1) for Lx in Ls])
    xs = [X[:, d].contiguous() for d in range(2)]
    scan = [t.clone() for t in plan.step(theta, m.clone(), Lcat, plan.bin(xs, y, run_cap=16), None, 1.2)]
    dense = plan.step(theta, m.clone(), Lcat, xs, y, 1.2)
    for a, b in zip(scan, dense):
        assert torch.allclose(a, b, rtol=1e-8, atol=1e-10)
    mean, var = plan.predict(xs)                   # B0 point prediction in scan form
    assert torch.isfinite(mean).all() and (var > 0).all()


@pytest.mark.parametrize("family,layout", [("asvgp", "packed"), ("asvgp", "binned"), ("gridded", "dense"), ("gridded", "binned")])
def test_model_training_loop_over_the_emulator(host, monkeypatch, family, layout):
    """The drop-in model classes (reference constructor signatures, parameters(), -_elbo().backward(), Adam) on the CPU:
    the two CUDA gates of the product (GridPlan's constructor, the model's device check) are replaced for this test only,
    everything else -- parameter containers, softplus constraints, the layout opt-in, the autograd bridge -- is the
    product's code.  ELBO and raw-parameter gradients against the oracle; a few Adam steps increase the bound."""
    plan_mod, EmuGridPlan, L = host
    gmod = importlib.import_module(PKG + ".models._gridded")
    gks = importlib.import_module(PKG + ".models.sparse.gridded_kronecker_structure")
    monkeypatch.setattr(gmod, "GridPlan", lambda fam, meshes, dtype, device: EmuGridPlan(fam, meshes, dtype))
    monkeypatch.setattr(gmod.GriddedVariationalGP, "_device_dtype", lambda self: (torch.device("cpu"), self.variational_mean.dtype))
    if layout == "binned":
        monkeypatch.setenv("VGGP_OBS_LAYOUT", "binned")
    else:
        monkeypatch.setenv("VGGP_OBS_LAYOUT", "packed")       # the round-1 layouts (B1: packed runs, B0: dense-feature kernel)
    g = torch.Generator().manual_seed(0)
    N = 400
    X = torch.rand(N, 2, generator=g, dtype=torch.float64)
    y = torch.sin(5 * X[:, 0]) + torch.cos(7 * X[:, 1]) + 0.05 * torch.randn(N, generator=g, dtype=torch.float64)
    if family == "asvgp":
        model = gks.GriddedMatern12ASVGP(X, y, 6, 1, (0, 1), (0, 1)).to(torch.float64)
        ofam, meshes = O.B1_ASVGP, [O.make_padded_mesh(0, 1, 6, 1)] * 2
    else:
        model = gks.Matern12GriddedGP(X, y, 7, (0, 1), (0, 1)).to(torch.float64)
        ofam, meshes = O.B0_GRIDDED, [O.make_mesh(0, 1, 7)] * 2
    with torch.no_grad():
        model.variational_mean.normal_(0, 0.1)
        model.kernel_1.base_kernel.lengthscale = 0.4
        model.likelihood.noise = 0.05
    elbo = model._elbo()
    assert elbo.dim() == 0 and elbo.requires_grad
    (-elbo).backward()
    packed = model._packed
    assert (packed is None) == (layout == "dense")
    if layout == "binned":
        assert type(packed).__name__ == "BinnedObs"
    raw_l = torch.stack([model.kernel_1.base_kernel.raw_lengthscale.detach().reshape(()),
                         model.kernel_2.base_kernel.raw_lengthscale.detach().reshape(())]).requires_grad_(True)
    raw_s = torch.stack([model.kernel_1.raw_outputscale.detach(), model.kernel_2.raw_outputscale.detach()]).requires_grad_(True)
    raw_n = model.likelihood.noise_covar.raw_noise.detach().reshape(()).requires_grad_(True)
    l, s2, noise = O.constrain(raw_l, raw_s, raw_n)
    ref = O.elbo_structured(ofam, meshes, X, y, l, s2, noise, model.variational_mean.detach(),
                            [model.variational_chol_1.detach(), model.variational_chol_2.detach()], ref_quirks=False)
    gl, gs, gn = torch.autograd.grad(-ref, [raw_l, raw_s, raw_n])
    assert abs(elbo.item() - ref.item()) < 1e-8 * abs(ref.item())
    assert abs(model.kernel_1.base_kernel.raw_lengthscale.grad.item() - gl[0].item()) < 1e-6 * abs(gl[0].item())
    assert abs(model.kernel_2.raw_outputscale.grad.item() - gs[1].item()) < 1e-6 * abs(gs[1].item())
    assert abs(model.likelihood.noise_covar.raw_noise.grad.item() - gn.item()) < 1e-6 * abs(gn.item())
    opt = torch.optim.Adam(model.parameters(), lr=0.05)
    losses = []
    for _ in range(6):
        opt.zero_grad()
        loss = -model._elbo()
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert losses[-1] < losses[0]
    post = model.posterior(X[:50])
    assert torch.isfinite(post.mean).all() and (post.variance > 0).all()


# ---------------------------------------------------------------------------------------------------------
# the reference's closed-form quantities (q_u / q_v of the collapsed bound) against numbers produced by the reference's own
# code (tests/golden/reference_models.npz, oracle/make_golden.py)
# ---------------------------------------------------------------------------------------------------------
def _patched_models(host, monkeypatch):
    plan_mod, EmuGridPlan, L = host
    gmod = importlib.import_module(PKG + ".models._gridded")
    monkeypatch.setattr(gmod, "GridPlan", lambda fam, meshes, dtype, device: EmuGridPlan(fam, meshes, dtype))
    monkeypatch.setattr(gmod.GriddedVariationalGP, "_device_dtype", lambda self: (torch.device("cpu"), self.variational_mean.dtype))
    return gmod


def _set_raw(model, spec):
    sd = dict(model.named_parameters())
    for k, v in spec.items():
        sd[k].data.fill_(v)


@pytest.mark.parametrize("pset", ["raw0", "raw1"])
def test_closed_form_q_1d_matches_reference_run(host, monkeypatch, golden_dir, pset):
    """gridded_univariate_structure.Matern12GriddedGP: q_u_optimal() / q_v(optimal=True) reproduce the reference's q_v(), and
    after set_optimal_q() the uncollapsed `_elbo()` IS the reference's collapsed `_elbo()` (case G3 of the golden file)."""
    import os
    _patched_models(host, monkeypatch)
    gus = importlib.import_module(PKG + ".models.sparse.gridded_univariate_structure")
    ref = np.load(os.path.join(golden_dir, "reference_models.npz"))
    key = f"G3_griddedgp1d.{pset}"
    x, y = torch.from_numpy(ref["g3.x"]), torch.from_numpy(ref["g3.y"])
    model = gus.Matern12GriddedGP(x, y, 32, (0., 2.)).to(torch.float64)
    _set_raw(model, {"raw0": {}, "raw1": {"kernel.raw_outputscale": 0.5, "kernel.base_kernel.raw_lengthscale": -0.3,
                                          "likelihood.noise_covar.raw_noise": -3.0}}[pset])
    assert relerr(model._Kuu(), torch.from_numpy(ref[key + ".Kuu"])) < 1e-12
    q = model.q_v(optimal=True)
    assert relerr(q.mean, torch.from_numpy(ref[key + ".q_mean"])) < 1e-8
    assert relerr(q.covariance_matrix, torch.from_numpy(ref[key + ".q_cov"])) < 1e-8
    Sig = model._sigma()
    assert Sig.shape == (32, 32) and torch.allclose(Sig, Sig.T)
    model.set_optimal_q()
    elbo = model._elbo()
    assert abs(elbo.item() - float(ref[key + ".elbo"])) < 1e-6 * abs(float(ref[key + ".elbo"]))
    model.check_factorisation()
    terms = model.elbo_terms()
    assert abs(terms[0].item() - elbo.item()) < 1e-12 * abs(elbo.item()) and abs((terms[1] - terms[2]).item() - elbo.item()) < 1e-9 * abs(elbo.item())
    assert terms[3].item() == 600


@pytest.mark.parametrize("case,pset", [("G1_griddedgp2d", "raw0"), ("G1_griddedgp2d", "raw1"), ("K_b0gridded2d", "raw1")])
def test_closed_form_q_2d_matches_reference_run(host, monkeypatch, golden_dir, case, pset):
    """2-D B0 models (gridded_kronecker_structure.Matern12GriddedGP, kronecker_structure.Matern12B0SplineGriddedGP): the dense
    closed-form q(u) against the reference's own q_u() / q_v() output; set_optimal_q() loads m* and the nearest Kronecker
    factorisation of S*, whose bound cannot exceed the collapsed one and must be close to it."""
    import os
    _patched_models(host, monkeypatch)
    ref = np.load(os.path.join(golden_dir, "reference_models.npz"))
    X, y = torch.from_numpy(ref["nb5.X"]).to(torch.float64), torch.from_numpy(ref["nb5.y"]).to(torch.float64)
    if case.startswith("G1"):
        mod = importlib.import_module(PKG + ".models.sparse.gridded_kronecker_structure")
        model = mod.Matern12GriddedGP(X, y, 11, (0, 1), (0, 1)).to(torch.float64)
    else:
        mod = importlib.import_module(PKG + ".models.sparse.kronecker_structure")
        model = mod.Matern12B0SplineGriddedGP(X, y, 9, (0, 1), (0, 1)).to(torch.float64)
    _set_raw(model, {"raw0": {}, "raw1": {"kernel_1.raw_outputscale": 0.3, "kernel_1.base_kernel.raw_lengthscale": -0.7,
                                          "kernel_2.raw_outputscale": -0.2, "kernel_2.base_kernel.raw_lengthscale": 0.4,
                                          "likelihood.noise_covar.raw_noise": -2.0}}[pset])
    key = f"{case}.{pset}"
    assert relerr(model._Kuu(), torch.from_numpy(ref[key + ".Kuu"])) < 1e-12
    q = model.q_v(optimal=True)
    assert relerr(q.mean, torch.from_numpy(ref[key + ".q_mean"])) < 1e-7
    assert relerr(q.covariance_matrix, torch.from_numpy(ref[key + ".q_cov"])) < 1e-7
    model.set_optimal_q()
    elbo, collapsed = model._elbo().item(), float(ref[key + ".elbo"])
    assert elbo <= collapsed + 1e-8 * abs(collapsed)          # the collapsed bound is the maximum over q(u)
    assert collapsed - elbo < 0.2 * abs(collapsed)            # Kronecker S is a restriction, not a different model
    # full-covariance posterior of the dense formulas against the marginals of the structured path
    xs = X[:40]
    dense = model.posterior_dense(xs)
    marg = model.posterior(xs)
    assert relerr(dense.mean, marg.mean) < 1e-8 and relerr(dense.variance, marg.variance) < 1e-7


@pytest.mark.parametrize("pset", ["raw0", "raw1"])
def test_svgp_model_class_matches_reference_run(host, monkeypatch, golden_dir, pset):
    """kronecker_structure.Matern12SVGP (product-grid SVGP, :287-338) as a drop-in class: its Kuu equals the reference's own
    `_Kuu()` output, the uncollapsed bound at the closed-form q(u) stays below the reference's collapsed ELBO and close to it,
    ELBO and raw hyper-parameter gradients equal the oracle's, Adam improves the bound, posterior() equals the dense formulas."""
    import os
    _patched_models(host, monkeypatch)
    ref = np.load(os.path.join(golden_dir, "reference_models.npz"))
    X, y = torch.from_numpy(ref["nb5.X"]).to(torch.float64), torch.from_numpy(ref["nb5.y"]).to(torch.float64)
    Z = torch.from_numpy(ref["svgp.Z"])
    mod = importlib.import_module(PKG + ".models.sparse.kronecker_structure")
    model = mod.Matern12SVGP(X, y, Z).to(torch.float64)
    assert model.m_per_dim == [7, 7] and model._packed is None
    _set_raw(model, {"raw0": {}, "raw1": {"kernel_1.raw_outputscale": 0.3, "kernel_1.base_kernel.raw_lengthscale": -0.7,
                                          "kernel_2.raw_outputscale": -0.2, "kernel_2.base_kernel.raw_lengthscale": 0.4,
                                          "likelihood.noise_covar.raw_noise": -2.0}}[pset])
    key = f"K_svgp2d.{pset}"
    assert relerr(model._Kuu(), torch.from_numpy(ref[key + ".Kuu"])) < 1e-12
    model.set_optimal_q()
    elbo, collapsed = model._elbo().item(), float(ref[key + ".elbo"])
    assert elbo <= collapsed + 1e-8 * abs(collapsed)
    assert collapsed - elbo < 0.2 * abs(collapsed)
    # a generic point: ELBO and raw-parameter gradients against the oracle
    with torch.no_grad():
        model.variational_mean.add_(0.05 * torch.randn(model.M, generator=torch.Generator().manual_seed(1), dtype=torch.float64))
    model.zero_grad()
    e = model._elbo()
    (-e).backward()
    raw_l = torch.stack([model.kernel_1.base_kernel.raw_lengthscale.detach().reshape(()),
                         model.kernel_2.base_kernel.raw_lengthscale.detach().reshape(())]).requires_grad_(True)
    raw_s = torch.stack([model.kernel_1.raw_outputscale.detach(), model.kernel_2.raw_outputscale.detach()]).requires_grad_(True)
    raw_n = model.likelihood.noise_covar.raw_noise.detach().reshape(()).requires_grad_(True)
    l, s2, noise = O.constrain(raw_l, raw_s, raw_n)
    meshes = [model.Z[:, 0].clone(), model.Z[:, 1].clone()]
    oref = O.elbo_structured(O.SVGP_GRID, meshes, X, y, l, s2, noise, model.variational_mean.detach(),
                             [model.variational_chol_1.detach(), model.variational_chol_2.detach()], ref_quirks=False)
    gl, gs, gn = torch.autograd.grad(-oref, [raw_l, raw_s, raw_n])
    assert abs(e.item() - oref.item()) < 1e-8 * abs(oref.item())
    assert abs(model.kernel_1.base_kernel.raw_lengthscale.grad.item() - gl[0].item()) < 1e-6 * abs(gl[0].item())
    assert abs(model.kernel_2.raw_outputscale.grad.item() - gs[1].item()) < 1e-6 * abs(gs[1].item())
    assert abs(model.likelihood.noise_covar.raw_noise.grad.item() - gn.item()) < 1e-6 * abs(gn.item())
    xs = X[:40]
    dense = model.posterior_dense(xs)
    marg = model.posterior(xs)
    assert relerr(dense.mean, marg.mean) < 1e-8 and relerr(dense.variance, marg.variance) < 1e-7
    opt = torch.optim.Adam(model.parameters(), lr=0.02)
    losses = []
    for _ in range(5):
        opt.zero_grad()
        loss = -model._elbo()
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert losses[-1] < losses[0]


@pytest.mark.parametrize("pset", ["raw0", "raw1"])
def test_vff_model_class_matches_reference_run(host, monkeypatch, golden_dir, pset):
    """kronecker_structure.Matern12VFFGP (variational Fourier features, :347-514) as a drop-in class: its Kuu equals the
    reference's own (float32-computed) `_Kuu()` output to float32 accuracy, the uncollapsed bound at the closed-form q(u) stays
    below the reference's collapsed ELBO and close to it, ELBO and raw hyper-parameter gradients equal the oracle's (float64
    semantics), posterior() equals the dense formulas, Adam improves the bound.  The second domain is smaller than the data."""
    import os
    _patched_models(host, monkeypatch)
    ref = np.load(os.path.join(golden_dir, "reference_models.npz"))
    X, y = torch.from_numpy(ref["nb5.X"]).to(torch.float64), torch.from_numpy(ref["nb5.y"]).to(torch.float64)
    mod = importlib.import_module(PKG + ".models.sparse.kronecker_structure")
    model = mod.Matern12VFFGP(X, y, 4, (-0.125, 1.125), (0.25, 0.75)).to(torch.float64)
    assert model.m_per_dim == [9, 9] and model._packed is None
    _set_raw(model, {"raw0": {}, "raw1": {"kernel_1.raw_outputscale": 0.3, "kernel_1.base_kernel.raw_lengthscale": -0.7,
                                          "kernel_2.raw_outputscale": -0.2, "kernel_2.base_kernel.raw_lengthscale": 0.4,
                                          "likelihood.noise_covar.raw_noise": -2.0}}[pset])
    key = f"K_vff2d.{pset}"
    assert relerr(model._Kuu(), torch.from_numpy(ref[key + ".Kuu"])) < 1e-6
    model.set_optimal_q()
    elbo, collapsed = model._elbo().item(), float(ref[key + ".elbo"])
    assert elbo <= collapsed + 1e-6 * abs(collapsed)
    assert collapsed - elbo < 0.2 * abs(collapsed)
    with torch.no_grad():
        model.variational_mean.add_(0.05 * torch.randn(model.M, generator=torch.Generator().manual_seed(1), dtype=torch.float64))
    model.zero_grad()
    e = model._elbo()
    (-e).backward()
    raw_l = torch.stack([model.kernel_1.base_kernel.raw_lengthscale.detach().reshape(()),
                         model.kernel_2.base_kernel.raw_lengthscale.detach().reshape(())]).requires_grad_(True)
    raw_s = torch.stack([model.kernel_1.raw_outputscale.detach(), model.kernel_2.raw_outputscale.detach()]).requires_grad_(True)
    raw_n = model.likelihood.noise_covar.raw_noise.detach().reshape(()).requires_grad_(True)
    l, s2, noise = O.constrain(raw_l, raw_s, raw_n)
    meshes = [torch.linspace(-0.125, 1.125, 9), torch.linspace(0.25, 0.75, 9)]
    oref = O.elbo_structured(O.VFF_GRID, meshes, X, y, l, s2, noise, model.variational_mean.detach(),
                             [model.variational_chol_1.detach(), model.variational_chol_2.detach()], ref_quirks=False)
    gl, gs, gn = torch.autograd.grad(-oref, [raw_l, raw_s, raw_n])
    assert abs(e.item() - oref.item()) < 1e-8 * abs(oref.item())
    assert abs(model.kernel_2.base_kernel.raw_lengthscale.grad.item() - gl[1].item()) < 1e-6 * abs(gl[1].item())
    assert abs(model.kernel_1.raw_outputscale.grad.item() - gs[0].item()) < 1e-6 * abs(gs[0].item())
    assert abs(model.likelihood.noise_covar.raw_noise.grad.item() - gn.item()) < 1e-6 * abs(gn.item())
    xs = X[:40]
    dense = model.posterior_dense(xs)
    marg = model.posterior(xs)
    assert relerr(dense.mean, marg.mean) < 1e-8 and relerr(dense.variance, marg.variance) < 1e-7
    opt = torch.optim.Adam(model.parameters(), lr=0.02)
    losses = []
    for _ in range(5):
        opt.zero_grad()
        loss = -model._elbo()
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert losses[-1] < losses[0]


def test_gridded_part_dense_matrices_asvgp(host, monkeypatch):
    """GriddedMatern12ASVGP (2-D) and its 1-D twin: _Kvu / _Kvv / p_v_u / q_v against the reference's constructions restated
    with torch (gridded_kronecker_structure.py:831-947, gridded_univariate_structure.py:595-700)."""
    _patched_models(host, monkeypatch)
    gks = importlib.import_module(PKG + ".models.sparse.gridded_kronecker_structure")
    gus = importlib.import_module(PKG + ".models.sparse.gridded_univariate_structure")
    g = torch.Generator().manual_seed(3)
    X = torch.rand(300, 2, generator=g, dtype=torch.float64)
    y = torch.sin(4 * X[:, 0]) * torch.cos(3 * X[:, 1]) + 0.05 * torch.randn(300, generator=g, dtype=torch.float64)
    model = gks.GriddedMatern12ASVGP(X, y, 6, 1, (0, 1), (0, 1)).to(torch.float64)
    # reference construction of Kvu_along_dim: padded [delta, delta] rolled by the cell index
    delta = model.b1_basis_1.delta
    nk = model.b1_basis_1.n_basis_functions
    first = torch.nn.functional.pad(torch.tensor([delta, delta]), (1, nk - 3))
    Kvu_ref = torch.vstack([torch.roll(first, i) for i in range(6)]).to(torch.float64)
    assert torch.equal(model._Kvu_along_dim(0), Kvu_ref)
    assert model._Kvu().shape == (36, nk * nk) and model._Kvv().shape == (36, 36)
    assert relerr(model._Kvv_along_dim(0), O.kuu_b0(model.b0_mesh_1, model.kernel_1.base_kernel.lengthscale.detach().reshape(()),
                                                    model.kernel_1.outputscale.detach().reshape(()))) < 1e-12
    with torch.no_grad():
        model.variational_mean.normal_(0, 0.3)
    qd = model.q_v_dense()
    qm = model.q_v()                       # structured marginals (band formulas) of the same q(u)
    assert relerr(qd.mean, qm.mean) < 1e-9 and relerr(qd.variance, qm.variance) < 1e-8
    pv = model.p_v_u()
    assert relerr(pv.mean, qd.mean) < 1e-12 and (torch.diagonal(pv.covariance_matrix) <= qd.variance + 1e-12).all()
    # optimal q(u): the reference's formulas through Sigma^-1
    qo = model.q_v_dense(optimal=True)
    Kuu, Kvu, Kvv = model._Kuu(), model._Kvu(), model._Kvv()
    Kuf = model._Kuf(X).to(torch.float64)
    noise = model.likelihood.noise.detach().reshape(())
    Sig = Kuu + Kuf @ Kuf.T / noise
    mean_ref = Kvu @ torch.linalg.solve(Sig, Kuf @ y) / noise
    cov_ref = Kvv - Kvu @ torch.linalg.solve(Kuu, Kvu.T) + Kvu @ torch.linalg.solve(Sig, Kvu.T)
    assert relerr(qo.mean, mean_ref) < 1e-7 and relerr(qo.covariance_matrix, cov_ref) < 1e-6
    # 1-D twin
    x1 = torch.rand(200, generator=g, dtype=torch.float64) * 2
    y1 = torch.sin(x1) + 0.05 * torch.randn(200, generator=g, dtype=torch.float64)
    m1 = gus.GriddedMatern12ASVGP(x1, y1, 8, 3, (0., 2.)).to(torch.float64)
    Kvu1 = m1._Kvu()
    assert Kvu1.shape == (8, m1.b1_basis_1.n_basis_functions)
    d1 = m1.b1_basis_1.delta.to(torch.float64)
    assert torch.allclose(Kvu1.sum(1), 4 * d1.expand(8))        # delta/2 + 3 delta + delta/2 = the cell width
    q1 = m1.q_v(optimal=True)
    Kuu1, Kuf1 = m1._Kuu(), m1._Kuf(x1.reshape(-1, 1)).to(torch.float64)
    n1 = m1.likelihood.noise.detach().reshape(())
    Sig1 = Kuu1 + Kuf1 @ Kuf1.T / n1
    assert relerr(q1.mean, Kvu1 @ torch.linalg.solve(Sig1, Kuf1 @ y1) / n1) < 1e-7
    cov1 = m1._Kvv() - Kvu1 @ torch.linalg.solve(Kuu1, Kvu1.T) + Kvu1 @ torch.linalg.solve(Sig1, Kvu1.T)
    assert relerr(q1.covariance_matrix, cov1) < 1e-6


def test_model_level_deterministic_mode_over_the_emulator(host, monkeypatch):
    """GriddedVariationalGP.set_deterministic (vggp_set_deterministic): same ELBO and raw-parameter gradients as the atomics path
    to rounding, identical bits from two evaluations, and a clear error for a family the mode does not cover."""
    gmod = _patched_models(host, monkeypatch)
    gks = importlib.import_module(PKG + ".models.sparse.gridded_kronecker_structure")
    monkeypatch.setenv("VGGP_OBS_LAYOUT", "binned")
    g = torch.Generator().manual_seed(1)
    N = 500
    X = torch.rand(N, 2, generator=g, dtype=torch.float64)
    y = torch.sin(5 * X[:, 0]) + torch.cos(7 * X[:, 1]) + 0.05 * torch.randn(N, generator=g, dtype=torch.float64)
    model = gks.GriddedMatern12ASVGP(X, y, 6, 1, (0, 1), (0, 1)).to(torch.float64)
    with torch.no_grad():
        model.variational_mean.normal_(0, 0.1)

    def value_and_grads():
        for q in model.parameters():
            q.grad = None
        e = model._elbo()
        (-e).backward()
        return [e.detach().clone()] + [q.grad.detach().clone() for q in model.parameters()]

    plain = value_and_grads()
    model.set_deterministic(True)
    a, b = value_and_grads(), value_and_grads()
    for u, v in zip(a, b):
        assert torch.equal(u, v)
    for u, v in zip(a, plain):
        assert torch.allclose(u, v, rtol=1e-9, atol=1e-12)
    model.set_deterministic(False)
    other = gks.Matern12GriddedGP(X, y, 7, (0, 1), (0, 1)).to(torch.float64)
    with pytest.raises(RuntimeError):
        other.set_deterministic(True)

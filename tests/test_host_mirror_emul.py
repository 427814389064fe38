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

        # no CUDA events / pinned memory on the CPU: the emulated library is synchronous, read the flag directly
        def arm_info_check(self):
            self._armed = True

        def poll_info(self, wait):
            if not self._armed:
                return None
            self._armed = False
            return self.read_info()

    # status checking of the product binding, against the emulated library
    def check(status):
        if status != 0:
            raise RuntimeError(f"libvggp status {status}: {lib.vggp_last_error().decode()}")
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

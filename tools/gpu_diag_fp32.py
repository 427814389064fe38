"""Diagnostic (not product): where the float32 error of the config-2 step comes from.  Prints, for N = 2^24 tracks on
512 x 512, the relative error of every gradient block against the chunked oracle for the binned / packed layouts, and the
error of the raw d alpha sums against a float64 torch evaluation of the definition."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
import vggp_b200 as vg
from oracle import vggp_oracle as O
from chunked_oracle import elbo_and_grads_chunked

def rel(a, b):
    a = a.detach().cpu().double().reshape(-1); b = b.detach().cpu().double().reshape(-1)
    return ((a - b).norm() / b.norm()).item()

dev = torch.device("cuda", 0)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
meshes = [torch.linspace(0, 1, k) for k in bench.KNOTS]
xs, y = bench.make_tracks(0, N, N, dev, torch.float32)
theta, m, Ls = bench.make_params(meshes, dev)
X = torch.stack([x.cpu() for x in xs], 1)
eref, gref = elbo_and_grads_chunked(O.B1_ASVGP, meshes, X, y.cpu(), theta[:2].clone(), theta[2:4].clone(), theta[4].clone(), m, Ls, chunk=1 << 21)
Lcat = torch.cat([L.reshape(-1) for L in Ls]).to(dev).contiguous()
# float64 definition of the raw sums g_alpha = sum_n r_n w_n (corner form), on the GPU with torch
K = bench.KNOTS[0]
t = meshes[0].double().to(dev)
def sten(x):
    x = x.double()
    c = (torch.searchsorted(t, x, right=False) - 1).clamp(0, K - 2)
    return c, (x - t[c]) / (t[c + 1] - t[c])
plan64 = vg.GridPlan(vg.B1_ASVGP, meshes, torch.float64, dev)
plan64.grid_forward(theta.to(dev), m.to(dev), Lcat)
alpha = plan64.workspace(vg._lib.WS_ALPHA).reshape(K, K)
c1, a1 = sten(xs[0]); c2, a2 = sten(xs[1])
mu = alpha[c1, c2] * (1 - a1) * (1 - a2) + alpha[c1, c2 + 1] * (1 - a1) * a2 + alpha[c1 + 1, c2] * a1 * (1 - a2) + alpha[c1 + 1, c2 + 1] * a1 * a2
r = y.double() - mu
g64 = torch.zeros(K * K, dtype=torch.float64, device=dev)
for w, di, dj in (((1 - a1) * (1 - a2), 0, 0), ((1 - a1) * a2, 0, 1), (a1 * (1 - a2), 1, 0), (a1 * a2, 1, 1)):
    g64.index_add_(0, (c1 + di) * K + (c2 + dj), r * w)
del mu, r, c1, c2, a1, a2
for dtype in (torch.float32, torch.float64):
    plan = vg.GridPlan(vg.B1_ASVGP, meshes, dtype, dev)
    xd = [x.to(dtype) for x in xs]; yd = y.to(dtype)
    for name in ("binned", "packed"):
        obs = plan.bin(xd, yd, run_cap=256) if name == "binned" else plan.pack(xd, yd, sort_by_cell=True)
        out, dth, dm, dL = plan.step(theta.to(dev), m.to(dev), Lcat, obs, None)
        gobs, gsc = plan.gbuf_views()
        ga = gobs[:K * K].double()
        print(f"{dtype} {name}: elbo {abs(out[0].item() - eref.item()) / abs(eref.item()):.2e} dl {rel(dth[:2], gref[0]):.2e} ds2 {rel(dth[2:4], gref[1]):.2e} "
              f"dnoise {rel(dth[4], gref[2]):.2e} dm {rel(dm, gref[3]):.2e} dL0 {rel(torch.tril(dL[:K*K].reshape(K,K)), torch.tril(gref[4])):.2e} "
              f"dL1 {rel(torch.tril(dL[K*K:].reshape(K,K)), torch.tril(gref[5])):.2e} | g_alpha vs f64 definition {rel(ga, g64):.2e} "
              f"|dm| {dm.norm().item():.3e} |alpha| {alpha.norm().item():.3e}")
        del obs

"""Diagnostic (not product): where the cycles of a fibre pass go.  Runs bench-config steps with the library's phase stamps
switched on (vggp_debug_fp_stamps) for one pass at a time and prints, per task kind, the mean cycles between the phase
boundaries of a CTA: start -> generators loaded -> tile loaded -> recurrences done -> epilogue done."""
import ctypes as C, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import vggp_b200 as vg

dev = torch.device("cuda", 0)
lib = vg._lib.load()
lib.vggp_debug_fp_stamps.argtypes = [C.c_void_p]
N = 1 << 24
meshes = [torch.linspace(0, 1, k) for k in bench.KNOTS]
plan = vg.GridPlan(vg.B1_ASVGP, meshes, torch.float32, dev)
xs, y = bench.make_tracks(0, N, N, dev, torch.float32)
theta, m, Ls = bench.make_params(meshes, dev)
theta, m = theta.to(dev), m.to(dev)
L = torch.cat([l.reshape(-1) for l in Ls]).to(dev).contiguous()
obs = plan.bin(xs, y)
for _ in range(5):
    plan.step(theta, m, L, obs, None)
torch.cuda.synchronize()
buf = torch.zeros(4096 * 8, dtype=torch.int64, device=dev)
KIND = ["R", "PROD", "ALPHA", "GA", "GAONLY", "DM", "DL", "QROW"]
def show(name, fn):
    buf.zero_()
    torch.cuda.synchronize()
    lib.vggp_debug_fp_stamps(buf.data_ptr())
    fn()
    torch.cuda.synchronize()
    lib.vggp_debug_fp_stamps(None)
    b = buf.view(-1, 8).cpu()
    b = b[b[:, 4] != 0]          # CTAs of the LAST pass of the call overwrite earlier ones: rows are per blockIdx
    if b.shape[0] == 0:
        print(f"== {name}: no stamps (the fast kernel is not instrumented; vggp_debug_fp_fast(0) selects the generic one)")
        return
    t0 = b[:, 6].min()
    print(f"== {name}: {b.shape[0]} stamped CTAs, span of CTA starts {(b[:, 6].max() - t0).item() / 1e3:.1f} us")
    for k in sorted(set(b[:, 7].tolist())):
        r = b[b[:, 7] == k].double()
        d = [(r[:, i + 1] - r[:, i]).mean().item() for i in range(4)]
        print(f"   {KIND[int(k)]:7s} n={r.shape[0]:4d}  gens {d[0]:8.0f}  load {d[1]:8.0f}  recur {d[2]:8.0f}  epilogue {d[3]:8.0f}  total {sum(d):8.0f} cycles")
show("forward (last pass: F2)", lambda: plan.grid_forward(theta, m, L))
plan.obs_fwd_bwd(obs)
show("backward (last pass: B2)", lambda: plan.grid_backward(theta, m, L, 1.0))
# theta kernel: it runs last in grid_backward and overwrites rows 0..D-1 of the buffer with its own stamps
buf.zero_(); torch.cuda.synchronize()
lib.vggp_debug_fp_stamps(buf.data_ptr())
plan.grid_backward(theta, m, L, 1.0)
torch.cuda.synchronize()
lib.vggp_debug_fp_stamps(None)
b = buf.view(-1, 8).cpu()[:2].double()
names = ["stage loads", "recurrences (warps 0,1) + dK/dtheta constants", "barrier wait", "band loop + reduce", "scalar tail"]
for d_ in range(2):
    print(f"== k_b1_theta block {d_}: " + ", ".join(f"{names[i]} {(b[d_, i + 1] - b[d_, i]).item():.0f}" for i in range(5)) + " cycles")

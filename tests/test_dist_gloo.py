"""Host-side multi-rank logic on the CPU: world_size = 2, gloo backend.

What is checked: the contiguous observation sharding (every observation on exactly one rank), and that ONE
sum-all-reduce of the per-observation gradient buffer -- laid out exactly as the CUDA kernel writes it:
[d alpha (M) | per-dimension band blocks | pad | float64 scalars] -- reproduces the single-process buffer.
The per-shard buffers are produced by a torch restatement of the kernel's outputs built on the oracle's stencil
(test infrastructure only)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import vggp_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def kernel_outputs_cpu(meshes, X, y, alpha, bands, obs_dtype):
    """What vggp_obs_fwd_bwd accumulates for one shard (B1 family): raw d alpha, band sums, E and n."""
    D = len(meshes)
    Ms = [m.numel() for m in meshes]
    M = int(np.prod(Ms))
    strides = [int(np.prod(Ms[d + 1:])) for d in range(D)]
    N = y.numel()
    sten = [O.b1_stencil(meshes[d], X[:, d]) for d in range(D)]
    inside = torch.ones(N, dtype=torch.bool)
    for c, _, _ in sten:
        inside &= c >= 0
    mu = torch.zeros(N, dtype=torch.float64)
    ps, qs = [], []
    for d in range(D):
        c, wl, wh = sten[d]
        cc = c.clamp(min=0)
        pd, po, qd, qo = bands[d]
        ps.append(wl * wl * pd[cc] + 2 * wl * wh * po[cc] + wh * wh * pd[cc + 1])
        qs.append(wl * wl * qd[cc] + 2 * wl * wh * qo[cc] + wh * wh * qd[cc + 1])
    corner_w, corner_idx = [], []
    for corner in range(2 ** D):
        w = torch.ones(N, dtype=torch.float64)
        idx = torch.zeros(N, dtype=torch.long)
        for d in range(D):
            hi = (corner >> (D - 1 - d)) & 1
            c, wl, wh = sten[d]
            w = w * (wh if hi else wl)
            idx = idx + (c.clamp(min=0) + hi) * strides[d]
        w = torch.where(inside, w, torch.zeros_like(w))
        corner_w.append(w)
        corner_idx.append(idx)
        mu = mu + w * alpha[idx]
    r = y - mu
    ga = torch.zeros(M, dtype=torch.float64)
    for w, idx in zip(corner_w, corner_idx):
        ga.index_add_(0, idx, w * r)
    gband = []
    for d in range(D):
        c, wl, wh = sten[d]
        cc = c.clamp(min=0)
        op = torch.ones(N, dtype=torch.float64)
        oq = torch.ones(N, dtype=torch.float64)
        for e in range(D):
            if e != d:
                op = op * ps[e]
                oq = oq * qs[e]
        n = Ms[d]
        blk = torch.zeros(4 * n, dtype=torch.float64)
        for off, wgt, val in ((0, wl * wl, op), (n, wl * wh, op), (0 + 1, wh * wh, op),
                              (2 * n, wl * wl, oq), (3 * n, wl * wh, oq), (2 * n + 1, wh * wh, oq)):
            blk.index_add_(0, cc + off, torch.where(inside, wgt * val, torch.zeros_like(val)))
        gband.append(blk)
    pp = torch.stack(ps).prod(0)
    qq = torch.stack(qs).prod(0)
    E = (r * r - pp + qq).sum()
    obs = torch.cat([ga] + gband).to(obs_dtype)
    scal = torch.zeros(8, dtype=torch.float64)
    scal[0] = E
    scal[1] = float(N)
    return obs, scal


def pack_gbuf(obs, scal):
    """[obs values | pad to 8 bytes | 8 float64] as one uint8 allocation + the two typed views (plan.gbuf_views)."""
    esz = obs.element_size()
    soff = (obs.numel() * esz + 7) // 8 * 8
    raw = torch.zeros(soff + 64, dtype=torch.uint8)
    ov = raw[: obs.numel() * esz].view(obs.dtype)
    sv = raw[soff: soff + 64].view(torch.float64)
    ov.copy_(obs)
    sv.copy_(scal)
    return raw, ov, sv


def _problem(seed=0):
    g = torch.Generator().manual_seed(seed)
    meshes = [torch.linspace(0, 1, 9), torch.linspace(0, 1, 7)]
    N = 1001
    X = torch.rand(N, 2, generator=g, dtype=torch.float64) * 1.1 - 0.05
    y = torch.randn(N, generator=g, dtype=torch.float64)
    alpha = torch.randn(63, generator=g, dtype=torch.float64)
    bands = [[torch.rand(m.numel(), generator=g, dtype=torch.float64) for _ in range(4)] for m in meshes]
    return meshes, X, y, alpha, bands


def _worker(rank, world, port, obs_dtype_name, out_dir):
    import importlib
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    vdist = importlib.import_module("variational-gridded-gaussian-processes_b200.dist")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        obs_dtype = getattr(torch, obs_dtype_name)
        meshes, X, y, alpha, bands = _problem()
        lo, hi = vdist.shard_bounds(y.numel(), rank, world)
        obs, scal = kernel_outputs_cpu(meshes, X[lo:hi], y[lo:hi], alpha, bands, obs_dtype)
        raw, ov, sv = pack_gbuf(obs, scal)
        vdist.allreduce_gbuf_views(raw, ov, sv, None)
        torch.save({"obs": ov.clone(), "scal": sv.clone(), "lo": lo, "hi": hi}, os.path.join(out_dir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("obs_dtype_name", ["float64", "float32"])
def test_two_rank_allreduce_matches_single_process(tmp_path, obs_dtype_name):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, obs_dtype_name, str(tmp_path)), nprocs=world, join=True)
    meshes, X, y, alpha, bands = _problem()
    obs_dtype = getattr(torch, obs_dtype_name)
    obs_ref, scal_ref = kernel_outputs_cpu(meshes, X, y, alpha, bands, obs_dtype)
    res = [torch.load(os.path.join(str(tmp_path), f"rank{r}.pt")) for r in range(world)]
    assert res[0]["lo"] == 0 and res[0]["hi"] == res[1]["lo"] and res[1]["hi"] == y.numel()
    tol = 1e-12 if obs_dtype_name == "float64" else 2e-5
    for r in res:
        assert torch.allclose(r["obs"].to(torch.float64), obs_ref.to(torch.float64), rtol=tol, atol=tol)
        assert torch.allclose(r["scal"], scal_ref, rtol=1e-12, atol=1e-9)
    # replicas stay bitwise identical: both ranks hold the same reduced buffer
    assert torch.equal(res[0]["obs"], res[1]["obs"]) and torch.equal(res[0]["scal"], res[1]["scal"])


@pytest.mark.parametrize("n,world", [(10, 3), (0, 4), (7, 8), (1 << 26, 8), (1001, 2)])
def test_shard_bounds_partition(n, world):
    import importlib
    vdist = importlib.import_module("variational-gridded-gaussian-processes_b200.dist")
    prev = 0
    sizes = []
    for r in range(world):
        lo, hi = vdist.shard_bounds(n, r, world)
        assert lo == prev and hi >= lo
        sizes.append(hi - lo)
        prev = hi
    assert prev == n and max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        vdist.shard_bounds(n, world, world)


def _reshard_worker(rank, world, port, out_dir, balance="cells"):
    import importlib
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    vdist = importlib.import_module("variational-gridded-gaussian-processes_b200.dist")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(100 + rank)
        n = 500 + 37 * rank
        n_keys = 48
        keys = torch.randint(0, n_keys + 1, (n,), generator=g)          # n_keys = "outside the mesh"
        if balance == "count":                                           # uneven coverage: most observations in few cells
            keys = (keys.double() / (n_keys + 1)).pow(3.0).mul(n_keys + 1).floor().long().clamp(0, n_keys)
        x1 = torch.rand(n, generator=g, dtype=torch.float64)
        x2 = keys.to(torch.float64) + 0.25                               # lets the test recover the key afterwards
        y = torch.rand(n, generator=g, dtype=torch.float64) + rank
        xs, yy = vdist.spatial_reshard([x1, x2], y, keys, n_keys, None, balance=balance)
        torch.save({"x1": xs[0], "x2": xs[1], "y": yy, "in_x1": x1, "in_y": y}, os.path.join(out_dir, f"reshard{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_spatial_reshard_two_ranks(tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_reshard_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    res = [torch.load(os.path.join(str(tmp_path), f"reshard{r}.pt")) for r in range(world)]
    n_keys = 48
    # every observation is on exactly one rank afterwards (multiset equality of (x1, y) pairs)
    before = torch.sort(torch.cat([r["in_x1"] * 3.0 + r["in_y"] for r in res]))[0]
    after = torch.sort(torch.cat([r["x1"] * 3.0 + r["y"] for r in res]))[0]
    assert torch.equal(before, after)
    # rank r owns the keys k with k * world // (n_keys + 1) == r
    for r, d in enumerate(res):
        keys = (d["x2"] - 0.25).round().to(torch.int64)
        assert torch.all((keys * world) // (n_keys + 1) == r)
        assert d["x1"].numel() == d["y"].numel() == d["x2"].numel()


def test_spatial_reshard_count_balanced_two_ranks(tmp_path):
    """balance="count": ranks own disjoint contiguous cell ranges cut at the quantiles of the global histogram, so the
    observation counts are even although three quarters of the observations sit in the first third of the cells."""
    world = 2
    mp.spawn(_reshard_worker, args=(world, _free_port(), str(tmp_path), "count"), nprocs=world, join=True)
    res = [torch.load(os.path.join(str(tmp_path), f"reshard{r}.pt")) for r in range(world)]
    before = torch.sort(torch.cat([r["in_x1"] * 3.0 + r["in_y"] for r in res]))[0]
    after = torch.sort(torch.cat([r["x1"] * 3.0 + r["y"] for r in res]))[0]
    assert torch.equal(before, after)
    keys = [(d["x2"] - 0.25).round().to(torch.int64) for d in res]
    assert keys[0].max() < keys[1].min()                 # contiguous, disjoint cell ranges in rank order
    n = [k.numel() for k in keys]
    biggest_cell = max(torch.bincount(torch.cat(keys)).max().item(), 1)
    assert abs(n[0] - n[1]) <= 2 * biggest_cell          # balanced up to one cell's worth of observations
    assert min(n) > 0.35 * sum(n)

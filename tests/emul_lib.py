"""TEST INFRASTRUCTURE ONLY -- builds and wraps the CPU emulation of the whole C ABI (tests/host_emul): csrc/vggp.cu with
its kernel launches rewritten for the SIMT emulator, compiled by g++.  "Device" pointers are numpy arrays.  Used to
exercise device code and host glue that have not run on a GPU yet; never imported by the product package."""
import ctypes as C
import importlib
import os
import shutil
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
EMU = os.path.join(HERE, "host_emul")
BUILD = os.path.join(EMU, "_build")
CSRC = os.path.join(ROOT, "variational-gridded-gaussian-processes_b200", "csrc")
OUT = os.path.join(BUILD, "libvggp_full_emul.so")
GEN = os.path.join(BUILD, "vggp_emul.cpp")


def _deps():
    d = [os.path.join(EMU, f) for f in ("emul_core.cpp", "make_full_emul.py")]
    d.append(os.path.join(EMU, "fake_cuda", "cuda_runtime.h"))
    d.append(os.path.join(EMU, "fake_cuda", "cub", "device", "device_radix_sort.cuh"))
    d += [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    d.append(os.path.join(ROOT, "include", "vggp.h"))
    return d


def build():
    if shutil.which("g++") is None:
        return None
    os.makedirs(BUILD, exist_ok=True)
    if os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in _deps()):
        return OUT
    subprocess.run([sys.executable, os.path.join(EMU, "make_full_emul.py"), os.path.join(CSRC, "vggp.cu"), GEN, CSRC],
                   check=True, stdout=subprocess.DEVNULL)
    subprocess.run(["g++", "-O1", "-std=c++17", "-fPIC", "-ffp-contract=off", "-fno-extern-tls-init", "-shared",
                    "-Wno-unknown-pragmas", "-Wno-attributes", "-I", os.path.join(EMU, "fake_cuda"), "-o", OUT,
                    os.path.join(EMU, "emul_core.cpp"), GEN], check=True)
    return OUT


def load():
    path = build()
    if path is None:
        return None
    L = importlib.import_module("variational-gridded-gaussian-processes_b200._lib")
    lib = C.CDLL(path)
    for name, (res, args) in L.SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib, L


def ptr(a):
    return a.ctypes.data if a is not None else None


class EmuPlan:
    """The calls of plan.GridPlan against the emulated library, numpy in / numpy out."""

    def __init__(self, lib, L, family, meshes, dtype):
        self.lib, self.L, self.family, self.dtype = lib, L, family, np.dtype(dtype)
        self.D = len(meshes)
        self.meshes = [np.ascontiguousarray(m, dtype=np.float32) for m in meshes]
        nk = (C.c_int * self.D)(*[m.size for m in self.meshes])
        ptrs = (C.POINTER(C.c_float) * self.D)(*[m.ctypes.data_as(C.POINTER(C.c_float)) for m in self.meshes])
        h = C.c_void_p()
        self.check(lib.vggp_plan_create(C.byref(h), family, self.D, nk, ptrs, L.F32 if self.dtype == np.float32 else L.F64, 0))
        self.h = h
        dims, M, Dd = (C.c_int * 3)(), C.c_int64(), C.c_int()
        self.check(lib.vggp_plan_dims(h, C.byref(Dd), dims, C.byref(M)))
        self.m_per_dim = [int(dims[d]) for d in range(self.D)]
        self.M = int(M.value)
        self.L_total = sum(n * n for n in self.m_per_dim)
        ne, so, ns, tot = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        self.check(lib.vggp_gbuf_layout(h, C.byref(ne), C.byref(so), C.byref(ns), C.byref(tot)))
        self.gbuf_obs_elems, self.gbuf_scalar_offset, self.gbuf_bytes = int(ne.value), int(so.value), int(tot.value)
        self.gbuf = np.zeros(self.gbuf_bytes, dtype=np.uint8)

    def check(self, rc):
        if rc != 0:
            raise RuntimeError(f"status {rc}: {self.lib.vggp_last_error().decode()}")

    def close(self):
        if self.h is not None:
            self.lib.vggp_plan_destroy(self.h)
            self.h = None

    def grid_forward(self, theta, m, Lcat):
        self._keep = (theta, m, Lcat)
        self.check(self.lib.vggp_grid_forward(self.h, ptr(theta), ptr(m), ptr(Lcat), None))

    def _xptrs(self, xs):
        return (C.c_void_p * self.D)(*[ptr(x) for x in xs])

    def pack(self, xs, y, sort_by_cell=True):
        n = y.size
        npk, run = C.c_int64(), C.c_int()
        self.check(self.lib.vggp_obs_pack_geometry(self.h, n, C.byref(npk), C.byref(run)))
        xp = [np.zeros(int(npk.value), dtype=self.dtype) for _ in range(self.D)]
        yp = np.zeros(int(npk.value), dtype=self.dtype)
        if n > 0:
            self.check(self.lib.vggp_obs_pack(self.h, self._xptrs(xs), ptr(y), n, 1 if sort_by_cell else 0,
                                              self._xptrs(xp), ptr(yp), None))
        return ("packed", xp, yp, n)

    def bin(self, xs, y, run_cap=256):
        desc = self.L.BinnedDesc()
        self.check(self.lib.vggp_obs_bin_prepare(self.h, self._xptrs(xs), y.size, run_cap, C.byref(desc), None))
        raw = np.zeros(int(desc.bytes) + 256, dtype=np.uint8)
        off = (-raw.ctypes.data) % 256
        buf = raw[off:off + int(desc.bytes)]
        self.check(self.lib.vggp_obs_bin_pack(self.h, C.byref(desc), self._xptrs(xs), ptr(y), ptr(buf), None))
        return ("binned", buf, desc, raw)

    def obs_fwd_bwd(self, obs, y=None, gbuf=None):
        g = self.gbuf if gbuf is None else gbuf
        if isinstance(obs, tuple) and obs[0] == "packed":
            self.check(self.lib.vggp_obs_fwd_bwd_packed(self.h, self._xptrs(obs[1]), ptr(obs[2]), obs[3], ptr(g), None))
        elif isinstance(obs, tuple) and obs[0] == "binned":
            self.check(self.lib.vggp_obs_fwd_bwd_binned(self.h, C.byref(obs[2]), ptr(obs[1]), ptr(g), None))
        else:
            self.check(self.lib.vggp_obs_fwd_bwd(self.h, self._xptrs(obs), ptr(y), y.size, ptr(g), None))

    def grid_backward(self, theta, m, Lcat, ell_scale):
        out = np.zeros(4)
        dtheta = np.zeros(2 * self.D + 1)
        dm = np.zeros(self.M)
        dL = np.zeros(self.L_total)
        self.check(self.lib.vggp_grid_backward(self.h, ptr(theta), ptr(m), ptr(Lcat), ptr(self.gbuf), float(ell_scale),
                                               ptr(out), ptr(dtheta), ptr(dm), ptr(dL), None))
        return out, dtheta, dm, dL

    def step(self, theta, m, Lcat, obs, y=None, ell_scale=1.0):
        self.grid_forward(theta, m, Lcat)
        self.obs_fwd_bwd(obs, y)
        return self.grid_backward(theta, m, Lcat, ell_scale)

    def b1_stencil(self, dim, x):
        x = np.ascontiguousarray(x, dtype=self.dtype)
        c = np.zeros(x.size, dtype=np.int32)
        wl, wh = np.zeros_like(x), np.zeros_like(x)
        self.check(self.lib.vggp_b1_stencil(self.h, dim, ptr(x), x.size, ptr(c), ptr(wl), ptr(wh), None))
        return c, wl, wh

    def features_dense(self, dim, x, theta=None):
        x = np.ascontiguousarray(x, dtype=self.dtype)
        phi = np.zeros((self.m_per_dim[dim], x.size), dtype=self.dtype)
        self.check(self.lib.vggp_features_dense(self.h, dim, ptr(x), x.size, ptr(theta), ptr(phi), None))
        return phi

    def predict(self, xs):
        xs = [np.ascontiguousarray(x, dtype=self.dtype) for x in xs]
        mean, var = np.zeros_like(xs[0]), np.zeros_like(xs[0])
        self.check(self.lib.vggp_predict(self.h, self._xptrs(xs), xs[0].size, ptr(mean), ptr(var), None))
        return mean, var

    def predict_metrics(self, xs, y):
        xs = [np.ascontiguousarray(x, dtype=self.dtype) for x in xs]
        y = np.ascontiguousarray(y, dtype=self.dtype)
        out = np.full(4, np.nan)
        self.check(self.lib.vggp_predict_metrics(self.h, self._xptrs(xs), ptr(y), y.size, ptr(out), None))
        return out

    def read_info(self):
        info = C.c_int(0)
        self.check(self.lib.vggp_read_info(self.h, C.byref(info), None))
        return int(info.value)

import ctypes, json, os, sys, torch
sys.path.insert(0, "/root/repo")
import vggp_b200 as vg
lib = vg._lib.load()
dev = torch.device("cuda", 0)
st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
def run(batch, m, n, k, sk, auto):
    lib.vggp_debug_auto_splitk(auto)
    A = torch.randn(batch, m, k, dtype=torch.float64, device=dev); B = torch.randn(batch, k, n, dtype=torch.float64, device=dev)
    C = torch.zeros(batch, m, n, dtype=torch.float64, device=dev)
    def f():
        rc = lib.vggp_gemm_f64(1, batch, m, n, k, ctypes.c_double(1.0), ctypes.c_void_p(A.data_ptr()), k, 1, m * k if batch > 1 else 0,
                               ctypes.c_void_p(B.data_ptr()), n, 1, k * n if batch > 1 else 0, ctypes.c_double(0.0),
                               ctypes.c_void_p(C.data_ptr()), n, 1, m * n, sk, st)
        assert rc == 0
    for _ in range(3): f()
    torch.cuda.synchronize()
    # graph to remove CPU launch overhead
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50): f()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 50
    print(json.dumps({"batch": batch, "mnk": [m, n, k], "splitk": sk, "auto": auto, "us": ms * 1e3, "tflops": 2.0 * batch * m * n * k / ms / 1e9}), flush=True)
for args in [(1,511,511,511,1,0),(1,511,511,511,1,1),(1,511,511,511,2,0),(1,511,511,511,4,0),(2,511,511,511,1,0),(4,511,511,511,1,0),(4,511,511,128,1,0),(1,512,512,512,1,1),(8,511,511,511,1,0)]:
    run(*args)

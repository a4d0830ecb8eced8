"""Micro-benchmark of the tcgen05 GEMM at the NAR shapes of BASELINE configs[2] (development aid)."""
import ctypes as C, sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mars5_tts_b200 import capi
from mars5_tts_b200.capi import ptr
lib = capi.load(); cfg = capi.ModelCfg(); ctx = C.c_void_p(); assert lib.m5_create(0, C.byref(cfg), None, 0, C.byref(ctx)) == 0
stream = torch.cuda.ExternalStream(lib.m5_stream(ctx))
DEV = "cuda:0"
def bench(M, N, K, mode, acc=0, bn=0, iters=8):
    A = torch.randn(M, K, device=DEV).half(); W = (torch.randn(N, K, device=DEV) * 0.03).half()
    bias = torch.randn(N, device=DEV)
    if mode == capi.OUT_F32: ldc = (N + 3) // 4 * 4; out = torch.zeros(M, ldc, device=DEV)
    elif mode == capi.OUT_F16: out = torch.zeros(M, N, device=DEV, dtype=torch.float16); ldc = N
    else: out = torch.zeros(M, N // 2, device=DEV, dtype=torch.float16); ldc = N // 2
    torch.cuda.synchronize()
    def run(): assert lib.m5_dbg_gemm(ctx, ptr(A), ptr(W), M, N, K, 0, ptr(bias), None, ptr(out), None, ldc, mode, 0, acc, bn) == 0
    for _ in range(2): run()
    lib.m5_sync(ctx)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(iters): run()
    e1.record(stream); lib.m5_sync(ctx); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    # cuBLAS for comparison
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(2): A @ W.T
    torch.cuda.synchronize(); t0.record()
    for _ in range(iters): A @ W.T
    t1.record(); torch.cuda.synchronize()
    cms = t0.elapsed_time(t1) / iters
    fl = 2.0 * M * N * K
    print(f"M={M:7d} N={N:5d} K={K:5d} mode={mode} acc={acc} bn={bn or 'auto'}: {ms:8.3f} ms {fl/ms/1e9:8.1f} TF/s | cuBLAS f16 {cms:8.3f} ms {fl/cms/1e9:8.1f} TF/s", flush=True)
M = 153552
if len(sys.argv) > 1 and sys.argv[1] == "one":
    bench(M, 3072, 1024, capi.OUT_F16, 0, iters=2)
    sys.exit(0)
for (N, K, mode, acc) in [(3072, 1024, capi.OUT_F16, 0), (1024, 1024, capi.OUT_F32, 1), (1024, 1024, capi.OUT_F16, 0), (6144, 1024, capi.OUT_SWIGLU_F16, 0),
                          (1024, 3072, capi.OUT_F32, 1), (1024, 3072, capi.OUT_F32, 0), (2048, 1024, capi.OUT_F16, 0), (1025, 1024, capi.OUT_F32, 0)]:
    bench(M, N, K, mode, acc)
for (N, K, mode, acc) in [(3072, 1024, capi.OUT_F16, 0), (1024, 1024, capi.OUT_F32, 1), (6144, 1024, capi.OUT_SWIGLU_F16, 0), (1024, 3072, capi.OUT_F32, 1)]:
    bench(M, N, K, mode, acc, bn=256)
bench(M, 3072, 1024, capi.OUT_F16, 0, bn=128)
bench(4352, 3072, 1024, capi.OUT_F16, 0)
bench(18784, 4608, 1536, capi.OUT_F16, 0)

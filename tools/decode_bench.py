"""Per-kernel timing of the AR decode step at B=32 (development aid): skinny GEMMs, decode attention, norm."""
import ctypes as C, sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mars5_tts_b200 import capi
from mars5_tts_b200.capi import ptr
lib = capi.load(); cfg = capi.ModelCfg(); ctx = C.c_void_p(); assert lib.m5_create(0, C.byref(cfg), None, 0, C.byref(ctx)) == 0
stream = torch.cuda.ExternalStream(lib.m5_stream(ctx)); DEV = "cuda:0"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
def timeit(fn, iters=20):
    for _ in range(3): fn()
    lib.m5_sync(ctx); ts = []
    for _ in range(iters):
        flush.zero_(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); fn(); e1.record(stream); lib.m5_sync(ctx); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort(); return ts[len(ts) // 2]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
tot = 0.0
for name, N, K, sw in [("wqkv", 4608, 1536, 0), ("wo", 1536, 1536, 0), ("w13", 7168, 1536, 1), ("w2", 1536, 3584, 0), ("vocab", 8000, 1536, 0)]:
    X = torch.randn(B, K, device=DEV).half(); W = (torch.randn(N, K, device=DEV) * 0.03).half()
    o32 = torch.zeros(B, N, device=DEV); o16 = torch.zeros(B, N // 2, device=DEV, dtype=torch.float16)
    f = (lambda: lib.m5_dbg_skinny(ctx, ptr(X), ptr(W), B, N, K, None, ptr(o16), N // 2, 1, 0)) if sw else (lambda: lib.m5_dbg_skinny(ctx, ptr(X), ptr(W), B, N, K, ptr(o32), None, N, 0, 1))
    us = timeit(f); gb = N * K * 2 / 1e9
    print(f"skinny {name:6s} N={N} K={K}: {us:7.1f} us  {gb / us * 1e6 / 1e3:6.2f} TB/s (cold L2)", flush=True)
    tot += us * (26 if name != "vocab" else 1)
H, D = 24, 1536
for L in (600, 1300, 2000):
    W_ = 2100
    q = torch.randn(B, D, device=DEV).half(); kc = torch.randn(B, W_, D, device=DEV).half(); vc = torch.randn(B, W_, D, device=DEV).half()
    kv = torch.full((B,), L, dtype=torch.int32, device=DEV); out = torch.zeros(B, D, device=DEV, dtype=torch.float16)
    us = timeit(lambda: lib.m5_dbg_decode_attn(ctx, ptr(q), ptr(kc), ptr(vc), B, H, W_, ptr(kv), ptr(out), 17))
    gb = B * L * D * 2 * 2 / 1e9
    print(f"decode_attn L={L}: {us:7.1f} us  {gb / us * 1e6 / 1e3:6.2f} TB/s", flush=True)
    if L == 1300: tot += us * 26
x = torch.randn(B, D, device=DEV); g = torch.randn(D, device=DEV); o = torch.zeros(B, D, device=DEV, dtype=torch.float16)
us = timeit(lambda: lib.m5_dbg_norm(ctx, ptr(x), B, D, ptr(g), None, 1e-5, 1, ptr(o), None)); print(f"rmsnorm [B,{D}]: {us:6.1f} us"); tot += us * 53
print(f"sum of kernel times per decode step (L=1300): {tot / 1e3:.2f} ms (+ rope/embed/sampler ~0.15 ms)")

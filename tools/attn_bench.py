"""Attention micro-benchmark at the NAR shapes (development aid): mma.sync kernel (impl 1) vs tcgen05 kernel (impl 2)."""
import ctypes as C, sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mars5_tts_b200 import capi
from mars5_tts_b200.capi import ptr
lib = capi.load(); cfg = capi.ModelCfg(); ctx = C.c_void_p(); assert lib.m5_create(0, C.byref(cfg), None, 0, C.byref(ctx)) == 0
stream = torch.cuda.ExternalStream(lib.m5_stream(ctx)); DEV = "cuda:0"
H, D = 16, 1024
def bench(nseq, S, Kl, impl, iters=5):
    Q = torch.randn(nseq * S, 3 * D, device=DEV).half()
    if Kl is None:
        Kp, Vp, ldk, krows, kl = Q[:, D:], Q[:, 2 * D:], 3 * D, nseq * S, S
    else:
        KV = torch.randn(nseq * Kl, 2 * D, device=DEV).half(); Kp, Vp, ldk, krows, kl = KV, KV[:, D:], 2 * D, nseq * Kl, Kl
    O = torch.zeros(nseq * S, D, device=DEV, dtype=torch.float16)
    i32 = lambda v: torch.tensor(v, dtype=torch.int32, device=DEV)
    qs, ql = i32([i * S for i in range(nseq)]), i32([S] * nseq); ks, kls = i32([i * kl for i in range(nseq)]), i32([kl] * nseq)
    def run():
        rc = lib.m5_dbg_attn(ctx, ptr(Q), C.c_void_p(Kp.data_ptr()), C.c_void_p(Vp.data_ptr()), 3 * D, ldk, ldk, ptr(O), D, H, nseq, S,
                             ptr(qs), ptr(ql), ptr(ks), ptr(kls), 0, impl, nseq * S, krows)
        assert rc == 0
    for _ in range(2): run()
    lib.m5_sync(ctx)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(iters): run()
    e1.record(stream); lib.m5_sync(ctx); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    fl = 4.0 * 64 * H * nseq * S * kl
    print(f"nseq={nseq} S={S} K={kl} impl={impl}: {ms:8.3f} ms {fl / ms / 1e9:8.1f} TF/s", flush=True)
    return O
def bench_split(nseq, S, kpair, iters=5):
    """tcgen05 kernel with value pairs (and key pairs when kpair): the `mixed` / `mixed8` (kpair) and `mixed8k` self-attention."""
    Q = torch.randn(nseq * S, 3 * D, device=DEV).half(); L = (torch.randn(nseq * S, 3 * D, device=DEV) * 2e-4).half()
    O, Ol = torch.zeros(nseq * S, D, device=DEV, dtype=torch.float16), torch.zeros(nseq * S, D, device=DEV, dtype=torch.float16)
    i32 = lambda v: torch.tensor(v, dtype=torch.int32, device=DEV)
    qs, ql = i32([i * S for i in range(nseq)]), i32([S] * nseq)
    at = lambda t, off: C.c_void_p(t.data_ptr() + 2 * off)
    def run():
        rc = lib.m5_dbg_attn_split(ctx, ptr(Q), at(Q, D), at(Q, 2 * D), at(L, D) if kpair else None, at(L, 2 * D), 3 * D, 3 * D, 3 * D,
                                   ptr(O), ptr(Ol), D, H, nseq, S, ptr(qs), ptr(ql), ptr(qs), ptr(ql), nseq * S, nseq * S)
        assert rc == 0
    for _ in range(2): run()
    lib.m5_sync(ctx)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(iters): run()
    e1.record(stream); lib.m5_sync(ctx); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    fl = 4.0 * 64 * H * nseq * S * S
    print(f"nseq={nseq} S={S} split kernel, keys {'pairs' if kpair else 'single'}: {ms:8.3f} ms {fl / ms / 1e9:8.1f} TF/s algorithmic "
          f"({fl * (2.0 if kpair else 1.5) / ms / 1e9:.1f} executed)", flush=True)
if len(sys.argv) > 1 and sys.argv[1] == 'split':
    bench(64, 2399, None, 2); bench_split(64, 2399, True); bench_split(64, 2399, False); sys.exit(0)
if len(sys.argv) > 1 and sys.argv[1] == 'one':
    bench(64, 2399, None, 2, iters=2); sys.exit(0)
for impl in (1, 2):
    bench(64, 2399, None, impl)
    bench(64, 2399, 137, impl)
    bench(64, 137, None, impl)
    bench(33, 451, None, impl)

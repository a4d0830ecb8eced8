"""Checks and times the tcgen05 flash-attention kernel through the C ABI (development aid).

Compares against an fp32 torch softmax(QK^T/8)V on ragged sequences, including keys whose magnitude grows tile after
tile (forces the exact/rescale path of the lazy softmax on every tile), keys that shrink (stale maximum stays valid) and
one late outlier tile; then times the NAR self-attention shape (64 sequences x 2399 tokens x 16 heads).
"""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mars5_tts_b200 import capi
from mars5_tts_b200.capi import ptr

lib = capi.load(); cfg = capi.ModelCfg(); ctx = C.c_void_p()
assert lib.m5_create(0, C.byref(cfg), None, 0, C.byref(ctx)) == 0
stream = torch.cuda.ExternalStream(lib.m5_stream(ctx)); DEV = "cuda:0"
H, D = 8, 512


def run(Q, K, V, qlens, klens, iters=0):
    nseq = len(qlens)
    O = torch.zeros(Q.shape[0], D, device=DEV, dtype=torch.float16)
    i32 = lambda v: torch.tensor(v, dtype=torch.int32, device=DEV)
    cs = lambda v: [sum(v[:i]) for i in range(len(v))]
    qs, ql, ks, kl = i32(cs(qlens)), i32(qlens), i32(cs(klens)), i32(klens)
    def go():
        rc = lib.m5_dbg_attn(ctx, ptr(Q), ptr(K), ptr(V), D, D, D, ptr(O), D, H, nseq, max(qlens), ptr(qs), ptr(ql), ptr(ks), ptr(kl),
                             0, 2, Q.shape[0], K.shape[0])
        assert rc == 0, rc
    go(); lib.m5_sync(ctx)
    ms = None
    if iters:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(iters): go()
        e1.record(stream); lib.m5_sync(ctx); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
    return O, ms


def ref(Q, K, V, qlens, klens):
    out = torch.zeros(Q.shape[0], D, device=DEV)
    qo = ko = 0
    for ql, kl in zip(qlens, klens):
        q = Q[qo:qo + ql].float().view(ql, H, 64).transpose(0, 1)
        k = K[ko:ko + kl].float().view(kl, H, 64).transpose(0, 1)
        v = V[ko:ko + kl].float().view(kl, H, 64).transpose(0, 1)
        p = torch.softmax(q @ k.transpose(1, 2) / 8.0, dim=-1)
        out[qo:qo + ql] = (p @ v).transpose(0, 1).reshape(ql, D)
        qo += ql; ko += kl
    return out


def case(name, qlens, klens, kscale=None, tol=4e-3):
    torch.manual_seed(1)
    Q = torch.randn(sum(qlens), D, device=DEV).half()
    K = torch.randn(sum(klens), D, device=DEV)
    if kscale is not None:
        ko = 0
        for kl in klens:
            t = torch.arange(kl, device=DEV) // 128
            K[ko:ko + kl] *= kscale(t, (kl + 127) // 128).unsqueeze(1)
            ko += kl
    K = K.half()
    V = torch.randn(sum(klens), D, device=DEV).half()
    O, _ = run(Q, K, V, qlens, klens)
    R = ref(Q, K, V, qlens, klens)
    err = (O.float() - R).abs().max().item()
    ok = err < tol and bool(torch.isfinite(O.float()).all())
    print(f"{name:28s} max|err|={err:.2e} {'ok' if ok else 'FAIL'}", flush=True)
    return ok


ok = True
ok &= case("ragged self", [300, 129, 1, 511, 128], [300, 129, 1, 511, 128])
ok &= case("cross, short keys", [700, 256], [37, 130])
ok &= case("growing keys (rescale)", [384, 200], [900, 515], kscale=lambda t, n: 1.0 + 4.0 * t)
ok &= case("shrinking keys (stale max)", [384, 200], [900, 515], kscale=lambda t, n: 1.0 + 4.0 * (n - 1 - t))
ok &= case("one late outlier tile", [256], [1280], kscale=lambda t, n: torch.where(t == 7, 12.0, 1.0))

# timing at the NAR shape (64 sequences x 2399 tokens, 16 heads): the kernel is head-count agnostic, use H*2 = 16 heads
H, D = 16, 1024
nseq, S = 64, 2399
QKV = torch.randn(nseq * S, 3 * D, device=DEV).half()
O = torch.zeros(nseq * S, D, device=DEV, dtype=torch.float16)
i32 = lambda v: torch.tensor(v, dtype=torch.int32, device=DEV)
qs, ql = i32([i * S for i in range(nseq)]), i32([S] * nseq)
def go():
    rc = lib.m5_dbg_attn(ctx, ptr(QKV), C.c_void_p(QKV.data_ptr() + 2 * D), C.c_void_p(QKV.data_ptr() + 4 * D), 3 * D, 3 * D, 3 * D, ptr(O), D, H,
                         nseq, S, ptr(qs), ptr(ql), ptr(qs), ptr(ql), 0, 2, nseq * S, nseq * S)
    assert rc == 0
for _ in range(2): go()
lib.m5_sync(ctx)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(5): go()
e1.record(stream); lib.m5_sync(ctx); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"NAR shape 64x2399x16h: {ms:.3f} ms  {4.0 * 64 * H * nseq * S * S / ms / 1e9:.1f} TFLOP/s  {'PARITY OK' if ok else 'PARITY FAIL'}", flush=True)
sys.exit(0 if ok else 1)

"""Kernel-level parity on the B200: every hand-written kernel is called through the C ABI (m5_dbg_*) and compared
with a plain PyTorch fp32 evaluation of the same op on the same seeded inputs."""
import ctypes as C

import pytest
import torch

from mars5_tts_b200 import capi
from mars5_tts_b200.capi import ptr

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _sync(lib, ctx):
    capi.check(ctx, lib.m5_sync(ctx), "m5_sync")


def _gemm(lib, ctx, A, W, *, bias=None, colscale=None, mode=capi.OUT_F32, act=0, accumulate=0, out=None, kwrap=0,
          force_bn=0, out_lo=None):
    M, K = A.shape
    N = W.shape[0]
    if out is None:
        if mode == capi.OUT_F32:
            out = torch.empty(M, N, device=DEV, dtype=torch.float32)
        elif mode in (capi.OUT_F16, capi.OUT_F16_SPLIT):
            out = torch.empty(M, N, device=DEV, dtype=torch.float16)
        else:
            out = torch.empty(M, N // 2, device=DEV, dtype=torch.float16)
    rc = lib.m5_dbg_gemm(ctx, ptr(A), ptr(W), M, N, K, kwrap, ptr(bias), ptr(colscale), ptr(out), ptr(out_lo),
                         out.stride(0), mode, act, accumulate, force_bn)
    capi.check(ctx, rc, "m5_dbg_gemm")
    _sync(lib, ctx)
    return out


@pytest.mark.parametrize("M,N,K,bn", [(128, 256, 64, 0), (128, 128, 128, 128), (300, 1025, 1024, 0),
                                      (4096, 3072, 1024, 0), (586, 4608, 1536, 0), (77, 384, 1152, 0),
                                      (1000, 64, 128, 64), (2500, 6144, 1024, 256), (2500, 6144, 1024, 512),
                                      (300, 1025, 1024, 512), (128, 256, 64, 512), (40000, 3072, 1024, 0)])
def test_gemm_f32(m5lib, bare_ctx, M, N, K, bn):
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N)
    A = (torch.randn(M, K, generator=g) * 0.5).half().to(DEV)
    W = (torch.randn(N, K, generator=g) * 0.05).half().to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    out = torch.zeros(M, (N + 3) // 4 * 4, device=DEV)
    _gemm(m5lib, bare_ctx, A, W, bias=bias, out=out, force_bn=bn)
    ref = A.float() @ W.float().T + bias
    err = (out[:, :N] - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= 2e-3 * max(scale, 1.0), f"gemm {M}x{N}x{K}: max err {err} (scale {scale})"


def test_gemm_accumulate_and_f16(m5lib, bare_ctx):
    g = torch.Generator(device="cpu").manual_seed(3)
    M, N, K = 700, 1024, 3072
    A = (torch.randn(M, K, generator=g) * 0.3).half().to(DEV)
    W = (torch.randn(N, K, generator=g) * 0.02).half().to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    resid = torch.randn(M, N, generator=g).to(DEV)
    out = resid.clone()
    _gemm(m5lib, bare_ctx, A, W, bias=bias, out=out, accumulate=1)
    ref = resid + A.float() @ W.float().T + bias
    assert (out - ref).abs().max().item() < 2e-3
    o16 = _gemm(m5lib, bare_ctx, A, W, bias=bias, mode=capi.OUT_F16)
    ref16 = (A.float() @ W.float().T + bias)
    assert (o16.float() - ref16).abs().max().item() < 5e-3


def test_gemm_swiglu_gelu_colscale(m5lib, bare_ctx):
    g = torch.Generator(device="cpu").manual_seed(4)
    M, D, F = 513, 1024, 3072
    A = (torch.randn(M, D, generator=g) * 0.5).half().to(DEV)
    Ww = (torch.randn(F, D, generator=g) * 0.03).half()
    Wv = (torch.randn(F, D, generator=g) * 0.03).half()
    inter = torch.stack([Ww, Wv], dim=1).reshape(2 * F, D).contiguous().to(DEV)  # rows (W_0, V_0, W_1, V_1, ...)
    o = _gemm(m5lib, bare_ctx, A, inter, mode=capi.OUT_SWIGLU_F16)
    ref = torch.nn.functional.silu(A.float() @ Ww.float().to(DEV).T) * (A.float() @ Wv.float().to(DEV).T)
    assert (o.float() - ref).abs().max().item() < 5e-3 * max(1.0, ref.abs().max().item())
    # gelu + bias then colscale, fp32 accumulate into residual (Vocos ConvNeXt pointwise convs)
    W1 = (torch.randn(1152, 384, generator=g) * 0.05).half().to(DEV)
    b1 = torch.randn(1152, generator=g).to(DEV)
    A2 = (torch.randn(M, 384, generator=g)).half().to(DEV)
    o2 = _gemm(m5lib, bare_ctx, A2, W1, bias=b1, act=capi.ACT_GELU, mode=capi.OUT_F16)
    ref2 = torch.nn.functional.gelu(A2.float() @ W1.float().T + b1)
    assert (o2.float() - ref2).abs().max().item() < 5e-3 * max(1.0, ref2.abs().max().item())
    cs = torch.randn(1152, generator=g).to(DEV)
    o3 = _gemm(m5lib, bare_ctx, A2, W1, bias=b1, colscale=cs)
    ref3 = (A2.float() @ W1.float().T + b1) * cs
    assert (o3 - ref3).abs().max().item() < 2e-3 * max(1.0, ref3.abs().max().item())


def test_gemm_split_precision(m5lib, bare_ctx):
    """kwrap mode: A = [hi | lo] halves of an fp32 activation -> fp32-class accuracy."""
    g = torch.Generator(device="cpu").manual_seed(5)
    M, N, K = 515, 1024, 1024
    X = torch.randn(M, K, generator=g).to(DEV)
    W = (torch.randn(N, K, generator=g) * 0.03).half().to(DEV)
    hi = X.half()
    lo = (X - hi.float()).half()
    A = torch.cat([hi, lo], dim=1).contiguous()
    o = _gemm(m5lib, bare_ctx, A, W, kwrap=K)
    ref = (X.double() @ W.double().T).float()
    err = (o - ref).abs().max().item()
    plain = (_gemm(m5lib, bare_ctx, hi.contiguous(), W) - ref).abs().max().item()
    assert err < 3e-5, (err, plain)
    assert err < plain


@pytest.mark.parametrize("rms", [0, 1])
def test_norm_rows(m5lib, bare_ctx, rms):
    g = torch.Generator(device="cpu").manual_seed(6)
    M, D = 1000, 1536 if rms else 1024
    x = (torch.randn(M, D, generator=g) * 3 + 0.5).to(DEV)
    gamma = torch.randn(D, generator=g).to(DEV)
    beta = None if rms else torch.randn(D, generator=g).to(DEV)
    out = torch.empty(M, D, device=DEV, dtype=torch.float16)
    lo = torch.empty_like(out)
    eps = 1e-5 if rms else 4e-5
    rc = m5lib.m5_dbg_norm(bare_ctx, ptr(x), M, D, ptr(gamma), ptr(beta), eps, rms, ptr(out), ptr(lo))
    capi.check(bare_ctx, rc, "norm")
    _sync(m5lib, bare_ctx)
    if rms:
        ref = x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps) * gamma
    else:
        ref = torch.nn.functional.layer_norm(x, (D,), gamma, beta, eps)
    assert (out.float() + lo.float() - ref).abs().max().item() < 1e-4
    assert (out.float() - ref).abs().max().item() < 4e-3 * ref.abs().max().item()


def _attn_ref(q, k, v, causal):
    # q [Lq, H, 64], k/v [Lk, H, 64] fp32
    s = torch.einsum("qhd,khd->hqk", q, k) * 0.125
    if causal:
        Lq, Lk = q.shape[0], k.shape[0]
        i = torch.arange(Lq, device=q.device)[:, None]
        j = torch.arange(Lk, device=q.device)[None, :]
        s = s.masked_fill(j > i + (Lk - Lq), float("-inf"))
    p = s.softmax(-1)
    return torch.einsum("hqk,khd->qhd", p, v)


@pytest.mark.parametrize("causal,impl", [(0, 1), (1, 1), (0, 2)])
def test_flash_attn_varlen(m5lib, bare_ctx, causal, impl):
    g = torch.Generator(device="cpu").manual_seed(7 + causal)
    H = 16
    q_lens = [1, 63, 64, 65, 200, 451]
    k_lens = q_lens if causal else [5, 64, 130, 1, 333, 451]
    D = H * 64
    Q = torch.randn(sum(q_lens), 3 * D, generator=g).half().to(DEV)  # packed qkv rows: q | k | v
    KV = torch.randn(sum(k_lens), 2 * D, generator=g).half().to(DEV)
    qs = torch.tensor([sum(q_lens[:i]) for i in range(len(q_lens))], dtype=torch.int32, device=DEV)
    ks = torch.tensor([sum(k_lens[:i]) for i in range(len(k_lens))], dtype=torch.int32, device=DEV)
    ql = torch.tensor(q_lens, dtype=torch.int32, device=DEV)
    kl = torch.tensor(k_lens, dtype=torch.int32, device=DEV)
    O = torch.zeros(sum(q_lens), D, device=DEV, dtype=torch.float16)
    if causal:
        Kp, Vp, ldk = Q[:, D:], Q[:, 2 * D:], 3 * D
    else:
        Kp, Vp, ldk = KV, KV[:, D:], 2 * D
    rc = m5lib.m5_dbg_attn(bare_ctx, ptr(Q), C.c_void_p(Kp.data_ptr()), C.c_void_p(Vp.data_ptr()), 3 * D, ldk, ldk,
                           ptr(O), D, H, len(q_lens), max(q_lens), ptr(qs), ptr(ql), ptr(ks), ptr(kl), causal, impl,
                           sum(q_lens), sum(q_lens) if causal else sum(k_lens))
    capi.check(bare_ctx, rc, "attn")
    _sync(m5lib, bare_ctx)
    for i in range(len(q_lens)):
        q = Q[qs[i]:qs[i] + q_lens[i], :D].float().view(-1, H, 64)
        if causal:
            k = Q[qs[i]:qs[i] + q_lens[i], D:2 * D].float().view(-1, H, 64)
            v = Q[qs[i]:qs[i] + q_lens[i], 2 * D:].float().view(-1, H, 64)
        else:
            k = KV[ks[i]:ks[i] + k_lens[i], :D].float().view(-1, H, 64)
            v = KV[ks[i]:ks[i] + k_lens[i], D:].float().view(-1, H, 64)
        ref = _attn_ref(q, k, v, causal).reshape(-1, D)
        got = O[qs[i]:qs[i] + q_lens[i]].float()
        err = (got - ref).abs().max().item()
        assert err < 4e-3, f"seq {i} (q {q_lens[i]}, k {k_lens[i]}): {err}"


def test_decode_attn(m5lib, bare_ctx):
    g = torch.Generator(device="cpu").manual_seed(9)
    B, H, W = 5, 24, 700
    D = H * 64
    q = torch.randn(B, D, generator=g).half().to(DEV)
    kc = torch.randn(B, W, D, generator=g).half().to(DEV)
    vc = torch.randn(B, W, D, generator=g).half().to(DEV)
    lens = [1, 17, 128, 699, 700]
    kv_len = torch.tensor(lens, dtype=torch.int32, device=DEV)
    for n_split in (1, 4, 7):
        out = torch.zeros(B, D, device=DEV, dtype=torch.float16)
        rc = m5lib.m5_dbg_decode_attn(bare_ctx, ptr(q), ptr(kc), ptr(vc), B, H, W, ptr(kv_len), ptr(out), n_split)
        capi.check(bare_ctx, rc, "decode_attn")
        _sync(m5lib, bare_ctx)
        for b in range(B):
            ref = _attn_ref(q[b].float().view(1, H, 64), kc[b, :lens[b]].float().view(-1, H, 64),
                            vc[b, :lens[b]].float().view(-1, H, 64), 0).reshape(-1)
            err = (out[b].float() - ref).abs().max().item()
            assert err < 3e-3, (n_split, b, err)

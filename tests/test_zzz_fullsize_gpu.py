"""Parity at the sizes of BASELINE configs[2] (deep clone, batch 32: 64 packed sequences of 2399 tokens = 153,552 rows in
the NAR transformer; 32 rows x 24 heads x 2000-token contexts in the AR decode step).  The pipelines are pinned on the
tiny model in test_pipeline_gpu.py; here the kernels that carry the step are run at full size through the C ABI and
compared with plain PyTorch fp32 evaluations of the same op -- whole-output for the GEMMs, per-sequence for attention."""
import ctypes as C

import pytest
import torch

from mars5_tts_b200 import capi
from mars5_tts_b200.capi import ptr

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
M_NAR = 153552          # 2 (cond, uncond) x 32 utterances x 2399 decoder positions


def _sync(lib, ctx):
    capi.check(ctx, lib.m5_sync(ctx), "m5_sync")


def _rand(shape, scale, seed):
    g = torch.Generator(device=DEV).manual_seed(seed)
    return (torch.randn(*shape, device=DEV, generator=g) * scale).half()


def _gemm(lib, ctx, A, W, out, *, bias=None, mode=capi.OUT_F32, accumulate=0):
    M, K = A.shape
    rc = lib.m5_dbg_gemm(ctx, ptr(A), ptr(W), M, W.shape[0], K, 0, ptr(bias), None, ptr(out), None, out.stride(0), mode, 0,
                         accumulate, 0)
    capi.check(ctx, rc, "m5_dbg_gemm")
    _sync(lib, ctx)


def _ref_rows(A, W, rows):
    return A[rows].float() @ W.float().T


def test_gemm_nar_shapes_full_m(m5lib, bare_ctx):
    """QKV projection (fp16 out), out-projection (fp32 residual accumulate) and SwiGLU at M = 153,552: every output row
    block is checked against fp32 matmuls of the same fp16 operands (chunked to bound the reference's memory)."""
    A = _rand((M_NAR, 1024), 0.5, 1)
    chunks = [slice(i, min(i + 16384, M_NAR)) for i in range(0, M_NAR, 16384)]
    # fp16 output, N = 3072 (CTA-pair kernel)
    W = _rand((3072, 1024), 0.04, 2)
    bias = torch.randn(3072, device=DEV)
    out = torch.empty(M_NAR, 3072, device=DEV, dtype=torch.float16)
    _gemm(m5lib, bare_ctx, A, W, out, bias=bias, mode=capi.OUT_F16)
    for rows in chunks:
        ref = _ref_rows(A, W, rows) + bias
        err = (out[rows].float() - ref).abs().max().item()
        assert err < 5e-3 * max(1.0, ref.abs().max().item()), (rows, err)
    del out
    # fp32 residual accumulate, N = K = 1024
    W2 = _rand((1024, 1024), 0.04, 3)
    g = torch.Generator(device=DEV).manual_seed(4)
    resid = torch.randn(M_NAR, 1024, device=DEV, generator=g)
    acc = resid.clone()
    _gemm(m5lib, bare_ctx, A, W2, acc, accumulate=1)
    for rows in chunks:
        ref = resid[rows] + _ref_rows(A, W2, rows)
        assert (acc[rows] - ref).abs().max().item() < 2e-3 * max(1.0, ref.abs().max().item()), rows
    del acc, resid
    # SwiGLU over interleaved (W_j, V_j) rows, N = 2 x 3072 -> 3072 fp16 columns
    Ww, Wv = _rand((3072, 1024), 0.03, 5), _rand((3072, 1024), 0.03, 6)
    inter = torch.stack([Ww, Wv], dim=1).reshape(6144, 1024).contiguous()
    o = torch.empty(M_NAR, 3072, device=DEV, dtype=torch.float16)
    _gemm(m5lib, bare_ctx, A, inter, o, mode=capi.OUT_SWIGLU_F16)
    for rows in chunks:
        ref = torch.nn.functional.silu(_ref_rows(A, Ww, rows)) * _ref_rows(A, Wv, rows)
        assert (o[rows].float() - ref).abs().max().item() < 5e-3 * max(1.0, ref.abs().max().item()), rows


def _attn_ref(q, k, v):
    p = torch.softmax(torch.einsum("qhd,khd->hqk", q, k) / 8.0, dim=-1)
    return torch.einsum("hqk,khd->qhd", p, v)


def test_flash_attention_nar_lengths(m5lib, bare_ctx):
    """tcgen05 attention at the NAR decoder length (2399 = 37 key tiles + a 31-key tail), self and cross (137 text
    keys), 16 heads, a few packed sequences; every sequence against an fp32 softmax(QK^T/8)V."""
    H, D = 16, 1024
    q_lens = [2399, 2399, 1500, 2399]
    for k_lens, seed in ((q_lens, 11), ([137, 88, 137, 1], 12)):
        self_attn = k_lens is q_lens
        Q = _rand((sum(q_lens), 3 * D), 1.0, seed)
        KV = Q[:, D:] if self_attn else _rand((sum(k_lens), 2 * D), 1.0, seed + 100)
        Kp, Vp, ldk = (Q[:, D:], Q[:, 2 * D:], 3 * D) if self_attn else (KV, KV[:, D:], 2 * D)
        i32 = lambda v: torch.tensor(v, dtype=torch.int32, device=DEV)
        cs = lambda v: [sum(v[:i]) for i in range(len(v))]
        qs, ql, ks, kl = i32(cs(q_lens)), i32(q_lens), i32(cs(k_lens)), i32(k_lens)
        O = torch.zeros(sum(q_lens), D, device=DEV, dtype=torch.float16)
        rc = m5lib.m5_dbg_attn(bare_ctx, ptr(Q), C.c_void_p(Kp.data_ptr()), C.c_void_p(Vp.data_ptr()), 3 * D, ldk, ldk, ptr(O), D, H,
                               len(q_lens), max(q_lens), ptr(qs), ptr(ql), ptr(ks), ptr(kl), 0, 2, sum(q_lens), sum(k_lens))
        capi.check(bare_ctx, rc, "attn")
        _sync(m5lib, bare_ctx)
        assert torch.isfinite(O.float()).all()
        for i, (qn, kn) in enumerate(zip(q_lens, k_lens)):
            q0, k0 = cs(q_lens)[i], cs(k_lens)[i]
            q = Q[q0:q0 + qn, :D].float().view(qn, H, 64)
            k = Kp[k0:k0 + kn, :D].float().reshape(kn, H, 64)
            v = Vp[k0:k0 + kn, :D].float().reshape(kn, H, 64)
            err = (O[q0:q0 + qn].float() - _attn_ref(q, k, v).reshape(qn, D)).abs().max().item()
            assert err < 4e-3, (self_attn, i, err)


def test_decode_attention_full_batch(m5lib, bare_ctx):
    """Split-KV decode attention at B = 32 rows x 24 heads with contexts up to 2000 cached tokens (the AR window of
    BASELINE configs[2]: 135 + 450 + 1500 tokens < 2100)."""
    B, H, W = 32, 24, 2100
    D = H * 64
    q = _rand((B, D), 1.0, 21)
    kc, vc = _rand((B, W, D), 1.0, 22), _rand((B, W, D), 1.0, 23)
    lens = [1 + (b * 67) % 2000 for b in range(B)]
    lens[0], lens[1], lens[2] = 2000, 256, 257
    kv_len = torch.tensor(lens, dtype=torch.int32, device=DEV)
    out = torch.zeros(B, D, device=DEV, dtype=torch.float16)
    n_split = (W + 255) // 256
    rc = m5lib.m5_dbg_decode_attn(bare_ctx, ptr(q), ptr(kc), ptr(vc), B, H, W, ptr(kv_len), ptr(out), n_split)
    capi.check(bare_ctx, rc, "decode_attn")
    _sync(m5lib, bare_ctx)
    for b in range(B):
        ref = _attn_ref(q[b].float().view(1, H, 64), kc[b, :lens[b]].float().view(-1, H, 64),
                        vc[b, :lens[b]].float().view(-1, H, 64)).reshape(-1)
        err = (out[b].float() - ref).abs().max().item()
        assert err < 3e-3, (b, lens[b], err)

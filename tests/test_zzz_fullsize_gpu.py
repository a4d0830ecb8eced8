"""Parity at the sizes of BASELINE configs[2] (deep clone, batch 32: 64 packed sequences of 2399 tokens = 153,552 rows in
the NAR transformer; 32 rows x 24 heads x 2000-token contexts in the AR decode step).  The pipelines are pinned on the
tiny model in test_pipeline_gpu.py; here the kernels that carry the step are run at full size through the C ABI and
compared with plain PyTorch fp32 evaluations of the same op -- whole-output for the GEMMs, per-sequence for attention."""
import ctypes as C

import numpy as np
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


def test_flash_attention_split_kv(m5lib, bare_ctx):
    """"mixed" numerics: keys and values as fp16 (hi, lo) pairs, queries / probabilities single fp16, output a pair.
    Against an fp32 softmax(Q K^T / 8) V of the UNROUNDED keys / values the pair kernel must be several times closer
    than the plain fp16 kernel, at the NAR decoder length and at the cross-attention length."""
    H, D = 16, 1024
    q_lens = [2399, 700, 2399]
    for k_lens, seed in ((q_lens, 31), ([137, 137, 5], 32)):
        g = torch.Generator(device=DEV).manual_seed(seed)
        nq, nk = sum(q_lens), sum(k_lens)
        Q = (torch.randn(nq, D, device=DEV, generator=g)).half()
        K32, V32 = torch.randn(nk, D, device=DEV, generator=g), torch.randn(nk, D, device=DEV, generator=g)
        Kh, Vh = K32.half(), V32.half()
        Kl, Vl = (K32 - Kh.float()).half(), (V32 - Vh.float()).half()
        i32 = lambda v: torch.tensor(v, dtype=torch.int32, device=DEV)
        cs = lambda v: [sum(v[:i]) for i in range(len(v))]
        qs, ql, ks, kl = i32(cs(q_lens)), i32(q_lens), i32(cs(k_lens)), i32(k_lens)
        O, Ol = torch.zeros(nq, D, device=DEV, dtype=torch.float16), torch.zeros(nq, D, device=DEV, dtype=torch.float16)
        rc = m5lib.m5_dbg_attn_split(bare_ctx, ptr(Q), ptr(Kh), ptr(Vh), ptr(Kl), ptr(Vl), D, D, D, ptr(O), ptr(Ol), D, H, len(q_lens),
                                     max(q_lens), ptr(qs), ptr(ql), ptr(ks), ptr(kl), nq, nk)
        capi.check(bare_ctx, rc, "attn_split")
        # mixed8k: keys single fp16 (Klo = NULL), values pairs
        Ok, Okl = torch.zeros(nq, D, device=DEV, dtype=torch.float16), torch.zeros(nq, D, device=DEV, dtype=torch.float16)
        rc = m5lib.m5_dbg_attn_split(bare_ctx, ptr(Q), ptr(Kh), ptr(Vh), None, ptr(Vl), D, D, D, ptr(Ok), ptr(Okl), D, H, len(q_lens),
                                     max(q_lens), ptr(qs), ptr(ql), ptr(ks), ptr(kl), nq, nk)
        capi.check(bare_ctx, rc, "attn_split (single keys)")
        Op = torch.zeros(nq, D, device=DEV, dtype=torch.float16)
        rc = m5lib.m5_dbg_attn(bare_ctx, ptr(Q), ptr(Kh), ptr(Vh), D, D, D, ptr(Op), D, H, len(q_lens), max(q_lens), ptr(qs), ptr(ql),
                               ptr(ks), ptr(kl), 0, 2, nq, nk)
        capi.check(bare_ctx, rc, "attn")
        _sync(m5lib, bare_ctx)
        e_split = e_plain = e_ksingle = e_kh = 0.0
        for i, (qn, kn) in enumerate(zip(q_lens, k_lens)):
            q0, k0 = cs(q_lens)[i], cs(k_lens)[i]
            qv = Q[q0:q0 + qn].float().view(qn, H, 64)
            ref = _attn_ref(qv, K32[k0:k0 + kn].view(kn, H, 64), V32[k0:k0 + kn].view(kn, H, 64)).reshape(qn, D)
            # what the single-key kernel computes exactly: fp16-rounded keys, unrounded values
            ref_kh = _attn_ref(qv, Kh[k0:k0 + kn].float().view(kn, H, 64), V32[k0:k0 + kn].view(kn, H, 64)).reshape(qn, D)
            e_split = max(e_split, ((O[q0:q0 + qn].float() + Ol[q0:q0 + qn].float()) - ref).abs().max().item())
            e_plain = max(e_plain, (Op[q0:q0 + qn].float() - ref).abs().max().item())
            e_ksingle = max(e_ksingle, ((Ok[q0:q0 + qn].float() + Okl[q0:q0 + qn].float()) - ref).abs().max().item())
            e_kh = max(e_kh, ((Ok[q0:q0 + qn].float() + Okl[q0:q0 + qn].float()) - ref_kh).abs().max().item())
        print(f"split-KV attention: max-abs {e_split:.2e} (plain fp16 kernel {e_plain:.2e}; single keys + value pairs {e_ksingle:.2e}, "
              f"{e_kh:.2e} against the softmax of the ROUNDED keys), keys {k_lens}")
        assert e_split < 1e-3 and e_split < 0.5 * e_plain, (k_lens, e_split, e_plain)
        # single keys: the same kernel minus the K_lo pass -- against the softmax of the rounded keys it is as close as the pair
        # kernel is against the unrounded one (kernel correctness); against the unrounded keys it sits between pair and plain
        assert e_kh < 1e-3 and e_kh < 0.5 * e_plain and e_ksingle <= 1.05 * e_plain + 1e-6, (k_lens, e_kh, e_ksingle, e_plain)


# ------------------------------------------------------------------------------------------------ full-size pipelines
@pytest.fixture(scope="module")
def full_engine():
    """FULL model dims (AR 1536 x 26, NAR 1024 x 8/16/3, V = 8000), seeded synthetic reference-format checkpoints."""
    from mars5_tts_b200 import synth, weights
    from mars5_tts_b200.engine import Engine
    torch.set_grad_enabled(False)
    size = synth.FULL
    ar_sd, nar_sd, voc_sd = synth.make_ar_state(size), synth.make_nar_state(size), synth.make_vocos_state(size)
    eng = Engine(ar_sd, nar_sd, voc_sd, size["n_text"], device=0, max_pos=4096)
    cfg = weights.dims_from_state(ar_sd, nar_sd, voc_sd, size["n_text"])
    yield size, ar_sd, nar_sd, eng, cfg
    eng.close()


def test_nar_forward_full_dims_absolute_tolerance(full_engine):
    """BASELINE north_star: <= 1e-3 MAX-ABS on the NAR logits against the reference's fp32 arithmetic, at the full model
    dims and a BASELINE configs[1]-sized sequence (S = 1650 decoder rows, 86 text tokens, 450 reference frames).  `mixed`
    (the mode bench.py and Mars5TTS run) and `precise` must hold the absolute bound; `fast` is reported and held to the
    relative bound only (it does NOT meet the north_star tolerance: fp16 activations have a 2^-11 relative error floor)."""
    from mars5_tts_b200 import capi as cp
    from oracle import nar_oracle
    size, _, nar_sd, eng, cfg = full_engine
    g = torch.Generator().manual_seed(17)
    S, Tc, Pf, t = 1650, 86, 450, 100
    text = torch.randint(0, size["n_text"], (Tc,), generator=g)
    codes = torch.randint(0, 1024, (Pf, 8), generator=g)
    x = torch.randint(0, 1025, (S, 8), generator=g)
    torch.set_num_threads(min(32, torch.get_num_threads()))
    for drop in (False, True):   # conditional and unconditional pass of the classifier-free guidance pair
        ref = nar_oracle.nar_forward(nar_sd, cfg, text, codes, x, t, drop_cond=drop).numpy()
        scale = float(np.abs(ref).max())
        errs = {}
        for name, mode in (("fast", cp.NUM_FAST), ("mixed", cp.NUM_MIXED), ("mixed8", cp.NUM_MIXED8), ("mixed8k", cp.NUM_MIXED8K),
                           ("precise", cp.NUM_PRECISE)):
            # mixed8's fp8 pass lives in the CTA-pair GEMM, which needs >= 74 tile pairs: three copies of the utterance in one
            # packed batch (4950 decoder rows) make the shapes eligible; every copy must give the same logits
            rep = 3 if mode in (cp.NUM_MIXED8, cp.NUM_MIXED8K) else 1
            outs = eng.nar_forward([text.numpy()] * rep, [codes.numpy()] * rep, [x.numpy()] * rep, t, drop_cond=drop, precise=mode)
            got = outs[0]
            if rep > 1:
                assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])
            errs[name] = float(np.abs(got - ref).max())
        print(f"NAR logits at full dims (S={S}, drop_cond={drop}): max|logit| {scale:.2f}; max-abs error " +
              ", ".join(f"{k} {v:.2e}" for k, v in errs.items()))
        assert errs["mixed"] < 1e-3, errs
        assert errs["mixed8"] < 1e-3, errs
        assert errs["mixed8k"] < 1e-3, errs
        assert errs["precise"] < 1e-3, errs
        assert errs["fast"] < 1e-3 * max(1.0, scale), errs


def test_nar_forward_full_dims_uncond_and_batch(full_engine):
    """Unconditional pass (speaker encoder sees only the identity token) and a second, shorter utterance in the same
    packed batch: rows of a batch equal single-utterance runs, mixed numerics, absolute bound."""
    from mars5_tts_b200 import capi as cp
    from oracle import nar_oracle
    size, _, nar_sd, eng, cfg = full_engine
    g = torch.Generator().manual_seed(18)
    texts = [torch.randint(0, size["n_text"], (n,), generator=g) for n in (40, 12)]
    codes = [torch.randint(0, 1024, (n, 8), generator=g) for n in (300, 450)]
    xs = [torch.randint(0, 1025, (n, 8), generator=g) for n in (700, 300)]
    got = eng.nar_forward([a.numpy() for a in texts], [c.numpy() for c in codes], [x.numpy() for x in xs], 7, drop_cond=True,
                          precise=cp.NUM_MIXED)
    for i in range(2):
        ref = nar_oracle.nar_forward(nar_sd, cfg, texts[i], codes[i], xs[i], 7, drop_cond=True).numpy()
        err = float(np.abs(got[i] - ref).max())
        print(f"uncond utterance {i}: max-abs {err:.2e}")
        assert err < 1e-3, (i, err)


def test_ar_forward_full_dims(full_engine):
    """CodecLM.forward at the full dims over a BASELINE-sized prompt (135 text + 450 speech tokens, 450-frame speaker
    reference) against the fp32 oracle.  The AR engine mirrors the reference's GPU arithmetic (fp16 autocast,
    inference.py:263 / SURVEY B.1: fp16 GEMM operands and KV cache, fp32 residual stream), which the reference itself
    only holds to fp16 accuracy against its own CPU fp32 path -- the max-abs error is reported and bounded relative to
    the logit scale; the sampled tokens are pinned bit-exact against the reference in test_pipeline_gpu.py."""
    from oracle import ar_oracle
    size, ar_sd, _, eng, cfg = full_engine
    g = torch.Generator().manual_seed(19)
    P, Pf = 586, 450
    prompt = torch.cat([torch.randint(0, size["n_text"], (136,), generator=g),
                        torch.randint(size["n_text"], size["n_text"] + 1024, (P - 136,), generator=g)])
    spk = torch.randint(0, 1024, (Pf, 8), generator=g)
    got = eng.ar_forward([prompt.tolist()], [spk.numpy()])[0]
    ref = ar_oracle.codeclm_forward(ar_sd, cfg, prompt, spk).numpy()
    err, scale = float(np.abs(got - ref).max()), float(np.abs(ref).max())
    print(f"AR logits at full dims (P={P}): max|logit| {scale:.2f}; max-abs error {err:.2e} (rel {err / scale:.2e})")
    assert err < 2e-3 * max(1.0, scale), (err, scale)
    assert (got.argmax(-1) == ref.argmax(-1)).mean() > 0.995


def test_nar_infer_full_dims_code_agreement(full_engine):
    """Whole reverse loop at the full dims (T = 20, CFG, in-kernel Philox noise keyed by utterance id): the codes of the
    `mixed` and `fast` modes against the `precise` mode (which is pinned bit-exact against the reference's own outputs on
    the tiny model, test_pipeline_gpu.py).  Sampling feeds back through x_t, so a single flipped Gumbel arg-max early on
    decorrelates later steps -- the mismatch RATE is the parity figure."""
    from mars5_tts_b200 import capi as cp
    from mars5_tts_b200.engine import InferenceConfig
    size, _, _, eng, _ = full_engine
    g = torch.Generator().manual_seed(23)
    B, Tc, Pf, N = 2, 60, 200, 1000   # S = 1200 rows: the tcgen05 pair attention (>= 1024) and the fp8 lo pass (>= 74 tile pairs) apply
    texts = [torch.randint(0, size["n_text"], (Tc,), generator=g).numpy().astype(np.int32) for _ in range(B)]
    codes = [torch.randint(0, 1024, (Pf, 8), generator=g).numpy().astype(np.int32) for _ in range(B)]
    l0 = [torch.randint(0, 1024, (N,), generator=g).numpy().astype(np.int32) for _ in range(B)]
    out = {}
    for name, mode in (("precise", cp.NUM_PRECISE), ("mixed", cp.NUM_MIXED), ("mixed8", cp.NUM_MIXED8), ("mixed8k", cp.NUM_MIXED8K),
                       ("fast", cp.NUM_FAST)):
        ncfg = eng.make_nar_cfg(InferenceConfig(), T=20, precise=mode)
        out[name] = np.stack(eng.nar_infer(texts, codes, l0, ncfg, seed=5, utt_ids=[100, 101]))
    mm = {k: float((out[k] != out["precise"]).mean()) for k in ("mixed", "mixed8", "mixed8k", "fast")}
    print(f"NAR codes after T=20 at full dims vs precise: mixed {mm['mixed']:.4%} differ, mixed8 {mm['mixed8']:.4%}, "
          f"mixed8k {mm['mixed8k']:.4%}, fast {mm['fast']:.4%} differ")
    assert mm["mixed"] <= 0.02 and mm["mixed8"] <= 0.02 and mm["mixed8k"] <= 0.02, mm
    assert mm["mixed"] <= mm["fast"] + 1e-9 or mm["fast"] < 0.02, mm


def test_gemm_fp8_lo_pass(m5lib, bare_ctx):
    """mixed8: C = A_hi W^T (fp16 UMMA) + A_lo8 W8^T (kind::f8f6f4 UMMA, e5m2 x e4m3, scales 2^-2 x 2^+2) in ONE TMEM
    accumulator.  Against the fp32 product of the unrounded A the result must be an order of magnitude closer than the
    fp16-only product, and close to the fp16-pair product it replaces."""
    M, N, K = 40960, 3072, 1024
    g = torch.Generator(device=DEV).manual_seed(71)
    A32 = torch.randn(M, K, device=DEV, generator=g)
    W = (torch.randn(N, K, device=DEV, generator=g) * 0.03).half()
    Ah = A32.half()
    lo = A32 - Ah.float()
    A8 = (lo * 0.25).to(torch.float8_e5m2).view(torch.uint8).contiguous()
    W8 = (W.float() * 4.0).to(torch.float8_e4m3fn).view(torch.uint8).contiguous()
    out = torch.zeros(M, N, device=DEV)
    rc = m5lib.m5_dbg_gemm_f8lo(bare_ctx, ptr(Ah), K, ptr(A8), ptr(W), ptr(W8), M, N, K, ptr(out), N)
    capi.check(bare_ctx, rc, "gemm_f8lo")
    _sync(m5lib, bare_ctx)
    e8 = e16 = 0.0
    for r0 in range(0, M, 8192):
        ref = A32[r0:r0 + 8192] @ W.float().T
        e8 = max(e8, (out[r0:r0 + 8192] - ref).abs().max().item())
        e16 = max(e16, (Ah[r0:r0 + 8192].float() @ W.float().T - ref).abs().max().item())
    print(f"fp8 lo pass: max-abs {e8:.2e} vs fp32 (fp16-only operands: {e16:.2e})")
    assert e8 < 0.2 * e16, (e8, e16)

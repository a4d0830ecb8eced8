"""CPU suite, build container only: the oracle (oracle/*.py, the checker of every GPU parity test) against the UNMODIFIED
reference run LIVE (/root/reference/mars5/*) on freshly drawn inputs -- other seeds, other lengths and the edge cases (one text
token, one reference frame, one decoder row, t = 0, no guidance, typical-p, tiny penalty windows) than the committed fixtures
of tests/golden/ hold.  Skipped where the reference tree does not exist (the GPU box): nothing there may read /root/reference."""
import os
import sys

import numpy as np
import pytest
import torch

from mars5_tts_b200 import synth, weights
from oracle import ar_oracle, nar_oracle

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "mars5")),
                                reason="the unmodified reference tree exists only in the build container")
torch.set_grad_enabled(False)


@pytest.fixture(scope="module")
def ref():
    sys.path.insert(0, REF)
    try:
        from mars5 import ar_generate as ref_ar
        from mars5 import diffuser as ref_diff
        from mars5 import samplers as ref_smp
        from mars5.model import CodecLM, ResidualTransformer
    finally:
        sys.path.remove(REF)
    size = synth.TINY
    ar_sd, nar_sd = synth.make_ar_state(size), synth.make_nar_state(size)
    V = size["n_text"] + size["n_speech"]
    lm = CodecLM(n_vocab=V, dim=size["ar_dim"], nhead=size["ar_dim"] // 64, n_layers=size["ar_layers"],
                 n_spk_layers=size["ar_spk_layers"], dim_ff_scale=7 / 3).eval()
    nar = ResidualTransformer(n_text_vocab=size["n_text"] + 1, n_quant=1025, dim=size["nar_dim"], nhead=size["nar_dim"] // 64,
                              enc_layers=size["nar_enc_layers"], dec_layers=size["nar_dec_layers"],
                              n_spk_layers=size["nar_spk_layers"], t_emb_dim=size["nar_dim"], p_cond_drop=0, dropout=0).eval()
    lm.load_state_dict(ar_sd, strict=True)
    nar.load_state_dict(nar_sd, strict=True)
    cfg = weights.dims_from_state(ar_sd, nar_sd, None, size["n_text"])
    tt, st = synth.ByteTextTok(), synth.CodeSpeechTok()
    n_text = len(tt.vocab)
    return dict(ar=ref_ar, diff=ref_diff, smp=ref_smp, lm=lm, nar=nar, ar_sd=ar_sd, nar_sd=nar_sd, cfg=cfg, tt=tt, st=st,
                n_text=n_text, V=n_text + len(st.vocab), eos=n_text + st.special_tokens["<|endofspeech|>"])


# ------------------------------------------------------------------------------------------------ model evaluations
@pytest.mark.parametrize("P,Pf,seed", [(1, 1, 0), (2, 3, 1), (17, 1, 2), (33, 26, 3)])
def test_codeclm_forward_live(ref, P, Pf, seed):
    g = torch.Generator().manual_seed(1000 + seed)
    ids = torch.randint(0, ref["V"], (P,), generator=g)
    spk = torch.randint(0, 1024, (Pf, 8), generator=g)
    want = ref["lm"](ids[None], None, spk_reference=spk[None])[0]
    got = ar_oracle.codeclm_forward(ref["ar_sd"], ref["cfg"], ids, spk)
    assert got.shape == want.shape
    assert (got - want).abs().max() < 2e-5


@pytest.mark.parametrize("pad_at", [0, 4, 11])
def test_codeclm_forward_live_speaker_padding_quirk(ref, pad_at):
    """the AR speaker encoder masks every reference frame from the first one whose codebook-0 entry is the pad
    class 1024 (construct_padding_mask = cumsum of the equality), whatever comes after it."""
    g = torch.Generator().manual_seed(1500 + pad_at)
    ids = torch.randint(0, ref["V"], (9,), generator=g)
    spk = torch.randint(0, 1024, (12, 8), generator=g)
    spk[pad_at, 0] = 1024
    want = ref["lm"](ids[None], None, spk_reference=spk[None])[0]
    got = ar_oracle.codeclm_forward(ref["ar_sd"], ref["cfg"], ids, spk)
    assert (got - want).abs().max() < 2e-5
    spk2 = spk.clone()
    spk2[pad_at + 1:] = torch.randint(0, 1024, spk2[pad_at + 1:].shape, generator=g)   # frames behind the pad do not matter
    assert (ar_oracle.codeclm_forward(ref["ar_sd"], ref["cfg"], ids, spk2) - got).abs().max() == 0 or pad_at == 11


@pytest.mark.parametrize("Tc,Pf,S,t,seed", [(1, 1, 1, 0, 0), (6, 2, 5, 3, 1), (13, 9, 23, 9, 2), (3, 17, 2, 199, 3)])
@pytest.mark.parametrize("drop", [False, True])
def test_residual_transformer_forward_live(ref, Tc, Pf, S, t, seed, drop):
    g = torch.Generator().manual_seed(2000 + seed)
    c_text = torch.randint(0, ref["n_text"], (Tc,), generator=g)
    c_codes = torch.randint(0, 1024, (Pf, 8), generator=g)
    x = torch.randint(0, 1025, (S, 8), generator=g)
    want = ref["nar"](c_text[None], c_codes[None].clone(), torch.tensor([Tc]), torch.tensor([Pf]), x[None],
                      torch.zeros(1, S, dtype=torch.bool), torch.tensor([t]), drop_cond=drop).permute(0, 1, 3, 2)[0]
    got = nar_oracle.nar_forward(ref["nar_sd"], ref["cfg"], c_text, c_codes, x, t, drop)
    assert got.shape == want.shape
    assert (got - want).abs().max() < 5e-5


# ------------------------------------------------------------------------------------------------ AR loop
AR_CASES = [
    dict(temperature=0.7, top_k=200, top_p=0.2, typical_p=1.0, alpha_frequency=3, alpha_presence=0.4, penalty_window=80,
         eos_penalty_decay=0.5, eos_penalty_factor=1),
    dict(temperature=1.0, top_k=50, top_p=0.95, typical_p=1.0, alpha_frequency=1.5, alpha_presence=0.0, penalty_window=2,
         eos_penalty_decay=0.8, eos_penalty_factor=2),
    dict(temperature=1.3, top_k=1000, top_p=1.0, typical_p=0.6, alpha_frequency=0.0, alpha_presence=0.9, penalty_window=7,
         eos_penalty_decay=0.5, eos_penalty_factor=0),
]


@pytest.mark.parametrize("case", range(len(AR_CASES)))
@pytest.mark.parametrize("seed", [0, 1])
def test_ar_generate_live_token_exact(ref, case, seed):
    """ar_generate with torch.multinomial replaced by the exponential race on injected Exp(1) noise:
    the oracle's token sequence must equal the reference's, EOS handling and max_len included."""
    sc = AR_CASES[case]
    g = torch.Generator().manual_seed(3000 + 10 * case + seed)
    Pf, n_txt, n_sp = (5, 4, 2) if seed == 0 else (14, 11, 6)
    spk = torch.randint(0, 1024, (Pf, 8), generator=g)
    text_ids = [256] + torch.randint(0, 256, (n_txt,), generator=g).tolist() + [257]
    prompt = torch.tensor(text_ids + (torch.randint(0, 1024, (n_sp,), generator=g) + ref["n_text"]).tolist())
    steps = 10
    max_len = len(prompt) + steps
    noise = torch.empty(steps, ref["V"]).exponential_(1, generator=g)
    n_ph = 3 + 4 * seed
    calls = {"n": 0}
    real = torch.multinomial

    def fake_multinomial(p, num_samples, replacement=False):
        q = p / noise[calls["n"]]
        calls["n"] += 1
        return q.argmax(dim=-1, keepdim=True)

    torch.multinomial = fake_multinomial
    try:
        want = ref["ar"].ar_generate(ref["tt"], ref["st"], ref["lm"], prompt, spk, len(text_ids) + 1, max_len=max_len, fp16=False,
                                     temperature=sc["temperature"], topk=sc["top_k"], top_p=sc["top_p"], typical_p=sc["typical_p"],
                                     alpha_frequency=sc["alpha_frequency"], alpha_presence=sc["alpha_presence"],
                                     penalty_window=sc["penalty_window"], eos_penalty_decay=sc["eos_penalty_decay"],
                                     eos_penalty_factor=sc["eos_penalty_factor"], n_phones_gen=n_ph, vocode=False, use_kv_cache=True)
    finally:
        torch.multinomial = real
    got, _ = ar_oracle.ar_generate(ref["ar_sd"], ref["cfg"], prompt, spk, sc, noise, max_len, n_ph, ref["eos"])
    np.testing.assert_array_equal(got.numpy(), want.numpy())


# ------------------------------------------------------------------------------------------------ NAR loop
@pytest.mark.parametrize("deep,w,q0,T,seed", [(True, 3.0, 2, 5, 0), (False, 3.0, 0, 4, 1), (True, 1.0, 1, 3, 2), (False, 2.0, 20, 6, 3)])
def test_perform_simple_inference_live_code_exact(ref, deep, w, q0, T, seed):
    """perform_simple_inference with torch.randint / torch.rand_like replaced by injected draws: the oracle's codes equal the
    reference's -- with and without classifier-free guidance (guidance_w = 1 skips the unconditional pass,
    diffuser.py:361), q0 override steps beyond T, deep and shallow clone."""
    g = torch.Generator().manual_seed(4000 + seed)
    Pf, Tc, N = 3 + 4 * seed, 2 + 3 * seed, 1 + 3 * seed
    c_text = torch.randint(0, ref["n_text"], (Tc,), generator=g)
    c_codes = torch.randint(0, 1024, (Pf, 8), generator=g)
    x_l0 = torch.randint(0, 1024, (N,), generator=g)
    x_init = torch.randint(0, 1025, (N, 8), generator=g)
    S_tot = N + (Pf if deep else 0)
    u = torch.rand(T, 2, S_tot, 8, 1025, generator=g)
    st_ = {"i": 0}
    real_randint, real_rand_like = torch.randint, torch.rand_like

    def fake_randint(lo, hi, shape, **kw):
        return x_init[None].clone()

    def fake_rand_like(t_, **kw):
        step, draw = divmod(st_["i"], 2)
        st_["i"] += 1
        return u[step, draw][None].clone()

    rd = ref["diff"]
    torch.randint, torch.rand_like = fake_randint, fake_rand_like
    try:
        diff = rd.MultinomialDiffusion(1025, timesteps=T)
        dsh = rd.DSH(last_greedy=True, x_0_temp=0.7, guidance_w=w, deep_clone=deep, jump_len=1, jump_n_sample=1,
                     q0_override_steps=q0, enable_kevin_scaled_inference=True, progress=False)
        _x = x_l0[None, :, None].repeat(1, 1, 8)
        want = rd.perform_simple_inference(ref["nar"], (c_text[None], c_codes[None].clone(), torch.tensor([Tc]), torch.tensor([Pf]),
                                                        _x, torch.zeros(1, N, dtype=torch.bool)),
                                           diff, T, torch.float16, dsh=dsh, retain_quant0=True)[0]
    finally:
        torch.randint, torch.rand_like = real_randint, real_rand_like
    assert st_["i"] == 2 * T - 1
    ncfg = dict(T=T, deep_clone=deep, guidance_w=w, x0_temp=0.7, q0_override_steps=q0)
    got = nar_oracle.nar_infer(ref["nar_sd"], ref["cfg"], c_text, c_codes, x_l0, ncfg, x_init, u)
    np.testing.assert_array_equal(got.numpy(), want.numpy())


# ------------------------------------------------------------------------------------------------ small pieces
@pytest.mark.parametrize("T,jl,jn", [(200, 1, 1), (6, 2, 2), (10, 3, 2), (12, 5, 3), (7, 10, 10)])
def test_get_schedule_live(ref, T, jl, jn):
    assert nar_oracle.get_schedule(T, jl, jn) == list(ref["diff"].get_schedule(T, jump_len=jl, jump_n_sample=jn))


@pytest.mark.parametrize("T", [3, 64, 128, 256])
def test_diffusion_tables_live(ref, T):
    d = ref["diff"].MultinomialDiffusion(1025, timesteps=T)
    want = torch.stack([d.log_alpha, d.log_1_min_alpha, d.log_cumprod_alpha, d.log_1_min_cumprod_alpha]).numpy()
    np.testing.assert_array_equal(torch.stack(nar_oracle.diffusion_tables(T)).numpy(), want)
    np.testing.assert_array_equal(weights.diffusion_schedule(T).numpy(), want)


# ------------------------------------------------------------------------------------------------ tokenisers
@pytest.mark.parametrize("seed", [7, 8])
def test_tokenisers_live(seed, tmp_path):
    """Fresh minbpe-v1 models trained, saved and re-loaded by the reference's own classes on another seeded corpus; the native
    merge engine behind mars5_tts_b200.bpe must encode / decode exactly like them (integer work)."""
    import random
    sys.path.insert(0, REF)
    try:
        from mars5.minbpe.codebook import CodebookTokenizer as RefCB
        from mars5.minbpe.regex import GPT4_SPLIT_PATTERN, RegexTokenizer as RefRT
    finally:
        sys.path.remove(REF)
    from mars5_tts_b200 import bpe
    rng = random.Random(seed)
    syl = ["ka", "to", "mi", "ra", "sen", "lo", "vi", "the", "ing", "qu", "é", "ü", "ñ", "漢", "字", "🙂", "'s", "12", "0"]
    sep = [" ", " ", ", ", ". ", "! ", "?\n", "\n\n", "  ", "\t", " - ", "'ll ", ""]

    def text(n):
        return "".join("".join(rng.choice(syl) for _ in range(rng.randint(1, 4))) + rng.choice(sep) for _ in range(n))

    # ---- text tokeniser
    rt = RefRT()
    rt.train(text(1500), 256 + 150)
    rt.register_special_tokens({"<|startoftext|>": 406, "<|endoftext|>": 407})
    rt.save(str(tmp_path / "text"))
    ref_t, my_t = RefRT(), bpe.RegexTokenizer()
    ref_t.load(str(tmp_path / "text.model"))
    my_t.load(str(tmp_path / "text.model"))
    assert len(my_t.vocab) == len(ref_t.vocab) and my_t.special_tokens == ref_t.special_tokens
    samples = [text(rng.randint(1, 40)) for _ in range(25)] + ["", " ", "\n", "a", "<|startoftext|>hello<|endoftext|>", "   x   "]
    for s in samples:
        ids = ref_t.encode(s, allowed_special="all")
        assert my_t.encode(s, allowed_special="all") == ids, repr(s[:40])
        assert my_t.encode_ordinary(s) == ref_t.encode_ordinary(s)
        assert my_t.decode(ids) == ref_t.decode(ids)
    plain = [s for s in samples if "<|" not in s]
    assert my_t.encode_batch(plain) == [ref_t.encode_ordinary(s) for s in plain]
    # ---- speech tokeniser
    hot = [rng.randrange(1024) for _ in range(30)]
    codes = lambda n: [rng.choice(hot) if rng.random() < 0.8 else rng.randrange(1024) for _ in range(n)]
    cb = RefCB(GPT4_SPLIT_PATTERN)
    cb.train(" ".join(map(str, codes(6000))), 1024 + 200)
    cb.register_special_tokens({"<|endofspeech|>": 1224})
    cb.save(str(tmp_path / "speech"))
    ref_s, my_s = RefCB(GPT4_SPLIT_PATTERN), bpe.CodebookTokenizer(bpe.GPT4_SPLIT_PATTERN)
    ref_s.load(str(tmp_path / "speech.model"))
    my_s.load(str(tmp_path / "speech.model"))
    seqs = [codes(n) for n in (1, 2, 5, 31, 450, 1499)] + [[hot[0]] * 11 + [hot[1]] * 6]
    for c in seqs:
        txt = " ".join(map(str, c))
        ids = ref_s.encode(txt)
        assert my_s.encode(txt) == ids
        assert my_s.decode_int(ids) == ref_s.decode_int(ids)
        assert my_s.decode(ids) == ref_s.decode(ids)
        mixed = ids[: len(ids) // 2] + [1224] + ids[len(ids) // 2:]
        assert my_s.decode_int(mixed) == ref_s.decode_int(mixed)
    assert my_s.encode_codes_batch(seqs) == [ref_s.encode(" ".join(map(str, c))) for c in seqs]


# ------------------------------------------------------------------------------------------------ silence trim
def test_trim_bounds_live():
    """m5_trim_bounds (csrc/trim.cu host code, behind mars5_tts_b200.trim) against the reference's trim() on freshly drawn
    waveforms: bursts at random places and levels, several top_db values, the shortest legal clip.  mars5/trim.py itself does not
    run under numpy 2 (np.array(x, copy=False), trim.py:546): as in tests/golden/make_trim_golden.py the untouched module is handed
    a numpy proxy with the numpy-1 meaning of that one call."""
    sys.path.insert(0, REF)
    try:
        import mars5.trim as ref_trim
    finally:
        sys.path.remove(REF)
    from mars5_tts_b200.trim import trim_bounds_batch

    class Numpy1:
        def __getattr__(self, name):
            return getattr(np, name)

        @staticmethod
        def array(x, copy=True, subok=False, **kw):
            return np.asarray(x) if copy is False else np.array(x, copy=copy, subok=subok, **kw)

    real_np, ref_trim.np = ref_trim.np, Numpy1()
    try:
        g = torch.Generator().manual_seed(77)
        wavs, want = [], []
        for i in range(24):
            n = int(torch.randint(1025, 60000, (1,), generator=g)) if i else 1025
            y = torch.randn(n, generator=g) * (10.0 ** float(-torch.rand(1, generator=g) * 5))          # noise floor 1 .. 1e-5
            for _ in range(int(torch.randint(0, 4, (1,), generator=g))):
                a = int(torch.randint(0, n, (1,), generator=g))
                b = min(n, a + int(torch.randint(1, 20000, (1,), generator=g)))
                y[a:b] += torch.randn(b - a, generator=g) * float(torch.rand(1, generator=g))
            top_db = [27, 60, 10, 40][i % 4]
            _, idx = ref_trim.trim(y, top_db=top_db)
            wavs.append((y.numpy(), top_db))
            want.append((int(idx[0]), int(idx[1])))
    finally:
        ref_trim.np = real_np
    for top_db in (27, 60, 10, 40):
        sel = [i for i, (_, d) in enumerate(wavs) if d == top_db]
        got = trim_bounds_batch([wavs[i][0] for i in sel], top_db=top_db)
        assert got == [want[i] for i in sel], top_db

"""Generates tests/golden/*.npz by running the UNMODIFIED reference (/root/reference/mars5/*) on CPU fp32.

Run in the build container only (the GPU box has no /root/reference):   python tests/golden/make_golden.py
Inputs are the seeded synthetic, reference-format checkpoints of mars5_tts_b200.synth (TINY size) loaded with
``load_state_dict(strict=True)`` into the reference's own CodecLM / ResidualTransformer, so the fixtures pin

  * the reference's state-dict key set / shapes            (weights.repack consumes exactly these)
  * CodecLM.forward logits, ar_generate token sequence with injected Exp(1) noise (torch.multinomial patched)
  * ResidualTransformer.forward logits (cond / drop_cond), perform_simple_inference codes with injected
    randint / rand_like draws
  * MultinomialDiffusion tables (T = 10, 200), reverse_diffusion posterior on random logits
  * the logit warpers (samplers.py) on random logits

The vocoder has no fixture: `vocos` is not installed here (see oracle/vocos_oracle.py, "parity unpinned").
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from mars5 import ar_generate as ref_ar  # noqa: E402
from mars5 import diffuser as ref_diff  # noqa: E402
from mars5.model import CodecLM, ResidualTransformer  # noqa: E402
from mars5.samplers import early_eos_penalty, freq_rep_penalty, top_k_top_p_filtering  # noqa: E402

from mars5_tts_b200 import synth  # noqa: E402

torch.set_grad_enabled(False)
torch.set_num_threads(8)


def build_reference_models(size):
    ar_sd, nar_sd = synth.make_ar_state(size), synth.make_nar_state(size)
    V = size["n_text"] + size["n_speech"]
    lm = CodecLM(n_vocab=V, dim=size["ar_dim"], nhead=size["ar_dim"] // 64, n_layers=size["ar_layers"],
                 n_spk_layers=size["ar_spk_layers"], dim_ff_scale=7 / 3).eval()
    nar = ResidualTransformer(n_text_vocab=size["n_text"] + 1, n_quant=1025, dim=size["nar_dim"], nhead=size["nar_dim"] // 64,
                              enc_layers=size["nar_enc_layers"], dec_layers=size["nar_dec_layers"],
                              n_spk_layers=size["nar_spk_layers"], t_emb_dim=size["nar_dim"], p_cond_drop=0, dropout=0).eval()
    for name, mod, sd in (("ar", lm, ar_sd), ("nar", nar, nar_sd)):
        ref_keys = {k: tuple(v.shape) for k, v in mod.state_dict().items()}
        my_keys = {k: tuple(v.shape) for k, v in sd.items()}
        assert ref_keys == my_keys, (name, set(ref_keys) ^ set(my_keys))
        mod.load_state_dict(sd, strict=True)
    return lm, nar, ar_sd, nar_sd


def main():
    size = synth.TINY
    lm, nar, ar_sd, nar_sd = build_reference_models(size)
    tt, st = synth.ByteTextTok(), synth.CodeSpeechTok()
    n_text = len(tt.vocab)
    V = n_text + len(st.vocab)
    eos = n_text + st.special_tokens["<|endofspeech|>"]
    g = torch.Generator().manual_seed(1234)
    out = {}

    # ---------------- AR: forward logits + generation with injected noise
    Pf, n_txt, n_sp = 12, 9, 5
    spk = torch.randint(0, 1024, (Pf, 8), generator=g)
    text_ids = [256] + torch.randint(0, 256, (n_txt,), generator=g).tolist() + [257]
    speech_ids = (torch.randint(0, 1024, (n_sp,), generator=g) + n_text).tolist()
    prompt = torch.tensor(text_ids + speech_ids)
    logits = lm(prompt[None], None, spk_reference=spk[None])[0]
    out["ar_spk"], out["ar_prompt"], out["ar_logits"] = spk.numpy(), prompt.numpy(), logits.numpy()

    max_len, steps = len(prompt) + 14, 14
    noise = torch.empty(steps, V).exponential_(1, generator=g)
    calls = {"n": 0}
    real_multinomial = torch.multinomial

    def fake_multinomial(p, num_samples, replacement=False):
        q = p / noise[calls["n"]]
        calls["n"] += 1
        return q.argmax(dim=-1, keepdim=True)

    torch.multinomial = fake_multinomial
    try:
        for use_cache in (True, False):
            calls["n"] = 0
            seq = ref_ar.ar_generate(tt, st, lm, prompt, spk, len(text_ids) + 1, max_len=max_len, fp16=False, temperature=0.7,
                                     topk=200, top_p=0.2, alpha_frequency=3, alpha_presence=0.4, penalty_window=80,
                                     eos_penalty_decay=0.5, eos_penalty_factor=1, n_phones_gen=7, vocode=False,
                                     use_kv_cache=use_cache)
            out["ar_gen_cache" if use_cache else "ar_gen_nocache"] = seq.numpy()
        # a run with a wide nucleus so that sampling actually depends on the noise
        calls["n"] = 0
        seq = ref_ar.ar_generate(tt, st, lm, prompt, spk, len(text_ids) + 1, max_len=max_len, fp16=False, temperature=1.0,
                                 topk=50, top_p=0.95, alpha_frequency=3, alpha_presence=0.4, penalty_window=4,
                                 eos_penalty_decay=0.5, eos_penalty_factor=1, n_phones_gen=7, vocode=False, use_kv_cache=True)
        out["ar_gen_wide"] = seq.numpy()
    finally:
        torch.multinomial = real_multinomial
    out["ar_noise"], out["ar_max_len"] = noise.numpy(), np.int64(max_len)

    # ---------------- samplers on random logits
    lg = torch.randn(3, V, generator=g) * 3
    prev = torch.randint(n_text, V, (3, 30), generator=g)
    a = freq_rep_penalty(lg.clone(), previous=prev, alpha_frequency=3, alpha_presence=0.4, penalty_window=20)
    b = early_eos_penalty(a.clone(), 5, 9, 0.5, 1, eos_index=eos)
    c = top_k_top_p_filtering(b.clone() / 0.7, top_k=200, top_p=0.2)
    d = top_k_top_p_filtering(b.clone() / 0.9, top_k=40, top_p=0.9)
    out.update(smp_logits=lg.numpy(), smp_prev=prev.numpy(), smp_rep=a.numpy(), smp_eos=b.numpy(), smp_kp=c.numpy(),
               smp_kp2=d.numpy())

    # ---------------- NAR: forward logits
    Tc, S = 11, 19
    c_text = torch.randint(0, n_text, (Tc,), generator=g)
    c_codes = torch.randint(0, 1024, (Pf, 8), generator=g)
    x = torch.randint(0, 1025, (S, 8), generator=g)
    args = lambda: (c_text[None], c_codes[None].clone(), torch.tensor([Tc]), torch.tensor([Pf]), x[None],
                    torch.zeros(1, S, dtype=torch.bool), torch.tensor([7]))
    lg_c = nar(*args()).permute(0, 1, 3, 2)[0]
    lg_u = nar(*args(), drop_cond=True).permute(0, 1, 3, 2)[0]
    out.update(nar_c_text=c_text.numpy(), nar_c_codes=c_codes.numpy(), nar_x=x.numpy(), nar_t=np.int64(7),
               nar_logits_cond=lg_c.numpy(), nar_logits_uncond=lg_u.numpy())

    # ---------------- diffusion tables
    for T in (10, 200):
        d_ = ref_diff.MultinomialDiffusion(1025, timesteps=T)
        out[f"diff_tables_{T}"] = torch.stack([d_.log_alpha, d_.log_1_min_alpha, d_.log_cumprod_alpha,
                                               d_.log_1_min_cumprod_alpha]).numpy()

    # ---------------- full NAR loop with injected randomness (deep and shallow clone)
    T = 6
    N = 7
    x_l0 = torch.randint(0, 1024, (N,), generator=g)
    x_init = torch.randint(0, 1025, (N, 8), generator=g)
    real_randint, real_rand_like = torch.randint, torch.rand_like
    for deep in (True, False):
        S_tot = N + (Pf if deep else 0)
        u = torch.rand(T, 2, S_tot, 8, 1025, generator=g)
        st_ = {"i": 0}

        def fake_randint(lo, hi, shape, **kw):
            return x_init[None].clone()

        def fake_rand_like(t_, **kw):
            step, draw = divmod(st_["i"], 2)
            st_["i"] += 1
            return u[step, draw][None].clone()

        torch.randint, torch.rand_like = fake_randint, fake_rand_like
        try:
            diff = ref_diff.MultinomialDiffusion(1025, timesteps=T)
            dsh = ref_diff.DSH(last_greedy=True, x_0_temp=0.7, guidance_w=3, deep_clone=deep, jump_len=1, jump_n_sample=1,
                               q0_override_steps=2, enable_kevin_scaled_inference=True, progress=False)
            _x = x_l0[None, :, None].repeat(1, 1, 8)
            res = ref_diff.perform_simple_inference(nar, (c_text[None], c_codes[None].clone(), torch.tensor([Tc]),
                                                          torch.tensor([Pf]), _x, torch.zeros(1, N, dtype=torch.bool)),
                                                    diff, T, torch.float16, dsh=dsh, retain_quant0=True)
        finally:
            torch.randint, torch.rand_like = real_randint, real_rand_like
        # the last step (t = 0) draws only once -> 2*T - 1 calls
        assert st_["i"] == 2 * T - 1, st_["i"]
        tag = "deep" if deep else "shallow"
        out[f"nar_loop_{tag}_u"] = u.numpy().astype(np.float32)
        out[f"nar_loop_{tag}_codes"] = res[0].numpy()
    out.update(nar_loop_x_l0=x_l0.numpy(), nar_loop_x_init=x_init.numpy(), nar_loop_T=np.int64(T))

    # ---------------- one posterior step on random logits (reverse_diffusion with a stub model)
    Sx = 9
    cond = torch.randn(Sx, 8, 1025, generator=g) * 2
    uncond = torch.randn(Sx, 8, 1025, generator=g) * 2
    xt = torch.randint(0, 1025, (Sx, 8), generator=g)
    xk = torch.randint(0, 1024, (Sx, 8), generator=g)
    m = torch.rand(Sx, 8, generator=g) < 0.4
    u2 = torch.rand(2, Sx, 8, 1025, generator=g)

    class Stub:
        def __call__(self, *a, drop_cond=False):
            return (uncond if drop_cond else cond).permute(0, 2, 1)[None]

    for t_ in (0, 3, 199):
        k = {"i": 0}

        def fake_rand_like2(t__, **kw):
            k["i"] += 1
            return u2[k["i"] - 1][None].clone()

        torch.rand_like = fake_rand_like2
        try:
            diff = ref_diff.MultinomialDiffusion(1025, timesteps=200)
            dsh = ref_diff.DSH(x_0_temp=0.7, guidance_w=3)
            batch = (None, None, None, None, xt[None], None, torch.tensor([t_]))
            xo, _ = ref_diff.reverse_diffusion(diff, Stub(), batch, xk[None], m[None], temperature=0.7,
                                               alphas=torch.linspace(1, 0, 200), ensemble_size=1, dsh=dsh)
        finally:
            torch.rand_like = real_rand_like
        out[f"post_out_t{t_}"] = xo[0].numpy()
    out.update(post_cond=cond.numpy(), post_uncond=uncond.numpy(), post_xt=xt.numpy(), post_xk=xk.numpy(),
               post_m=m.numpy(), post_u=u2.numpy())

    # Large random INPUT tensors are not stored: tests regenerate them from the same seeded generator
    # (tests/golden/inputs.py replays this script's draw order) and verify these checksums first.
    big = [k for k in out if k.endswith("_u") or k in ("ar_noise", "post_cond", "post_uncond")]
    for k in big:
        out["chk_" + k] = np.float64(np.asarray(out[k], dtype=np.float64).sum())
        del out[k]
    path = os.path.join(HERE, "reference_tiny.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB;", len(out), "arrays")


if __name__ == "__main__":
    main()

"""CPU suite, build container only: the drop-in surface END TO END against the unmodified reference run live.

`/root/reference/inference.py` (Mars5TTS.tts, lines 201-307) is imported with stub modules standing in for the packages that
are not installable here (vocos, encodec, librosa) and driven with tiny seeded models, freshly trained minbpe tokenisers (speech
tokens that expand to SEVERAL Encodec codes), a stub codec / vocoder and injected randomness (torch.multinomial, randint,
rand_like replaced as in SURVEY Appendix D).  Our Mars5TTS.tts_batch runs the SAME inputs through its real host glue
(engine.py: _prepare, prompt assembly, first_codec_idx, speech-BPE decode, both crops, cfg -> C structs) with a stand-in engine
whose compute stages are the CPU oracle -- the CUDA stages themselves are pinned to that oracle by the -m gpu suite.  The AR L0
codes and the final 8-codebook codes (through a deterministic stub vocoder) must be identical, deep and shallow clone."""
import os
import random
import sys
import types

import numpy as np
import pytest
import torch

from mars5_tts_b200 import bpe, synth, weights
from mars5_tts_b200.engine import Engine, InferenceConfig, Mars5TTS
from oracle import ar_oracle, nar_oracle

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "inference.py")),
                                reason="the unmodified reference tree exists only in the build container")
torch.set_grad_enabled(False)
T_STEPS, AR_STEPS, HOP = 3, 9, 320
_REAL_RANDINT = torch.randint   # the reference run patches torch.randint / rand_like / multinomial


def _stub_wave(codes):
    """deterministic stand-in vocoder: codes (N, 8) -> (HOP * N,) float32"""
    w = (np.asarray(codes, dtype=np.float32) * np.arange(1, 9, dtype=np.float32)).sum(1)
    return np.repeat(w, HOP)


class StubCodec:
    """EncodecModel.encode stand-in: [(codes (1, 8, T), scale)], T = ceil(samples / 320), seeded by the clip length"""

    def encode(self, wav):
        T = (wav.shape[-1] + HOP - 1) // HOP
        g = torch.Generator().manual_seed(int(wav.shape[-1]))
        return [(_REAL_RANDINT(0, 1024, (1, 8, T), generator=g), None)]


class StubVocos:
    def codes_to_features(self, tokens):            # (n_q, L)
        return tokens

    def decode(self, features, bandwidth_id=None):
        assert int(bandwidth_id[0]) == 1            # inference.py:169
        return torch.from_numpy(_stub_wave(features.T.cpu().numpy()))[None]


def _noise(kind, i, shape):
    g = torch.Generator().manual_seed({"ar": 5000, "init": 6000, "u": 7000}[kind] + i)
    if kind == "ar":
        return torch.empty(shape).exponential_(1, generator=g)
    if kind == "init":
        return _REAL_RANDINT(0, 1025, shape, generator=g)
    return torch.rand(shape, generator=g)


class OracleEngine:
    """The Engine surface tts_batch uses, with the CPU oracle as the compute stages and the injected noise of the reference run."""
    has_encodec = False
    dims = {"voc_hop": HOP}

    def __init__(self, ar_sd, nar_sd, cfg):
        self.ar_sd, self.nar_sd, self.cfg, self._sched_cache = ar_sd, nar_sd, cfg, {}

    make_ar_cfg, make_nar_cfg, schedule = Engine.make_ar_cfg, Engine.make_nar_cfg, Engine.schedule   # the real C-struct builders

    def ar_generate(self, prompts, spks, n_phones, acfg, seed=0, utt_ids=None):
        sc = dict(temperature=acfg.temperature, top_k=acfg.top_k, top_p=acfg.top_p, typical_p=acfg.typical_p,
                  alpha_frequency=acfg.alpha_frequency, alpha_presence=acfg.alpha_presence, penalty_window=acfg.penalty_window,
                  eos_penalty_decay=acfg.eos_penalty_decay, eos_penalty_factor=acfg.eos_penalty_factor)
        ids, hits = [], []
        for p, s, n in zip(prompts, spks, n_phones):
            noise = torch.stack([_noise("ar", i, (self.cfg["ar_vocab"],)) for i in range(acfg.max_len)])
            seq, hit = ar_oracle.ar_generate(self.ar_sd, self.cfg, torch.tensor(p), torch.from_numpy(np.asarray(s)).long(), sc, noise,
                                             acfg.max_len, n, acfg.eos_id)
            ids.append(seq.numpy()); hits.append(hit)
        return ids, hits, None

    def nar_infer(self, texts, spks, l0s, ncfg, seed=0, utt_ids=None):
        out = []
        for t, s, l0 in zip(texts, spks, l0s):
            N, S = len(l0), len(l0) + (len(s) if ncfg.deep_clone else 0)
            x_init = _noise("init", 0, (N, 8))
            u = torch.stack([torch.stack([_noise("u", 2 * st + d, (S, 8, 1025)) for d in range(2)]) for st in range(ncfg.T)])
            nc = dict(T=ncfg.T, deep_clone=bool(ncfg.deep_clone), guidance_w=ncfg.guidance_w, x0_temp=ncfg.x0_temp,
                      q0_override_steps=ncfg.q0_override_steps)
            out.append(nar_oracle.nar_infer(self.nar_sd, self.cfg, torch.tensor(t), torch.from_numpy(np.asarray(s)).long(),
                                            torch.from_numpy(np.asarray(l0)).long(), nc, x_init, u).numpy())
        return out

    def vocode_trim(self, outs, top_db, bandwidth_id=1):
        assert bandwidth_id == 1
        wavs = [_stub_wave(o) for o in outs]
        return wavs, [(0, len(w)) for w in wavs]


@pytest.fixture(scope="module")
def world(tmp_path_factory):
    return build_world(tmp_path_factory.mktemp("tok"))


def build_world(tmp):
    stubbed = []
    for name, attrs in (("vocos", {"Vocos": object}), ("encodec", {"EncodecModel": object}), ("librosa", {})):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__dict__.update(attrs)
            sys.modules[name] = m
            stubbed.append(name)
    sys.path.insert(0, REF)
    try:
        import inference as ref_inf
        from mars5.minbpe.codebook import CodebookTokenizer as RefCB
        from mars5.minbpe.regex import GPT4_SPLIT_PATTERN, RegexTokenizer as RefRT
    finally:
        sys.path.remove(REF)
        for name in stubbed:
            del sys.modules[name]
    rng = random.Random(11)
    syl = ["the", "qui", "ck", "bro", "wn", "rat", "we", "meet", "de", "mand", "ac", "tu", "al", "ly", "ha", "ven't", "ma", "na", "ged",
           "to", "é", "so", "on", "ing", "st", "ou", "er", "12", "ü"]
    corpus = " ".join("".join(rng.choice(syl) for _ in range(rng.randint(1, 4))) for _ in range(1500))
    rt = RefRT()
    rt.train(corpus, 256 + 60)
    rt.register_special_tokens({"<|startoftext|>": 316, "<|endoftext|>": 317})
    rt.save(str(tmp / "text"))
    hot = [rng.randrange(1024) for _ in range(12)]
    cb = RefCB(GPT4_SPLIT_PATTERN)
    cb.train(" ".join(str(rng.choice(hot) if rng.random() < 0.9 else rng.randrange(1024)) for _ in range(5000)), 1024 + 80)
    cb.register_special_tokens({"<|endofspeech|>": 1104})
    cb.save(str(tmp / "speech"))
    ref_t, ref_s = RefRT(), RefCB(GPT4_SPLIT_PATTERN)
    ref_t.load(str(tmp / "text.model")); ref_s.load(str(tmp / "speech.model"))
    my_t, my_s = bpe.RegexTokenizer(bpe.GPT4_SPLIT_PATTERN), bpe.CodebookTokenizer(bpe.GPT4_SPLIT_PATTERN)
    my_t.load(str(tmp / "text.model")); my_s.load(str(tmp / "speech.model"))
    size = dict(synth.TINY, n_text=len(ref_t.vocab), n_speech=len(ref_s.vocab))
    ar_sd, nar_sd = synth.make_ar_state(size), synth.make_nar_state(size)
    from mars5.model import CodecLM, ResidualTransformer   # already imported by inference.py
    lm = CodecLM(n_vocab=size["n_text"] + size["n_speech"], dim=size["ar_dim"], nhead=size["ar_dim"] // 64, n_layers=size["ar_layers"],
                 n_spk_layers=size["ar_spk_layers"], dim_ff_scale=7 / 3).eval()
    nar = ResidualTransformer(n_text_vocab=size["n_text"] + 1, n_quant=1025, dim=size["nar_dim"], nhead=size["nar_dim"] // 64,
                              enc_layers=size["nar_enc_layers"], dec_layers=size["nar_dec_layers"],
                              n_spk_layers=size["nar_spk_layers"], t_emb_dim=size["nar_dim"], p_cond_drop=0, dropout=0).eval()
    lm.load_state_dict(ar_sd, strict=True); nar.load_state_dict(nar_sd, strict=True)
    cfg = weights.dims_from_state(ar_sd, nar_sd, None, size["n_text"])
    # the reference object, without its __init__ (which downloads Encodec / Vocos)
    r = object.__new__(ref_inf.Mars5TTS)
    torch.nn.Module.__init__(r)
    r.device, r.sr, r.texttok, r.speechtok, r.codec, r.vocos = torch.device("cpu"), 24000, ref_t, ref_s, StubCodec(), StubVocos()
    r.codeclm, r.codecnar, r.default_T, r.diffusion_n_classes = lm, nar, T_STEPS, 1025
    ref_inf.trim = lambda wav, top_db=27: (wav, None)      # the silence trim has its own fixtures (tests/test_trim_cpu.py)
    # ours, without its __init__ (which creates the CUDA engine)
    o = object.__new__(Mars5TTS)
    o.texttok, o.speechtok, o.codec, o.engine = my_t, my_s, StubCodec(), OracleEngine(ar_sd, nar_sd, cfg)
    o.device, o.default_T, o.sr, o.latent_sr, o._calls = torch.device("cpu"), T_STEPS, 24000, 75, 0
    return ref_inf, r, o, size


def _run_reference(ref_inf, r, text, ref_audio, transcript, cfg):
    calls = {"ar": 0, "u": 0}
    real = torch.multinomial, torch.randint, torch.rand_like

    def fake_multinomial(p, num_samples, replacement=False):
        q = p / _noise("ar", calls["ar"], (p.shape[-1],))
        calls["ar"] += 1
        return q.argmax(dim=-1, keepdim=True)

    def fake_randint(lo, hi, shape, **kw):
        return _noise("init", 0, tuple(shape[1:]))[None]

    def fake_rand_like(t, **kw):
        calls["u"] += 1
        return _noise("u", calls["u"] - 1, tuple(t.shape[1:]))[None]

    torch.multinomial, torch.randint, torch.rand_like = fake_multinomial, fake_randint, fake_rand_like
    try:
        return r.tts(text, ref_audio, transcript, cfg)
    finally:
        torch.multinomial, torch.randint, torch.rand_like = real


@pytest.mark.parametrize("deep", [True, False])
@pytest.mark.parametrize("text,transcript,n_ref,channels,pad", [("the quick brown rat", "we actually haven't managed", 2300, 0, 0.0),
                                                                ("so on é", "to meet demand", 961, 0, 0.0),
                                                                ("  we meet  ", "the rat", 1500, 2, 0.03)])   # stereo clip, left zero-pad (B.4-17)
def test_tts_glue_equals_live_reference(world, deep, text, transcript, n_ref, channels, pad):
    ref_inf, r, o, size = world
    shape = (channels, n_ref) if channels else (n_ref,)
    ref_audio = torch.randn(*shape, generator=torch.Generator().manual_seed(n_ref)) * 0.1
    probe = o._prepare(text, ref_audio, transcript, InferenceConfig(deep_clone=deep, ref_audio_pad=pad))
    kw = dict(deep_clone=deep, ref_audio_pad=pad, generate_max_len_override=len(probe["prompt"]) + AR_STEPS, temperature=1.0, top_k=60,
              top_p=0.9, rep_penalty_window=5, q0_override_steps=2)
    want_codes, want_wav = _run_reference(ref_inf, r, text, ref_audio, transcript, ref_inf.InferenceConfig(**kw))
    got_codes, got_wav = o.tts_batch([text], [ref_audio], [transcript], InferenceConfig(**kw))[0]
    assert want_codes.numel() > 0
    np.testing.assert_array_equal(got_codes.numpy(), want_codes.numpy())
    np.testing.assert_array_equal(got_wav.numpy(), want_wav.numpy())

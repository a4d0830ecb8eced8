"""CPU suite: host-side logic that needs no GPU (weight repack layout, schedule tables, tokenizer stand-ins, bench workload)."""
import pytest
import numpy as np
import torch

from mars5_tts_b200 import synth, weights


def test_repack_layouts():
    size = synth.TINY
    ar, nar, voc = synth.make_ar_state(size), synth.make_nar_state(size), synth.make_vocos_state(size)
    d = weights.dims_from_state(ar, nar, voc, size["n_text"])
    t, alphas = weights.repack(ar, nar, voc, d, max_pos=64)
    D, F = d["ar_dim"], d["ar_hidden"]
    assert t["ar.l0.wqkv"].shape == (3 * D, D) and t["ar.l0.wqkv"].dtype == torch.float16
    assert torch.equal(t["ar.l0.wqkv"][D:2 * D].float(), ar["ar.layers.0.attention.wk.weight"])  # fp16-exact checkpoints
    w13 = t["ar.l0.w13"].float()
    assert torch.equal(w13[0::2], ar["ar.layers.0.feed_forward.w1.weight"]) and torch.equal(w13[1::2], ar["ar.layers.0.feed_forward.w3.weight"])
    Dn = d["nar_dim"]
    assert torch.equal(t["nar.dec.l1.ca_kv_w"].float(), nar["tfm.decoder.layers.1.multihead_attn.in_proj_weight"][Dn:])
    hw = voc["head.out.weight"]
    h3 = t["voc.head_w"].float()
    K = hw.shape[1]
    assert (h3[:, :K] + h3[:, 2 * K:] - hw).abs().max() < 1e-6 and torch.equal(h3[:, :K], h3[:, K:2 * K])
    assert abs(alphas["ar_pos_alpha"] - 0.75) < 1e-6
    cfg = weights.make_cfg(d, alphas, 64)
    assert cfg.ar_layers == size["ar_layers"] and abs(cfg.ln_eps - 4e-5) < 1e-9


def test_pe_and_timestep_tables_match_reference_formulas():
    pe = weights.sine_pe(50, 128)
    pos = torch.arange(50, dtype=torch.float32)[:, None]
    div = torch.exp(torch.arange(0, 128, 2, dtype=torch.float32) * -(np.log(10000.0) / 128))
    assert torch.equal(pe[:, 0::2], torch.sin(pos * div)) and torch.equal(pe[:, 1::2], torch.cos(pos * div))
    tt = weights.timestep_table(10, 128)
    assert tt.shape == (10, 128) and torch.allclose(tt[0, :64], torch.ones(64)) and torch.allclose(tt[0, 64:], torch.zeros(64))


def test_synthetic_tokenizers_and_bench_workload():
    tt, st = synth.ByteTextTok(), synth.CodeSpeechTok()
    ids = tt.encode("<|startoftext|>hi<|endoftext|>")
    assert ids == [256, ord("h"), ord("i"), 257] and len(tt.vocab) == 258 and len(st.vocab) == 1025
    assert st.encode("3 7 1023") == [3, 7, 1023] and st.decode_int([5, 1024, 9]) == [5, 9]
    import bench
    wl = bench.make_workload(synth.FULL, 2, 0)
    assert wl["N"] == 1500 and len(wl["prompts"][0]) == 137 + 450 and wl["first_codec_idx"] == 138
    assert abs(wl["audio_s"] - 2 * 1499 / 75) < 1e-9 and wl["max_len"] == 587 + 1502


def test_prepare_builds_the_reference_prompt_with_native_tokenisers():
    """Host glue of Mars5TTS.tts (inference.py:212-258) with the bpe.py tokenisers: prompt = text tokens (+ offset speech
    tokens of the reference clip when deep cloning), first_codec_idx, character-count EOS estimate."""
    import io, json, os
    import torch
    from mars5_tts_b200 import bpe
    from mars5_tts_b200.engine import InferenceConfig, Mars5TTS
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "bpe_golden.json"), encoding="utf-8"))
    tt = bpe.RegexTokenizer(); tt.load(gold["text"]["model"])
    st = bpe.CodebookTokenizer(bpe.GPT4_SPLIT_PATTERN); st.load(io.BytesIO(gold["speech"]["model"].encode("utf-8")))
    case = gold["speech"]["cases"][5]                      # 450 prompt frames with reference-made ids

    class Codec:                                           # EncodecModel.encode stand-in: (1, 8, T) codes
        def encode(self, wav):
            codes = torch.zeros(1, 8, len(case["codes"]), dtype=torch.long)
            codes[0, 0] = torch.tensor(case["codes"])
            return [(codes, None)]

    m = object.__new__(Mars5TTS)                           # no GPU context: only the host glue is exercised
    m.texttok, m.speechtok, m.codec, m.sr = tt, st, Codec(), 24000
    cfg = InferenceConfig(deep_clone=True)
    p = m._prepare("Hello world, it's me!", torch.zeros(24000), "ka to mi", cfg)
    text_ids = tt.encode("<|startoftext|>ka to mi Hello world, it's me!<|endoftext|>", allowed_special="all")
    assert p["text_tokens"] == text_ids
    assert p["prompt"] == text_ids + [i + len(tt.vocab) for i in case["ids"]]
    assert p["first_codec_idx"] == len(text_ids) + 1
    assert p["spk_ref"].shape == (450, 8) and p["n_phones"] == round(cfg.eos_estimated_gen_length_factor * len("Hello world, it's me!"))
    shallow = m._prepare("Hello", torch.zeros(24000), None, InferenceConfig(deep_clone=False))
    assert shallow["prompt"] == tt.encode("<|startoftext|>Hello<|endoftext|>", allowed_special="all")
    assert shallow["first_codec_idx"] == len(shallow["prompt"]) + 1
    import pytest
    with pytest.raises(AssertionError):
        m._prepare("x", torch.zeros(24000), None, cfg)     # deep clone without a transcript (inference.py:212-214)


def test_repack_consumes_the_released_checkpoint_layout():
    """Every key of the reference's real-size state dicts (tests/golden/state_manifest.json: CodecLM / ResidualTransformer
    built with inference.py:105-110's arguments on the meta device) is read by weights.repack and nothing else is
    expected; the derived dimensions are the released models' (SURVEY.md Appendix A)."""
    import json, os

    m = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "state_manifest.json")))
    assert len(m["ar"]) == 273 and len(m["nar"]) == 487

    class Spy(dict):
        def __init__(self, *a):
            super().__init__(*a)
            self.used = set()

        def __getitem__(self, k):
            self.used.add(k)
            return super().__getitem__(k)

    def tensors(shapes):   # shapes only (meta device); the learned PE scalars are read with .item() and must be real
        return Spy({k: torch.full(v, 0.5) if k.endswith("alpha") else torch.empty(v, device="meta") for k, v in shapes.items()})

    ar, nar = tensors(m["ar"]), tensors(m["nar"])
    d = weights.dims_from_state(ar, nar, None, m["n_text"])
    assert (d["ar_dim"], d["ar_heads"], d["ar_layers"], d["ar_hidden"], d["ar_vocab"], d["ar_spk_layers"], d["ar_spk_ff"]) == \
        (1536, 24, 26, 3584, 8000, 2, 4608)
    assert (d["nar_dim"], d["nar_heads"], d["nar_enc_layers"], d["nar_dec_layers"], d["nar_spk_layers"], d["nar_ff"], d["n_classes"],
            d["n_quant"], d["nar_text_vocab"]) == (1024, 16, 8, 16, 3, 3072, 1025, 8, 2049)
    t, alphas = weights.repack(ar, nar, None, d, max_pos=64)
    assert set(ar) == ar.used and set(nar) == nar.used            # nothing in a real checkpoint is ignored
    assert t["ar.l25.wqkv"].shape == (4608, 1536) and t["ar.l25.w13"].shape == (7168, 1536) and t["ar.output"].shape == (8000, 1536)
    assert t["nar.dec.l15.ca_kv_w"].shape == (2048, 1024) and alphas["nar_ref_alpha"] == 0.5
    # the synthetic FULL-size checkpoints of bench.py have exactly this layout
    full = synth.FULL
    assert full["n_text"] + full["n_speech"] == d["ar_vocab"] and full["ar_layers"] == d["ar_layers"] and full["nar_dec_layers"] == d["nar_dec_layers"]


def test_vocos_codebook_stride_is_the_encodec_bin_count():
    """the released vocos table concatenates the codebooks of the MAXIMUM bandwidth (16 x 1024 rows);
    codes_to_features offsets codebook q by q * 1024 whatever the table holds.  dims / oracle must not derive the stride
    from the row count."""
    from oracle import vocos_oracle
    size = synth.TINY
    ar, nar = synth.make_ar_state(size), synth.make_nar_state(size)
    v16, v8 = synth.make_vocos_state(size, n_codebooks=16), synth.make_vocos_state(size, n_codebooks=8)
    assert v16["feature_extractor.codebook_weights"].shape[0] == 16 * 1024
    assert weights.dims_from_state(ar, nar, v16, size["n_text"])["voc_codebook"] == 1024
    assert weights.dims_from_state(ar, nar, v8, size["n_text"])["voc_codebook"] == 1024
    # same first 8 x 1024 rows -> same audio, independent of how many codebooks follow
    v8["feature_extractor.codebook_weights"] = v16["feature_extractor.codebook_weights"][: 8 * 1024].clone()
    for k in v16:
        if k != "feature_extractor.codebook_weights":
            v8[k] = v16[k]
    codes = torch.randint(0, 1024, (5, 8), generator=torch.Generator().manual_seed(0))
    assert torch.equal(vocos_oracle.vocos_forward(v16, codes, 1), vocos_oracle.vocos_forward(v8, codes, 1))
    import pytest
    bad = dict(v16)
    bad["feature_extractor.codebook_weights"] = v16["feature_extractor.codebook_weights"][:4096]
    with pytest.raises(ValueError):
        weights.dims_from_state(ar, nar, bad, size["n_text"])


def test_repack_encodec_folds_weight_norm_and_oracle_shapes():
    """weights.repack_encodec accepts EncodecModel.state_dict() with weight norm still attached (weight_g / weight_v) or
    folded; the oracle yields ceil(samples / 320) frames of 8 codes."""
    from oracle import encodec_oracle
    sd = synth.make_encodec_state(n_filters=4, dimension=16, n_codebooks=8, bins=32)
    t = weights.repack_encodec(sd)
    wn = {}
    for k, v in sd.items():
        if k.endswith(".conv.conv.weight"):
            norm = v.flatten(1).norm(dim=1).reshape(-1, 1, 1)
            wn[k[:-len("weight")] + "weight_g"], wn[k[:-len("weight")] + "weight_v"] = norm * 1.0, v * 3.0   # g = |w|, v = any multiple
        else:
            wn[k] = v
    t2 = weights.repack_encodec(wn)
    assert set(t) == set(t2)
    for k in t:
        assert torch.allclose(t[k], t2[k], atol=1e-6), k
    assert t["enc.c0.w"].shape == (4, 1, 7) and t["enc.d3.w"].shape == (64, 32, 16) and t["enc.lstm1.hh.w"].shape == (256, 64)
    for n in (1, 320, 321, 999):
        assert encodec_oracle.encode(sd, torch.randn(n)).shape == ((n + 319) // 320, 8)


def test_nar_cfg_and_ar_cfg_carry_every_inference_field():
    from mars5_tts_b200 import capi
    from mars5_tts_b200.engine import Engine, InferenceConfig
    e = object.__new__(Engine)
    e._sched_cache = {}
    ic = InferenceConfig(typical_p=0.6, top_k=50, x_0_temp=0.5, nar_guidance_w=2.0, q0_override_steps=7, deep_clone=False)
    a = Engine.make_ar_cfg(e, ic, 321, 99)
    assert abs(a.typical_p - 0.6) < 1e-7 and a.top_k == 50 and a.max_len == 321 and a.eos_id == 99
    n = Engine.make_nar_cfg(e, ic, T=10, jump_len=2, jump_n_sample=3)
    assert (n.T, n.q0_override_steps, n.deep_clone, n.precise, n.jump_len, n.jump_n_sample, n.scaled_forward) == (10, 7, 0, capi.NUM_DEFAULT, 2, 3, 0)
    assert abs(n.x0_temp - 0.5) < 1e-7 and abs(n.guidance_w - 2.0) < 1e-7


def test_numerics_modes_agree_between_header_capi_and_bench():
    """m5_nar_cfg.precise: the names bench.py accepts, the constants of capi.py, the enum of csrc/layers.h and the comment of
    include/mars5_b200.h list the same modes; the default is one of the modes that hold the absolute 1e-3 logit bound
    (tests/test_zzz_fullsize_gpu.py asserts it for mixed, mixed8, mixed8k, precise)."""
    import os
    import re
    import bench
    from mars5_tts_b200 import capi
    assert bench.MODES == capi.NUM_NAMES
    assert sorted(capi.NUM_NAMES.values()) == list(range(len(capi.NUM_NAMES)))
    assert capi.NUM_DEFAULT in (capi.NUM_MIXED, capi.NUM_MIXED8, capi.NUM_MIXED8K, capi.NUM_PRECISE)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    enum = re.search(r"enum \{ (M5_NUM_FAST[^}]*)\}", open(os.path.join(root, "mars5-tts_b200", "csrc", "layers.h")).read()).group(1)
    vals = {k.strip()[len("M5_NUM_"):].lower(): int(v) for k, v in (kv.split("=") for kv in enum.split(","))}
    assert vals == capi.NUM_NAMES
    hdr = open(os.path.join(root, "include", "mars5_b200.h")).read()
    for name, v in capi.NUM_NAMES.items():
        assert re.search(rf"\b{v} {name}\b", hdr), (name, v)


def test_phase_ranges_never_swallow_errors():
    """engine._phase (NVTX range around every C-ABI call) pops its range on the way out and lets exceptions through."""
    from mars5_tts_b200.engine import _phase
    with _phase("m5_test") as r:
        assert r.name == "m5_test"
    with pytest.raises(ValueError):
        with _phase("m5_test"):
            raise ValueError("must propagate")

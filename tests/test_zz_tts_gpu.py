"""End-to-end drop-in surface on the GPU: Mars5TTS.tts / tts_batch (inference.py:201-307) = tokenisers (bpe.py) -> AR
-> speech-BPE decode -> NAR -> vocoder -> silence trim, on the tiny synthetic models.  Checks the glue, not the numerics
(those are pinned kernel by kernel in test_pipeline_gpu.py): a batch row equals the single call, the stages chained by
hand through Engine give the same codes and the same waveform, and the returned audio is the oracle's trim of it."""
import numpy as np
import pytest
import torch

from mars5_tts_b200 import bpe, synth
from mars5_tts_b200.engine import InferenceConfig, Mars5TTS
from oracle import trim_oracle

pytestmark = pytest.mark.gpu

TEXT_MODEL = "minbpe v1\n" + bpe.GPT4_SPLIT_PATTERN + "\n2\n<|startoftext|> 256\n<|endoftext|> 257\n"
SPEECH_MODEL = "minbpe v1\n\n1\n<|endofspeech|> 1024\n"
PF = 12


class _Codec:
    """EncodecModel.encode stand-in (inference.py:233): deterministic (1, 8, PF) codes from the clip's first sample."""

    def encode(self, wav):
        g = torch.Generator().manual_seed(int(abs(float(wav.flatten()[0])) * 1000) + 7)
        return [(torch.randint(0, 1024, (1, 8, PF), generator=g), None)]


@pytest.fixture(scope="module")
def tts():
    size = synth.TINY
    ar = {"model": synth.make_ar_state(size), "vocab": {"texttok.model": TEXT_MODEL, "speechtok.model": SPEECH_MODEL}}
    nar = {"model": synth.make_nar_state(size)}
    m = Mars5TTS(ar, nar, "cuda:0", vocos_state=synth.make_vocos_state(size), codec=_Codec())
    assert len(m.texttok.vocab) == 258 and len(m.speechtok.vocab) == 1025
    yield m
    m.engine.close()


def test_tts_batch_rows_equal_single_calls_and_hand_chained_stages(tts):
    cfg = InferenceConfig(generate_max_len_override=56, deep_clone=True)
    texts, refs, trs = ["hello there", "b200"], [torch.full((2400,), 0.25), torch.full((2400,), 0.5)], ["ab cd", "xy"]
    out = tts.tts_batch(texts, refs, trs, cfg)
    assert len(out) == 2
    for codes, wav in out:
        assert codes.dtype == torch.long and codes.is_cuda and codes.dim() == 1 and codes.numel() > 0
        assert wav.dtype == torch.float32 and not wav.is_cuda and wav.dim() == 1 and torch.isfinite(wav).all()
    # single call == row 0 of the batch (results are keyed by utterance id, not by batch position or size)
    tts._calls = 0
    c0, w0 = tts.tts(texts[0], refs[0], trs[0], cfg, seed=0)
    assert torch.equal(c0, out[0][0]) and torch.equal(w0, out[0][1])
    # like the reference (unseeded torch generator) the next call draws fresh randomness: call counter -> utterance id
    c1, w1 = tts.tts(texts[0], refs[0], trs[0], cfg, seed=0)
    assert tts._calls == 2 and not (c1.numel() == c0.numel() and torch.equal(c1, c0) and torch.equal(w1, w0))
    # the same stages chained by hand through Engine
    eng = tts.engine
    preps = [tts._prepare(t, a, r, cfg) for t, a, r in zip(texts, refs, trs)]
    eos = len(tts.texttok.vocab) + tts.speechtok.special_tokens["<|endofspeech|>"]
    ids, hit, _ = eng.ar_generate([p["prompt"] for p in preps], [p["spk_ref"] for p in preps], [p["n_phones"] for p in preps],
                                  eng.make_ar_cfg(cfg, 56, eos), seed=0, utt_ids=[0, 1])
    l0s = []
    for p, seq in zip(preps, ids):
        toks = np.clip(seq.astype(np.int64) - 258, 0, None)[p["first_codec_idx"]:].tolist()
        l0s.append(np.asarray([c for c in tts.speechtok.decode_int(toks) if type(c) == int], dtype=np.int32))
    codes = eng.nar_infer([p["text_tokens"] for p in preps], [p["spk_ref"] for p in preps], l0s, eng.make_nar_cfg(cfg, T=tts.default_T),
                          seed=0, utt_ids=[0, 1])
    wavs = eng.vocode([c[PF:] for c in codes], bandwidth_id=1)
    for (got_codes, got_wav), l0, w in zip(out, l0s, wavs):
        assert got_codes.cpu().tolist() == l0.tolist()
        a, b = trim_oracle.trim_bounds(torch.from_numpy(w), cfg.trim_db)
        assert got_wav.numel() == b - a and torch.equal(got_wav, torch.from_numpy(w)[a:b])


def test_device_trim_equals_host_trim_and_oracle(tts):
    """m5_vocode_trim (frame powers on the device, behind the overlap-add) returns the very bounds of the host entry point
    m5_trim_bounds and of the float32 oracle (pinned to the reference's own trim() by tests/golden/trim_golden.json) on
    vocoder output of different lengths, including a near-silent one and an all-pad (constant) utterance."""
    from mars5_tts_b200.trim import trim_bounds_batch
    eng = tts.engine
    g = torch.Generator().manual_seed(3)
    codes = [torch.randint(0, 1024, (n, 8), generator=g).numpy().astype(np.int32) for n in (4, 9, 40, 150)]
    codes.append(np.zeros((30, 8), dtype=np.int32))
    for db in (27.0, 5.0, 60.0):
        wavs, bounds = eng.vocode_trim(codes, db)
        ref_wavs = eng.vocode(codes)
        host = trim_bounds_batch(ref_wavs, top_db=db)
        for w, rw, b, h in zip(wavs, ref_wavs, bounds, host):
            assert np.array_equal(w, rw)
            assert b == h == trim_oracle.trim_bounds(torch.from_numpy(rw), db), (db, b, h)


def test_results_do_not_depend_on_sharding(tts):
    """Multi-GPU invariant: utterances are keyed by (seed, utterance id), so serving ids 0..3 as one batch
    of four, or as two shards of two (what ranks 0 and 1 of a 2-GPU job would each run, `utt_base` = first id of the
    shard), yields the same L0 codes and the same waveforms."""
    cfg = InferenceConfig(generate_max_len_override=48, deep_clone=True)
    texts = ["one", "two two", "three", "4"]
    refs = [torch.full((2400,), 0.1 * (i + 1)) for i in range(4)]
    trs = ["a b", "c", "d e f", "g"]
    whole = tts.tts_batch(texts, refs, trs, cfg, seed=11, utt_base=0)
    shard0 = tts.tts_batch(texts[:2], refs[:2], trs[:2], cfg, seed=11, utt_base=0)
    shard1 = tts.tts_batch(texts[2:], refs[2:], trs[2:], cfg, seed=11, utt_base=2)
    for (c, w), (c2, w2) in zip(whole, shard0 + shard1):
        assert torch.equal(c, c2)
        # same discrete codes => same audio up to the vocoder GEMMs' tile choice (it depends on the batch's total frame count
        # and changes the fp32 accumulation order): a flipped NAR code would show as an O(0.1) difference
        assert w.shape == w2.shape and float((w - w2).abs().max()) < 1e-4 * max(1.0, float(w.abs().max()))
    # and a different seed or utterance id is a different stream
    other = tts.tts_batch(texts[:1], refs[:1], trs[:1], cfg, seed=12, utt_base=0)
    assert not (other[0][1].shape == whole[0][1].shape and torch.equal(other[0][1], whole[0][1]))

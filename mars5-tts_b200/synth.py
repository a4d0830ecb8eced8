"""Synthetic, seeded, reference-FORMAT checkpoints (there is no network for the real MARS5 / Vocos weights).

The state dicts carry exactly the key names and shapes of the reference modules (CodecLM / ResidualTransformer,
SURVEY.md Appendix A; vocos module names for the vocoder), with values rounded to fp16 like the released checkpoints,
so the very same `weights.repack` path serves real checkpoints.  tests/golden/make_golden.py asserts the key sets
against the reference's own `state_dict()`.
"""
import torch

FULL = dict(ar_dim=1536, ar_layers=26, ar_spk_layers=2, nar_dim=1024, nar_enc_layers=8, nar_dec_layers=16,
            nar_spk_layers=3, n_text=2048, n_speech=5952, voc_feat=128, voc_dim=384, voc_inter=1152, voc_layers=8)
TINY = dict(ar_dim=192, ar_layers=2, ar_spk_layers=1, nar_dim=128, nar_enc_layers=1, nar_dec_layers=2,
            nar_spk_layers=1, n_text=258, n_speech=1025, voc_feat=64, voc_dim=128, voc_inter=256, voc_layers=2)
MID = dict(ar_dim=768, ar_layers=4, ar_spk_layers=2, nar_dim=512, nar_enc_layers=2, nar_dec_layers=4,
           nar_spk_layers=2, n_text=258, n_speech=1025, voc_feat=128, voc_dim=384, voc_inter=1152, voc_layers=3)


class _Init:
    def __init__(self, seed):
        self.g = torch.Generator().manual_seed(seed)

    def lin(self, out_f, in_f, scale=1.0):
        return (torch.randn(out_f, in_f, generator=self.g) * (scale / in_f ** 0.5)).half().float()

    def vec(self, n, mean=0.0, std=0.02):
        return (torch.randn(n, generator=self.g) * std + mean).half().float()

    def emb(self, n, d, std=1.0):
        return (torch.randn(n, d, generator=self.g) * std).half().float()


def _enc_layer(sd, it, p, dim, ff):
    sd[p + "self_attn.in_proj_weight"] = it.lin(3 * dim, dim)
    sd[p + "self_attn.in_proj_bias"] = it.vec(3 * dim)
    sd[p + "self_attn.out_proj.weight"] = it.lin(dim, dim)
    sd[p + "self_attn.out_proj.bias"] = it.vec(dim)
    sd[p + "linear2.weight"] = it.lin(dim, ff)
    sd[p + "linear2.bias"] = it.vec(dim)
    for n in ("norm1", "norm2"):
        sd[p + n + ".weight"] = it.vec(dim, 1.0, 0.05)
        sd[p + n + ".bias"] = it.vec(dim)
    sd[p + "activation.V.weight"] = it.lin(ff, dim)
    sd[p + "activation.W.weight"] = it.lin(ff, dim)


def make_ar_state(size, seed=0):
    """CodecLM(n_vocab, dim, nhead=dim/64, n_layers, n_spk_layers, dim_ff_scale=7/3) state dict (model.py:44-67)."""
    it = _Init(seed)
    dim, V = size["ar_dim"], size["n_text"] + size["n_speech"]
    hidden, spk_ff = int(dim * 7 / 3), int(dim * 4 * (3 / 4))
    sd = {}
    for l in range(size["ar_layers"]):
        p = f"ar.layers.{l}."
        for n in ("wq", "wk", "wv", "wo"):
            sd[p + f"attention.{n}.weight"] = it.lin(dim, dim)
        sd[p + "feed_forward.w1.weight"] = it.lin(hidden, dim)
        sd[p + "feed_forward.w2.weight"] = it.lin(dim, hidden)
        sd[p + "feed_forward.w3.weight"] = it.lin(hidden, dim)
        sd[p + "attention_norm.weight"] = it.vec(dim, 1.0, 0.05)
        sd[p + "ffn_norm.weight"] = it.vec(dim, 1.0, 0.05)
    sd["ar.norm.weight"] = it.vec(dim, 1.0, 0.05)
    sd["ar.output.weight"] = it.lin(V, dim, scale=2.0)
    sd["embed.weight"] = it.emb(V, dim)
    sd["pos_embedding.alpha"] = torch.tensor([0.75])
    for q in range(8):
        sd[f"ref_chunked_emb.embs.{q}.weight"] = it.emb(1025, dim // 8)
    sd["spk_identity_emb.weight"] = it.emb(1, dim)
    for l in range(size["ar_spk_layers"]):
        _enc_layer(sd, it, f"spk_encoder.layers.{l}.", dim, spk_ff)
    sd["spk_encoder.norm.weight"] = it.vec(dim, 1.0, 0.05)
    sd["spk_encoder.norm.bias"] = it.vec(dim)
    return sd


def make_nar_state(size, seed=1):
    """ResidualTransformer(n_text_vocab, n_quant=1025, dim, nhead=dim/64, ...) state dict (model.py:165-242)."""
    it = _Init(seed)
    dim = size["nar_dim"]
    ff = int(dim * 4 * (3 / 4))
    sd = {}
    for l in range(size["nar_enc_layers"]):
        _enc_layer(sd, it, f"tfm.encoder.layers.{l}.", dim, ff)
    sd["tfm.encoder.norm.weight"], sd["tfm.encoder.norm.bias"] = it.vec(dim, 1.0, 0.05), it.vec(dim)
    for l in range(size["nar_dec_layers"]):
        p = f"tfm.decoder.layers.{l}."
        _enc_layer(sd, it, p, dim, ff)
        sd[p + "multihead_attn.in_proj_weight"] = it.lin(3 * dim, dim)
        sd[p + "multihead_attn.in_proj_bias"] = it.vec(3 * dim)
        sd[p + "multihead_attn.out_proj.weight"] = it.lin(dim, dim)
        sd[p + "multihead_attn.out_proj.bias"] = it.vec(dim)
        sd[p + "norm3.weight"], sd[p + "norm3.bias"] = it.vec(dim, 1.0, 0.05), it.vec(dim)
    sd["tfm.decoder.norm.weight"], sd["tfm.decoder.norm.bias"] = it.vec(dim, 1.0, 0.05), it.vec(dim)
    for n in ("timestep_encoder_emb", "timestep_decoder_emb"):
        sd[n + ".0.weight"], sd[n + ".0.bias"] = it.lin(dim, dim), it.vec(dim)
        sd[n + ".2.weight"], sd[n + ".2.bias"] = it.lin(dim, dim), it.vec(dim)
    sd["text_embed.weight"] = it.emb(size["n_text"] + 1, dim)
    for q in range(8):
        sd[f"ref_embedder.embs.{q}.weight"] = it.emb(1025, dim // 8)
        sd[f"residual_encoder.embs.{q}.weight"] = it.emb(1025, dim // 8)
    for n, a in (("cond_pos_embedding", 0.8), ("ref_pos_embedding", 1.1), ("pos_embedding", 0.9)):
        sd[n + ".alpha"] = torch.tensor([a]).half().float()
    sd["spk_identity_emb.weight"] = it.emb(1, dim)
    for l in range(size["nar_spk_layers"]):
        _enc_layer(sd, it, f"spk_encoder.layers.{l}.", dim, ff)
    sd["spk_encoder.norm.weight"], sd["spk_encoder.norm.bias"] = it.vec(dim, 1.0, 0.05), it.vec(dim)
    for q in range(8):
        sd[f"residual_decoder.{q}.0.weight"], sd[f"residual_decoder.{q}.0.bias"] = it.vec(dim, 1.0, 0.05), it.vec(dim)
        sd[f"residual_decoder.{q}.1.weight"] = it.lin(1025, dim, scale=2.0)
        sd[f"residual_decoder.{q}.1.bias"] = it.vec(1025)
    return sd


def make_vocos_state(size, seed=2, n_fft=1280, n_bw=4, n_codebooks=16):
    """Vocos encodec-24khz state dict with vocos' own module names (fp32 values, not fp16-exact, like the real one)."""
    g = torch.Generator().manual_seed(seed)
    feat, dim, inter = size["voc_feat"], size["voc_dim"], size["voc_inter"]

    def rn(*shape, std=1.0):
        return torch.randn(*shape, generator=g) * std

    # like the released checkpoint: the codebooks of the maximum bandwidth (12 kbps = 16 x 1024 rows), 8 of which are used
    sd = {"feature_extractor.codebook_weights": rn(n_codebooks * 1024, feat, std=0.5)}
    sd["backbone.embed.weight"] = rn(dim, feat, 7, std=1.0 / (7 * feat) ** 0.5)
    sd["backbone.embed.bias"] = rn(dim, std=0.02)
    sd["backbone.norm.scale.weight"] = 1.0 + rn(n_bw, dim, std=0.05)
    sd["backbone.norm.shift.weight"] = rn(n_bw, dim, std=0.05)
    for l in range(size["voc_layers"]):
        p = f"backbone.convnext.{l}."
        sd[p + "dwconv.weight"] = rn(dim, 1, 7, std=0.35)
        sd[p + "dwconv.bias"] = rn(dim, std=0.02)
        sd[p + "norm.scale.weight"] = 1.0 + rn(n_bw, dim, std=0.05)
        sd[p + "norm.shift.weight"] = rn(n_bw, dim, std=0.05)
        sd[p + "pwconv1.weight"], sd[p + "pwconv1.bias"] = rn(inter, dim, std=1.0 / dim ** 0.5), rn(inter, std=0.02)
        sd[p + "pwconv2.weight"], sd[p + "pwconv2.bias"] = rn(dim, inter, std=1.0 / inter ** 0.5), rn(dim, std=0.02)
        sd[p + "gamma"] = 0.3 + rn(dim, std=0.05)
    sd["backbone.final_layer_norm.weight"], sd["backbone.final_layer_norm.bias"] = 1.0 + rn(dim, std=0.05), rn(dim, std=0.02)
    sd["head.out.weight"] = rn(n_fft + 2, dim, std=0.6 / dim ** 0.5)
    sd["head.out.bias"] = rn(n_fft + 2, std=0.1)
    return sd


def make_encodec_state(seed=3, n_filters=32, dimension=128, n_codebooks=32, bins=1024):
    """Encodec 24 kHz encoder + RVQ state dict with encodec's own module names after weight norm has been folded
    (oracle/encodec_oracle.py header): seeded random fp32 values at the real shapes (n_filters / dimension can be shrunk for
    tests; the released model is n_filters=32, dimension=128, 32 codebooks of 1024 x 128)."""
    g = torch.Generator().manual_seed(seed)

    def conv(co, ci, k):
        return torch.randn(co, ci, k, generator=g) * (1.0 / (ci * k) ** 0.5), torch.randn(co, generator=g) * 0.02

    sd, p = {}, "encoder.model."
    sd[p + "0.conv.conv.weight"], sd[p + "0.conv.conv.bias"] = conv(n_filters, 1, 7)
    idx, mult = 1, 1
    for ratio in (2, 4, 5, 8):
        dim = mult * n_filters
        sd[p + f"{idx}.block.1.conv.conv.weight"], sd[p + f"{idx}.block.1.conv.conv.bias"] = conv(dim // 2, dim, 3)
        sd[p + f"{idx}.block.3.conv.conv.weight"], sd[p + f"{idx}.block.3.conv.conv.bias"] = conv(dim, dim // 2, 1)
        sd[p + f"{idx}.shortcut.conv.conv.weight"], sd[p + f"{idx}.shortcut.conv.conv.bias"] = conv(dim, dim, 1)
        sd[p + f"{idx + 2}.conv.conv.weight"], sd[p + f"{idx + 2}.conv.conv.bias"] = conv(2 * dim, dim, 2 * ratio)
        idx += 3
        mult *= 2
    H = mult * n_filters
    for layer in range(2):
        for nm, shape in (("weight_ih", (4 * H, H)), ("weight_hh", (4 * H, H)), ("bias_ih", (4 * H,)), ("bias_hh", (4 * H,))):
            sd[p + f"{idx}.lstm.{nm}_l{layer}"] = torch.randn(*shape, generator=g) * (1.0 / H ** 0.5)
    sd[p + f"{idx + 2}.conv.conv.weight"], sd[p + f"{idx + 2}.conv.conv.bias"] = conv(dimension, H, 7)
    for q in range(n_codebooks):
        sd[f"quantizer.vq.layers.{q}._codebook.embed"] = torch.randn(bins, dimension, generator=g) * (0.7 ** q)
    return sd


class ByteTextTok:
    """Minimal stand-in with the reference tokenizers' interface (minbpe v1, no merges): 256 bytes + 2 specials."""

    def __init__(self):
        self.special_tokens = {"<|startoftext|>": 256, "<|endoftext|>": 257}
        self.vocab = {i: bytes([i]) for i in range(256)}
        self.vocab.update({256: b"<|startoftext|>", 257: b"<|endoftext|>"})

    def encode(self, text, allowed_special="all"):
        out, i = [], 0
        while i < len(text):
            for s, idx in self.special_tokens.items():
                if text.startswith(s, i):
                    out.append(idx)
                    i += len(s)
                    break
            else:
                out.extend(text[i].encode("utf-8"))
                i += 1
        return out


class CodeSpeechTok:
    """Speech-BPE stand-in: 1024 codes + <|endofspeech|>, one token per Encodec frame (ratio 1)."""

    def __init__(self):
        self.special_tokens = {"<|endofspeech|>": 1024}
        self.vocab = {i: (i,) for i in range(1024)}
        self.vocab[1024] = "<|endofspeech|>"
        # len(vocab) = 1025 ; n_vocab = len(text vocab) + 1025

    def encode(self, s, allowed_special="all"):
        return [int(t) for t in s.split()] if s.strip() else []

    def decode_int(self, ids):
        return [i for i in ids if i < 1024]

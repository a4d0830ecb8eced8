"""Weight repacking: reference-format checkpoints -> the named device tensors libmars5_b200.so consumes.

Input state dicts use the reference's own key names (``ar_ckpt['model']`` for CodecLM, ``nar_ckpt['model']`` for
ResidualTransformer -- /root/reference/inference.py:105-111, SURVEY.md Appendix A -- and vocos' module names for the
vocoder).  Output: fp16 GEMM operands (the checkpoints are fp16-exact, README.md:163-164), fused QKV, SwiGLU pairs
interleaved row-wise so the GEMM epilogue can gate adjacent columns, fp32 norms/biases, plus the derived tables
(sinusoidal position tables, RoPE frequencies, timestep embeddings, iSTFT twiddles) computed with torch so that they
are bit-identical to what the reference computes at run time.  This is load-time plumbing; no hot-path arithmetic.
"""
import math

import torch

from . import capi


def dims_from_state(ar_sd, nar_sd, voc_sd, text_vocab_len):
    """Derive m5_model_cfg integers from tensor shapes (nothing is hard-coded to the released 750M/450M sizes)."""
    d = {}
    d["ar_vocab"], d["ar_dim"] = ar_sd["embed.weight"].shape
    d["ar_heads"] = d["ar_dim"] // 64
    d["ar_layers"] = 1 + max(int(k.split(".")[2]) for k in ar_sd if k.startswith("ar.layers."))
    d["ar_hidden"] = ar_sd["ar.layers.0.feed_forward.w1.weight"].shape[0]
    d["ar_text_vocab"] = text_vocab_len
    d["ar_spk_layers"] = 1 + max(int(k.split(".")[2]) for k in ar_sd if k.startswith("spk_encoder.layers."))
    d["ar_spk_ff"] = ar_sd["spk_encoder.layers.0.activation.W.weight"].shape[0]
    d["nar_text_vocab"], d["nar_dim"] = nar_sd["text_embed.weight"].shape
    d["nar_heads"] = d["nar_dim"] // 64
    d["nar_enc_layers"] = 1 + max(int(k.split(".")[3]) for k in nar_sd if k.startswith("tfm.encoder.layers."))
    d["nar_dec_layers"] = 1 + max(int(k.split(".")[3]) for k in nar_sd if k.startswith("tfm.decoder.layers."))
    d["nar_spk_layers"] = 1 + max(int(k.split(".")[2]) for k in nar_sd if k.startswith("spk_encoder.layers."))
    d["nar_ff"] = nar_sd["tfm.decoder.layers.0.activation.W.weight"].shape[0]
    d["n_classes"] = nar_sd["residual_decoder.0.1.weight"].shape[0]
    d["n_quant"] = 1 + max(int(k.split(".")[1]) for k in nar_sd if k.startswith("residual_decoder."))
    if voc_sd is not None:
        d["voc_dim"], d["voc_feat"], _ = voc_sd["backbone.embed.weight"].shape
        d["voc_inter"] = voc_sd["backbone.convnext.0.pwconv1.weight"].shape[0]
        d["voc_layers"] = 1 + max(int(k.split(".")[2]) for k in voc_sd if k.startswith("backbone.convnext."))
        d["voc_nfft"] = voc_sd["head.out.weight"].shape[0] - 2
        d["voc_hop"] = d["voc_nfft"] // 4
        d["voc_n_bw"] = voc_sd["backbone.norm.scale.weight"].shape[0]
        # vocos' codes_to_features offsets codebook q by q * quantizer.bins (= 1024 Encodec codes = n_classes - 1); the
        # released charactr/vocos-encodec-24khz table concatenates the codebooks of the MAXIMUM bandwidth (16 x 1024 rows),
        # of which an 8-codebook input touches the first 8 x 1024
        d["voc_codebook"] = d["n_classes"] - 1
        rows = voc_sd["feature_extractor.codebook_weights"].shape[0]
        if rows < d["n_quant"] * d["voc_codebook"]:
            raise ValueError(f"vocos codebook_weights has {rows} rows, need at least n_quant * bins = {d['n_quant'] * d['voc_codebook']}")
    return d


def sine_pe(n, dim):
    """Position table of SinePositionalEmbedding (nn_future.py:51-76), same fp32 op order."""
    pe = torch.zeros(n, dim)
    position = torch.arange(0, n, dtype=torch.float32).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, dim, 2, dtype=torch.float32) * -(math.log(10000.0) / dim))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe


def timestep_table(n_t, dim, max_period=10000):
    """timestep_embedding (model.py:18-35) for t = 0..n_t-1."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(start=0, end=half) / half)
    args = torch.arange(n_t)[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def diffusion_schedule(T):
    """MultinomialDiffusion tables (diffuser.py:76-109) -> float32 array (4, T):
    log_alpha, log_1_min_alpha, log_cumprod_alpha, log_1_min_cumprod_alpha."""
    x = torch.linspace(0, T, T + 1)
    ac = torch.cos(((x / T) + 0.008) / (1 + 0.008) * torch.pi * 0.5) ** 2
    ac = ac / ac[0]
    alphas = torch.sqrt(torch.clamp(ac[1:] / ac[:-1], 0.001, 1.0)).to(torch.float64)
    la = alphas.log()
    lc = torch.cumsum(la, dim=-1)

    def l1m(a):
        return torch.log((1 - a.exp()).clamp_(min=1e-30))

    return torch.stack([la, l1m(la), lc, l1m(lc)]).to(torch.float32).contiguous()


def _interleave(w, v):
    """rows (W_0, V_0, W_1, V_1, ...): the GEMM epilogue gates adjacent output columns."""
    return torch.stack([w, v], dim=1).reshape(2 * w.shape[0], w.shape[1])


def _split3(w):
    """fp32 weight -> [W_hi | W_hi | W_lo] fp16 (three-term split product against [A_hi | A_lo])."""
    hi = w.half()
    lo = (w - hi.float()).half()
    return torch.cat([hi, hi, lo], dim=1)


def _enc_layer(out, sd, src, dst):
    out[dst + "n1w"], out[dst + "n1b"] = sd[src + "norm1.weight"].float(), sd[src + "norm1.bias"].float()
    out[dst + "n2w"], out[dst + "n2b"] = sd[src + "norm2.weight"].float(), sd[src + "norm2.bias"].float()
    out[dst + "in_w"], out[dst + "in_b"] = sd[src + "self_attn.in_proj_weight"].half(), sd[src + "self_attn.in_proj_bias"].float()
    out[dst + "out_w"], out[dst + "out_b"] = sd[src + "self_attn.out_proj.weight"].half(), sd[src + "self_attn.out_proj.bias"].float()
    out[dst + "wv"] = _interleave(sd[src + "activation.W.weight"], sd[src + "activation.V.weight"]).half()
    out[dst + "w2"], out[dst + "b2"] = sd[src + "linear2.weight"].half(), sd[src + "linear2.bias"].float()


def repack(ar_sd, nar_sd, voc_sd, dims, max_pos=4096, n_t=1000):
    """Returns (tensors: name -> CPU tensor, alphas: dict)."""
    t = {}
    D = dims["ar_dim"]
    # ---- AR (CodecLM) ------------------------------------------------------------------------------------
    t["ar.embed"] = ar_sd["embed.weight"].half()
    t["ar.output"] = ar_sd["ar.output.weight"].half()
    t["ar.norm"] = ar_sd["ar.norm.weight"].float()
    for i in range(dims["ar_layers"]):
        s, p = f"ar.layers.{i}.", f"ar.l{i}."
        t[p + "attn_norm"] = ar_sd[s + "attention_norm.weight"].float()
        t[p + "ffn_norm"] = ar_sd[s + "ffn_norm.weight"].float()
        t[p + "wqkv"] = torch.cat([ar_sd[s + "attention.wq.weight"], ar_sd[s + "attention.wk.weight"],
                                   ar_sd[s + "attention.wv.weight"]], dim=0).half()
        t[p + "wo"] = ar_sd[s + "attention.wo.weight"].half()
        t[p + "w13"] = _interleave(ar_sd[s + "feed_forward.w1.weight"], ar_sd[s + "feed_forward.w3.weight"]).half()
        t[p + "w2"] = ar_sd[s + "feed_forward.w2.weight"].half()
    t["ar.spk.tables"] = torch.stack([ar_sd[f"ref_chunked_emb.embs.{q}.weight"] for q in range(dims["n_quant"])]).half()
    t["ar.spk.identity"] = ar_sd["spk_identity_emb.weight"].reshape(-1).float()
    for i in range(dims["ar_spk_layers"]):
        _enc_layer(t, ar_sd, f"spk_encoder.layers.{i}.", f"ar.spk.l{i}.")
    t["ar.spk.norm_w"], t["ar.spk.norm_b"] = ar_sd["spk_encoder.norm.weight"].float(), ar_sd["spk_encoder.norm.bias"].float()
    # ---- NAR (ResidualTransformer) -----------------------------------------------------------------------
    Dn = dims["nar_dim"]
    t["nar.text_embed"] = nar_sd["text_embed.weight"].half()
    t["nar.ref.tables"] = torch.stack([nar_sd[f"ref_embedder.embs.{q}.weight"] for q in range(dims["n_quant"])]).half()
    t["nar.res.tables"] = torch.stack([nar_sd[f"residual_encoder.embs.{q}.weight"] for q in range(dims["n_quant"])]).half()
    t["nar.spk_identity"] = nar_sd["spk_identity_emb.weight"].reshape(-1).float()
    for i in range(dims["nar_spk_layers"]):
        _enc_layer(t, nar_sd, f"spk_encoder.layers.{i}.", f"nar.spk.l{i}.")
    t["nar.spk.norm_w"], t["nar.spk.norm_b"] = nar_sd["spk_encoder.norm.weight"].float(), nar_sd["spk_encoder.norm.bias"].float()
    for i in range(dims["nar_enc_layers"]):
        _enc_layer(t, nar_sd, f"tfm.encoder.layers.{i}.", f"nar.enc.l{i}.")
    t["nar.enc.norm_w"], t["nar.enc.norm_b"] = nar_sd["tfm.encoder.norm.weight"].float(), nar_sd["tfm.encoder.norm.bias"].float()
    for i in range(dims["nar_dec_layers"]):
        s, p = f"tfm.decoder.layers.{i}.", f"nar.dec.l{i}."
        for j in (1, 2, 3):
            t[p + f"n{j}w"], t[p + f"n{j}b"] = nar_sd[s + f"norm{j}.weight"].float(), nar_sd[s + f"norm{j}.bias"].float()
        t[p + "sa_in_w"], t[p + "sa_in_b"] = nar_sd[s + "self_attn.in_proj_weight"].half(), nar_sd[s + "self_attn.in_proj_bias"].float()
        t[p + "sa_out_w"], t[p + "sa_out_b"] = nar_sd[s + "self_attn.out_proj.weight"].half(), nar_sd[s + "self_attn.out_proj.bias"].float()
        cw, cb = nar_sd[s + "multihead_attn.in_proj_weight"], nar_sd[s + "multihead_attn.in_proj_bias"]
        t[p + "ca_q_w"], t[p + "ca_q_b"] = cw[:Dn].half(), cb[:Dn].float()
        t[p + "ca_kv_w"], t[p + "ca_kv_b"] = cw[Dn:].half(), cb[Dn:].float()
        t[p + "ca_out_w"], t[p + "ca_out_b"] = nar_sd[s + "multihead_attn.out_proj.weight"].half(), nar_sd[s + "multihead_attn.out_proj.bias"].float()
        t[p + "wv"] = _interleave(nar_sd[s + "activation.W.weight"], nar_sd[s + "activation.V.weight"]).half()
        t[p + "w2"], t[p + "b2"] = nar_sd[s + "linear2.weight"].half(), nar_sd[s + "linear2.bias"].float()
        # e4m3 copies scaled by 2^+2 for the fp8 lo pass of the `mixed8` numerics: the product with
        # the e5m2 lo halves (scaled by 2^-2) lands at true scale in the fp16 pass's TMEM accumulator
        if hasattr(torch, "float8_e4m3fn") and t[p + "wv"].device.type != "meta":
            for nm in ("sa_in_w", "sa_out_w", "ca_out_w", "wv", "w2"):
                t[p + nm + "8"] = (t[p + nm].float() * 4.0).clamp(-448.0, 448.0).to(torch.float8_e4m3fn).view(torch.uint8)
    t["nar.dec.norm_w"], t["nar.dec.norm_b"] = nar_sd["tfm.decoder.norm.weight"].float(), nar_sd["tfm.decoder.norm.bias"].float()
    for nm, src in (("t_enc", "timestep_encoder_emb"), ("t_dec", "timestep_decoder_emb")):
        t[f"nar.{nm}.w0"], t[f"nar.{nm}.b0"] = nar_sd[src + ".0.weight"].half(), nar_sd[src + ".0.bias"].float()
        t[f"nar.{nm}.w1"], t[f"nar.{nm}.b1"] = nar_sd[src + ".2.weight"].half(), nar_sd[src + ".2.bias"].float()
    for q in range(dims["n_quant"]):
        t[f"nar.head.{q}.ln_w"], t[f"nar.head.{q}.ln_b"] = nar_sd[f"residual_decoder.{q}.0.weight"].float(), nar_sd[f"residual_decoder.{q}.0.bias"].float()
        t[f"nar.head.{q}.w"], t[f"nar.head.{q}.b"] = nar_sd[f"residual_decoder.{q}.1.weight"].half(), nar_sd[f"residual_decoder.{q}.1.bias"].float()
    # ---- derived tables ----------------------------------------------------------------------------------
    t["tab.pe_ar"] = sine_pe(max_pos, D)
    t["tab.pe_nar"] = sine_pe(max_pos, Dn)
    t["tab.rope_inv_freq"] = (1.0 / (10000.0 ** (torch.arange(0, 64, 2)[:32].float() / 64))).float()
    t["tab.t_emb"] = timestep_table(n_t, Dn)
    alphas = {
        "ar_pos_alpha": float(ar_sd["pos_embedding.alpha"].reshape(-1)[0]),
        "nar_pos_alpha": float(nar_sd["pos_embedding.alpha"].reshape(-1)[0]),
        "nar_cond_alpha": float(nar_sd["cond_pos_embedding.alpha"].reshape(-1)[0]),
        "nar_ref_alpha": float(nar_sd["ref_pos_embedding.alpha"].reshape(-1)[0]),
    }
    # ---- Vocos -------------------------------------------------------------------------------------------
    if voc_sd is not None:
        t["voc.codebook"] = voc_sd["feature_extractor.codebook_weights"].float()
        w = voc_sd["backbone.embed.weight"].float()  # (dim, feat, 7) -> im2col layout (dim, 7*feat), column = tap*feat + c
        t["voc.embed_w"] = _split3(w.permute(0, 2, 1).reshape(w.shape[0], -1))
        t["voc.embed_b"] = voc_sd["backbone.embed.bias"].float()
        t["voc.norm_scale"] = voc_sd["backbone.norm.scale.weight"].float()
        t["voc.norm_shift"] = voc_sd["backbone.norm.shift.weight"].float()
        for i in range(dims["voc_layers"]):
            s, p = f"backbone.convnext.{i}.", f"voc.l{i}."
            t[p + "dw_w"] = voc_sd[s + "dwconv.weight"].float().reshape(-1, 7)
            t[p + "dw_b"] = voc_sd[s + "dwconv.bias"].float()
            t[p + "norm_scale"], t[p + "norm_shift"] = voc_sd[s + "norm.scale.weight"].float(), voc_sd[s + "norm.shift.weight"].float()
            t[p + "pw1_w"], t[p + "pw1_b"] = _split3(voc_sd[s + "pwconv1.weight"].float()), voc_sd[s + "pwconv1.bias"].float()
            t[p + "pw2_w"], t[p + "pw2_b"] = _split3(voc_sd[s + "pwconv2.weight"].float()), voc_sd[s + "pwconv2.bias"].float()
            t[p + "gamma"] = voc_sd[s + "gamma"].float()
        t["voc.final_ln_w"], t["voc.final_ln_b"] = voc_sd["backbone.final_layer_norm.weight"].float(), voc_sd["backbone.final_layer_norm.bias"].float()
        t["voc.head_w"], t["voc.head_b"] = _split3(voc_sd["head.out.weight"].float()), voc_sd["head.out.bias"].float()
        k = torch.arange(640, dtype=torch.float64)
        t["voc.w1280"] = torch.stack([torch.cos(2 * math.pi * k / 1280), torch.sin(2 * math.pi * k / 1280)], dim=1).float()
        t["voc.w640"] = torch.stack([torch.cos(2 * math.pi * k / 640), torch.sin(2 * math.pi * k / 640)], dim=1).float()
        t["voc.window"] = torch.hann_window(1280, dtype=torch.float64).float()
    return {k: v.contiguous() for k, v in t.items()}, alphas


def encodec_keys_from_hf(sd):
    """State dict of `transformers`' EncodecModel (facebook/encodec_24khz: `encoder.layers.{i}...`, `quantizer.layers.{q}.codebook.embed`)
    -> the key names of the `encodec` package (`encoder.model.{i}...conv.conv...`, `quantizer.vq.layers.{q}._codebook.embed`) that
    repack_encodec and oracle/encodec_oracle.py use.  Tensors are shared, decoder / bookkeeping entries dropped; a dict that already
    has encodec's names is returned unchanged."""
    import re
    if not any(k.startswith("encoder.layers.") for k in sd):
        return sd
    out = {}
    for k, v in sd.items():
        m = re.match(r"encoder\.layers\.(\d+)\.(block\.\d+\.|shortcut\.)?conv\.(.+)$", k)
        if m:
            out[f"encoder.model.{m.group(1)}.{m.group(2) or ''}conv.conv.{m.group(3)}"] = v
            continue
        m = re.match(r"encoder\.layers\.(\d+)\.lstm\.(.+)$", k)
        if m:
            out[f"encoder.model.{m.group(1)}.lstm.{m.group(2)}"] = v
            continue
        m = re.match(r"quantizer\.layers\.(\d+)\.codebook\.embed$", k)
        if m:
            out[f"quantizer.vq.layers.{m.group(1)}._codebook.embed"] = v
    return out


def repack_encodec(enc_sd, n_q=8):
    """Encodec 24 kHz encoder + RVQ state dict -> the "enc.*" fp32 tensors csrc/encodec.cu consumes.  Accepts
    EncodecModel.state_dict() of the `encodec` package (key names: oracle/encodec_oracle.py header) or of `transformers`'
    EncodecModel, with the weight norm folded (`.weight`), in torch's old form (`.weight_g` / `.weight_v`) or as a
    parametrization (`.parametrizations.weight.original0` = g, `.original1` = v)."""
    enc_sd = encodec_keys_from_hf(enc_sd)

    def w_of(prefix):
        if prefix + ".weight" in enc_sd:
            return enc_sd[prefix + ".weight"].float()
        if prefix + ".weight_g" in enc_sd:
            g, v = enc_sd[prefix + ".weight_g"].float(), enc_sd[prefix + ".weight_v"].float()   # torch weight_norm, dim=0
        else:
            g, v = (enc_sd[prefix + ".parametrizations.weight.original0"].float(),
                    enc_sd[prefix + ".parametrizations.weight.original1"].float())
        return v * (g.reshape(-1, *([1] * (v.dim() - 1))) / v.flatten(1).norm(dim=1).reshape(-1, *([1] * (v.dim() - 1))))

    t, p = {}, "encoder.model."
    t["enc.c0.w"], t["enc.c0.b"] = w_of(p + "0.conv.conv"), enc_sd[p + "0.conv.conv.bias"].float()
    idx = 1
    for s in range(4):
        for nm, key in (("a", f"{idx}.block.1.conv.conv"), ("b", f"{idx}.block.3.conv.conv"), ("s", f"{idx}.shortcut.conv.conv")):
            t[f"enc.r{s}.{nm}.w"], t[f"enc.r{s}.{nm}.b"] = w_of(p + key), enc_sd[p + key + ".bias"].float()
        t[f"enc.d{s}.w"], t[f"enc.d{s}.b"] = w_of(p + f"{idx + 2}.conv.conv"), enc_sd[p + f"{idx + 2}.conv.conv.bias"].float()
        idx += 3
    for l in range(2):
        t[f"enc.lstm{l}.ih.w"], t[f"enc.lstm{l}.ih.b"] = enc_sd[p + f"{idx}.lstm.weight_ih_l{l}"].float(), enc_sd[p + f"{idx}.lstm.bias_ih_l{l}"].float()
        t[f"enc.lstm{l}.hh.w"], t[f"enc.lstm{l}.hh.b"] = enc_sd[p + f"{idx}.lstm.weight_hh_l{l}"].float(), enc_sd[p + f"{idx}.lstm.bias_hh_l{l}"].float()
    t["enc.final.w"], t["enc.final.b"] = w_of(p + f"{idx + 2}.conv.conv"), enc_sd[p + f"{idx + 2}.conv.conv.bias"].float()
    for q in range(n_q):
        t[f"enc.cb{q}"] = enc_sd[f"quantizer.vq.layers.{q}._codebook.embed"].float()
    return {k: v.contiguous() for k, v in t.items()}


def make_cfg(dims, alphas, max_pos):
    cfg = capi.ModelCfg()
    for k, v in dims.items():
        setattr(cfg, k, int(v))
    for k, v in alphas.items():
        setattr(cfg, k, float(v))
    cfg.ar_norm_eps = 1e-5     # ModelArgs.norm_eps (nn_future.py:154)
    cfg.ln_eps = 4e-5          # LAYERNORM_EPS (model.py:13)
    cfg.head_ln_eps = 1e-5     # nn.LayerNorm default in residual_decoder (model.py:237)
    cfg.max_pos = max_pos
    return cfg

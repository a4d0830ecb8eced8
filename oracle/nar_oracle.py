"""TEST INFRASTRUCTURE ONLY -- CPU fp32 restatement of the reference NAR path (never imported by the product package).

Restates over a reference-format ``state_dict``:
  * ResidualTransformer.forward          /root/reference/mars5/model.py:264-343
  * timestep_embedding                   /root/reference/mars5/model.py:18-35
  * nn.Transformer (pre-LN encoder/decoder layers with FNNSwiGLU, linear1 = Identity) model.py:179-204
  * MultinomialDiffusion tables + q_pred / q_pred_one_timestep / q_posterior / log_sample_categorical / q_sample
                                          /root/reference/mars5/diffuser.py:62-236
  * reverse_diffusion / perform_simple_inference / get_schedule   /root/reference/mars5/diffuser.py:318-472

Pinned against the reference by tests/golden/make_golden.py; fixtures replayed by tests/test_oracle_golden.py.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from .ar_oracle import LAYERNORM_EPS, chunked_embed, encoder_layer, mha, sine_pe, speaker_vector, swiglu_ffn

MIN_LOG_ARG = 1e-7  # diffuser.py:18


def timestep_embedding(t, dim, max_period=10000):
    """model.py:18-35 for a scalar timestep."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(start=0, end=half) / half)
    args = torch.tensor([float(t)])[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)[0]


def decoder_layer(sd, p, x, mem, nhead, eps=LAYERNORM_EPS):
    """nn.TransformerDecoderLayer(norm_first=True): self-attn, cross-attn, feed-forward, each pre-normed."""
    D = x.shape[-1]
    h = F.layer_norm(x, (D,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], eps)
    x = x + mha(h, h, sd[p + "self_attn.in_proj_weight"], sd[p + "self_attn.in_proj_bias"],
                sd[p + "self_attn.out_proj.weight"], sd[p + "self_attn.out_proj.bias"], nhead)
    h = F.layer_norm(x, (D,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], eps)
    x = x + mha(h, mem, sd[p + "multihead_attn.in_proj_weight"], sd[p + "multihead_attn.in_proj_bias"],
                sd[p + "multihead_attn.out_proj.weight"], sd[p + "multihead_attn.out_proj.bias"], nhead)
    h = F.layer_norm(x, (D,), sd[p + "norm3.weight"], sd[p + "norm3.bias"], eps)
    return x + swiglu_ffn(sd, p, h)


def nar_forward(sd, cfg, c_text, c_codes, x, t, drop_cond=False, collect=None):
    """ResidualTransformer.forward for one utterance (bs=1): returns logits (S, n_quant, K) -- i.e. the reference's
    (1, S, K, 8) output permuted (0,1,3,2) as reverse_diffusion consumes it (diffuser.py:359)."""
    D, H, Q = cfg["nar_dim"], cfg["nar_heads"], cfg["n_quant"]
    text = sd["text_embed.weight"][c_text]
    if drop_cond:  # model.py:291-296: codes := pad, length := 0 -> only the identity token is visible
        codes = torch.full_like(c_codes, 1024)
        key_len = 1
    else:
        codes, key_len = c_codes, None
    spk = speaker_vector(sd, codes, H, cfg["nar_spk_layers"], "ref_embedder", "ref_pos_embedding.alpha", key_len=key_len)
    t_emb = timestep_embedding(t, D)

    def mlp(prefix):
        h = F.silu(t_emb @ sd[prefix + ".0.weight"].T + sd[prefix + ".0.bias"])
        return h @ sd[prefix + ".2.weight"].T + sd[prefix + ".2.bias"]

    t_enc, t_dec = mlp("timestep_encoder_emb"), mlp("timestep_decoder_emb")
    c = torch.cat([spk[None], text], dim=0)
    c = c + sd["cond_pos_embedding.alpha"] * sine_pe(c.shape[0], D)
    xe = chunked_embed(sd, "residual_encoder", x, Q)
    xe = xe + sd["pos_embedding.alpha"] * sine_pe(xe.shape[0], D)
    xe = xe + t_dec[None]
    c = c + t_enc[None]
    for l in range(cfg["nar_enc_layers"]):
        c = encoder_layer(sd, f"tfm.encoder.layers.{l}.", c, H)
    mem = F.layer_norm(c, (D,), sd["tfm.encoder.norm.weight"], sd["tfm.encoder.norm.bias"], LAYERNORM_EPS)
    for l in range(cfg["nar_dec_layers"]):
        xe = decoder_layer(sd, f"tfm.decoder.layers.{l}.", xe, mem, H)
        if collect is not None:
            collect.append(xe.clone())
    out = F.layer_norm(xe, (D,), sd["tfm.decoder.norm.weight"], sd["tfm.decoder.norm.bias"], LAYERNORM_EPS)
    heads = []
    for q in range(Q):
        h = F.layer_norm(out, (D,), sd[f"residual_decoder.{q}.0.weight"], sd[f"residual_decoder.{q}.0.bias"], 1e-5)
        heads.append(h @ sd[f"residual_decoder.{q}.1.weight"].T + sd[f"residual_decoder.{q}.1.bias"])
    return torch.stack(heads, dim=1)  # (S, Q, K)


# ------------------------------------------------------------------------------------------------ diffusion
def cosine_alpha_schedule(timesteps, s=0.008):
    """MultinomialDiffusion.cosine_beta_schedule (diffuser.py:97-109); note the sqrt of the clamped ratio."""
    steps = timesteps + 1
    x = torch.linspace(0, timesteps, steps)
    ac = torch.cos(((x / timesteps) + s) / (1 + s) * torch.pi * 0.5) ** 2
    ac = ac / ac[0]
    alphas = torch.clamp(ac[1:] / ac[:-1], 0.001, 1.0)
    return torch.sqrt(alphas)


def diffusion_tables(T):
    """MultinomialDiffusion.__init__ (diffuser.py:76-95): fp64 log tables stored as fp32.
    Returns (log_alpha, log_1_min_alpha, log_cumprod_alpha, log_1_min_cumprod_alpha), each (T,) float32."""
    alphas = cosine_alpha_schedule(T).to(torch.float64)
    log_alpha = alphas.log()
    log_cum = torch.cumsum(log_alpha, dim=-1)

    def l1m(a):
        return torch.log((1 - a.exp()).clamp_(min=1e-30))

    return tuple(v.to(torch.float32) for v in (log_alpha, l1m(log_alpha), log_cum, l1m(log_cum)))


def log_add_exp(a, b):
    m = torch.max(a, b)
    return m + torch.log(torch.exp(a - m) + torch.exp(b - m))


def log_onehot(x, K):
    return torch.log(F.one_hot(x, K).float().clamp(min=MIN_LOG_ARG))


def gumbel_argmax(logp, u):
    """log_sample_categorical with the uniforms injected (diffuser.py:219-228)."""
    g = -torch.log((-torch.log(u.clamp(min=MIN_LOG_ARG))).clamp(min=MIN_LOG_ARG))
    return (g + logp).argmax(dim=-1)


def posterior_logprobs(tabs, log_x0, x_t, t, K):
    """p_pred -> q_posterior (diffuser.py:176-217) for integer t shared by all rows."""
    la, l1ma, lc, l1mc = tabs
    lnK = np.log(K)
    if t == 0:
        ev = log_x0
    else:
        ev = log_add_exp(log_x0 + lc[t - 1], l1mc[t - 1] - lnK)
    one = log_add_exp(log_onehot(x_t, K) + la[t], l1ma[t] - lnK)
    un = ev + one
    return un - torch.logsumexp(un, dim=-1, keepdim=True)


def reverse_step(tabs, cond, uncond, x, x_known, m, t, guidance_w, temp, u_unknown, u_known, K):
    """reverse_diffusion (diffuser.py:345-394) given the model outputs cond/uncond (S, Q, K)."""
    x0 = cond
    if guidance_w != 1:
        x0 = guidance_w * cond + (1 - guidance_w) * uncond
    x0 = x0 / temp
    log_x0 = F.log_softmax(x0, dim=-1)
    logp = posterior_logprobs(tabs, log_x0, x, t, K)
    x_unknown = gumbel_argmax(logp, u_unknown)
    if t == 0:
        x_kn = x_known
    else:
        la, l1ma, lc, l1mc = tabs
        q = log_add_exp(log_onehot(x_known, K) + lc[t], l1mc[t] - np.log(K))  # q_sample -> q_pred (diffuser.py:161-174)
        x_kn = gumbel_argmax(q, u_known)
    return x_kn * m.long() + x_unknown * (1 - m.long()), logp


def get_schedule(t_T, jump_len=10, jump_n_sample=10):
    """RePaint resampling schedule (diffuser.py:318-333)."""
    jumps = {j: jump_n_sample - 1 for j in range(0, t_T - jump_len, jump_len)}
    t, ts = t_T, []
    while t >= 1:
        t -= 1
        ts.append(t)
        if jumps.get(t, 0) > 0:
            jumps[t] -= 1
            for _ in range(jump_len):
                t += 1
                ts.append(t)
    ts.append(-1)
    return ts


def forward_step(tabs, x, t, u, K):
    """forward_diffusion with c=None (diffuser.py:336-342): x_{t+1} ~ q_pred_one_timestep(onehot(x_t), t), every entry."""
    la, l1ma, _, _ = tabs
    return gumbel_argmax(log_add_exp(log_onehot(x, K) + la[t], l1ma[t] - np.log(K)), u)


def nar_infer(sd, cfg, c_text, c_codes, x_l0, ncfg, x_init, noise):
    """perform_simple_inference (diffuser.py:398-472) for one utterance.
    x_init: (N, Q) initial randint draw; noise: (n_steps, 2, S, Q, K) uniforms, n_steps = len(get_schedule) - 1 (= T
    without jumps): a reverse step uses draw 0 (unknown sample) and draw 1 (known re-noise), a forward step draw 0.
    ncfg may carry jump_len / jump_n_sample (RePaint; unscaled forward diffusion, enable_kevin_scaled_inference=False).
    Returns codes (N, Q) after the deep-clone crop."""
    K, Q, T = cfg["n_classes"], cfg["n_quant"], ncfg["T"]
    tabs = diffusion_tables(T)
    x = x_init.clone()
    x[:, 0] = x_l0
    x_known = torch.zeros_like(x)
    x_known[:, 0] = x[:, 0]
    m = torch.zeros_like(x).bool()
    m[:, 0] = True
    x_q0 = x_l0.clone()
    offset = 0
    if ncfg["deep_clone"]:
        x = torch.cat([c_codes, x], dim=0)
        x_known = torch.cat([c_codes, x_known], dim=0)
        m = torch.cat([torch.ones_like(c_codes).bool(), m], dim=0)
        x_q0 = torch.cat([c_codes[:, 0], x_q0], dim=0)
        offset = c_codes.shape[0]
    times = get_schedule(T, ncfg.get("jump_len", 1), ncfg.get("jump_n_sample", 1))
    for step, (t, t_cur) in enumerate(zip(times[:-1], times[1:])):
        if t_cur < t:
            cond = nar_forward(sd, cfg, c_text, c_codes, x, t, drop_cond=False)
            uncond = nar_forward(sd, cfg, c_text, c_codes, x, t, drop_cond=True) if ncfg["guidance_w"] != 1 else None
            x, _ = reverse_step(tabs, cond, uncond, x, x_known, m, t, ncfg["guidance_w"], ncfg["x0_temp"], noise[step, 0],
                                noise[step, 1], K)
        else:
            x = forward_step(tabs, x, t, noise[step, 0], K)
        if ncfg["q0_override_steps"] < t:
            x[:, 0] = x_q0
    return x[offset:]

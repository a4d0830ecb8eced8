"""TEST INFRASTRUCTURE ONLY -- CPU fp32 restatement of the reference AR path (never imported by the product package).

Restates, as plain functional torch-on-CPU code over a reference-format ``state_dict``:
  * CodecLM.forward                       /root/reference/mars5/model.py:95-141
  * MistralTransformer / TransformerBlock /root/reference/mars5/nn_future.py:315-398
  * Attention (RoPE, causal SDPA)         /root/reference/mars5/nn_future.py:166-198, 235-274
  * RMSNorm / FeedForward                 /root/reference/mars5/nn_future.py:277-312
  * speaker encoder (pre-LN TransformerEncoderLayer with FNNSwiGLU) model.py:56-67,109-127; nn_future.py:13-29,35-83
  * logit warpers + sampling + stop rule  /root/reference/mars5/ar_generate.py:62-165, /root/reference/mars5/samplers.py:20-93

Pinned against the reference itself by tests/golden/make_golden.py (run in the build container, where
/root/reference is importable); the resulting fixtures are replayed by tests/test_oracle_golden.py.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
import math

import torch
import torch.nn.functional as F

LAYERNORM_EPS = 4e-5  # model.py:13


def sine_pe(n, dim):
    """SinePositionalEmbedding.extend_pe (nn_future.py:51-76)."""
    pe = torch.zeros(n, dim)
    position = torch.arange(0, n, dtype=torch.float32).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, dim, 2, dtype=torch.float32) * -(math.log(10000.0) / dim))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe


def chunked_embed(sd, prefix, codes, n_q):
    """ChunkedEmbedding.forward (model.py:152-159). codes (L, n_q)."""
    return torch.cat([sd[f"{prefix}.embs.{i}.weight"][codes[:, i]] for i in range(n_q)], dim=-1)


def mha(x_q, x_kv, w_in, b_in, w_out, b_out, nhead, key_len=None):
    """nn.MultiheadAttention forward for one sequence: x_q (Lq, D), x_kv (Lk, D); keys >= key_len are masked."""
    D = x_q.shape[-1]
    q = x_q @ w_in[:D].T + b_in[:D]
    k = x_kv @ w_in[D:2 * D].T + b_in[D:2 * D]
    v = x_kv @ w_in[2 * D:].T + b_in[2 * D:]
    hd = D // nhead
    q = q.view(-1, nhead, hd).transpose(0, 1)
    k = k.view(-1, nhead, hd).transpose(0, 1)
    v = v.view(-1, nhead, hd).transpose(0, 1)
    s = (q @ k.transpose(1, 2)) / math.sqrt(hd)
    if key_len is not None and key_len < k.shape[1]:
        s[:, :, key_len:] = float("-inf")
    o = (s.softmax(-1) @ v).transpose(0, 1).reshape(-1, D)
    return o @ w_out.T + b_out


def swiglu_ffn(sd, p, x):
    """FNNSwiGLU (nn_future.py:13-29) as the activation of a layer whose linear1 is Identity, then linear2."""
    g = F.silu(x @ sd[p + "activation.W.weight"].T) * (x @ sd[p + "activation.V.weight"].T)
    return g @ sd[p + "linear2.weight"].T + sd[p + "linear2.bias"]


def encoder_layer(sd, p, x, nhead, key_len=None, eps=LAYERNORM_EPS):
    """nn.TransformerEncoderLayer(norm_first=True): x + sa(norm1(x)); x + ff(norm2(x))."""
    D = x.shape[-1]
    h = F.layer_norm(x, (D,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], eps)
    x = x + mha(h, h, sd[p + "self_attn.in_proj_weight"], sd[p + "self_attn.in_proj_bias"],
                sd[p + "self_attn.out_proj.weight"], sd[p + "self_attn.out_proj.bias"], nhead, key_len)
    h = F.layer_norm(x, (D,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], eps)
    return x + swiglu_ffn(sd, p, h)


def speaker_vector(sd, spk_codes, nhead, n_layers, emb_prefix, pos_alpha_key, enc_prefix="spk_encoder", key_len=None):
    """model.py:109-127 (AR) / model.py:298-310 (NAR): identity token + chunked codes -> encoder -> row 0."""
    n_q = spk_codes.shape[1]
    seq = torch.cat([sd["spk_identity_emb.weight"], chunked_embed(sd, emb_prefix, spk_codes, n_q)], dim=0)
    dim = seq.shape[-1]
    seq = seq + sd[pos_alpha_key] * sine_pe(seq.shape[0], dim)
    for l in range(n_layers):
        seq = encoder_layer(sd, f"{enc_prefix}.layers.{l}.", seq, nhead, key_len)
    seq = F.layer_norm(seq, (dim,), sd[f"{enc_prefix}.norm.weight"], sd[f"{enc_prefix}.norm.bias"], LAYERNORM_EPS)
    return seq[0]


def ar_spk_key_len(spk_codes):
    """construct_padding_mask on codebook 0 == 1024 (model.py:119-125, utils.py:41-42)."""
    pad = (spk_codes[:, 0] == 1024).nonzero()
    return 1 + (int(pad[0]) if len(pad) else spk_codes.shape[0])


def rope(x, positions, head_dim=64, theta=10000.0):
    """apply_rotary_emb with precompute_freqs_cis (nn_future.py:181-198). x (L, H, hd)."""
    freqs = 1.0 / (theta ** (torch.arange(0, head_dim, 2)[: head_dim // 2].float() / head_dim))
    ang = torch.outer(positions.to(freqs.dtype) if positions.is_floating_point() else positions, freqs).float()
    fc = torch.polar(torch.ones_like(ang), ang)
    xc = torch.view_as_complex(x.float().reshape(*x.shape[:-1], -1, 2))
    return torch.view_as_real(xc * fc[:, None, :]).flatten(2)


def rmsnorm(x, w, eps):
    return x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps) * w


def mistral_forward(sd, h, nhead, n_layers, eps=1e-5, collect=None):
    """MistralTransformer.forward without cache over one full sequence h (L, D) (nn_future.py:369-398)."""
    L, D = h.shape
    pos = torch.arange(L)
    mask = torch.log(torch.tril(torch.ones(L, L)))  # banded by sliding_window >= L: plain causal (nn_future.py:381-392)
    for l in range(n_layers):
        p = f"ar.layers.{l}."
        x = rmsnorm(h, sd[p + "attention_norm.weight"], eps)
        q = (x @ sd[p + "attention.wq.weight"].T).view(L, nhead, 64)
        k = (x @ sd[p + "attention.wk.weight"].T).view(L, nhead, 64)
        v = (x @ sd[p + "attention.wv.weight"].T).view(L, nhead, 64)
        q, k = rope(q, pos), rope(k, pos)
        s = torch.einsum("qhd,khd->hqk", q, k) / 8.0 + mask
        o = torch.einsum("hqk,khd->qhd", s.softmax(-1), v).reshape(L, D)
        h = h + o @ sd[p + "attention.wo.weight"].T
        x = rmsnorm(h, sd[p + "ffn_norm.weight"], eps)
        f = F.silu(x @ sd[p + "feed_forward.w1.weight"].T) * (x @ sd[p + "feed_forward.w3.weight"].T)
        h = h + f @ sd[p + "feed_forward.w2.weight"].T
        if collect is not None:
            collect.append(h.clone())
    return rmsnorm(h, sd["ar.norm.weight"], eps) @ sd["ar.output.weight"].T


def codeclm_forward(sd, cfg, ids, spk_codes):
    """CodecLM.forward(x, spk_reference=...) without cache (model.py:95-141): logits (len(ids), V)."""
    x = sd["embed.weight"][ids]
    spk = speaker_vector(sd, spk_codes, cfg["ar_heads"], cfg["ar_spk_layers"], "ref_chunked_emb", "pos_embedding.alpha",
                         key_len=ar_spk_key_len(spk_codes))
    h = torch.cat([spk[None], x], dim=0)
    return mistral_forward(sd, h, cfg["ar_heads"], cfg["ar_layers"])[1:]


# ------------------------------------------------------------------------------------------------ sampler
def warp_logits(logits, prev_ids, scfg, text_vocab, eos_idx, n_phones_gen):
    """ar_generate.py:73-98 for one row; returns log-probs (V,). scfg keys mirror InferenceConfig."""
    z = logits.clone().float()
    if len(prev_ids) > 1:  # samplers.py:20-36
        prev = torch.tensor(prev_ids[-scfg["penalty_window"]:], dtype=torch.long)
        vals, cnts = prev.unique(return_counts=True)
        c = torch.zeros_like(z, dtype=torch.long)
        c[vals] = cnts
        z = z - c * scfg["alpha_frequency"] - (c > 0).to(z.dtype) * scfg["alpha_presence"]
    z[: text_vocab - 1] = float("-inf")
    if n_phones_gen is not None and len(prev_ids) <= n_phones_gen:  # samplers.py:39-56
        penalty = max(n_phones_gen - len(prev_ids), 1)
        z[eos_idx] -= scfg["eos_penalty_factor"] * (penalty ** scfg["eos_penalty_decay"])
    z = z / scfg["temperature"]
    if scfg["top_k"] > 0:  # samplers.py:70-74
        k = min(max(scfg["top_k"], 1), z.numel())
        z[z < torch.topk(z, k)[0][-1]] = float("-inf")
    if scfg["top_p"] < 1.0:  # samplers.py:76-91
        sl, si = torch.sort(z, descending=True)
        cp = torch.cumsum(F.softmax(sl, dim=-1), dim=-1)
        rm = cp > scfg["top_p"]
        rm[1:] = rm[:-1].clone()
        rm[0] = False
        z[si[rm]] = float("-inf")
    tp = scfg.get("typical_p", 1.0)
    if tp <= 0.999:  # apply_typical_p (samplers.py:96-122), applied after top-k / top-p (ar_generate.py:93)
        z = typical_p_filter(z, tp)
    z[: text_vocab - 1] = float("-inf")
    return z.log_softmax(-1)


def typical_p_filter(logits, mass):
    """apply_typical_p (samplers.py:96-122) for one row of (already filtered) logits."""
    normalized = logits.log_softmax(-1)
    p = normalized.exp()
    ent = -(normalized * p).nansum(-1, keepdim=True)
    shifted = ((-normalized) - ent).abs()
    sorted_scores, sorted_idx = torch.sort(shifted, descending=False)
    cum = logits[sorted_idx].softmax(-1).cumsum(-1)
    last = int((cum < mass).sum())
    thr = sorted_scores[min(last, len(sorted_scores) - 1)]
    out = logits.clone()
    out[shifted > thr] = float("-inf")
    return out


def sample_token(logprobs, exp_noise):
    """torch.multinomial(p, 1) == argmax(p / e), e ~ Exp(1) (ar_generate.py:114-118)."""
    return int((logprobs.exp() / exp_noise).argmax())


def ar_generate(sd, cfg, prompt, spk_codes, scfg, noise, max_len, n_phones_gen, eos_idx, return_logits=False):
    """ar_generate.py:15-165 for one utterance (beam 1), recomputing the full forward each step (the reference's
    KV cache is numerically equivalent).  noise: (steps, V) Exp(1) draws."""
    ids = [int(t) for t in prompt]
    prev, all_logits = [], []
    step = 0
    hit = False
    while len(ids) < max_len:
        logits = codeclm_forward(sd, cfg, torch.tensor(ids), spk_codes)[-1]
        if return_logits:
            all_logits.append(logits.clone())
        lp = warp_logits(logits, prev, scfg, cfg["ar_text_vocab"], eos_idx, n_phones_gen)
        tok = sample_token(lp, noise[step])
        step += 1
        if tok == eos_idx:
            break
        ids.append(tok)
        prev.append(tok)
    if len(ids) >= max_len - 1:
        hit = True
    out = (torch.tensor(ids), hit)
    return out + (torch.stack(all_logits),) if return_logits else out


# ------------------------------------------------------------------------------------------------ KV-cached stepping
class KVCache:
    """Per-layer K/V of everything fed so far (RotatingBufferCache without the wrap, nn_future.py:89-134)."""

    def __init__(self):
        self.k, self.v, self.n = {}, {}, 0


def codeclm_step(sd, cfg, new_ids, spk_codes, cache):
    """One CodecLM.forward call of the reference's cached loop (ar_generate.py:62-71, model.py:95-141): the speaker
    vector is recomputed on EVERY call exactly like the reference does; on the first call (empty cache) the whole prompt
    plus the speaker slot is processed, afterwards only the newest token.  Returns logits of the last position."""
    nhead, n_layers, eps = cfg["ar_heads"], cfg["ar_layers"], 1e-5
    spk = speaker_vector(sd, spk_codes, nhead, cfg["ar_spk_layers"], "ref_chunked_emb", "pos_embedding.alpha",
                         key_len=ar_spk_key_len(spk_codes))
    x = sd["embed.weight"][new_ids]
    if cache.n == 0:
        h = torch.cat([spk[None], x], dim=0)
    else:
        h = x[-1:]
    L, D = h.shape
    pos = torch.arange(cache.n, cache.n + L)
    for l in range(n_layers):
        p = f"ar.layers.{l}."
        xn = rmsnorm(h, sd[p + "attention_norm.weight"], eps)
        q = rope((xn @ sd[p + "attention.wq.weight"].T).view(L, nhead, 64), pos)
        k = rope((xn @ sd[p + "attention.wk.weight"].T).view(L, nhead, 64), pos)
        v = (xn @ sd[p + "attention.wv.weight"].T).view(L, nhead, 64)
        cache.k[l] = k if cache.n == 0 else torch.cat([cache.k[l], k], dim=0)
        cache.v[l] = v if cache.n == 0 else torch.cat([cache.v[l], v], dim=0)
        s = torch.einsum("qhd,khd->hqk", q, cache.k[l]) / 8.0
        if L > 1:
            s = s + torch.log(torch.tril(torch.ones(L, L)))
        o = torch.einsum("hqk,khd->qhd", s.softmax(-1), cache.v[l]).reshape(L, D)
        h = h + o @ sd[p + "attention.wo.weight"].T
        xn = rmsnorm(h, sd[p + "ffn_norm.weight"], eps)
        h = h + (F.silu(xn @ sd[p + "feed_forward.w1.weight"].T) * (xn @ sd[p + "feed_forward.w3.weight"].T)) @ sd[p + "feed_forward.w2.weight"].T
    cache.n += L
    return rmsnorm(h[-1], sd["ar.norm.weight"], eps) @ sd["ar.output.weight"].T

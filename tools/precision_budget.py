"""CPU experiment behind DESIGN.md section 5: which fp16 roundings of the NAR forward cost how much logit error.

Emulates the GPU data path of csrc/nar.cu on the CPU (fp32 accumulate, operands rounded to fp16 at selectable points)
on the full-size synthetic ResidualTransformer and reports max-abs / rms logit error against the fp32 oracle for a set
of rounding configurations.  "split" = the operand is carried as an fp16 (hi, lo) pair, i.e. effectively unrounded.

    python tools/precision_budget.py [--size full|tiny] [--S 800]
"""
import argparse
import math
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from mars5_tts_b200 import synth, weights  # noqa: E402
from oracle import nar_oracle  # noqa: E402
from oracle.ar_oracle import LAYERNORM_EPS, chunked_embed, sine_pe, speaker_vector  # noqa: E402

POINTS = ("ln_sa", "ln_cq", "ln_ffn", "q", "k", "v", "p", "att_sa", "att_ca", "g", "mem", "head", "kvmem")


def r16(x, on):
    return x.half().float() if on else x


def mha(rc, x_q, x_kv, w_in, b_in, w_out, b_out, nhead, cross):
    D = x_q.shape[-1]
    q = r16(x_q @ w_in[:D].T + b_in[:D], rc["q"])
    k = r16(x_kv @ w_in[D:2 * D].T + b_in[D:2 * D], rc["kvmem"] if cross else rc["k"])
    v = r16(x_kv @ w_in[2 * D:].T + b_in[2 * D:], rc["kvmem"] if cross else rc["v"])
    hd = D // nhead
    q = q.view(-1, nhead, hd).transpose(0, 1)
    k = k.view(-1, nhead, hd).transpose(0, 1)
    v = v.view(-1, nhead, hd).transpose(0, 1)
    s = (q @ k.transpose(1, 2)) / math.sqrt(hd)
    m = s.max(-1, keepdim=True).values
    p = torch.exp(s - m)
    l = p.sum(-1, keepdim=True)           # the kernel sums the unrounded fp32 p
    o = (r16(p, rc["p"]) @ v) / l
    o = r16(o.transpose(0, 1).reshape(-1, D), rc["att_ca"] if cross else rc["att_sa"])
    return o @ w_out.T + b_out


def ffn(rc, sd, p, h):
    g = F.silu(h @ sd[p + "activation.W.weight"].T) * (h @ sd[p + "activation.V.weight"].T)
    return r16(g, rc["g"]) @ sd[p + "linear2.weight"].T + sd[p + "linear2.bias"]


def enc_layer(rc, sd, p, x, H):
    D = x.shape[-1]
    h = r16(F.layer_norm(x, (D,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], LAYERNORM_EPS), rc["ln_sa"])
    x = x + mha(rc, h, h, sd[p + "self_attn.in_proj_weight"], sd[p + "self_attn.in_proj_bias"],
                sd[p + "self_attn.out_proj.weight"], sd[p + "self_attn.out_proj.bias"], H, False)
    h = r16(F.layer_norm(x, (D,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], LAYERNORM_EPS), rc["ln_ffn"])
    return x + ffn(rc, sd, p, h)


def dec_layer(rc, sd, p, x, mem, H):
    D = x.shape[-1]
    h = r16(F.layer_norm(x, (D,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], LAYERNORM_EPS), rc["ln_sa"])
    x = x + mha(rc, h, h, sd[p + "self_attn.in_proj_weight"], sd[p + "self_attn.in_proj_bias"],
                sd[p + "self_attn.out_proj.weight"], sd[p + "self_attn.out_proj.bias"], H, False)
    h = r16(F.layer_norm(x, (D,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], LAYERNORM_EPS), rc["ln_cq"])
    x = x + mha(rc, h, mem, sd[p + "multihead_attn.in_proj_weight"], sd[p + "multihead_attn.in_proj_bias"],
                sd[p + "multihead_attn.out_proj.weight"], sd[p + "multihead_attn.out_proj.bias"], H, True)
    h = r16(F.layer_norm(x, (D,), sd[p + "norm3.weight"], sd[p + "norm3.bias"], LAYERNORM_EPS), rc["ln_ffn"])
    return x + ffn(rc, sd, p, h)


def forward(rc, sd, cfg, spk, c_text, x, t):
    D, H, Q = cfg["nar_dim"], cfg["nar_heads"], cfg["n_quant"]
    t_emb = nar_oracle.timestep_embedding(t, D)

    def mlp(prefix):
        h = F.silu(t_emb @ sd[prefix + ".0.weight"].T + sd[prefix + ".0.bias"])
        return h @ sd[prefix + ".2.weight"].T + sd[prefix + ".2.bias"]

    c = torch.cat([spk[None], sd["text_embed.weight"][c_text]], dim=0)
    c = c + sd["cond_pos_embedding.alpha"] * sine_pe(c.shape[0], D) + mlp("timestep_encoder_emb")[None]
    xe = chunked_embed(sd, "residual_encoder", x, Q)
    xe = xe + sd["pos_embedding.alpha"] * sine_pe(xe.shape[0], D) + mlp("timestep_decoder_emb")[None]
    for l in range(cfg["nar_enc_layers"]):
        c = enc_layer(rc, sd, f"tfm.encoder.layers.{l}.", c, H)
    mem = r16(F.layer_norm(c, (D,), sd["tfm.encoder.norm.weight"], sd["tfm.encoder.norm.bias"], LAYERNORM_EPS), rc["mem"])
    for l in range(cfg["nar_dec_layers"]):
        xe = dec_layer(rc, sd, f"tfm.decoder.layers.{l}.", xe, mem, H)
    out = F.layer_norm(xe, (D,), sd["tfm.decoder.norm.weight"], sd["tfm.decoder.norm.bias"], LAYERNORM_EPS)
    heads = []
    for q in range(1, Q):
        h = r16(F.layer_norm(out, (D,), sd[f"residual_decoder.{q}.0.weight"], sd[f"residual_decoder.{q}.0.bias"], 1e-5), rc["head"])
        heads.append(h @ sd[f"residual_decoder.{q}.1.weight"].T + sd[f"residual_decoder.{q}.1.bias"])
    return torch.stack(heads, dim=1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", default="full")
    ap.add_argument("--S", type=int, default=800)
    ap.add_argument("--configs", default="")
    args = ap.parse_args()
    torch.set_grad_enabled(False)
    size = {"full": synth.FULL, "tiny": synth.TINY, "mid": synth.MID}[args.size]
    sd = synth.make_nar_state(size)
    cfg = weights.dims_from_state(synth.make_ar_state(dict(size, ar_layers=1, ar_spk_layers=1)), sd, None, size["n_text"])
    g = torch.Generator().manual_seed(5)
    Pf, Tc = (450, 136) if args.size == "full" else (12, 10)
    spk_codes = torch.randint(0, 1024, (Pf, 8), generator=g)
    c_text = torch.randint(0, size["n_text"], (Tc,), generator=g)
    x = torch.randint(0, 1025, (args.S, 8), generator=g)
    spk = speaker_vector(sd, spk_codes, cfg["nar_heads"], cfg["nar_spk_layers"], "ref_embedder", "ref_pos_embedding.alpha")
    none = {k: 0 for k in POINTS}
    t0 = time.time()
    ref = forward(none, sd, cfg, spk, c_text, x, 100)
    print(f"fp32 forward {time.time() - t0:.1f}s  max|logit| {ref.abs().max():.3f} rms {ref.pow(2).mean().sqrt():.3f}")
    allon = {k: 1 for k in POINTS}
    named = {"fast (all fp16)": allon}
    for k in POINTS:
        named["only " + k] = dict(none, **{k: 1})
    named["mixed A: split ln_ffn, att_*, g, head, v"] = dict(allon, ln_ffn=0, att_sa=0, att_ca=0, g=0, head=0, v=0)
    named["mixed B: split ln_ffn, att_*, g, head"] = dict(allon, ln_ffn=0, att_sa=0, att_ca=0, g=0, head=0)
    named["mixed C: split all GEMM A operands (ln_*, att_*, g, mem, head)"] = dict(none, q=1, k=1, v=1, p=1, kvmem=1)
    named["mixed D: C + v split"] = dict(none, q=1, k=1, p=1, kvmem=1)
    named["mixed F: only q, p fp16 (GEMM A operands, K, V split)"] = dict(none, q=1, p=1)
    named["mixed E: B + ln_sa split"] = dict(allon, ln_ffn=0, ln_sa=0, att_sa=0, att_ca=0, g=0, head=0)
    sel = [s.strip() for s in args.configs.split(";") if s.strip()]
    for name, rc in named.items():
        if sel and name not in sel:
            continue
        t0 = time.time()
        out = forward(rc, sd, cfg, spk, c_text, x, 100)
        d = (out - ref).abs()
        print(f"{name:52s} max-abs {d.max():.2e}  rms {d.pow(2).mean().sqrt():.2e}  ({time.time() - t0:.1f}s)")


if __name__ == "__main__":
    main()

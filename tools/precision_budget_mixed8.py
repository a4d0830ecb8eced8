"""CPU emulation of the `mixed8` NAR numerics on the full-size synthetic model: every big decoder GEMM as
fp16 hi halves x fp16 weights + e5m2 lo halves (x 2^-2) x e4m3 weights (x 2^+2), queries and probabilities single fp16, everything
else fp32.  Prints the max-abs / rms logit error against the fp32 forward (round 2: 4.4e-4 / 8.9e-5 at S = 800; the GPU measured
5.3e-4 at S = 1650).  M5_PB_KSINGLE=1 / M5_PB_VSINGLE=1 additionally round the decoder self-attention keys / values to single fp16
(`mixed8k` = keys: 4.65e-4 / 8.90e-5, i.e. free; values too: 5.16e-4).   python tools/precision_budget_mixed8.py"""
import os, sys, math, torch, torch.nn.functional as F
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import precision_budget as pb
from mars5_tts_b200 import synth, weights
from oracle.ar_oracle import LAYERNORM_EPS, speaker_vector
torch.set_grad_enabled(False)
W8C = {}
def lin8(a, W, b=None, f8=True):
    hi = a.half().float(); lo = a - hi
    if f8:
        key = id(W)
        if key not in W8C: W8C[key] = (W*4).to(torch.float8_e4m3fn).float()/4
        lo = (lo*0.25).to(torch.float8_e5m2).float()*4
        out = hi @ W.T + lo @ W8C[key].T
    else:
        out = hi @ W.T + lo.half().float() @ W.T
    return out if b is None else out + b
def mha8(x_q, x_kv, w_in, b_in, w_out, b_out, H, cross):
    D = x_q.shape[-1]
    if cross:
        q = (x_q.half().float() @ w_in[:D].T + b_in[:D]).half().float()        # hi halves only
        k = lin8(x_kv, w_in[D:2*D], b_in[D:2*D], f8=False); v = lin8(x_kv, w_in[2*D:], b_in[2*D:], f8=False)
    else:
        q = lin8(x_q, w_in[:D], b_in[:D]).half().float(); k = lin8(x_kv, w_in[D:2*D], b_in[D:2*D]); v = lin8(x_kv, w_in[2*D:], b_in[2*D:])
        if os.environ.get("M5_PB_KSINGLE"): k = k.half().float()   # mixed8k: single-fp16 keys in the decoder self-attention
        if os.environ.get("M5_PB_VSINGLE"): v = v.half().float()
    hd = D // H
    q,k,v = [t.view(-1,H,hd).transpose(0,1) for t in (q,k,v)]
    s = (q @ k.transpose(1,2)) / math.sqrt(hd)
    m = s.max(-1, keepdim=True).values; p = torch.exp(s-m); l = p.sum(-1, keepdim=True)
    o = ((p.half().float() @ v) / l).transpose(0,1).reshape(-1, D)
    return lin8(o, w_out, b_out)
def ffn8(sd, p, h):
    a = lin8(h, sd[p+"activation.W.weight"]); c = lin8(h, sd[p+"activation.V.weight"])
    return lin8(F.silu(a)*c, sd[p+"linear2.weight"], sd[p+"linear2.bias"])
def dec8(sd, p, x, mem, H):
    D = x.shape[-1]
    ln = lambda n: F.layer_norm(x, (D,), sd[p+n+".weight"], sd[p+n+".bias"], LAYERNORM_EPS)
    x = x + mha8(ln("norm1"), ln("norm1"), sd[p+"self_attn.in_proj_weight"], sd[p+"self_attn.in_proj_bias"], sd[p+"self_attn.out_proj.weight"], sd[p+"self_attn.out_proj.bias"], H, False)
    ln = lambda n: F.layer_norm(x, (D,), sd[p+n+".weight"], sd[p+n+".bias"], LAYERNORM_EPS)
    x = x + mha8(ln("norm2"), mem, sd[p+"multihead_attn.in_proj_weight"], sd[p+"multihead_attn.in_proj_bias"], sd[p+"multihead_attn.out_proj.weight"], sd[p+"multihead_attn.out_proj.bias"], H, True)
    ln = lambda n: F.layer_norm(x, (D,), sd[p+n+".weight"], sd[p+n+".bias"], LAYERNORM_EPS)
    return x + ffn8(sd, p, ln("norm3"))
pb.dec_layer = lambda rc, sd, p, x, mem, H: dec8(sd, p, x, mem, H)
size = synth.FULL
sd = synth.make_nar_state(size)
cfg = weights.dims_from_state(synth.make_ar_state(dict(size, ar_layers=1, ar_spk_layers=1)), sd, None, size["n_text"])
g = torch.Generator().manual_seed(5)
spk_codes = torch.randint(0,1024,(450,8),generator=g); c_text = torch.randint(0,size["n_text"],(136,),generator=g); x = torch.randint(0,1025,(800,8),generator=g)
spk = speaker_vector(sd, spk_codes, cfg["nar_heads"], cfg["nar_spk_layers"], "ref_embedder", "ref_pos_embedding.alpha")
none = {k:0 for k in pb.POINTS}
orig = pb.dec_layer
out8 = pb.forward(none, sd, cfg, spk, c_text, x, 100)
import importlib; importlib.reload(pb)
ref = pb.forward(none, sd, cfg, spk, c_text, x, 100)
d = (out8-ref).abs(); print(f"mixed8 emulation (decoder GEMMs fp16 hi + e5m2 lo x e4m3 W, q/p fp16): max-abs {d.max():.2e} rms {d.pow(2).mean().sqrt():.2e}")

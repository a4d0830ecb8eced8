"""TEST INFRASTRUCTURE ONLY -- CPU fp32 restatement of the Encodec 24 kHz ENCODER + residual vector quantiser that
Mars5TTS runs on the reference clip before the loops (/root/reference/inference.py:87-88, 233:
``self.codec.encode(ref_audio[None])`` with ``EncodecModel.encodec_model_24khz()`` at 6 kbps -> (1, 8, T) codes).

PARITY PINNED AGAINST AN INDEPENDENT IMPLEMENTATION, UNPINNED AGAINST THE RELEASED WEIGHTS: the algorithm lives in the
third-party package ``encodec`` (requirements.txt, unpinned; 0.1.1 at the time of the reference) whose source and weights are
absent from /root/reference and from this image.  ``transformers`` (installed) ships an independent port of the same model
(EncodecModel; EncodecConfig's defaults are facebook/encodec_24khz): tests/test_encodec_hf_cpu.py loads the synthetic state dict
into it and requires identical codes and embeddings to 1e-6 on clips of 1 ... 4801 samples -- which is how the CAUSAL padding of
the 24 kHz model was found (round 2; the first restatement used the 48 kHz model's split padding).  This file restates
the published encodec 0.1.1 algorithm:
  * encodec/modules/conv.py   SConv1d with ``causal=True`` (EncodecModel.encodec_model_24khz -> _get_model(..., causal=True)),
                              ``pad_mode='reflect'``: total padding = effective_kernel - stride, ALL of it on the left, plus the
                              extra right padding that makes the last window full; pad1d's zero-extension for inputs not longer
                              than the reflect pad
  * encodec/modules/seanet.py SEANetEncoder(channels=1, dimension=128, n_filters=32, n_residual_layers=1, ratios=[8,5,4,2]
                              (applied reversed: 2,4,5,8), ELU(alpha=1), kernel_size=7, residual_kernel_size=3,
                              last_kernel_size=7, dilation_base=2, compress=2, true_skip=False, lstm=2)
  * encodec/modules/lstm.py   SLSTM: 2-layer nn.LSTM over time with a skip connection (y = lstm(x) + x)
  * encodec/quantization/core_vq.py  EuclideanCodebook.quantize (arg-max of -(|x|^2 - 2 x.e + |e|^2)),
                              ResidualVectorQuantization.encode (first n_q = 8 of the 32 layers at 6 kbps / 75 Hz)
State-dict keys follow encodec's module names after weight norm has been folded (what the reference's
``nuke_weight_norm`` leaves): ``encoder.model.{i}.conv.conv.{weight,bias}``, ``encoder.model.{i}.block.{1,3}.conv.conv.*``,
``encoder.model.{i}.shortcut.conv.conv.*``, ``encoder.model.13.lstm.{weight_ih,weight_hh,bias_ih,bias_hh}_l{0,1}``,
``quantizer.vq.layers.{q}._codebook.embed``.  tests/golden/make_encodec_golden.py produces the pinning fixture from the
real package wherever it is installed.
"""
import math

import torch
import torch.nn.functional as F

RATIOS = (2, 4, 5, 8)  # reversed([8, 5, 4, 2])


def _extra_padding(length, kernel_size, stride, padding_total):
    n_frames = (length - kernel_size + padding_total) / stride + 1
    ideal = (math.ceil(n_frames) - 1) * stride + (kernel_size - padding_total)
    return ideal - length


def _pad1d_reflect(x, left, right):
    length = x.shape[-1]
    max_pad = max(left, right)
    extra = 0
    if length <= max_pad:
        extra = max_pad - length + 1
        x = F.pad(x, (0, extra))
    padded = F.pad(x, (left, right), mode="reflect")
    return padded[..., : padded.shape[-1] - extra]


def _sconv1d(x, w, b, stride=1, dilation=1, causal=True):
    """x (1, C_in, L) -> (1, C_out, ceil(L / stride)).  encodec_model_24khz is the CAUSAL model (`causal=True` in
    EncodecModel._get_model): the whole padding_total goes to the left, only the extra padding that completes the last window to
    the right.  (causal=False = the split padding of the 48 kHz model, kept for the cross-check against the HF implementation.)"""
    k = (w.shape[-1] - 1) * dilation + 1
    total = k - stride
    extra = _extra_padding(x.shape[-1], k, stride, total)
    if causal:
        left, right = total, 0
    else:
        right = total // 2
        left = total - right
    return F.conv1d(_pad1d_reflect(x, left, right + extra), w, b, stride=stride, dilation=dilation)


def encoder_forward(sd, wav, causal=True):
    """wav (L,) float -> embeddings (128, T), T = ceil(L / 320)."""
    p = "encoder.model."
    x = wav[None, None].float()
    sconv1d = lambda *a, **k: _sconv1d(*a, causal=causal, **k)
    x = sconv1d(x, sd[p + "0.conv.conv.weight"], sd[p + "0.conv.conv.bias"])
    idx = 1
    for ratio in RATIOS:
        r = p + f"{idx}."
        y = sconv1d(F.elu(x), sd[r + "block.1.conv.conv.weight"], sd[r + "block.1.conv.conv.bias"])
        y = sconv1d(F.elu(y), sd[r + "block.3.conv.conv.weight"], sd[r + "block.3.conv.conv.bias"])
        x = sconv1d(x, sd[r + "shortcut.conv.conv.weight"], sd[r + "shortcut.conv.conv.bias"]) + y
        d = p + f"{idx + 2}."
        x = sconv1d(F.elu(x), sd[d + "conv.conv.weight"], sd[d + "conv.conv.bias"], stride=ratio)
        idx += 3
    # SLSTM: (B, C, T) -> (T, B, C) -> 2-layer LSTM -> + skip
    seq = x[0].T  # (T, C)
    h = seq
    for layer in range(2):
        w_ih, w_hh = sd[p + f"{idx}.lstm.weight_ih_l{layer}"], sd[p + f"{idx}.lstm.weight_hh_l{layer}"]
        b_ih, b_hh = sd[p + f"{idx}.lstm.bias_ih_l{layer}"], sd[p + f"{idx}.lstm.bias_hh_l{layer}"]
        H = w_hh.shape[1]
        ht, ct, outs = torch.zeros(H), torch.zeros(H), []
        pre = h @ w_ih.T + b_ih
        for t in range(h.shape[0]):
            g = pre[t] + w_hh @ ht + b_hh
            i, f, gg, o = g[:H].sigmoid(), g[H:2 * H].sigmoid(), g[2 * H:3 * H].tanh(), g[3 * H:].sigmoid()
            ct = f * ct + i * gg
            ht = o * ct.tanh()
            outs.append(ht)
        h = torch.stack(outs)
    x = (h + seq).T[None]
    x = sconv1d(F.elu(x), sd[p + f"{idx + 2}.conv.conv.weight"], sd[p + f"{idx + 2}.conv.conv.bias"])
    return x[0]


def rvq_encode(sd, emb, n_q=8):
    """emb (128, T) -> codes (T, n_q): greedy residual nearest-neighbour search (first maximal index on ties)."""
    residual = emb.T.clone()  # (T, 128)
    codes = []
    for q in range(n_q):
        e = sd[f"quantizer.vq.layers.{q}._codebook.embed"]  # (1024, 128)
        dist = -(residual.pow(2).sum(1, keepdim=True) - 2 * residual @ e.T + e.pow(2).sum(1)[None])
        ind = dist.max(dim=-1).indices
        residual = residual - e[ind]
        codes.append(ind)
    return torch.stack(codes, dim=1)


def encode(sd, wav, n_q=8, causal=True):
    """EncodecModel.encode on one mono clip (no normalisation, no segmenting: encodec_model_24khz defaults)."""
    return rvq_encode(sd, encoder_forward(sd, wav, causal), n_q)


sconv1d = _sconv1d

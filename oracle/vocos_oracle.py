"""TEST INFRASTRUCTURE ONLY -- CPU fp32 restatement of the Vocos encodec-24khz vocoder called by Mars5TTS.vocode
(/root/reference/inference.py:119,160-172).

PARITY UNPINNED: the algorithm lives in the third-party package ``vocos`` (requirements.txt:7, unpinned; 0.1.0 resolved
in the reference's notebook log) whose source and weights (``charactr/vocos-encodec-24khz``) are absent from
/root/reference and from this image.  This file restates the published vocos 0.1.0 algorithm
(vocos/pretrained.py codes_to_features + decode, vocos/models.py VocosBackbone, vocos/modules.py ConvNeXtBlock +
AdaLayerNorm, vocos/heads.py ISTFTHead, vocos/spectral_ops.py ISTFT with "same" padding) as summarised in SURVEY.md
Appendix C; it could not be checked against the real package.  Its iSTFT head is pinned against torch.istft on the span the
"same" and "center" paddings share (tests/test_vocos_istft_cpu.py).
State-dict keys follow vocos' module names so that a real checkpoint could be dropped in.
"""
import torch
import torch.nn.functional as F


def vocos_forward(sd, codes, bandwidth_id, n_fft=1280, hop=320, return_spec=False):
    """codes (N, 8) -> waveform (320 * N,)."""
    N, Q = codes.shape
    cb = sd["feature_extractor.codebook_weights"]  # (n_codebooks_max * 1024, 128): 16 x 1024 rows in the released checkpoint
    bins = 1024  # encodec.quantizer.bins -- vocos offsets codebook q by q * bins regardless of how many codebooks the table holds
    feats = sum(cb[codes[:, q] + q * bins] for q in range(Q))  # (N, 128)
    x = feats.T[None]  # (1, 128, N)
    x = F.conv1d(x, sd["backbone.embed.weight"], sd["backbone.embed.bias"], padding=3)

    def adanorm(x_t, prefix):  # x_t (1, N, C): AdaLayerNorm = LN(no affine, eps 1e-6) * scale[id] + shift[id]
        h = F.layer_norm(x_t, (x_t.shape[-1],), eps=1e-6)
        return h * sd[prefix + ".scale.weight"][bandwidth_id] + sd[prefix + ".shift.weight"][bandwidth_id]

    x = adanorm(x.transpose(1, 2), "backbone.norm").transpose(1, 2)
    n_layers = len([k for k in sd if k.startswith("backbone.convnext.") and k.endswith(".gamma")])
    for l in range(n_layers):
        p = f"backbone.convnext.{l}."
        res = x
        h = F.conv1d(x, sd[p + "dwconv.weight"], sd[p + "dwconv.bias"], padding=3, groups=x.shape[1])
        h = adanorm(h.transpose(1, 2), p + "norm")
        h = F.gelu(h @ sd[p + "pwconv1.weight"].T + sd[p + "pwconv1.bias"])
        h = h @ sd[p + "pwconv2.weight"].T + sd[p + "pwconv2.bias"]
        h = sd[p + "gamma"] * h
        x = res + h.transpose(1, 2)
    x = F.layer_norm(x.transpose(1, 2), (x.shape[1],), sd["backbone.final_layer_norm.weight"],
                     sd["backbone.final_layer_norm.bias"], eps=1e-6)  # (1, N, C)
    spec = x @ sd["head.out.weight"].T + sd["head.out.bias"]  # (1, N, n_fft + 2)
    wav = istft_head(spec[0], n_fft, hop)
    return (wav, spec[0]) if return_spec else wav


def istft_head(spec, n_fft=1280, hop=320):
    """ISTFTHead + ISTFT(padding="same"): spec (N, n_fft+2) = [log-magnitude | phase] -> (hop*N,)."""
    N = spec.shape[0]
    mag, ph = spec.T.chunk(2, dim=0)  # (n_fft/2+1, N) each
    mag = torch.clip(torch.exp(mag), max=1e2)
    S = mag * (torch.cos(ph) + 1j * torch.sin(ph))
    window = torch.hann_window(n_fft)
    pad = (n_fft - hop) // 2
    ifft = torch.fft.irfft(S[None], n_fft, dim=1, norm="backward") * window[None, :, None]  # (1, n_fft, N)
    out_size = (N - 1) * hop + n_fft
    y = F.fold(ifft, output_size=(1, out_size), kernel_size=(1, n_fft), stride=(1, hop))[0, 0, 0, pad:-pad]
    wsq = window.square().expand(1, N, -1).transpose(1, 2)
    env = F.fold(wsq, output_size=(1, out_size), kernel_size=(1, n_fft), stride=(1, hop)).squeeze()[pad:-pad]
    return y / env

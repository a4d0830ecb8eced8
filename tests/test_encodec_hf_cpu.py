"""CPU suite: pins oracle/encodec_oracle.py (the checker of tests/test_*encodec_gpu.py) against an INDEPENDENT implementation of
the Encodec 24 kHz encoder + RVQ -- `transformers`' EncodecModel, whose EncodecConfig defaults are facebook/encodec_24khz
(causal reflect-padded SEANet, 2-layer LSTM with skip, 32 x 1024 x 128 residual codebooks, 8 of them at 6 kbps).  The synthetic
reference-format state dict (encodec's own key names, weight norm folded) is loaded into the HF modules; codes must be identical
and the embeddings equal to 1e-5 on clips from 1 sample (zero-extended reflect padding) to several frames.  This cross-check
found the causal padding of the 24 kHz model (the first restatement had used the 48 kHz model's split padding).  The released
WEIGHTS remain unpinned (tests/golden/make_encodec_golden.py needs the `encodec` package)."""
import pytest
import torch

from mars5_tts_b200 import synth
from oracle import encodec_oracle as eo

hf = pytest.importorskip("transformers.models.encodec.modeling_encodec")
torch.set_grad_enabled(False)


def load_into_hf(sd, n_filters, dimension, causal):
    from transformers import EncodecConfig, EncodecModel
    m = EncodecModel(EncodecConfig(num_filters=n_filters, hidden_size=dimension, codebook_dim=dimension, use_causal_conv=causal)).eval()

    def put_conv(mod, w, b):
        conv = mod.conv
        if hasattr(conv, "parametrizations"):   # torch.nn.utils.parametrizations.weight_norm: original0 = g, original1 = v
            g, v = conv.parametrizations.weight.original0, conv.parametrizations.weight.original1
        else:
            g, v = conv.weight_g, conv.weight_v
        assert v.shape == w.shape, (v.shape, w.shape)
        v.copy_(w)
        g.copy_(w.flatten(1).norm(dim=1).view_as(g))   # g = |v| per output channel: the folded weight is v itself
        conv.bias.copy_(b)

    p, L = "encoder.model.", m.encoder.layers
    put_conv(L[0], sd[p + "0.conv.conv.weight"], sd[p + "0.conv.conv.bias"])
    for idx in (1, 4, 7, 10):
        put_conv(L[idx].block[1], sd[p + f"{idx}.block.1.conv.conv.weight"], sd[p + f"{idx}.block.1.conv.conv.bias"])
        put_conv(L[idx].block[3], sd[p + f"{idx}.block.3.conv.conv.weight"], sd[p + f"{idx}.block.3.conv.conv.bias"])
        put_conv(L[idx].shortcut, sd[p + f"{idx}.shortcut.conv.conv.weight"], sd[p + f"{idx}.shortcut.conv.conv.bias"])
        put_conv(L[idx + 2], sd[p + f"{idx + 2}.conv.conv.weight"], sd[p + f"{idx + 2}.conv.conv.bias"])
    for layer in range(2):
        for nm in ("weight_ih", "weight_hh", "bias_ih", "bias_hh"):
            getattr(L[13].lstm, f"{nm}_l{layer}").copy_(sd[p + f"13.lstm.{nm}_l{layer}"])
    put_conv(L[15], sd[p + "15.conv.conv.weight"], sd[p + "15.conv.conv.bias"])
    assert len(m.quantizer.layers) == 32
    for q in range(32):
        m.quantizer.layers[q].codebook.embed.copy_(sd[f"quantizer.vq.layers.{q}._codebook.embed"])
    return m


@pytest.mark.parametrize("causal", [True, False])   # True = encodec_model_24khz (what Mars5TTS loads); False = the 48 kHz padding
def test_oracle_equals_hf_encodec(causal):
    nf, dim = 8, 32
    sd = synth.make_encodec_state(seed=5, n_filters=nf, dimension=dim)
    m = load_into_hf(sd, nf, dim, causal)
    for n in (1, 5, 319, 320, 321, 999, 2400, 4801):
        wav = torch.randn(n, generator=torch.Generator().manual_seed(n)) * 0.3
        want = m.encode(wav[None, None], bandwidth=6.0).audio_codes[0, 0].T          # (T, 8)
        got = eo.encode(sd, wav, causal=causal)
        assert got.shape == want.shape == ((n + 319) // 320, 8)
        assert torch.equal(got, want), (causal, n)
        emb_want, emb = m.encoder(wav[None, None])[0], eo.encoder_forward(sd, wav, causal)
        assert (emb - emb_want).abs().max() < 1e-5, (causal, n)


def test_oracle_equals_hf_encodec_at_released_sizes():
    """The real shapes (32 filters, 512-wide LSTM, 128-dim codebooks), one 0.1 s clip."""
    sd = synth.make_encodec_state(seed=3)
    m = load_into_hf(sd, 32, 128, True)
    wav = torch.randn(2400, generator=torch.Generator().manual_seed(9)) * 0.2
    want = m.encode(wav[None, None], bandwidth=6.0).audio_codes[0, 0].T
    assert torch.equal(eo.encode(sd, wav), want)


def test_cuda_conv_padding_arithmetic_equals_oracle():
    """csrc/encodec.cu computes every convolution input through `enc_padded` (index into the virtually padded clip) with
    left = padding_total, right = 0, extra = (Lout - 1) stride + (k - total) - Lin and the zero-extension `ext`.  The same
    arithmetic restated in Python must reproduce the causal SConv1d of the oracle for every kernel / stride of the encoder,
    including clips shorter than the padding."""
    def enc_padded(x, ln, left, ext, j):
        Lp = ln + ext
        s = j - left
        if s < 0:
            s = -s
        elif s >= Lp:
            s = 2 * (Lp - 1) - s
        return x[s] if 0 <= s < ln else 0.0

    def kernel_conv(x, w, b, stride):
        (Ci, Lin), (Co, _, k) = x.shape, w.shape
        Lout = (Lin + stride - 1) // stride
        total = k - stride
        right, left = 0, total
        extra = (Lout - 1) * stride + (k - total) - Lin
        max_pad = max(left, right + extra)
        ext = max_pad - Lin + 1 if Lin <= max_pad else 0
        y = torch.zeros(Co, Lout)
        for t in range(Lout):
            for ci in range(Ci):
                for kk in range(k):
                    pj = t * stride + kk
                    v = enc_padded(x[ci], Lin, left, ext, pj) if pj < Lin + left + right + extra else 0.0
                    y[:, t] += w[:, ci, kk] * v
        return y + b[:, None]

    g = torch.Generator().manual_seed(0)
    for k, stride in [(7, 1), (3, 1), (1, 1), (4, 2), (8, 4), (10, 5), (16, 8)]:
        for Lin in (1, 2, 3, 5, 7, 8, 9, 15, 16, 17, 33):
            x, w, b = torch.randn(2, Lin, generator=g), torch.randn(3, 2, k, generator=g), torch.randn(3, generator=g)
            want = eo.sconv1d(x[None], w, b, stride=stride, causal=True)[0]
            got = kernel_conv(x, w, b, stride)
            assert got.shape == want.shape and (got - want).abs().max() < 1e-4, (k, stride, Lin)


def test_repack_accepts_the_hf_state_dict():
    """weights.repack_encodec on the state dict of `transformers`' EncodecModel (its own key names, weight norm as a
    parametrization) gives the tensors it gives for the equivalent `encodec`-package dict: a user who has facebook/encodec_24khz
    through transformers can pass `encodec_state=EncodecModel.from_pretrained(...).state_dict()` to Mars5TTS."""
    from mars5_tts_b200 import weights
    nf, dim = 8, 32
    sd = synth.make_encodec_state(seed=6, n_filters=nf, dimension=dim)
    m = load_into_hf(sd, nf, dim, True)
    hf_sd = m.state_dict()
    assert any(k.startswith("encoder.layers.") for k in hf_sd) and not any(k.startswith("encoder.model.") for k in hf_sd)
    want, got = weights.repack_encodec(sd), weights.repack_encodec(hf_sd)
    assert want.keys() == got.keys()
    for k in want:
        assert want[k].shape == got[k].shape and (want[k] - got[k]).abs().max() < 1e-6, k
    # and the oracle reads the renamed dict like the original one
    ren = weights.encodec_keys_from_hf(hf_sd)
    folded = {k.replace(".parametrizations.weight.original1", ".weight"): v for k, v in ren.items() if "original0" not in k}
    wav = torch.randn(1000, generator=torch.Generator().manual_seed(2)) * 0.3
    for k in [k for k in folded if k.endswith(".weight") and k.replace(".weight", ".parametrizations.weight.original0") in ren]:
        g, v = ren[k.replace(".weight", ".parametrizations.weight.original0")], folded[k]
        folded[k] = v * (g / v.flatten(1).norm(dim=1).view_as(g))
    assert torch.equal(eo.encode(folded, wav), eo.encode(sd, wav))

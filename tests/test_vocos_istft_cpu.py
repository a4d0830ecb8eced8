"""CPU suite: the iSTFT head of the Vocos restatement (oracle/vocos_oracle.py::istft_head -- ISTFTHead + ISTFT(padding="same") of
vocos 0.1.0: the checker of the device's warp-cooperative inverse DFT + overlap-add kernels) against an INDEPENDENT implementation
of the same transform, torch.istft.  torch.istft only knows `center` padding: it drops n_fft / 2 samples at both ends where "same"
drops (n_fft - hop) / 2, so the two outputs cover the same overlap-added signal shifted by (n_fft / 2 - pad) = 160 samples and must
agree on the whole span torch.istft returns (inverse DFT, periodic Hann window, overlap-add and window-envelope normalisation are
all exercised).  The learned part of the vocoder (ConvNeXt backbone, codebook features) stays unpinned: `vocos` is not installable
here and no other implementation of it is in the image."""
import pytest
import torch

from oracle import vocos_oracle


@pytest.mark.parametrize("N,n_fft,hop,seed", [(9, 1280, 320, 0), (40, 1280, 320, 1), (5, 64, 16, 2), (2, 1280, 320, 3)])
def test_istft_same_padding_equals_torch_istft_on_the_common_span(N, n_fft, hop, seed):
    g = torch.Generator().manual_seed(seed)
    logmag = torch.randn(N, n_fft // 2 + 1, generator=g) * 1.5          # some bins above log(100): the clip is exercised
    logmag[0, :4] = 6.0
    phase = (torch.rand(N, n_fft // 2 + 1, generator=g) * 2 - 1) * 3.14159
    spec = torch.cat([logmag, phase], dim=1)                             # (N, n_fft + 2) = [log-magnitude | phase]
    got = vocos_oracle.istft_head(spec, n_fft, hop)
    assert got.shape == (hop * N,)
    mag = torch.clip(torch.exp(logmag), max=1e2)
    S = (mag * (torch.cos(phase) + 1j * torch.sin(phase))).T             # (bins, N)
    want = torch.istft(S, n_fft=n_fft, hop_length=hop, win_length=n_fft, window=torch.hann_window(n_fft), center=True)
    assert want.shape == ((N - 1) * hop,)
    shift = n_fft // 2 - (n_fft - hop) // 2
    seg = got[shift: shift + want.numel()]
    scale = float(want.abs().max())
    assert (seg - want).abs().max() < 2e-5 * max(1.0, scale), float((seg - want).abs().max())

"""Pins the vocoder oracle against the REAL `vocos` package where it is importable (it is not in the build image:
oracle/vocos_oracle.py says "parity unpinned" until this script has been run somewhere with `pip install vocos`).

    python tests/golden/make_vocos_golden.py            ->  tests/golden/vocos_golden.npz

Builds vocos' own modules (EncodecFeatures is bypassed exactly like Mars5TTS.vocode does: codes_to_features ->
backbone -> head, inference.py:160-172), loads the seeded synthetic state dict of mars5_tts_b200.synth.make_vocos_state
into them with strict=True (so the key names / shapes the engine consumes are pinned too) and stores
Vocos.decode(codes_to_features(codes), bandwidth_id=1) for seeded codes.  tests/test_vocos_golden.py replays the
fixture against the oracle on the CPU and against m5_vocode on the GPU when the file exists, and reports "unpinned"
(skip with that reason) when it does not.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from mars5_tts_b200 import synth  # noqa: E402


def main():
    try:
        from vocos.heads import ISTFTHead
        from vocos.models import VocosBackbone
    except ImportError as e:
        print(f"vocos is not importable here ({e}); the vocoder oracle stays unpinned")
        return 2
    torch.set_grad_enabled(False)
    size = synth.FULL
    sd = synth.make_vocos_state(size)
    backbone = VocosBackbone(input_channels=size["voc_feat"], dim=size["voc_dim"], intermediate_dim=size["voc_inter"],
                             num_layers=size["voc_layers"], adanorm_num_embeddings=4).eval()
    head = ISTFTHead(dim=size["voc_dim"], n_fft=1280, hop_length=320, padding="same").eval()
    backbone.load_state_dict({k[len("backbone."):]: v for k, v in sd.items() if k.startswith("backbone.")}, strict=True)
    head.load_state_dict({k[len("head."):]: v for k, v in sd.items() if k.startswith("head.") and "window" not in k}, strict=False)
    cb = sd["feature_extractor.codebook_weights"]
    g = torch.Generator().manual_seed(77)
    out = {}
    for i, n in enumerate((1, 9, 150)):
        codes = torch.randint(0, 1024, (n, 8), generator=g)
        # Vocos.codes_to_features (vocos/pretrained.py): offsets of quantizer.bins = 1024 per codebook, sum over codebooks
        idx = codes.T + (torch.arange(8) * 1024)[:, None]
        feats = torch.nn.functional.embedding(idx, cb).sum(dim=0).T[None]          # (1, 128, n)
        x = backbone(feats, bandwidth_id=torch.tensor([1]))
        wav = head(x)[0]
        out[f"codes_{i}"], out[f"wav_{i}"] = codes.numpy(), wav.numpy()
    path = os.path.join(HERE, "vocos_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path)
    return 0


if __name__ == "__main__":
    sys.exit(main())

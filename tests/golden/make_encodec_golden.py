"""Pins the Encodec encoder oracle against the REAL `encodec` package where it is importable (not in the build image:
oracle/encodec_oracle.py says "parity unpinned" until this has been run somewhere with `pip install encodec`).

    python tests/golden/make_encodec_golden.py          ->  tests/golden/encodec_golden.npz

Loads the seeded synthetic state dict of mars5_tts_b200.synth.make_encodec_state into the package's own
EncodecModel.encodec_model_24khz(pretrained=False) modules (weight norm folded, strict key check on the encoder and the
first 8 quantiser layers) and stores model.encode(wav[None, None]) codes at 6 kbps for seeded clips.
tests/test_encodec_golden.py (CPU: oracle, GPU: m5_encodec_encode) replays the fixture when it exists and skips with the
reason "encodec parity unpinned" when it does not.
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
        from encodec import EncodecModel
        from encodec.utils import convert_audio  # noqa: F401
    except ImportError as e:
        print(f"encodec is not importable here ({e}); the encoder oracle stays unpinned")
        return 2
    torch.set_grad_enabled(False)
    model = EncodecModel.encodec_model_24khz(pretrained=False).eval()
    model.set_target_bandwidth(6.0)
    for m in model.modules():   # fold weight norm like the reference's nuke_weight_norm (mars5/utils.py)
        try:
            torch.nn.utils.remove_weight_norm(m)
        except ValueError:
            pass
    sd = synth.make_encodec_state()
    own = model.state_dict()
    missing = [k for k in sd if k not in own]
    assert not missing, missing[:5]
    for k, v in sd.items():
        assert tuple(own[k].shape) == tuple(v.shape), (k, own[k].shape, v.shape)
    model.load_state_dict({**own, **sd}, strict=True)
    g = torch.Generator().manual_seed(123)
    out = {}
    for i, n in enumerate((24000, 12345, 641)):
        wav = torch.randn(n, generator=g) * 0.3
        codes = model.encode(wav[None, None])[0][0][0]        # (n_q, T)
        out[f"wav_{i}"], out[f"codes_{i}"] = wav.numpy(), codes.T.numpy()
    np.savez_compressed(os.path.join(HERE, "encodec_golden.npz"), **out)
    print("wrote encodec_golden.npz")
    return 0


if __name__ == "__main__":
    sys.exit(main())

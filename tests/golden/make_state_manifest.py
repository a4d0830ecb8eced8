"""Writes tests/golden/state_manifest.json: every state-dict key and shape of the reference's CodecLM and
ResidualTransformer built on the meta device with the constructor arguments of inference.py:101-110 (V = 8000,
2048 text ids), i.e. the layout of the released checkpoints.  Run in the build container:
    python tests/golden/make_state_manifest.py
"""
import json
import os
import sys

import torch

sys.path.insert(0, "/root/reference")
from mars5.model import CodecLM, ResidualTransformer  # noqa: E402

n_text, n_speech = 2048, 5952
with torch.device("meta"):
    lm = CodecLM(n_vocab=n_text + n_speech, dim=1536, dim_ff_scale=7 / 3)              # inference.py:105
    nar = ResidualTransformer(n_text_vocab=n_text + 1, n_quant=1025)                    # inference.py:109-110
out = {"n_text": n_text, "n_speech": n_speech,
       "ar": {k: list(v.shape) for k, v in lm.state_dict().items()},
       "nar": {k: list(v.shape) for k, v in nar.state_dict().items()}}
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "state_manifest.json")
json.dump(out, open(path, "w"))
print("wrote", path, len(out["ar"]), "AR tensors,", len(out["nar"]), "NAR tensors")

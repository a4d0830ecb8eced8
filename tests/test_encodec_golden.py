"""Encodec pin: replays tests/golden/encodec_golden.npz (tests/golden/make_encodec_golden.py, needs the real `encodec`
package) against the oracle (CPU) and m5_encodec_encode (GPU).  Absent fixture -> SKIP with "encodec parity unpinned"."""
import os

import numpy as np
import pytest
import torch

PATH = os.path.join(os.path.dirname(__file__), "golden", "encodec_golden.npz")
UNPINNED = "encodec parity unpinned: tests/golden/encodec_golden.npz absent (run tests/golden/make_encodec_golden.py where `encodec` is installed)"


@pytest.mark.skipif(not os.path.exists(PATH), reason=UNPINNED)
def test_oracle_matches_real_encodec():
    from mars5_tts_b200 import synth
    from oracle import encodec_oracle
    g, sd = np.load(PATH), synth.make_encodec_state()
    for i in range(3):
        got = encodec_oracle.encode(sd, torch.from_numpy(g[f"wav_{i}"])).numpy()
        assert (got != g[f"codes_{i}"]).mean() <= 0.01


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(PATH), reason=UNPINNED)
def test_m5_encodec_matches_real_encodec():
    from mars5_tts_b200 import synth
    from mars5_tts_b200.engine import Engine
    g, size = np.load(PATH), synth.TINY
    eng = Engine(synth.make_ar_state(size), synth.make_nar_state(size), synth.make_vocos_state(size), size["n_text"], device=0, max_pos=512,
                 enc_sd=synth.make_encodec_state())
    got = eng.encodec_encode([g[f"wav_{i}"] for i in range(3)])
    for i in range(3):
        assert (got[i] != g[f"codes_{i}"]).mean() <= 0.03
    eng.close()

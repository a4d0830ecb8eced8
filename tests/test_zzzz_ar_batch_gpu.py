"""Fused AR decode kernel at the batch widths of its three template instantiations (ar_decode_kernel<1> for B <= 8, <2> for
B <= 16, <4> for B <= 32 = the benchmarked configuration): every row of a wide batch must equal, bit for bit, the run of that
utterance alone -- and the solo runs are pinned to the unmodified reference's tokens in tests/test_pipeline_gpu.py.  Rows stop
at different steps (EOS / max_len), prompts and speaker references have different lengths, randomness is keyed by utterance id."""
import numpy as np
import pytest
import torch

from mars5_tts_b200 import synth
from mars5_tts_b200.engine import Engine, InferenceConfig
from tests.golden.inputs import make_inputs

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)


@pytest.fixture(scope="module")
def tiny_engine():
    inp = make_inputs()
    size = inp["size"]
    eng = Engine(synth.make_ar_state(size), synth.make_nar_state(size), synth.make_vocos_state(size), inp["n_text"], device=0, max_pos=512)
    yield inp, eng
    eng.close()


@pytest.mark.parametrize("B", [12, 32])
def test_ar_decode_rows_of_wide_batches_equal_solo_runs(tiny_engine, B):
    inp, eng = tiny_engine
    g = torch.Generator().manual_seed(100 + B)
    lo, hi = inp["n_text"], inp["n_text"] + 1024          # speech-token range of the tiny vocabulary
    plen = torch.randint(6, 60, (B,), generator=g).tolist()
    slen = torch.randint(3, 30, (B,), generator=g).tolist()
    prompts = [torch.randint(lo, hi, (n,), generator=g).tolist() for n in plen]
    spks = [torch.randint(0, 1024, (n, 8), generator=g).numpy() for n in slen]
    n_ph = torch.randint(3, 50, (B,), generator=g).tolist()
    utt = list(range(200, 200 + B))
    acfg = eng.make_ar_cfg(InferenceConfig(temperature=1.0, top_k=50, top_p=0.95), 90, inp["eos"], sync_every=3)
    batch, _, _ = eng.ar_generate(prompts, spks, n_ph, acfg, seed=5, utt_ids=utt)
    assert len(batch) == B
    for i in range(B):
        solo, _, _ = eng.ar_generate([prompts[i]], [spks[i]], [n_ph[i]], acfg, seed=5, utt_ids=[utt[i]])
        np.testing.assert_array_equal(batch[i], solo[0], err_msg=f"row {i} of a batch of {B}")
    assert any(len(b) > len(p) for b, p in zip(batch, prompts))

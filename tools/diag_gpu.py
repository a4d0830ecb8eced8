"""Concise GPU diagnostics (development aid): prints error levels of the pipeline stages against the oracle."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mars5_tts_b200 import synth, weights
from mars5_tts_b200.engine import Engine, InferenceConfig
from oracle import ar_oracle, nar_oracle
from tests.golden.inputs import make_inputs
torch.set_grad_enabled(False)
GOLD = np.load("tests/golden/reference_tiny.npz")
inp = make_inputs(); size = inp["size"]
ar_sd, nar_sd, voc_sd = synth.make_ar_state(size), synth.make_nar_state(size), synth.make_vocos_state(size)
eng = Engine(ar_sd, nar_sd, voc_sd, inp["n_text"], device=0, max_pos=512)
cfg = weights.dims_from_state(ar_sd, nar_sd, voc_sd, inp["n_text"])
g = torch.Generator().manual_seed(11)
prompts = [inp["ar_prompt"].tolist(), torch.randint(258, 1282, (33,), generator=g).tolist(), [256, 65, 257, 300]]
spks = [inp["ar_spk"].numpy(), torch.randint(0, 1024, (40, 8), generator=g).numpy(), torch.randint(0, 1024, (3, 8), generator=g).numpy()]
spks[1][25:, :] = 1024
outs = eng.ar_forward(prompts, spks)
for i, (pr, sp, got) in enumerate(zip(prompts, spks, outs)):
    ref = ar_oracle.codeclm_forward(ar_sd, cfg, torch.tensor(pr), torch.from_numpy(sp)).numpy()
    d = np.abs(got - ref)
    print(f"ar_forward[{i}] max|ref|={np.abs(ref).max():.3f} maxerr={d.max():.5f} per-row max:", np.round(d.max(1)[:8], 4))
one = eng.ar_forward([prompts[0]], [spks[0]])[0]
print("ar_forward single vs batch:", np.abs(one - outs[0]).max(), " vs golden:", np.abs(one - GOLD["ar_logits"]).max())
acfg = eng.make_ar_cfg(InferenceConfig(), inp["ar_max_len"], inp["eos"], sync_every=4)
ids, hit, dump = eng.ar_generate([inp["ar_prompt"].tolist()], [inp["ar_spk"].numpy()], [7], acfg, noise=inp["ar_noise"][None].numpy(), dump_steps=4)
print("gen  :", ids[0].tolist()); print("gold :", GOLD["ar_gen_cache"].tolist()); print("hit", hit)
seq = GOLD["ar_gen_cache"]; P = len(inp["ar_prompt"])
for s in range(4):
    ref = ar_oracle.codeclm_forward(ar_sd, cfg, torch.from_numpy(seq[:P + s]), inp["ar_spk"])[-1].numpy()
    print(f"  step {s}: dump-vs-oracle maxerr {np.abs(dump[0, s] - ref).max():.5f} argmax {dump[0,s].argmax()} {ref.argmax()}")
for precise in (0, 1):
    got = eng.nar_forward([inp["nar_c_text"].numpy()], [inp["nar_c_codes"].numpy()], [inp["nar_x"].numpy()], 7, precise=bool(precise))[0]
    d = np.abs(got - GOLD["nar_logits_cond"])
    print(f"nar_forward precise={precise}: maxerr {d.max():.5f} mean {d.mean():.6f} max|ref| {np.abs(GOLD['nar_logits_cond']).max():.2f}")

"""AR decode micro-benchmark on the full-size model: B=32 rows, prompt of P tokens, N forced decode steps.  Prints ms per
decode step and the achieved fraction of the measured HBM peak (profile kind 2 of the C ABI); with M5_AR_PROFILE=1 the
library also prints the per-phase timeline of the fused kernel.   python tools/ar_decode_bench.py [--P 1200] [--N 200]"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from mars5_tts_b200 import dist, synth  # noqa: E402
from mars5_tts_b200.engine import Engine, InferenceConfig  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--P", type=int, default=1200)
ap.add_argument("--N", type=int, default=200)
ap.add_argument("--B", type=int, default=32)
ap.add_argument("--size", default="full")
args = ap.parse_args()
size = {"full": synth.FULL, "mid": synth.MID, "tiny": synth.TINY}[args.size]
eng = Engine(device=0, packed=dist.build_or_receive_weights(size, 0, 1, 0, max_pos=4096))
g = torch.Generator().manual_seed(0)
n_text = size["n_text"]
prompts = [torch.randint(n_text, n_text + 1024, (args.P,), generator=g).numpy().astype(np.int32) for _ in range(args.B)]
spk = [torch.randint(0, 1024, (450, 8), generator=g).numpy().astype(np.int32) for _ in range(args.B)]
eos = eng.dims["ar_vocab"] - 1
acfg = eng.make_ar_cfg(InferenceConfig(), args.P + args.N + 2, eos, force_len=args.N + 8, sync_every=64)  # rows still running at the last step: its profile is a normal step
peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6569.6
from bench import ClockSampler  # noqa: E402  (nvidia-smi clock samples during the decode loops)
ap_reps = int(os.environ.get("M5_AR_BENCH_REPS", "3"))
for rep in range(ap_reps):
    cs = ClockSampler(0)
    cs.start()
    eng.lib.m5_profile_enable(eng.ctx, 1)
    ids, _, _ = eng.ar_generate(prompts, spk, [500] * args.B, acfg, seed=rep)
    a, b, c, n = C.c_double(), C.c_double(), C.c_double(), C.c_int64()
    eng.lib.m5_profile_read(eng.ctx, 2, C.byref(a), C.byref(b), C.byref(c), C.byref(n))
    eng.lib.m5_profile_enable(eng.ctx, 0)
    gbs = c.value / max(a.value, 1e-9) / 1e6
    print(f"rep {rep}: {n.value} decode steps, {a.value / max(n.value, 1):.3f} ms/step, {c.value / max(n.value, 1) / 1e9:.2f} GB/step algorithmic, "
          f"{gbs:.0f} GB/s = {gbs / peak:.3f} of {peak:.0f}; generated {len(ids[0]) - args.P} tokens; clocks {cs.stop()}", flush=True)

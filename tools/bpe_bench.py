"""CPU micro-benchmark of the tokeniser stage between the AR and NAR loops at the BASELINE
configs[2] batch: 32 utterances x (450-frame prompt L0 codes -> speech BPE, ~135-token text -> text BPE, ~1000 generated
speech tokens -> decode_int).  Native merge engine (csrc/bpe.cu, batched, threaded) vs -- when /root/reference is present
(build container only) -- the unmodified reference classes, whose outputs are also compared.  Tokeniser models: the golden fixtures' (trained by the reference on synthetic corpora).

    python tools/bpe_bench.py > profiles/r1_bpe_cpu.txt
"""
import json
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mars5_tts_b200 import bpe  # noqa: E402

gold = json.load(open(os.path.join(ROOT, "tests", "golden", "bpe_golden.json"), encoding="utf-8"))
rng = random.Random(7)
B = 32
sp = bpe.CodebookTokenizer(bpe.GPT4_SPLIT_PATTERN); sp.load(gold["speech"]["model"])
tx = bpe.RegexTokenizer(); tx.load(gold["text"]["model"])
# inputs with the statistics of the golden corpus (so that merges actually fire)
hot = sorted({c for case in gold["speech"]["cases"] for c in case["codes"]})[:40]
prompts = [[rng.choice(hot) if rng.random() < 0.85 else rng.randrange(1024) for _ in range(450)] for _ in range(B)]
texts = [gold["text"]["cases"][5]["text"][: 600 + 10 * i] for i in range(B)]
gen = [sp.encode_codes_batch([[rng.choice(hot) if rng.random() < 0.85 else rng.randrange(1024) for _ in range(1500)]])[0] for _ in range(B)]


def best(fn, reps=5):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    return min(ts)


rows = []
t_native_enc = best(lambda: sp.encode_codes_batch(prompts))
rows.append(("speech BPE encode, 32 x 450 codes", t_native_enc))
t_native_dec = best(lambda: sp.decode_int_batch(gen))
rows.append((f"decode_int, 32 x {len(gen[0])} tokens -> 1500 codes", t_native_dec))
t_native_txt = best(lambda: tx.encode_batch(texts))
rows.append((f"text BPE encode, 32 texts (~{sum(map(len, texts)) // B} chars)", t_native_txt))

ref = {}
if os.path.isdir("/root/reference/mars5"):
    sys.path.insert(0, "/root/reference")
    import io
    from mars5.minbpe.codebook import CodebookTokenizer
    from mars5.minbpe.regex import RegexTokenizer
    rsp = CodebookTokenizer(bpe.GPT4_SPLIT_PATTERN); rsp.load(io.BytesIO(gold["speech"]["model"].encode()))
    rtx = RegexTokenizer(); rtx.load(io.BytesIO(gold["text"]["model"].encode()))
    strs = [" ".join(map(str, p)) for p in prompts]
    assert [rsp.encode(s) for s in strs] == sp.encode_codes_batch(prompts)
    assert [rtx.encode(t) for t in texts] == tx.encode_batch(texts)
    assert [rsp.decode_int(g) for g in gen] == sp.decode_int_batch(gen)
    ref["speech BPE encode, 32 x 450 codes"] = best(lambda: [rsp.encode(s) for s in strs], reps=2)
    ref[rows[1][0]] = best(lambda: [rsp.decode_int(g) for g in gen], reps=2)
    ref[rows[2][0]] = best(lambda: [rtx.encode(t) for t in texts], reps=2)

print(f"# tools/bpe_bench.py on {os.cpu_count()} host threads; best of 5 (native) / 2 (reference); batch = {B} utterances")
print(f"# {'stage':58s} {'native':>10s} {'reference':>11s}  speed-up")
for name, tn in rows:
    tr = ref.get(name)
    print(f"  {name:58s} {tn * 1e3:8.2f} ms " + (f"{tr * 1e3:8.1f} ms  {tr / tn:7.1f}x" if tr else "          -"))
tot_n = sum(r[1] for r in rows)
tot_r = sum(ref.values()) if ref else None
print(f"# host time per batch of 32 between/around the GPU stages: native {tot_n * 1e3:.1f} ms"
      + (f", reference {tot_r * 1e3:.0f} ms ({tot_r / tot_n:.0f}x); the GPU step itself is 37.4 s" if tot_r else ""))

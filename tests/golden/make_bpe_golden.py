"""Generates tests/golden/bpe_golden.json by running the UNMODIFIED reference tokenisers (mars5/minbpe) in the build
container:  python tests/golden/make_bpe_golden.py

The reference ships no tokenizer models (they sit inside the checkpoints, which are not available offline), so two
small "minbpe v1" models are trained with the reference's own `train()` on seeded synthetic corpora, saved with its own
`save()`, re-loaded with its own `load()`, and its `encode` / `decode` / `decode_int` outputs on seeded inputs are stored
next to the model text.  tests/test_bpe_cpu.py replays them against mars5_tts_b200.bpe and oracle/bpe_oracle.py.
"""
import json
import os
import random
import sys
import tempfile

sys.path.insert(0, "/root/reference")
from mars5.minbpe.codebook import CodebookTokenizer  # noqa: E402
from mars5.minbpe.regex import GPT4_SPLIT_PATTERN, RegexTokenizer  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def synth_text(rng, n_words):
    syl = ["ka", "to", "mi", "ra", "sen", "lo", "vi", "ne", "shu", "ba", "the", "ing", "qu", "st", "é", "ü", "ñ", "漢", "字"]
    punct = [" ", " ", " ", ", ", ". ", "! ", "? ", "'s ", "'ll ", "\n", "\n\n", "  ", " - ", " 12", " 2024 ", "\t"]
    out = []
    for _ in range(n_words):
        w = "".join(rng.choice(syl) for _ in range(rng.randint(1, 4)))
        if rng.random() < 0.15:
            w = w.capitalize()
        out.append(w + rng.choice(punct))
    return "".join(out)


def main():
    rng = random.Random(1234)
    tmp = tempfile.mkdtemp()
    gold = {}

    # ---- speech tokeniser: 1024 codes, 400 merges, end-of-speech special right after the merges
    hot = [rng.randrange(1024) for _ in range(40)]
    corpus = [rng.choice(hot) if rng.random() < 0.85 else rng.randrange(1024) for _ in range(12000)]
    sp = CodebookTokenizer(GPT4_SPLIT_PATTERN)
    sp.train(" ".join(map(str, corpus)), 1024 + 400)
    sp.register_special_tokens({"<|endofspeech|>": 1024 + 400})
    sp.save(os.path.join(tmp, "speechtok"))
    sp2 = CodebookTokenizer(GPT4_SPLIT_PATTERN)
    sp2.load(os.path.join(tmp, "speechtok.model"))
    cases = []
    for n in [1, 2, 3, 7, 64, 450, 451, 1500]:
        seq = [rng.choice(hot) if rng.random() < 0.85 else rng.randrange(1024) for _ in range(n)]
        txt = " ".join(map(str, seq))
        ids = sp2.encode(txt)
        with_special = ids[: len(ids) // 2] + [1024 + 400] + ids[len(ids) // 2:]
        cases.append({"codes": seq, "ids": ids, "encode_all": sp2.encode(txt + "<|endofspeech|>", allowed_special="all"),
                      "decode_int": sp2.decode_int(with_special), "decode_int_in": with_special,
                      "decode": sp2.decode(with_special)})
    # overlapping-pair edge case: runs of one code
    run = [hot[0]] * 9 + [hot[1]] * 5 + [hot[0]] * 4
    cases.append({"codes": run, "ids": sp2.encode(" ".join(map(str, run))), "encode_all": sp2.encode(" ".join(map(str, run)), allowed_special="all"),
                  "decode_int": sp2.decode_int(sp2.encode(" ".join(map(str, run)))), "decode_int_in": sp2.encode(" ".join(map(str, run))),
                  "decode": sp2.decode(sp2.encode(" ".join(map(str, run))))})
    gold["speech"] = {"model": open(os.path.join(tmp, "speechtok.model"), encoding="utf-8").read(), "vocab_len": len(sp2.vocab),
                      "cases": cases}

    # ---- text tokeniser: GPT-4 split pattern, 300 merges, two specials
    tx = RegexTokenizer()
    tx.train(synth_text(rng, 6000), 256 + 300)
    tx.register_special_tokens({"<|startoftext|>": 556, "<|endoftext|>": 557})
    tx.save(os.path.join(tmp, "texttok"))
    tx2 = RegexTokenizer()
    tx2.load(os.path.join(tmp, "texttok.model"))
    tcases = []
    texts = ["", " ", "a", "Hello world, it's me!", synth_text(rng, 40), synth_text(rng, 400),
             "<|startoftext|>" + synth_text(rng, 25) + " " + synth_text(rng, 30).strip() + "<|endoftext|>",
             "tabs\tand\r\nnewlines\n\n\n   trailing   ", "numbers 1234567 and 12 and 2024's", "ünïcödé 漢字漢字 ñandú"]
    for t in texts:
        ids = tx2.encode(t, allowed_special="all")
        tcases.append({"text": t, "ids": ids, "ordinary": tx2.encode_ordinary(t), "decode": tx2.decode(ids)})
    gold["text"] = {"model": open(os.path.join(tmp, "texttok.model"), encoding="utf-8").read(), "vocab_len": len(tx2.vocab),
                    "cases": tcases}

    with open(os.path.join(HERE, "bpe_golden.json"), "w", encoding="utf-8") as f:
        json.dump(gold, f, ensure_ascii=False)
    print("wrote", os.path.join(HERE, "bpe_golden.json"), os.path.getsize(os.path.join(HERE, "bpe_golden.json")), "bytes")


if __name__ == "__main__":
    main()

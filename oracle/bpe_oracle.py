"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): CPU restatement of the reference's "minbpe v1" encode / decode
loops, pinned to outputs of the unmodified reference by tests/golden/bpe_golden.json (made by
tests/golden/make_bpe_golden.py).  Used to check mars5_tts_b200.bpe (native merge engine) on inputs larger and more
adversarial than the golden set, and as the pure-Python timing baseline of tools/bpe_bench.py.

Follows:
  * chunk encoding  -- RegexTokenizer._encode_chunk / CodebookTokenizer._encode_chunk (mars5/minbpe/regex.py:92-111,
                       codebook.py:96-115) with get_stats / merge of mars5/minbpe/base.py:14-41
  * expansion       -- Tokenizer._build_vocab (base.py:88-96), CodebookTokenizer._build_vocab (codebook.py:207-215),
                       decode_int (codebook.py:88-94)
"""
from typing import Dict, List, Sequence, Tuple


def merge_table(pairs: Sequence[Tuple[int, int]], base: int) -> Dict[Tuple[int, int], int]:
    """pair -> token id, in model-file order; a repeated pair keeps the last id (dict assignment in load())."""
    table = {}
    for i, (a, b) in enumerate(pairs):
        table[(a, b)] = base + i
    return table


def encode_chunk(ids: Sequence[int], table: Dict[Tuple[int, int], int]) -> List[int]:
    """Repeat: among the adjacent pairs present, take the one with the smallest token id in the table; replace every
    non-overlapping occurrence of it, scanning left to right; stop when no present pair is in the table."""
    ids = list(ids)
    while len(ids) >= 2:
        best = None
        for pair in zip(ids, ids[1:]):
            tok = table.get(pair)
            if tok is not None and (best is None or tok < best[1]):
                best = (pair, tok)
        if best is None:
            break
        (a, b), tok = best
        out, i = [], 0
        while i < len(ids):
            if i + 1 < len(ids) and ids[i] == a and ids[i + 1] == b:
                out.append(tok)
                i += 2
            else:
                out.append(ids[i])
                i += 1
        ids = out
    return ids


def expansions(pairs: Sequence[Tuple[int, int]], base: int) -> Dict[int, Tuple[int, ...]]:
    """token id -> tuple of base symbols (bytes for the text tokeniser, codebook indices for the speech tokeniser)."""
    exp = {i: (i,) for i in range(base)}
    for (a, b), tok in merge_table(pairs, base).items():
        exp[tok] = exp[a] + exp[b]
    return exp


def decode_int(ids: Sequence[int], pairs: Sequence[Tuple[int, int]], base: int, specials: Dict[str, int]) -> list:
    """Codebook indices of the sequence, special tokens as their strings, in order (specials win over merge ids)."""
    exp = expansions(pairs, base)
    inv = {v: k for k, v in specials.items()}
    out: list = []
    for t in ids:
        if t in inv:
            out.append(inv[t])
        elif t in exp:
            out.extend(exp[t])
        else:
            raise ValueError(f"invalid token id: {t}")
    return out

"""Text and speech tokenisers of the MARS5 front/back end ("minbpe v1" model files) on the native merge engine.

Mirrors the reference's ``RegexTokenizer`` (mars5/minbpe/regex.py) and ``CodebookTokenizer`` (mars5/minbpe/codebook.py):
same constructor arguments, ``load`` / ``register_special_tokens`` / ``encode`` / ``encode_ordinary`` / ``decode`` /
``decode_int``, same ``vocab`` / ``merges`` / ``special_tokens`` attributes, same results (tests/test_bpe_cpu.py checks
them against outputs of the unmodified reference).  What differs is where the time goes: the merge loops run in
``libmars5_b200.so`` (csrc/bpe.cu, ``m5_bpe_*`` in include/mars5_b200.h) over a whole batch of sequences at once
(``encode_batch`` / ``decode_int_batch``), instead of one Python ``while`` loop per utterance between the AR and NAR stages
(inference.py:237-243,272-275).  Training a tokeniser is not part of inference and is not provided.
"""
import ctypes as C
import io
import os
from typing import Dict, Iterable, List, Optional, Sequence, Tuple, Union

import numpy as np

from . import capi

# the split patterns the reference ships (mars5/minbpe/regex.py:17-18); a model file carries its own pattern line
GPT2_SPLIT_PATTERN = r"""'(?:[sdmt]|ll|ve|re)| ?\p{L}+| ?\p{N}+| ?[^\s\p{L}\p{N}]+|\s+(?!\S)|\s+"""
GPT4_SPLIT_PATTERN = r"""'(?i:[sdmt]|ll|ve|re)|[^\r\n\p{L}\p{N}]?+\p{L}+|\p{N}{1,3}| ?[^\s\p{L}\p{N}]++[\r\n]*|\s*[\r\n]|\s+(?!\S)|\s+"""


def _regex():
    import regex  # third-party `regex` (Unicode classes, possessive quantifiers), the module the reference uses too
    return regex


def parse_model(model_file) -> Tuple[str, Dict[str, int], List[Tuple[int, int]]]:
    """Reads a "minbpe v1" model: (pattern, special tokens in file order, merge pairs in file order).
    Accepts a path ending in .model, an io.BytesIO / text stream, or the file's content as str / bytes."""
    if isinstance(model_file, (bytes, bytearray)):
        text = bytes(model_file).decode("utf-8")
    elif isinstance(model_file, io.BytesIO):
        text = model_file.getvalue().decode("utf-8")
    elif hasattr(model_file, "read"):
        text = model_file.read()
        text = text.decode("utf-8") if isinstance(text, bytes) else text
    elif isinstance(model_file, str) and model_file.startswith("minbpe v1"):
        text = model_file
    else:
        path = os.fspath(model_file)
        assert path.endswith(".model")
        with open(path, encoding="utf-8") as f:
            text = f.read()
    lines = text.split("\n")
    assert lines[0].strip() == "minbpe v1"
    pattern = lines[1].strip()
    n_special = int(lines[2].strip())
    specials: Dict[str, int] = {}
    for ln in lines[3:3 + n_special]:
        name, idx = ln.strip().split()
        specials[name] = int(idx)
    merges = []
    for ln in lines[3 + n_special:]:
        if not ln.strip():
            continue
        a, b = ln.split()
        merges.append((int(a), int(b)))
    return pattern, specials, merges


class _MergeEngine:
    """Owner of one native m5_bpe handle."""

    def __init__(self, base: int, pairs: Sequence[Tuple[int, int]]):
        self.lib = capi.load()
        self.base, self.n_merges = base, len(pairs)
        arr = np.ascontiguousarray(np.asarray(pairs, dtype=np.int32).reshape(-1, 2))
        self.handle = C.c_void_p()
        rc = self.lib.m5_bpe_create(base, arr.ctypes.data_as(C.c_void_p), len(pairs), C.byref(self.handle))
        if rc != 0:
            raise ValueError(f"invalid merge table (m5_bpe_create rc={rc})")

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h:
            self.lib.m5_bpe_destroy(h)

    def encode_packed(self, ids: np.ndarray, offsets: np.ndarray, n_threads: int = 0) -> List[np.ndarray]:
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        n_seq = len(offsets) - 1
        out = np.empty_like(ids)
        out_len = np.zeros(max(n_seq, 1), dtype=np.int32)
        rc = self.lib.m5_bpe_encode(self.handle, ids.ctypes.data_as(C.c_void_p), offsets.ctypes.data_as(C.c_void_p), n_seq,
                                    out.ctypes.data_as(C.c_void_p), out_len.ctypes.data_as(C.c_void_p), n_threads)
        if rc != 0:
            raise ValueError(f"m5_bpe_encode rc={rc} (ids outside the vocabulary?)")
        return [out[offsets[s]:offsets[s] + out_len[s]] for s in range(n_seq)]

    def expand_packed(self, ids: np.ndarray, offsets: np.ndarray, special_ids: Sequence[int]) -> Tuple[np.ndarray, np.ndarray]:
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        sp = np.ascontiguousarray(np.asarray(list(special_ids), dtype=np.int32))
        n_seq = len(offsets) - 1
        args = (self.handle, ids.ctypes.data_as(C.c_void_p), offsets.ctypes.data_as(C.c_void_p), n_seq,
                sp.ctypes.data_as(C.c_void_p), len(sp))
        n = self.lib.m5_bpe_expand(*args, None, None, 0)
        if n < 0:
            raise ValueError("invalid token id")
        out = np.empty(max(int(n), 1), dtype=np.int32)
        out_off = np.zeros(n_seq + 1, dtype=np.int64)
        n2 = self.lib.m5_bpe_expand(*args, out.ctypes.data_as(C.c_void_p), out_off.ctypes.data_as(C.c_void_p), int(n))
        assert n2 == n
        return out[:n], out_off


def _pack(seqs: Sequence[np.ndarray]) -> Tuple[np.ndarray, np.ndarray]:
    lens = np.fromiter((len(s) for s in seqs), dtype=np.int64, count=len(seqs))
    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    ids = np.concatenate([np.asarray(s, dtype=np.int32) for s in seqs]) if len(seqs) and offsets[-1] > 0 else np.zeros(0, np.int32)
    return ids, offsets


class _Tokenizer:
    """State shared by both tokenisers: merge table, specials, the derived vocab (Tokenizer in minbpe/base.py:67-96)."""

    base = 256

    def __init__(self):
        self.merges: Dict[Tuple[int, int], int] = {}
        self.pattern = ""
        self.special_tokens: Dict[str, int] = {}
        self.inverse_special_tokens: Dict[int, str] = {}
        self._engine: Optional[_MergeEngine] = None
        self.vocab = self._build_vocab()

    # ---- model state
    def _base_bytes(self, idx: int) -> bytes:
        return bytes([idx])

    def _build_vocab(self) -> Dict[int, bytes]:
        vocab = {idx: self._base_bytes(idx) for idx in range(self.base)}
        for (p0, p1), idx in self.merges.items():
            vocab[idx] = vocab[p0] + vocab[p1]
        for name, idx in self.special_tokens.items():
            vocab[idx] = name.encode("utf-8")
        return vocab

    def _set_merges(self, pairs: Sequence[Tuple[int, int]]):
        self.merges = {}
        for i, pr in enumerate(pairs):
            self.merges[(int(pr[0]), int(pr[1]))] = self.base + i
        self._engine = _MergeEngine(self.base, pairs)

    def load(self, model_file):
        pattern, specials, pairs = parse_model(model_file)
        self.pattern = pattern
        self.special_tokens = specials
        self.inverse_special_tokens = {v: k for k, v in specials.items()}
        self._set_merges(pairs)
        self.vocab = self._build_vocab()

    def register_special_tokens(self, special_tokens: Dict[str, int]):
        self.special_tokens = dict(special_tokens)
        self.inverse_special_tokens = {v: k for k, v in special_tokens.items()}

    def train(self, *a, **k):
        raise NotImplementedError("tokeniser training is not part of the inference path")

    @property
    def engine(self) -> _MergeEngine:
        if self._engine is None:
            self._engine = _MergeEngine(self.base, [])
        return self._engine

    # ---- special-token policy of encode() (regex.py:122-143)
    def _allowed(self, text: str, allowed_special) -> Dict[str, int]:
        if allowed_special == "all":
            return self.special_tokens
        if allowed_special == "none":
            return {}
        if allowed_special == "none_raise":
            assert all(token not in text for token in self.special_tokens)
            return {}
        if isinstance(allowed_special, set):
            return {k: v for k, v in self.special_tokens.items() if k in allowed_special}
        raise ValueError(f"allowed_special={allowed_special} not understood")

    def _split_special(self, text: str, special: Dict[str, int]) -> List[str]:
        if not special:
            return [text]
        re = _regex()
        return re.split("(" + "|".join(re.escape(k) for k in special) + ")", text)

    def decode(self, ids: Iterable[int]) -> str:
        parts = []
        for idx in ids:
            idx = int(idx)
            if idx in self.vocab:
                parts.append(self.vocab[idx])
            elif idx in self.inverse_special_tokens:
                parts.append(self.inverse_special_tokens[idx].encode("utf-8"))
            else:
                raise ValueError(f"invalid token id: {idx}")
        return b"".join(parts).decode("utf-8", errors="replace")


class RegexTokenizer(_Tokenizer):
    """Byte-level BPE over regex-split chunks (mars5/minbpe/regex.py:21-164)."""

    base = 256

    def __init__(self, pattern: Optional[str] = None):
        super().__init__()
        self.pattern = GPT4_SPLIT_PATTERN if pattern is None else pattern
        self._ctor_pattern = self.pattern   # what the reference compiles ONCE in __init__ (regex.py:32-33)
        self._compiled = None

    def load(self, model_file):
        # like the reference: Tokenizer.load (base.py:140-168) overwrites self.pattern with the model file's line but the
        # regex compiled in the constructor keeps being used -- the split pattern of a loaded model never takes effect
        super().load(model_file)

    @property
    def compiled_pattern(self):
        if self._compiled is None:
            self._compiled = _regex().compile(self._ctor_pattern)
        return self._compiled

    # -> for every text a list of segments: an int (special id) or a list of chunk byte strings
    def _segments(self, text: str, allowed_special):
        special = self._allowed(text, allowed_special)
        segs = []
        for part in self._split_special(text, special):
            if special and part in special:
                segs.append(special[part])
            else:
                segs.append([ch.encode("utf-8") for ch in self.compiled_pattern.findall(part)])
        return segs

    def encode_batch(self, texts: Sequence[str], allowed_special="none_raise", n_threads: int = 0) -> List[List[int]]:
        all_segs = [self._segments(t, allowed_special) for t in texts]
        chunks = [ch for segs in all_segs for seg in segs if not isinstance(seg, int) for ch in seg]
        lens = np.fromiter((len(c) for c in chunks), dtype=np.int64, count=len(chunks))
        offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        flat = np.frombuffer(b"".join(chunks), dtype=np.uint8).astype(np.int32)
        merged = iter(self.engine.encode_packed(flat, offsets, n_threads)) if len(chunks) else iter(())
        out = []
        for segs in all_segs:
            ids: List[int] = []
            for seg in segs:
                if isinstance(seg, int):
                    ids.append(seg)
                else:
                    for _ in seg:
                        ids.extend(next(merged).tolist())
            out.append(ids)
        return out

    def encode_ordinary(self, text: str) -> List[int]:
        return self.encode_batch([text], allowed_special="none")[0]

    def encode(self, text: str, allowed_special="none_raise") -> List[int]:
        return self.encode_batch([text], allowed_special)[0]


class CodebookTokenizer(_Tokenizer):
    """BPE over Encodec codebook indices written as space-separated integers (mars5/minbpe/codebook.py:14-215)."""

    def __init__(self, pattern: Optional[str] = None, codebook_size: int = 1024):
        self.base = codebook_size
        self.codebook_size = codebook_size
        super().__init__()
        self.pattern = pattern

    def _base_bytes(self, idx: int) -> bytes:
        return f" {idx:04d}".encode("utf-8")

    @staticmethod
    def _parse_codes(chunk: str) -> np.ndarray:
        return np.asarray([int(tok) for tok in chunk.split(" ")], dtype=np.int32)   # ValueError on "" like the reference

    def _segments(self, text: str, allowed_special):
        special = self._allowed(text, allowed_special)
        if not special:
            return [self._parse_codes(text)]
        segs = []
        for part in self._split_special(text, special):
            part = part.strip()
            if len(part) == 0:
                continue
            segs.append(special[part] if part in special else self._parse_codes(part))
        return segs

    def _encode_segments(self, all_segs, n_threads: int) -> List[List[int]]:
        seqs = [seg for segs in all_segs for seg in segs if not isinstance(seg, int)]
        merged = iter(self.engine.encode_packed(*_pack(seqs), n_threads)) if seqs else iter(())
        out = []
        for segs in all_segs:
            ids: List[int] = []
            for seg in segs:
                if isinstance(seg, int):
                    ids.append(seg)
                else:
                    ids.extend(next(merged).tolist())
            out.append(ids)
        return out

    def encode_batch(self, texts: Sequence[str], allowed_special="none_raise", n_threads: int = 0) -> List[List[int]]:
        return self._encode_segments([self._segments(t, allowed_special) for t in texts], n_threads)

    def encode_codes_batch(self, codes: Sequence[Union[np.ndarray, Sequence[int]]], n_threads: int = 0) -> List[List[int]]:
        """encode() of `' '.join(map(str, codes))` without the detour through a string (inference.py:236-238)."""
        return self._encode_segments([[np.asarray(c, dtype=np.int32)] for c in codes], n_threads)

    def encode_ordinary(self, text: str) -> List[int]:
        return self.encode_batch([text], allowed_special="none")[0]

    def encode(self, text: str, allowed_special="none_raise") -> List[int]:
        return self.encode_batch([text], allowed_special)[0]

    def decode_int_batch(self, ids_batch: Sequence[Sequence[int]]) -> List[list]:
        """decode_int() (codebook.py:88-94) for a batch: codebook integers, special tokens as their strings."""
        names = list(self.special_tokens)
        flat, offsets = _pack([np.asarray(i, dtype=np.int32) for i in ids_batch])
        # a special id always decodes to its name, even when it collides with a merge id (the reference's _build_vocab
        # writes specials last): route every special through an id beyond the merge table
        n_tok = self.base + self.engine.n_merges
        flat = flat.copy()
        sentinels = []
        for k, name in enumerate(names):
            flat[flat == self.special_tokens[name]] = n_tok + k
            sentinels.append(n_tok + k)
        syms, out_off = self.engine.expand_packed(flat, offsets, sentinels)
        out = []
        for s in range(len(ids_batch)):
            row = syms[out_off[s]:out_off[s + 1]]
            if (row >= 0).all():
                out.append(row.tolist())
            else:
                out.append([int(v) if v >= 0 else names[-int(v) - 1] for v in row])
        return out

    def decode_int(self, ids: Sequence[int]) -> list:
        return self.decode_int_batch([ids])[0]

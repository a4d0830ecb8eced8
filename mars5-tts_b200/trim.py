"""Silence trim of synthesised audio: the step after the vocoder in the reference's ``tts()`` (inference.py:304-305,
mars5/trim.py:110-177 -- a port of ``librosa.effects.trim`` that no longer runs under numpy 2).  Same call shape as the
reference's ``trim(y, top_db=...)`` for mono input; the work is done by ``m5_trim_bounds`` (csrc/trim.cu, host code,
a batch of waveforms in parallel)."""
import ctypes as C
from typing import List, Sequence, Tuple

import numpy as np
import torch

from . import capi


def trim_bounds_batch(wavs: Sequence[np.ndarray], top_db: float = 60, frame_length: int = 2048, hop_length: int = 512,
                      n_threads: int = 0) -> List[Tuple[int, int]]:
    """[start, end) of the non-silent region of every waveform (1-D float32)."""
    lib = capi.load()
    arrs = [np.ascontiguousarray(np.asarray(w, dtype=np.float32).reshape(-1)) for w in wavs]
    offsets = np.concatenate([[0], np.cumsum([len(a) for a in arrs])]).astype(np.int64)
    flat = np.concatenate(arrs) if arrs else np.zeros(0, np.float32)
    start = np.zeros(max(len(arrs), 1), dtype=np.int64)
    end = np.zeros(max(len(arrs), 1), dtype=np.int64)
    rc = lib.m5_trim_bounds(len(arrs), flat.ctypes.data_as(C.c_void_p), offsets.ctypes.data_as(C.c_void_p), float(top_db),
                            frame_length, hop_length, start.ctypes.data_as(C.c_void_p), end.ctypes.data_as(C.c_void_p), n_threads)
    if rc != 0:
        raise RuntimeError(f"m5_trim_bounds rc={rc}: every waveform must be longer than frame_length // 2 = {frame_length // 2} samples "
                           "(reflect padding), top_db >= 0")
    return [(int(s), int(e)) for s, e in zip(start[:len(arrs)], end[:len(arrs)])]


def trim(y, top_db: float = 60, frame_length: int = 2048, hop_length: int = 512):
    """``y_trimmed, index`` like the reference's trim(): ``y`` a mono waveform (1-D, or (1, n) / (c, n) which is averaged
    over channels to find the range, as the reference does); ``index`` = tensor([start, end])."""
    t = y if isinstance(y, torch.Tensor) else torch.as_tensor(np.asarray(y))
    mono = t.float().mean(dim=0) if t.dim() > 1 else t.float()
    (start, end), = trim_bounds_batch([mono.detach().cpu().numpy()], top_db, frame_length, hop_length)
    return t[..., start:end], torch.asarray([start, end])

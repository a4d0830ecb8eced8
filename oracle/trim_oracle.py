"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): float32 restatement of the reference's silence trim, pinned to
outputs of the reference itself by tests/golden/trim_golden.json (made by tests/golden/make_trim_golden.py).

Follows mars5/trim.py: trim :110-177, _signal_to_frame_nonsilent :69-107, rms :180-289 (centered, reflect padding, mean of
squares per 2048-sample frame every 512 samples, sqrt and squared again), power_to_db :579-699 with ref = max and
amin = 1e-10, frames_to_samples :702-741.
"""
import torch
import torch.nn.functional as F


def trim_bounds(y: torch.Tensor, top_db: float = 60, frame_length: int = 2048, hop_length: int = 512):
    y = y.float()
    pad = frame_length // 2
    yp = F.pad(y[None, None], (pad, pad), mode="reflect")[0, 0]
    frames = yp.unfold(0, frame_length, hop_length)               # (n_frames, frame_length)
    mse = torch.sqrt(torch.mean(frames.abs() ** 2, dim=1)) ** 2   # rms(...) ** 2, float32 like the reference
    amin = torch.tensor(1e-10)
    db = 10.0 * torch.log10(torch.maximum(amin, mse)) - 10.0 * torch.log10(torch.maximum(amin, mse.max()))
    idx = torch.nonzero(db > -top_db).flatten()
    if idx.numel() == 0:
        return 0, 0
    return int(idx[0]) * hop_length, min(y.shape[-1], (int(idx[-1]) + 1) * hop_length)

"""Multi-GPU plumbing (one process per GPU, torch.distributed / NCCL over NVLink): the path shards by utterance with no
collective inside any loop.  Exactly two collectives exist: one broadcast of the repacked weight blob
at start-up and one all-gather of the finished waveforms."""
import numpy as np
import torch


def _flatten(tensors):
    """name -> tensor  ==>  (uint8 blob, manifest) with 256-byte aligned segments."""
    manifest, off = [], 0
    for k, v in tensors.items():
        n = v.numel() * v.element_size()
        manifest.append((k, str(v.dtype), tuple(v.shape), off, n))
        off += (n + 255) & ~255
    blob = torch.empty(off, dtype=torch.uint8)
    for (k, _, _, o, n), v in zip(manifest, tensors.values()):
        blob[o:o + n] = v.contiguous().view(torch.uint8).reshape(-1)
    return blob, manifest


def _unflatten(blob, manifest):
    out = {}
    for k, dt, shape, o, n in manifest:
        dtype = {"torch.float16": torch.float16, "torch.float32": torch.float32, "torch.uint8": torch.uint8}[dt]
        out[k] = blob[o:o + n].view(dtype).reshape(shape)
    return out


def _device(local):
    """int / "cuda:N" -> that GPU (NCCL); "cpu" -> host tensors (gloo, the CPU test-suite)."""
    return torch.device("cuda", local) if isinstance(local, int) else torch.device(local)


def broadcast_packed(packed, rank, world, local):
    """ONE broadcast of the repacked weight blob from rank 0 (plus a small object broadcast of its manifest).  `packed` is
    the dict Engine takes ({"dims", "alphas", "tensors", "max_pos"}) on rank 0 and ignored elsewhere; every rank returns
    that dict with tensors living on its own device (views into the received blob)."""
    import torch.distributed as td
    if world == 1:
        return packed
    dev = _device(local)
    meta = [None]
    if rank == 0:
        blob, manifest = _flatten(packed["tensors"])
        meta = [(packed["dims"], packed["alphas"], manifest, blob.numel(), packed["max_pos"])]
    td.broadcast_object_list(meta, src=0)
    dims, alphas, manifest, nbytes, max_pos = meta[0]
    dblob = blob.to(dev) if rank == 0 else torch.empty(nbytes, dtype=torch.uint8, device=dev)
    td.broadcast(dblob, src=0)
    return {"dims": dims, "alphas": alphas, "tensors": _unflatten(dblob, manifest), "max_pos": max_pos}


def build_or_receive_weights(size, rank, world, local, max_pos=4096, seed=0):
    """Rank 0 builds the synthetic reference-format checkpoints and repacks them; every other rank receives the packed
    blob with ONE broadcast (weights never touch the other ranks' host memory)."""
    from . import synth, weights
    packed = None
    if rank == 0:
        ar_sd, nar_sd, voc_sd = synth.make_ar_state(size, seed), synth.make_nar_state(size, seed + 1), synth.make_vocos_state(size, seed + 2)
        dims = weights.dims_from_state(ar_sd, nar_sd, voc_sd, size["n_text"])
        tensors, alphas = weights.repack(ar_sd, nar_sd, voc_sd, dims, max_pos=max_pos)
        packed = {"dims": dims, "alphas": alphas, "tensors": tensors, "max_pos": max_pos}
    return broadcast_packed(packed, rank, world, local)


def shard_utterances(costs, world):
    """Longest-processing-time-first assignment of utterances to ranks (mixed-length batches, BASELINE configs[3]).
    Returns a list of index lists, one per rank; deterministic."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    loads, out = [0.0] * world, [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda j: (loads[j], j))
        out[r].append(i)
        loads[r] += costs[i]
    return [sorted(x) for x in out]


def all_gather_waveforms(wavs, local):
    """All-gather of variable-length waveforms from shards of possibly DIFFERENT sizes (mixed-length batches, LPT sharding):
    the per-rank utterance counts first, then the lengths (padded to the largest count), then one padded
    (max_count, max_len) fp32 tensor per rank.
    Returns (list over ranks of (B_r, max_len) tensors, list over ranks of (B_r,) length tensors)."""
    import torch.distributed as td
    world = td.get_world_size()
    dev = _device(local)
    cnt = torch.tensor([len(wavs)], dtype=torch.int64, device=dev)
    cnts = [torch.empty_like(cnt) for _ in range(world)]
    td.all_gather(cnts, cnt)
    counts = [int(c.item()) for c in cnts]
    mc = max(counts)
    lens = torch.zeros(mc, dtype=torch.int64, device=dev)
    if len(wavs):
        lens[:len(wavs)] = torch.tensor([len(w) for w in wavs], dtype=torch.int64, device=dev)
    all_lens = [torch.empty_like(lens) for _ in range(world)]
    td.all_gather(all_lens, lens)
    mx = int(max(int(l.max()) for l in all_lens)) if mc else 0
    pad = torch.zeros(mc, mx, dtype=torch.float32, device=dev)
    for i, w in enumerate(wavs):
        pad[i, :len(w)] = torch.as_tensor(w, device=dev) if not torch.is_tensor(w) else w.to(dev)
    out = [torch.empty_like(pad) for _ in range(world)]
    td.all_gather(out, pad)
    return [o[:c] for o, c in zip(out, counts)], [l[:c] for l, c in zip(all_lens, counts)]

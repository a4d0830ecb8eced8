"""CPU suite: the multi-GPU host logic (sharding, weight-blob broadcast, waveform all-gather) on world_size-2 gloo."""
import os
import sys

import numpy as np
import torch
import torch.distributed as td
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_lpt_sharding_is_balanced_and_complete():
    from mars5_tts_b200 import dist
    rng = np.random.RandomState(0)
    costs = rng.randint(20, 121, size=256).astype(float) ** 2
    for world in (1, 2, 4, 8):
        shards = dist.shard_utterances(list(costs), world)
        flat = sorted(i for s in shards for i in s)
        assert flat == list(range(256))
        loads = [sum(costs[i] for i in s) for s in shards]
        assert max(loads) <= 1.02 * (sum(costs) / world) + max(costs)
    assert dist.shard_utterances([3.0, 1.0], 4) == [[0], [1], [], []]


def test_flatten_roundtrip():
    from mars5_tts_b200 import dist
    t = {"a": torch.randn(3, 5).half(), "b": torch.randn(7), "c": torch.randn(2, 2, 2).half()}
    blob, man = dist._flatten(t)
    back = dist._unflatten(blob, man)
    for k in t:
        assert back[k].dtype == t[k].dtype and torch.equal(back[k], t[k])


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    td.init_process_group("gloo", rank=rank, world_size=world)
    from mars5_tts_b200 import dist
    # weight blob: rank 0 flattens, everybody receives the same bytes
    meta = [None]
    if rank == 0:
        g = torch.Generator().manual_seed(0)
        tensors = {"w": torch.randn(33, 17, generator=g).half(), "b": torch.randn(9, generator=g)}
        blob, man = dist._flatten(tensors)
        meta = [(man, blob.numel())]
    td.broadcast_object_list(meta, src=0)
    man, n = meta[0]
    buf = blob if rank == 0 else torch.empty(n, dtype=torch.uint8)
    td.broadcast(buf, src=0)
    got = dist._unflatten(buf, man)
    # waveform gather on CPU tensors (same code path as the NCCL one, device-agnostic parts)
    wavs = [torch.full((100 + 10 * rank + i,), float(rank * 10 + i)) for i in range(2)]
    lens = torch.tensor([len(w) for w in wavs])
    all_lens = [torch.empty_like(lens) for _ in range(world)]
    td.all_gather(all_lens, lens)
    mx = int(max(int(l.max()) for l in all_lens))
    pad = torch.zeros(2, mx)
    for i, w in enumerate(wavs):
        pad[i, :len(w)] = w
    out = [torch.empty_like(pad) for _ in range(world)]
    td.all_gather(out, pad)
    q.put((rank, float(got["w"].float().sum()), [int(x) for l in all_lens for x in l], float(out[1 - rank][1, 0])))
    td.destroy_process_group()


def test_two_rank_gloo_broadcast_and_gather():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == res[1][1]                        # identical weights on both ranks
    assert res[0][2] == res[1][2] == [100, 101, 110, 111]
    assert res[0][3] == 11.0 and res[1][3] == 1.0         # each rank sees the other's second waveform

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
    """Drives the PRODUCT functions (dist.build_or_receive_weights / broadcast_packed / all_gather_waveforms /
    shard_utterances) over gloo with CPU tensors -- the same code bench.py and the engine run over NCCL."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    td.init_process_group("gloo", rank=rank, world_size=world)
    from mars5_tts_b200 import dist, synth
    packed = dist.build_or_receive_weights(synth.TINY, rank, world, "cpu", max_pos=64)
    chk = float(sum(v.double().abs().sum() for v in packed["tensors"].values()))
    n_t = len(packed["tensors"])
    # sharding: every rank computes the same deterministic assignment, takes its own shard, "synthesises" waveforms whose
    # content encodes the GLOBAL utterance id, and the all-gather must hand every rank every utterance exactly once
    costs = [float((7 * i) % 13 + 1) for i in range(9)]
    shards = dist.shard_utterances(costs, world)
    mine = shards[rank]
    wavs = [torch.full((50 + 3 * u,), float(u)) for u in mine]
    out, all_lens = dist.all_gather_waveforms(wavs, "cpu")
    seen = {}
    for r in range(world):
        for j, u in enumerate(shards[r]):
            n = int(all_lens[r][j])
            w = out[r][j, :n]
            assert n == 50 + 3 * u and bool((w == float(u)).all()), (r, j, u)
            assert bool((out[r][j, n:] == 0).all())
            seen[u] = n
    q.put((rank, chk, n_t, sorted(seen), packed["dims"]["ar_dim"], packed["max_pos"]))
    td.destroy_process_group()


def test_two_rank_gloo_broadcast_and_gather():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == res[1][1] and res[0][2] == res[1][2] > 50      # identical packed weights on both ranks
    assert res[0][3] == res[1][3] == list(range(9))                     # every utterance reached every rank exactly once
    assert res[0][4] == res[1][4] == 192 and res[0][5] == res[1][5] == 64

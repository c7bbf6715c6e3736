"""CPU: utterance sharding logic, single process and world_size-2 over gloo."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from promptttspp_b200 import dist as pdist


def test_shards_are_a_balanced_partition():
    g = torch.Generator().manual_seed(5)
    lengths = torch.randint(32, 257, (256,), generator=g).tolist()
    for world in (1, 2, 4, 8):
        shards = [pdist.shard_indices(lengths, world, r) for r in range(world)]
        flat = sorted(i for s in shards for i in s)
        assert flat == list(range(256))  # disjoint cover
        cost = [sum(lengths[i] for i in s) for s in shards]
        assert max(cost) - min(cost) <= 256  # round-robin over a sorted list: within one utterance
        assert shards == [pdist.shard_indices(lengths, world, r) for r in range(world)]  # deterministic
    with pytest.raises(ValueError):
        pdist.shard_indices(lengths, 2, 2)


def test_batch_shards_keep_composition_and_balance():
    """The batch list is cut once; ranks get whole batches (LPT on the padded cost): same composition at every world
    size, disjoint cover, balanced within one batch."""
    g = torch.Generator().manual_seed(6)
    lengths = torch.randint(64, 257, (256,), generator=g).tolist()
    batches = pdist.make_batches(range(256), lengths, 16)
    cost = lambda b: len(b) * max(lengths[i] for i in b)  # noqa: E731
    for world in (1, 2, 4, 8):
        shards = [pdist.shard_batches(batches, lengths, world, r) for r in range(world)]
        got = sorted(tuple(b) for s in shards for b in s)
        assert got == sorted(tuple(b) for b in batches)               # every batch exactly once, unchanged
        loads = [sum(cost(b) for b in s) for s in shards]
        assert max(loads) - min(loads) <= max(cost(b) for b in batches)
        assert shards == [pdist.shard_batches(batches, lengths, world, r) for r in range(world)]
    with pytest.raises(ValueError):
        pdist.shard_batches(batches, lengths, 2, 2)


def test_batches_group_neighbouring_lengths_and_ragged_tail():
    lengths = [5, 50, 7, 48, 6, 49, 100]
    batches = pdist.make_batches(range(7), lengths, 3)
    assert [len(b) for b in batches] == [3, 3, 1]
    assert batches[0] == [6, 1, 5] and batches[2] == [0]
    padded, lens = pdist.pad_batch([torch.arange(1, n + 1) for n in (3, 1)])
    assert padded.tolist() == [[1, 2, 3], [1, 0, 0]] and lens.tolist() == [3, 1]
    assert pdist.make_batches([], lengths, 4) == []


def _fake_synth(padded, lens, idx):
    # a per-utterance function of its own tokens only: 8 frames per phoneme, checksum of the ids
    return [(int(lens[b]) * 8, int(padded[b, : int(lens[b])].sum())) for b in range(len(idx))]


def _worker(rank, world, port, phonemes, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        res = pdist.synthesize_sharded(phonemes, _fake_synth, batch_size=4, world_size=world, rank=rank)
        counts = pdist.gather_frame_counts({i: r[0] for i, r in res.items()}, len(phonemes))
        q.put((rank, res, counts))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo_matches_single_process():
    g = torch.Generator().manual_seed(9)
    phonemes = [torch.randint(3, 90, (int(n),), generator=g) for n in torch.randint(4, 40, (19,), generator=g)]
    single = pdist.synthesize_sharded(phonemes, _fake_synth, batch_size=4)
    assert sorted(single) == list(range(19))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, phonemes, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    merged = {}
    for rank, res, counts in got:
        lens = [p.numel() for p in phonemes]
        mine = pdist.shard_batches(pdist.make_batches(range(19), lens, 4), lens, 2, rank)
        assert set(res) == {i for b in mine for i in b}
        merged.update(res)
        assert counts == [single[i][0] for i in range(19)]  # every rank sees every frame count
    assert merged == single  # sharding must not change any utterance's result

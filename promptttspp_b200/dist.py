"""Utterance-batch sharding for multi-GPU batched synthesis.

Utterances are independent in inference (eval-mode BatchNorm, per-utterance attention and masks), so the
path shards with no data-path collective: one process per GPU (torchrun), every rank synthesises its own
utterances; torch.distributed (NCCL over NVLink on GPUs, gloo in the CPU tests) is used only to exchange
lengths / results.  What is sharded is the list of BATCHES, not of utterances: the batches are cut once from the
length-sorted request list (neighbouring lengths share a batch, minimal padding) and then dealt to the ranks by
longest-processing-time-first on their padded size.  An utterance's result depends on the batch it sits in (the padded
frames of a batch are live inside the diffusion sampler), so keeping the batch composition
independent of the world size is what makes 1/2/4/8-GPU runs produce identical outputs.

The reference has no batched-inference caller at all (egs/proposed/bin/synthesize.py:132 loops one
utterance at a time); its trainer shards token-bucketed batches with x[rank::num_replicas]
(promptttspp/trainers/tts.py:138-142), which is the precedent for the round-robin deal.
"""
from typing import Callable, Dict, List, Sequence

import torch


def shard_indices(lengths: Sequence[int], world_size: int, rank: int) -> List[int]:
    """Indices of the utterances rank `rank` owns: sort by length (desc, stable), deal round-robin."""
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside world of {world_size}")
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    return order[rank::world_size]


def make_batches(indices: Sequence[int], lengths: Sequence[int], batch_size: int) -> List[List[int]]:
    """Cut a rank's (length-sorted) utterance list into batches of neighbouring lengths."""
    if batch_size < 1:
        raise ValueError("batch_size must be >= 1")
    idx = sorted(indices, key=lambda i: (-int(lengths[i]), i))
    return [idx[k:k + batch_size] for k in range(0, len(idx), batch_size)]


def shard_batches(batches: Sequence[Sequence[int]], lengths: Sequence[int], world_size: int, rank: int) -> List[List[int]]:
    """The batches rank `rank` owns.  Longest-processing-time-first on the padded cost len(batch) * longest utterance:
    batches in decreasing cost order go to the currently least loaded rank (ties: lowest rank) -- deterministic, within
    one batch of the optimum, and the batch COMPOSITION is the same for every world size."""
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside world of {world_size}")
    cost = [len(b) * max((int(lengths[i]) for i in b), default=0) for b in batches]
    order = sorted(range(len(batches)), key=lambda k: (-cost[k], k))
    load = [0] * world_size
    mine: List[int] = []
    for k in order:
        r = min(range(world_size), key=lambda q: (load[q], q))
        load[r] += cost[k]
        if r == rank:
            mine.append(k)
    return [list(batches[k]) for k in sorted(mine)]


def pad_batch(seqs: Sequence[torch.Tensor], device=None):
    """List of 1-D LongTensors -> (padded [B, Tmax] with 0, lengths [B])."""
    lens = torch.tensor([int(s.numel()) for s in seqs], dtype=torch.int64)
    out = torch.zeros(len(seqs), int(lens.max()) if len(seqs) else 0, dtype=torch.int64)
    for b, s in enumerate(seqs):
        out[b, : s.numel()] = s
    if device is not None:
        out, lens = out.to(device), lens.to(device)
    return out, lens


def synthesize_sharded(phonemes: Sequence[torch.Tensor], synth_batch: Callable, batch_size: int, world_size: int = 1,
                       rank: int = 0) -> Dict[int, object]:
    """Run `synth_batch(padded, lengths, utterance_indices) -> list of per-utterance results` over this rank's share
    of the batches (`shard_batches`).  Returns {utterance index: result} for the utterances this rank owns."""
    lengths = [int(p.numel()) for p in phonemes]
    batches = make_batches(range(len(phonemes)), lengths, batch_size)
    results: Dict[int, object] = {}
    for batch in shard_batches(batches, lengths, world_size, rank):
        padded, lens = pad_batch([phonemes[i] for i in batch])
        outs = synth_batch(padded, lens, batch)
        if len(outs) != len(batch):
            raise RuntimeError("synth_batch must return one result per utterance")
        for i, o in zip(batch, outs):
            results[i] = o
    return results


def gather_frame_counts(local: Dict[int, int], n_utts: int, group=None) -> List[int]:
    """All ranks learn every utterance's frame count (one small all_reduce; the only collective of the path)."""
    import torch.distributed as dist

    t = torch.zeros(n_utts, dtype=torch.int64)
    for i, n in local.items():
        t[i] = int(n)
    if dist.is_available() and dist.is_initialized():
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else t.device
        t = t.to(dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        t = t.cpu()
    return t.tolist()

"""Batched synthesis front end -- the caller of `infer_batch` the reference does not have.

`egs/proposed/bin/synthesize.py:132-193` loops over utterances one at a time (`model.infer`, `vocoder`, `.cpu()` per
utterance).  `BatchedSynthesizer` takes a list of requests, cuts it into length-bucketed batches under a padded-token
budget (the trainer's `batch_by_size` rule, `promptttspp/datasets/utils.py:55-112`: a batch is full when
`(n + 1) * longest > max_tokens` or `n == max_sentences`), runs acoustic model -> f0 post-processing -> vocoder per
batch exactly as `app.py:56-81` does for one utterance, and returns the waveforms in request order.  With
`world_size > 1` each rank serves its round-robin share of the length-sorted requests (`dist.shard_indices`).
"""
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import torch

from .dist import pad_batch, shard_indices
from .utils.model import lowpass_filter


def token_buckets(lengths: Sequence[int], max_tokens: int, max_sentences: Optional[int] = None,
                  indices: Optional[Sequence[int]] = None) -> List[List[int]]:
    """Length-sorted greedy batching under a PADDED token budget: a batch of n requests costs n * longest tokens.
    Requests longer than `max_tokens` get a batch of their own.  Returns lists of request indices."""
    if max_tokens < 1:
        raise ValueError("max_tokens must be >= 1")
    cap = max_sentences if max_sentences is not None else 1 << 30
    if cap < 1:
        raise ValueError("max_sentences must be >= 1")
    idx = list(range(len(lengths))) if indices is None else list(indices)
    idx.sort(key=lambda i: (-int(lengths[i]), i))
    batches, cur, longest = [], [], 0
    for i in idx:
        n = int(lengths[i])
        new_longest = max(longest, n)
        if cur and ((len(cur) + 1) * new_longest > max_tokens or len(cur) >= cap):
            batches.append(cur)
            cur, new_longest = [], n
        cur.append(i)
        longest = new_longest
    if cur:
        batches.append(cur)
    return batches


@dataclass
class MelStats:
    """mean / std of stats.yaml (compute_mel.py:60-68); the acoustic model emits normalised mels (app.py:80)."""
    mean: float = 0.0
    std: float = 1.0


class BatchedSynthesizer:
    def __init__(self, model, vocoder, stats: MelStats = MelStats(), max_tokens: int = 4096,
                 max_sentences: Optional[int] = 32, use_lowpass: bool = True, noise_scale: float = 0.5,
                 frame_rate: int = 100):
        self.model, self.vocoder, self.stats = model, vocoder, stats
        self.max_tokens, self.max_sentences = max_tokens, max_sentences
        self.use_lowpass, self.noise_scale, self.frame_rate = use_lowpass, noise_scale, frame_rate
        self.f0_aware = hasattr(vocoder, "m_source")

    @torch.no_grad()
    def synthesize(self, phonemes: Sequence[torch.Tensor], style_prompts, world_size: int = 1, rank: int = 0,
                   device=None) -> Dict[int, torch.Tensor]:
        """phonemes: list of 1-D LongTensors (text_to_sequence output); style_prompts: list of strings or a tensor
        [N, prompt_dim] of sentence embeddings.  Returns {request index: waveform [samples] on the host} for the requests
        this rank owns (all of them when world_size == 1)."""
        device = torch.device(device) if device is not None else next(self.model.parameters()).device
        lengths = [int(p.numel()) for p in phonemes]
        mine = shard_indices(lengths, world_size, rank)
        hop = getattr(self.vocoder, "hop", 240)
        out: Dict[int, torch.Tensor] = {}
        for batch in token_buckets(lengths, self.max_tokens, self.max_sentences, mine):
            padded, lens = pad_batch([phonemes[i] for i in batch])
            padded, lens = padded.pin_memory().to(device, non_blocking=True), lens.pin_memory().to(device, non_blocking=True)
            if isinstance(style_prompts, torch.Tensor):
                prompts = style_prompts[batch].to(device)
            else:
                prompts = [style_prompts[i] for i in batch]
            mel, log_cf0, vuv, frame_len = self.model.infer_batch(padded, lens, style_prompt=prompts, use_max=True,
                                                                  noise_scale=self.noise_scale, return_f0=True)
            dec = mel * self.stats.std + self.stats.mean                       # app.py:80
            if self.f0_aware:
                if self.use_lowpass:
                    # app.py:76-77, per utterance: every row is filtered over its own frames only
                    log_cf0 = lowpass_filter(log_cf0, self.frame_rate, cutoff=20, lengths=frame_len)
                f0 = log_cf0.exp()
                f0[vuv < 0.5] = 0                                              # app.py:78-79
                wav = self.vocoder(dec, f0)
            else:
                wav = self.vocoder(dec)
            wav = wav.squeeze(1).cpu()                                          # one D2H per batch (app.py:81)
            n_frames = frame_len.cpu().long().tolist()
            for b, i in enumerate(batch):
                out[i] = wav[b, : n_frames[b] * hop].clone()
        return out

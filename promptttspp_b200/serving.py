"""Batched synthesis front end -- the caller of `infer_batch` the reference does not have.

`egs/proposed/bin/synthesize.py:132-193` loops over utterances one at a time (`model.infer`, `vocoder`, `.cpu()` per
utterance).  `BatchedSynthesizer` takes a list of requests, cuts it into length-bucketed batches under a padded-token
budget (the trainer's `batch_by_size` rule, `promptttspp/datasets/utils.py:55-112`: a batch is full when
`(n + 1) * longest > max_tokens` or `n == max_sentences`), runs acoustic model -> f0 post-processing -> vocoder per
batch exactly as `app.py:56-81` does for one utterance, and returns the waveforms in request order.  With
`world_size > 1` each rank serves its share of the BATCHES (`dist.shard_batches`), so an utterance sits in the same batch
whatever the world size.
"""
import csv
import queue
import struct
import threading
from dataclasses import dataclass
from pathlib import Path
from typing import Callable, Dict, Iterable, List, Optional, Sequence

import torch

from .dist import pad_batch, shard_batches
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


def cost_buckets(lengths: Sequence[int], max_tokens: int, max_sentences: Optional[int] = None,
                 indices: Optional[Sequence[int]] = None, min_tokens: int = 2048,
                 overhead_tokens: int = 256) -> List[List[int]]:
    """Length-sorted batching that MINIMISES the padded work instead of filling the budget greedily.

    Padded rows are live in the diffusion sampler and the vocoder (a batch of n requests costs n * longest), so one
    budget-filling batch of mixed lengths wastes up to half of its frames (cfg4: 32 requests of 32..256 phonemes in one
    8192-token batch = 1.8x the valid tokens).  The batch boundaries over the length-sorted requests are chosen by dynamic
    programming over  cost(batch) = max(n * longest, min_tokens) + overhead_tokens  -- `min_tokens`: below this a batch
    no longer fills the GPU (the fused DiffNet kernel needs ~74 units of 256 frames), `overhead_tokens`: the per-batch
    text side, launches and host round trip -- under the same constraints as `token_buckets` (n * longest <= max_tokens,
    n <= max_sentences, an oversize request alone).  Deterministic: the result depends on the multiset of lengths only,
    so every rank of a sharded job cuts the same batches."""
    if max_tokens < 1:
        raise ValueError("max_tokens must be >= 1")
    cap = max_sentences if max_sentences is not None else 1 << 30
    if cap < 1:
        raise ValueError("max_sentences must be >= 1")
    idx = list(range(len(lengths))) if indices is None else list(indices)
    idx.sort(key=lambda i: (-int(lengths[i]), i))
    n_req = len(idx)
    if n_req == 0:
        return []
    ln = [int(lengths[i]) for i in idx]
    inf = float("inf")
    best = [0.0] + [inf] * n_req  # best[k]: minimal cost of batching the k longest requests
    cut = [0] * (n_req + 1)       # cut[k]: start of the last batch in that optimum
    for k in range(1, n_req + 1):
        for j in range(k - 1, -1, -1):  # last batch = sorted requests j .. k-1, its longest is ln[j]
            n = k - j
            if n > cap or (n > 1 and n * ln[j] > max_tokens):
                break
            c = best[j] + max(n * ln[j], min_tokens) + overhead_tokens
            if c < best[k]:
                best[k], cut[k] = c, j
    batches, k = [], n_req
    while k > 0:
        batches.append(idx[cut[k]:k])
        k = cut[k]
    batches.reverse()
    return batches


@dataclass
class MelStats:
    """mean / std of stats.yaml (compute_mel.py:60-68); the acoustic model emits normalised mels (app.py:80)."""
    mean: float = 0.0
    std: float = 1.0

    @classmethod
    def from_yaml(cls, path) -> "MelStats":
        """`stats.yaml` as egs/proposed/bin/compute_mel.py:60-68 writes it (keys min, max, mean, std, var) and
        app.py:133 / synthesize.py:102 read it."""
        import yaml

        with open(path, "r") as f:
            d = yaml.safe_load(f)
        return cls(mean=float(d["mean"]), std=float(d["std"]))


# ---- the evaluation-set formats of egs/proposed/bin/synthesize.py -----------------------------------------------------
EVAL_COLUMNS = ["spk_id", "item_name", "gender", "pitch", "speaking_speed", "energy", "style_prompt", "style_prompt_key",
                "seq"]  # synthesize.py:118-130


def read_eval_csv(path) -> List[dict]:
    """The label file of synthesize.py:131-135: one request per row; `seq` (space-separated phoneme ids) becomes a
    LongTensor under "phonemes"."""
    rows = []
    with open(path, newline="") as f:
        for r in csv.DictReader(f):
            missing = [c for c in EVAL_COLUMNS if c not in r]
            if missing:
                raise ValueError(f"{path}: missing columns {missing}")
            row = {c: r[c] for c in EVAL_COLUMNS}
            row["phonemes"] = torch.tensor([int(t) for t in r["seq"].split()], dtype=torch.int64)
            rows.append(row)
    return rows


def read_prompt_candidate(path) -> Dict[str, List[str]]:
    """`style_key|prompt a; prompt b; ...` lines -> {style_key: [lower-cased prompts]} (synthesize.py:64-76)."""
    out = {}
    with open(path) as f:
        for line in f:
            line = line.rstrip("\n")
            if not line:
                continue
            key, prompts = line.split("|", 1)
            out[key] = [p.lower().strip() for p in prompts.split(";")]
    return out


def read_spk_prompt_candidate(path) -> Dict[str, List[str]]:
    """`spk|word,word,...` lines -> {spk: [words]} (synthesize.py:79-84)."""
    out = {}
    with open(path) as f:
        for line in f:
            line = line.rstrip("\n")
            if line:
                spk, words = line.split("|", 1)
                out[spk] = words.split(",")
    return out


def add_spk_prompt(style_prompt: str, words: str) -> str:
    """synthesize.py:87-90"""
    return f"{style_prompt}. The speaker identity can be described as {words}."


def eval_prompts(rows: Sequence[dict], prompt_candidate, spk_prompt_candidate, use_spk_prompt=True) -> List[str]:
    """The prompt string of every evaluation row (synthesize.py:137-150)."""
    prompts = []
    for r in rows:
        style = prompt_candidate[r["style_prompt_key"]][0]
        spk = r["spk_id"]
        if use_spk_prompt and spk in spk_prompt_candidate:
            style = add_spk_prompt(style, ", ".join(spk_prompt_candidate[spk]))
        prompts.append(style)
    return prompts


def write_wav(path, wav: torch.Tensor, sample_rate: int):
    """Mono float32 RIFF/WAVE (format tag 3), what torchaudio.save(path, float tensor, sr) produces by default
    (synthesize.py:195, :214) -- written with the standard library only."""
    data = wav.detach().to(torch.float32).reshape(-1).cpu().numpy().astype("<f4").tobytes()
    hdr = b"RIFF" + struct.pack("<I", 50 + len(data)) + b"WAVE"
    hdr += b"fmt " + struct.pack("<IHHIIHHH", 18, 3, 1, sample_rate, sample_rate * 4, 4, 32, 0)
    hdr += b"fact" + struct.pack("<II", 4, len(data) // 4)
    hdr += b"data" + struct.pack("<I", len(data))
    with open(path, "wb") as f:
        f.write(hdr + data)


class WavWriter:
    """Streamed write-back: finished waveforms are handed to a background thread that encodes and writes them while the
    GPU already runs the next batch (the reference's loop blocks on torchaudio.save per utterance)."""

    def __init__(self, out_dir, sample_rate=24000, namer: Optional[Callable[[int], str]] = None, depth: int = 64):
        self.out_dir, self.sample_rate = Path(out_dir), sample_rate
        self.namer = namer or (lambda i: f"{i:06d}.wav")
        self.q: "queue.Queue" = queue.Queue(maxsize=depth)
        self.error = None
        self.written: List[Path] = []
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _run(self):
        while True:
            item = self.q.get()
            if item is None:
                return
            try:
                i, wav = item
                path = self.out_dir / self.namer(i)
                path.parent.mkdir(parents=True, exist_ok=True)
                write_wav(path, wav, self.sample_rate)
                self.written.append(path)
            except Exception as e:  # surfaced by close()
                self.error = e

    def put(self, index: int, wav: torch.Tensor):
        self.q.put((index, wav))

    def close(self):
        self.q.put(None)
        self.thread.join()
        if self.error is not None:
            raise self.error
        return self.written


class BatchedSynthesizer:
    def __init__(self, model, vocoder, stats: MelStats = MelStats(), max_tokens: int = 4096,
                 max_sentences: Optional[int] = 32, use_lowpass: bool = True, noise_scale: float = 0.5,
                 frame_rate: int = 100, seed: Optional[int] = None, bucketing: str = "cost",
                 min_tokens: int = 2048, overhead_tokens: int = 256):
        """bucketing: "cost" (default) = `cost_buckets`, batch boundaries minimising the padded work; "greedy" =
        `token_buckets`, the trainer's budget-filling rule."""
        if bucketing not in ("cost", "greedy"):
            raise ValueError(f"bucketing must be 'cost' or 'greedy', got {bucketing!r}")
        self.model, self.vocoder, self.stats = model, vocoder, stats
        self.max_tokens, self.max_sentences = max_tokens, max_sentences
        self.bucketing, self.min_tokens, self.overhead_tokens = bucketing, min_tokens, overhead_tokens
        self.use_lowpass, self.noise_scale, self.frame_rate = use_lowpass, noise_scale, frame_rate
        self.f0_aware = hasattr(vocoder, "m_source")
        # seed: every batch draws its noise from torch.manual_seed(seed + first request index of the batch), so a
        # request's audio does not depend on which rank (or in which order) its batch is served
        self.seed = seed

    @torch.no_grad()
    def synthesize(self, phonemes: Sequence[torch.Tensor], style_prompts, world_size: int = 1, rank: int = 0,
                   device=None, sink: Optional[Callable[[int, torch.Tensor], None]] = None) -> Dict[int, torch.Tensor]:
        """phonemes: list of 1-D LongTensors (text_to_sequence output); style_prompts: list of strings or a tensor
        [N, prompt_dim] of sentence embeddings.  Returns {request index: waveform [samples] on the host} for the requests
        this rank owns (all of them when world_size == 1).  `sink(index, waveform)` -- e.g. `WavWriter.put` -- receives
        every waveform as soon as its batch has left the device (streamed write-back)."""
        device = torch.device(device) if device is not None else next(self.model.parameters()).device
        lengths = [int(p.numel()) for p in phonemes]
        hop = getattr(self.vocoder, "hop", 240)
        out: Dict[int, torch.Tensor] = {}
        # batches are cut from ALL requests and then dealt to the ranks: their composition does not depend on world_size
        if self.bucketing == "cost":
            buckets = cost_buckets(lengths, self.max_tokens, self.max_sentences, min_tokens=self.min_tokens,
                                   overhead_tokens=self.overhead_tokens)
        else:
            buckets = token_buckets(lengths, self.max_tokens, self.max_sentences)
        for batch in shard_batches(buckets, lengths, world_size, rank):
            if self.seed is not None:
                torch.manual_seed(self.seed + batch[0])
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
                if sink is not None:
                    sink(i, out[i])
        return out

    def synthesize_texts(self, texts: Iterable[str], style_prompts, **kw) -> Dict[int, torch.Tensor]:
        """Phoneme strings (g2p output, `HH AH0 L OW1 ...`) instead of id tensors: text.eng.text_to_sequence per request
        (app.py:60-62), then `synthesize`."""
        from .text import text_to_sequence

        return self.synthesize([torch.tensor(text_to_sequence(t), dtype=torch.int64) for t in texts], style_prompts, **kw)

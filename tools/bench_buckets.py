"""cfg4 (32 variable-length requests, text ids -> waveform) under different bucketing cost models."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
from promptttspp_b200.serving import BatchedSynthesizer, MelStats, cost_buckets  # noqa: E402

torch.set_grad_enabled(False)
dev = torch.device("cuda", 0)
model, voc = bench.build_models(dev)
phonemes, emb = bench.var_len_requests(32, seed=41)
L = [int(p.numel()) for p in phonemes]
for kw in (dict(min_tokens=1536, overhead_tokens=512), dict(min_tokens=1536, overhead_tokens=384),
           dict(min_tokens=1280, overhead_tokens=512), dict(min_tokens=1792, overhead_tokens=512),
           dict(min_tokens=1536, overhead_tokens=768), dict(min_tokens=1024, overhead_tokens=768),
           dict(min_tokens=2048, overhead_tokens=512)):
    srv = BatchedSynthesizer(model, voc, MelStats(mean=-5.0, std=2.0), max_tokens=8192, max_sentences=32, **kw)

    def run():
        torch.manual_seed(11)
        return srv.synthesize(phonemes, emb, device=dev)

    run()
    ms, _ = bench._event_time(run, dev)
    b = [len(x) for x in cost_buckets(L, 8192, 32, min_tokens=kw.get("min_tokens", 2048),
                                      overhead_tokens=kw.get("overhead_tokens", 256))] if kw.get("bucketing") != "greedy" else [32]
    print(f"{kw}: {ms:7.1f} ms  batches {b}")

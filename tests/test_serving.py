"""Batched-synthesis front end: bucketing rule on the CPU, end-to-end on the GPU."""
import pytest
import torch

from promptttspp_b200.serving import cost_buckets, token_buckets


def test_token_buckets_cover_budget_and_order():
    g = torch.Generator().manual_seed(0)
    lengths = torch.randint(5, 257, (97,), generator=g).tolist()
    batches = token_buckets(lengths, max_tokens=1024, max_sentences=8)
    flat = [i for b in batches for i in b]
    assert sorted(flat) == list(range(97))                       # a disjoint cover
    for b in batches:
        longest = max(lengths[i] for i in b)
        assert len(b) <= 8 and (len(b) * longest <= 1024 or len(b) == 1)
    firsts = [lengths[b[0]] for b in batches]
    assert firsts == sorted(firsts, reverse=True)                # neighbouring lengths share a batch
    # the padded-token waste of bucketing is far below one arbitrary split of the same sizes
    waste = sum(len(b) * max(lengths[i] for i in b) - sum(lengths[i] for i in b) for b in batches)
    naive = [list(range(k, min(k + 8, 97))) for k in range(0, 97, 8)]
    waste_naive = sum(len(b) * max(lengths[i] for i in b) - sum(lengths[i] for i in b) for b in naive)
    assert waste < 0.35 * waste_naive
    assert token_buckets([300, 10], max_tokens=128) == [[0], [1]]  # an oversize request gets its own batch
    assert token_buckets([], 64) == []
    with pytest.raises(ValueError):
        token_buckets([1], 0)
    # a rank's shard only
    assert sorted(i for b in token_buckets(lengths, 1024, 8, indices=[3, 5, 7]) for i in b) == [3, 5, 7]


def test_cost_buckets_minimise_padded_work():
    import random

    def cost(batches, lengths, min_tokens=2048, overhead=256):
        return sum(max(len(b) * max(lengths[i] for i in b), min_tokens) + overhead for b in batches)

    rng = random.Random(3)
    for trial in range(20):
        n = rng.randint(1, 70)
        lengths = [rng.randint(1, 300) for _ in range(n)]
        for max_tokens, cap in ((8192, 32), (1024, 8), (200, None)):
            cb = cost_buckets(lengths, max_tokens, cap)
            tb = token_buckets(lengths, max_tokens, cap)
            assert sorted(i for b in cb for i in b) == list(range(n))             # every request exactly once
            for b in cb:
                longest = max(lengths[i] for i in b)
                assert len(b) == 1 or len(b) * longest <= max_tokens                # same budget rule as the greedy cutter
                assert cap is None or len(b) <= cap
                assert all(lengths[b[0]] >= lengths[i] for i in b)                  # length-sorted, longest first
            assert cost(cb, lengths) <= cost(tb, lengths)                           # never worse than budget filling
            assert cb == cost_buckets(lengths, max_tokens, cap)                     # deterministic
    # cfg4's shape: 32 requests of 32..256 phonemes -- one budget-filling batch pads to 8192 tokens, the optimum does not
    lengths = [32 + 7 * i for i in range(32)]
    assert len(token_buckets(lengths, 8192, 32)) == 1
    cb = cost_buckets(lengths, 8192, 32)
    assert len(cb) >= 2 and sum(len(b) * max(lengths[i] for i in b) for b in cb) < 0.85 * 32 * max(lengths)
    assert cost_buckets([300, 10], max_tokens=128) == [[0], [1]]
    assert cost_buckets([], 64) == []
    assert sorted(i for b in cost_buckets(lengths, 1024, 8, indices=[3, 5, 7]) for i in b) == [3, 5, 7]


@pytest.mark.gpu
def test_batched_synthesizer_matches_single_calls():
    from golden_cases import F0_KWARGS
    from promptttspp_b200.modules.prompt_encoder import FixedPromptEmbedding
    from promptttspp_b200.serving import BatchedSynthesizer, MelStats
    from promptttspp_b200.utils.synthetic import build_acoustic, build_vocoder_f0, synthetic_state_dict

    torch.set_grad_enabled(False)
    g = torch.Generator().manual_seed(4)
    phonemes = [torch.randint(3, 90, (n,), generator=g) for n in (12, 5, 9, 12, 7)]
    emb = torch.randn(len(phonemes), 768, generator=g)
    model = build_acoustic(bert=FixedPromptEmbedding(emb), K_step=6)
    model.load_state_dict(synthetic_state_dict(model, seed=1234, frames_per_phoneme=3.0), strict=True)
    model = model.cuda().eval()
    voc = build_vocoder_f0(**F0_KWARGS)
    voc.load_state_dict(synthetic_state_dict(voc, seed=4322), strict=True)
    voc = voc.cuda().eval()
    srv = BatchedSynthesizer(model, voc, MelStats(mean=-5.0, std=2.0), max_tokens=30, max_sentences=4)
    torch.manual_seed(0)
    out = srv.synthesize(phonemes, emb)
    assert sorted(out) == list(range(len(phonemes)))
    for i, w in out.items():
        assert w.dim() == 1 and w.numel() % 240 == 0 and w.numel() > 0 and torch.isfinite(w).all()
    # two ranks serve disjoint shares
    a = srv.synthesize(phonemes, emb, world_size=2, rank=0)
    b = srv.synthesize(phonemes, emb, world_size=2, rank=1)
    assert sorted(list(a) + list(b)) == list(range(len(phonemes)))
    # same seed, same batch composition -> identical audio (the style draw depends on the batch, so shares differ)
    torch.manual_seed(0)
    again = srv.synthesize(phonemes, emb)
    assert all(torch.equal(out[i], again[i]) for i in out)
    # seeded per batch + batch-level sharding: the merged output of two ranks equals the single-rank output bit for bit
    srv2 = BatchedSynthesizer(model, voc, MelStats(mean=-5.0, std=2.0), max_tokens=30, max_sentences=4, seed=77)
    one = srv2.synthesize(phonemes, emb)
    two = dict(srv2.synthesize(phonemes, emb, world_size=2, rank=1))
    two.update(srv2.synthesize(phonemes, emb, world_size=2, rank=0))
    assert sorted(two) == sorted(one) and all(torch.equal(one[i], two[i]) for i in one)


@pytest.mark.gpu
def test_batched_f0_path_equals_per_utterance_f0_path():
    """the f0 post-processing (low-pass -> exp -> unvoiced zeroing, app.py:76-79) of a ragged batch must equal
    the per-utterance flow on each row's own frames, and the vocoder's output for a row must not depend on the padding
    of the batch (valid samples compared)."""
    from golden_cases import F0_KWARGS
    from promptttspp_b200.utils.model import lowpass_filter
    from promptttspp_b200.utils.synthetic import build_vocoder_f0, synthetic_state_dict

    torch.set_grad_enabled(False)
    voc = build_vocoder_f0(**F0_KWARGS)
    voc.load_state_dict(synthetic_state_dict(voc, seed=4322), strict=True)
    voc = voc.cuda().eval()
    g = torch.Generator().manual_seed(8)
    lens = [96, 41, 70]
    B, T = len(lens), max(lens)
    log_cf0, vuv = torch.zeros(B, 1, T), torch.zeros(B, 1, T)
    for b, n in enumerate(lens):
        log_cf0[b, 0, :n] = 5.0 + 0.2 * torch.randn(n, generator=g)
        vuv[b, 0, :n] = (torch.rand(n, generator=g) > 0.3).float()
    flen = torch.tensor(lens)
    lf = lowpass_filter(log_cf0.cuda(), 100, cutoff=20, lengths=flen.cuda())
    f0 = lf.exp()
    f0[vuv.cuda() < 0.5] = 0
    for b, n in enumerate(lens):
        s = lowpass_filter(log_cf0[b:b + 1, :, :n].contiguous().cuda(), 100, cutoff=20).exp()
        s[vuv[b:b + 1, :, :n].cuda() < 0.5] = 0
        assert torch.equal(f0[b:b + 1, :, :n], s)


def test_eval_formats_roundtrip(tmp_path):
    """stats.yaml, the evaluation CSV, the prompt candidate files (synthesize.py:64-150) and the wav writer."""
    import struct

    from promptttspp_b200.serving import (EVAL_COLUMNS, MelStats, WavWriter, eval_prompts, read_eval_csv,
                                          read_prompt_candidate, read_spk_prompt_candidate)

    (tmp_path / "stats.yaml").write_text("min: -11.5\nmax: 2.0\nmean: -5.25\nstd: 2.125\nvar: 4.515625\n")
    st = MelStats.from_yaml(tmp_path / "stats.yaml")
    assert (st.mean, st.std) == (-5.25, 2.125)
    (tmp_path / "eval.csv").write_text(
        ",".join(EVAL_COLUMNS + ["extra"]) + "\n"
        + "p1,utt_a,F,high,fast,low,calm voice,k1,1 12 40 7 2,x\n"
        + "p9,utt_b,M,low,slow,high,angry voice,k2,1 5 2,y\n")
    rows = read_eval_csv(tmp_path / "eval.csv")
    assert [r["item_name"] for r in rows] == ["utt_a", "utt_b"] and rows[0]["phonemes"].tolist() == [1, 12, 40, 7, 2]
    (tmp_path / "prompts.txt").write_text("k1|A Calm Voice ; quiet\nk2|An angry voice\n")
    (tmp_path / "spk.txt").write_text("p1|bright,young\n")
    pc, sc = read_prompt_candidate(tmp_path / "prompts.txt"), read_spk_prompt_candidate(tmp_path / "spk.txt")
    assert pc["k1"] == ["a calm voice", "quiet"] and sc == {"p1": ["bright", "young"]}
    prompts = eval_prompts(rows, pc, sc)
    assert prompts == ["a calm voice. The speaker identity can be described as bright, young.", "an angry voice"]
    assert eval_prompts(rows, pc, sc, use_spk_prompt=False)[0] == "a calm voice"
    w = WavWriter(tmp_path / "out", 24000, namer=lambda i: f"{rows[i]['spk_id']}/{rows[i]['item_name']}.wav")
    wav = torch.linspace(-1, 1, 480)
    w.put(0, wav)
    w.put(1, wav[:100])
    files = w.close()
    assert sorted(p.name for p in files) == ["utt_a.wav", "utt_b.wav"]
    raw = (tmp_path / "out" / "p1" / "utt_a.wav").read_bytes()
    assert raw[:4] == b"RIFF" and raw[8:12] == b"WAVE" and struct.unpack("<H", raw[20:22])[0] == 3
    body = torch.frombuffer(bytearray(raw[-480 * 4:]), dtype=torch.float32)
    assert torch.equal(body, wav)
    try:
        import torchaudio

        back, sr = torchaudio.load(str(tmp_path / "out" / "p1" / "utt_a.wav"))
        assert sr == 24000 and torch.equal(back.view(-1), wav)
    except (ImportError, RuntimeError):
        pass

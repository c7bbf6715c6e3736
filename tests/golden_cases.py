"""Case definitions shared by tests/golden/make_golden.py (reference side) and the parity tests.

Everything here is regenerated from seeds with CPU generators, so only the reference's outputs
need to be stored in tests/golden/*.npz.
"""
import torch

from promptttspp_b200.models.prompttts_mdn_v2_final.model import InferNoise

_randn = torch.randn  # bound at import: make_golden.py patches torch.randn while the reference runs

ACOUSTIC_CASES = {
    # B=3 ragged batch, legacy rel-pos (the released demo checkpoint's flavour)
    "legacy_b3": dict(api="infer_batch", rel_pos_type="legacy", lengths=[10, 7, 4], weight_seed=1234,
                      frames_per_phoneme=3.0, input_seed=2, noise_seed=102, noise_scale=1.0, K_step=100),
    # new rel-pos flavour (training yaml), B=2
    "new_b2": dict(api="infer_batch", rel_pos_type="new", lengths=[6, 9], weight_seed=1235,
                   frames_per_phoneme=3.0, input_seed=3, noise_seed=103, noise_scale=0.5, K_step=100),
    # single utterance through .infer (app.py path), BOS/EOS framed like text_to_sequence
    "legacy_infer": dict(api="infer", rel_pos_type="legacy", lengths=[16], weight_seed=1234,
                         frames_per_phoneme=3.0, input_seed=0, noise_seed=100, noise_scale=0.5, K_step=100),
}

# Benchmark-scale cases: large enough that the kernels bench.py runs are the kernels under test --
# >= 9.5 k padded frames selects the cta_group::2 DiffNet kernels at cfg2's shape class (100 steps, legacy rel-pos).
ACOUSTIC_LARGE_CASES = {
    "legacy_b8_bench": dict(api="infer_batch", rel_pos_type="legacy", lengths=[131, 160, 97, 152, 118, 143, 104, 126],
                            weight_seed=1234, frames_per_phoneme=8.0, input_seed=21, noise_seed=121, noise_scale=1.0,
                            K_step=100),
}
# cfg2's text side exactly (B=16, Tx in [128, 256], one row at 256): >= 3 k phonemes whose integer durations must be
# bit-exact; the reference runs only up to the duration predictor
TEXT_CASES = {
    "cfg2_text": dict(api="infer_batch", rel_pos_type="legacy", weight_seed=1234, frames_per_phoneme=8.0, input_seed=2,
                      noise_seed=122, noise_scale=1.0, K_step=100,
                      lengths=[256, 203, 141, 187, 250, 129, 233, 176, 198, 160, 221, 149, 244, 135, 212, 169]),
}

# the same text side with the training yaml's rel-pos flavour (RelPositionalEncoding, 2T-1 position rows): exercises the
# two-window path of the fused attention kernel at T = 256
TEXT_CASES["cfg2_text_new"] = dict(TEXT_CASES["cfg2_text"], rel_pos_type="new", weight_seed=1235, input_seed=4,
                                   noise_seed=123)

VOCODER_CASES = {
    "b2_t12": dict(weight_seed=4321, input_seed=3, B=2, T=12, remove_weight_norm=True),
    "b1_t33": dict(weight_seed=4321, input_seed=4, B=1, T=33),
}
# cfg3's per-utterance shape (1024 frames -> 245 760 samples): the reference's own output at the benchmarked length
VOCODER_LARGE_CASES = {
    "b1_t1024": dict(weight_seed=4321, input_seed=3, B=1, T=1024),
}

# F0-aware vocoder (conf/vocoder/bigvgan_f0.yaml): mel + frame-level f0 with unvoiced stretches + injected source noise
VOCODER_F0_CASES = {
    "b2_t12": dict(weight_seed=4322, input_seed=5, noise_seed=205, B=2, T=12),
    "b1_t40": dict(weight_seed=4322, input_seed=6, noise_seed=206, B=1, T=40),
}
F0_KWARGS = dict(sampling_rate=24000, harmonic_num=8)

AA_CASES = {"c4_l37": (2, 4, 37), "c32_l3": (1, 32, 3), "c3_l1": (1, 3, 1), "c8_l64": (1, 8, 64)}


def acoustic_inputs(case):
    g = torch.Generator().manual_seed(case["input_seed"])
    lengths = torch.tensor(case["lengths"], dtype=torch.int64)
    B, Tx = len(case["lengths"]), int(lengths.max())
    phoneme = torch.zeros(B, Tx, dtype=torch.int64)
    for b, n in enumerate(case["lengths"]):
        ids = torch.randint(3, 90, (n,), generator=g)
        if case["api"] == "infer":
            ids[0], ids[-1] = 1, 2  # BOS / EOS (promptttspp/text/eng.py:117)
        phoneme[b, :n] = ids
    cls_emb = torch.randn(B, 768, generator=g)
    return phoneme, lengths, cls_emb


def golden_noise(case, B, Ty, C=256, mel=80):
    """z_style always; x_T / z once Ty is known (None otherwise)."""
    g = torch.Generator().manual_seed(case["noise_seed"])
    z_style = _randn(B, 1, C, generator=g)
    if Ty is None:
        return InferNoise(z_style, None, None)
    x_T = _randn(B, mel, Ty, generator=g)
    z = _randn(case["K_step"], B, mel, Ty, generator=g)
    return InferNoise(z_style, x_T, z)


def vocoder_inputs(case):
    g = torch.Generator().manual_seed(case["input_seed"])
    mel = torch.randn(case["B"], 80, case["T"], generator=g) * 2.0 - 5.0
    return mel.clamp(-11.5, 2.0)


def vocoder_f0_inputs(case, hop=240):
    """mel, f0 [B, 1, T] (Hz; about a third of the frames unvoiced = 0), rand_ini [B, 9], noise [B, L, 9]."""
    mel = vocoder_inputs(case)
    g = torch.Generator().manual_seed(case["input_seed"] + 1000)
    B, T = case["B"], case["T"]
    f0 = 80.0 + 320.0 * torch.rand(B, 1, T, generator=g)
    f0 = f0 * (torch.rand(B, 1, T, generator=g) > 0.33).float()
    gn = torch.Generator().manual_seed(case["noise_seed"])
    H = F0_KWARGS["harmonic_num"] + 1
    rand_ini = torch.rand(B, H, generator=gn)
    noise = _randn(B, T * hop, H, generator=gn)
    return mel, f0, rand_ini, noise


def lowpass_inputs():
    """log-f0-like contours [B, 1, T]: a long ragged batch, a single row, and one short enough for the pass-through."""
    g = torch.Generator().manual_seed(77)
    return [5.0 + 0.5 * torch.randn(3, 1, 400, generator=g).cumsum(-1) * 0.1, 4.0 + torch.randn(1, 1, 57, generator=g),
            torch.randn(2, 1, 17, generator=g)]


def mel_inputs():
    """Speech-like waveforms [B, L] in [-1, 1]: harmonic stacks + noise; one L a multiple of the hop, one not, one with a
    silent stretch (exercises the 1e-5 log floor)."""
    g = torch.Generator().manual_seed(88)
    out = []
    for B, L in ((2, 24000), (1, 5003), (3, 1300)):
        t = torch.arange(L) / 24000.0
        f0 = 90.0 + 120.0 * torch.rand(B, 1, generator=g)
        w = sum(torch.sin(2 * torch.pi * f0 * h * t[None]) / h for h in range(1, 9)) * 0.2
        w = w + 0.02 * _randn(B, L, generator=g)
        w[:, L // 3: L // 3 + 700] = 0.0
        out.append(w.clamp(-1, 1).contiguous())
    return out


# reference-mel style path: ragged batch of normalised reference mels
STYLE_CASE = dict(weight_seed=1234, input_seed=11, lengths=[200, 133, 64], rel_pos_type="legacy")
ACOUSTIC_REFMEL_CASE = dict(api="infer_batch", rel_pos_type="legacy", lengths=[9, 6], weight_seed=1234,
                            frames_per_phoneme=3.0, input_seed=7, noise_seed=107, noise_scale=1.0, K_step=100,
                            ref_lengths=[150, 97])


def style_inputs(case=None):
    case = case or STYLE_CASE
    lens = case["lengths"] if "ref_lengths" not in case else case["ref_lengths"]
    g = torch.Generator().manual_seed(case["input_seed"] + 500)
    mel = torch.randn(len(lens), 80, max(lens), generator=g)
    for b, n in enumerate(lens):
        mel[b, :, n:] = 0
    return mel, torch.tensor(lens, dtype=torch.int64)


# use_max=False: per-dimension Categorical draw of the style MDN component (mdn.py:226-257), realised from uniforms
ACOUSTIC_SAMPLED_CASE = dict(api="infer_batch", rel_pos_type="legacy", lengths=[8, 5], weight_seed=1234,
                             frames_per_phoneme=3.0, input_seed=9, noise_seed=109, noise_scale=0.7, K_step=100)


def component_uniforms(case, B, C=256):
    g = torch.Generator().manual_seed(case["noise_seed"] + 5000)
    return torch.rand(B, C, generator=g)


# BERT: a small config whose weights fit in the golden file; bert-base dimensions are exercised on the
# GPU against the oracle with seeded weights
BERT_SMALL = dict(vocab_size=120, hidden_size=128, num_hidden_layers=2, num_attention_heads=4, intermediate_size=256,
                  max_position_embeddings=64, type_vocab_size=2, layer_norm_eps=1e-12)


def bert_inputs(vocab=120, B=3, T=11, seed=13):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(1, vocab, (B, T), generator=g)
    lens = [T, T - 4, 3]
    mask = torch.zeros(B, T, dtype=torch.int64)
    for b, n in enumerate(lens[:B]):
        mask[b, :n] = 1
        ids[b, n:] = 0
    return ids, mask

"""CPU arm of bench.py: the UNMODIFIED reference modules (line/promptttspp) timed on the host cores.

The reference is installed once into the git-ignored `baseline/_ref/` by `baseline/install_reference.sh`
(`pip install --no-index --no-deps --target baseline/_ref <copy of /root/reference>`); that directory travels to the GPU
box with the snapshot, `/root/reference` does not.  Nothing from promptttspp_b200's kernels or engine is on this path:
the modules are the reference's own `PromptTTSMDNDurCFG.infer / infer_batch` (models/prompttts_mdn_v2_final/model.py:
198-325) and `BigVGAN.forward` (vocoders/bigvgan.py:120-131), built with the kwargs of the shipped yaml
(promptttspp_b200/utils/synthetic.py only supplies those kwargs and the seeded synthetic checkpoint, loaded strict).
The one substitution is `BertWrapper` -> a fixed sentence-embedding provider: BASELINE.json's configs use a fixed
style-prompt embedding and no BERT weights exist offline.

When `baseline/_ref` is absent (never installed) the functions fall back to the oracle port and say `kind: "port"`.
"""
import os
import statistics
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
REF_DIR = ROOT / "baseline" / "_ref"
UNIT = "frames/s"


class _FixedBert(torch.nn.Module):
    """Stands in for promptttspp.modules.prompt_encoder.BertWrapper: returns the given [N, 768] embeddings."""

    table = None

    def __init__(self, *a, **k):
        super().__init__()

    def forward(self, prompts, device):
        return _FixedBert.table[: len(prompts)].to(device)


def available():
    return (REF_DIR / "promptttspp" / "__init__.py").exists() or (REF_DIR / "promptttspp").is_dir()


def _namespace():
    if str(REF_DIR) not in sys.path:
        sys.path.insert(0, str(REF_DIR))
    import promptttspp.modules.prompt_encoder as pe

    assert Path(pe.__file__).resolve().is_relative_to(REF_DIR.resolve()), pe.__file__
    pe.BertWrapper = _FixedBert
    from promptttspp.layers.embedding import PhonemeEmbedding
    from promptttspp.models.prompttts_mdn_v2_final.model import PromptTTSMDNDurCFG
    from promptttspp.modules.denoiser import DiffNet
    from promptttspp.modules.diffusion import GaussianDiffusion
    from promptttspp.modules.esp import ConformerEncoder
    from promptttspp.modules.frame_prior import FramePriorNetwork
    from promptttspp.modules.mdn import MDNLayer
    from promptttspp.modules.prompt_encoder import PromptEncoder
    from promptttspp.modules.style_encoder import StyleEncoder
    from promptttspp.modules.variance_adaptor import MDNPredictor, Predictor, VarianceAdaptor

    return dict(PhonemeEmbedding=PhonemeEmbedding, PromptTTSMDNDurCFG=PromptTTSMDNDurCFG, DiffNet=DiffNet,
                GaussianDiffusion=GaussianDiffusion, ConformerEncoder=ConformerEncoder,
                FramePriorNetwork=FramePriorNetwork, MDNLayer=MDNLayer, PromptEncoder=PromptEncoder,
                StyleEncoder=StyleEncoder, MDNPredictor=MDNPredictor, Predictor=Predictor,
                VarianceAdaptor=VarianceAdaptor)


def reference_acoustic(seed=1234):
    """The reference's PromptTTSMDNDurCFG (legacy rel-pos demo config) with the synthetic checkpoint, on the CPU."""
    from promptttspp_b200.utils.synthetic import build_acoustic, synthetic_state_dict

    model = build_acoustic(ns=_namespace()).eval()
    model.load_state_dict(synthetic_state_dict(model, seed=seed), strict=True)
    return model


def reference_vocoder(seed=4321):
    from promptttspp_b200.utils.synthetic import build_vocoder, synthetic_state_dict

    _namespace()
    import promptttspp.vocoders as ref_voc

    voc = build_vocoder(ns=ref_voc).eval()
    voc.load_state_dict(synthetic_state_dict(voc, seed=seed), strict=True)
    return voc


def _inputs(seed, B, lo, hi):
    g = torch.Generator().manual_seed(seed)
    lengths = torch.randint(lo, hi, (B,), generator=g)
    lengths[0] = hi - 1
    Tx = int(lengths.max())
    phoneme = torch.zeros(B, Tx, dtype=torch.int64)
    for b in range(B):
        phoneme[b, : int(lengths[b])] = torch.randint(3, 90, (int(lengths[b]),), generator=g)
    return phoneme, lengths, torch.randn(B, 768, generator=g)


def _time(fn, steps, warmup):
    times, out = [], None
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        out = fn()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return statistics.median(times), times, out


@torch.no_grad()
def acoustic_sample(steps=3, warmup=1, B=4, lo=48, hi=65):
    """Bounded sample of cfg2 (cost is linear in padded frames x 100 diffusion steps): B=4, Tx in [48, 64] ->
    >= 2 k padded frames per step.  Returns (cpu_baseline dict, seconds per step)."""
    torch.set_num_threads(os.cpu_count() or 1)
    phoneme, lengths, cls_emb = _inputs(2, B, lo, hi)
    if available():
        model = reference_acoustic()
        _FixedBert.table = cls_emb

        def run():
            torch.manual_seed(7)
            return model.infer_batch(phoneme, lengths, style_prompt=["p"] * B, use_max=True, noise_scale=0.5,
                                     return_f0=True)

        t, times, out = _time(run, steps, warmup)
        mel, flen = out[0], out[3]
        kind, what = "reference", "baseline/_ref promptttspp PromptTTSMDNDurCFG.infer_batch (unmodified reference modules)"
    else:
        sys.path.insert(0, str(ROOT))
        from oracle import oracle
        from promptttspp_b200.modules.prompt_encoder import FixedPromptEmbedding
        from promptttspp_b200.utils.synthetic import build_acoustic, synthetic_state_dict

        sd = synthetic_state_dict(build_acoustic(bert=FixedPromptEmbedding(torch.zeros(1, 768))), seed=1234)
        g = torch.Generator().manual_seed(7)
        z_style = torch.randn(B, 1, 256, generator=g)

        def run():
            return oracle.acoustic_infer_batch(sd, dict(oracle.ACOUSTIC_CFG), phoneme, lengths, cls_emb, z_style,
                                               noise_fn=lambda s: torch.randn(s))

        t, times, out = _time(run, steps, warmup)
        mel, flen = out[0], out[3]
        kind, what = "port", "oracle/oracle.py acoustic_infer_batch (baseline/_ref not installed)"
    frames, padded = float(flen.sum()), mel.shape[0] * mel.shape[-1]
    return dict(value=frames / t, unit=UNIT, cores=torch.get_num_threads(), kind=kind,
                sample=f"{what}, B={B} Tx<={hi - 1} -> {int(frames)} valid / {padded} padded frames, 100 diffusion "
                       f"steps, {warmup} warm-up + median of {len(times)} ({t:.2f} s each)"), t


@torch.no_grad()
def cfg1_sample(steps=3, warmup=1, n_phonemes=50):
    """BASELINE.json configs[0] in full: ONE utterance of 50 phonemes through the reference's `infer` on the CPU."""
    torch.set_num_threads(os.cpu_count() or 1)
    phoneme, lengths, cls_emb = _inputs(11, 1, n_phonemes, n_phonemes + 1)
    if not available():
        return None
    model = reference_acoustic()
    _FixedBert.table = cls_emb

    def run():
        torch.manual_seed(7)
        return model.infer(phoneme, style_prompt="p", use_max=True, noise_scale=0.5)

    t, times, mel = _time(run, steps, warmup)
    frames = mel.shape[-1]
    return dict(latency_ms=t * 1e3, frames=frames, frames_per_sec=frames / t, cores=torch.get_num_threads(),
                kind="reference", sample=f"baseline/_ref PromptTTSMDNDurCFG.infer, 1 x {n_phonemes} phonemes -> {frames} "
                                         f"frames, {warmup} warm-up + median of {len(times)}")


@torch.no_grad()
def bigvgan_sample(steps=3, warmup=1, B=2, T=256):
    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(3)
    mel = (torch.randn(B, 80, T, generator=g) * 2.0 - 5.0).clamp(-11.5, 2.0)
    if available():
        voc = reference_vocoder()
        t, times, _ = _time(lambda: voc(mel), steps, warmup)
        kind, what = "reference", "baseline/_ref promptttspp BigVGAN.forward"
    else:
        sys.path.insert(0, str(ROOT))
        from oracle import oracle
        from promptttspp_b200.utils.synthetic import build_vocoder, synthetic_state_dict

        sd = synthetic_state_dict(build_vocoder(), seed=4321)
        t, times, _ = _time(lambda: oracle.bigvgan_forward(sd, oracle.VOCODER_CFG, mel), steps, warmup)
        kind, what = "port", "oracle/oracle.py bigvgan_forward (baseline/_ref not installed)"
    audio_s = B * T / 100.0
    return dict(rtf=t / audio_s, frames_per_sec=B * T / t, cores=torch.get_num_threads(), kind=kind,
                sample=f"{what}, B={B} x {T} frames ({audio_s:.2f} s audio), {warmup} warm-up + median of "
                       f"{len(times)} ({t:.2f} s each)")

"""Generate the golden vectors under tests/golden/ by running the REFERENCE's own modules.

Run where /root/reference exists (the build container), never on the GPU box:

    python tests/golden/make_golden.py

The reference has no tests, fixtures or pretrained weights, so parity is pinned
by executing its nn.Modules on CPU with
  * the synthetic, seeded checkpoints of promptttspp_b200/utils/synthetic.py (loaded strict into the
    reference modules -- which also proves the state_dict key/shape contract),
  * BERT replaced by a fixed sentence-embedding provider (no network / weights offline),
  * Gaussian noise injected at the three places the reference draws it (model.py:191,
    diffusion.py:332, diffusion.py:30-38/218), regenerated from a seed by `golden_noise`.
Only inputs that cannot be regenerated from seeds and the reference OUTPUTS are stored (npz).
"""
import sys
import warnings
from contextlib import contextmanager
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
REF = Path("/root/reference")

from golden_cases import (ACOUSTIC_CASES, ACOUSTIC_LARGE_CASES, AA_CASES, TEXT_CASES, VOCODER_CASES,  # noqa: E402
                          VOCODER_LARGE_CASES, acoustic_inputs, golden_noise, vocoder_inputs)

from promptttspp_b200.utils.synthetic import build_acoustic, build_vocoder, synthetic_state_dict  # noqa: E402

warnings.filterwarnings("ignore")
torch.set_grad_enabled(False)
torch.set_num_threads(8)
OUT = Path(__file__).resolve().parent


class _FixedBert(torch.nn.Module):
    """Stands in for promptttspp.modules.prompt_encoder.BertWrapper: returns the given embeddings."""

    table = None

    def __init__(self, *a, **k):
        super().__init__()

    def forward(self, prompts, device):
        return _FixedBert.table[: len(prompts)].to(device)


def reference_namespace():
    sys.path.insert(0, str(REF))
    import promptttspp.modules.prompt_encoder as pe

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


@contextmanager
def injected_noise(z_style, x_T_fn, z_fn):
    """Patch the reference's three noise sources.  x_T / z depend on Ty, known only mid-call."""
    import promptttspp.modules.diffusion as dmod

    state = {"n": 0}
    orig_randn_like, orig_randn, orig_noise_like = torch.randn_like, torch.randn, dmod.noise_like

    def randn_like(t, *a, **k):
        assert tuple(t.shape) == tuple(z_style.shape), (t.shape, z_style.shape)
        return z_style.clone()

    def randn(*shape, **k):
        shape = tuple(shape[0]) if len(shape) == 1 and not isinstance(shape[0], int) else tuple(shape)
        return x_T_fn(shape)

    def noise_like(shape, noise_fn, device, repeat=False):
        out = z_fn(tuple(shape), state["n"])
        state["n"] += 1
        return out

    torch.randn_like, torch.randn, dmod.noise_like = randn_like, randn, noise_like
    try:
        yield
    finally:
        torch.randn_like, torch.randn, dmod.noise_like = orig_randn_like, orig_randn, orig_noise_like


class _StopAfterDurations(Exception):
    pass


def make_text(ns):
    """cfg2's text side through the reference up to the duration predictor (the rest of infer_batch is skipped by
    raising from the hooked predictor): log-durations and the integer durations of >= 3 k phonemes."""
    for name, case in TEXT_CASES.items():
        print("text", name)
        model = build_acoustic(rel_pos_type=case["rel_pos_type"], ns=ns, K_step=case["K_step"]).eval()
        sd = synthetic_state_dict(build_acoustic(rel_pos_type=case["rel_pos_type"], bert=_FixedBert(),
                                                 K_step=case["K_step"]),
                                  seed=case["weight_seed"], frames_per_phoneme=case["frames_per_phoneme"])
        model.load_state_dict(sd, strict=True)
        phoneme, lengths, cls_emb = acoustic_inputs(case)
        _FixedBert.table = cls_emb
        B = phoneme.shape[0]
        inter = {}
        va = model.variance_adaptor
        orig_dur = va.duration_predictor.infer

        def hooked(x, m):
            inter["log_d"] = orig_dur(x, m)
            raise _StopAfterDurations()

        va.duration_predictor.infer = hooked
        z_style = golden_noise(case, B, None).z_style
        try:
            with injected_noise(z_style, None, None):
                model.infer_batch(phoneme, lengths, style_prompt=["p"] * B, use_max=True,
                                  noise_scale=case["noise_scale"], return_f0=True)
        except _StopAfterDurations:
            pass
        log_d = inter["log_d"]
        dur = log_d.exp().round().clamp_min(1).long().squeeze(1)
        dur = dur * (torch.arange(phoneme.shape[1])[None] < lengths[:, None]).long()
        print("   phonemes", int(lengths.sum()), "frames", dur.sum(1).tolist())
        np.savez_compressed(OUT / f"text_{name}.npz", log_d=log_d.numpy(), duration=dur.numpy())


def make_acoustic(ns, cases=None, store_cond=True):
    for name, case in (cases or ACOUSTIC_CASES).items():
        print("acoustic", name, case)
        model = build_acoustic(rel_pos_type=case["rel_pos_type"], ns=ns, K_step=case["K_step"]).eval()
        sd = synthetic_state_dict(build_acoustic(rel_pos_type=case["rel_pos_type"], bert=_FixedBert(),
                                                 K_step=case["K_step"]),
                                  seed=case["weight_seed"], frames_per_phoneme=case["frames_per_phoneme"])
        missing = model.load_state_dict(sd, strict=True)  # key/shape contract with the reference
        assert not missing.missing_keys and not missing.unexpected_keys
        phoneme, lengths, cls_emb = acoustic_inputs(case)
        _FixedBert.table = cls_emb
        B = phoneme.shape[0]
        noise = {}

        def x_T_fn(shape):
            noise["full"] = golden_noise(case, B, shape[-1])
            return noise["full"].x_T.clone()

        def z_fn(shape, n):
            return noise["full"].z[n].clone()

        z_style = golden_noise(case, B, None).z_style
        inter = {}
        va = model.variance_adaptor
        orig_dur = va.duration_predictor.infer
        va.duration_predictor.infer = lambda x, m: inter.setdefault("log_d", orig_dur(x, m))
        orig_dec = model.decoder.inference
        model.decoder.inference = lambda c, l, g=None: orig_dec(inter.setdefault("cond", c), l, g=g)
        with injected_noise(z_style, x_T_fn, z_fn):
            if case["api"] == "infer":
                mel, log_cf0, vuv = model.infer(phoneme, style_prompt=["p"] * B, use_max=True,
                                                noise_scale=case["noise_scale"], return_f0=True)
                frame_lengths = torch.tensor([mel.shape[-1]], dtype=torch.float32)
            else:
                mel, log_cf0, vuv, frame_lengths = model.infer_batch(
                    phoneme, lengths, style_prompt=["p"] * B, use_max=True, noise_scale=case["noise_scale"],
                    return_f0=True)
        log_d = inter["log_d"]
        dur = log_d.exp().round().clamp_min(1).long().squeeze(1)
        if case["api"] != "infer":
            dur = dur * (torch.arange(phoneme.shape[1])[None] < lengths[:, None]).long()
        print("   Ty", mel.shape[-1], "frame_lengths", frame_lengths.tolist(), "mel absmax", float(mel.abs().max()))
        extra = dict(cond=inter["cond"].numpy()) if store_cond else {}
        np.savez_compressed(
            OUT / f"acoustic_{name}.npz",
            mel=mel.numpy(), log_cf0=log_cf0.numpy(), vuv=vuv.numpy(), frame_lengths=frame_lengths.numpy(),
            log_d=log_d.numpy(), duration=dur.numpy(), **extra,
        )


def make_vocoder(cases=None):
    sys.path.insert(0, str(REF))
    import promptttspp.vocoders as ref_voc

    for name, case in (cases or VOCODER_CASES).items():
        print("vocoder", name, case)
        voc = build_vocoder(ns=ref_voc).eval()
        sd = synthetic_state_dict(build_vocoder(), seed=case["weight_seed"])
        voc.load_state_dict(sd, strict=True)
        mel = vocoder_inputs(case)
        wav = voc(mel)
        print("   wav", tuple(wav.shape), "rms", float(wav.pow(2).mean().sqrt()))
        out = {"wav": wav.numpy()}
        if case.get("remove_weight_norm"):
            from promptttspp.utils.model import remove_weight_norm_

            voc.apply(remove_weight_norm_)
            out["wav_nowm"] = voc(mel).numpy()
        np.savez_compressed(OUT / f"vocoder_{name}.npz", **out)


def make_vocoder_f0():
    """Reference F0AwareBigVGAN with its three random draws (nsf.py:64 torch.rand, :143 and :205 torch.randn_like)
    replaced by the seeded tensors of tests/golden_cases.py."""
    sys.path.insert(0, str(REF))
    import promptttspp.vocoders as ref_voc
    from golden_cases import F0_KWARGS, VOCODER_F0_CASES, vocoder_f0_inputs
    from promptttspp_b200.utils.synthetic import build_vocoder_f0

    for name, case in VOCODER_F0_CASES.items():
        print("vocoder_f0", name, case)
        voc = build_vocoder_f0(ns=ref_voc, **F0_KWARGS).eval()
        sd = synthetic_state_dict(build_vocoder_f0(**F0_KWARGS), seed=case["weight_seed"])
        voc.load_state_dict(sd, strict=True)
        mel, f0, rand_ini, noise = vocoder_f0_inputs(case)
        real_rand, real_randn_like = torch.rand, torch.randn_like
        calls = []

        def fake_rand(*shape, **k):
            calls.append(("rand", tuple(shape)))
            assert tuple(shape) == tuple(rand_ini.shape), shape
            return rand_ini.clone()

        def fake_randn_like(t, *a, **k):
            calls.append(("randn_like", tuple(t.shape)))
            if tuple(t.shape) == tuple(noise.shape):
                return noise.clone()
            return torch.zeros_like(t)  # nsf.py:205: the noise-branch source, unused by the vocoder

        torch.rand, torch.randn_like = fake_rand, fake_randn_like
        try:
            with torch.no_grad():
                wav = voc(mel, f0)
                f0_up = voc.f0_up(f0).transpose(-1, -2)
                har, _, _ = voc.m_source(f0_up)
        finally:
            torch.rand, torch.randn_like = real_rand, real_randn_like
        print("   calls", calls[:3], "wav", tuple(wav.shape), "rms", float(wav.pow(2).mean().sqrt()))
        np.savez_compressed(OUT / f"vocoder_f0_{name}.npz", wav=wav.numpy(), har=har.numpy())


def make_lowpass():
    """promptttspp.utils.model.lowpass_filter (the reference's own function) on seeded log-f0-like contours."""
    sys.path.insert(0, str(REF))
    from promptttspp.utils.model import lowpass_filter as ref_lowpass
    from golden_cases import lowpass_inputs

    out = {}
    for i, x in enumerate(lowpass_inputs()):
        out[f"y{i}"] = ref_lowpass(x, 100, cutoff=20).numpy()
    np.savez_compressed(OUT / "lowpass.npz", **out)
    print("lowpass", {k: v.shape for k, v in out.items()})


def make_mel():
    """promptttspp.transforms.MelSpectrogramTransform with the kwargs of conf/transforms/mel.yaml on seeded waveforms."""
    sys.path.insert(0, str(REF))
    import yaml
    from promptttspp.transforms import MelSpectrogramTransform
    from golden_cases import mel_inputs

    kw = yaml.safe_load(open(REF / "egs/proposed/bin/conf/transforms/mel.yaml"))
    kw.pop("_target_")
    to_mel = MelSpectrogramTransform(**kw).eval()
    out = {}
    for i, w in enumerate(mel_inputs()):
        out[f"mel{i}"] = to_mel(w).numpy()
        out[f"spec{i}"] = to_mel.to_spec(w).numpy()
    np.savez_compressed(OUT / "mel_transform.npz", **out)
    print("mel", {k: v.shape for k, v in out.items()})


def make_style(ns):
    """Reference StyleEncoder on a ragged batch, and infer_batch(reference_mel=...) end to end."""
    from golden_cases import ACOUSTIC_REFMEL_CASE, STYLE_CASE, style_inputs

    case = STYLE_CASE
    model = build_acoustic(rel_pos_type=case["rel_pos_type"], ns=ns).eval()
    sd = synthetic_state_dict(build_acoustic(rel_pos_type=case["rel_pos_type"], bert=_FixedBert()), seed=case["weight_seed"])
    model.load_state_dict(sd, strict=True)
    mel, lens = style_inputs()
    with torch.no_grad():
        style = model.reference_encoder(mel, lens)
        style_full = model.reference_encoder(mel[:1], None)
    print("style", tuple(style.shape), float(style.abs().max()))
    out = {"style": style.numpy(), "style_nolen": style_full.numpy()}

    case = ACOUSTIC_REFMEL_CASE
    model = build_acoustic(rel_pos_type=case["rel_pos_type"], ns=ns, K_step=case["K_step"]).eval()
    sd = synthetic_state_dict(build_acoustic(rel_pos_type=case["rel_pos_type"], bert=_FixedBert(), K_step=case["K_step"]),
                              seed=case["weight_seed"], frames_per_phoneme=case["frames_per_phoneme"])
    model.load_state_dict(sd, strict=True)
    phoneme, lengths, _ = acoustic_inputs(case)
    ref_mel, ref_lens = style_inputs(case)
    B = phoneme.shape[0]
    noise = {}

    def x_T_fn(shape):
        noise["full"] = golden_noise(case, B, shape[-1])
        return noise["full"].x_T.clone()

    def z_fn(shape, n):
        return noise["full"].z[n].clone()

    with injected_noise(golden_noise(case, B, None).z_style, x_T_fn, z_fn):
        mel_out, log_cf0, vuv, frame_lengths = model.infer_batch(phoneme, lengths, reference_mel=ref_mel,
                                                                 ref_lengths=ref_lens, use_max=True, return_f0=True)
    print("   refmel acoustic Ty", mel_out.shape[-1], frame_lengths.tolist())
    out.update(mel=mel_out.numpy(), log_cf0=log_cf0.numpy(), vuv=vuv.numpy(), frame_lengths=frame_lengths.numpy())
    np.savez_compressed(OUT / "style_refmel.npz", **out)


def make_sampled(ns):
    """infer_batch(use_max=False): the reference's Categorical(probs=pi).sample() replaced by the inverse CDF of seeded
    uniforms (same distribution; the draw becomes an input like the Gaussian ones)."""
    from golden_cases import ACOUSTIC_SAMPLED_CASE, component_uniforms

    case = ACOUSTIC_SAMPLED_CASE
    model = build_acoustic(rel_pos_type=case["rel_pos_type"], ns=ns, K_step=case["K_step"]).eval()
    sd = synthetic_state_dict(build_acoustic(rel_pos_type=case["rel_pos_type"], bert=_FixedBert(), K_step=case["K_step"]),
                              seed=case["weight_seed"], frames_per_phoneme=case["frames_per_phoneme"])
    model.load_state_dict(sd, strict=True)
    phoneme, lengths, cls_emb = acoustic_inputs(case)
    _FixedBert.table = cls_emb
    B = phoneme.shape[0]
    u = component_uniforms(case, B)
    noise = {}

    def x_T_fn(shape):
        noise["full"] = golden_noise(case, B, shape[-1])
        return noise["full"].x_T.clone()

    def z_fn(shape, n):
        return noise["full"].z[n].clone()

    real_sample = torch.distributions.Categorical.sample

    def fake_sample(self, sample_shape=torch.Size()):
        probs = self.probs                                   # (B, C, G), normalised by Categorical
        assert tuple(probs.shape[:2]) == tuple(u.shape), probs.shape
        cdf = torch.cumsum(probs, dim=-1)
        return (u.unsqueeze(-1) >= cdf).sum(-1).clamp(max=probs.shape[-1] - 1)

    torch.distributions.Categorical.sample = fake_sample
    try:
        with injected_noise(golden_noise(case, B, None).z_style, x_T_fn, z_fn):
            mel, log_cf0, vuv, frame_lengths = model.infer_batch(phoneme, lengths, style_prompt=["p"] * B, use_max=False,
                                                                 noise_scale=case["noise_scale"], return_f0=True)
    finally:
        torch.distributions.Categorical.sample = real_sample
    print("   sampled acoustic Ty", mel.shape[-1], frame_lengths.tolist())
    np.savez_compressed(OUT / "acoustic_sampled_b2.npz", mel=mel.numpy(), log_cf0=log_cf0.numpy(), vuv=vuv.numpy(),
                        frame_lengths=frame_lengths.numpy())


def make_bert():
    """HF transformers BertModel (the reference's dependency, prompt_encoder.py:25) with a small seeded config: weights
    and last_hidden_state stored together."""
    from golden_cases import BERT_SMALL, bert_inputs
    from transformers import BertConfig, BertModel

    torch.manual_seed(99)
    hf = BertModel(BertConfig(**BERT_SMALL, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)).eval()
    with torch.no_grad():
        for n, p in hf.named_parameters():  # de-degenerate LayerNorm / biases
            if "LayerNorm.weight" in n:
                p.copy_(1.0 + 0.1 * torch.randn_like(p))
            elif n.endswith("bias"):
                p.copy_(0.05 * torch.randn_like(p))
            else:
                p.mul_(4.0)
    ids, mask = bert_inputs(BERT_SMALL["vocab_size"])
    with torch.no_grad():
        out = hf(input_ids=ids, attention_mask=mask).last_hidden_state
    blob = {"w::" + k: v.numpy() for k, v in hf.state_dict().items()}
    blob["last_hidden_state"] = out.numpy()
    np.savez_compressed(OUT / "bert_small.npz", **blob)
    print("bert", tuple(out.shape), float(out.abs().max()))


def make_ops():
    """Op-level vectors from the reference layers: pin the closed forms used by the kernels."""
    sys.path.insert(0, str(REF))
    from promptttspp.layers.activations import AntiAliasActivation
    from promptttspp.modules.esp.transformer.attention import (LegacyRelPositionMultiHeadedAttention,
                                                               RelPositionMultiHeadedAttention)
    from promptttspp.modules.esp.transformer.embedding import LegacyRelPositionalEncoding, RelPositionalEncoding
    from promptttspp.utils.model import generate_path, sequence_mask

    out = {}
    g = torch.Generator().manual_seed(77)
    for name, (B, C, L) in AA_CASES.items():
        act = AntiAliasActivation(C).eval()
        act.act.alpha.data = torch.rand(1, C, 1, generator=g) - 0.5
        x = torch.randn(B, C, L, generator=g) * 2
        out[f"aa_{name}_x"] = x.numpy()
        out[f"aa_{name}_alpha"] = act.act.alpha.data.numpy()
        out[f"aa_{name}_y"] = act(x).numpy()
    out["aa_up_filter"] = AntiAliasActivation(1).up.filter.numpy()
    out["aa_down_filter"] = AntiAliasActivation(1).down.lowpass.filter.numpy()
    # rel_shift, both flavours, and the positional tables
    T = 7
    x = torch.randn(2, 2, T, T, generator=g)
    out["rel_shift_legacy_in"] = x.numpy()
    out["rel_shift_legacy_out"] = LegacyRelPositionMultiHeadedAttention(2, 8, 0.0).rel_shift(x).numpy()
    x = torch.randn(2, 2, T, 2 * T - 1, generator=g)
    out["rel_shift_new_in"] = x.numpy()
    out["rel_shift_new_out"] = RelPositionMultiHeadedAttention(2, 8, 0.0).rel_shift(x).numpy()
    d = torch.zeros(1, 9, 16)
    out["pos_legacy"] = LegacyRelPositionalEncoding(16, 0.0).eval()(d)[1][0].numpy()
    out["pos_new"] = RelPositionalEncoding(16, 0.0).eval()(d)[1][0].numpy()
    # length regulator
    dur = torch.tensor([[2, 1, 3, 0, 0], [1, 1, 1, 4, 2]])
    xlen = torch.tensor([3, 5])
    flen = dur.sum(1)
    pm = sequence_mask(xlen, 5).unsqueeze(1).float()
    fm = sequence_mask(flen).unsqueeze(1).float()
    path = generate_path(dur, (pm.unsqueeze(-1) * fm.unsqueeze(2)).squeeze(1))
    out["lr_dur"] = dur.numpy()
    out["lr_path"] = path.numpy()
    np.savez_compressed(OUT / "ops.npz", **out)


if __name__ == "__main__":
    which = sys.argv[1:] or ["ops", "vocoder", "vocoder_f0", "lowpass", "mel", "style", "sampled", "bert", "acoustic"]
    if "ops" in which:
        make_ops()
    if "vocoder" in which:
        make_vocoder()
    if "vocoder_f0" in which:
        make_vocoder_f0()
    if "lowpass" in which:
        make_lowpass()
    if "mel" in which:
        make_mel()
    if "style" in which:
        make_style(reference_namespace())
    if "sampled" in which:
        make_sampled(reference_namespace())
    if "bert" in which:
        make_bert()
    if "acoustic" in which:
        make_acoustic(reference_namespace())
    # benchmark-scale cases (minutes of CPU time): only on request
    if "vocoder_large" in which:
        make_vocoder(VOCODER_LARGE_CASES)
    if "text" in which:
        make_text(reference_namespace())
    if "acoustic_large" in which:
        make_acoustic(reference_namespace(), ACOUSTIC_LARGE_CASES, store_cond=False)
    print("done")

"""Front-end pieces either side of the hot path (SURVEY.md 8 rows b, f3, f4): yaml `_target_` instantiation, the phoneme
symbol table, the mel-spectrogram transform (oracle vs the reference's golden on the CPU, CUDA kernel vs golden on the
GPU) and generate_style_emb."""
from pathlib import Path

import numpy as np
import pytest
import torch

from golden_cases import mel_inputs
from oracle import oracle
from promptttspp_b200.utils import config as cfgmod

torch.set_grad_enabled(False)
REF_CONF = Path("/root/reference/egs/proposed/bin/conf")


# ---- yaml resolver ------------------------------------------------------------------------------------------------------
def test_resolve_relative_and_absolute_interpolations():
    cfg = {"a": {"c": 256, "n": {"x": "${..c}", "y": "${...top}", "z": "${.x}", "s": "v${a.c}_${.x}"}}, "top": [1, 2],
           "lst": [{"k": "${...top}"}]}
    r = cfgmod.resolve(cfg)
    assert r["a"]["n"] == {"x": 256, "y": [1, 2], "z": 256, "s": "v256_256"}
    assert r["lst"][0]["k"] == [1, 2]
    with pytest.raises(KeyError):
        cfgmod.resolve({"a": {"b": "${....x}"}})
    with pytest.raises(ValueError):
        cfgmod.resolve({"a": "${b}", "b": "${a}"})


def test_instantiate_builds_nested_targets_depth_first():
    cfg = {"_target_": "torch.nn.Sequential", "_args_unused": None}
    cfg = {"_target_": "collections.OrderedDict", "conv": {"_target_": "torch.nn.Conv1d", "in_channels": 1,
                                                           "out_channels": "${..width}", "kernel_size": 1},
           "width": 8}
    out = cfgmod.instantiate(cfg)
    assert isinstance(out["conv"], torch.nn.Conv1d) and out["conv"].out_channels == 8 and out["width"] == 8
    assert cfgmod.locate("promptttspp.vocoders.BigVGAN").__module__.startswith("promptttspp_b200.")
    part = cfgmod.instantiate({"_target_": "torch.zeros", "_partial_": True, "size": [2]})
    assert part().shape == (2,)


@pytest.mark.skipif(not REF_CONF.exists(), reason="the reference's yaml files exist only in the build container")
def test_demo_yaml_instantiates_the_shims_with_the_checkpoint_contract():
    """app.py:28-39: instantiate(cfg.model) / instantiate(cfg.vocoder) / instantiate(cfg.transforms) on the REAL
    demo.yaml with the `promptttspp.` prefix re-pointed; the result must carry exactly the state_dict contract of the
    hand-typed kwargs in utils/synthetic.py (which the goldens pin to the reference's modules)."""
    from promptttspp_b200.utils.synthetic import build_acoustic, build_vocoder_f0

    cfg = cfgmod.load_config(REF_CONF / "demo.yaml")
    assert set(cfg) >= {"model", "vocoder", "transforms", "model_ckpt_path", "mel_stats_file"}
    model = cfgmod.instantiate(cfg["model"])
    want = build_acoustic(rel_pos_type="legacy")
    got_sd, want_sd = model.state_dict(), want.state_dict()
    assert list(got_sd) == list(want_sd)
    assert all(got_sd[k].shape == want_sd[k].shape for k in want_sd)
    assert type(model).__module__ == "promptttspp_b200.models.prompttts_mdn_v2_final.model"
    assert model.encoder.rel_pos_type == "legacy" and model.norm_style_emb is True
    voc = cfgmod.instantiate(cfg["vocoder"])
    assert list(voc.state_dict()) == list(build_vocoder_f0(sampling_rate=24000, harmonic_num=8).state_dict())
    to_mel = cfgmod.instantiate(cfg["transforms"])
    assert (to_mel.sample_rate, to_mel.n_fft, to_mel.hop_length, to_mel.n_mels) == (24000, 512, 240, 80)
    # the training-flavour model yaml (new rel-pos) and the plain vocoder resolve too
    m2 = cfgmod.instantiate(cfgmod.load_yaml(REF_CONF / "model" / "prompttts_mdn_v2_wo_erg_final.yaml"))
    assert list(m2.state_dict()) == list(build_acoustic(rel_pos_type=m2.encoder.rel_pos_type).state_dict())
    v2 = cfgmod.instantiate(cfgmod.load_yaml(REF_CONF / "vocoder" / "bigvgan.yaml"))
    assert type(v2).__name__ == "BigVGAN"


# ---- phoneme symbols ----------------------------------------------------------------------------------------------------
def test_symbol_table_and_sequences():
    from promptttspp_b200.text import eng

    assert eng.num_vocab() == 90 and eng.symbols[:3] == ["_", "^", "$"] and eng.symbols[-3:] == ["spn", "sil", "sp"]
    assert eng.symbols[3:7] == ["AA", "AA0", "AA1", "AA2"] and eng.symbol_to_id("ZH") == 86
    seq = eng.text_to_sequence("HH AH0 L OW1 sp")
    assert seq[0] == 1 and seq[-1] == 2 and len(seq) == 7
    assert eng.sequence_to_text(seq, remove_special_token=True) == ["HH", "AH0", "L", "OW1", "sp"]
    assert eng.text_to_sequence("K", add_special_token=False) == [eng.symbol_to_id("K")]
    with pytest.raises(KeyError):
        eng.text_to_sequence("QQ")
    ids, lens = eng.batch_text_to_sequence(["HH AH0", "K AE1 T sp"], pin=False)
    assert ids.shape == (2, 6) and lens.tolist() == [4, 6] and ids[0, 4:].tolist() == [0, 0]
    ref = Path("/root/reference/promptttspp/text/eng.py")
    if ref.exists():
        import importlib.util

        spec = importlib.util.spec_from_file_location("_ref_eng", ref)
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        assert m.symbols == eng.symbols


# ---- mel transform ------------------------------------------------------------------------------------------------------
def test_mel_oracle_matches_reference_golden(golden_dir):
    gold = np.load(golden_dir / "mel_transform.npz")
    for i, w in enumerate(mel_inputs()):
        spec, mel = oracle.mel_spectrogram(w)
        rs, rm = torch.from_numpy(gold[f"spec{i}"]), torch.from_numpy(gold[f"mel{i}"])
        assert spec.shape == rs.shape and mel.shape == rm.shape
        assert float((spec - rs).abs().max()) < 2e-4 * float(rs.abs().max())
        assert float((mel - rm).abs().max()) < 2e-3, float((mel - rm).abs().max())
        assert float((rm == np.log(1e-5).astype(np.float32)).float().mean()) > 0.0 or i != 0  # the floor is exercised


def test_mel_filterbank_matches_oracle():
    from promptttspp_b200.transforms.mel import melscale_fbanks

    a = melscale_fbanks(257, 63, 12000, 80, 24000, "slaney", "slaney")
    b = oracle.mel_filterbank(257, 63, 12000, 80, 24000)
    assert torch.allclose(a, b, atol=1e-9, rtol=1e-6)


@pytest.mark.gpu
def test_mel_transform_kernel_matches_reference_golden(golden_dir):
    from promptttspp_b200.transforms import MelSpectrogramTransform

    gold = np.load(golden_dir / "mel_transform.npz")
    to_mel = MelSpectrogramTransform(**{k: v for k, v in oracle.MEL_CFG.items()}).cuda().eval()
    for i, w in enumerate(mel_inputs()):
        rs, rm = torch.from_numpy(gold[f"spec{i}"]), torch.from_numpy(gold[f"mel{i}"])
        mel = to_mel(w.cuda()).cpu()
        spec = to_mel.to_spec(w.cuda())
        mel2 = to_mel.spec_to_mel(spec).cpu()
        e_spec = float((spec.cpu() - rs).abs().max()) / float(rs.abs().max())
        e_mel = float((mel - rm).abs().max())
        print(f"mel case {i}: spec rel err {e_spec:.2e}, log-mel max-abs err {e_mel:.2e}")
        assert mel.shape == rm.shape and e_spec < 2e-5 and e_mel < 2e-3
        assert torch.allclose(mel, mel2, atol=1e-5)
    one = to_mel(mel_inputs()[1][0].cuda())           # an un-batched waveform keeps its rank (torchaudio semantics)
    assert one.shape == (80, 21)
    with pytest.raises(RuntimeError):
        to_mel(torch.zeros(1, 1000))                   # host tensor: no CPU fallback
    with pytest.raises(RuntimeError):
        to_mel(torch.zeros(1, 200).cuda())             # too short for the reflect padding


@pytest.mark.gpu
def test_generate_style_emb_matches_infer_style_and_oracle():
    """model.py:327-344: prompt_emb == the style vector infer_batch adds to the encoder output (normalised), ref_emb ==
    the normalised reference-mel style encoder output."""
    from golden_cases import STYLE_CASE, style_inputs
    from promptttspp_b200.models.prompttts_mdn_v2_final.model import InferNoise
    from promptttspp_b200.modules.prompt_encoder import FixedPromptEmbedding
    from promptttspp_b200.utils.synthetic import build_acoustic, synthetic_state_dict

    g = torch.Generator().manual_seed(31)
    cls = torch.randn(3, 768, generator=g)
    model = build_acoustic(bert=FixedPromptEmbedding(cls), K_step=2)
    sd = synthetic_state_dict(model, seed=STYLE_CASE["weight_seed"])
    model.load_state_dict(sd, strict=True)
    model = model.cuda().eval()
    mel, lens = style_inputs()
    ref_mel = mel[:1, :, : int(lens[0])].contiguous()
    z = torch.randn(3, 1, 256, generator=g)
    p_emb, r_emb = model.generate_style_emb(["a", "b", "c"], ref_mel.cuda(), use_max=True, noise_scale=0.7,
                                            noise=InferNoise(z, None, None))
    assert p_emb.shape == (3, 256, 1) and r_emb.shape == (1, 256, 1)
    want = oracle.style_embedding({k: v.cpu() for k, v in sd.items()}, oracle.ACOUSTIC_CFG, cls, z, 0.7)
    want = torch.nn.functional.normalize(want.reshape(3, 256, 1), dim=1)
    assert float((p_emb.cpu() - want).abs().max()) < 2e-6
    want_r = torch.nn.functional.normalize(oracle.style_encoder({k: v.cpu() for k, v in sd.items()}, ref_mel), dim=1)
    assert float((r_emb.cpu() - want_r.reshape(1, 256, 1)).abs().max()) < 5e-6

"""Synthetic, seeded, de-degenerated checkpoints and the shipped model configurations.

There are no pretrained weights for the reference offline, so tests, golden vectors and the
benchmark all run on checkpoints produced here (SURVEY.md section 8d "Synthetic weights").
Default initialisations hide large parts of the path (DiffNet.output_projection is zero,
Snake alpha is 0, LayerNorm is identity, BatchNorm stats are (0, 1), weight-norm g = ||v||), so
every tensor is re-drawn with a spread that keeps activations O(1) through the whole network.
The generator only looks at key names and shapes: the same state_dict loads (strict) into the
reference modules and into the promptttspp_b200 shims.
"""
import math

import torch

ACOUSTIC_YAML = "egs/proposed/bin/conf/model/prompttts_mdn_v2_wo_erg_final_demo.yaml"
VOCODER_YAML = "egs/proposed/bin/conf/vocoder/bigvgan.yaml"

VOCODER_KWARGS = dict(
    in_channel=80,
    upsample_initial_channel=512,
    upsample_rates=[6, 5, 4, 2],
    upsample_kernel_sizes=[12, 10, 8, 4],
    resblock_kernel_sizes=[3, 7, 11],
    resblock_dilations=[[1, 3, 5], [1, 3, 5], [1, 3, 5]],
)


def build_vocoder(ns=None, **overrides):
    """BigVGAN with conf/vocoder/bigvgan.yaml kwargs; `ns` = module namespace providing BigVGAN."""
    if ns is None:
        from .. import vocoders as ns
    kw = dict(VOCODER_KWARGS)
    kw.update(overrides)
    return ns.BigVGAN(**kw)


def build_vocoder_f0(ns=None, sampling_rate=24000, harmonic_num=8, **overrides):
    """F0AwareBigVGAN with conf/vocoder/bigvgan_f0.yaml kwargs; `ns` = module namespace providing the class."""
    if ns is None:
        from .. import vocoders as ns
    kw = dict(VOCODER_KWARGS)
    kw.update(overrides)
    return ns.F0AwareBigVGAN(sampling_rate=sampling_rate, harmonic_num=harmonic_num, **kw)


def build_acoustic(rel_pos_type="legacy", bert=None, K_step=100, ns=None):
    """PromptTTSMDNDurCFG with the kwargs of prompttts_mdn_v2_wo_erg_final[_demo].yaml.

    `ns` maps role -> class; default = this package.  (tests/golden/make_golden.py passes the
    reference's classes to build the very same architecture from /root/reference.)
    """
    if ns is None:
        from ..layers.embedding import PhonemeEmbedding
        from ..models.prompttts_mdn_v2_final.model import PromptTTSMDNDurCFG
        from ..modules.denoiser import DiffNet
        from ..modules.diffusion import GaussianDiffusion
        from ..modules.esp import ConformerEncoder
        from ..modules.frame_prior import FramePriorNetwork
        from ..modules.mdn import MDNLayer
        from ..modules.prompt_encoder import PromptEncoder
        from ..modules.style_encoder import StyleEncoder
        from ..modules.variance_adaptor import MDNPredictor, Predictor, VarianceAdaptor

        ns = dict(PhonemeEmbedding=PhonemeEmbedding, PromptTTSMDNDurCFG=PromptTTSMDNDurCFG, DiffNet=DiffNet,
                  GaussianDiffusion=GaussianDiffusion, ConformerEncoder=ConformerEncoder,
                  FramePriorNetwork=FramePriorNetwork, MDNLayer=MDNLayer, PromptEncoder=PromptEncoder,
                  StyleEncoder=StyleEncoder, MDNPredictor=MDNPredictor, Predictor=Predictor,
                  VarianceAdaptor=VarianceAdaptor)
    C = 256
    pe_kwargs = dict(model_name="bert-base-uncased", in_channels=768, mid_channels=512, out_channels=C)
    if bert is not None:
        pe_kwargs["bert"] = bert
    return ns["PromptTTSMDNDurCFG"](
        phoneme_embedding=ns["PhonemeEmbedding"](num_vocab=90, channels=C, do_scale=False, init_normal=False),
        encoder=ns["ConformerEncoder"](
            idim=C, attention_dim=C, attention_heads=2, linear_units=1024, num_blocks=4,
            positionwise_layer_type="conv1d", positionwise_conv_kernel_size=9, dropout_rate=0.2,
            pos_enc_layer_type="rel_pos", selfattention_layer_type="rel_selfattn", activation_type="swish",
            macaron_style=True, use_cnn_module=True, cnn_module_kernel=7, return_mask=False,
            rel_pos_type=rel_pos_type),
        variance_adaptor=ns["VarianceAdaptor"](
            duration_predictor=ns["MDNPredictor"](channels=C, out_channels=1, kernel_size=3, dropout=0.5,
                                                  num_layers=2, num_gaussians=4, detach=True, disable_amp=True),
            pitch_predictor=ns["Predictor"](channels=C, out_channels=2, kernel_size=5, dropout=0.5, num_layers=5,
                                            detach=False),
            pitch_emb=torch.nn.Conv1d(1, C, 1),
            energy_predictor=None, energy_emb=None,
            frame_prior_network=ns["FramePriorNetwork"](out_channels=C, hidden_channels=C, n_layers=6,
                                                        kernel_size=17, p_dropout=0.1)),
        reference_encoder=ns["StyleEncoder"](idim=80, gst_tokens=10, gst_heads=4, conv_layers=6,
                                             conv_chans_list=[128, 128, 256, 256, 512, 512], conv_kernel_size=3,
                                             conv_stride=2, gru_layers=1, gru_units=C),
        prompt_encoder=ns["PromptEncoder"](**pe_kwargs),
        decoder=ns["GaussianDiffusion"](
            in_dim=C, out_dim=80, norm_scale=6.0, K_step=K_step,
            denoise_fn=ns["DiffNet"](in_dim=80, encoder_hidden_dim=C, residual_layers=20, residual_channels=256,
                                     kernel_size=3, dilation_cycle_length=4)),
        style_mdn=ns["MDNLayer"](in_dim=C, out_dim=C, num_gaussians=10, dim_wise=True),
        norm_style_emb=True, mdn_disable_amp=True,
    )


_KEEP = ("filter", "num_batches_tracked", "decoder.betas", "decoder.alphas_cumprod", "decoder.sqrt_",
         "decoder.log_one_minus", "decoder.posterior_")


def synthetic_state_dict(module, seed=1234, frames_per_phoneme=8.0):
    """Deterministic re-draw of every learnable tensor of `module` (CPU generator, key order)."""
    gen = torch.Generator().manual_seed(seed)
    sd = module.state_dict()
    out = {}

    def normal(shape, std):
        return torch.randn(shape, generator=gen) * std

    def uniform(shape, lo, hi):
        return torch.rand(shape, generator=gen) * (hi - lo) + lo

    for k, v in sd.items():
        if k.startswith("prompt_encoder.bert."):
            continue  # external BERT weights are never part of the synthetic checkpoint
        if any(tag in k for tag in _KEEP) or not v.dtype.is_floating_point:
            out[k] = v.clone()
            continue
        shape = tuple(v.shape)
        leaf = k.rsplit(".", 1)[-1]
        if leaf == "alpha":  # Snake log-alpha
            t = uniform(shape, -0.5, 0.5)
        elif leaf == "running_mean":
            t = normal(shape, 0.1)
        elif leaf == "running_var":
            t = uniform(shape, 0.5, 1.5)
        elif leaf in ("gamma",) or (leaf == "weight" and v.dim() == 1):  # LayerNorm / BatchNorm scale
            t = 1.0 + normal(shape, 0.1)
        elif leaf in ("beta", "bias") or "bias" in leaf:
            t = normal(shape, 0.02)
        elif leaf in ("pos_bias_u", "pos_bias_v"):
            t = normal(shape, 0.05)
        elif leaf == "weight_g":
            t = None  # filled in after weight_v
        elif leaf == "gst_embs":
            t = normal(shape, 0.5)
        elif k.endswith("emb.weight") and v.dim() == 2:  # phoneme embedding table
            t = normal(shape, 1.0)
            t[0].zero_()  # padding_idx row
        elif v.dim() >= 2:
            # fan-in scaled Gaussian; ConvTranspose1d weight_v is [Cin][Cout][k] -> fan_in = Cin*k/stride-ish
            if "upsamples" in k:
                fan_in = shape[0] * 2
            else:
                fan_in = math.prod(shape[1:])
            t = normal(shape, 1.0 / math.sqrt(fan_in))
        else:
            t = normal(shape, 0.1)
        out[k] = t
    # weight-norm gains: g = ||v|| * U(0.8, 1.2) over all dims but 0
    for k in list(out):
        if k.endswith("weight_g"):
            v = out[k[:-1] + "v"]
            norm = v.flatten(1).norm(dim=1).view(sd[k].shape)
            out[k] = norm * uniform(tuple(sd[k].shape), 0.8, 1.2)
    # BigVGAN: the residual stages grow the activations to a standard deviation of ~30 in front of conv_post; with a
    # unit-gain head the final tanh saturates (|wav| > 0.99 on 92 % of the samples) and squashes every upstream error,
    # which would make the waveform parity bar meaningless.  A head gain of 0.012 gives a speech-like output
    # (pre-tanh standard deviation ~0.3), so errors anywhere in the stack reach the waveform linearly.
    if "conv_post.weight_g" in out:
        out["conv_post.weight_g"] = out["conv_post.weight_g"] * 0.012
    # observable denoiser head (zero-initialised upstream) and a speech-like duration head
    k = "decoder.denoise_fn.output_projection.weight"
    if k in out:
        out[k] = out[k] * 0.5
    pre = "variance_adaptor.duration_predictor.out_layer."
    if pre + "mu.weight" in out:
        out[pre + "mu.weight"] = out[pre + "mu.weight"] * 0.02 * math.sqrt(out[pre + "mu.weight"].shape[1])
        out[pre + "mu.bias"] = torch.full_like(out[pre + "mu.bias"], math.log(frames_per_phoneme)) + normal(
            tuple(out[pre + "mu.bias"].shape), 0.05)
        out[pre + "log_sigma.weight"] = out[pre + "log_sigma.weight"] * 0.2
        out[pre + "log_sigma.bias"] = torch.full_like(out[pre + "log_sigma.bias"], -1.5)
    return out

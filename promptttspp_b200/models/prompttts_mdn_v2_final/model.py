"""PromptTTSMDNDurCFG -- drop-in for promptttspp.models.prompttts_mdn_v2_final.model.

Same constructor (children passed in as modules, conf/model/prompttts_mdn_v2_wo_erg_final*.yaml),
same state_dict keys, same `infer` / `infer_batch` / `sample_style_emb`-level behaviour
(reference: promptttspp/models/prompttts_mdn_v2_final/model.py:28-325).  Inference runs in two
native calls (csrc/acoustic.cu): `encode` (text side, up to integer durations) and `decode`
(length regulator ... diffusion sampler); between them the host reads the frame lengths once to
size the output tensors -- the only device->host synchronisation of a call (the reference has
three: utils/model.py:32, variance_adaptor.py:154, reference_encoder.py:119).
"""
import ctypes as C
import math
from typing import NamedTuple, Optional

import torch
from torch import nn

from ... import _abi
from ..._engine import NativeHandle
from ...modules.denoiser import DiffNet
from ...modules.diffusion import GaussianDiffusion
from ...modules.esp import ConformerEncoder


class InferNoise(NamedTuple):
    """Pre-drawn Gaussian noise, in the order the reference consumes torch's generator.

    z_style: [B, 1, C]           torch.randn_like(sigma)          (model.py:191)
    x_T:     [B, mel, Ty]        torch.randn(shape)               (diffusion.py:332)
    z:       [K_step, B, mel, Ty] noise_like(x.shape) per step    (diffusion.py:218), z[0] first
    """

    z_style: torch.Tensor
    x_T: Optional[torch.Tensor]
    z: Optional[torch.Tensor]
    comp_u: Optional[torch.Tensor] = None  # [B, C] uniforms: the per-dimension component draw of use_max=False


def _sinusoid_table(positions: torch.Tensor, d_model: int) -> torch.Tensor:
    """sin/cos(position * div_term), the fp32 recipe of esp/transformer/embedding.py:68-77."""
    div_term = torch.exp(torch.arange(0, d_model, 2, dtype=torch.float32) * -(math.log(10000.0) / d_model))
    pe = torch.zeros(positions.numel(), d_model)
    ang = positions.to(torch.float32).unsqueeze(1) * div_term
    pe[:, 0::2] = torch.sin(ang)
    pe[:, 1::2] = torch.cos(ang)
    return pe


class PromptTTSMDNDurCFG(nn.Module):
    def __init__(self, phoneme_embedding, encoder, variance_adaptor, reference_encoder, prompt_encoder, decoder,
                 out_conv=None, style_mdn=None, norm_style_emb=False, mdn_disable_amp=False, loss_dec_scale=8.0):
        super().__init__()
        self.phoneme_emb = phoneme_embedding
        self.encoder = encoder
        self.variance_adaptor = variance_adaptor
        self.reference_encoder = reference_encoder
        self.prompt_encoder = prompt_encoder
        self.style_mdn = style_mdn
        self.decoder = decoder
        self.out_conv = out_conv
        self.norm_style_emb = norm_style_emb
        self.mdn_disable_amp = mdn_disable_amp
        self.loss_dec_scale = loss_dec_scale
        assert self.variance_adaptor.frame_prior_network is not None  # model.py:70
        if not isinstance(encoder, ConformerEncoder):
            raise NotImplementedError("only the ConformerEncoder text encoder is supported")
        if not isinstance(decoder, GaussianDiffusion) or not isinstance(decoder.denoise_fn, DiffNet):
            raise NotImplementedError("only the GaussianDiffusion(DiffNet) decoder is supported")
        if style_mdn is None or not style_mdn.dim_wise:
            raise NotImplementedError("style_mdn with dim_wise=True is required (shipped configs)")
        self._native = None
        self._tensor_view = None
        self._pe_cache = {}

    # ---- training forward: out of the accelerated path ------------------------------------
    def forward(self, batch):
        raise NotImplementedError(
            "promptttspp_b200 accelerates inference (infer / infer_batch); the training loss "
            "(reference model.py:72-183) is out of scope"
        )

    # ---- native handle ---------------------------------------------------------------------
    def _config(self):
        va = self.variance_adaptor
        dn = self.decoder.denoise_fn
        cfg = _abi.AcousticConfig()
        cfg.num_vocab = self.phoneme_emb.num_vocab
        cfg.channels = self.phoneme_emb.channels
        cfg.emb_do_scale = int(self.phoneme_emb.do_scale)
        cfg.enc_heads = self.encoder.heads
        cfg.enc_linear_units = self.encoder.linear_units
        cfg.enc_blocks = self.encoder.num_blocks
        cfg.enc_ff_kernel = self.encoder.ff_kernel
        cfg.enc_cnn_kernel = self.encoder.cnn_kernel
        cfg.rel_pos_legacy = int(self.encoder.rel_pos_type == "legacy")
        cfg.dur_layers = len(va.duration_predictor.layers)
        cfg.dur_kernel = va.duration_predictor.kernel_size
        cfg.dur_gaussians = va.duration_predictor.out_layer.num_gaussians
        cfg.pitch_layers = len(va.pitch_predictor.layers)
        cfg.pitch_kernel = va.pitch_predictor.kernel_size
        cfg.fp_layers = va.frame_prior_network.n_layers
        cfg.fp_kernel = va.frame_prior_network.kernel_size
        ad = self.prompt_encoder.adaptor
        cfg.prompt_in = ad[0].in_features
        cfg.prompt_mid = ad[0].out_features
        cfg.style_gaussians = self.style_mdn.num_gaussians
        cfg.norm_style_emb = int(self.norm_style_emb)
        cfg.mel_dim = self.decoder.out_dim
        cfg.K_step = self.decoder.K_step
        cfg.diff_layers = len(dn.residual_layers)
        cfg.diff_channels = dn.residual_channels
        cfg.diff_kernel = dn.kernel_size
        cfg.diff_dilation_cycle = dn.dilation_cycle_length
        cfg.diff_scale = float(dn.scale)
        cfg.norm_scale = float(self.decoder.norm_scale) if self.decoder.norm_scale is not None else 0.0
        cfg.a_min = float(self.decoder.a_min)
        cfg.a_max = float(self.decoder.a_max)
        if va.duration_predictor.out_layer.out_dim != 1 or va.pitch_predictor.out_layer.out_channels != 2:
            raise NotImplementedError("duration head must be 1-D, pitch head 2-D (log_cf0, vuv)")
        return cfg

    def _handle(self, device):
        if self._native is None:
            self._native = NativeHandle("acoustic", self._config())
        # the {key: tensor} view of the module tree is rebuilt only when the tree changed (load_state_dict keeps the
        # Parameter objects, .to() / remove_weight_norm_ replace them): cheap identity probe instead of a state_dict()
        # per call on the B=1 latency path
        probe = tuple(id(p) for p in self.parameters(recurse=True)) + tuple(id(b) for b in self.buffers(recurse=True))
        if self._tensor_view is None or self._tensor_view[0] != probe:
            tensors = {
                k: v for k, v in self.state_dict(keep_vars=True).items()
                if not k.startswith("prompt_encoder.bert.") and not k.startswith("reference_encoder.")
                and v.dtype.is_floating_point
            }
            self._tensor_view = (probe, tensors)
        self._native.sync(self._tensor_view[1], device)
        return self._native

    def refresh_weights(self):
        """Force a re-upload of the packed native weights on the next call.  Needed only after writes that bypass
        autograd's version counter (`param.data.copy_(...)`, e.g. an EMA swap): load_state_dict, `.to()`, in-place ops
        on the parameters and remove_weight_norm_ are detected automatically."""
        self._tensor_view = None
        if self._native is not None:
            self._native.invalidate()

    def _pos_table(self, kind, T, device):
        """Positional tables built on the host in fp32 exactly like the reference, cached per length."""
        key = (kind, T, str(device))
        tab = self._pe_cache.get(key)
        if tab is None:
            C_ = self.phoneme_emb.channels
            if kind == "legacy":
                # LegacyRelPositionalEncoding: a reversed table of max_len 5000 sliced from the front,
                # i.e. pe[k] = sinusoid(max(5000, T) - 1 - k)  (esp/transformer/embedding.py:58-79, :256)
                n = max(5000, T)
                tab = _sinusoid_table(torch.arange(n - 1, -1, -1.0)[:T], C_)
            elif kind == "new":
                # RelPositionalEncoding: row k <-> relative position T-1-k, k in [0, 2T-2] (:283-331)
                pos = _sinusoid_table(torch.arange(0, T, dtype=torch.float32), C_)
                neg = _sinusoid_table(-1 * torch.arange(0, T, dtype=torch.float32), C_)
                tab = torch.cat([torch.flip(pos, [0]), neg[1:]], dim=0)
            else:  # absolute table of the frame-prior network (modules/embedding.py:57-78)
                tab = _sinusoid_table(torch.arange(0, T, dtype=torch.float32), C_)
            tab = tab.contiguous().to(device)
            if len(self._pe_cache) > 64:
                self._pe_cache.clear()
            self._pe_cache[key] = tab
        return tab

    # ---- inference ---------------------------------------------------------------------------
    @torch.no_grad()
    def _synthesize(self, phoneme, phone_lengths, style_prompt, reference_mel, use_max, noise_scale, noise,
                    ref_lengths=None):
        assert (style_prompt is not None) ^ (reference_mel is not None), "One of style inputs must not be None."
        _abi.require_cuda(phoneme, "PromptTTSMDNDurCFG.infer")
        device = phoneme.device
        B, Tx = phoneme.shape
        Cc = self.phoneme_emb.channels
        M = self.decoder.out_dim
        K = self.decoder.K_step
        phoneme = phoneme.to(torch.int64).contiguous()
        phone_lengths = phone_lengths.to(device=device, dtype=torch.int64).contiguous()
        with torch.cuda.device(device):
            nat = self._handle(device)
            lib = _abi.lib()
            stream = _abi.stream_ptr(device)
            style_in = None
            if reference_mel is not None:
                # reference-mel style path (model.py:232-235 / :297-301): StyleEncoder on the normalised mel
                if ref_lengths is None:
                    ref_lengths = torch.full((B,), reference_mel.shape[-1], dtype=torch.int64, device=device)
                style_in = self.reference_encoder(reference_mel.to(device), ref_lengths)[:, :, 0].float().contiguous()
                if style_in.shape[0] != B:
                    raise ValueError(f"{style_in.shape[0]} reference mels for a batch of {B}")
            else:
                cls = self.prompt_encoder.sentence_embedding(style_prompt, device).float().contiguous()
                if cls.shape[0] != B:
                    raise ValueError(f"{cls.shape[0]} style prompts for a batch of {B}")
                # RNG draw #1 (model.py:191)
                z_style = noise.z_style if noise is not None else torch.randn(B, 1, Cc, device=device)
                z_style = z_style.to(device=device, dtype=torch.float32).reshape(B, Cc).contiguous()
                comp_u = None
                if not use_max:
                    # mdn_sample_sigma_and_mu (mdn.py:226-257): one component draw per (utterance, dimension), supplied
                    # as uniforms (the reference's Categorical.sample consumes its own generator stream)
                    comp_u = getattr(noise, "comp_u", None) if noise is not None else None
                    if comp_u is None:
                        comp_u = torch.rand(B, Cc, device=device)
                    comp_u = comp_u.to(device=device, dtype=torch.float32).reshape(B, Cc).contiguous()
            legacy = self.encoder.rel_pos_type == "legacy"
            pos = self._pos_table("legacy" if legacy else "new", Tx, device)
            enc_state = torch.empty(B, Tx, Cc, device=device)
            dur = torch.empty(B, Tx, dtype=torch.int64, device=device)
            frame_len = torch.empty(B, dtype=torch.int64, device=device)
            log_dur = torch.empty(B, Tx, device=device)
            ws = nat.workspace(lib.pttspp_acoustic_encode_workspace_bytes(nat.h, B, Tx), device)
            if style_in is not None:
                _abi.check(lib.pttspp_acoustic_encode_ref(
                    nat.h, _abi.ptr(phoneme), _abi.ptr(phone_lengths), B, Tx, _abi.ptr(pos), pos.shape[0],
                    _abi.ptr(style_in), _abi.ptr(enc_state), _abi.ptr(dur), _abi.ptr(frame_len), _abi.ptr(log_dur),
                    _abi.ptr(ws), C.c_size_t(ws.numel()), stream))
            elif not use_max:
                _abi.check(lib.pttspp_acoustic_encode_sampled(
                    nat.h, _abi.ptr(phoneme), _abi.ptr(phone_lengths), B, Tx, _abi.ptr(pos), pos.shape[0],
                    _abi.ptr(cls), _abi.ptr(z_style), _abi.ptr(comp_u), float(noise_scale), _abi.ptr(enc_state),
                    _abi.ptr(dur), _abi.ptr(frame_len), _abi.ptr(log_dur), None, _abi.ptr(ws),
                    C.c_size_t(ws.numel()), stream))
            else:
                _abi.check(lib.pttspp_acoustic_encode(
                    nat.h, _abi.ptr(phoneme), _abi.ptr(phone_lengths), B, Tx, _abi.ptr(pos), pos.shape[0],
                    _abi.ptr(cls), _abi.ptr(z_style), float(noise_scale), int(use_max), _abi.ptr(enc_state),
                    _abi.ptr(dur), _abi.ptr(frame_len), _abi.ptr(log_dur), None, _abi.ptr(ws),
                    C.c_size_t(ws.numel()), stream))
            Ty = int(frame_len.max().item())  # the one device->host sync: sizes the outputs
            pe_abs = self._pos_table("abs", max(Ty, 1), device)
            # RNG draws #2 .. #K+2, same shapes and order as diffusion.py:332 and :218
            if noise is not None and noise.x_T is not None:
                x_T = noise.x_T.to(device=device, dtype=torch.float32).contiguous()
                z = noise.z.to(device=device, dtype=torch.float32).contiguous()
                if tuple(x_T.shape) != (B, M, Ty) or tuple(z.shape) != (K, B, M, Ty):
                    raise ValueError(f"injected noise has shapes {tuple(x_T.shape)}, {tuple(z.shape)}; "
                                     f"expected {(B, M, Ty)}, {(K, B, M, Ty)}")
            else:
                x_T = torch.randn((B, M, Ty), device=device)
                z = None  # drawn step by step inside the native loop, from this generator's Philox stream
            mel = torch.empty(B, M, Ty, device=device)
            log_cf0 = torch.empty(B, 1, Ty, device=device)
            vuv = torch.empty(B, 1, Ty, device=device)
            if Ty > 0:
                ws = nat.workspace(lib.pttspp_acoustic_decode_workspace_bytes(nat.h, B, Tx, Ty), device)
                if z is not None:
                    _abi.check(lib.pttspp_acoustic_decode(
                        nat.h, _abi.ptr(enc_state), _abi.ptr(dur), _abi.ptr(frame_len), B, Tx, Ty, _abi.ptr(pe_abs),
                        _abi.ptr(x_T), _abi.ptr(z), _abi.ptr(mel), _abi.ptr(log_cf0), _abi.ptr(vuv), None,
                        _abi.ptr(ws), C.c_size_t(ws.numel()), stream))
                else:
                    # the K per-step draws of diffusion.py:218, bit-identical to K x torch.randn((B, M, Ty)) from this
                    # generator state; afterwards the generator is moved past them
                    gen = torch.cuda.default_generators[device.index if device.index is not None
                                                        else torch.cuda.current_device()]
                    off_out = C.c_uint64(0)
                    _abi.check(lib.pttspp_acoustic_decode_rng(
                        nat.h, _abi.ptr(enc_state), _abi.ptr(dur), _abi.ptr(frame_len), B, Tx, Ty, _abi.ptr(pe_abs),
                        _abi.ptr(x_T), C.c_uint64(gen.initial_seed()), C.c_uint64(gen.get_offset()), C.byref(off_out),
                        _abi.ptr(mel), _abi.ptr(log_cf0), _abi.ptr(vuv), None, _abi.ptr(ws), C.c_size_t(ws.numel()),
                        stream))
                    gen.set_offset(off_out.value)
        self.last_durations = dur
        self.last_log_durations = log_dur
        # the reference returns frame_mask.sum(dim=(1, 2)): a float tensor (model.py:311)
        return mel, log_cf0, vuv, frame_len.to(torch.float32)

    @torch.no_grad()
    def generate_style_emb(self, style_prompt, reference_mel, use_max=True, noise_scale=1.0, *,
                           noise: Optional[InferNoise] = None):
        """(prompt_emb, ref_emb), both [B, C, 1] -- model.py:327-344: the style vector the prompt path would add to the
        encoder output (adaptor -> normalise -> style MDN -> sample -> normalise) and the reference-mel style encoder's,
        normalised.  The prompt side runs the text-side native call on a one-phoneme dummy (its `style_emb` output)."""
        _abi.require_cuda(reference_mel, "PromptTTSMDNDurCFG.generate_style_emb")
        device = reference_mel.device
        Cc = self.phoneme_emb.channels
        with torch.cuda.device(device):
            nat = self._handle(device)
            lib = _abi.lib()
            stream = _abi.stream_ptr(device)
            cls = self.prompt_encoder.sentence_embedding(style_prompt, device).float().contiguous()
            B = cls.shape[0]
            z_style = noise.z_style if noise is not None else torch.randn(B, 1, Cc, device=device)
            z_style = z_style.to(device=device, dtype=torch.float32).reshape(B, Cc).contiguous()
            phoneme = torch.ones(B, 1, dtype=torch.int64, device=device)
            lengths = torch.ones(B, dtype=torch.int64, device=device)
            legacy = self.encoder.rel_pos_type == "legacy"
            pos = self._pos_table("legacy" if legacy else "new", 1, device)
            enc_state = torch.empty(B, 1, Cc, device=device)
            dur = torch.empty(B, 1, dtype=torch.int64, device=device)
            frame_len = torch.empty(B, dtype=torch.int64, device=device)
            style = torch.empty(B, Cc, device=device)
            ws = nat.workspace(lib.pttspp_acoustic_encode_workspace_bytes(nat.h, B, 1), device)
            if use_max:
                _abi.check(lib.pttspp_acoustic_encode(
                    nat.h, _abi.ptr(phoneme), _abi.ptr(lengths), B, 1, _abi.ptr(pos), pos.shape[0], _abi.ptr(cls),
                    _abi.ptr(z_style), float(noise_scale), 1, _abi.ptr(enc_state), _abi.ptr(dur), _abi.ptr(frame_len),
                    None, _abi.ptr(style), _abi.ptr(ws), C.c_size_t(ws.numel()), stream))
            else:
                comp_u = getattr(noise, "comp_u", None) if noise is not None else None
                if comp_u is None:
                    comp_u = torch.rand(B, Cc, device=device)
                comp_u = comp_u.to(device=device, dtype=torch.float32).reshape(B, Cc).contiguous()
                _abi.check(lib.pttspp_acoustic_encode_sampled(
                    nat.h, _abi.ptr(phoneme), _abi.ptr(lengths), B, 1, _abi.ptr(pos), pos.shape[0], _abi.ptr(cls),
                    _abi.ptr(z_style), _abi.ptr(comp_u), float(noise_scale), _abi.ptr(enc_state), _abi.ptr(dur),
                    _abi.ptr(frame_len), None, _abi.ptr(style), _abi.ptr(ws), C.c_size_t(ws.numel()), stream))
            prompt_emb = style.unsqueeze(-1)
            if self.norm_style_emb:  # the second normalisation of model.py:338-339 (a no-op up to rounding)
                prompt_emb = torch.nn.functional.normalize(prompt_emb, dim=1)
            ref_lengths = torch.full((reference_mel.shape[0],), reference_mel.shape[-1], dtype=torch.int64, device=device)
            ref_emb = self.reference_encoder(reference_mel, ref_lengths).float()
            if self.norm_style_emb:
                ref_emb = torch.nn.functional.normalize(ref_emb, dim=1)
        return prompt_emb, ref_emb

    def infer(self, x, style_prompt=None, reference_mel=None, use_max=True, noise_scale=1.0, return_f0=False,
              *, noise: Optional[InferNoise] = None):
        """x: LongTensor [1, L] -> mel [1, 80, Ty] (and log_cf0, vuv [1, 1, Ty])  (model.py:198-259)."""
        lengths = torch.full((x.shape[0],), x.shape[-1], dtype=torch.int64, device=x.device)
        mel, log_cf0, vuv, _ = self._synthesize(x, lengths, style_prompt, reference_mel, use_max, noise_scale, noise)
        if return_f0:
            return mel, log_cf0, vuv
        return mel

    def infer_batch(self, phoneme, phone_lengths, style_prompt=None, reference_mel=None, ref_lengths=None,
                    use_max=True, noise_scale=1.0, return_f0=False, *, noise: Optional[InferNoise] = None):
        """Batched inference (model.py:261-325): returns (mel, [log_cf0, vuv,] frame_lengths)."""
        if reference_mel is not None:
            assert ref_lengths is not None  # model.py:296
        mel, log_cf0, vuv, frame_lengths = self._synthesize(
            phoneme, phone_lengths, style_prompt, reference_mel, use_max, noise_scale, noise, ref_lengths)
        if return_f0:
            return mel, log_cf0, vuv, frame_lengths
        return mel, frame_lengths

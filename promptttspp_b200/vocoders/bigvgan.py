"""BigVGAN generator -- drop-in for promptttspp.vocoders.BigVGAN.

Same constructor kwargs (conf/vocoder/bigvgan.yaml), same state_dict keys (weight-normed
convs, `act*.act.alpha`, `act*.up.filter`, `act*.down.lowpass.filter`), same
`forward(mel[B, 80, T]) -> wav[B, 1, 240*T]` (reference: promptttspp/vocoders/bigvgan.py:71-131).
The sub-modules below hold parameters only; the arithmetic is pttspp_bigvgan_forward
(csrc/bigvgan.cu).
"""
import ctypes as C

import torch
from torch import nn
from torch.nn.utils import weight_norm

from .. import _abi
from .._engine import NativeHandle
from ..layers.activations import AntiAliasActivation


class AMPLayer(nn.Module):
    def __init__(self, channels, kernel_size, dilation):
        super().__init__()
        self.conv1 = weight_norm(nn.Conv1d(channels, channels, kernel_size, dilation=dilation,
                                           padding=(kernel_size * dilation - dilation) // 2))
        self.conv2 = weight_norm(nn.Conv1d(channels, channels, kernel_size, padding=kernel_size // 2))
        self.act1 = AntiAliasActivation(channels)
        self.act2 = AntiAliasActivation(channels)


class AMPBlock(nn.Module):
    def __init__(self, channels, kernel_size, dilations):
        super().__init__()
        self.layers = nn.ModuleList([AMPLayer(channels, kernel_size, d) for d in dilations])


class BigVGAN(nn.Module):
    def __init__(self, in_channel, upsample_initial_channel, upsample_rates, upsample_kernel_sizes,
                 resblock_kernel_sizes, resblock_dilations):
        super().__init__()
        self.in_channel = in_channel
        self.upsample_initial_channel = upsample_initial_channel
        self.upsample_rates = list(upsample_rates)
        self.upsample_kernel_sizes = list(upsample_kernel_sizes)
        self.resblock_kernel_sizes = list(resblock_kernel_sizes)
        self.resblock_dilations = [list(d) for d in resblock_dilations]
        self.num_kernels = len(self.resblock_kernel_sizes)
        if len({len(d) for d in self.resblock_dilations}) != 1:
            raise NotImplementedError("all AMP blocks must have the same number of layers")

        ch = upsample_initial_channel
        self.conv_pre = weight_norm(nn.Conv1d(in_channel, ch, kernel_size=7, stride=1, padding=3))
        self.upsamples = nn.ModuleList()
        self.mrfs = nn.ModuleList()
        for i, (u, k) in enumerate(zip(self.upsample_rates, self.upsample_kernel_sizes)):
            self.upsamples.append(weight_norm(nn.ConvTranspose1d(
                ch // (2 ** i), ch // (2 ** (i + 1)), kernel_size=k, stride=u,
                padding=u // 2 + u % 2, output_padding=u % 2)))
            self.mrfs.append(nn.ModuleList([
                AMPBlock(ch // (2 ** (i + 1)), kernel_size=rk, dilations=rd)
                for rk, rd in zip(self.resblock_kernel_sizes, self.resblock_dilations)]))
        last = ch // (2 ** len(self.upsample_rates))
        self.act_post = AntiAliasActivation(last)
        self.conv_post = weight_norm(nn.Conv1d(last, 1, kernel_size=7, stride=1, padding=3))
        self._native = None
        self._tensor_view = None

    # -- native handle -------------------------------------------------------------------
    def _config(self):
        cfg = _abi.BigVGANConfig()
        cfg.in_channel = self.in_channel
        cfg.upsample_initial_channel = self.upsample_initial_channel
        cfg.num_upsamples = len(self.upsample_rates)
        for i, (u, k) in enumerate(zip(self.upsample_rates, self.upsample_kernel_sizes)):
            cfg.upsample_rates[i] = u
            cfg.upsample_kernel_sizes[i] = k
        cfg.num_kernels = self.num_kernels
        cfg.num_dilations = len(self.resblock_dilations[0])
        for j, (rk, rd) in enumerate(zip(self.resblock_kernel_sizes, self.resblock_dilations)):
            cfg.resblock_kernel_sizes[j] = rk
            for l, d in enumerate(rd):
                cfg.resblock_dilations[j][l] = d
        return cfg

    def _handle(self, device):
        if self._native is None:
            self._native = NativeHandle("bigvgan", self._config())
        probe = tuple(id(p) for p in self.parameters(recurse=True)) + tuple(id(b) for b in self.buffers(recurse=True))
        if self._tensor_view is None or self._tensor_view[0] != probe:
            self._tensor_view = (probe, dict(self.state_dict(keep_vars=True)))
        self._native.sync(self._tensor_view[1], device)
        return self._native

    def refresh_weights(self):
        """Force a re-upload of the packed native weights (after writes through `.data`, see NativeHandle.invalidate)."""
        self._tensor_view = None
        if self._native is not None:
            self._native.invalidate()

    @property
    def hop(self):
        r = 1
        for u in self.upsample_rates:
            r *= u
        return r

    @torch.no_grad()
    def forward(self, x):
        """x: mel [B, in_channel, T] (float32, CUDA) -> waveform [B, 1, T * prod(upsample_rates)]."""
        _abi.require_cuda(x, "BigVGAN.forward")
        if x.dim() != 3 or x.shape[1] != self.in_channel:
            raise ValueError(f"expected mel of shape [B, {self.in_channel}, T], got {tuple(x.shape)}")
        x = x.contiguous().float()
        B, _, T = x.shape
        out = torch.empty(B, 1, T * self.hop, dtype=torch.float32, device=x.device)
        if B == 0 or T == 0:
            return out
        with torch.cuda.device(x.device):
            nat = self._handle(x.device)
            lib = _abi.lib()
            nbytes = lib.pttspp_bigvgan_workspace_bytes(nat.h, B, T)
            ws = nat.workspace(nbytes, x.device)
            _abi.check(lib.pttspp_bigvgan_forward(nat.h, _abi.ptr(x), B, T, _abi.ptr(out), _abi.ptr(ws),
                                                  C.c_size_t(ws.numel()), _abi.stream_ptr(x.device)))
        return out

    def remove_weight_norm(self):
        from ..utils.model import remove_weight_norm_

        self.apply(remove_weight_norm_)
